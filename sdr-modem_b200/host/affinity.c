/*
 * Host thread placement next to a GPU.
 *
 * The host-buffer entry points move 8 bytes per sample over PCIe; on a two-socket box a staging buffer that was pinned on the
 * far socket sends every copy over the inter-socket link, which is what bounds eight GPUs fed at once (measured: 8 x B200,
 * 22.5 Gsamples/s in total against 6.6 per GPU alone). Pinned memory is placed where the allocating thread runs (first
 * touch), so a thread that is about to allocate staging buffers for device `d`, or to feed it, first moves itself onto the
 * CPUs the kernel lists as local to that device.
 *
 * The reference has no equivalent: its dsp_worker threads are unpinned (src/dsp_worker.c:199-227) and all of its
 * arithmetic runs on the CPU.
 */
#include <ctype.h>
#include <errno.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "sdrm/sdrm_batch.h"

/* "0-23,48-71" -> cpu set; returns the number of CPUs parsed, or -1 for a malformed list */
static int parse_cpulist(const char *text, cpu_set_t *set) {
    int count = 0;
    CPU_ZERO(set);
    const char *p = text;
    while (*p != '\0') {
        while (*p == ',' || isspace((unsigned char) *p)) {
            p++;
        }
        if (*p == '\0') {
            break;
        }
        char *end = NULL;
        long first = strtol(p, &end, 10);
        if (end == p || first < 0) {
            return -1;
        }
        long last = first;
        p = end;
        if (*p == '-') {
            p++;
            last = strtol(p, &end, 10);
            if (end == p || last < first) {
                return -1;
            }
            p = end;
        }
        for (long c = first; c <= last; c++) {
            if (c < CPU_SETSIZE) {
                CPU_SET((int) c, set);
                count++;
            }
        }
    }
    return count;
}

int sdrm_cpulist_parse_count(const char *text) {
    cpu_set_t set;
    if (text == NULL) {
        return -1;
    }
    return parse_cpulist(text, &set);
}

int sdrm_bind_thread_near_device(int device) {
    char bus_id[32];
    if (cudaDeviceGetPCIBusId(bus_id, (int) sizeof(bus_id), device) != cudaSuccess) {
        cudaGetLastError();
        return -ENODEV;
    }
    for (char *c = bus_id; *c != '\0'; c++) {
        *c = (char) tolower((unsigned char) *c);
    }
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus_id);
    FILE *f = fopen(path, "r");
    if (f == NULL) {
        return 0; /* no topology information (container without sysfs): leave the thread where it is */
    }
    char text[1024];
    const size_t got = fread(text, 1, sizeof(text) - 1, f);
    fclose(f);
    text[got] = '\0';
    cpu_set_t local;
    if (parse_cpulist(text, &local) <= 0) {
        return 0;
    }
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) {
        return -errno;
    }
    cpu_set_t both;
    CPU_AND(&both, &local, &allowed);
    const int n = CPU_COUNT(&both);
    if (n == 0) {
        return 0; /* the cgroup keeps this process off the device's socket */
    }
    if (sched_setaffinity(0, sizeof(both), &both) != 0) {
        return -errno;
    }
    return n;
}
