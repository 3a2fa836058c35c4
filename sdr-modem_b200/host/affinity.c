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
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

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

/* reads /sys/bus/pci/devices/<bus id of `device`>/<leaf> into text; 0 on success */
static int read_pci_attribute(int device, const char *leaf, char *text, size_t cap) {
    char bus_id[32];
    if (cudaDeviceGetPCIBusId(bus_id, (int) sizeof(bus_id), device) != cudaSuccess) {
        cudaGetLastError();
        return -ENODEV;
    }
    for (char *c = bus_id; *c != '\0'; c++) {
        *c = (char) tolower((unsigned char) *c);
    }
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/%s", bus_id, leaf);
    FILE *f = fopen(path, "r");
    if (f == NULL) {
        return -ENOENT; /* no topology information (container without sysfs) */
    }
    const size_t got = fread(text, 1, cap - 1, f);
    fclose(f);
    text[got] = '\0';
    return 0;
}

int sdrm_device_numa_node(int device) {
    char text[64];
    if (read_pci_attribute(device, "numa_node", text, sizeof(text)) != 0) {
        return -1;
    }
    char *end = NULL;
    const long node = strtol(text, &end, 10);
    return end == text || node < 0 ? -1 : (int) node;
}

int sdrm_bind_thread_near_device(int device) {
    char text[1024];
    const int read_code = read_pci_attribute(device, "local_cpulist", text, sizeof(text));
    if (read_code == -ENODEV) {
        return -ENODEV;
    }
    if (read_code != 0) {
        return 0; /* leave the thread where it is */
    }
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

/*
 * Pinned staging memory with an explicit NUMA policy.
 *
 * cudaHostAlloc places pages by first touch of the calling thread, which sdrm_bind_thread_near_device can only steer when
 * the kernel exposes the device's local CPUs. This allocator asks for the node itself: anonymous mapping, 2 MB aligned,
 * MPOL_BIND to the device's NUMA node (mbind through the raw system call: libnuma is not a dependency), transparent huge
 * pages requested, every page touched, then registered with the CUDA driver (cudaHostRegister, portable so that any device's
 * copy engine may use it). Where the platform reports no node for the device (numa_node = -1: single-node hosts and most
 * virtual machines, including the 8 x B200 boxes this was measured on) there is nothing to bind to and the allocation is
 * an ordinary cudaHostAlloc. SDRM_PINNED_MODE = cuda | mmap | thp overrides the choice (measurements).
 */
#define SDRM_MPOL_BIND 2
#define SDRM_HUGE (2u << 20)

struct pinned_region {
    void *base;
    size_t bytes;
    struct pinned_region *next;
};

static struct pinned_region *pinned_regions;
static pthread_mutex_t pinned_lock = PTHREAD_MUTEX_INITIALIZER;

static void *pinned_mmap(size_t bytes, int node, int huge) {
    const size_t len = (bytes + SDRM_HUGE - 1) / SDRM_HUGE * SDRM_HUGE;
    /* over-allocate to align the region to 2 MB, then trim */
    char *raw = mmap(NULL, len + SDRM_HUGE, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (raw == MAP_FAILED) {
        return NULL;
    }
    char *base = (char *) (((uintptr_t) raw + SDRM_HUGE - 1) & ~((uintptr_t) SDRM_HUGE - 1));
    if (base > raw) {
        munmap(raw, (size_t) (base - raw));
    }
    const size_t tail = (size_t) (raw + len + SDRM_HUGE - (base + len));
    if (tail > 0) {
        munmap(base + len, tail);
    }
    if (node >= 0 && node < 1024) {
        unsigned long mask[1024 / (8 * sizeof(unsigned long))];
        memset(mask, 0, sizeof(mask));
        mask[(size_t) node / (8 * sizeof(unsigned long))] |= 1UL << ((size_t) node % (8 * sizeof(unsigned long)));
        /* best effort: without the capability or the node the pages fall back to first touch */
        syscall(SYS_mbind, base, len, SDRM_MPOL_BIND, mask, (unsigned long) (node + 2), 0UL);
    }
#ifdef MADV_HUGEPAGE
    if (huge) {
        madvise(base, len, MADV_HUGEPAGE);
    }
#else
    (void) huge;
#endif
    for (size_t off = 0; off < len; off += 4096) {
        base[off] = 0;
    }
    if (cudaHostRegister(base, len, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        munmap(base, len);
        return NULL;
    }
    struct pinned_region *r = malloc(sizeof(*r));
    if (r == NULL) {
        cudaHostUnregister(base);
        munmap(base, len);
        return NULL;
    }
    r->base = base;
    r->bytes = len;
    pthread_mutex_lock(&pinned_lock);
    r->next = pinned_regions;
    pinned_regions = r;
    pthread_mutex_unlock(&pinned_lock);
    return base;
}

void *sdrm_pinned_alloc_near_device(size_t bytes, int device) {
    if (bytes == 0) {
        bytes = 1;
    }
    const char *mode = getenv("SDRM_PINNED_MODE");
    const int node = device >= 0 ? sdrm_device_numa_node(device) : -1;
    int use_mmap = node >= 0;
    int huge = 1;
    if (mode != NULL) {
        if (strcmp(mode, "cuda") == 0) {
            use_mmap = 0;
        } else if (strcmp(mode, "mmap") == 0) {
            use_mmap = 1;
            huge = 0;
        } else if (strcmp(mode, "thp") == 0) {
            use_mmap = 1;
        }
    }
    if (use_mmap) {
        void *p = pinned_mmap(bytes, node, huge);
        if (p != NULL) {
            return p;
        }
    }
    void *p = NULL;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return NULL;
    }
    return p;
}

/* returns 1 when p was one of pinned_mmap's regions (now released), 0 otherwise */
int sdrm_pinned_region_release(void *p) {
    struct pinned_region *found = NULL;
    pthread_mutex_lock(&pinned_lock);
    for (struct pinned_region **link = &pinned_regions; *link != NULL; link = &(*link)->next) {
        if ((*link)->base == p) {
            found = *link;
            *link = found->next;
            break;
        }
    }
    pthread_mutex_unlock(&pinned_lock);
    if (found == NULL) {
        return 0;
    }
    cudaHostUnregister(found->base);
    munmap(found->base, found->bytes);
    free(found);
    return 1;
}
