/*
 * Multi-GPU entry points (include/sdrm/sdrm_multi.h): channel sessions partitioned over a device list, one host thread,
 * stream set and pinned ingest ring per device, no collective.
 *
 * Reference analogue: src/sdr_worker.c:31-55 hands each SDR block to every dsp_worker of the stream, and every dsp_worker
 * is a thread of its own (src/dsp_worker.c:199-227). Here a "worker" is a GPU with a slice of the sessions.
 */

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/sdrm/sdrm_multi.h"
#include "sdrm_internal.h"

enum shard_command { CMD_NONE = 0, CMD_SUBMIT, CMD_SUBMIT_I16, CMD_FETCH, CMD_SYNC, CMD_STOP };

struct shard {
    uint32_t first; /* first channel of the slice */
    uint32_t count;
    int device;
    uint32_t max_len;
    sdrm_fsk_demod_batch *batch;

    pthread_t thread;
    int started;
    pthread_mutex_t lock;
    pthread_cond_t wake;
    pthread_cond_t finished;
    int command;
    int running;
    int result;

    /* arguments of the command in flight (pointers are those of the whole job; the shard offsets them by `first`) */
    const void *input;
    size_t in_stride;
    size_t input_len;
    float scalar;
    int8_t *output;
    float *soft;
    size_t out_stride;
    uint32_t *output_len;

    /* pinned ingest ring: pageable caller memory is packed into it by this shard's thread before the copy */
    void *stage[SDRM_MAX_IN_FLIGHT];
    size_t stage_bytes;
    uint64_t submitted;
};

struct sdrm_fsk_demod_multi_t {
    uint32_t n_shards;
    uint32_t n_channels;
    struct shard *shards;
};

static int is_pageable(const void *p) {
    struct cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

/* packs rows [first, first + count) of a pageable buffer into the shard's pinned slot; returns the slot and its stride */
static int stage_rows(struct shard *s, const char *input, size_t in_stride, size_t input_len, size_t elem, const void **staged,
                      size_t *staged_stride) {
    const int slot = (int) (s->submitted % SDRM_MAX_IN_FLIGHT);
    const size_t stride = (input_len + 1) & ~(size_t) 1;
    const size_t need = (size_t) s->count * ((size_t) s->max_len + 2) * 8;
    if (s->stage[slot] == NULL) {
        s->stage[slot] = sdrm_pinned_alloc_near_device(need, s->device);
        if (s->stage[slot] == NULL) {
            return -ENOMEM;
        }
        s->stage_bytes = need;
    }
    char *dst = s->stage[slot];
    for (uint32_t c = 0; c < s->count; c++) {
        memcpy(dst + (size_t) c * stride * elem, input + ((size_t) (s->first + c) * in_stride) * elem, input_len * elem);
    }
    *staged = dst;
    *staged_stride = stride;
    return 0;
}

static int run_command(struct shard *s, int command) {
    switch (command) {
    case CMD_SUBMIT:
    case CMD_SUBMIT_I16: {
        const size_t elem = command == CMD_SUBMIT ? 8 : 4;
        const void *src = (const char *) s->input + (size_t) s->first * s->in_stride * elem;
        size_t stride = s->in_stride;
        if (s->input_len > 0 && is_pageable(s->input)) {
            int code = stage_rows(s, s->input, s->in_stride, s->input_len, elem, &src, &stride);
            if (code != 0) {
                return code;
            }
        }
        int code = command == CMD_SUBMIT
                       ? sdrm_fsk_demod_batch_submit(s->batch, (const float complex *) src, stride, s->input_len)
                       : sdrm_fsk_demod_batch_submit_i16(s->batch, (const int16_t *) src, stride, s->input_len, s->scalar);
        if (code == 0) {
            s->submitted++;
        }
        return code;
    }
    case CMD_FETCH:
        return sdrm_fsk_demod_batch_fetch(s->batch, s->output != NULL ? s->output + (size_t) s->first * s->out_stride : NULL,
                                          s->soft != NULL ? s->soft + (size_t) s->first * s->out_stride : NULL, s->out_stride,
                                          s->output_len != NULL ? s->output_len + s->first : NULL);
    case CMD_SYNC:
        return sdrm_fsk_demod_batch_sync(s->batch);
    default:
        return 0;
    }
}

static void *shard_thread(void *arg) {
    struct shard *s = arg;
    /* this thread feeds one GPU: run next to it, so that the staging ring it allocates is local to that GPU too */
    cudaSetDevice(s->device);
    sdrm_bind_thread_near_device(s->device);
    pthread_mutex_lock(&s->lock);
    while (1) {
        while (s->command == CMD_NONE) {
            pthread_cond_wait(&s->wake, &s->lock);
        }
        const int command = s->command;
        pthread_mutex_unlock(&s->lock);
        const int result = command == CMD_STOP ? 0 : run_command(s, command);
        pthread_mutex_lock(&s->lock);
        s->result = result;
        s->command = CMD_NONE;
        s->running = 0;
        pthread_cond_signal(&s->finished);
        if (command == CMD_STOP) {
            break;
        }
    }
    pthread_mutex_unlock(&s->lock);
    return NULL;
}

/* hands `command` to every shard's thread and waits for all of them; the first failure is returned */
static int broadcast(sdrm_fsk_demod_multi *m, int command) {
    for (uint32_t g = 0; g < m->n_shards; g++) {
        struct shard *s = &m->shards[g];
        pthread_mutex_lock(&s->lock);
        s->command = command;
        s->running = 1;
        pthread_cond_signal(&s->wake);
        pthread_mutex_unlock(&s->lock);
    }
    int code = 0;
    for (uint32_t g = 0; g < m->n_shards; g++) {
        struct shard *s = &m->shards[g];
        pthread_mutex_lock(&s->lock);
        while (s->running) {
            pthread_cond_wait(&s->finished, &s->lock);
        }
        if (code == 0) {
            code = s->result;
        }
        pthread_mutex_unlock(&s->lock);
    }
    return code;
}

int sdrm_fsk_demod_multi_create(const sdrm_fsk_demod_batch_config *config, const int *devices, uint32_t n_devices,
                                sdrm_fsk_demod_multi **multi) {
    if (config == NULL || devices == NULL || multi == NULL || n_devices == 0 || n_devices > config->n_channels) {
        return -1;
    }
    sdrm_fsk_demod_multi *m = calloc(1, sizeof(*m));
    if (m == NULL) {
        return -ENOMEM;
    }
    m->shards = calloc(n_devices, sizeof(struct shard));
    if (m->shards == NULL) {
        free(m);
        return -ENOMEM;
    }
    m->n_shards = n_devices;
    m->n_channels = config->n_channels;
    int code = 0;
    for (uint32_t g = 0; g < n_devices && code == 0; g++) {
        struct shard *s = &m->shards[g];
        /* static partition by channel id */
        s->first = (uint32_t) ((uint64_t) g * config->n_channels / n_devices);
        s->count = (uint32_t) ((uint64_t) (g + 1) * config->n_channels / n_devices) - s->first;
        s->device = devices[g];
        s->max_len = config->max_input_buffer_length;
        sdrm_fsk_demod_batch_config shard_config = *config;
        shard_config.n_channels = s->count;
        shard_config.device = s->device;
        code = sdrm_fsk_demod_batch_create(&shard_config, &s->batch);
        if (code != 0) {
            break;
        }
        pthread_mutex_init(&s->lock, NULL);
        pthread_cond_init(&s->wake, NULL);
        pthread_cond_init(&s->finished, NULL);
        if (pthread_create(&s->thread, NULL, &shard_thread, s) != 0) {
            code = -1;
            break;
        }
        s->started = 1;
    }
    if (code != 0) {
        sdrm_fsk_demod_multi_destroy(m);
        return code;
    }
    *multi = m;
    return 0;
}

int sdrm_fsk_demod_multi_submit(sdrm_fsk_demod_multi *m, const float complex *input, size_t in_stride, size_t input_len) {
    if (m == NULL || (input == NULL && input_len > 0)) {
        return -1;
    }
    for (uint32_t g = 0; g < m->n_shards; g++) {
        struct shard *s = &m->shards[g];
        s->input = input;
        s->in_stride = in_stride;
        s->input_len = input_len;
    }
    return broadcast(m, CMD_SUBMIT);
}

int sdrm_fsk_demod_multi_submit_i16(sdrm_fsk_demod_multi *m, const int16_t *input, size_t in_stride, size_t input_len, float scalar) {
    if (m == NULL || (input == NULL && input_len > 0)) {
        return -1;
    }
    for (uint32_t g = 0; g < m->n_shards; g++) {
        struct shard *s = &m->shards[g];
        s->input = input;
        s->in_stride = in_stride;
        s->input_len = input_len;
        s->scalar = scalar;
    }
    return broadcast(m, CMD_SUBMIT_I16);
}

int sdrm_fsk_demod_multi_fetch(sdrm_fsk_demod_multi *m, int8_t *output, float *soft, size_t out_stride, uint32_t *output_len) {
    if (m == NULL) {
        return -1;
    }
    for (uint32_t g = 0; g < m->n_shards; g++) {
        struct shard *s = &m->shards[g];
        s->output = output;
        s->soft = soft;
        s->out_stride = out_stride;
        s->output_len = output_len;
    }
    return broadcast(m, CMD_FETCH);
}

int sdrm_fsk_demod_multi_process(sdrm_fsk_demod_multi *m, const float complex *input, size_t in_stride, size_t input_len,
                                 int8_t *output, float *soft, size_t out_stride, uint32_t *output_len) {
    int code = sdrm_fsk_demod_multi_submit(m, input, in_stride, input_len);
    if (code != 0) {
        return code;
    }
    return sdrm_fsk_demod_multi_fetch(m, output, soft, out_stride, output_len);
}

int sdrm_fsk_demod_multi_sync(sdrm_fsk_demod_multi *m) { return m == NULL ? -1 : broadcast(m, CMD_SYNC); }

uint32_t sdrm_fsk_demod_multi_device_count(const sdrm_fsk_demod_multi *m) { return m == NULL ? 0 : m->n_shards; }

int sdrm_fsk_demod_multi_shard(const sdrm_fsk_demod_multi *m, uint32_t g, uint32_t *first, uint32_t *count, int *device) {
    if (m == NULL || g >= m->n_shards) {
        return -1;
    }
    if (first != NULL) *first = m->shards[g].first;
    if (count != NULL) *count = m->shards[g].count;
    if (device != NULL) *device = m->shards[g].device;
    return 0;
}

sdrm_fsk_demod_batch *sdrm_fsk_demod_multi_batch(sdrm_fsk_demod_multi *m, uint32_t g) {
    return m == NULL || g >= m->n_shards ? NULL : m->shards[g].batch;
}

uint64_t sdrm_fsk_demod_multi_launch_count(const sdrm_fsk_demod_multi *m) {
    uint64_t total = 0;
    for (uint32_t g = 0; m != NULL && g < m->n_shards; g++) {
        total += sdrm_fsk_demod_batch_launch_count(m->shards[g].batch);
    }
    return total;
}

int sdrm_fsk_demod_multi_error_flags(sdrm_fsk_demod_multi *m) {
    if (m == NULL) {
        return -1;
    }
    int flags = 0;
    for (uint32_t g = 0; g < m->n_shards; g++) {
        const int f = sdrm_fsk_demod_batch_error_flags(m->shards[g].batch);
        if (f < 0) {
            return f;
        }
        flags |= f;
    }
    return flags;
}

void sdrm_fsk_demod_multi_destroy(sdrm_fsk_demod_multi *m) {
    if (m == NULL) {
        return;
    }
    for (uint32_t g = 0; g < m->n_shards; g++) {
        struct shard *s = &m->shards[g];
        if (s->started) {
            pthread_mutex_lock(&s->lock);
            s->command = CMD_STOP;
            s->running = 1;
            pthread_cond_signal(&s->wake);
            pthread_mutex_unlock(&s->lock);
            pthread_join(s->thread, NULL);
            pthread_mutex_destroy(&s->lock);
            pthread_cond_destroy(&s->wake);
            pthread_cond_destroy(&s->finished);
        }
        sdrm_fsk_demod_batch_destroy(s->batch);
        for (int k = 0; k < SDRM_MAX_IN_FLIGHT; k++) {
            sdrm_pinned_free(s->stage[k]);
        }
    }
    free(m->shards);
    free(m);
}

/* ---- the sdr_worker fan-out over several devices ------------------------------------------------------------------------- */

struct sdrm_rx_multi_t {
    uint32_t n_groups;
    sdrm_rx_group **groups;
};

int sdrm_rx_multi_create(const sdrm_rx_group_config *config, const sdrm_rx_session *sessions, uint32_t n_sessions,
                         const int *devices, uint32_t n_devices, sdrm_rx_multi **multi) {
    if (config == NULL || sessions == NULL || devices == NULL || multi == NULL || n_devices == 0 || n_devices > n_sessions) {
        return -1;
    }
    sdrm_rx_multi *m = calloc(1, sizeof(*m));
    if (m == NULL) {
        return -ENOMEM;
    }
    m->groups = calloc(n_devices, sizeof(sdrm_rx_group *));
    if (m->groups == NULL) {
        free(m);
        return -ENOMEM;
    }
    m->n_groups = n_devices;
    for (uint32_t g = 0; g < n_devices; g++) {
        const uint32_t first = (uint32_t) ((uint64_t) g * n_sessions / n_devices);
        const uint32_t count = (uint32_t) ((uint64_t) (g + 1) * n_sessions / n_devices) - first;
        sdrm_rx_group_config group_config = *config;
        group_config.device = devices[g];
        const int code = sdrm_rx_group_create(&group_config, sessions + first, count, &m->groups[g]);
        if (code != 0) {
            sdrm_rx_multi_destroy(m);
            return code;
        }
    }
    *multi = m;
    return 0;
}

void sdrm_rx_multi_put(float complex *block, size_t len, sdrm_rx_multi *m) {
    for (uint32_t g = 0; g < m->n_groups; g++) {
        sdrm_rx_group_put(block, len, m->groups[g]);
    }
}

void sdrm_rx_multi_shutdown(sdrm_rx_multi *m) {
    for (uint32_t g = 0; m != NULL && g < m->n_groups; g++) {
        sdrm_rx_group_shutdown(m->groups[g]);
    }
}

uint64_t sdrm_rx_multi_blocks_done(const sdrm_rx_multi *m) {
    uint64_t least = UINT64_MAX;
    for (uint32_t g = 0; m != NULL && g < m->n_groups; g++) {
        const uint64_t done = sdrm_rx_group_blocks_done(m->groups[g]);
        if (done < least) {
            least = done;
        }
    }
    return least == UINT64_MAX ? 0 : least;
}

int sdrm_rx_multi_failed(const sdrm_rx_multi *m) {
    for (uint32_t g = 0; m != NULL && g < m->n_groups; g++) {
        if (sdrm_rx_group_failed(m->groups[g])) {
            return 1;
        }
    }
    return 0;
}

void sdrm_rx_multi_destroy(sdrm_rx_multi *m) {
    if (m == NULL) {
        return;
    }
    for (uint32_t g = 0; g < m->n_groups; g++) {
        sdrm_rx_group_destroy(m->groups[g]);
    }
    free(m->groups);
    free(m);
}
