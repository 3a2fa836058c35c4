/*
 * Batched RX dispatcher (include/sdrm/rx_group.h): the sdr_worker -> N x dsp_worker fan-out of the reference
 * (src/sdr_worker.c:25-55, src/dsp_worker.c:44-106) as one queue, one thread and one set of launches per block.
 */
#define _POSIX_C_SOURCE 200809L

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../../include/sdrm/queue.h"
#include "../../include/sdrm/rx_group.h"
#include "../../include/sdrm/sdrm_batch.h"
#include "sdrm_internal.h"

#define IN_FLIGHT 2

struct sdrm_rx_group_t {
    int device;
    uint32_t n;        /* sessions */
    uint32_t n_dop;    /* the first n_dop rows are the sessions with doppler correction */
    sdrm_rx_session *rows; /* sessions in row order */
    uint32_t buffer_size;
    size_t stride;     /* float2 per row of the corrected buffers */
    queue *queue;
    sdrm_doppler_batch *dopp;
    sdrm_fsk_demod_batch *demod;
    void *d_block[IN_FLIGHT];     /* the SDR block as it arrived */
    void *d_corrected[IN_FLIGHT]; /* cf32 [n][stride]: demodulator input */
    cudaStream_t copy_stream;
    cudaEvent_t copied[IN_FLIGHT];
    cudaEvent_t prepared[IN_FLIGHT]; /* doppler / fan-out of a slot finished */
    cudaEvent_t filtered[IN_FLIGHT]; /* the demodulator's filters have consumed a slot */
    int8_t *h_symbols;   /* [n][cap] */
    uint32_t *h_lens;
    size_t cap;
    pthread_t thread;
    int thread_started;
    volatile uint64_t blocks_done;
    int failed;
    uint8_t *session_failed; /* [n]: a session whose client could not be written to is no longer served (the reference's
                                dsp_worker thread ends on the first failed write, src/dsp_worker.c:93-103) */
    volatile uint32_t sessions_failed;
};

static int write_all(const uint8_t *buffer, size_t len, int fd) {
    size_t left = len;
    while (left > 0) {
        const ssize_t written = write(fd, buffer + (len - left), left);
        if (written < 0) {
            return -1;
        }
        left -= (size_t) written;
    }
    return 0;
}

/* oldest call in flight -> host -> sinks */
static int deliver(sdrm_rx_group *g) {
    int code = sdrm_fsk_demod_batch_fetch(g->demod, g->h_symbols, NULL, g->cap, g->h_lens);
    if (code != 0) {
        return code;
    }
    for (uint32_t r = 0; r < g->n; r++) {
        const sdrm_rx_session *s = &g->rows[r];
        const int8_t *symbols = g->h_symbols + (size_t) r * g->cap;
        const size_t len = g->h_lens[r];
        if (len == 0 || g->session_failed[r]) {
            continue;
        }
        if (s->sink != NULL) {
            s->sink(s->sink_ctx, s->id, symbols, len);
        } else if (s->client_socket >= 0) {
            if (write_all((const uint8_t *) symbols, len, s->client_socket) != 0) {
                SDRM_LOG_ERROR("[%d] unable to write demod data to the client: session stopped", s->id);
                g->session_failed[r] = 1;
                g->sessions_failed++;
            }
        }
    }
    g->blocks_done++;
    return 0;
}

static int enqueue_block(sdrm_rx_group *g, const float complex *input, size_t len, int slot) {
    cudaStream_t fir = (cudaStream_t) sdrm_fsk_demod_batch_stream(g->demod);
    cudaStream_t prep = g->dopp != NULL ? (cudaStream_t) sdrm_doppler_batch_stream(g->dopp) : g->copy_stream;
    /* d_block[slot] was last read by the preparation of the call two blocks ago */
    SDRM_CUDA_TRY(cudaStreamWaitEvent(g->copy_stream, g->prepared[slot], 0));
    SDRM_CUDA_TRY(cudaMemcpyAsync(g->d_block[slot], input, len * 8, cudaMemcpyHostToDevice, g->copy_stream));
    SDRM_CUDA_TRY(cudaEventRecord(g->copied[slot], g->copy_stream));
    /* d_corrected[slot] is free once the filters of the call two blocks ago are through with it */
    SDRM_CUDA_TRY(cudaStreamWaitEvent(prep, g->copied[slot], 0));
    SDRM_CUDA_TRY(cudaStreamWaitEvent(prep, g->filtered[slot], 0));
    if (g->n_dop > 0) {
        /* in_stride 0: every session reads the same block */
        int code = sdrm_doppler_batch_process_device(g->dopp, 1, g->d_block[slot], 0, len, g->d_corrected[slot], g->stride);
        if (code != 0) {
            return code;
        }
    }
    for (uint32_t r = g->n_dop; r < g->n; r++) {
        SDRM_CUDA_TRY(cudaMemcpyAsync((char *) g->d_corrected[slot] + (size_t) r * g->stride * 8, g->d_block[slot], len * 8,
                                      cudaMemcpyDeviceToDevice, prep));
    }
    SDRM_CUDA_TRY(cudaEventRecord(g->prepared[slot], prep));
    SDRM_CUDA_TRY(cudaStreamWaitEvent(fir, g->prepared[slot], 0));
    int code = sdrm_fsk_demod_batch_process_device(g->demod, g->d_corrected[slot], g->stride, len);
    if (code != 0) {
        return code;
    }
    SDRM_CUDA_TRY(cudaEventRecord(g->filtered[slot], fir));
    /* the pinned queue slot goes back to the producer as soon as the copy has left it */
    SDRM_CUDA_TRY(cudaEventSynchronize(g->copied[slot]));
    return 0;
}

static void *rx_group_thread(void *arg) {
    sdrm_rx_group *g = arg;
    if (cudaSetDevice(g->device) != cudaSuccess) {
        g->failed = 1;
        return NULL;
    }
    int in_flight = 0;
    uint64_t k = 0;
    while (true) {
        float complex *input = NULL;
        size_t len = 0;
        take_buffer_for_processing(&input, &len, g->queue);
        if (input == NULL) {
            break; /* poison pill and nothing left */
        }
        int code = 0;
        if (len > g->buffer_size) {
            SDRM_LOG_ERROR("requested buffer %zu is more than max: %u", len, g->buffer_size);
            code = -1;
        } else if (len > 0) {
            code = enqueue_block(g, input, len, (int) (k % IN_FLIGHT));
        }
        complete_buffer_processing(g->queue);
        if (code != 0 && !(len > g->buffer_size)) {
            /* a block that could not be enqueued leaves every session's stream with a hole: stop, do not skip */
            SDRM_LOG_ERROR("rx group: block %llu failed with %d, group stopped", (unsigned long long) k, code);
            g->failed = 1;
            break;
        }
        if (code != 0 || len == 0) {
            continue; /* oversize block: rejected like the reference's blocks reject it (NULL / 0 output) */
        }
        k++;
        in_flight++;
        while (in_flight > 0 && (in_flight == IN_FLIGHT || !sdrm_queue_has_data(g->queue))) {
            if (deliver(g) != 0) {
                g->failed = 1;
                break;
            }
            in_flight--;
        }
        if (g->failed) {
            break;
        }
    }
    while (!g->failed && in_flight > 0) {
        if (deliver(g) != 0) {
            g->failed = 1;
        }
        in_flight--;
    }
    return NULL;
}

int sdrm_rx_group_create(const sdrm_rx_group_config *config, const sdrm_rx_session *sessions, uint32_t n_sessions,
                         sdrm_rx_group **out) {
    if (config == NULL || sessions == NULL || n_sessions == 0 || out == NULL || config->buffer_size == 0) {
        return -1;
    }
    sdrm_rx_group *g = calloc(1, sizeof(*g));
    if (g == NULL) {
        return -ENOMEM;
    }
    g->n = n_sessions;
    g->buffer_size = config->buffer_size;
    g->stride = sdrm_round_up((size_t) config->buffer_size, 2) + 2;
    g->rows = calloc(n_sessions, sizeof(*g->rows));
    sdrm_doppler_channel *channels = calloc(n_sessions, sizeof(*channels));
    int code = (g->rows == NULL || channels == NULL) ? -ENOMEM : 0;
    if (code == 0) {
        if (config->device >= 0) {
            g->device = config->device;
        } else {
            code = sdrm_cuda_code(cudaGetDevice(&g->device), "cudaGetDevice");
        }
    }
    if (code == 0) code = sdrm_cuda_code(cudaSetDevice(g->device), "cudaSetDevice");
    if (code == 0) {
        /* rows: sessions with doppler first, each group in the caller's order */
        for (uint32_t i = 0; i < n_sessions; i++) {
            if (sessions[i].has_doppler) {
                const sdrm_rx_session *s = &sessions[i];
                sdrm_doppler_channel *c = &channels[g->n_dop];
                /* same scalings as src/dsp_worker.c:130 */
                c->latitude = s->doppler_latitude / 10E6;
                c->longitude = s->doppler_longitude / 10E6;
                c->altitude = s->doppler_altitude / 10E3;
                c->constant_offset = 0;
                c->start_time_seconds = s->file_start_time_seconds;
                memcpy(c->tle, s->doppler_tle, sizeof(c->tle));
                g->rows[g->n_dop++] = *s;
            }
        }
        uint32_t r = g->n_dop;
        for (uint32_t i = 0; i < n_sessions; i++) {
            if (!sessions[i].has_doppler) {
                g->rows[r++] = sessions[i];
            }
        }
    }
    if (code == 0 && g->n_dop > 0) {
        code = sdrm_doppler_batch_create(g->n_dop, channels, config->rx_sampling_freq, config->rx_center_freq, config->buffer_size,
                                         g->device, &g->dopp);
        if (code != 0) {
            SDRM_LOG_ERROR("unable to create doppler correction for %u sessions", g->n_dop);
        }
    }
    free(channels);
    if (code == 0) {
        sdrm_fsk_demod_batch_config dc;
        memset(&dc, 0, sizeof(dc));
        dc.n_channels = n_sessions;
        dc.sampling_freq = config->rx_sampling_freq;
        dc.baud_rate = config->demod_baud_rate;
        dc.deviation = config->demod_fsk_deviation;
        dc.decimation = (uint8_t) config->demod_decimation;
        dc.transition_width = config->demod_fsk_transition_width;
        dc.use_dc_block = config->demod_fsk_use_dc_block;
        dc.max_input_buffer_length = config->buffer_size;
        dc.device = g->device;
        code = sdrm_fsk_demod_batch_create(&dc, &g->demod);
        if (code != 0) {
            SDRM_LOG_ERROR("unable to create demodulator for %u sessions", n_sessions);
        }
    }
    g->cap = config->buffer_size;
    for (int s = 0; s < IN_FLIGHT && code == 0; s++) {
        code = sdrm_dev_zalloc(&g->d_block[s], g->stride * 8);
        if (code == 0) code = sdrm_dev_zalloc(&g->d_corrected[s], (size_t) n_sessions * g->stride * 8);
        if (code == 0) code = sdrm_cuda_code(cudaEventCreateWithFlags(&g->copied[s], cudaEventDisableTiming), "event");
        if (code == 0) code = sdrm_cuda_code(cudaEventCreateWithFlags(&g->prepared[s], cudaEventDisableTiming), "event");
        if (code == 0) code = sdrm_cuda_code(cudaEventCreateWithFlags(&g->filtered[s], cudaEventDisableTiming), "event");
    }
    if (code == 0) code = sdrm_cuda_code(cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking), "stream");
    if (code == 0) {
        g->h_symbols = malloc((size_t) n_sessions * g->cap);
        g->h_lens = calloc(n_sessions, sizeof(uint32_t));
        g->session_failed = calloc(n_sessions, 1);
        if (g->h_symbols == NULL || g->h_lens == NULL || g->session_failed == NULL) {
            code = -ENOMEM;
        }
    }
    if (code == 0) code = create_queue(config->buffer_size, config->queue_size, config->blocking_queue, &g->queue);
    if (code == 0 && pthread_create(&g->thread, NULL, &rx_group_thread, g) != 0) {
        code = -1;
    }
    if (code != 0) {
        sdrm_rx_group_destroy(g);
        return code;
    }
    g->thread_started = 1;
    *out = g;
    return 0;
}

void sdrm_rx_group_put(float complex *block, size_t len, sdrm_rx_group *g) { queue_put(block, len, g->queue); }

void sdrm_rx_group_shutdown(sdrm_rx_group *g) {
    if (g != NULL) {
        interrupt_waiting_the_data(g->queue);
    }
}

uint64_t sdrm_rx_group_blocks_done(const sdrm_rx_group *g) { return g->blocks_done; }

int sdrm_rx_group_failed(const sdrm_rx_group *g) { return g == NULL ? -1 : g->failed; }

uint32_t sdrm_rx_group_sessions_failed(const sdrm_rx_group *g) { return g == NULL ? 0 : g->sessions_failed; }

void sdrm_rx_group_destroy(sdrm_rx_group *g) {
    if (g == NULL) {
        return;
    }
    if (g->queue != NULL) {
        interrupt_waiting_the_data(g->queue);
    }
    if (g->thread_started) {
        pthread_join(g->thread, NULL);
    }
    if (g->queue != NULL) {
        destroy_queue(g->queue);
    }
    sdrm_fsk_demod_batch_destroy(g->demod);
    sdrm_doppler_batch_destroy(g->dopp);
    for (int s = 0; s < IN_FLIGHT; s++) {
        cudaFree(g->d_block[s]);
        cudaFree(g->d_corrected[s]);
        if (g->copied[s] != NULL) cudaEventDestroy(g->copied[s]);
        if (g->prepared[s] != NULL) cudaEventDestroy(g->prepared[s]);
        if (g->filtered[s] != NULL) cudaEventDestroy(g->filtered[s]);
    }
    if (g->copy_stream != NULL) {
        cudaStreamDestroy(g->copy_stream);
    }
    free(g->h_symbols);
    free(g->h_lens);
    free(g->session_failed);
    free(g->rows);
    free(g);
}
