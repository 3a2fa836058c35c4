/*
 * Batched TX dispatcher (include/sdrm/tx_group.h): the transmit chain of the reference's tcp_worker
 * (src/tcp_server.c:175-241, built in src/tcp_server.c:491-570) for N sessions, device-resident between the stages.
 */
#define _POSIX_C_SOURCE 200809L

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/sdrm/sdrm_batch.h"
#include "../../include/sdrm/tx_group.h"
#include "sdrm_internal.h"

#define IN_FLIGHT 2
#define STATUS_INTERNAL_ERROR 3 /* RESPONSE_DETAILS_INTERNAL_ERROR, src/api.h:18 */

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

struct tx_row {
    sdrm_tx_session session;
    uint32_t caller_index;
    FILE *dump;
    int failed; /* during the current process call */
};

struct sdrm_tx_group_t {
    int device;
    uint32_t n;     /* sessions = rows */
    uint32_t n_dop; /* rows [0, n_dop): doppler_process_tx */
    uint32_t n_off; /* rows [n_dop, n_dop + n_off): sig_source_multiply(tx_offset); the rest goes out as modulated */
    struct tx_row *rows;
    int64_t *offsets; /* tx_offset of the n_off rows */
    uint32_t buffer_size;
    size_t per_byte;  /* samples per input byte */
    size_t stride;    /* complex samples per row of the device and host sample buffers */
    bool output_int16;
    float scalar;
    bool any_dump;
    sdrm_gfsk_mod_batch *mod;
    sdrm_doppler_batch *dopp;
    sdrm_nco_batch *nco;
    uint8_t *h_bytes[IN_FLIGHT]; /* pinned [n][buffer_size], row order */
    void *d_bytes[IN_FLIGHT];
    void *d_mod[IN_FLIGHT]; /* cf32 [n][stride]: modulator output */
    void *d_mix[IN_FLIGHT]; /* cf32 [n_dop + n_off][stride]: after the frequency correction */
    void *d_i16[IN_FLIGHT]; /* int16 pairs [n][stride] */
    float complex *h_cf32[IN_FLIGHT]; /* pinned [n][stride]; only filled when a sink or a dump file wants cf32 */
    int16_t *h_i16[IN_FLIGHT];        /* pinned [n][stride][2] */
    size_t batch_bytes[IN_FLIGHT];
    cudaStream_t copy_stream;
    cudaStream_t out_stream;
    cudaEvent_t ev_in[IN_FLIGHT];
    cudaEvent_t ev_mod[IN_FLIGHT];
    cudaEvent_t ev_mix[IN_FLIGHT][2];
    cudaEvent_t ev_out[IN_FLIGHT];
};

static int enqueue_batch(sdrm_tx_group *g, const uint8_t *data, size_t stride, size_t offset, size_t batch, int slot) {
    /* gather the sessions' bytes in row order: one contiguous host->device copy */
    for (uint32_t r = 0; r < g->n; r++) {
        memcpy(g->h_bytes[slot] + (size_t) r * g->buffer_size, data + (size_t) g->rows[r].caller_index * stride + offset, batch);
    }
    const size_t n_out = batch * g->per_byte;
    g->batch_bytes[slot] = batch;
    SDRM_CUDA_TRY(cudaMemcpyAsync(g->d_bytes[slot], g->h_bytes[slot], (size_t) g->n * g->buffer_size, cudaMemcpyHostToDevice,
                                  g->copy_stream));
    SDRM_CUDA_TRY(cudaEventRecord(g->ev_in[slot], g->copy_stream));
    cudaStream_t mod_in = (cudaStream_t) sdrm_gfsk_mod_batch_input_stream(g->mod);
    cudaStream_t mod_out = (cudaStream_t) sdrm_gfsk_mod_batch_stream(g->mod);
    SDRM_CUDA_TRY(cudaStreamWaitEvent(mod_in, g->ev_in[slot], 0));
    int code = sdrm_gfsk_mod_batch_process_device(g->mod, g->d_bytes[slot], g->buffer_size, batch, g->d_mod[slot], g->stride);
    if (code != 0) {
        return code;
    }
    SDRM_CUDA_TRY(cudaEventRecord(g->ev_mod[slot], mod_out));
    SDRM_CUDA_TRY(cudaStreamWaitEvent(g->out_stream, g->ev_mod[slot], 0));
    if (g->n_dop > 0) {
        cudaStream_t s = (cudaStream_t) sdrm_doppler_batch_stream(g->dopp);
        SDRM_CUDA_TRY(cudaStreamWaitEvent(s, g->ev_mod[slot], 0));
        code = sdrm_doppler_batch_process_device(g->dopp, -1, g->d_mod[slot], g->stride, n_out, g->d_mix[slot], g->stride);
        if (code != 0) {
            return code;
        }
        SDRM_CUDA_TRY(cudaEventRecord(g->ev_mix[slot][0], s));
        SDRM_CUDA_TRY(cudaStreamWaitEvent(g->out_stream, g->ev_mix[slot][0], 0));
    }
    if (g->n_off > 0) {
        cudaStream_t s = (cudaStream_t) sdrm_nco_batch_stream(g->nco);
        SDRM_CUDA_TRY(cudaStreamWaitEvent(s, g->ev_mod[slot], 0));
        const size_t skip = (size_t) g->n_dop * g->stride * 8;
        code = sdrm_nco_batch_process_device(g->nco, g->offsets, (const char *) g->d_mod[slot] + skip, g->stride, n_out,
                                             (char *) g->d_mix[slot] + skip, g->stride);
        if (code != 0) {
            return code;
        }
        SDRM_CUDA_TRY(cudaEventRecord(g->ev_mix[slot][1], s));
        SDRM_CUDA_TRY(cudaStreamWaitEvent(g->out_stream, g->ev_mix[slot][1], 0));
    }
    /* rows [0, mixed) leave from d_mix, the others straight from the modulator's buffer */
    const uint32_t mixed = g->n_dop + g->n_off;
    const size_t row_bytes = g->stride * 8;
    if (g->output_int16) {
        if (mixed > 0) {
            code = sdrm_samples_cf32_to_i16_device(g->d_mix[slot], g->stride, g->d_i16[slot], g->stride, g->scalar, n_out, mixed,
                                                   g->out_stream);
        }
        if (code == 0 && mixed < g->n) {
            code = sdrm_samples_cf32_to_i16_device((const char *) g->d_mod[slot] + mixed * row_bytes, g->stride,
                                                   (char *) g->d_i16[slot] + (size_t) mixed * g->stride * 4, g->stride, g->scalar,
                                                   n_out, g->n - mixed, g->out_stream);
        }
        if (code != 0) {
            return code;
        }
        SDRM_CUDA_TRY(cudaMemcpy2DAsync(g->h_i16[slot], g->stride * 4, g->d_i16[slot], g->stride * 4, n_out * 4, g->n,
                                        cudaMemcpyDeviceToHost, g->out_stream));
    }
    if (!g->output_int16 || g->any_dump) {
        if (mixed > 0) {
            SDRM_CUDA_TRY(cudaMemcpy2DAsync(g->h_cf32[slot], row_bytes, g->d_mix[slot], row_bytes, n_out * 8, mixed,
                                            cudaMemcpyDeviceToHost, g->out_stream));
        }
        if (mixed < g->n) {
            SDRM_CUDA_TRY(cudaMemcpy2DAsync((char *) g->h_cf32[slot] + mixed * row_bytes, row_bytes,
                                            (const char *) g->d_mod[slot] + mixed * row_bytes, row_bytes, n_out * 8, g->n - mixed,
                                            cudaMemcpyDeviceToHost, g->out_stream));
        }
    }
    SDRM_CUDA_TRY(cudaEventRecord(g->ev_out[slot], g->out_stream));
    return 0;
}

/* a finished batch -> dump files and sinks, session by session (src/tcp_server.c:214-230) */
static int deliver_batch(sdrm_tx_group *g, int slot) {
    SDRM_CUDA_TRY(cudaEventSynchronize(g->ev_out[slot]));
    const size_t n_out = g->batch_bytes[slot] * g->per_byte;
    for (uint32_t r = 0; r < g->n; r++) {
        struct tx_row *row = &g->rows[r];
        if (row->failed) {
            continue;
        }
        if (row->dump != NULL) {
            const float complex *v = g->h_cf32[slot] + (size_t) r * g->stride;
            if (fwrite(v, sizeof(float complex), n_out, row->dump) < n_out) {
                SDRM_LOG_ERROR("[%u] unable to write tx data", row->session.id); /* full disk: keep transmitting */
            }
        }
        if (row->session.sink != NULL) {
            const void *samples = g->output_int16 ? (const void *) (g->h_i16[slot] + (size_t) r * g->stride * 2)
                                                  : (const void *) (g->h_cf32[slot] + (size_t) r * g->stride);
            if (row->session.sink(row->session.sink_ctx, row->session.id, samples, n_out) != 0) {
                SDRM_LOG_ERROR("[%u] unable to transmit request fully", row->session.id);
                row->failed = 1;
            }
        }
    }
    return 0;
}

int sdrm_tx_group_process(sdrm_tx_group *g, const uint8_t *data, size_t stride, size_t len, int *session_status) {
    if (g == NULL || (data == NULL && len > 0) || stride < len) {
        return -1;
    }
    SDRM_CUDA_TRY(cudaSetDevice(g->device));
    for (uint32_t r = 0; r < g->n; r++) {
        g->rows[r].failed = 0;
    }
    size_t left = len;
    size_t processed = 0;
    uint64_t k = 0;
    int in_flight = 0;
    int code = 0;
    while (left > 0 && code == 0) {
        const size_t batch = left < g->buffer_size ? left : g->buffer_size;
        if (in_flight == IN_FLIGHT) {
            code = deliver_batch(g, (int) ((k - IN_FLIGHT) % IN_FLIGHT));
            in_flight--;
            if (code != 0) {
                break;
            }
        }
        code = enqueue_batch(g, data, stride, processed, batch, (int) (k % IN_FLIGHT));
        if (code == 0) {
            k++;
            in_flight++;
        }
        left -= batch;
        processed += batch;
    }
    while (in_flight > 0) {
        const int c = deliver_batch(g, (int) ((k - (uint64_t) in_flight) % IN_FLIGHT));
        if (code == 0) {
            code = c;
        }
        in_flight--;
    }
    if (session_status != NULL) {
        for (uint32_t r = 0; r < g->n; r++) {
            session_status[g->rows[r].caller_index] = g->rows[r].failed ? STATUS_INTERNAL_ERROR : 0;
        }
    }
    return code;
}

size_t sdrm_tx_group_samples_per_byte(const sdrm_tx_group *g) { return g->per_byte; }

static int open_dump(struct tx_row *row, const char *base_path) {
    if (base_path == NULL) {
        return -1;
    }
    char path[4096];
    snprintf(path, sizeof(path), "%s/tx.mod2sdr.%u.cf32", base_path, row->session.id);
    row->dump = fopen(path, "wb");
    if (row->dump == NULL) {
        SDRM_LOG_ERROR("[%u] unable to open file for tx output: %s", row->session.id, path);
        return -1;
    }
    return 0;
}

int sdrm_tx_group_create(const sdrm_tx_group_config *config, const sdrm_tx_session *sessions, uint32_t n_sessions,
                         sdrm_tx_group **out) {
    if (config == NULL || sessions == NULL || n_sessions == 0 || out == NULL || config->buffer_size == 0 ||
        config->mod_baud_rate == 0 || config->tx_sampling_freq == 0) {
        return -1;
    }
    /* src/tcp_server.c:529,537: the modulator gets the float ratio, the buffers the truncated one */
    const float sps = (float) ((double) config->tx_sampling_freq / config->mod_baud_rate);
    const int sps_int = (int) ((double) config->tx_sampling_freq / config->mod_baud_rate);
    if (sps_int < 1 || sps_int > 255) {
        return -1; /* interp_fir_filter takes the interpolation as uint8_t (src/dsp/interp_fir_filter.h:9) */
    }
    sdrm_tx_group *g = calloc(1, sizeof(*g));
    if (g == NULL) {
        return -ENOMEM;
    }
    g->n = n_sessions;
    g->buffer_size = config->buffer_size;
    g->per_byte = (size_t) 8 * (size_t) sps_int;
    g->stride = sdrm_round_up((size_t) config->buffer_size * g->per_byte, 2);
    g->output_int16 = config->output_int16;
    g->scalar = config->int16_scalar != 0.0f ? config->int16_scalar : 32768.0f;
    g->rows = calloc(n_sessions, sizeof(*g->rows));
    g->offsets = calloc(n_sessions, sizeof(*g->offsets));
    sdrm_doppler_channel *channels = calloc(n_sessions, sizeof(*channels));
    int code = (g->rows == NULL || g->offsets == NULL || channels == NULL) ? -ENOMEM : 0;
    if (code == 0) {
        if (config->device >= 0) {
            g->device = config->device;
        } else {
            code = sdrm_cuda_code(cudaGetDevice(&g->device), "cudaGetDevice");
        }
    }
    if (code == 0) code = sdrm_cuda_code(cudaSetDevice(g->device), "cudaSetDevice");
    if (code == 0) {
        /* rows: doppler sessions, then offset-only sessions, then the rest; each class in the caller's order */
        uint32_t r = 0;
        for (int cls = 0; cls < 3; cls++) {
            for (uint32_t i = 0; i < n_sessions; i++) {
                const sdrm_tx_session *s = &sessions[i];
                const int is = s->has_doppler ? 0 : (s->tx_offset != 0 ? 1 : 2);
                if (is != cls) {
                    continue;
                }
                if (cls == 0) {
                    sdrm_doppler_channel *c = &channels[g->n_dop++];
                    /* same scalings and constant offset as src/tcp_server.c:549 */
                    c->latitude = s->doppler_latitude / 10E6;
                    c->longitude = s->doppler_longitude / 10E6;
                    c->altitude = s->doppler_altitude / 10E3;
                    c->constant_offset = s->tx_offset;
                    c->start_time_seconds = s->file_start_time_seconds;
                    memcpy(c->tle, s->doppler_tle, sizeof(c->tle));
                } else if (cls == 1) {
                    g->offsets[g->n_off++] = s->tx_offset;
                }
                g->rows[r].session = *s;
                g->rows[r].caller_index = i;
                r++;
            }
        }
    }
    const uint32_t max_samples = (uint32_t) g->stride;
    if (code == 0) {
        code = sdrm_gfsk_mod_batch_create(n_sessions, sps, (float) (2 * M_PI * (double) config->mod_fsk_deviation /
                                                                   (double) config->tx_sampling_freq),
                                          0.5F, config->buffer_size, g->device, &g->mod);
        if (code != 0) {
            SDRM_LOG_ERROR("unable to create fsk modulator for %u sessions", n_sessions);
        }
    }
    if (code == 0 && g->n_dop > 0) {
        code = sdrm_doppler_batch_create(g->n_dop, channels, config->tx_sampling_freq, config->tx_center_freq, max_samples,
                                         g->device, &g->dopp);
        if (code != 0) {
            SDRM_LOG_ERROR("unable to create tx doppler correction for %u sessions", g->n_dop);
        }
    }
    free(channels);
    if (code == 0 && g->n_off > 0) {
        code = sdrm_nco_batch_create(g->n_off, 1.0F, config->tx_sampling_freq, max_samples, g->device, &g->nco);
        if (code != 0) {
            SDRM_LOG_ERROR("unable to create freq correction for %u sessions", g->n_off);
        }
    }
    for (uint32_t r = 0; r < n_sessions && code == 0; r++) {
        if (g->rows[r].session.tx_dump_file) {
            g->any_dump = true;
            code = open_dump(&g->rows[r], config->base_path);
        }
    }
    const size_t sample_bytes = (size_t) n_sessions * g->stride * 8;
    const uint32_t mixed = g->n_dop + g->n_off;
    for (int s = 0; s < IN_FLIGHT && code == 0; s++) {
        g->h_bytes[s] = sdrm_pinned_alloc((size_t) n_sessions * config->buffer_size);
        if (g->h_bytes[s] == NULL) code = -ENOMEM;
        if (code == 0) code = sdrm_dev_zalloc(&g->d_bytes[s], (size_t) n_sessions * config->buffer_size);
        if (code == 0) code = sdrm_dev_zalloc(&g->d_mod[s], sample_bytes);
        if (code == 0 && mixed > 0) code = sdrm_dev_zalloc(&g->d_mix[s], (size_t) mixed * g->stride * 8);
        if (code == 0 && g->output_int16) {
            code = sdrm_dev_zalloc(&g->d_i16[s], sample_bytes / 2);
            if (code == 0) {
                g->h_i16[s] = sdrm_pinned_alloc(sample_bytes / 2);
                if (g->h_i16[s] == NULL) code = -ENOMEM;
            }
        }
        if (code == 0 && (!g->output_int16 || g->any_dump)) {
            g->h_cf32[s] = sdrm_pinned_alloc(sample_bytes);
            if (g->h_cf32[s] == NULL) code = -ENOMEM;
        }
        if (code == 0) code = sdrm_cuda_code(cudaEventCreateWithFlags(&g->ev_in[s], cudaEventDisableTiming), "event");
        if (code == 0) code = sdrm_cuda_code(cudaEventCreateWithFlags(&g->ev_mod[s], cudaEventDisableTiming), "event");
        if (code == 0) code = sdrm_cuda_code(cudaEventCreateWithFlags(&g->ev_mix[s][0], cudaEventDisableTiming), "event");
        if (code == 0) code = sdrm_cuda_code(cudaEventCreateWithFlags(&g->ev_mix[s][1], cudaEventDisableTiming), "event");
        if (code == 0) code = sdrm_cuda_code(cudaEventCreateWithFlags(&g->ev_out[s], cudaEventDisableTiming), "event");
    }
    if (code == 0) code = sdrm_cuda_code(cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking), "stream");
    if (code == 0) code = sdrm_cuda_code(cudaStreamCreateWithFlags(&g->out_stream, cudaStreamNonBlocking), "stream");
    if (code != 0) {
        sdrm_tx_group_destroy(g);
        return code;
    }
    *out = g;
    return 0;
}

void sdrm_tx_group_destroy(sdrm_tx_group *g) {
    if (g == NULL) {
        return;
    }
    cudaSetDevice(g->device);
    if (g->copy_stream != NULL) cudaStreamSynchronize(g->copy_stream);
    if (g->out_stream != NULL) cudaStreamSynchronize(g->out_stream);
    sdrm_gfsk_mod_batch_destroy(g->mod);
    sdrm_doppler_batch_destroy(g->dopp);
    sdrm_nco_batch_destroy(g->nco);
    for (int s = 0; s < IN_FLIGHT; s++) {
        sdrm_pinned_free(g->h_bytes[s]);
        sdrm_pinned_free(g->h_cf32[s]);
        sdrm_pinned_free(g->h_i16[s]);
        cudaFree(g->d_bytes[s]);
        cudaFree(g->d_mod[s]);
        cudaFree(g->d_mix[s]);
        cudaFree(g->d_i16[s]);
        if (g->ev_in[s] != NULL) cudaEventDestroy(g->ev_in[s]);
        if (g->ev_mod[s] != NULL) cudaEventDestroy(g->ev_mod[s]);
        if (g->ev_mix[s][0] != NULL) cudaEventDestroy(g->ev_mix[s][0]);
        if (g->ev_mix[s][1] != NULL) cudaEventDestroy(g->ev_mix[s][1]);
        if (g->ev_out[s] != NULL) cudaEventDestroy(g->ev_out[s]);
    }
    if (g->copy_stream != NULL) cudaStreamDestroy(g->copy_stream);
    if (g->out_stream != NULL) cudaStreamDestroy(g->out_stream);
    if (g->rows != NULL) {
        for (uint32_t r = 0; r < g->n; r++) {
            if (g->rows[r].dump != NULL) {
                fclose(g->rows[r].dump);
            }
        }
    }
    free(g->rows);
    free(g->offsets);
    free(g);
}
