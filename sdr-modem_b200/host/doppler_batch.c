/*
 * Batched Doppler correction and the reference's doppler handle on top of it (a batch of one).
 * Reference: src/dsp/doppler.c:31-42 (shift from the range rate), :44-114 (create), :116-190 (process: the frequency is
 * re-evaluated every `sampling_freq` samples with SGP4, interpolated linearly in between, held constant and truncated
 * to an integer number of Hz per call segment, then mixed in by sig_source_multiply).
 *
 * The segment logic below is the reference's, per channel in double precision on the host; all channels of a batch
 * have consumed the same number of samples, so they share the segment boundaries and each segment is one NCO launch.
 */
#define _POSIX_C_SOURCE 200809L

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/sdrm/doppler.h"
#include "../../include/sdrm/sdrm_batch.h"
#include "orbit.h"
#include "sdrm_internal.h"

static const double SPEED_OF_LIGHT_KM_S = 2.99792458E5;
static const double DEG_TO_RAD = 1.74532925E-2; /* the reference's Radians() constant */
static const double SECONDS_PER_DAY = 8.6400E4;

struct channel_state {
    sdrm_orbit orbit;
    double lat;
    double lon;
    double alt;
    int64_t constant_offset;
    double jul_start_time;
    double jul_utc;
    double current_freq_difference;
    double next_freq_difference;
    double freq_difference_per_sample;
};

struct sdrm_doppler_batch_t {
    int device; /* the NCO batch's device: staging buffers and copies must be issued with it current */
    uint32_t n_ch;
    uint32_t max_len;
    double sampling_freq;
    int64_t center_freq;
    uint64_t update_interval_samples;
    uint64_t current_samples;
    struct channel_state *ch;
    int64_t *freq;
    sdrm_nco_batch *nco;
    void *d_in;
    void *d_out;
    size_t stride;
};

static double shift_hz(struct sdrm_doppler_batch_t *b, struct channel_state *c, int direction) {
    const double range_rate = sdrm_orbit_range_rate(&c->orbit, c->jul_utc, c->lat, c->lon, c->alt);
    return (direction * (b->center_freq - b->center_freq * (SPEED_OF_LIGHT_KM_S - range_rate) / SPEED_OF_LIGHT_KM_S)) + c->constant_offset;
}

int sdrm_doppler_batch_create(uint32_t n_channels, const sdrm_doppler_channel *channels, uint64_t sampling_freq,
                              uint64_t center_freq, uint32_t max_output_buffer_length, int device, sdrm_doppler_batch **batch) {
    if (n_channels == 0 || channels == NULL || batch == NULL || sampling_freq == 0) {
        return -1;
    }
    struct sdrm_doppler_batch_t *b = calloc(1, sizeof(*b));
    if (b == NULL) {
        return -ENOMEM;
    }
    b->n_ch = n_channels;
    b->max_len = max_output_buffer_length;
    b->sampling_freq = (double) sampling_freq;
    b->center_freq = (int64_t) center_freq;
    b->update_interval_samples = sampling_freq; /* re-evaluate once per second of signal */
    b->current_samples = b->update_interval_samples;
    b->ch = calloc(n_channels, sizeof(struct channel_state));
    b->freq = calloc(n_channels, sizeof(int64_t));
    if (b->ch == NULL || b->freq == NULL) {
        sdrm_doppler_batch_destroy(b);
        return -ENOMEM;
    }
    for (uint32_t i = 0; i < n_channels; i++) {
        struct channel_state *c = &b->ch[i];
        c->lat = channels[i].latitude * DEG_TO_RAD;
        c->lon = channels[i].longitude * DEG_TO_RAD;
        c->alt = channels[i].altitude;
        c->constant_offset = channels[i].constant_offset;
        if (channels[i].start_time_seconds == 0) {
            c->jul_start_time = 0.0;
        } else {
            struct tm cdate;
            time_t start = (time_t) channels[i].start_time_seconds;
            if (gmtime_r(&start, &cdate) == NULL) {
                sdrm_doppler_batch_destroy(b);
                return -1;
            }
            c->jul_start_time = sdrm_julian_date(cdate.tm_year + 1900, cdate.tm_mon + 1, cdate.tm_mday, cdate.tm_hour, cdate.tm_min,
                                                 cdate.tm_sec);
        }
        const int code = sdrm_orbit_init(channels[i].tle, &c->orbit);
        if (code != 0) {
            SDRM_LOG_ERROR("invalid tle configuration");
            sdrm_doppler_batch_destroy(b);
            return -1;
        }
    }
    int code = 0;
    if (device >= 0) {
        b->device = device;
    } else {
        code = sdrm_cuda_code(cudaGetDevice(&b->device), "cudaGetDevice");
    }
    if (code == 0) {
        code = sdrm_nco_batch_create(n_channels, 1.0F, sampling_freq, max_output_buffer_length, b->device, &b->nco);
    }
    if (code != 0) {
        sdrm_doppler_batch_destroy(b);
        return code;
    }
    *batch = b;
    return 0;
}

/* One pass of the reference's while loop (doppler.c:132-186) for every channel; returns the segment length. */
static size_t next_segment(struct sdrm_doppler_batch_t *b, int direction, size_t remaining) {
    size_t batch_len;
    if (b->update_interval_samples < remaining + b->current_samples) {
        if (b->current_samples >= b->update_interval_samples) {
            batch_len = b->update_interval_samples < remaining ? b->update_interval_samples : remaining;
        } else {
            batch_len = b->update_interval_samples - b->current_samples;
        }
    } else {
        batch_len = remaining;
    }
    const int rollover = b->current_samples >= b->update_interval_samples;
    if (rollover) {
        b->current_samples = 0;
    }
    for (uint32_t i = 0; i < b->n_ch; i++) {
        struct channel_state *c = &b->ch[i];
        if (rollover) {
            if (c->next_freq_difference == 0) {
                if (c->jul_start_time == 0.0) {
                    /* lazily, as the reference: the stream starts when the first block arrives */
                    struct tm t;
                    time_t now = time(NULL);
                    gmtime_r(&now, &t);
                    c->jul_start_time = sdrm_julian_date(t.tm_year + 1900, t.tm_mon + 1, t.tm_mday, t.tm_hour, t.tm_min, t.tm_sec);
                }
                c->jul_utc = c->jul_start_time;
                c->current_freq_difference = shift_hz(b, c, direction);
            } else {
                c->current_freq_difference = c->next_freq_difference;
            }
            c->jul_utc += (double) b->update_interval_samples / b->sampling_freq / SECONDS_PER_DAY;
            c->next_freq_difference = shift_hz(b, c, direction);
            c->freq_difference_per_sample = (c->next_freq_difference - c->current_freq_difference) / b->update_interval_samples;
        } else {
            c->current_freq_difference += c->freq_difference_per_sample * (double) batch_len;
        }
        b->freq[i] = (int64_t) c->current_freq_difference;
    }
    b->current_samples += batch_len;
    return batch_len;
}

int sdrm_doppler_batch_process_device(sdrm_doppler_batch *b, int direction, const void *d_input, size_t in_stride, size_t len,
                                      void *d_output, size_t out_stride) {
    if (b == NULL || d_input == NULL || d_output == NULL) {
        return -1;
    }
    if (len > b->max_len) {
        SDRM_LOG_ERROR("requested buffer %zu is more than max: %u", len, b->max_len);
        return -1;
    }
    size_t processed = 0;
    size_t remaining = len;
    while (processed < len) {
        const size_t seg = next_segment(b, direction, remaining);
        int code = sdrm_nco_batch_process_device(b->nco, b->freq, (const char *) d_input + processed * 8, in_stride, seg,
                                                 (char *) d_output + processed * 8, out_stride);
        if (code != 0) {
            return code;
        }
        processed += seg;
        remaining -= seg;
    }
    return 0;
}

int sdrm_doppler_batch_process(sdrm_doppler_batch *b, int direction, const float complex *input, size_t in_stride, size_t len,
                               float complex *output, size_t out_stride) {
    if (b == NULL || input == NULL || output == NULL) {
        return -1;
    }
    if (len > b->max_len) {
        SDRM_LOG_ERROR("requested buffer %zu is more than max: %u", len, b->max_len);
        return -1;
    }
    if (len == 0) {
        return 0;
    }
    /* a worker thread's current device is 0 until it says otherwise: the staging buffers must live where the NCO runs */
    SDRM_CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t stream = (cudaStream_t) sdrm_nco_batch_stream(b->nco);
    if (b->d_in == NULL) {
        b->stride = sdrm_round_up((size_t) b->max_len, 2) + 2;
        int code = sdrm_dev_zalloc(&b->d_in, (size_t) b->n_ch * b->stride * 8);
        if (code == 0) code = sdrm_dev_zalloc(&b->d_out, (size_t) b->n_ch * b->stride * 8);
        if (code != 0) return code;
    }
    SDRM_CUDA_TRY(cudaMemcpy2DAsync(b->d_in, b->stride * 8, input, in_stride * 8, len * 8, b->n_ch, cudaMemcpyHostToDevice, stream));
    int code = sdrm_doppler_batch_process_device(b, direction, b->d_in, b->stride, len, b->d_out, b->stride);
    if (code != 0) return code;
    SDRM_CUDA_TRY(cudaMemcpy2DAsync(output, out_stride * 8, b->d_out, b->stride * 8, len * 8, b->n_ch, cudaMemcpyDeviceToHost, stream));
    SDRM_CUDA_TRY(cudaStreamSynchronize(stream));
    return 0;
}

int sdrm_doppler_batch_sync(sdrm_doppler_batch *b) { return b == NULL ? -1 : sdrm_nco_batch_sync(b->nco); }

void *sdrm_doppler_batch_stream(sdrm_doppler_batch *b) { return b == NULL ? NULL : sdrm_nco_batch_stream(b->nco); }

void sdrm_doppler_batch_destroy(sdrm_doppler_batch *b) {
    if (b == NULL) {
        return;
    }
    sdrm_nco_batch_destroy(b->nco);
    if (b->d_in != NULL || b->d_out != NULL) {
        cudaSetDevice(b->device);
    }
    cudaFree(b->d_in);
    cudaFree(b->d_out);
    free(b->ch);
    free(b->freq);
    free(b);
}

/* ---- the reference's single-stream handle ---------------------------------------------------------------------------------- */

struct doppler_t {
    sdrm_doppler_batch *batch;
    float complex *output;
    uint32_t output_len;
};

int doppler_create(double latitude, double longitude, double altitude, uint64_t sampling_freq, uint64_t center_freq,
                   int64_t constant_offset, time_t start_time_seconds, uint32_t max_output_buffer_length, char tle[3][80],
                   doppler **result) {
    struct doppler_t *d = calloc(1, sizeof(*d));
    if (d == NULL) {
        return -ENOMEM;
    }
    d->output_len = max_output_buffer_length;
    d->output = malloc(sizeof(float complex) * (max_output_buffer_length == 0 ? 1 : max_output_buffer_length));
    if (d->output == NULL) {
        doppler_destroy(d);
        return -ENOMEM;
    }
    sdrm_doppler_channel channel;
    memset(&channel, 0, sizeof(channel));
    channel.latitude = latitude;
    channel.longitude = longitude;
    channel.altitude = altitude;
    channel.constant_offset = constant_offset;
    channel.start_time_seconds = (int64_t) start_time_seconds;
    memcpy(channel.tle, tle, sizeof(channel.tle));
    int code = sdrm_doppler_batch_create(1, &channel, sampling_freq, center_freq, max_output_buffer_length, -1, &d->batch);
    if (code != 0) {
        doppler_destroy(d);
        return code;
    }
    *result = d;
    return 0;
}

static void doppler_run(float complex *input, size_t input_len, float complex **output, size_t *output_len, int direction, doppler *d) {
    *output = NULL;
    *output_len = 0;
    if (input == NULL || input_len == 0) {
        return; /* doppler.c:117-121 */
    }
    if (sdrm_doppler_batch_process(d->batch, direction, input, d->output_len, input_len, d->output, d->output_len) != 0) {
        return;
    }
    *output = d->output;
    *output_len = input_len;
}

void doppler_process_rx(float complex *input, size_t input_len, float complex **output, size_t *output_len, doppler *result) {
    doppler_run(input, input_len, output, output_len, 1, result);
}

void doppler_process_tx(float complex *input, size_t input_len, float complex **output, size_t *output_len, doppler *result) {
    doppler_run(input, input_len, output, output_len, -1, result);
}

void doppler_destroy(doppler *d) {
    if (d == NULL) {
        return;
    }
    sdrm_doppler_batch_destroy(d->batch);
    free(d->output);
    free(d);
}
