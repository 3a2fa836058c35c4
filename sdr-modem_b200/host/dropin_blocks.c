/*
 * The reference's per-block handles for the demod side, signature for signature, each running its kernel on the GPU
 * for one stream: lpf (src/dsp/lpf.h:10-14), quadrature_demod (quadrature_demod.h:9-13), dc_blocker (dc_blocker.h:7-11),
 * clock_mm (clock_recovery_mm.h:8-12), plus the host helpers create_low_pass_filter (lpf_taps.h:6) and fast_atan2f
 * (src/math/fast_atan2f.h:4). These exist so that code written against the reference's blocks links unchanged; the
 * throughput path is the fused batch in fsk_demod_batch.c.
 *
 * Contract kept from the reference: *_create returns 0 / -ENOMEM / -1 and writes *out only on success; *_process hands
 * back a buffer owned by the handle (valid until the next call), prints "<3>requested buffer …" and returns NULL / 0
 * when input_len exceeds the maximum; *_destroy accepts NULL.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/sdrm/clock_recovery_mm.h"
#include "../../include/sdrm/dc_blocker.h"
#include "../../include/sdrm/fast_atan2f.h"
#include "../../include/sdrm/fir_filter.h"
#include "../../include/sdrm/lpf.h"
#include "../../include/sdrm/lpf_taps.h"
#include "../../include/sdrm/quadrature_demod.h"
#include "sdrm_internal.h"

static void log_oversize(size_t requested, size_t max) {
    SDRM_LOG_ERROR("requested buffer %zu is more than max: %zu", requested, max);
}

/* ------------------------------------------------------------------------------------------------ tap design */

int create_low_pass_filter(float gain, uint64_t sampling_freq, uint64_t cutoff_freq, uint32_t transition_width,
                           float **taps, size_t *len) {
    return sdrm_design_low_pass(gain, sampling_freq, cutoff_freq, transition_width, taps, len);
}

/* Host evaluation of the table arctangent (same table and operation order as csrc/device_math.cuh). */
float fast_atan2f(float y, float x) {
    const float *table = sdrm_host_atan_table();
    const float y_abs = fabsf(y);
    const float x_abs = fabsf(x);
    if (!((y_abs > 0.0f) || (x_abs > 0.0f))) {
        return 0.0F;
    }
    const float z = y_abs < x_abs ? y_abs / x_abs : x_abs / y_abs;
    float base;
    if ((double) z < 0.003921569) {
        base = z;
    } else {
        float alpha = z * 255.0f;
        const int index = ((int) alpha) & 0xff;
        alpha -= (float) index;
        base = table[index] + (table[index + 1] - table[index]) * alpha;
    }
    if (x_abs > y_abs) {
        if (x >= 0.0f) {
            return y >= 0.0f ? base : -base;
        }
        return y >= 0.0f ? 3.14159265358979323846F - base : base - 3.14159265358979323846F;
    }
    if (y >= 0.0f) {
        return x >= 0.0f ? 1.57079632679489661923F - base : 1.57079632679489661923F + base;
    }
    return x >= 0.0f ? -1.57079632679489661923F + base : -1.57079632679489661923F - base;
}

/* ------------------------------------------------------------------------------------------------ fir_filter / lpf */

/* One streaming decimating FIR over one stream (reference fir_filter.c:35-159): shared by fir_filter_* and lpf_*. */
struct fir_core {
    uint8_t decimation;
    size_t num_bytes;
    size_t max_len;
    int n_taps;
    int hist_len;
    int phase;
    int cur;
    void *d_taps;
    float *h_taps; /* host copy of the (h, h) pairs */
    void *d_hist[2];
    void *d_in;  /* float2 [max_len + 2] */
    void *d_out; /* float2 [max_len + 2] */
    float *h_pack; /* real input widened to (x, 0) pairs / complex output staging */
    void *output;
    cudaStream_t stream;
};

static void fir_core_free(struct fir_core *f) {
    if (f->stream != NULL) {
        cudaStreamSynchronize(f->stream);
        cudaStreamDestroy(f->stream);
    }
    cudaFree(f->d_taps);
    free(f->h_taps);
    cudaFree(f->d_hist[0]);
    cudaFree(f->d_hist[1]);
    cudaFree(f->d_in);
    cudaFree(f->d_out);
    free(f->h_pack);
    free(f->output);
    memset(f, 0, sizeof(*f));
}

/* taps in design order (h[0] multiplies the newest sample), not consumed */
static int fir_core_init(struct fir_core *f, uint8_t decimation, const float *taps, size_t taps_len, size_t max_len,
                         size_t num_bytes) {
    if (decimation == 0 || taps_len == 0 || (num_bytes != 8 && num_bytes != 4)) {
        return -1;
    }
    f->decimation = decimation;
    f->num_bytes = num_bytes;
    f->max_len = max_len;
    f->n_taps = (int) taps_len;
    f->hist_len = (int) sdrm_round_up(taps_len - 1, 2);
    if (f->hist_len == 0) {
        f->hist_len = 2;
    }
    int code = sdrm_upload_taps_dup(taps, taps_len, &f->d_taps);
    f->h_taps = sdrm_host_taps_dup(taps, taps_len);
    const size_t cap = sdrm_round_up(max_len, 2) + 2;
    for (int i = 0; i < 2 && code == 0; i++) {
        code = sdrm_dev_zalloc(&f->d_hist[i], (size_t) f->hist_len * 8);
    }
    if (code == 0) code = sdrm_dev_zalloc(&f->d_in, cap * 8);
    if (code == 0) code = sdrm_dev_zalloc(&f->d_out, cap * 8);
    if (code == 0) code = sdrm_cuda_code(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking), "stream");
    if (code == 0) {
        f->h_pack = malloc(cap * 8);
        f->output = malloc(cap * 8);
        if (f->h_pack == NULL || f->output == NULL) {
            code = -ENOMEM;
        }
    }
    return code;
}

static void fir_core_process(struct fir_core *f, const void *input, size_t input_len, void **output, size_t *output_len) {
    *output = NULL;
    *output_len = 0;
    if (input_len > f->max_len) {
        log_oversize(input_len, f->max_len);
        return;
    }
    const void *src = input;
    if (f->num_bytes == 4) {
        /* one real stream rides in the first lane of a channel pair */
        const float *x = (const float *) input;
        for (size_t i = 0; i < input_len; i++) {
            f->h_pack[2 * i] = x[i];
            f->h_pack[2 * i + 1] = 0.0f;
        }
        src = f->h_pack;
    }
    if (input_len > 0 && cudaMemcpyAsync(f->d_in, src, input_len * 8, cudaMemcpyHostToDevice, f->stream) != cudaSuccess) {
        return;
    }
    const int dec = f->decimation;
    const int n_in = (int) input_len;
    const int n_out = n_in > f->phase ? (n_in - f->phase + dec - 1) / dec : 0;
    sdrm_fir_args a;
    memset(&a, 0, sizeof(a));
    a.in = f->d_in;
    a.in_stride = sdrm_round_up(f->max_len, 2) + 2;
    a.hist = f->d_hist[f->cur];
    a.hist_len = f->hist_len;
    a.taps_dup = f->d_taps;
    a.h_taps_dup = f->h_taps;
    a.n_taps = f->n_taps;
    a.decimation = dec;
    a.phase = f->phase;
    a.n_in = n_in;
    a.n_out = n_out;
    a.rows = 1;
    a.out_mode = SDRM_FIR_OUT_ROWS;
    a.out = f->d_out;
    a.out_stride = a.in_stride;
    if (sdrm_launch_code(sdrm_cu_fir(&a, f->stream), "fir") != 0) {
        return;
    }
    if (sdrm_launch_code(sdrm_cu_hist_update(f->d_in, a.in_stride, f->d_hist[f->cur], f->d_hist[f->cur ^ 1], f->hist_len, n_in, 1,
                                             f->stream),
                         "fir history") != 0) {
        return;
    }
    f->cur ^= 1;
    f->phase = f->phase + n_out * dec - n_in;
    void *dst = f->num_bytes == 8 ? f->output : (void *) f->h_pack;
    if (n_out > 0 && cudaMemcpyAsync(dst, f->d_out, (size_t) n_out * 8, cudaMemcpyDeviceToHost, f->stream) != cudaSuccess) {
        return;
    }
    if (sdrm_cuda_code(cudaStreamSynchronize(f->stream), "fir_process") != 0) {
        return;
    }
    if (f->num_bytes == 4) {
        float *y = (float *) f->output;
        for (int i = 0; i < n_out; i++) {
            y[i] = f->h_pack[2 * i];
        }
    }
    *output = f->output;
    *output_len = (size_t) n_out;
}

/* fir_filter: reference src/dsp/fir_filter.h:29-35. The struct is opaque here (the reference's fields describe its
 * host working buffers; nothing outside src/dsp reads them). */
/* fir_filter: reference src/dsp/fir_filter.h:9-35. The public struct (the reference's field layout) is the head of the
 * handle; the device-side state follows it. */
struct fir_filter_impl {
    struct fir_filter_t pub; /* first: a fir_filter* is the address of this member */
    struct fir_core core;
    float *reversed; /* taps_len floats, reversed (fir_filter.c:27): pub.taps[0], and fir_filter_process_float_single */
};

int fir_filter_create(uint8_t decimation, float *taps, size_t taps_len, size_t max_input_buffer_length, size_t num_bytes,
                      fir_filter **filter) {
    if (taps == NULL) {
        return -1;
    }
    struct fir_filter_impl *f = calloc(1, sizeof(*f));
    if (f == NULL) {
        return -ENOMEM;
    }
    /* from here on the filter owns the taps, also when create fails (fir_filter.c:58 followed by fir_filter_destroy) */
    f->pub.original_taps = taps;
    f->pub.taps_len = taps_len;
    f->pub.alignment = 16;
    /* the reference rejects an unparsable VOLK_ALIGNMENT override here (fir_filter.c:44-51); the value has no other use */
    const char *alignment = getenv("VOLK_ALIGNMENT");
    if (alignment != NULL) {
        f->pub.alignment = (size_t) strtol(alignment, NULL, 10);
        if (f->pub.alignment == 0) {
            SDRM_LOG_ERROR("invalid VOLK_ALIGNMENT specified: %s", alignment);
            fir_filter_destroy(&f->pub);
            return -1;
        }
    }
    int code = fir_core_init(&f->core, decimation, taps, taps_len, max_input_buffer_length, num_bytes);
    if (code == 0) {
        f->reversed = malloc(sizeof(float) * (taps_len == 0 ? 1 : taps_len));
        if (f->reversed == NULL) {
            code = -ENOMEM;
        }
    }
    if (code != 0) {
        fir_filter_destroy(&f->pub);
        return code;
    }
    for (size_t i = 0; i < taps_len; i++) {
        f->reversed[i] = taps[taps_len - 1 - i];
    }
    f->pub.decimation = decimation;
    f->pub.taps = &f->reversed;
    f->pub.aligned_taps_len = 1;
    f->pub.history_offset = taps_len - 1;
    f->pub.working_len_total = max_input_buffer_length + f->pub.history_offset;
    f->pub.max_input_buffer_length = max_input_buffer_length;
    f->pub.output = f->core.output;
    f->pub.output_len = max_input_buffer_length / decimation + 1;
    f->pub.num_bytes = num_bytes;
    *filter = &f->pub;
    return 0;
}

void fir_filter_process(const void *input, size_t input_len, void **output, size_t *output_len, fir_filter *filter) {
    fir_core_process(&((struct fir_filter_impl *) filter)->core, input, input_len, output, output_len);
}

/* One output from taps_len host floats, no state (fir_filter.c:116-121). The reference rounds the pointer down to its
 * 16-byte alignment and runs the dot product over (input - k .. input + taps_len) with k leading zero taps; the zeros
 * only matter when a sample in front of `input` is not finite, and that case is kept. */
float fir_filter_process_float_single(const float *input, fir_filter *filter) {
    const struct fir_filter_impl *f = (const struct fir_filter_impl *) filter;
    const size_t lead = ((size_t) input & 15u) / sizeof(float);
    const float *p = input - lead;
    float acc = 0.0f;
    for (size_t i = 0; i < lead; i++) {
        acc += p[i] * 0.0f;
    }
    for (size_t i = 0; i < f->pub.taps_len; i++) {
        acc += input[i] * f->reversed[i];
    }
    return acc;
}

void fir_filter_destroy(fir_filter *filter) {
    if (filter == NULL) {
        return;
    }
    struct fir_filter_impl *f = (struct fir_filter_impl *) filter;
    fir_core_free(&f->core);
    free(f->pub.original_taps);
    free(f->reversed);
    free(f);
}

struct lpf_t {
    struct fir_core core;
};

int lpf_create(uint8_t decimation, uint64_t sampling_freq, uint64_t cutoff_freq, uint32_t transition_width,
               size_t max_input_buffer_length, size_t num_bytes, lpf **filter) {
    if (decimation == 0 || (num_bytes != 8 && num_bytes != 4)) {
        return -1;
    }
    struct lpf_t *f = calloc(1, sizeof(*f));
    if (f == NULL) {
        return -ENOMEM;
    }
    float *taps = NULL;
    size_t taps_len = 0;
    int code = sdrm_design_low_pass(1.0F, sampling_freq, cutoff_freq, transition_width, &taps, &taps_len);
    if (code == 0) {
        code = fir_core_init(&f->core, decimation, taps, taps_len, max_input_buffer_length, num_bytes);
        free(taps);
    }
    if (code != 0) {
        lpf_destroy(f);
        return code;
    }
    *filter = f;
    return 0;
}

void lpf_process(const void *input, size_t input_len, void **output, size_t *output_len, lpf *f) {
    fir_core_process(&f->core, input, input_len, output, output_len);
}

void lpf_destroy(lpf *f) {
    if (f == NULL) {
        return;
    }
    fir_core_free(&f->core);
    free(f);
}

/* ------------------------------------------------------------------------------------------------ quadrature demod */

struct quadrature_demod_t {
    float gain;
    uint32_t max_len;
    float *d_atan;
    void *d_in;
    void *d_prev;
    float *d_out;
    float *output;
    cudaStream_t stream;
};

int quadrature_demod_create(float gain, uint32_t max_input_buffer_length, quadrature_demod **demod) {
    struct quadrature_demod_t *q = calloc(1, sizeof(*q));
    if (q == NULL) {
        return -ENOMEM;
    }
    q->gain = gain;
    q->max_len = max_input_buffer_length;
    int code = sdrm_upload_atan_table(&q->d_atan);
    if (code == 0) code = sdrm_dev_zalloc(&q->d_in, ((size_t) max_input_buffer_length + 1) * 8);
    if (code == 0) code = sdrm_dev_zalloc(&q->d_prev, 8);
    if (code == 0) code = sdrm_dev_zalloc((void **) &q->d_out, ((size_t) max_input_buffer_length + 1) * 4);
    if (code == 0) code = sdrm_cuda_code(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking), "stream");
    if (code == 0) {
        q->output = malloc(((size_t) max_input_buffer_length + 1) * sizeof(float));
        if (q->output == NULL) {
            code = -ENOMEM;
        }
    }
    if (code != 0) {
        quadrature_demod_destroy(q);
        return code;
    }
    *demod = q;
    return 0;
}

void quadrature_demod_process(float complex *input, size_t input_len, float **output, size_t *output_len, quadrature_demod *q) {
    *output = NULL;
    *output_len = 0;
    if (input_len > q->max_len) {
        log_oversize(input_len, q->max_len);
        return;
    }
    if (input_len > 0) {
        if (cudaMemcpyAsync(q->d_in, input, input_len * 8, cudaMemcpyHostToDevice, q->stream) != cudaSuccess) {
            return;
        }
        if (sdrm_launch_code(sdrm_cu_quad_demod(q->d_in, q->max_len + 1, q->d_prev, q->gain, q->d_atan, q->d_out, q->max_len + 1,
                                                (int) input_len, 1, q->stream),
                             "quadrature demod") != 0) {
            return;
        }
        if (cudaMemcpyAsync(q->output, q->d_out, input_len * 4, cudaMemcpyDeviceToHost, q->stream) != cudaSuccess) {
            return;
        }
        if (sdrm_cuda_code(cudaStreamSynchronize(q->stream), "quadrature_demod_process") != 0) {
            return;
        }
    }
    *output = q->output;
    *output_len = input_len;
}

void quadrature_demod_destroy(quadrature_demod *q) {
    if (q == NULL) {
        return;
    }
    if (q->stream != NULL) {
        cudaStreamSynchronize(q->stream);
        cudaStreamDestroy(q->stream);
    }
    cudaFree(q->d_atan);
    cudaFree(q->d_in);
    cudaFree(q->d_prev);
    cudaFree(q->d_out);
    free(q->output);
    free(q);
}

/* ------------------------------------------------------------------------------------------------ dc blocker */

/* The reference's handle has no maximum length (it works in place on the caller's buffer), so the device staging
 * ring grows on demand. */
struct dc_blocker_t {
    int length;
    float *d_delay;
    float *d_sums;
    float *d_ring;
    uint32_t ring_rows;
    long long head;
    int pos_l;
    int pos_x;
    float *h_pack;
    size_t h_cap;
    cudaStream_t stream;
};

int dc_blocker_create(int length, dc_blocker **blocker) {
    if (length < 2) {
        return -1;
    }
    struct dc_blocker_t *d = calloc(1, sizeof(*d));
    if (d == NULL) {
        return -ENOMEM;
    }
    d->length = length;
    int code = sdrm_dev_zalloc((void **) &d->d_delay, (size_t) (6 * length - 2) * 2 * sizeof(float));
    if (code == 0) code = sdrm_dev_zalloc((void **) &d->d_sums, 4 * 2 * sizeof(float));
    if (code == 0) code = sdrm_cuda_code(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking), "stream");
    if (code != 0) {
        dc_blocker_destroy(d);
        return code;
    }
    *blocker = d;
    return 0;
}

static int dc_reserve(struct dc_blocker_t *d, size_t n) {
    if (n <= d->ring_rows) {
        return 0;
    }
    const uint32_t rows = sdrm_next_pow2(n < 1024 ? 1024 : n);
    float *ring = NULL;
    int code = sdrm_dev_zalloc((void **) &ring, (size_t) rows * 2 * sizeof(float));
    if (code != 0) {
        return code;
    }
    float *pack = realloc(d->h_pack, (size_t) rows * 2 * sizeof(float));
    if (pack == NULL) {
        cudaFree(ring);
        return -ENOMEM;
    }
    cudaFree(d->d_ring);
    d->d_ring = ring;
    d->ring_rows = rows;
    d->h_pack = pack;
    d->head = 0; /* the ring carries no state between calls (all of it is in the delay lines) */
    return 0;
}

void dc_blocker_process(float *input, size_t input_len, float **output, size_t *output_len, dc_blocker *d) {
    *output = input; /* in place, as the reference (dc_blocker.c:115-117) */
    *output_len = input_len;
    if (input_len == 0) {
        return;
    }
    if (dc_reserve(d, input_len) != 0) {
        *output = NULL;
        *output_len = 0;
        return;
    }
    for (size_t i = 0; i < input_len; i++) {
        d->h_pack[2 * i] = input[i];
        d->h_pack[2 * i + 1] = 0.0f;
    }
    d->head = 0;
    int ok = cudaMemcpyAsync(d->d_ring, d->h_pack, input_len * 8, cudaMemcpyHostToDevice, d->stream) == cudaSuccess;
    ok = ok && sdrm_launch_code(sdrm_cu_dc_blocker(d->d_ring, 2, (int) d->ring_rows, 0, (int) input_len, 2, d->length, d->d_delay,
                                                   d->d_sums, d->pos_l, d->pos_x, 0, d->stream),
                                "dc blocker") == 0;
    ok = ok && cudaMemcpyAsync(d->h_pack, d->d_ring, input_len * 8, cudaMemcpyDeviceToHost, d->stream) == cudaSuccess;
    ok = ok && sdrm_cuda_code(cudaStreamSynchronize(d->stream), "dc_blocker_process") == 0;
    if (!ok) {
        *output = NULL;
        *output_len = 0;
        return;
    }
    d->pos_l = (int) (((long long) d->pos_l + (long long) input_len) % d->length);
    d->pos_x = (int) (((long long) d->pos_x + (long long) input_len) % (2 * d->length - 2));
    for (size_t i = 0; i < input_len; i++) {
        input[i] = d->h_pack[2 * i];
    }
}

void dc_blocker_destroy(dc_blocker *d) {
    if (d == NULL) {
        return;
    }
    if (d->stream != NULL) {
        cudaStreamSynchronize(d->stream);
        cudaStreamDestroy(d->stream);
    }
    cudaFree(d->d_delay);
    cudaFree(d->d_sums);
    cudaFree(d->d_ring);
    free(d->h_pack);
    free(d);
}

/* ------------------------------------------------------------------------------------------------ clock recovery */

struct clock_mm_t {
    size_t max_len;
    float omega_mid;
    float omega_lim;
    float gain_omega;
    float gain_mu;
    float *d_mmse;
    float *d_ring;
    uint32_t ring_rows;
    long long head;
    sdrm_clock_state *d_state;
    float *d_soft;
    uint32_t *d_out_len;
    int *d_error;
    float *h_pack;
    float *output;
    int history; /* host mirror of the carried sample count (the reference's history_offset) */
    cudaStream_t stream;
};

int clock_mm_create(float omega, float gain_omega, float mu, float gain_mu, float omega_relative_limit, size_t output_len,
                    clock_mm **clock) {
    struct clock_mm_t *c = calloc(1, sizeof(*c));
    if (c == NULL) {
        return -ENOMEM;
    }
    c->max_len = output_len;
    c->omega_mid = omega;
    c->omega_lim = omega * omega_relative_limit;
    c->gain_omega = gain_omega;
    c->gain_mu = gain_mu;
    /* the reference keeps at most output_len + 8 samples; one ring of twice that never overwrites carried history */
    c->ring_rows = sdrm_next_pow2(2 * (uint64_t) output_len + 64);
    int code = sdrm_upload_mmse_table(&c->d_mmse);
    if (code == 0) code = sdrm_dev_zalloc((void **) &c->d_ring, (size_t) c->ring_rows * 2 * sizeof(float));
    if (code == 0) code = sdrm_dev_zalloc((void **) &c->d_state, 2 * sizeof(sdrm_clock_state));
    if (code == 0) code = sdrm_dev_zalloc((void **) &c->d_soft, (output_len + 8) * sizeof(float));
    if (code == 0) code = sdrm_dev_zalloc((void **) &c->d_out_len, 2 * sizeof(uint32_t));
    if (code == 0) code = sdrm_dev_zalloc((void **) &c->d_error, sizeof(int));
    if (code == 0) code = sdrm_cuda_code(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "stream");
    if (code == 0) {
        sdrm_clock_state init = {mu, omega, 0.0f, 0};
        code = sdrm_cuda_code(cudaMemcpy(c->d_state, &init, sizeof(init), cudaMemcpyHostToDevice), "clock state");
    }
    if (code == 0) {
        c->h_pack = malloc((output_len + 8) * 2 * sizeof(float));
        c->output = malloc((output_len + 8) * sizeof(float));
        if (c->h_pack == NULL || c->output == NULL) {
            code = -ENOMEM;
        }
    }
    if (code != 0) {
        clock_mm_destroy(c);
        return code;
    }
    *clock = c;
    return 0;
}

void clock_mm_process(const float *input, size_t input_len, float **output, size_t *output_len, clock_mm *c) {
    *output = NULL;
    *output_len = 0;
    if (input_len > c->max_len) {
        log_oversize(input_len, c->max_len);
        return;
    }
    /* new samples go to ring rows [head, head + n), possibly in two pieces around the wrap */
    for (size_t i = 0; i < input_len; i++) {
        c->h_pack[2 * i] = input[i];
        c->h_pack[2 * i + 1] = 0.0f;
    }
    const size_t first_row = (size_t) (c->head & (c->ring_rows - 1));
    const size_t first = input_len < c->ring_rows - first_row ? input_len : c->ring_rows - first_row;
    int ok = 1;
    if (first > 0) {
        ok = cudaMemcpyAsync(c->d_ring + 2 * first_row, c->h_pack, first * 8, cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
    }
    if (ok && input_len > first) {
        ok = cudaMemcpyAsync(c->d_ring, c->h_pack + 2 * first, (input_len - first) * 8, cudaMemcpyHostToDevice, c->stream) ==
             cudaSuccess;
    }
    sdrm_clock_args a;
    memset(&a, 0, sizeof(a));
    a.ring = c->d_ring;
    a.tc_stride = 2;
    a.ring_rows = (int) c->ring_rows;
    a.head = c->head;
    a.n_rows = (int) input_len;
    a.n_ch = 1;
    a.max_history = (int) c->max_len + 8;
    a.omega_mid = c->omega_mid;
    a.omega_lim = c->omega_lim;
    a.gain_omega = c->gain_omega;
    a.gain_mu = c->gain_mu;
    a.mmse_taps = c->d_mmse;
    a.state = c->d_state;
    a.soft_out = c->d_soft;
    a.hard_out = NULL;
    a.out_stride = c->max_len + 8;
    a.out_len = c->d_out_len;
    a.max_out = (int) c->max_len;
    a.error_flag = c->d_error;
    ok = ok && sdrm_launch_code(sdrm_cu_clock_mm(&a, c->stream), "clock recovery") == 0;
    uint32_t produced = 0;
    sdrm_clock_state after;
    memset(&after, 0, sizeof(after));
    ok = ok && cudaMemcpyAsync(&produced, c->d_out_len, sizeof(produced), cudaMemcpyDeviceToHost, c->stream) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(&after, c->d_state, sizeof(after), cudaMemcpyDeviceToHost, c->stream) == cudaSuccess;
    ok = ok && sdrm_cuda_code(cudaStreamSynchronize(c->stream), "clock_mm_process") == 0;
    if (!ok) {
        return;
    }
    c->head += (long long) input_len;
    /* the reference returns NULL / 0 while history + input is still shorter than the interpolator's 8 taps
     * (clock_recovery_mm.c:94-99) and its output buffer, with any length including 0, otherwise */
    const size_t working_len = (size_t) c->history + input_len;
    c->history = after.history;
    if (working_len < 8) {
        return;
    }
    if (produced > 0 &&
        cudaMemcpy(c->output, c->d_soft, (size_t) produced * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) {
        return;
    }
    *output = c->output;
    *output_len = produced;
}

void clock_mm_destroy(clock_mm *c) {
    if (c == NULL) {
        return;
    }
    if (c->stream != NULL) {
        cudaStreamSynchronize(c->stream);
        cudaStreamDestroy(c->stream);
    }
    cudaFree(c->d_mmse);
    cudaFree(c->d_ring);
    cudaFree(c->d_state);
    cudaFree(c->d_soft);
    cudaFree(c->d_out_len);
    cudaFree(c->d_error);
    free(c->h_pack);
    free(c->output);
    free(c);
}

/* ------------------------------------------------------------------------------------------------ mmse interpolator */

#include "../../include/sdrm/mmse_fir_interpolator.h"

struct mmse_fir_interpolator_t {
    int taps_len;
    int steps;
};

int mmse_fir_interpolator_create(mmse_fir_interpolator **interp) {
    struct mmse_fir_interpolator_t *result = malloc(sizeof(*result));
    if (result == NULL) {
        return -ENOMEM;
    }
    result->taps_len = 8;
    result->steps = 128;
    *interp = result;
    return 0;
}

/* One output of the bank selected by rint(mu * 128) (mmse_fir_interpolator.c:188-191). The reference evaluates it with
 * an aligned dot product that starts at the 16-byte boundary at or below `input` and meets zero taps first
 * (fir_filter.c:116-121); those leading terms are reproduced because they matter for non-finite neighbours. */
float mmse_fir_interpolator_process(const float *input, float mu, mmse_fir_interpolator *interp) {
    const int imu = (int) rint(mu * interp->steps);
    const float *row = sdrm_host_mmse_table() + (size_t) imu * 8;
    const size_t lead = ((size_t) input & 15) / sizeof(float);
    float acc = 0.0f;
    for (size_t k = lead; k > 0; k--) {
        acc += *(input - k) * 0.0f;
    }
    for (int j = 0; j < 8; j++) {
        acc += input[j] * row[7 - j];
    }
    return acc;
}

int mmse_fir_interpolator_taps(mmse_fir_interpolator *interp) { return interp->taps_len; }

void mmse_fir_interpolator_destroy(mmse_fir_interpolator *interp) { free(interp); }
