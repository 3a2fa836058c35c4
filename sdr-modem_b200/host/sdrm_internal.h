/* Internal helpers of the C host layer. Not installed. */
#ifndef SDRM_INTERNAL_H
#define SDRM_INTERNAL_H

#include <cuda_runtime_api.h>
#include <errno.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#include "../csrc/sdrm_cuda.h"

/* stderr convention of the reference: "<3>" = syslog priority "error" for journald (e.g. src/dsp/fir_filter.c:148) */
#define SDRM_LOG_ERROR(...)                \
    do {                                   \
        fprintf(stderr, "<3>" __VA_ARGS__); \
        fputc('\n', stderr);               \
    } while (0)

/* CUDA failure -> reference-style negative code. There is deliberately no CPU fallback anywhere. */
static inline int sdrm_cuda_code(cudaError_t err, const char *what) {
    if (err == cudaSuccess) {
        return 0;
    }
    SDRM_LOG_ERROR("cuda failure in %s: %s", what, cudaGetErrorString(err));
    /* a failed runtime call also parks its code as the thread's "last error", where the cudaGetLastError() check behind the
     * next kernel launch of an unrelated handle would find it: it has been reported here, clear it */
    (void) cudaGetLastError();
    return err == cudaErrorMemoryAllocation ? -ENOMEM : -EIO;
}

#define SDRM_CUDA_TRY(call)                          \
    do {                                             \
        int sdrm_code_ = sdrm_cuda_code((call), #call); \
        if (sdrm_code_ != 0) {                       \
            return sdrm_code_;                       \
        }                                            \
    } while (0)

static inline int sdrm_launch_code(int code, const char *what) {
    if (code == 0) {
        return 0;
    }
    if (code <= -1000) {
        return sdrm_cuda_code((cudaError_t) (-(code + 1000)), what);
    }
    SDRM_LOG_ERROR("invalid launch arguments in %s (%d)", what, code);
    return -1;
}

/* Measurement aids of the batched demodulator (tools/probe_*.py, bench.py --debug-no-tail). They make the results invalid,
 * so they are not public flags: sdrm_fsk_demod_batch_create rejects unknown flag bits, and these are set through
 * sdrm_debug_set_measurement_aid, which no installed header declares. */
#define SDRM_AID_NO_CLOCK_LOOP 1u /* the tail runs, the clock loop emits nothing */
#define SDRM_AID_NO_TAIL 2u       /* filters only */
#define SDRM_AID_FETCH_BOUND_1 4u /* test aid, results stay valid: the fetch assumes one symbol per call, so that its second pass runs */
struct sdrm_fsk_demod_batch_t;
int sdrm_debug_set_measurement_aid(struct sdrm_fsk_demod_batch_t *batch, uint32_t mask);

/* cudaMalloc + zero fill */
int sdrm_dev_zalloc(void **p, size_t bytes);

/* Host tap design (taps.c) — double math, float taps, identical formulas and libm calls to the reference. */
int sdrm_design_low_pass(float gain, uint64_t sampling_freq, uint64_t cutoff_freq, uint32_t transition_width,
                         float **taps, size_t *len);
int sdrm_design_gaussian(double gain, double samples_per_symbol, double bt, size_t taps_len, float **taps);
int sdrm_convolve_full(const float *x, size_t x_len, const float *y, size_t y_len, float **out, size_t *out_len);

/* Uploads taps reversed and duplicated into float2 (h, h), padded with zeros to an even count + 2. */
int sdrm_upload_taps_dup(const float *taps, size_t len, void **d_taps);
/* The same (h, h) pairs, len float2, in host memory (malloc): short filters pass them to the kernel as parameters. */
float *sdrm_host_taps_dup(const float *taps, size_t len);

/* Markstein corrections needed so that the tail's division by `length` is an IEEE division (exhaustive check, ~20 ms). */
int sdrm_division_steps(int length);

/* Device copies of the constant tables (tables_data.h). */
int sdrm_upload_atan_table(float **d_table);
int sdrm_upload_mmse_table(float **d_table);
const float *sdrm_host_atan_table(void);
const float *sdrm_host_mmse_table(void);

/* handoff.c: true when take_buffer_for_processing would not block */
struct queue_t;
int sdrm_queue_has_data(struct queue_t *q);

static inline size_t sdrm_round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

static inline uint32_t sdrm_next_pow2(uint64_t v) {
    uint64_t p = 1;
    while (p < v) {
        p <<= 1;
    }
    return (uint32_t) p;
}

#endif
