/*
 * TEST INFRASTRUCTURE ONLY — never linked into the product library.
 *
 * Minimal stand-in for <volk/volk.h> so that the reference's own src/dsp sources
 * (compiled IN PLACE from /root/reference, see oracle/Makefile) build without libvolk,
 * which is an un-vendored third-party dependency of the reference
 * (reference CMakeLists.txt:49-52, README.md:41 "volk 2.x"; no version pin).
 *
 * It restates the published semantics of VOLK's *generic* (scalar) kernels — the ones
 * the reference's tests pin results at (test/resources/run_tests.sh:8 VOLK_GENERIC=1,
 * test/test_fsk_demod.c:130):
 *   - dot products: one accumulator per component, started at 0.0f, products added
 *     sequentially from index 0; multiply and add are separately rounded,
 *   - complex multiplies: C99 complex arithmetic,
 *   - float -> int8: scale, saturate to [-128,127], round-half-even (rintf).
 * Call sites in the reference: src/dsp/fir_filter.c:102,119,132,
 * src/dsp/quadrature_demod.c:65, src/dsp/sig_source.c:71, src/dsp/fsk_demod.c:106.
 *
 * With -DSDRM_SHIM_FMA the three dot products accumulate with fmaf() instead (same order).
 * That models the reference on FMA-contracting platforms and is the checker for the
 * library's optional "fast" arithmetic mode; it is NOT the parity-defining oracle.
 *
 * With -DSDRM_SHIM_SIMD the dot products keep 16 lane-partial sums that are added up at the
 * end, the structure of VOLK's SIMD kernels (e.g. volk_32fc_32f_dot_prod_32fc_a_avx: products
 * accumulated per vector lane, horizontal sum last), so that the compiler vectorises them.
 * A different summation order, hence different bits: used ONLY for the "tuned" CPU throughput
 * figure that bench.py reports beside the strict one, never as a checker.
 */
#ifndef SDRM_ORACLE_VOLK_SHIM_H
#define SDRM_ORACLE_VOLK_SHIM_H

#include <complex.h>
#include <math.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

typedef float complex lv_32fc_t;

static inline size_t volk_get_alignment(void) { return 16; }

static inline void *volk_malloc(size_t size, size_t alignment) {
    void *p = NULL;
    if (alignment < sizeof(void *)) {
        alignment = sizeof(void *);
    }
    if (posix_memalign(&p, alignment, size == 0 ? alignment : size) != 0) {
        return NULL;
    }
    return p;
}

static inline void volk_free(void *p) { free(p); }

#ifdef SDRM_SHIM_FMA
#define SDRM_SHIM_MAC(acc, a, b) ((acc) = fmaf((a), (b), (acc)))
#else
#define SDRM_SHIM_MAC(acc, a, b) ((acc) += (a) * (b))
#endif

#ifdef SDRM_SHIM_SIMD
static inline void volk_32f_x2_dot_prod_32f_u(float *result, const float *input, const float *taps, unsigned int num_points) {
    float lane[16] = {0};
    unsigned int i = 0;
    for (; i + 16 <= num_points; i += 16) {
        for (int k = 0; k < 16; k++) {
            lane[k] += input[i + k] * taps[i + k];
        }
    }
    float acc = 0.0f;
    for (; i < num_points; i++) {
        acc += input[i] * taps[i];
    }
    for (int k = 0; k < 16; k++) {
        acc += lane[k];
    }
    *result = acc;
}
#else
static inline void volk_32f_x2_dot_prod_32f_u(float *result, const float *input, const float *taps, unsigned int num_points) {
    float acc = 0.0f;
    for (unsigned int i = 0; i < num_points; i++) {
        SDRM_SHIM_MAC(acc, input[i], taps[i]);
    }
    *result = acc;
}
#endif

static inline void volk_32f_x2_dot_prod_32f_a(float *result, const float *input, const float *taps, unsigned int num_points) {
    volk_32f_x2_dot_prod_32f_u(result, input, taps, num_points);
}

#ifdef SDRM_SHIM_SIMD
typedef float sdrm_v8f __attribute__((vector_size(32), aligned(4)));
typedef int sdrm_v8i __attribute__((vector_size(32)));
static inline void volk_32fc_32f_dot_prod_32fc_u(lv_32fc_t *result, const lv_32fc_t *input, const float *taps, unsigned int num_points) {
    const float *in = (const float *) input;
    sdrm_v8f acc0 = {0};
    sdrm_v8f acc1 = {0}; /* lanes: (re, im) of 4 consecutive samples each */
    const sdrm_v8i lo = {0, 0, 1, 1, 2, 2, 3, 3};
    const sdrm_v8i hi = {4, 4, 5, 5, 6, 6, 7, 7};
    unsigned int i = 0;
    for (; i + 8 <= num_points; i += 8) {
        const sdrm_v8f t = *(const sdrm_v8f *) (taps + i);
        acc0 += *(const sdrm_v8f *) (in + 2 * i) * __builtin_shuffle(t, lo);
        acc1 += *(const sdrm_v8f *) (in + 2 * i + 8) * __builtin_shuffle(t, hi);
    }
    float re = 0.0f;
    float im = 0.0f;
    for (; i < num_points; i++) {
        re += in[2 * i] * taps[i];
        im += in[2 * i + 1] * taps[i];
    }
    acc0 += acc1;
    for (int k = 0; k < 8; k += 2) {
        re += acc0[k];
        im += acc0[k + 1];
    }
    float *out = (float *) result;
    out[0] = re;
    out[1] = im;
}
#else
static inline void volk_32fc_32f_dot_prod_32fc_u(lv_32fc_t *result, const lv_32fc_t *input, const float *taps, unsigned int num_points) {
    const float *in = (const float *) input;
    float re = 0.0f;
    float im = 0.0f;
    for (unsigned int i = 0; i < num_points; i++) {
        SDRM_SHIM_MAC(re, in[2 * i], taps[i]);
        SDRM_SHIM_MAC(im, in[2 * i + 1], taps[i]);
    }
    float *out = (float *) result;
    out[0] = re;
    out[1] = im;
}
#endif

static inline void volk_32fc_x2_multiply_conjugate_32fc(lv_32fc_t *c, const lv_32fc_t *a, const lv_32fc_t *b, unsigned int num_points) {
    for (unsigned int i = 0; i < num_points; i++) {
        c[i] = a[i] * conjf(b[i]);
    }
}

static inline void volk_32fc_x2_multiply_32fc(lv_32fc_t *c, const lv_32fc_t *a, const lv_32fc_t *b, unsigned int num_points) {
    for (unsigned int i = 0; i < num_points; i++) {
        c[i] = a[i] * b[i];
    }
}

static inline void volk_32f_s32f_convert_8i(int8_t *out, const float *in, const float scalar, unsigned int num_points) {
    for (unsigned int i = 0; i < num_points; i++) {
        float r = in[i] * scalar;
        if (r > 127.0f) {
            out[i] = 127;
        } else if (r < -128.0f) {
            out[i] = -128;
        } else {
            out[i] = (int8_t) rintf(r);
        }
    }
}

/* SDR sample formats of the PlutoSDR plugin (src/sdr/plutosdr.c:83,129), VOLK generic kernels:
 * 16i -> 32f: (float) in / scalar;  32f -> 16i: in * scalar, saturated to [-32768, 32767], rintf */
static inline void volk_16i_s32f_convert_32f(float *out, const int16_t *in, const float scalar, unsigned int num_points) {
    for (unsigned int i = 0; i < num_points; i++) {
        out[i] = (float) in[i] / scalar;
    }
}

static inline void volk_32f_s32f_convert_16i(int16_t *out, const float *in, const float scalar, unsigned int num_points) {
    for (unsigned int i = 0; i < num_points; i++) {
        float r = in[i] * scalar;
        if (r > 32767.0f) {
            out[i] = 32767;
        } else if (r < -32768.0f) {
            out[i] = -32768;
        } else {
            out[i] = (int16_t) rintf(r);
        }
    }
}

#endif
