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
 */
#ifndef SDRM_ORACLE_VOLK_SHIM_H
#define SDRM_ORACLE_VOLK_SHIM_H

#include <complex.h>
#include <math.h>
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

static inline void volk_32f_x2_dot_prod_32f_u(float *result, const float *input, const float *taps, unsigned int num_points) {
    float acc = 0.0f;
    for (unsigned int i = 0; i < num_points; i++) {
        SDRM_SHIM_MAC(acc, input[i], taps[i]);
    }
    *result = acc;
}

static inline void volk_32f_x2_dot_prod_32f_a(float *result, const float *input, const float *taps, unsigned int num_points) {
    volk_32f_x2_dot_prod_32f_u(result, input, taps, num_points);
}

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

#endif
