// Scalar device functions shared by the kernels. Everything here mirrors a reference routine operation by
// operation; the translation unit is compiled with -fmad=false and the rounded intrinsics below never contract.
#ifndef SDRM_DEVICE_MATH_CUH
#define SDRM_DEVICE_MATH_CUH

#include <cuda_runtime.h>

// Table arctangent, reference src/math/fast_atan2f.c:87-157 (GNU Radio fast_atan2f): octant fold, one true
// division, 255-step table with linear interpolation, small-angle shortcut below TAN_MAP_RES.
__device__ __forceinline__ float sdrm_fast_atan2f(float y, float x, const float *table) {
    const float y_abs = fabsf(y);
    const float x_abs = fabsf(x);
    if (!((y_abs > 0.0f) || (x_abs > 0.0f))) {
        return 0.0f;
    }
    const float z = (y_abs < x_abs) ? __fdiv_rn(y_abs, x_abs) : __fdiv_rn(x_abs, y_abs);
    float base_angle;
    if ((double) z < 0.003921569) {  // the reference compares against a double literal
        base_angle = z;
    } else {
        float alpha = __fmul_rn(z, 255.0f);
        const int index = ((int) alpha) & 0xff;
        alpha = __fsub_rn(alpha, (float) index);
        const float t0 = table[index];
        const float t1 = table[index + 1];
        base_angle = __fadd_rn(t0, __fmul_rn(__fsub_rn(t1, t0), alpha));
    }
    float angle;
    if (x_abs > y_abs) {
        if (x >= 0.0f) {
            angle = (y >= 0.0f) ? base_angle : -base_angle;
        } else {
            const float pi = 3.14159265358979323846f;
            angle = (y >= 0.0f) ? __fsub_rn(pi, base_angle) : __fsub_rn(base_angle, pi);
        }
    } else {
        const float half_pi = 1.57079632679489661923f;
        if (y >= 0.0f) {
            angle = (x >= 0.0f) ? __fsub_rn(half_pi, base_angle) : __fadd_rn(half_pi, base_angle);
        } else {
            angle = (x >= 0.0f) ? __fadd_rn(-half_pi, base_angle) : __fsub_rn(-half_pi, base_angle);
        }
    }
    return angle;
}

// gain * fast_atan2f(t), t = cur * conj(prev) in C99 complex arithmetic
// (reference src/dsp/quadrature_demod.c:65-67, volk_32fc_x2_multiply_conjugate_32fc generic kernel).
__device__ __forceinline__ float sdrm_quad_demod_sample(float2 cur, float2 prev, float gain, const float *table) {
    const float re = __fadd_rn(__fmul_rn(cur.x, prev.x), __fmul_rn(cur.y, prev.y));
    const float im = __fsub_rn(__fmul_rn(cur.y, prev.x), __fmul_rn(cur.x, prev.y));
    return __fmul_rn(gain, sdrm_fast_atan2f(im, re, table));
}

// (float) cos((double) p), (float) sin((double) p): what the reference's frequency modulator and signal source store for a
// float phase (frequency_modulator.c:56, sig_source.c). One definition for every kernel that needs it, so that the sweep over
// all float phases (sdrm_cu_selftest_sincos, tests/test_gpu_sincos_sweep.py) checks the code that runs.
__device__ __forceinline__ void sdrm_phase_sincos(float p, double *sin_out, double *cos_out) { sincos((double) p, sin_out, cos_out); }

#endif
