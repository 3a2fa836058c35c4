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
//
// CUDA's sincos() is accurate enough (the sweep shows it) but costs ~100 instructions per phase in these kernels, half of
// them moves that rebuild its 64-bit polynomial coefficients from immediates at every use (ncu: the kernels are bound by issue
// slots, the FP64 pipe is 27 % busy). The phase is a float that the wrap keeps inside [-2 pi, 2 pi]; for |p| <= 7 this is the
// textbook evaluation with the coefficients in constant memory, where DFMA reads them as operands:
//   k = rint(p * 2 / pi) in float (|k| <= 5);  y = (p - k * P1) - k * P2, two-constant Cody-Waite (the first product and
//   difference are exact: P1 holds the first 33 bits of pi / 2);  Taylor polynomials of sin y and cos y on |y| <= pi / 4 + eps
//   to y^17 and y^16 (remainders below 2^-53 relative);  quadrant k mod 4.
// Its double differs from libm's in the last place now and then, like sincos() does; after rounding to float there is no
// difference for any phase of the range — established by the sweep, not by this comment. Anything else (the handles accept any
// float) goes to sincos().
static __constant__ double sdrm_sin_taylor[8] = {-0.16666666666666666,    0.008333333333333333,   -0.0001984126984126984, 2.7557319223985893e-06,
                                                 -2.505210838544172e-08,  1.6059043836821613e-10, -7.647163731819816e-13, 2.8114572543455206e-15};
static __constant__ double sdrm_cos_taylor[8] = {-0.5,                    0.041666666666666664,   -0.001388888888888889,  2.48015873015873e-05,
                                                 -2.755731922398589e-07,  2.08767569878681e-09,   -1.1470745597729725e-11, 4.779477332387385e-14};
static __constant__ double sdrm_pio2_parts[2] = {1.57079632673412561417e+00, 6.07710050650619224932e-11};

__device__ __forceinline__ void sdrm_phase_sincos(float p, double *sin_out, double *cos_out) {
#ifndef SDRM_TRIG_LIBDEVICE  // A/B builds: define it to send every phase through sincos()
    if (fabsf(p) <= 7.0f && p != 0.0f) {  // a zero phase keeps its sign through sincos() (the reduction below would lose -0)
        const float kf = rintf(__fmul_rn(p, 0.636619772f));
        const double kd = (double) kf;
        double y = fma(kd, -sdrm_pio2_parts[0], (double) p);
        y = fma(kd, -sdrm_pio2_parts[1], y);
        const double z = y * y;
        double ps = sdrm_sin_taylor[7];
        double pc = sdrm_cos_taylor[7];
#pragma unroll
        for (int i = 6; i >= 0; i--) {
            ps = fma(ps, z, sdrm_sin_taylor[i]);
            pc = fma(pc, z, sdrm_cos_taylor[i]);
        }
        const double sin_y = fma(y * z, ps, y);
        const double cos_y = fma(z, pc, 1.0);
        const int q = (int) kf;
        const double s = (q & 1) ? cos_y : sin_y;
        const double c = (q & 1) ? sin_y : cos_y;
        *sin_out = (q & 2) ? -s : s;
        *cos_out = ((q + 1) & 2) ? -c : c;
        return;
    }
#endif
    sincos((double) p, sin_out, cos_out);
}

#endif
