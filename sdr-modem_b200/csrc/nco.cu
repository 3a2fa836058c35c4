// NCO / mixer for sm_100a: the reference's sig_source (src/dsp/sig_source.c:43-75), which doppler_process_rx/tx
// (src/dsp/doppler.c:180) and the TX offset mix (src/tcp_server.c:209) are built on.
//
//   out[i] = in[i] * amp * (cos p_i + j sin p_i),   p_{i+1} = wrap(p_i + w),   w = 2*pi*(float)f / fs   (all float)
//
// The phase is a float accumulator that is wrapped by +-2pi and never reset, so p_i is a serial recurrence whose
// rounding is part of the result: it is walked per channel (one lane per channel, nco_walk_kernel) and the phases are
// written out; the trigonometry (double precision, rounded to float, as the reference) and the complex multiply are
// then embarrassingly parallel over samples (nco_rotate_kernel). C99 complex product of the VOLK generic kernel:
// (ar*cr - ai*ci, ar*ci + ai*cr) with every product and sum rounded separately.
//
// Not fused into the FIR on purpose: the FIR is FP32-pipe bound, not HBM bound, so the 16 B/sample round
// trip costs nothing there, while rotating on load would redo the trigonometry for every tile's halo (for the 9325-tap
// filter of BASELINE config 3 the halo is 3.6x the tile).

#include <cuda_runtime.h>
#include <stdint.h>

#include "sdrm_cuda.h"

namespace {

constexpr float kTwoPi = 6.283185307179586476925286766559f;  // (float) (2 * M_PI), sig_source.c:11

// lane = channel; writes the phase used for sample i to phases[ch * stride + i] and carries the accumulator.
__global__ void nco_walk_kernel(const float *__restrict__ step, float *__restrict__ phase_state, float *__restrict__ phases,
                                size_t stride, int n, int n_ch) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= n_ch) {
        return;
    }
    const float w = step[ch];
    float p = phase_state[ch];
    float *out = phases + (size_t) ch * stride;
    for (int i = 0; i < n; i++) {
        out[i] = p;
        p = __fadd_rn(p, w);
        if (p < -kTwoPi) {
            p = __fadd_rn(p, kTwoPi);
        }
        if (p > kTwoPi) {
            p = __fsub_rn(p, kTwoPi);
        }
    }
    phase_state[ch] = p;
}

template <bool MULTIPLY>
__global__ void nco_rotate_kernel(const float2 *__restrict__ in, size_t in_stride, const float *__restrict__ phases,
                                  size_t phase_stride, const float *__restrict__ amplitude, float2 *__restrict__ out,
                                  size_t out_stride, int n) {
    const int ch = blockIdx.y;
    const double amp = (double) amplitude[ch];
    const float *ph = phases + (size_t) ch * phase_stride;
    const float2 *x = MULTIPLY ? in + (size_t) ch * in_stride : nullptr;
    float2 *y = out + (size_t) ch * out_stride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double s;
        double c;
        sincos((double) ph[i], &s, &c);
        const float cr = (float) (c * amp);
        const float ci = (float) (s * amp);
        if (MULTIPLY) {
            const float2 a = x[i];
            y[i] = make_float2(__fsub_rn(__fmul_rn(a.x, cr), __fmul_rn(a.y, ci)), __fadd_rn(__fmul_rn(a.x, ci), __fmul_rn(a.y, cr)));
        } else {
            y[i] = make_float2(cr, ci);
        }
    }
}

}  // namespace

extern "C" int sdrm_cu_nco(const sdrm_nco_args *a, void *stream_ptr) {
    if (a->n <= 0 || a->n_ch <= 0) {
        return 0;
    }
    cudaStream_t stream = (cudaStream_t) stream_ptr;
    nco_walk_kernel<<<(a->n_ch + 31) / 32, 32, 0, stream>>>(a->step, a->phase_state, a->phases, a->phase_stride, a->n, a->n_ch);
    int blocks_x = (a->n + 255) / 256;
    if (blocks_x > 64) {
        blocks_x = 64;
    }
    dim3 grid((unsigned) blocks_x, (unsigned) a->n_ch);
    if (a->in != nullptr) {
        nco_rotate_kernel<true><<<grid, 256, 0, stream>>>((const float2 *) a->in, a->in_stride, a->phases, a->phase_stride,
                                                          a->amplitude, (float2 *) a->out, a->out_stride, a->n);
    } else {
        nco_rotate_kernel<false><<<grid, 256, 0, stream>>>(nullptr, 0, a->phases, a->phase_stride, a->amplitude,
                                                           (float2 *) a->out, a->out_stride, a->n);
    }
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}
