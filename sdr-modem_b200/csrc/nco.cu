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

#include "device_math.cuh"
#include "sdrm_cuda.h"

namespace {

constexpr float kTwoPi = 6.283185307179586476925286766559f;  // (float) (2 * M_PI), sig_source.c:11

// phases[ch][i] = phase used for sample i; carries the accumulator. One CTA of two warps per 32 channels.
// Warp 0, lane = channel, walks: per sample only the dependent chain (add, add, sign-merge, select: wrap_chain.cu) and
// one conflict-free shared-memory store into a [time][channel] tile. Warp 1 writes the previous tile out transposed, so
// that the global stores are 128-byte runs of one channel's row (the first version stored 32 scattered words per sample
// from the walking lane itself: 44 cycles per sample; this one 22).
__device__ __forceinline__ void nco_store_tile(const float *tile, float *phases, size_t stride, int ch0, int rows, int i0, int count,
                                               int lane) {
    if (lane >= count) {
        return;
    }
    float *dst = phases + (size_t) ch0 * stride + i0 + lane;
    if (rows == 32) {
        // loads first, eight at a time: a rolled LDS -> STG loop waits out the shared-memory latency 32 times
#pragma unroll
        for (int r0 = 0; r0 < 32; r0 += 8) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                v[k] = tile[lane * 33 + r0 + k];
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                dst[(size_t) (r0 + k) * stride] = v[k];
            }
        }
    } else {
        for (int r = 0; r < rows; r++) {
            dst[(size_t) r * stride] = tile[lane * 33 + r];
        }
    }
}

__global__ void __launch_bounds__(64) nco_walk_kernel(const float *__restrict__ step, float *__restrict__ phase_state,
                                                      float *__restrict__ phases, size_t stride, int n, int n_ch) {
    __shared__ float tiles[2][32 * 33];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int ch0 = blockIdx.x * 32;
    const int ch = ch0 + lane;
    const bool valid = ch < n_ch;
    const int rows = min(32, n_ch - ch0);
    const int n_tiles = (n + 31) / 32;
    const float w = valid && warp == 0 ? step[ch] : 0.0f;
    float p = valid && warp == 0 ? phase_state[ch] : 0.0f;
    for (int k = 0; k <= n_tiles; k++) {
        if (warp == 0) {
            if (k < n_tiles) {
                float *tile = tiles[k & 1];
                const int count = min(32, n - k * 32);
                float after[33];  // after[32]: the state after `count` steps of a partial last tile
#pragma unroll
                for (int r = 0; r < 32; r++) {
                    tile[r * 33 + lane] = p;
                    if (r == count) {
                        after[32] = p;
                    }
                    // sig_source.c:49-55: p += w; p < -2pi -> p += 2pi; p > 2pi -> p -= 2pi. Both corrections are |p| - 2pi
                    // with p's sign, selected by one comparison on |p| (the reference's two tests can never both fire).
                    const float q = __fadd_rn(p, w);
                    const float m = __fsub_rn(fabsf(q), kTwoPi);
                    p = fabsf(q) > kTwoPi ? copysignf(m, q) : q;
                }
                if (count < 32) {
                    p = after[32];  // a partial last tile: the accumulator stops after `count` steps
                }
            }
        } else if (k > 0) {
            nco_store_tile(tiles[(k - 1) & 1], phases, stride, ch0, rows, (k - 1) * 32, min(32, n - (k - 1) * 32), lane);
        }
        __syncthreads();
    }
    if (valid && warp == 0) {
        phase_state[ch] = p;
    }
}

template <bool MULTIPLY>
__global__ void nco_rotate_kernel(const float2 *__restrict__ in, size_t in_stride, const float *__restrict__ phases,
                                  size_t phase_stride, const float *__restrict__ amplitude, float2 *__restrict__ out,
                                  size_t out_stride, int n, int n_ch) {
    for (int ch = blockIdx.y; ch < n_ch; ch += gridDim.y) {  // gridDim.y <= 65535
        const double amp = (double) amplitude[ch];
        const float *ph = phases + (size_t) ch * phase_stride;
        const float2 *x = MULTIPLY ? in + (size_t) ch * in_stride : nullptr;
        float2 *y = out + (size_t) ch * out_stride;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            double s;
            double c;
            sdrm_phase_sincos(ph[i], &s, &c);
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
}

}  // namespace

extern "C" int sdrm_cu_nco(const sdrm_nco_args *a, void *stream_ptr) {
    if (a->n <= 0 || a->n_ch <= 0) {
        return 0;
    }
    cudaStream_t stream = (cudaStream_t) stream_ptr;
    nco_walk_kernel<<<(a->n_ch + 31) / 32, 64, 0, stream>>>(a->step, a->phase_state, a->phases, a->phase_stride, a->n, a->n_ch);
    int blocks_x = (a->n + 255) / 256;
    if (blocks_x > 64) {
        blocks_x = 64;
    }
    dim3 grid((unsigned) blocks_x, (unsigned) (a->n_ch < 65535 ? a->n_ch : 65535));
    if (a->in != nullptr) {
        nco_rotate_kernel<true><<<grid, 256, 0, stream>>>((const float2 *) a->in, a->in_stride, a->phases, a->phase_stride,
                                                          a->amplitude, (float2 *) a->out, a->out_stride, a->n, a->n_ch);
    } else {
        nco_rotate_kernel<false><<<grid, 256, 0, stream>>>(nullptr, 0, a->phases, a->phase_stride, a->amplitude,
                                                           (float2 *) a->out, a->out_stride, a->n, a->n_ch);
    }
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}
