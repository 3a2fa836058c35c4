// GFSK modulator for sm_100a: the reference's gfsk_mod chain (src/dsp/gfsk_mod.c:102-132) as three passes over a batch
// of channels.
//
//   1. interp_shape_kernel  bytes -> +-1 bits (MSB first, gfsk_mod.c:109-120) -> polyphase interpolating FIR
//                           (interp_fir_filter.c:139-154: I branch filters h_p[k] = h[k*I + p], outputs interleaved),
//                           times the modulator sensitivity: d[m] = sens * shaped[m]. Parallel over output samples;
//                           every output is one thread's sequential K-tap dot product (VOLK generic order).
//   2. phase_walk_kernel    p = wrap(p + d[m]) in float (frequency_modulator.c:48-55). The accumulator is never reset
//                           and float addition is not associative, so this is a true serial recurrence per channel:
//                           one lane per channel; everything else is kept off its critical path.
//   3. phase_to_iq_kernel   out[m] = (float) cos(p[m]) + j (float) sin(p[m]) in double precision, as the reference
//                           (frequency_modulator.c:56). Parallel over samples, coalesced cf32 stores.
//
// Pass 2 bounds small batches: with C channels there are only C lanes of serial work (about 14 cycles per sample),
// so the HBM roofline of pass 3 (8 B per output sample) is reached only from a few thousand channels per GPU up
// (DESIGN.md has the numbers). The same kernels serve the reference's interp_fir_filter and frequency_modulator handles.

#include <cuda_runtime.h>
#include <stdint.h>

#include "sdrm_cuda.h"

namespace {

constexpr float kTwoPi = 6.283185307179586476925286766559f;  // (float) (2 * M_PI), frequency_modulator.c:11
constexpr int kMaxBranchTaps = 64;

// out[ch][k * I + p] = scale * sum_j w[k + j] * rev[p][j],  w = (K - 1 carried inputs, then this call's inputs).
// BITS: inputs are bits unpacked from bytes; otherwise floats. scale_on: multiply by the sensitivity (separately rounded).
template <bool BITS>
__global__ void interp_shape_kernel(const sdrm_interp_args a) {
    __shared__ float taps_s[kMaxBranchTaps * 32];  // [p][j], at most 2048 floats (checked by the launcher)
    const int K = a.branch_taps;
    const int I = a.interpolation;
    for (int i = threadIdx.x; i < K * I; i += blockDim.x) {
        taps_s[i] = a.taps_rev[i];
    }
    __syncthreads();
    const int ch = blockIdx.y;
    const long long n_out = (long long) a.n_in * I;
    const float *hist = a.history + (size_t) ch * (K - 1);
    const uint8_t *bytes = BITS ? (const uint8_t *) a.in + (size_t) ch * a.in_stride : nullptr;
    const float *fin = BITS ? nullptr : (const float *) a.in + (size_t) ch * a.in_stride;
    float *out = a.out + (size_t) ch * a.out_stride;
    for (long long m = (long long) blockIdx.x * blockDim.x + threadIdx.x; m < n_out; m += (long long) gridDim.x * blockDim.x) {
        const int k = (int) (m / I);
        const int p = (int) (m - (long long) k * I);
        const float *rev = taps_s + p * K;
        float acc = 0.0f;
        for (int j = 0; j < K; j++) {
            const int i = k + j - (K - 1);  // index into this call's inputs; negative -> carried history
            float w;
            if (i < 0) {
                w = hist[(K - 1) + i];
            } else if (BITS) {
                w = ((bytes[i >> 3] >> (7 - (i & 7))) & 1) ? 1.0f : -1.0f;
            } else {
                w = fin[i];
            }
            acc = __fadd_rn(acc, __fmul_rn(w, rev[j]));
        }
        out[m] = a.apply_scale ? __fmul_rn(a.scale, acc) : acc;
    }
}

// new_history = last K - 1 inputs of (history, this call's inputs)
template <bool BITS>
__global__ void interp_history_kernel(const sdrm_interp_args a) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= a.n_ch) {
        return;
    }
    const int H = a.branch_taps - 1;
    float next[kMaxBranchTaps];
    const float *hist = a.history + (size_t) ch * H;
    for (int q = 0; q < H; q++) {
        const int i = a.n_in - H + q;
        if (i < 0) {
            next[q] = hist[H + i];
        } else if (BITS) {
            const uint8_t *bytes = (const uint8_t *) a.in + (size_t) ch * a.in_stride;
            next[q] = ((bytes[i >> 3] >> (7 - (i & 7))) & 1) ? 1.0f : -1.0f;
        } else {
            next[q] = ((const float *) a.in)[(size_t) ch * a.in_stride + i];
        }
    }
    for (int q = 0; q < H; q++) {
        a.history[(size_t) ch * H + q] = next[q];
    }
}

__device__ __forceinline__ float wrap_step(float p, float d) {
    // phase += d; if (phase < -2pi) phase += 2pi; if (phase > 2pi) phase -= 2pi;   both candidates are formed
    // speculatively so that the serial chain is one add plus one select deep
    const float q = __fadd_rn(p, d);
    const float up = __fadd_rn(q, kTwoPi);
    const float down = __fsub_rn(q, kTwoPi);
    return q < -kTwoPi ? up : (q > kTwoPi ? down : q);
}

// lane = channel. pre_add: the phase is advanced before it is used (frequency_modulator) or after (sig_source).
// increments: [ch][stride] per-sample steps, rows 16-byte aligned. phases: same layout, the phase each sample uses.
__global__ void phase_walk_kernel(const float *increments, float *phases, size_t stride,
                                  float *phase_state, int n, int n_ch) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= n_ch) {
        return;
    }
    const float4 *d4 = reinterpret_cast<const float4 *>(increments + (size_t) ch * stride);
    float4 *p4 = reinterpret_cast<float4 *>(phases + (size_t) ch * stride);
    float p = phase_state[ch];
    const int n4 = n >> 2;
    int i = 0;
    for (; i + 4 <= n4; i += 4) {
        float4 d[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            d[u] = d4[i + u];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            float4 o;
            p = wrap_step(p, d[u].x);
            o.x = p;
            p = wrap_step(p, d[u].y);
            o.y = p;
            p = wrap_step(p, d[u].z);
            o.z = p;
            p = wrap_step(p, d[u].w);
            o.w = p;
            p4[i + u] = o;
        }
    }
    const float *dd = increments + (size_t) ch * stride;
    float *pp = phases + (size_t) ch * stride;
    for (int m = i * 4; m < n; m++) {
        p = wrap_step(p, dd[m]);
        pp[m] = p;
    }
    phase_state[ch] = p;
}

__global__ void phase_to_iq_kernel(const float *__restrict__ phases, size_t phase_stride, float2 *__restrict__ out,
                                   size_t out_stride, long long n) {
    const int ch = blockIdx.y;
    const float *ph = phases + (size_t) ch * phase_stride;
    float2 *y = out + (size_t) ch * out_stride;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        double s;
        double c;
        sincos((double) ph[i], &s, &c);
        y[i] = make_float2((float) c, (float) s);
    }
}

}  // namespace

extern "C" int sdrm_cu_interp_fir(const sdrm_interp_args *a, void *stream_ptr) {
    if (a->n_ch <= 0) {
        return 0;
    }
    if (a->branch_taps < 1 || a->branch_taps > kMaxBranchTaps || a->interpolation < 1 ||
        a->branch_taps * a->interpolation > kMaxBranchTaps * 32) {
        return -22;
    }
    cudaStream_t stream = (cudaStream_t) stream_ptr;
    const long long n_out = (long long) a->n_in * a->interpolation;
    if (n_out > 0) {
        long long bx = (n_out + 255) / 256;
        if (bx > 2048) {
            bx = 2048;
        }
        dim3 grid((unsigned) bx, (unsigned) a->n_ch);
        if (a->in_is_bytes) {
            interp_shape_kernel<true><<<grid, 256, 0, stream>>>(*a);
        } else {
            interp_shape_kernel<false><<<grid, 256, 0, stream>>>(*a);
        }
    }
    if (a->branch_taps > 1) {
        if (a->in_is_bytes) {
            interp_history_kernel<true><<<(a->n_ch + 127) / 128, 128, 0, stream>>>(*a);
        } else {
            interp_history_kernel<false><<<(a->n_ch + 127) / 128, 128, 0, stream>>>(*a);
        }
    }
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}

extern "C" int sdrm_cu_freq_mod(const float *increments, float *phases, size_t stride, float *phase_state, void *out,
                                size_t out_stride, long long n, int n_ch, void *stream_ptr) {
    if (n <= 0 || n_ch <= 0) {
        return 0;
    }
    if ((stride & 3) != 0) {
        return -22;
    }
    cudaStream_t stream = (cudaStream_t) stream_ptr;
    phase_walk_kernel<<<(n_ch + 31) / 32, 32, 0, stream>>>(increments, phases, stride, phase_state, (int) n, n_ch);
    long long bx = (n + 255) / 256;
    if (bx > 1024) {
        bx = 1024;
    }
    dim3 grid((unsigned) bx, (unsigned) n_ch);
    phase_to_iq_kernel<<<grid, 256, 0, stream>>>(phases, stride, (float2 *) out, out_stride, n);
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}
