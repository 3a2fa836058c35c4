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

#include "device_math.cuh"
#include "sdrm_cuda.h"

namespace {

constexpr float kTwoPi = 6.283185307179586476925286766559f;  // (float) (2 * M_PI), frequency_modulator.c:11
constexpr int kMaxBranchTaps = 64;
constexpr int kTileRows = 32;              // time steps per shared-memory tile of the walker

// Intermediate layout between the three passes ("GTC", as in the demod tail): float [group][time][32 channels], so that
// the serial pass, whose lanes are channels, moves 32 time steps of its 32 channels as one contiguous 4 KB tile.
__device__ __forceinline__ size_t gtc_index(int ch, long long m, size_t rows) {
    return (((size_t) (ch >> 5) * rows + (size_t) m) << 5) + (ch & 31);
}

// out[k * I + p] = scale * sum_j w[k + j] * rev[p][j],  w = (K - 1 carried inputs, then this call's inputs).
// One thread per INPUT k: its K-sample window is gathered once and feeds all I branch filters.
// BITS: inputs are bits unpacked from bytes; otherwise floats. GTC: lane = channel and the result goes to the grouped
// layout (batch path); otherwise lane = input index and the result goes to plain rows (interp_fir_filter handle).
// KU = 8: window held in registers (fully unrolled, guarded by j < K); KU = kMaxBranchTaps: any K, window in local memory.
template <bool BITS, bool GTC, int KU>
__global__ void interp_shape_kernel(const sdrm_interp_args a) {
    __shared__ float taps_s[kMaxBranchTaps * 32];  // [p][j], at most 2048 floats (checked by the launcher)
    const int K = a.branch_taps;
    const int I = a.interpolation;
    for (int i = threadIdx.x; i < K * I; i += blockDim.x) {
        taps_s[i] = a.taps_rev[i];
    }
    __syncthreads();
    int ch;
    int k_first;
    int k_step;
    if (GTC) {
        // blockIdx.y = channel group; 8 warps stride over the inputs, lanes are the group's channels
        ch = blockIdx.y * 32 + (threadIdx.x & 31);
        k_first = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        k_step = gridDim.x * (blockDim.x >> 5);
    } else {
        ch = blockIdx.y;
        k_first = blockIdx.x * blockDim.x + threadIdx.x;
        k_step = gridDim.x * blockDim.x;
    }
    const int chc = ch < a.n_ch ? ch : a.n_ch - 1;  // padding lanes shadow the last channel and write their own column
    const float *hist = a.history + (size_t) chc * (K - 1);
    const uint8_t *bytes = BITS ? (const uint8_t *) a.in + (size_t) chc * a.in_stride : nullptr;
    const float *fin = BITS ? nullptr : (const float *) a.in + (size_t) chc * a.in_stride;
    for (int k = k_first; k < a.n_in; k += k_step) {
        float w[KU];
#pragma unroll(KU == 8 ? 8 : 1)
        for (int j = 0; j < (KU == 8 ? 8 : K); j++) {
            const int i = k + j - (K - 1);  // index into this call's inputs; negative -> carried history
            if (j >= K) {
                w[j] = 0.0f;
            } else if (i < 0) {
                w[j] = hist[(K - 1) + i];
            } else if (BITS) {
                w[j] = ((bytes[i >> 3] >> (7 - (i & 7))) & 1) ? 1.0f : -1.0f;
            } else {
                w[j] = fin[i];
            }
        }
        for (int p = 0; p < I; p++) {
            const float *rev = taps_s + p * K;
            float acc = 0.0f;
#pragma unroll(KU == 8 ? 8 : 1)
            for (int j = 0; j < (KU == 8 ? 8 : K); j++) {
                if (j < K) {
                    acc = __fadd_rn(acc, __fmul_rn(w[j], rev[j]));
                }
            }
            const float v = a.apply_scale ? __fmul_rn(a.scale, acc) : acc;
            const size_t m = (size_t) k * I + p;
            if (GTC) {
                a.out[gtc_index(ch, (long long) m, a.out_stride)] = v;
            } else {
                a.out[(size_t) ch * a.out_stride + m] = v;
            }
        }
    }
}

// The batch modulator's first pass (bytes in, GTC out) for K <= 8: a warp takes one 32-bit word of its 32 channels'
// packets (lane = channel), so the bytes are read once, the +-1 samples never exist as floats (multiplying a tap by
// +-1 is a sign flip, exact) and every store is a full 128-byte row of the GTC layout.
// The first word also needs the carried history (floats, zero before the first call) and takes the general route.
// A shaped sample depends on K consecutive bits only: with the window's bits as an index, all 2^K x I values are computed once
// per CTA into shared memory — by the same sequence of additions, tap 0 first, and the same scaling the direct form uses, so
// the table holds bit for bit what the direct form would produce — and an output costs a shift, a mask, a load and a store
// instead of K sign flips and K additions per branch. LUT = false is the direct form (tables of more than 32 KB: K = 8 with
// more than 32 branches does not occur; kept for the A/B comparison and as the definition of the table's contents).
template <bool LUT>
__global__ void __launch_bounds__(256) bits_shape_kernel(const sdrm_interp_args a) {
    __shared__ float taps_s[8 * 32];  // [p][j]; the launcher checks K <= 8 and I <= 32
    extern __shared__ float lut_s[];  // LUT: [window bits][p]
    const int K = a.branch_taps;
    const int I = a.interpolation;
    for (int i = threadIdx.x; i < K * I; i += blockDim.x) {
        taps_s[i] = a.taps_rev[i];
    }
    __syncthreads();
    if (LUT) {
        // bit (K - 1 - j) of the index is window sample j; a clear bit is the sample -1, i.e. the tap with its sign flipped
        for (int e = threadIdx.x; e < (I << K); e += blockDim.x) {
            const int w = e / I;
            const int p = e - w * I;
            const float *rev = taps_s + p * K;
            float acc = 0.0f;
            for (int j = 0; j < K; j++) {
                const uint32_t flip = ((w >> (K - 1 - j)) & 1) ? 0u : 0x80000000u;
                acc = __fadd_rn(acc, __uint_as_float(__float_as_uint(rev[j]) ^ flip));
            }
            lut_s[e] = a.apply_scale ? __fmul_rn(a.scale, acc) : acc;
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int ch = blockIdx.y * 32 + lane;
    const int chc = ch < a.n_ch ? ch : a.n_ch - 1;
    const float *hist = a.history + (size_t) chc * (K - 1);
    const uint8_t *bytes = (const uint8_t *) a.in + (size_t) chc * a.in_stride;
    const int n_bytes = (a.n_in + 7) >> 3;
    const int n_words = (a.n_in + 31) >> 5;
    float *out = a.out + (((size_t) blockIdx.y * a.out_stride) << 5) + lane;
    auto load_word = [&](int wi) {  // 32 stream bits, first bit in the MSB
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int idx = wi * 4 + b;
            v = (v << 8) | (idx < n_bytes ? (uint32_t) bytes[idx] : 0u);
        }
        return v;
    };
    for (int wi = blockIdx.x * 8 + (threadIdx.x >> 5); wi < n_words; wi += gridDim.x * 8) {
        const uint32_t cur = load_word(wi);
        const int k0 = wi << 5;
        const int nk = min(32, a.n_in - k0);
        if (wi == 0) {
            for (int kk = 0; kk < nk; kk++) {
                float w[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int i = kk + j - (K - 1);
                    w[j] = j >= K ? 0.0f : (i < 0 ? hist[(K - 1) + i] : (((cur >> (31 - i)) & 1) ? 1.0f : -1.0f));
                }
                for (int p = 0; p < I; p++) {
                    const float *rev = taps_s + p * K;
                    float acc = 0.0f;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        if (j < K) {
                            acc = __fadd_rn(acc, __fmul_rn(w[j], rev[j]));
                        }
                    }
                    out[((size_t) kk * I + p) << 5] = a.apply_scale ? __fmul_rn(a.scale, acc) : acc;
                }
            }
            continue;
        }
        const uint64_t comb = ((uint64_t) load_word(wi - 1) << 32) | cur;
        if (LUT) {
            const uint32_t mask = (1u << K) - 1u;
            if (I == 2) {
                const float2 *lut2 = reinterpret_cast<const float2 *>(lut_s);
                for (int kk = 0; kk < nk; kk++) {
                    const float2 v = lut2[(uint32_t) (comb >> (31 - kk)) & mask];
                    float *dst = out + ((size_t) (k0 + kk) << 6);
                    dst[0] = v.x;
                    dst[32] = v.y;
                }
            } else {
                for (int kk = 0; kk < nk; kk++) {
                    const float *row = lut_s + ((uint32_t) (comb >> (31 - kk)) & mask) * I;
                    for (int p = 0; p < I; p++) {
                        out[((size_t) (k0 + kk) * I + p) << 5] = row[p];
                    }
                }
            }
        } else {
            for (int kk = 0; kk < nk; kk++) {
                // bit (K - 1 - j) of x is window sample j; a clear bit is the sample -1, i.e. a flipped tap sign
                const uint32_t flips = ~(uint32_t) (comb >> (31 - kk));
                uint32_t sgn[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    sgn[j] = (flips << (31 - (K - 1 - j))) & 0x80000000u;  // shift counts of absent taps are unused
                }
                for (int p = 0; p < I; p++) {
                    const float *rev = taps_s + p * K;
                    float acc = 0.0f;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        if (j < K) {
                            acc = __fadd_rn(acc, __uint_as_float(__float_as_uint(rev[j]) ^ sgn[j]));
                        }
                    }
                    out[((size_t) (k0 + kk) * I + p) << 5] = a.apply_scale ? __fmul_rn(a.scale, acc) : acc;
                }
            }
        }
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
    // frequency_modulator.c:50-55: q = p + d; q < -2pi -> q + 2pi; q > 2pi -> q - 2pi.
    // q - 2pi and q + 2pi are the same magnitude |q| - 2pi carrying q's sign (float add is symmetric in sign), so a
    // single comparison on |q| selects it: the serial chain is add, add, sign-merge, select (microbench:
    // tools/microbench/wrap_chain.cu, 20.5 cycles per step against 37 for the two-comparison form).
    const float q = __fadd_rn(p, d);
    const float s = __fsub_rn(fabsf(q), kTwoPi);
    return fabsf(q) > kTwoPi ? copysignf(s, q) : q;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MOD_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MOD_WAIT_DONE;\n"
        "bra MOD_WAIT_LOOP;\n"
        "MOD_WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Tile geometry of the walker, chosen by the launcher from the number of channel groups (= CTAs). Every tile costs its warp
// ~450 cycles of barrier wait, proxy fence and copy issue in series with the recurrence, so tiles should be as long as shared
// memory allows: 512 steps per tile (3 tiles, 192 KB) while there is at most one CTA per SM, shorter tiles when several
// CTAs have to share an SM to stay in one wave. C4 (1024 channels x 32768 steps): 0.49 ms per call with 128-step tiles,
// 0.40 with 256, 0.385 with 512; the recurrence alone is 0.34 ms.
template <int ROWS, int STAGES, int PREFETCH>
struct WalkCfg {
    static constexpr int kRows = ROWS;                   // time steps per tile
    static constexpr int kTile = ROWS * 32;              // floats per tile
    static constexpr int kStages = STAGES;               // tiles in the ring
    static constexpr int kPrefetch = PREFETCH;           // loads run this many tiles ahead; a stage is reloaded
                                                         // kStages - kPrefetch tiles after its store
    static constexpr int kSmem = STAGES * ROWS * 32 * 4 + STAGES * 8;
};

// The serial pass: phase = wrap(phase + increment[m]), one warp per group of 32 channels, lane = channel.
// The increments stream in as tiles of Cfg::kRows time steps x 32 channels (contiguous in the GTC layout) through a ring of
// kStages TMA bulk copies, the phases go back the same way in place, so that the recurrence itself — one dependent add
// and select per sample — is all the warp waits for.
template <class Cfg>
__global__ void __launch_bounds__(32) phase_walk_kernel(float *work, size_t rows, float *phase_state, long long n, int n_ch) {
    extern __shared__ __align__(128) unsigned char walk_smem[];
    float(*tiles)[Cfg::kTile] = reinterpret_cast<float(*)[Cfg::kTile]>(walk_smem);
    uint64_t *bars = reinterpret_cast<uint64_t *>(walk_smem + Cfg::kStages * Cfg::kTile * 4);
    const int lane = threadIdx.x;
    const int group = blockIdx.x;
    const int ch = group * 32 + lane;
    float *base = work + (size_t) group * rows * 32;
    const long long n_tiles = (n + Cfg::kRows - 1) / Cfg::kRows;
    if (lane == 0) {
        for (int s = 0; s < Cfg::kStages; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto fetch = [&](long long t) {
        const int s = (int) (t % Cfg::kStages);
        const int nr = (int) min((long long) Cfg::kRows, n - t * Cfg::kRows);
        const uint32_t bytes = (uint32_t) nr * 128u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(tiles[s])),
                     "l"(base + (size_t) t * Cfg::kTile), "r"(bytes), "r"(smem_u32(&bars[s]))
                     : "memory");
    };
    if (lane == 0) {
        for (long long t = 0; t < Cfg::kPrefetch && t < n_tiles; t++) {
            fetch(t);
        }
    }
    float p = ch < n_ch ? phase_state[ch] : 0.0f;
    for (long long t = 0; t < n_tiles; t++) {
        const int s = (int) (t % Cfg::kStages);
        if (lane == 0 && t + Cfg::kPrefetch < n_tiles) {
            // the stage being refilled was stored kStages - kPrefetch tiles ago: all but the newest
            // kStages - kPrefetch - 1 bulk stores must have finished reading shared memory
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(Cfg::kStages - Cfg::kPrefetch - 1) : "memory");
            fetch(t + Cfg::kPrefetch);
        }
        __syncwarp();
        mbar_wait(&bars[s], (uint32_t) ((t / Cfg::kStages) & 1));
        float *tile = tiles[s] + lane;
        const int nr = (int) min((long long) Cfg::kRows, n - t * Cfg::kRows);
        if (nr == Cfg::kRows) {
            for (int r0 = 0; r0 < Cfg::kRows; r0 += 32) {
#pragma unroll
                for (int r = 0; r < 32; r++) {
                    p = wrap_step(p, tile[(r0 + r) * 32]);
                    tile[(r0 + r) * 32] = p;
                }
            }
        } else {
            for (int r = 0; r < nr; r++) {
                p = wrap_step(p, tile[r * 32]);
                tile[r * 32] = p;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (size_t) t * Cfg::kTile),
                         "r"(smem_u32(tiles[s])), "r"((uint32_t) nr * 128u)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    if (ch < n_ch) {
        phase_state[ch] = p;
    }
}

// out[ch][m] = (float) cos(p) + j (float) sin(p), p from the GTC layout. A block takes a [32 time][32 channel] tile
// (coalesced: 32 channels are contiguous), turns it through shared memory so that lanes become time, and each warp
// writes 256-byte runs of one channel's cf32 row.
__global__ void __launch_bounds__(256) phase_to_iq_kernel(const float *__restrict__ phases, size_t rows, float2 *__restrict__ out,
                                                          size_t out_stride, long long n, int n_ch) {
    __shared__ float tile[kTileRows][33];
    const int group = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const float *base = phases + (size_t) group * rows * 32;
    const long long n_tiles = (n + kTileRows - 1) / kTileRows;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const long long m0 = t * kTileRows;
#pragma unroll
        for (int r = warp; r < kTileRows; r += 8) {
            if (m0 + r < n) {
                tile[r][lane] = base[(size_t) (m0 + r) * 32 + lane];
            }
        }
        __syncthreads();
        const long long m = m0 + lane;
#pragma unroll
        for (int c = warp; c < 32; c += 8) {
            const int ch = group * 32 + c;
            if (m < n && ch < n_ch) {
                double s;
                double co;
                sdrm_phase_sincos(tile[lane][c], &s, &co);
                out[(size_t) ch * out_stride + m] = make_float2((float) co, (float) s);
            }
        }
        __syncthreads();
    }
}

// self test: the (cos, sin) pairs the modulator stores for the floats with bit patterns [first_bits, first_bits + count)
__global__ void sincos_selftest_kernel(uint32_t first_bits, size_t count, float2 *out) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t) gridDim.x * blockDim.x) {
        double s;
        double co;
        sdrm_phase_sincos(__uint_as_float(first_bits + (uint32_t) i), &s, &co);
        out[i] = make_float2((float) co, (float) s);
    }
}

}  // namespace

extern "C" int sdrm_cu_selftest_sincos(uint32_t first_bits, size_t count, float *h_cos_sin) {
    if (count == 0) {
        return 0;
    }
    float2 *d = nullptr;
    if (cudaMalloc(&d, count * sizeof(float2)) != cudaSuccess) {
        (void) cudaGetLastError();
        return -12;
    }
    sincos_selftest_kernel<<<148 * 8, 256>>>(first_bits, count, d);
    cudaError_t err = cudaMemcpy(h_cos_sin, d, count * sizeof(float2), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}

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
        if (a->out_grouped) {
            long long bx = ((long long) a->n_in + 7) / 8;
            if (bx > 1024) {
                bx = 1024;
            }
            dim3 grid((unsigned) bx, (unsigned) ((a->n_ch + 31) / 32));
            if (a->in_is_bytes) {
                if (a->branch_taps <= 8 && a->interpolation <= 32) {
                    long long wx = (((long long) a->n_in + 31) / 32 + 7) / 8;
                    const long long want = (148LL * 8 + grid.y - 1) / grid.y;
                    dim3 wgrid((unsigned) (wx < want ? wx : want), grid.y);
                    const size_t lut_bytes = ((size_t) a->interpolation << a->branch_taps) * sizeof(float);
#ifdef MOD_NO_LUT  // A/B builds
                    bits_shape_kernel<false><<<wgrid, 256, 0, stream>>>(*a);
#else
                    if (lut_bytes <= 32768) {
                        bits_shape_kernel<true><<<wgrid, 256, lut_bytes, stream>>>(*a);
                    } else {
                        bits_shape_kernel<false><<<wgrid, 256, 0, stream>>>(*a);
                    }
#endif
                } else if (a->branch_taps <= 8) {
                    interp_shape_kernel<true, true, 8><<<grid, 256, 0, stream>>>(*a);
                } else {
                    interp_shape_kernel<true, true, kMaxBranchTaps><<<grid, 256, 0, stream>>>(*a);
                }
            } else {
                if (a->branch_taps <= 8) {
                    interp_shape_kernel<false, true, 8><<<grid, 256, 0, stream>>>(*a);
                } else {
                    interp_shape_kernel<false, true, kMaxBranchTaps><<<grid, 256, 0, stream>>>(*a);
                }
            }
        } else {
            long long bx = ((long long) a->n_in + 255) / 256;
            if (bx > 2048) {
                bx = 2048;
            }
            if (a->n_ch > 65535) {
                return -22;  // channel-major output puts the channel in gridDim.y; batches use the grouped layout
            }
            dim3 grid((unsigned) bx, (unsigned) a->n_ch);
            if (a->in_is_bytes) {
                if (a->branch_taps <= 8) {
                    interp_shape_kernel<true, false, 8><<<grid, 256, 0, stream>>>(*a);
                } else {
                    interp_shape_kernel<true, false, kMaxBranchTaps><<<grid, 256, 0, stream>>>(*a);
                }
            } else {
                if (a->branch_taps <= 8) {
                    interp_shape_kernel<false, false, 8><<<grid, 256, 0, stream>>>(*a);
                } else {
                    interp_shape_kernel<false, false, kMaxBranchTaps><<<grid, 256, 0, stream>>>(*a);
                }
            }
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

extern "C" int sdrm_cu_phase_walk(float *work, size_t rows, float *phase_state, long long n, int n_ch, void *stream_ptr) {
    if (n <= 0 || n_ch <= 0) {
        return 0;
    }
    if ((long long) rows < n || ((uintptr_t) work & 127) != 0) {
        return -22;
    }
    const int groups = (n_ch + 31) / 32;
    int device = 0, n_sms = 0;
    if (cudaGetDevice(&device) != cudaSuccess || cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
        n_sms = 1;
    }
    cudaStream_t stream = (cudaStream_t) stream_ptr;
    if (2 * groups <= n_sms) {
        // 192 KB per CTA: only while at least half of the SMs stay free for the shaping and trigonometry passes of the
        // neighbouring calls, which overlap this one (4096 channels with 512-step tiles: 1.11 ms per call against 0.91)
        using Cfg = WalkCfg<512, 3, 1>;
        // ... and the walker asks for all the shared memory of its SM, so that no CTA of those passes lands beside it and takes
        // issue slots from its single warp (C4: 0.383 -> 0.368 ms per call)
        int smem = Cfg::kSmem;
        int most = 0;
        if (cudaDeviceGetAttribute(&most, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) == cudaSuccess && most > smem) {
            smem = most;
        }
        cudaFuncSetAttribute(phase_walk_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        phase_walk_kernel<Cfg><<<groups, 32, smem, stream>>>(work, rows, phase_state, n, n_ch);
    } else if (groups <= 2 * n_sms) {
        using Cfg = WalkCfg<256, 3, 1>;  // 96 KB: two CTAs per SM
        cudaFuncSetAttribute(phase_walk_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
        phase_walk_kernel<Cfg><<<groups, 32, Cfg::kSmem, stream>>>(work, rows, phase_state, n, n_ch);
    } else {
        using Cfg = WalkCfg<128, 3, 1>;  // 48 KB: four CTAs per SM, one walker per scheduler
        cudaFuncSetAttribute(phase_walk_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
        phase_walk_kernel<Cfg><<<groups, 32, Cfg::kSmem, stream>>>(work, rows, phase_state, n, n_ch);
    }
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}

extern "C" int sdrm_cu_phase_to_iq(const float *work, size_t rows, void *out, size_t out_stride, long long n, int n_ch,
                                   void *stream_ptr) {
    if (n <= 0 || n_ch <= 0) {
        return 0;
    }
    if ((long long) rows < n) {
        return -22;
    }
    const int groups = (n_ch + 31) / 32;
    long long bx = (n + kTileRows - 1) / kTileRows;
    const long long want = (148LL * 8 + groups - 1) / groups;
    if (bx > want) {
        bx = want;
    }
    dim3 grid((unsigned) bx, (unsigned) groups);
    phase_to_iq_kernel<<<grid, 256, 0, (cudaStream_t) stream_ptr>>>(work, rows, (float2 *) out, out_stride, n, n_ch);
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}

extern "C" int sdrm_cu_freq_mod(float *work, size_t rows, float *phase_state, void *out, size_t out_stride, long long n, int n_ch,
                                void *stream_ptr) {
    int code = sdrm_cu_phase_walk(work, rows, phase_state, n, n_ch, stream_ptr);
    if (code != 0) {
        return code;
    }
    return sdrm_cu_phase_to_iq(work, rows, out, out_stride, n, n_ch, stream_ptr);
}
