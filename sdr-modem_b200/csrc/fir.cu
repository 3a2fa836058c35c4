// Streaming decimating FIR over float2 rows for sm_100a, with the quadrature-demod epilogue of the demod chain.
//
// What it replaces: reference src/dsp/fir_filter.c:93-144 (fir_filter_process_float / _complex, one
// volk_32fc_32f_dot_prod_32fc / volk_32f_x2_dot_prod_32f call per output sample) and, in QD_PAIR mode, the
// conj-multiply + fast_atan2f pass of src/dsp/quadrature_demod.c:57-73 that follows lpf1 in src/dsp/fsk_demod.c:83-87.
//
// Arithmetic contract (exact mode): every output is accumulated by ONE thread, sequentially from tap index 0,
// one accumulator per float2 component, multiply and add rounded separately — the VOLK generic order, so results
// are bit-identical to the reference. ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under
// --fmad=false (checked with cuobjdump on CUDA 12.9), so the separate roundings are expressed as two FFMA2 with
// runtime operands the compiler cannot fold:  p = fma(x, h, -0)  == RN(x*h),  acc = fma(acc, 1, p) == RN(acc + p).
// Fast mode uses a single FFMA2 per tap (same order, fused rounding).
//
// Structure: one CTA = one row x TILE outputs; each thread owns R consecutive outputs and slides a register
// window over the samples (one 16-byte LDS per two taps). The taps, duplicated (h, h) pairs, normally arrive in the
// kernel parameters (constant bank: uniform registers in FMA mode, constant loads in exact mode), 528 per launch for a
// short filter and 7 x 528 per launch for a long one, which then runs as several launches with its accumulators carried
// through a scratch buffer; callers without a host copy of the taps or without scratch get them broadcast from shared
// memory. Samples (tile + halo) are staged by 1-D TMA bulk copies (cp.async.bulk + mbarrier), tap-blocked so that
// shared memory does not grow with the filter length, double-buffered when there is more than one tap block.

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <new>

#include "sdrm_cuda.h"
#include "device_math.cuh"

namespace {

#ifndef FIR_R
#define FIR_R 10
#define FIR_TAPBLOCK 528
#define FIR_CTAS1 3
#define FIR_CTAS2 2
#endif
constexpr int kThreads = 256;
constexpr int kR = FIR_R;                 // outputs per thread (even)
constexpr int kTile = kThreads * kR;      // outputs per CTA
constexpr int kW1 = kR + 2;               // D=1: samples in the register window = taps per window period
constexpr int kW2 = kR + 1;               // D=2: sample pairs in the register window; 2 * kW2 taps per window period
constexpr int kTapBlock = FIR_TAPBLOCK;   // multiple of kW1 and 2 * kW2 (528 = 44 * 12 = 24 * 22 for 10 outputs per thread)
static_assert(kTapBlock == SDRM_FIR_TAP_BLOCK, "sdrm_cuda.h names the tap block for the host layer");
static_assert(kTapBlock % kW1 == 0 && kTapBlock % (2 * kW2) == 0 && kR % 2 == 0, "tap block must hold whole window periods");
constexpr int kSlack = 32;                // float2 of read-ahead slack after each staged sample window
constexpr int kMaxDevices = 64;           // per-device "function attributes are set" flags
constexpr int kMaxGridY = 65535;          // kernels that put rows in gridDim.y loop over the rows beyond it

struct FirCommon {
    const float2 *in;
    size_t in_stride;
    const float2 *hist;
    int hist_len;
    const float2 *taps_dup;
    int n_taps;
    int decimation;
    int phase;
    int n_in;
    int n_out;
    int rows;
    int out_mode;
    void *out;
    size_t out_stride;
    int tc_mask;
    long long tc_head;
    float qd_gain;
    const float *atan_table;
    int tile_stride;    // outputs advanced per tile (kTile, or kTile - 2 with the quad-demod overlap)
    int first_out;      // index of the first computed output of tile 0 (0, or -2 with the overlap)
    int stage_samples;  // float2 per stage for samples
    int n_stages;
    float2 one;         // (1, 1)    runtime so that the compiler cannot simplify the exact-mode FFMA2 pair
    float2 negzero;     // (-0, -0)
    // A long filter in FMA mode runs as several launches, each over the tap blocks [block_first, block_first + n_blocks_here) with
    // that range's taps in its parameters; the accumulators travel between the launches through `acc` (float2 [row][tile][kTile]):
    // read when block_first > 0, written unless this launch holds the filter's last block. The summation order does not change.
    int block_first;
    int n_blocks_here;
    float2 *acc;
};

// Filters of at most CAP taps (per launch) carry their (h, h) pairs in the kernel parameters: the taps are the same for every thread,
// so they can reach the FFMA2 through the constant bank and a uniform register instead of a shared-memory load per tap and warp.
// In FMA mode that is what lifts the pipe limit: FFMA2 acc = x * h + acc with three 64-bit register operands tops out at 74 % of
// the pipe (ncu: math pipe throttle at 74 % busy), with h in a uniform register it reads two.
template <int CAP>
struct FirParamsT : FirCommon {
    float2 taps_c[CAP];
};

constexpr int kLongBlocks = 7;  // tap blocks per launch of a long FMA-mode filter: 7 x 528 (h, h) pairs = 29.6 KB of the 32 KB of parameters
typedef FirParamsT<kTapBlock> FirParams;
typedef FirParamsT<kLongBlocks * kTapBlock> FirParamsLong;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier. 16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <bool FAST>
__device__ __forceinline__ float2 mac2(float2 acc, float2 x, float2 h, float2 one, float2 negzero) {
    if (FAST) {
        return __ffma2_rn(x, h, acc);
    }
    float2 p = __ffma2_rn(x, h, negzero);
    return __ffma2_rn(acc, one, p);
}

template <bool ALIGNED>
__device__ __forceinline__ float4 ld_pair(const float2 *p) {
    if (ALIGNED) {
        return *reinterpret_cast<const float4 *>(p);
    }
    float2 a = p[0];
    float2 b = p[1];
    return make_float4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ float2 lo(const float4 &v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi(const float4 &v) { return make_float2(v.z, v.w); }

// ---- decimation 1: acc[r] += s[j + r] * h[j]; register window of 12 samples, period 12 taps -------------------
template <bool FAST, bool ALIGNED, bool GUARD, bool TP, typename P>
__device__ __forceinline__ void fir_d1_body(float2 (&acc)[kR], float4 (&w)[kW1 / 2], const float2 *s, const float2 *hs, const P &p,
                                            int j, int ju, int n_taps, float2 one, float2 negzero) {
#pragma unroll
    for (int u = 0; u < kW1; u++) {
        if (!GUARD || j + u < n_taps) {
            float2 h = TP ? p.taps_c[ju + u] : hs[j + u];
#pragma unroll
            for (int r = 0; r < kR; r++) {
                const int e = (u + r) % kW1;
                float2 x = (e & 1) ? hi(w[e >> 1]) : lo(w[e >> 1]);
                acc[r] = mac2<FAST>(acc[r], x, h, one, negzero);
            }
            if (u & 1) {
                // samples j+u-1 and j+u are no longer needed: refill their slots with the pair one window further on
                w[(u - 1) >> 1] = ld_pair<ALIGNED>(s + j + u + kW1 - 1);
            }
        }
    }
}

// tap_base: index in p.taps_c of this block's first tap (0 for a single-block filter)
template <bool FAST, bool ALIGNED, bool TP, typename P>
__device__ __forceinline__ void fir_d1_block(float2 (&acc)[kR], const float2 *s, const float2 *hs, const P &p, int n_taps,
                                             int tap_base, float2 one, float2 negzero) {
    float4 w[kW1 / 2];
#pragma unroll
    for (int k = 0; k < kW1 / 2; k++) {
        w[k] = ld_pair<ALIGNED>(s + 2 * k);
    }
    int j = 0;
    // In FMA mode the parameter taps are indexed with a count of their own, hidden from the optimiser (which would merge it with j
    // again), so that it can live in a uniform register and the taps be loaded into uniform registers. Exact mode indexes them with
    // the sample counter (see the launcher).
    int ju = tap_base;
    if (FAST) {
        asm volatile("mov.u32 %0, %1;" : "=r"(ju) : "r"(tap_base));
    }
    for (; j + kW1 <= n_taps; j += kW1, ju += kW1) {
        fir_d1_body<FAST, ALIGNED, false, TP>(acc, w, s, hs, p, j, FAST ? ju : tap_base + j, n_taps, one, negzero);
    }
    if (j < n_taps) {
        fir_d1_body<FAST, ALIGNED, true, TP>(acc, w, s, hs, p, j, FAST ? ju : tap_base + j, n_taps, one, negzero);
    }
}

// ---- decimation 2: acc[r] += s[j + 2r] * h[j]; even/odd windows of 11 samples each, period 22 taps ------------
template <bool FAST, bool ALIGNED, bool GUARD, bool TP, typename P>
__device__ __forceinline__ void fir_d2_body(float2 (&acc)[kR], float4 (&w)[kW2], const float2 *s, const float2 *hs, const P &p,
                                            int j, int ju, int n_taps, float2 one, float2 negzero) {
#pragma unroll
    for (int q = 0; q < kW2; q++) {
        if (!GUARD || j + 2 * q < n_taps) {
            float2 h = TP ? p.taps_c[ju + 2 * q] : hs[j + 2 * q];
#pragma unroll
            for (int r = 0; r < kR; r++) {
                acc[r] = mac2<FAST>(acc[r], lo(w[(q + r) % kW2]), h, one, negzero);
            }
        }
        if (!GUARD || j + 2 * q + 1 < n_taps) {
            float2 h = TP ? p.taps_c[ju + 2 * q + 1] : hs[j + 2 * q + 1];
#pragma unroll
            for (int r = 0; r < kR; r++) {
                acc[r] = mac2<FAST>(acc[r], hi(w[(q + r) % kW2]), h, one, negzero);
            }
        }
        if (!GUARD || j + 2 * q + 2 < n_taps) {
            w[q] = ld_pair<ALIGNED>(s + j + 2 * q + 2 * kW2);
        }
    }
}

template <bool FAST, bool ALIGNED, bool TP, typename P>
__device__ __forceinline__ void fir_d2_block(float2 (&acc)[kR], const float2 *s, const float2 *hs, const P &p, int n_taps,
                                             int tap_base, float2 one, float2 negzero) {
    float4 w[kW2];
#pragma unroll
    for (int k = 0; k < kW2; k++) {
        w[k] = ld_pair<ALIGNED>(s + 2 * k);
    }
    int j = 0;
    int ju = tap_base;
    if (FAST) {
        asm volatile("mov.u32 %0, %1;" : "=r"(ju) : "r"(tap_base));
    }
    for (; j + 2 * kW2 <= n_taps; j += 2 * kW2, ju += 2 * kW2) {
        fir_d2_body<FAST, ALIGNED, false, TP>(acc, w, s, hs, p, j, FAST ? ju : tap_base + j, n_taps, one, negzero);
    }
    if (j < n_taps) {
        fir_d2_body<FAST, ALIGNED, true, TP>(acc, w, s, hs, p, j, FAST ? ju : tap_base + j, n_taps, one, negzero);
    }
}

// Issues the TMA copies for one tap block of one tile: samples v[start, start + count) and taps [j0, j0 + nt).
// start is even; the window may straddle the history / input boundary and the end of the input.
template <bool TP>
__device__ void stage_load(const FirCommon &p, int row, long long start, int count, int j0, int nt, float2 *smp, float2 *tps,
                           uint64_t *bar) {
    const float2 *hist_row = p.hist + (size_t) row * p.hist_len;
    const float2 *in_row = p.in + (size_t) row * p.in_stride;
    long long end = start + count;
    if (end > p.n_in) {
        end = p.n_in;
    }
    // history part [start, min(end, 0))
    long long h_end = end < 0 ? end : 0;
    long long h_cnt = h_end > start ? h_end - start : 0;
    // input part [max(start, 0), end)
    long long i_beg = start > 0 ? start : 0;
    long long i_cnt = end > i_beg ? end - i_beg : 0;
    long long i_bulk = i_cnt & ~1LL;
    long long h_bulk = h_cnt & ~1LL;  // h_cnt is even whenever end >= 0 (start and hist boundary are even)
    // odd leftovers (only at the very end of an odd-length input): plain stores, made visible to the waiters by
    // the release semantics of the mbarrier arrive that follows
    if (h_cnt & 1) {
        smp[h_bulk] = hist_row[p.hist_len + start + h_bulk];
    }
    if (i_cnt & 1) {
        smp[(i_beg - start) + i_bulk] = in_row[i_beg + i_bulk];
    }
    uint32_t tap_bytes = TP ? 0u : (uint32_t) (((nt + 1) & ~1) * sizeof(float2));
    uint32_t bytes = (uint32_t) ((h_bulk + i_bulk) * sizeof(float2)) + tap_bytes;
    mbar_expect_tx(bar, bytes);
    if (h_bulk > 0) {
        tma_load_1d(smp, hist_row + (p.hist_len + start), (uint32_t) (h_bulk * sizeof(float2)), bar);
    }
    if (i_bulk > 0) {
        tma_load_1d(smp + (i_beg - start), in_row + i_beg, (uint32_t) (i_bulk * sizeof(float2)), bar);
    }
    if (!TP) {
        tma_load_1d(tps, p.taps_dup + j0, tap_bytes, bar);
    }
}

template <int D, bool FAST, bool ALIGNED, bool TP, typename P>
__global__ void __launch_bounds__(kThreads, D == 1 ? FIR_CTAS1 : FIR_CTAS2) fir_tile_kernel(const P p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[2];
    __shared__ float2 last_out[kThreads];
    __shared__ float atan_s[257];

    const int row = blockIdx.x;
    const int tile = blockIdx.y;
    const int tid = threadIdx.x;

    float2 *const smem_f2 = reinterpret_cast<float2 *>(smem_raw);
    const int stage_f2 = p.stage_samples + kTapBlock;  // per stage: sample window, then the tap block

    // first computed output of this tile and the sample its tap 0 multiplies
    const long long m0 = (long long) tile * p.tile_stride + p.first_out;
    const long long v0 = (long long) p.phase + m0 * D - (p.n_taps - 1);
    const int odd = (int) (v0 & 1);
    const long long v0_al = v0 - odd;
    const int window = (kTile - 1) * D + 1;  // samples spanned by the tile's outputs for one tap
    // tap blocks of this launch: all of them, except for a long FMA-mode filter that runs as several launches
    const int n_blocks = p.n_blocks_here;
    const int block_first = p.block_first;
    const bool last_launch = (block_first + n_blocks) * kTapBlock >= p.n_taps;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (p.out_mode == SDRM_FIR_OUT_QD_PAIR && last_launch) {
        for (int i = tid; i < 257; i += kThreads) {
            atan_s[i] = p.atan_table[i];
        }
    }
    __syncthreads();

    auto issue = [&](int b, int st) {  // b: tap block of this launch
        const int j0 = (block_first + b) * kTapBlock;
        const int nt = min(kTapBlock, p.n_taps - j0);
        // taps [j0, j0 + nt) of this tile touch v0 + j0 ... v0 + j0 + window + nt - 2
        const int count = (odd + window + nt - 1 + 1) & ~1;
        float2 *smp = smem_f2 + st * stage_f2;
        stage_load<TP>(p, row, v0_al + j0, count, j0, nt, smp, smp + p.stage_samples, &bars[st]);
    };

    if (tid == 0) {
        issue(0, 0);
        if (p.n_stages > 1 && n_blocks > 1) {
            issue(1, 1);
        }
    }

    float2 acc[kR];
    // accumulators in flight between the launches of a long filter: float2 [row][tile][kTile], a thread's kR values contiguous
    float2 *const acc_mine = p.acc == nullptr ? nullptr : p.acc + (((size_t) row * gridDim.y + tile) * kThreads + tid) * kR;
    if (block_first > 0) {
#pragma unroll
        for (int r = 0; r < kR; r += 2) {
            const float4 v = *reinterpret_cast<const float4 *>(acc_mine + r);
            acc[r] = make_float2(v.x, v.y);
            acc[r + 1] = make_float2(v.z, v.w);
        }
    } else {
#pragma unroll
        for (int r = 0; r < kR; r++) {
            acc[r] = make_float2(0.0f, 0.0f);
        }
    }
    const float2 one = p.one;
    const float2 negzero = p.negzero;
    // The last tile of a row is usually ragged (131072 outputs are 51.2 tiles): warps whose outputs all lie beyond the end of
    // the row skip the arithmetic (they still take part in the barriers) instead of filtering samples nobody stores.
    const bool warp_has_outputs = m0 + (long long) (tid & ~31) * kR < p.n_out;

    for (int b = 0; b < n_blocks; b++) {
        const int st = (p.n_stages > 1) ? (b & 1) : 0;
        const int nt = min(kTapBlock, p.n_taps - (block_first + b) * kTapBlock);
        mbar_wait(&bars[st], (uint32_t) ((p.n_stages > 1 ? (b >> 1) : b) & 1));
        const float2 *smp = smem_f2 + st * stage_f2;
        const float2 *tps = smp + p.stage_samples;
        const float2 *s = smp + odd + tid * (kR * D);
        if (warp_has_outputs) {
            if (D == 1) {
                fir_d1_block<FAST, ALIGNED, TP>(acc, s, tps, p, nt, b * kTapBlock, one, negzero);
            } else {
                fir_d2_block<FAST, ALIGNED, TP>(acc, s, tps, p, nt, b * kTapBlock, one, negzero);
            }
        }
        if (b + p.n_stages < n_blocks) {
            __syncthreads();  // everyone is done reading this stage before TMA overwrites it
            if (tid == 0) {
                issue(b + p.n_stages, st);
            }
        }
    }

    if (!last_launch) {
#pragma unroll
        for (int r = 0; r < kR; r += 2) {
            *reinterpret_cast<float4 *>(acc_mine + r) = make_float4(acc[r].x, acc[r].y, acc[r + 1].x, acc[r + 1].y);
        }
        return;
    }
    const long long m_first = m0 + (long long) tid * kR;
    if (p.out_mode == SDRM_FIR_OUT_ROWS) {
        float2 *out = reinterpret_cast<float2 *>(p.out) + (size_t) row * p.out_stride;
#pragma unroll
        for (int r = 0; r < kR; r++) {
            long long m = m_first + r;
            if (m < p.n_out) {
                out[m] = acc[r];
            }
        }
    } else if (p.out_mode == SDRM_FIR_OUT_TC) {
        float *out = reinterpret_cast<float *>(p.out);
#pragma unroll
        for (int r = 0; r < kR; r++) {
            long long m = m_first + r;
            if (m < p.n_out) {
                size_t ring_row = (size_t) ((p.tc_head + m) & p.tc_mask);
                // GTC layout: [channel group of 32][ring row][32]
                *reinterpret_cast<float2 *>(out + (((size_t) (row >> 4) * (p.tc_mask + 1) + ring_row) << 5) + ((2 * row) & 31)) = acc[r];
            }
        }
    } else {
        // quadrature demod: out[o] = gain * atan2(y[o] * conj(y[o-1])); y[o-1] of a thread's first output comes
        // from its left neighbour; the tile's first two outputs exist only to provide that for the third.
        last_out[tid] = acc[kR - 1];
        __syncthreads();
        float2 prev = tid > 0 ? last_out[tid - 1] : make_float2(0.0f, 0.0f);
        float *out = reinterpret_cast<float *>(p.out) + (size_t) (row >> 1) * p.out_stride * 2 + (row & 1);
        const long long emit_from = m0 + 2;
#pragma unroll
        for (int r = 0; r < kR; r++) {
            long long m = m_first + r;
            float2 cur = acc[r];
            if (m >= emit_from && m < p.n_out) {
                out[2 * m] = sdrm_quad_demod_sample(cur, prev, p.qd_gain, atan_s);
            }
            prev = cur;
        }
    }
}

// Any decimation: one thread per output, samples read through L1/L2. Only used where the FIR is a negligible share
// of the chain (lpf2 behind a large decimation) or for unusual standalone filters.
template <bool FAST>
__global__ void fir_generic_kernel(const FirCommon p) {
    const long long m = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= p.n_out) {
        return;
    }
    for (int row = blockIdx.y; row < p.rows; row += gridDim.y) {  // gridDim.y <= 65535
        const float2 *hist_row = p.hist + (size_t) row * p.hist_len + p.hist_len;
        const float2 *in_row = p.in + (size_t) row * p.in_stride;
        long long v = (long long) p.phase + m * p.decimation - (p.n_taps - 1);
        float2 acc = make_float2(0.0f, 0.0f);
        for (int j = 0; j < p.n_taps; j++, v++) {
            float2 x = v < 0 ? hist_row[v] : in_row[v];
            acc = mac2<FAST>(acc, x, p.taps_dup[j], p.one, p.negzero);
        }
        if (p.out_mode == SDRM_FIR_OUT_ROWS) {
            reinterpret_cast<float2 *>(p.out)[(size_t) row * p.out_stride + m] = acc;
        } else {
            size_t ring_row = (size_t) ((p.tc_head + m) & p.tc_mask);
            *reinterpret_cast<float2 *>(reinterpret_cast<float *>(p.out) + (((size_t) (row >> 4) * (p.tc_mask + 1) + ring_row) << 5) +
                                        ((2 * row) & 31)) = acc;
        }
    }
}

// Decimation >= 3 with the input staged in shared memory: a CTA takes kDecOutputs consecutive outputs of one row, whose
// inputs form one contiguous span of (kDecOutputs - 1) * D + n_taps samples, read from global memory once and coalesced.
// Thread t accumulates output t sequentially over the taps (the reference's order); it reads sample t * D + j, i.e. lanes
// are D samples apart. The span is therefore stored in segments of D samples padded to an odd length, which spreads the 16
// lanes of a half-warp's 64-bit loads over all 32 banks for every D; the position of tap j inside the segments (j mod D,
// j div D) is tracked by two warp-uniform counters. Taps are broadcast from shared memory.
// Bound by the shared-memory crossbar (one 8-byte load per lane for two FFMA2), about a third of the exact-mode FP32
// rate: these are the light filters (lpf2 behind a large decimation, channel filters), 5x faster than reading the
// strided samples through L1 as fir_generic_kernel does.
constexpr int kDecOutputs = 128;

template <bool FAST>
__global__ void __launch_bounds__(kDecOutputs) fir_dec_kernel(const FirCommon p, int seg_len, int n_segs) {
    extern __shared__ __align__(16) float2 dec_smem[];
    float2 *taps_s = dec_smem;                  // n_taps (h, h) pairs
    float2 *span = dec_smem + p.n_taps;         // n_segs segments of seg_len (>= D, odd) samples
    const int D = p.decimation;
    const long long m0 = (long long) blockIdx.x * kDecOutputs;
    const int n_here = (int) min((long long) kDecOutputs, (long long) p.n_out - m0);
    // first input of the span: the oldest sample of output m0
    const long long v0 = (long long) p.phase + m0 * D - (p.n_taps - 1);
    const int span_len = (n_here - 1) * D + p.n_taps;
    for (int i = threadIdx.x; i < p.n_taps; i += kDecOutputs) {
        taps_s[i] = p.taps_dup[i];
    }
    for (int row = blockIdx.y; row < p.rows; row += gridDim.y) {  // gridDim.y <= 65535; one pass for smaller batches
        const float2 *hist_row = p.hist + (size_t) row * p.hist_len + p.hist_len;
        const float2 *in_row = p.in + (size_t) row * p.in_stride;
        for (int i = threadIdx.x; i < span_len; i += kDecOutputs) {
            const long long v = v0 + i;
            const float2 x = v < 0 ? hist_row[v] : in_row[v];
            span[(i / D) * seg_len + (i % D)] = x;
        }
        __syncthreads();
        if ((int) threadIdx.x < n_here) {
            float2 acc = make_float2(0.0f, 0.0f);
            const float2 *mine = span + (size_t) threadIdx.x * seg_len;  // sample t * D sits at the start of segment t
            int jr = 0;                                                  // j mod D
            const float2 *seg = mine;                                    // segment t + j div D
#pragma unroll 4
            for (int j = 0; j < p.n_taps; j++) {
                acc = mac2<FAST>(acc, seg[jr], taps_s[j], p.one, p.negzero);
                if (++jr == D) {
                    jr = 0;
                    seg += seg_len;
                }
            }
            const long long m = m0 + threadIdx.x;
            if (p.out_mode == SDRM_FIR_OUT_ROWS) {
                reinterpret_cast<float2 *>(p.out)[(size_t) row * p.out_stride + m] = acc;
            } else {
                size_t ring_row = (size_t) ((p.tc_head + m) & p.tc_mask);
                *reinterpret_cast<float2 *>(reinterpret_cast<float *>(p.out) + (((size_t) (row >> 4) * (p.tc_mask + 1) + ring_row) << 5) +
                                            ((2 * row) & 31)) = acc;
            }
        }
        __syncthreads();  // the span is overwritten by the next row
    }
}

__global__ void hist_update_kernel(const float2 *in, size_t in_stride, const float2 *hist, float2 *hist_next, int hist_len,
                                   int n_in, int rows) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= hist_len) {
        return;
    }
    const long long v = (long long) n_in - hist_len + k;
    for (int row = blockIdx.y; row < rows; row += gridDim.y) {  // gridDim.y <= 65535
        float2 x = v < 0 ? hist[(size_t) row * hist_len + hist_len + v] : in[(size_t) row * in_stride + v];
        hist_next[(size_t) row * hist_len + k] = x;
    }
}

__global__ void quad_demod_kernel(const float2 *in, size_t in_stride, float2 *prev, float gain, const float *atan_table,
                                  float *out, size_t out_stride, int n_in, int rows) {
    __shared__ float atan_s[257];
    for (int i = threadIdx.x; i < 257; i += blockDim.x) {
        atan_s[i] = atan_table[i];
    }
    __syncthreads();
    for (int row = blockIdx.y; row < rows; row += gridDim.y) {  // gridDim.y <= 65535
        const float2 *x = in + (size_t) row * in_stride;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += gridDim.x * blockDim.x) {
            float2 before = i > 0 ? x[i - 1] : prev[row];
            out[(size_t) row * out_stride + i] = sdrm_quad_demod_sample(x[i], before, gain, atan_s);
        }
    }
}

__global__ void quad_demod_carry_kernel(const float2 *in, size_t in_stride, float2 *prev, int n_in, int rows) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row < rows && n_in > 0) {
        prev[row] = in[(size_t) row * in_stride + n_in - 1];
    }
}

template <int D, bool FAST, bool ALIGNED, bool TP, typename P>
int launch_tile_tp(const P &p, int rows, int tiles, size_t smem, cudaStream_t stream) {
    auto kernel = fir_tile_kernel<D, FAST, ALIGNED, TP, P>;
    cudaError_t err = cudaSuccess;
    // Function attributes are per device and sticky: set them once per (instantiation, device) for the largest size this
    // instantiation can ask for (two stages), not on every launch (~10 us each, most of a single-handle call's launch time).
    static std::atomic<bool> configured[kMaxDevices];
    int device = 0;
    cudaGetDevice(&device);
    if (device < 0 || device >= kMaxDevices || !configured[device].load(std::memory_order_acquire)) {
        const size_t stage = (size_t) ((((kTile - 1) * D + 1 + kTapBlock + 2 + kSlack) + 1) & ~1) + kTapBlock;
        err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (2 * stage * sizeof(float2)));
        if (err != cudaSuccess) {
            return -(int) err - 1000;
        }
        // Same (maximum) shared-memory carveout as the tail kernel: an SM whose carveout was sized for one of the two kernels
        // alone cannot take CTAs of the other until it drains, which serialises the filters of call k+1 against the tail of call k.
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (device >= 0 && device < kMaxDevices) {
            configured[device].store(true, std::memory_order_release);
        }
    }
    dim3 grid((unsigned) rows, (unsigned) tiles);
    kernel<<<grid, kThreads, smem, stream>>>(p);
    err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}

template <int D, bool FAST, bool ALIGNED>
int launch_tile(const FirParams &p, bool taps_in_params, int rows, int tiles, size_t smem, cudaStream_t stream) {
    if (taps_in_params) {
        return launch_tile_tp<D, FAST, ALIGNED, true, FirParams>(p, rows, tiles, smem, stream);
    }
    return launch_tile_tp<D, FAST, ALIGNED, false, FirParams>(p, rows, tiles, smem, stream);
}

// A long filter: ceil(n_blocks / kLongBlocks) launches, each with its tap blocks in the kernel parameters (in FMA mode: uniform
// loads, h as the FFMA2's uniform operand, as for short filters) and the accumulators carried through `acc` in between. The
// parameter block is 30 KB, so it lives on the heap, one per calling thread.
template <int D, bool FAST, bool ALIGNED>
int launch_tile_long(const FirCommon &common, const float2 *h_taps_dup, int n_blocks, int rows, int tiles, size_t smem,
                     cudaStream_t stream) {
    static thread_local FirParamsLong *params = nullptr;
    if (params == nullptr) {
        params = new (std::nothrow) FirParamsLong;
        if (params == nullptr) {
            return -12;
        }
    }
    static_cast<FirCommon &>(*params) = common;
    for (int first = 0; first < n_blocks; first += kLongBlocks) {
        const int here = n_blocks - first < kLongBlocks ? n_blocks - first : kLongBlocks;
        const int tap0 = first * kTapBlock;
        const int n_here = common.n_taps - tap0 < here * kTapBlock ? common.n_taps - tap0 : here * kTapBlock;
        memcpy(params->taps_c, h_taps_dup + tap0, (size_t) n_here * sizeof(float2));
        for (int j = n_here; j < kLongBlocks * kTapBlock; j++) {
            params->taps_c[j] = make_float2(0.0f, 0.0f);
        }
        params->block_first = first;
        params->n_blocks_here = here;
        const int code = launch_tile_tp<D, FAST, ALIGNED, true, FirParamsLong>(*params, rows, tiles, smem, stream);
        if (code != 0) {
            return code;
        }
    }
    return 0;
}

}  // namespace

extern "C" size_t sdrm_cu_fir_scratch_bytes(int rows, int max_out) {
    const size_t tiles = ((size_t) max_out + (kTile - 2) - 1) / (kTile - 2);  // the quad-demod tiling has the shorter stride
    return (size_t) rows * tiles * kTile * sizeof(float2);
}

extern "C" int sdrm_cu_fir(const sdrm_fir_args *a, void *stream_ptr) {
    cudaStream_t stream = (cudaStream_t) stream_ptr;
    if (a->n_out <= 0 || a->rows <= 0) {
        return 0;
    }
    const bool qd = a->out_mode == SDRM_FIR_OUT_QD_PAIR;
    if (a->n_taps < 1 || a->decimation < 1 || (a->hist_len & 1) || a->hist_len < a->n_taps - 1 + (qd ? 2 : 0) ||
        (qd && a->decimation != 1)) {
        return -22;
    }
    FirParams p;
    p.in = (const float2 *) a->in;
    p.in_stride = a->in_stride;
    p.hist = (const float2 *) a->hist;
    p.hist_len = a->hist_len;
    p.taps_dup = (const float2 *) a->taps_dup;
    p.n_taps = a->n_taps;
    p.decimation = a->decimation;
    p.phase = a->phase;
    p.n_in = a->n_in;
    p.n_out = a->n_out;
    p.rows = a->rows;
    p.out_mode = a->out_mode;
    p.out = a->out;
    p.out_stride = a->out_stride;
    p.tc_mask = a->tc_ring_rows - 1;
    p.tc_head = a->tc_head;
    p.qd_gain = a->qd_gain;
    p.atan_table = a->atan_table;
    p.one = make_float2(1.0f, 1.0f);
    p.negzero = make_float2(-0.0f, -0.0f);
    p.block_first = 0;
    p.n_blocks_here = 0;
    p.acc = nullptr;

    if (a->decimation > 2) {
        if (qd) {
            return -22;
        }
        p.tile_stride = 0;
        p.first_out = 0;
        p.stage_samples = 0;
        p.n_stages = 0;
        // staged kernel when the span of 128 outputs and the taps fit in shared memory
        const int seg_len = a->decimation | 1;
        const int n_segs = kDecOutputs + (a->n_taps + a->decimation - 1) / a->decimation;
        const size_t dec_smem = ((size_t) n_segs * seg_len + a->n_taps) * sizeof(float2);
        if (dec_smem <= 200 * 1024) {
            dim3 dgrid((unsigned) ((a->n_out + kDecOutputs - 1) / kDecOutputs), (unsigned) (a->rows < kMaxGridY ? a->rows : kMaxGridY));
            cudaError_t derr;
            if (a->fast) {
                derr = cudaFuncSetAttribute(fir_dec_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dec_smem);
                fir_dec_kernel<true><<<dgrid, kDecOutputs, dec_smem, stream>>>(static_cast<const FirCommon &>(p), seg_len, n_segs);
            } else {
                derr = cudaFuncSetAttribute(fir_dec_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dec_smem);
                fir_dec_kernel<false><<<dgrid, kDecOutputs, dec_smem, stream>>>(static_cast<const FirCommon &>(p), seg_len, n_segs);
            }
            if (derr != cudaSuccess) {
                return -(int) derr - 1000;
            }
            derr = cudaGetLastError();
            return derr == cudaSuccess ? 0 : -(int) derr - 1000;
        }
        dim3 grid((unsigned) ((a->n_out + 127) / 128), (unsigned) (a->rows < kMaxGridY ? a->rows : kMaxGridY));
        if (a->fast) {
            fir_generic_kernel<true><<<grid, 128, 0, stream>>>(static_cast<const FirCommon &>(p));
        } else {
            fir_generic_kernel<false><<<grid, 128, 0, stream>>>(static_cast<const FirCommon &>(p));
        }
        cudaError_t err = cudaGetLastError();
        return err == cudaSuccess ? 0 : -(int) err - 1000;
    }

    const int D = a->decimation;
    p.tile_stride = qd ? kTile - 2 : kTile;
    p.first_out = qd ? -2 : 0;
    const int n_blocks = (a->n_taps + kTapBlock - 1) / kTapBlock;
    p.n_stages = n_blocks > 1 ? 2 : 1;
    p.stage_samples = (((kTile - 1) * D + 1 + kTapBlock + 2 + kSlack) + 1) & ~1;
    const size_t smem = (size_t) p.n_stages * (p.stage_samples + kTapBlock) * sizeof(float2);
    const int tiles = (a->n_out + p.tile_stride - 1) / p.tile_stride;
    // sample windows start on an even float2 (16 bytes) iff phase - (n_taps - 1) + first_out * D is even
    const long long v0 = (long long) a->phase + (long long) p.first_out * D - (a->n_taps - 1);
    const bool aligned = ((v0 & 1) == 0) && ((p.tile_stride * D) % 2 == 0) && ((a->in_stride & 1) == 0) &&
                         (((uintptr_t) a->in & 15) == 0) && (((uintptr_t) a->hist & 15) == 0);
    if ((((uintptr_t) a->in | (uintptr_t) a->hist | (uintptr_t) a->taps_dup) & 15) != 0 || (a->in_stride & 1)) {
        return -22;  // TMA bulk copies need 16-byte aligned rows
    }
    // Short filters carry their taps in the kernel parameters (see FirParams::taps_c). FMA mode indexes them with a counter of its
    // own, which the compiler keeps in a uniform register: uniform loads, h as the FFMA2's uniform operand, K1 4.78 -> 3.97 ms. Exact
    // mode indexes them with the sample counter and gets a register-indexed constant load into ordinary registers: its FFMA2 pair
    // already takes -0 and 1 from uniform registers and an instruction has one uniform operand (h there: 7.65 ms), but the constant
    // load still saves the shared-memory load per tap and warp: K1 7.50 -> 7.45 ms, lpf2 0.966 -> 0.947 ms.
    p.n_blocks_here = n_blocks;
    if (n_blocks > 1 && a->h_taps_dup != nullptr && a->acc_scratch != nullptr) {
        // long filter: taps through the kernel parameters, kLongBlocks tap blocks per launch (FMA mode: uniform registers, 71 -> 93 %
        // of the pipe; exact mode: constant loads instead of a shared-memory load per tap and warp, worth 3 % as for short filters)
        p.acc = (float2 *) a->acc_scratch;
        const float2 *taps = (const float2 *) a->h_taps_dup;
#define SDRM_LONG(DD, FF, AA) return launch_tile_long<DD, FF, AA>(p, taps, n_blocks, a->rows, tiles, smem, stream)
        if (D == 1) {
            if (a->fast) {
                if (aligned) SDRM_LONG(1, true, true);
                SDRM_LONG(1, true, false);
            }
            if (aligned) SDRM_LONG(1, false, true);
            SDRM_LONG(1, false, false);
        }
        if (a->fast) {
            if (aligned) SDRM_LONG(2, true, true);
            SDRM_LONG(2, true, false);
        }
        if (aligned) SDRM_LONG(2, false, true);
        SDRM_LONG(2, false, false);
#undef SDRM_LONG
    }
    const bool taps_in_params = a->h_taps_dup != nullptr && n_blocks == 1;
    if (taps_in_params) {
        memcpy(p.taps_c, a->h_taps_dup, (size_t) a->n_taps * sizeof(float2));
        for (int j = a->n_taps; j < kTapBlock; j++) {
            p.taps_c[j] = make_float2(0.0f, 0.0f);
        }
    }
#define SDRM_LAUNCH(DD, FF, AA) return launch_tile<DD, FF, AA>(p, taps_in_params, a->rows, tiles, smem, stream)
    if (D == 1) {
        if (a->fast) {
            if (aligned) SDRM_LAUNCH(1, true, true);
            SDRM_LAUNCH(1, true, false);
        }
        if (aligned) SDRM_LAUNCH(1, false, true);
        SDRM_LAUNCH(1, false, false);
    }
    if (a->fast) {
        if (aligned) SDRM_LAUNCH(2, true, true);
        SDRM_LAUNCH(2, true, false);
    }
    if (aligned) SDRM_LAUNCH(2, false, true);
    SDRM_LAUNCH(2, false, false);
#undef SDRM_LAUNCH
}

extern "C" int sdrm_cu_hist_update(const void *in, size_t in_stride, const void *hist, void *hist_next, int hist_len,
                                   int n_in, int rows, void *stream_ptr) {
    if (rows <= 0 || hist_len <= 0) {
        return 0;
    }
    dim3 grid((unsigned) ((hist_len + 255) / 256), (unsigned) (rows < kMaxGridY ? rows : kMaxGridY));
    hist_update_kernel<<<grid, 256, 0, (cudaStream_t) stream_ptr>>>((const float2 *) in, in_stride, (const float2 *) hist,
                                                                    (float2 *) hist_next, hist_len, n_in, rows);
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}

extern "C" int sdrm_cu_quad_demod(const void *in, size_t in_stride, void *prev, float gain, const float *atan_table,
                                  float *out, size_t out_stride, int n_in, int rows, void *stream_ptr) {
    if (rows <= 0 || n_in <= 0) {
        return 0;
    }
    cudaStream_t stream = (cudaStream_t) stream_ptr;
    int blocks = (n_in + 255) / 256;
    if (blocks > 1024) {
        blocks = 1024;
    }
    dim3 grid((unsigned) blocks, (unsigned) (rows < kMaxGridY ? rows : kMaxGridY));
    quad_demod_kernel<<<grid, 256, 0, stream>>>((const float2 *) in, in_stride, (float2 *) prev, gain, atan_table, out,
                                                out_stride, n_in, rows);
    quad_demod_carry_kernel<<<(rows + 127) / 128, 128, 0, stream>>>((const float2 *) in, in_stride, (float2 *) prev, n_in, rows);
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}
