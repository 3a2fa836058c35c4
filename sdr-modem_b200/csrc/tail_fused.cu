// Fused serial tail of the demod chain for sm_100a: DC blocker -> Mueller & Mueller clock recovery -> int8, one kernel.
//
// Same arithmetic as tail.cu (which keeps the two stages as separate kernels for the reference's per-block handles);
// this kernel is the throughput path behind fsk_demod. The running sums of the dc blocker and the timing loop are
// feedback loops whose float rounding is part of the result, so a stream is never split in time: one LANE owns one
// channel. What is parallel is the cascade: the four moving averages of the dc blocker and the clock loop are five
// pipeline stages, each run by its own WARP of the CTA (32 channels per CTA), skewed by one 32-row block:
//
//     step t:   warp 0  MA0 on block t      (rows from the TC ring in global memory)
//               warp 1  MA1 on block t-1    (input from warp 0 through shared memory, double buffered)
//               warp 2  MA2 on block t-2
//               warp 3  MA3 on block t-3    -> x[n-(2L-2)] - y4, appended to the per-lane sample ring in shared memory
//               warp 4  clock recovery over everything complete, interpolating straight out of that ring
//     barrier among the four producer warps (bar.sync 1)
//
// The clock warp is NOT part of that step barrier: a step of the producers costs about the same every time, the clock's
// cost varies from step to step (lanes sit at different symbol phases, so the same number of new rows is a different number
// of trips of the warp), and in lock step every step cost the maximum of the two. The sample ring is therefore a proper
// producer / consumer queue of step-sized blocks with one "full" and one "empty" mbarrier per block slot: the last producer
// waits for a slot to be released before it overwrites it and signals it full, the clock warp takes whatever is complete and
// releases the blocks every active lane has moved past. The two sides then cost max(sum, sum) instead of sum(max).
//
// Data movement is all TMA bulk copies (cp.async.bulk). Every global array the loop touches is blocked by groups of 32
// channels ([group][row][32]), so a block of 32 rows of one group is 4 KB of contiguous memory and moves with ONE copy
// issued by one lane (two when it wraps around a ring): a stage's delayed inputs are fetched into shared memory one
// block ahead of the arithmetic and complete on an mbarrier; the stage's own inputs go back to its delay line straight
// from the shared-memory tile they arrived in. No per-row address arithmetic or register staging is left in the loop
// (per-row LDG/STG, then per-row cp.async, then per-row bulk copies were each measured and were each the bottleneck).
// Inside a stage and block, only the running sum y[n] = d[n] + y[n-1] is serial (one FADD per row); the 32 subtractions
// and the 32 divisions by L are independent and unrolled. The division by the constant L is branch-free (two Markstein
// corrections of sum * RN(1/L), checked against IEEE division by sdrm_cu_selftest_div); values outside a safe exponent
// range redo their block with __fdiv_rn.
// What the reference carries from call to call in its working buffer (clock_recovery_mm.c:127-135) is saved from the
// shared-memory ring into a small per-channel carry array at the end of the call and reloaded at the start of the next.
//
// Mirrors: reference src/dsp/dc_blocker.c:52-64,105-119; src/dsp/clock_recovery_mm.c:78-139;
// src/dsp/mmse_fir_interpolator.c:188-191; src/dsp/fir_filter.c:116-121; src/dsp/fsk_demod.c:106.

#include <cuda_runtime.h>

#include <atomic>
#include <stdint.h>

#include "sdrm_cuda.h"

namespace {

// Register budget of the tail kernel. ptxas takes 168 registers when only the block size bounds it; capped at 144 it spills 8
// bytes and the kernel is 4 % faster (1.835 -> 1.764 ms at C2; 152: 2.05 ms, 136: 2.17 ms: the schedule changes with the
// allocation, measured, not predicted). What it does NOT change is what the tail costs the filter kernel beside it (K1 7.52 ms
// against 7.30 alone at every budget): that cost is not a matter of how many K1 CTAs fit next to a tail CTA.
#ifndef TAIL_REGS
#define TAIL_REGS 144
#endif
#define TAIL_BOUNDS(threads) __maxnreg__(TAIL_REGS)
constexpr int kBlockRows = 32;             // rows per pipeline step
constexpr int kTile = kBlockRows * 32;     // floats per [row][lane] tile (4 KB)
constexpr int kRowBytes = 32 * 4;          // one row of 32 channels
constexpr int kGuard = 16;                 // slack between what a lane may lag behind and the size of its ring
constexpr int kMirror = 16;                // ring rows [0, kMirror) are kept twice, again at [slots, slots + kMirror), so
                                           // that the clock's 11-sample window never wraps
constexpr int kTapsFloats = 129 * 8 + 24;  // MMSE bank, padded to a multiple of 128 bytes

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
        "TAIL_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TAIL_WAIT_DONE;\n"
        "bra TAIL_WAIT_LOOP;\n"
        "TAIL_WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void producers_barrier(int threads) { asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory"); }

// Rows [first, first + n) of a circular array of `len` rows of 32 channels (128 bytes each, contiguous) <-> a tile in
// shared memory: one TMA bulk copy, or two when the span wraps around the end of the array.
__device__ __forceinline__ void bulk_load_rows(float *smem_dst, const float *gmem_rows, int first, int n, int len, uint64_t *bar) {
    const int part = min(n, len - first);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_rows + (size_t) first * 32), "r"(part * kRowBytes), "r"(smem_u32(bar))
                 : "memory");
    if (part < n) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(smem_dst + part * 32)),
                     "l"(gmem_rows), "r"((n - part) * kRowBytes), "r"(smem_u32(bar))
                     : "memory");
    }
}

__device__ __forceinline__ void bulk_store_rows(float *gmem_rows, const float *smem_src, int first, int n, int len) {
    const int part = min(n, len - first);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_rows + (size_t) first * 32),
                 "r"(smem_u32(smem_src)), "r"(part * kRowBytes)
                 : "memory");
    if (part < n) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_rows), "r"(smem_u32(smem_src + part * 32)),
                     "r"((n - part) * kRowBytes)
                     : "memory");
    }
}

// one lane of a converged warp, chosen by the hardware: the compiler then issues the bulk copies from uniform registers
// without the per-thread "waterfall" loop it builds around `if (lane == 0)`
__device__ __forceinline__ bool elect_one() {
    uint32_t is_leader;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(is_leader));
    return is_leader != 0;
}

__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }

// End-of-step wait of the lane that issued a block's bulk stores. The stores have finished READING their shared-memory
// tile (it is rewritten in the next step); their global writes must be complete before the same slots are fetched
// again, which is (L - 64) / 32 steps later at the earliest, so up to `slack` younger store groups may stay in flight.
__device__ __forceinline__ void bulk_wait_step(int slack) {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (slack >= 2) {
        asm volatile("cp.async.bulk.wait_group 2;" ::: "memory");
    } else if (slack == 1) {
        asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
    } else {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// generic-proxy writes to shared memory become visible to the async proxy (bulk stores issued after the next barrier)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float slice_pm1(float x) { return x < 0.0f ? -1.0f : 1.0f; }

__device__ __forceinline__ float branchless_clip(float x, float clip) {
    return __fmul_rn(0.5f, __fsub_rn(fabsf(__fadd_rn(x, clip)), fabsf(__fsub_rn(x, clip))));
}

// the arithmetic mode is a template parameter: as a run-time flag every tap of the interpolator was issued in both forms (an
// FFMA, then a predicated FADD over the same register) inside the loop's dependent chain
template <bool FAST>
__device__ __forceinline__ float dot_step(float acc, float v, float t) {
    return FAST ? __fmaf_rn(v, t, acc) : __fadd_rn(acc, __fmul_rn(v, t));
}

// sum / L, correctly rounded, for a constant L with rcp = RN(1/L): Markstein corrections of q0 = RN(sum * rcp).
// Valid inside the exponent range where the residuals are exact (see SumRange). One correction is enough for most
// lengths, every length the reference's parameters produce among them, but not provably for all: the host checks the
// length at create over all 2^23 mantissas (host/taps.c, sdrm_division_steps) and passes STEPS = 1 or 2.
// A zero sum: +0 gives q0 = e0 = q1 = ... = +0. (-0 would come back as +0, but the running sums of this kernel are never -0:
// they start at +0 and RN(x + (-x)) = +0; the per-block dc_blocker handle in tail.cu keeps an explicit test.)
template <int STEPS>
__device__ __forceinline__ float div_by_length(float sum, float length_f, float rcp) {
    if (STEPS == 0) {
        return __fdiv_rn(sum, length_f);
    }
    const float q0 = __fmul_rn(sum, rcp);
    const float e0 = __fmaf_rn(-q0, length_f, sum);
    const float q1 = __fmaf_rn(e0, rcp, q0);
    if (STEPS == 1) {
        return q1;
    }
    const float e1 = __fmaf_rn(-q1, length_f, sum);
    return __fmaf_rn(e1, rcp, q1);
}

// Sums for which div_by_length is not proven (huge, non-finite, or tiny but non-zero) make their block use __fdiv_rn.
// The test runs over the 32 sums of a block as two unsigned reductions of the magnitude bits shifted left by one (the
// sign falls off, order is preserved): the largest, and the smallest of (bits - 1), where the wrap-around sends a zero
// sum — which the fast path divides correctly — to the far end. Three integer instructions per sum instead of three
// comparisons and two predicate merges.
struct SumRange {
    uint32_t largest = 0u;
    uint32_t smallest_m1 = 0xffffffffu;
    __device__ __forceinline__ void add(float sum) {
        const uint32_t u = __float_as_uint(sum) << 1;
        largest = max(largest, u);
        smallest_m1 = min(smallest_m1, u - 1u);
    }
    // 1.0e18f = 0x5D5E0B6B, 1.0e-18f = 0x219392EF
    __device__ __forceinline__ bool outside() const {
        return largest >= (0x5D5E0B6Bu << 1) || smallest_m1 < (0x219392EFu << 1) - 1u;
    }
};

// the same test on one value (self test)
__device__ __forceinline__ bool outside_fast_division_range(float sum) {
    SumRange r;
    r.add(sum);
    return r.outside();
}

// Shared-memory map of one CTA.
struct Layout {
    float *taps;     // MMSE bank
    float *ring;     // [ring_slots][32] clock sample ring
    float *pipe;     // [PROD - 1][2] tiles handed from stage to stage
    float *rows;     // [2] tiles: TC ring rows for warp 0
    float *line;     // [PROD][2] tiles: delayed inputs of each stage
    float *dx;       // [2] tiles: group delay line values for the last stage
    uint64_t *bars;  // [PROD][2] mbarriers of the producers' fetches
    uint64_t *full;  // [ring blocks] the last producer has written this block of the sample ring
    uint64_t *empty; // [ring blocks] every active lane of the clock warp has moved past this block
};

// Global arrays are blocked by groups of 32 channels ([group][row][32]), so a block of 32 rows of one group is 4 KB of
// contiguous memory. These return the group's slice of each array.
__device__ __forceinline__ const float *group_rows(const sdrm_tail_args &a, int group) {
    return a.rows + (size_t) group * a.ring_rows * 32;
}
__device__ __forceinline__ float *group_line(const sdrm_tail_args &a, int group, int stage) {
    return a.delay + ((size_t) stage * a.n_groups + group) * a.dc_length * 32;
}
__device__ __forceinline__ float *group_dx(const sdrm_tail_args &a, int group) {
    return a.delay + (size_t) 4 * a.n_groups * a.dc_length * 32 + (size_t) group * a.dx_length * 32;
}

// One elected lane fetches block b for this warp's stage: its rows of the TC ring (warp 0), the stage's own inputs of L
// rows ago (its delay line) and, for the last stage, x[n - (2L - 2)] from the group delay line. Completion is counted on
// the warp's mbarrier of parity b & 1.
// The group's slices of the global arrays a producer warp touches, resolved once per kernel (recomputing the 64-bit
// addresses for every copy was most of what the elected lane did).
struct Arrays {
    const float *rows;  // lpf2 output ring (first stage)
    float *line;        // this stage's delay line
    float *dx;          // group delay line (first stage writes, last stage reads)
};

// Cursors of a producer warp into the circular arrays, for the block it is working on. They advance by one block per
// step with a compare-and-subtract (all lengths are >= 32 rows); a 64-bit modulo per copy, as a first version had, was a
// quarter of all instructions the kernel executed, on one lane, in front of everything else.
struct Cursors {
    int line;      // stage delay line: slot of the block's first row
    int dx_store;  // group delay line: where the first stage writes
    int dx_load;   // group delay line: where the last stage reads x[n - (2L - 2)]
};

__device__ __forceinline__ int advance(int pos, int len, int rows) {
    pos += rows;
    return pos >= len ? pos - len : pos;
}

template <int SB>
__device__ __forceinline__ Cursors next_block(const sdrm_tail_args &a, const Cursors &c) {
    Cursors n;
    n.line = advance(c.line, a.dc_length, SB * kBlockRows);
    n.dx_store = advance(c.dx_store, a.dx_length, SB * kBlockRows);
    n.dx_load = advance(c.dx_load, a.dx_length, SB * kBlockRows);
    return n;
}

template <int PROD, int SB>
__device__ __forceinline__ void fetch_block(const sdrm_tail_args &a, const Layout &s, const Arrays &g, int warp, int lane, int b,
                                            const Cursors &c) {
    if (!elect_one()) {
        return;
    }
    const bool has_dc = PROD == 4;
    const int row0 = b * SB * kBlockRows;
    const int nr = min(SB * kBlockRows, a.n_rows - row0);
    const int parity = b & 1;
    uint64_t *bar = s.bars + warp * 2 + parity;
    const int copies = (warp == 0 ? 1 : 0) + (has_dc ? 1 : 0) + (has_dc && warp == PROD - 1 ? 1 : 0);
    mbar_expect_tx(bar, (uint32_t) (copies * nr * kRowBytes));
    if (warp == 0) {
        const int first = (int) ((a.head + row0) & (a.ring_rows - 1));
        bulk_load_rows(s.rows + parity * SB * kTile, g.rows, first, nr, a.ring_rows, bar);
    }
    if (has_dc) {
        bulk_load_rows(s.line + (warp * 2 + parity) * SB * kTile, g.line, c.line, nr, a.dc_length, bar);
        if (warp == PROD - 1) {
            bulk_load_rows(s.dx + parity * SB * kTile, g.dx, c.dx_load, nr, a.dx_length, bar);
        }
    }
}

// A sample enters the clock's ring: row `pos` and, for the first kMirror rows, its copy behind the end.
__device__ __forceinline__ void ring_put(float *ring_lane, int pos, int ring_slots, float v) {
    ring_lane[pos * 32] = v;
    if (pos < kMirror) {
        ring_lane[(pos + ring_slots) * 32] = v;
    }
}

// One producer warp's arithmetic for one block whose inputs are staged:
//   moving average  y = in - in[n-L] + y_prev ; out = y / L                           (dc_blocker.c:52-64)
//   last stage      x[n-(2L-2)] - y4 appended to the clock's sample ring              (dc_blocker.c:110-114)
// FULL blocks carry no per-row guards, so the 32 rows form one basic block that ptxas interleaves freely.
template <int PROD, int SB, bool FULL, int DIVSTEPS>
__device__ __forceinline__ void producer_block(const sdrm_tail_args &a, const Layout &s, int warp, int lane, int b, int h, int nr,
                                               float &sum, float rcp) {
    const bool has_dc = PROD == 4;
    const int ring_mask = a.ring_slots - 1;
    const int row0 = (b * SB + h) * kBlockRows;  // h: which 32-row block of the step's SB
    const int parity = b & 1;
    const float *in_tile = (warp == 0 ? s.rows + parity * SB * kTile : s.pipe + ((warp - 1) * 2 + parity) * SB * kTile) + h * kTile;
    const float *src = in_tile + lane;
    // this block's 32 rows of the clock's ring: aligned to the block size, so they never wrap (ring_slots is a power of two)
    float *ring_block = s.ring + lane + (row0 & ring_mask) * 32;
    const bool mirrored = (row0 & ring_mask) == 0;  // rows [0, kMirror) are stored a second time behind the ring's end
    if (!has_dc) {
#pragma unroll
        for (int r = 0; r < kBlockRows; r++) {
            if (FULL || r < nr) {
                ring_block[r * 32] = src[r * 32];
            }
        }
        if (mirrored) {
#pragma unroll
            for (int r = 0; r < kMirror; r++) {
                if (FULL || r < nr) {
                    ring_block[(r + a.ring_slots) * 32] = src[r * 32];
                }
            }
        }
        return;
    }

    const float length_f = (float) a.dc_length;
    const float *delayed = s.line + (warp * 2 + parity) * SB * kTile + h * kTile + lane;
    float y[kBlockRows];
#pragma unroll
    for (int r = 0; r < kBlockRows; r++) {
        if (FULL || r < nr) {
            y[r] = __fsub_rn(src[r * 32], delayed[r * 32]);
        }
    }
#pragma unroll
    for (int r = 0; r < kBlockRows; r++) {
        if (FULL || r < nr) {
            sum = __fadd_rn(y[r], sum);
            y[r] = sum;
        }
    }
    // The last stage also needs x[n - (2L - 2)]: fetched now, while registers are free, so that the loads are not strung
    // between the stores at the end (they were: LDS, FADD, STS one after the other, 35 cycles a row).
    float xv[kBlockRows];
    if (warp == PROD - 1) {
        const float *xd = s.dx + parity * SB * kTile + h * kTile + lane;
#pragma unroll
        for (int r = 0; r < kBlockRows; r++) {
            if (FULL || r < nr) {
                xv[r] = xd[r * 32];
            }
        }
    }
    // All 32 quotients as pure arithmetic, in place: 32 independent FMUL + 4 FFMA chains that ptxas interleaves. (With the
    // stores and the stage test inside this loop every row became its own basic block behind a branch, the chains ran one
    // after the other through the same two registers and a block took 5600 cycles instead of a few hundred.)
    SumRange range;
#pragma unroll
    for (int r = 0; r < kBlockRows; r++) {
        if (FULL || r < nr) {
            range.add(y[r]);
        }
    }
    if (DIVSTEPS == 0 || __any_sync(0xffffffffu, range.outside())) {
#pragma unroll
        for (int r = 0; r < kBlockRows; r++) {
            if (FULL || r < nr) {
                y[r] = __fdiv_rn(y[r], length_f);
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < kBlockRows; r++) {
            if (FULL || r < nr) {
                y[r] = div_by_length<DIVSTEPS>(y[r], length_f, rcp);
            }
        }
    }
    if (warp < PROD - 1) {
        float *dst = s.pipe + (warp * 2 + parity) * SB * kTile + h * kTile + lane;
#pragma unroll
        for (int r = 0; r < kBlockRows; r++) {
            if (FULL || r < nr) {
                dst[r * 32] = y[r];
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < kBlockRows; r++) {
            if (FULL || r < nr) {
                y[r] = __fsub_rn(xv[r], y[r]);
                ring_block[r * 32] = y[r];
            }
        }
        if (mirrored) {
#pragma unroll
            for (int r = 0; r < kMirror; r++) {
                if (FULL || r < nr) {
                    ring_block[(r + a.ring_slots) * 32] = y[r];
                }
            }
        }
    }
}

// PROD producer warps (4 moving averages, or 1 plain copier when the dc blocker is off) + 1 clock warp.
// All per-channel arrays are padded to a multiple of 32 channels, so every lane owns real memory.
// SB = 32-row blocks per pipeline step. A step has fixed costs that do not depend on its size (issuing and retiring the bulk
// copies, the fences, the step barrier: about half of a 32-row step), and the clock warp loses fewer trips to lanes that sit
// at different symbol phases when more rows arrive at once, so two blocks per step are used wherever shared memory and the
// delay-line lengths allow it (8 KB tiles).
template <int PROD, int DIVSTEPS, bool FAST, int SB>
__global__ void TAIL_BOUNDS((PROD + 1) * 32) demod_tail_kernel(const sdrm_tail_args a) {
    extern __shared__ __align__(128) float smem[];
    Layout s;
    s.taps = smem;
    s.ring = s.taps + kTapsFloats;
    s.pipe = s.ring + (size_t) (a.ring_slots + kMirror) * 32;
    s.rows = s.pipe + (PROD - 1) * 2 * SB * kTile;
    s.line = s.rows + 2 * SB * kTile;
    s.dx = s.line + PROD * 2 * SB * kTile;
    s.bars = reinterpret_cast<uint64_t *>(s.dx + 2 * SB * kTile);
    constexpr int kStepRows = SB * kBlockRows;
    constexpr int kStepShift = SB == 1 ? 5 : 6;       // log2(kStepRows)
    const int ring_blocks = a.ring_slots >> kStepShift;  // step-sized block slots of the sample ring (a power of two >= 2)
    s.full = s.bars + PROD * 2;
    s.empty = s.full + ring_blocks;
    const int ring_mask = a.ring_slots - 1;
    for (int i = threadIdx.x; i < 129 * 8; i += blockDim.x) {
        s.taps[i] = a.mmse_taps[i];
    }
    // The warp's index decides its role, and every address the role computes is the same for its 32 lanes. Taken from
    // threadIdx the compiler cannot know that: it computes them per lane and moves each operand of a bulk copy into a
    // uniform register one after the other (R2UR, ~25 cycles apiece). A warp-wide reduction returns its result in a
    // uniform register, which makes the role, the branches on it and the address arithmetic uniform-datapath work
    // (63 -> 12 R2UR in the kernel; the producers' step 2310 -> 2040 cycles).
#ifdef TAIL_IDLE_WARPS  // A/B builds: TAIL_IDLE_WARPS empty warps between the producers and the clock warp, which moves the
                        // clock warp to another scheduler of the SM (warp index modulo 4)
    const int hw_warp = (int) __reduce_max_sync(0xffffffffu, threadIdx.x >> 5);
    const int warp = hw_warp < PROD ? hw_warp : (hw_warp < PROD + TAIL_IDLE_WARPS ? PROD + 1 : PROD);
#else
    const int warp = (int) __reduce_max_sync(0xffffffffu, threadIdx.x >> 5);
#endif
    const int lane = threadIdx.x & 31;
    const int ch0 = blockIdx.x * 32;
    const int ch = ch0 + lane;
    const bool valid = ch < a.n_ch;
    if (threadIdx.x < PROD * 2) {
        mbar_init(s.bars + threadIdx.x, 1);
    }
    if ((int) threadIdx.x < 2 * ring_blocks) {
        // full[] and empty[] are adjacent. Every lane of the signalling warp arrives for itself (count 32): each lane's own
        // ring accesses are then ordered by its own arrival, which is also what compute-sanitizer's racecheck can follow.
        mbar_init(s.full + threadIdx.x, 32);
    }
    if (threadIdx.x == 0) {
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float *ring_lane = s.ring + lane;

    const sdrm_clock_state st = a.state[ch];
    int history = st.history;
    const int max_carry = a.ring_slots - 2 * kStepRows - kGuard;
    if (history > max_carry) {  // cannot happen unless a previous call already raised the error flag
        history = max_carry;
    }
    // Ring position of working-buffer index i (0 = oldest carried sample) is (i - history) & mask: this call's row n sits at
    // n & mask for every lane, whatever the lane carried over, so the producers' stores are uniform across the warp.
    if (warp == PROD) {
        for (int k = 0; k < history; k++) {
            ring_put(ring_lane, (k - history) & ring_mask, a.ring_slots, a.carry[(size_t) k * a.delay_stride + ch]);
        }
    }
    const int n_blocks = (a.n_rows + kStepRows - 1) / kStepRows;  // steps' worth of rows ("blocks" of SB x 32 rows below)
    const int n_steps = n_blocks + PROD;

    // producer state
    const bool has_dc = PROD == 4;
    float sum = 0.0f;
    float rcp = 0.0f;
    if (has_dc && warp < PROD) {
        sum = a.sums[(size_t) warp * a.delay_stride + ch];
        rcp = __frcp_rn((float) a.dc_length);
    }
    Arrays arrays;
    arrays.rows = group_rows(a, blockIdx.x);
    arrays.line = has_dc ? group_line(a, blockIdx.x, warp < PROD ? warp : 0) : nullptr;
    arrays.dx = has_dc ? group_dx(a, blockIdx.x) : nullptr;
    Cursors cur;
    cur.line = has_dc ? (int) (a.pos_l % a.dc_length) : 0;
    cur.dx_store = has_dc ? (int) (a.pos_x % a.dx_length) : 0;
    cur.dx_load = has_dc ? (int) ((a.pos_x + a.dx_length - (2 * a.dc_length - 2)) % a.dx_length) : 0;
    // a block's delay-line slots may be fetched one block ahead only if the previous block does not write them
    const bool lookahead = !has_dc || a.dc_length >= 2 * kStepRows;
    const int store_slack = has_dc ? max(0, min(2, (a.dc_length - 2 * kStepRows) / kStepRows)) : 0;

    // clock state (warp PROD)
    float *soft = a.soft_out != nullptr ? a.soft_out + (size_t) ch * a.out_stride : nullptr;
    int8_t *hard = a.hard_out != nullptr ? a.hard_out + (size_t) ch * a.out_stride : nullptr;
    const int working_len = history + a.n_rows;
    // clock_recovery_mm.c:94-99: fewer than 8 samples are only buffered (padding lanes never run the loop)
    const bool run_clock = valid && working_len >= 8;
    int ii = 0;
    int oo = 0;
    int previous = 0;
    float mu = st.mu;
    float omega = st.omega;
    float last_sample = st.last_sample;
    bool overflow = false;

    const uint32_t taps_base = smem_u32(s.taps);
    __syncthreads();
#ifdef TAIL_IDLE_WARPS
    if (warp > PROD) {
        return;
    }
#endif
    if (warp < PROD) {
        // ---- producers: PROD pipeline stages in lock step among themselves -------------------------------------------------
        for (int t = 0; t < n_steps - 1; t++) {
            const int b = t - warp;
            if (b >= 0 && b < n_blocks) {
                const int nr = min(kStepRows, a.n_rows - b * kStepRows);
                const Cursors nxt = next_block<SB>(a, cur);
                if (b == 0 || !lookahead) {
                    fetch_block<PROD, SB>(a, s, arrays, warp, lane, b, cur);
                }
                if (lookahead && b + 1 < n_blocks) {
                    fetch_block<PROD, SB>(a, s, arrays, warp, lane, b + 1, nxt);
                }
                mbar_wait(s.bars + warp * 2 + (b & 1), (uint32_t) ((b >> 1) & 1));
                if (has_dc) {
                    // the step's own inputs replace the ones it is about to consume: the input tile goes back to the stage's
                    // delay line (slots are distinct: L >= the step's rows) and, for the first stage, to the group delay line,
                    // which is long enough that these writes never reach what the last stage still has to read
                    if (elect_one()) {  // always the same lane for the full mask: it owns the warp's bulk groups
                        const float *in_tile =
                            warp == 0 ? s.rows + (b & 1) * SB * kTile : s.pipe + ((warp - 1) * 2 + (b & 1)) * SB * kTile;
                        bulk_store_rows(arrays.line, in_tile, cur.line, nr, a.dc_length);
                        if (warp == 0) {
                            bulk_store_rows(arrays.dx, in_tile, cur.dx_store, nr, a.dx_length);
                        }
                        bulk_commit();
                    }
                }
                if (warp == PROD - 1) {
                    // this block's slot of the sample ring still holds block b - ring_blocks (for the first blocks: the samples
                    // carried over from the previous call, or nothing): wait until the clock warp has let go of it. The wait is
                    // over before it starts unless the clock has fallen a whole ring behind. (Probing the barrier at the top
                    // of the step and waiting only after a failed probe was measured: slower, 1.80 -> 1.89 ms.)
                    mbar_wait(s.empty + (b & (ring_blocks - 1)), (uint32_t) ((b / ring_blocks) & 1));
                }
#pragma unroll
                for (int h = 0; h < SB; h++) {
                    const int nr_h = nr - h * kBlockRows;
                    if (nr_h >= kBlockRows) {
                        producer_block<PROD, SB, true, DIVSTEPS>(a, s, warp, lane, b, h, kBlockRows, sum, rcp);
                    } else if (nr_h > 0) {
                        producer_block<PROD, SB, false, DIVSTEPS>(a, s, warp, lane, b, h, nr_h, sum, rcp);
                    }
                }
                if (has_dc && warp < PROD - 1) {
                    fence_async_smem();  // the next stage sends this tile to its delay line with a bulk store
                }
                if (warp == PROD - 1) {
                    mbar_arrive(s.full + (b & (ring_blocks - 1)));  // every lane announces its rows of the block (release, CTA scope)
                }
                cur = nxt;
                // this step's delay-line stores are done before the step ends: their source tile is recycled two steps
                // later and the lines are read again at the earliest one step later
                if (has_dc && elect_one()) {  // always the same lane for the full mask: it owns the warp's bulk groups
                    bulk_wait_step(store_slack);
                }
            }
            if (PROD > 1) {
                producers_barrier(PROD * 32);
            }
        }
    } else {
        // ---- Mueller & Mueller loop: consumer of the sample ring -----------------------------------------------------------
        int seen = 0;  // blocks known to be complete
        // Blocks handed back to the producer so far. Rows before this call's row 0 count as blocks -1, -2, ...: they sit in the
        // ring's last slots and hold the carried samples, so those slots are released like any other, once the lanes have moved
        // past them; the slots in front of them are free from the start. (Block j >= -ring_blocks lives in slot j mod ring_blocks.)
        int released = -ring_blocks;
        {
            const int first_held = __reduce_min_sync(0xffffffffu, (-history - 3) >> kStepShift);  // floor: block of the oldest row
            for (int j = released; j < first_held; j++) {
                mbar_arrive(s.empty + (j & (ring_blocks - 1)));
            }
            released = first_held;
        }
        while (true) {
            if (seen < n_blocks) {
                mbar_wait(s.full + (seen & (ring_blocks - 1)), (uint32_t) ((seen / ring_blocks) & 1));
                seen++;
            }
            const int avail = history + min(a.n_rows, seen * kStepRows);
            // The body is branch-free: the reference's three data-dependent paths (leading zero taps, NaN output, loop
            // update) are computed side by side and selected, so that lanes in different situations stay converged and the
            // iteration is one dependent chain of ~40 operations instead of a sequence of divergent regions.
            // (A variant in which every lane makes the warp's trips with its stores and updates switched off was measured:
            // idle lanes must then be kept away from ring rows that are still being written, and the extra select on the
            // window address costs more than the uniform control flow saves.)
            while (run_clock && !overflow && ii >= 0 && ii + 7 < avail && oo < a.max_out) {
                if (ii - 3 - history < released * kStepRows) {  // the lane went back into rows it had let go of (pathological input)
                    overflow = true;
                    break;
                }
                // imu = (int) rint(mu * 128) (mmse_fir_interpolator.c:189), mu in [0, 1): mu * 128 is exact, and adding 1.5 * 2^23
                // rounds it to an integer (to nearest even, like rint) in the low mantissa bits: one FFMA on the loop's critical
                // path instead of a multiply and a conversion. A NaN mu gives row 0, as the conversion does.
                const int imu = __float_as_int(__fmaf_rn(mu, 128.0f, 12582912.0f)) & 0xff;
                // the tap row through an explicit shared-memory address: with a generic pointer the compiler rebuilds the
                // shared window base (S2UR + ULEA, ~20 cycles on the critical path) in every iteration
                float4 t_lo, t_hi;
                asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t_lo.x), "=f"(t_lo.y), "=f"(t_lo.z), "=f"(t_lo.w) : "r"(taps_base + imu * 32));
                asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+16];" : "=f"(t_hi.x), "=f"(t_hi.y), "=f"(t_hi.z), "=f"(t_hi.w) : "r"(taps_base + imu * 32));
                const float tp[8] = {t_lo.x, t_lo.y, t_lo.z, t_lo.w, t_hi.x, t_hi.y, t_hi.z, t_hi.w};
                // 11 samples from ii - 3 on; the mirror rows make the window contiguous
                const float *win = ring_lane + ((ii - 3 - history) & ring_mask) * 32;
                float v[11];
#pragma unroll
                for (int k = 0; k < 11; k++) {
                    v[k] = win[k * 32];
                }
                // aligned dot product of fir_filter_process_float_single: (ii & 3) earlier samples meet zero taps first.
                // Their products are +-0, or NaN for a non-finite sample; a skipped product is replaced by +0, which leaves
                // the accumulator (+0, or already NaN) unchanged.
                const int lead = ii & 3;
                float acc = 0.0f;
                acc = dot_step<FAST>(acc, lead >= 3 ? v[0] : 0.0f, 0.0f);
                acc = dot_step<FAST>(acc, lead >= 2 ? v[1] : 0.0f, 0.0f);
                acc = dot_step<FAST>(acc, lead >= 1 ? v[2] : 0.0f, 0.0f);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    acc = dot_step<FAST>(acc, v[3 + k], tp[7 - k]);
                }
                const bool nan = isnan(acc);  // clock_recovery_mm.c:107-113: output 0, skip the loop update
                const float out = nan ? 0.0f : acc;
                if (soft != nullptr) soft[oo] = out;
                if (hard != nullptr) {
                    // fsk_demod.c:106, volk_32f_s32f_convert_8i: saturate(rint(x * 127)). The clamped value plus 1.5 * 2^23
                    // carries rint(x) in two's complement in its low mantissa bits (out is never NaN here).
                    const float scaled = fminf(fmaxf(__fmul_rn(out, 127.0f), -128.0f), 127.0f);
                    // (packing four symbols into one 32-bit store was measured: slower, 1.80 -> 1.85 ms)
                    hard[oo] = (int8_t) (__float_as_int(__fadd_rn(scaled, 12582912.0f)) & 0xff);
                }
                const float mm_val = __fsub_rn(__fmul_rn(slice_pm1(last_sample), out), __fmul_rn(slice_pm1(out), last_sample));
                float omega_next = __fadd_rn(omega, __fmul_rn(a.gain_omega, mm_val));
                omega_next = __fadd_rn(a.omega_mid, branchless_clip(__fsub_rn(omega_next, a.omega_mid), a.omega_lim));
                const float mu_next = __fadd_rn(__fadd_rn(mu, omega_next), __fmul_rn(a.gain_mu, mm_val));
                const float whole = floorf(nan ? omega : mu_next);
                previous = ii;
                ii += (int) whole;
                last_sample = nan ? last_sample : out;
                mu = nan ? mu : __fsub_rn(mu_next, whole);
                omega = nan ? omega : omega_next;
                oo++;
            }
            if (seen >= n_blocks) {
                break;  // everything the producers will ever deliver has been looked at
            }
            // Blocks that no active lane will read again go back to the producer: the oldest row a lane still needs is the start
            // of its window, ii - 3 (in rows of this call: minus the carried samples). Lanes that have stopped for good
            // (padding, symbol capacity reached, error) do not hold anything.
            const bool active = run_clock && !overflow && ii >= 0 && oo < a.max_out;
            const int oldest_row = __reduce_min_sync(0xffffffffu, active ? ii - 3 - history : 0x7fffffff);
            const int free_to = min(oldest_row >> kStepShift, seen);  // arithmetic shift: floor, also among the carried samples
            for (int j = released; j < free_to; j++) {
                mbar_arrive(s.empty + (j & (ring_blocks - 1)));  // every lane: its own reads of the block are behind it
            }
            released = max(released, free_to);
        }
    }

    if (warp < PROD) {
        if (has_dc) {
            a.sums[(size_t) warp * a.delay_stride + ch] = sum;
        }
        return;
    }
    if (!valid) {
        return;
    }
    // clock_recovery_mm.c:127-135: what is left of the working buffer is carried into the next call
    long long last_index;
    if (!run_clock) {
        last_index = 0;
    } else {
        last_index = ((unsigned long long) (long long) ii > (unsigned long long) working_len) ? previous : ii;
    }
    long long carried = working_len - last_index;
    if (carried > max_carry || carried < 0 || overflow) {
        atomicOr(a.error_flag, 1);
        carried = carried < 0 ? 0 : (carried > max_carry ? (long long) max_carry : carried);
        last_index = working_len - carried;
    }
    if (oo >= a.max_out && run_clock && ii >= 0 && ii + 7 < working_len) {
        atomicOr(a.error_flag, 2);  // symbol capacity reached with input left over
    }
    for (int k = 0; k < (int) carried; k++) {
        a.carry[(size_t) k * a.delay_stride + ch] = ring_lane[(((int) last_index + k - history) & ring_mask) * 32];
    }
    sdrm_clock_state next;
    next.mu = mu;
    next.omega = omega;
    next.last_sample = last_sample;
    next.history = (int) carried;
    a.state[ch] = next;
    a.out_len[ch] = (uint32_t) oo;
}

// Self test of div_by_length: compares the branch-free form with __fdiv_rn on pseudo-random sums (all exponents the
// fast path accepts, both signs) and counts disagreements, including "redo" requests inside the accepted range.
template <int STEPS>
__global__ void div_selftest_kernel(int length, uint32_t seed, int per_thread, unsigned long long *mismatches) {
    const float length_f = (float) length;
    const float rcp = __frcp_rn(length_f);
    uint32_t x = seed ^ (0x9E3779B9u * (blockIdx.x * blockDim.x + threadIdx.x + 1));
    unsigned long long bad = 0;
    for (int i = 0; i < per_thread; i++) {
        x ^= x << 13;
        x ^= x >> 17;
        x ^= x << 5;
        // exponent in [127 - 58, 127 + 58], random sign and mantissa
        const uint32_t expo = 69u + (x >> 8) % 117u;
        const uint32_t bits = (x & 0x807FFFFFu) | (expo << 23);
        const float sum = __uint_as_float(bits);
        const bool redo = outside_fast_division_range(sum);
        const float fast = div_by_length<STEPS>(sum, length_f, rcp);
        const float exact = __fdiv_rn(sum, length_f);
        if (redo || __float_as_uint(fast) != __float_as_uint(exact)) {
            bad++;
        }
    }
    if (bad != 0) {
        atomicAdd(mismatches, bad);
    }
}

}  // namespace

extern "C" int sdrm_cu_selftest_div(int length, int steps, uint32_t seed, int blocks, int per_thread,
                                    unsigned long long *h_mismatches) {
    unsigned long long *d = nullptr;
    if (cudaMalloc(&d, sizeof(*d)) != cudaSuccess) {
        return -12;
    }
    cudaMemset(d, 0, sizeof(*d));
    if (steps == 1) {
        div_selftest_kernel<1><<<blocks, 256>>>(length, seed, per_thread, d);
    } else if (steps == 0) {
        div_selftest_kernel<0><<<blocks, 256>>>(length, seed, per_thread, d);
    } else {
        div_selftest_kernel<2><<<blocks, 256>>>(length, seed, per_thread, d);
    }
    cudaError_t err = cudaMemcpy(h_mismatches, d, sizeof(*d), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}

extern "C" int sdrm_cu_demod_tail(const sdrm_tail_args *args, void *stream_ptr) {
    if (args->n_ch <= 0) {
        return 0;
    }
    const bool has_dc = args->dc_length != 0;
    if ((args->ring_rows & (args->ring_rows - 1)) != 0 || args->ring_slots < 128 || (args->ring_slots & (args->ring_slots - 1)) != 0 ||
        (args->delay_stride & 31) != 0 || args->n_groups * 32 != (int) args->delay_stride || (size_t) args->n_ch > args->delay_stride ||
        (((uintptr_t) args->rows | (uintptr_t) args->delay) & 127) != 0 ||
        (has_dc && (args->dc_length < kBlockRows || args->dx_length < 2 * args->dc_length - 2 + 8 * kBlockRows))) {
        return -22;
    }
    cudaStream_t stream = (cudaStream_t) stream_ptr;
    const int blocks = (args->n_ch + 31) / 32;
    const int prod = has_dc ? 4 : 1;
    int device = 0;
    cudaGetDevice(&device);
    int optin = 0;
    cudaError_t err = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (err != cudaSuccess) {
        return -(int) err - 1000;
    }
    // two 32-row blocks per step where everything is long enough for 64-row steps and the tiles fit in shared memory
#ifdef TAIL_FORCE_SB  // A/B builds (tools/ab_build.sh)
    int sb = TAIL_FORCE_SB;
#else
    int sb = 2;
#endif
    size_t smem = 0;
    for (; sb >= 1; sb--) {
        const size_t tiles = (size_t) (prod - 1) * 2 + 2 + (size_t) prod * 2 + 2;
        const size_t floats = (size_t) kTapsFloats + (size_t) (args->ring_slots + kMirror) * 32 + tiles * sb * kTile;
        smem = floats * sizeof(float) + ((size_t) prod * 2 + 2 * (size_t) (args->ring_slots / (sb * kBlockRows))) * sizeof(uint64_t);
        const int step_rows = sb * kBlockRows;
        const bool fits = smem <= (size_t) optin && args->ring_slots >= 2 * step_rows + kGuard + 16 &&
                          (!has_dc || (args->dc_length >= step_rows && args->dx_length >= 2 * args->dc_length - 2 + 8 * step_rows));
        if (fits) {
            break;
        }
    }
    if (sb < 1) {
        return -22;
    }
    const int steps_idx = !has_dc ? 3 : (args->div_steps == 1 ? 1 : (args->div_steps == 2 ? 2 : 0));
    typedef void (*tail_kernel_t)(const sdrm_tail_args);
    // [fast][division variant][blocks per step - 1]
    static const tail_kernel_t kernels[2][4][2] = {
        {{demod_tail_kernel<4, 0, false, 1>, demod_tail_kernel<4, 0, false, 2>},
         {demod_tail_kernel<4, 1, false, 1>, demod_tail_kernel<4, 1, false, 2>},
         {demod_tail_kernel<4, 2, false, 1>, demod_tail_kernel<4, 2, false, 2>},
         {demod_tail_kernel<1, 2, false, 1>, demod_tail_kernel<1, 2, false, 2>}},
        {{demod_tail_kernel<4, 0, true, 1>, demod_tail_kernel<4, 0, true, 2>},
         {demod_tail_kernel<4, 1, true, 1>, demod_tail_kernel<4, 1, true, 2>},
         {demod_tail_kernel<4, 2, true, 1>, demod_tail_kernel<4, 2, true, 2>},
         {demod_tail_kernel<1, 2, true, 1>, demod_tail_kernel<1, 2, true, 2>}}};
    tail_kernel_t kernel = kernels[args->fast ? 1 : 0][steps_idx][sb - 1];
    // function attributes once per (variant, device), sized for the largest ring (fir.cu does the same): not per launch
    static std::atomic<bool> configured[16][64];
    const int variant = ((args->fast ? 4 : 0) + steps_idx) * 2 + (sb - 1);
    if (device < 0 || device >= 64 || !configured[variant][device].load(std::memory_order_acquire)) {
        err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
        if (err != cudaSuccess) {
            return -(int) err - 1000;
        }
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);  // see fir.cu
        if (device >= 0 && device < 64) {
            configured[variant][device].store(true, std::memory_order_release);
        }
    }
#ifdef TAIL_IDLE_WARPS
    kernel<<<blocks, (prod + 1 + TAIL_IDLE_WARPS) * 32, smem, stream>>>(*args);
#else
    kernel<<<blocks, (prod + 1) * 32, smem, stream>>>(*args);
#endif
    err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}
