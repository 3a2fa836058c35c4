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
//     __syncthreads()
//
// Inside a stage and block, only the running sum y[n] = d[n] + y[n-1] is serial (one FADD per row); the 32 subtractions
// and the 32 divisions by L are independent and unrolled, so a warp has plenty of ILP. The division by the constant L
// is done branch-free (two Markstein corrections of sum * RN(1/L), checked against IEEE division by
// sdrm_cu_selftest_div); values outside a safe exponent range redo their block with __fdiv_rn.
// What the reference carries from call to call in its working buffer (clock_recovery_mm.c:127-135) is saved from the
// shared-memory ring into a small per-channel carry array at the end of the call and reloaded at the start of the next.
//
// Mirrors: reference src/dsp/dc_blocker.c:52-64,105-119; src/dsp/clock_recovery_mm.c:78-139;
// src/dsp/mmse_fir_interpolator.c:188-191; src/dsp/fir_filter.c:116-121; src/dsp/fsk_demod.c:106.

#include <cuda_runtime.h>
#include <stdint.h>

#include "sdrm_cuda.h"

namespace {

constexpr int kBlockRows = 32;  // rows per pipeline step
constexpr int kGuard = 16;      // slack between what a lane may lag behind and the size of its ring

__device__ __forceinline__ float slice_pm1(float x) { return x < 0.0f ? -1.0f : 1.0f; }

__device__ __forceinline__ float branchless_clip(float x, float clip) {
    return __fmul_rn(0.5f, __fsub_rn(fabsf(__fadd_rn(x, clip)), fabsf(__fsub_rn(x, clip))));
}

__device__ __forceinline__ float dot_step(bool fast, float acc, float v, float t) {
    return fast ? __fmaf_rn(v, t, acc) : __fadd_rn(acc, __fmul_rn(v, t));
}

// sum / L, correctly rounded, for a constant L with rcp = RN(1/L): two Markstein corrections of q0 = RN(sum * rcp).
// `redo` is raised outside the exponent range where the residuals are exact; the caller then uses __fdiv_rn.
__device__ __forceinline__ float div_by_length(float sum, float length_f, float rcp, bool &redo) {
    const float q0 = __fmul_rn(sum, rcp);
    const float e0 = __fmaf_rn(-q0, length_f, sum);
    const float q1 = __fmaf_rn(e0, rcp, q0);
    const float e1 = __fmaf_rn(-q1, length_f, sum);
    const float q2 = __fmaf_rn(e1, rcp, q1);
    const float mag = fabsf(sum);
    redo |= !(mag < 1.0e18f) || (mag < 1.0e-18f && sum != 0.0f);
    return sum == 0.0f ? sum : q2;  // keeps the sign of a zero sum
}

__device__ __forceinline__ void cp_async_f32(float *smem_dst, const float *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t) __cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
                 : "memory");
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }

template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// What a producer warp needs from global memory for one block, staged into its shared-memory fetch buffers with
// cp.async so that all of a block's loads are in flight together (ptxas otherwise sinks each register load next to its
// use and pays one memory round trip per row) and, when the delay lines allow it, one block ahead of the arithmetic:
//   warp 0            the block's rows of the TC ring,
//   every MA warp     the stage's own inputs of L rows ago (its delay line),
//   last MA warp      x[n - (2L - 2)] from the group delay line.
struct FetchBuffers {
    float *rows;  // [2][32][32]   (warp 0)
    float *line;  // [2][32][32]   (this warp's delay line values)
    float *dx;    // [2][32][32]   (last producer warp)
};

template <int PROD>
__device__ __forceinline__ void fetch_block(const sdrm_tail_args &a, const FetchBuffers &f, int warp, int lane, int ch, int b,
                                            const float *line, const float *dx) {
    const bool has_dc = PROD == 4;
    const int row0 = b * kBlockRows;
    const int nr = min(kBlockRows, a.n_rows - row0);
    const int parity = (b & 1) * kBlockRows * 32;
    if (warp == 0) {
        const int tc_mask = a.ring_rows - 1;
#pragma unroll 8
        for (int r = 0; r < nr; r++) {
            cp_async_f32(f.rows + parity + r * 32 + lane, a.rows + (size_t) ((a.head + row0 + r) & tc_mask) * a.tc_stride + ch);
        }
    }
    if (has_dc) {
        int slot = (int) (((long long) a.pos_l + row0) % a.dc_length);
#pragma unroll 8
        for (int r = 0; r < nr; r++) {
            cp_async_f32(f.line + parity + r * 32 + lane, line + (size_t) slot * a.delay_stride);
            slot = slot + 1 == a.dc_length ? 0 : slot + 1;
        }
        if (warp == PROD - 1) {
            const int len_x = a.dx_length;
            int sx = (int) (((long long) a.pos_x + row0 + len_x - (2 * a.dc_length - 2)) % len_x);
#pragma unroll 8
            for (int r = 0; r < nr; r++) {
                cp_async_f32(f.dx + parity + r * 32 + lane, dx + (size_t) sx * a.delay_stride);
                sx = sx + 1 == len_x ? 0 : sx + 1;
            }
        }
    }
    cp_async_commit();
}

// One producer warp's arithmetic for one block whose inputs are already staged: moving average
//   y = in - in[n-L] + y_prev ; out = y / L   (dc_blocker.c:52-64)
// then hand-over to the next stage's buffer or (last stage) x[n-(2L-2)] - y4 into the clock's sample ring
// (dc_blocker.c:110-114). FULL blocks carry no per-row guards, so the 32 rows form one basic block.
template <int PROD, bool FULL>
__device__ __forceinline__ void producer_block(const sdrm_tail_args &a, const FetchBuffers &f, int warp, int lane, bool valid,
                                               int b, int row0, int nr, int history, float &sum, float rcp, float *line,
                                               float *dx, float *pipe_s, float *ring_lane) {
    const bool has_dc = PROD == 4;
    const int ring_mask = a.ring_slots - 1;
    const int parity = (b & 1) * kBlockRows * 32;
    const float *src = warp == 0 ? f.rows + parity + lane : pipe_s + ((size_t) (warp - 1) * 2 + (b & 1)) * kBlockRows * 32 + lane;
    float *dst = warp < PROD - 1 ? pipe_s + ((size_t) warp * 2 + (b & 1)) * kBlockRows * 32 + lane : nullptr;
    if (!has_dc) {
#pragma unroll
        for (int r = 0; r < kBlockRows; r++) {
            if (FULL || r < nr) {
                ring_lane[((history + row0 + r) & ring_mask) * 32] = src[r * 32];
            }
        }
        return;
    }
    const float length_f = (float) a.dc_length;
    // only y[] lives in registers across the block; inputs are re-read from shared memory where they are needed again
    float y[kBlockRows];
#pragma unroll
    for (int r = 0; r < kBlockRows; r++) {
        if (FULL || r < nr) {
            y[r] = __fsub_rn(src[r * 32], f.line[parity + r * 32 + lane]);
        }
    }
    if (valid) {
        // the block's own inputs replace the ones it just consumed (slots are distinct: L >= 32)
        int slot = (int) (((long long) a.pos_l + row0) % a.dc_length);
#pragma unroll
        for (int r = 0; r < kBlockRows; r++) {
            if (FULL || r < nr) {
                line[(size_t) slot * a.delay_stride] = src[r * 32];
            }
            slot = slot + 1 == a.dc_length ? 0 : slot + 1;
        }
        if (warp == 0) {
            // group delay line: x[n] goes in now, the last warp reads x[n - (2L - 2)] three steps later; the line is
            // 2L - 2 + 256 slots long so that the newest writes never reach the oldest reads
            const int len_x = a.dx_length;
            int sx = (int) (((long long) a.pos_x + row0) % len_x);
#pragma unroll
            for (int r = 0; r < kBlockRows; r++) {
                if (FULL || r < nr) {
                    dx[(size_t) sx * a.delay_stride] = src[r * 32];
                }
                sx = sx + 1 == len_x ? 0 : sx + 1;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kBlockRows; r++) {
        if (FULL || r < nr) {
            sum = __fadd_rn(y[r], sum);
            y[r] = sum;
        }
    }
    bool redo = false;
#pragma unroll
    for (int r = 0; r < kBlockRows; r++) {
        if (FULL || r < nr) {
            const float q = div_by_length(y[r], length_f, rcp, redo);
            if (warp < PROD - 1) {
                dst[r * 32] = q;
            } else {
                ring_lane[((history + row0 + r) & ring_mask) * 32] = __fsub_rn(f.dx[parity + r * 32 + lane], q);
            }
        }
    }
    if (__any_sync(0xffffffffu, redo)) {
#pragma unroll
        for (int r = 0; r < kBlockRows; r++) {
            if (FULL || r < nr) {
                const float q = __fdiv_rn(y[r], length_f);
                if (warp < PROD - 1) {
                    dst[r * 32] = q;
                } else {
                    ring_lane[((history + row0 + r) & ring_mask) * 32] = __fsub_rn(f.dx[parity + r * 32 + lane], q);
                }
            }
        }
    }
}

// PROD producer warps (4 moving averages, or 1 plain copier when the dc blocker is off) + 1 clock warp.
template <int PROD>
__global__ void __launch_bounds__((PROD + 1) * 32) demod_tail_kernel(const sdrm_tail_args a) {
    extern __shared__ float smem[];
    float *taps_s = smem;                                 // 129 * 8
    float *ring_s = smem + 129 * 8 + 8;                   // [ring_slots][32]
    float *pipe_s = ring_s + (size_t) a.ring_slots * 32;  // [PROD - 1][2][32 rows][32 lanes]
    float *fetch_s = pipe_s + (size_t) (PROD - 1) * 2 * kBlockRows * 32;  // [PROD + 2][2][32][32]
    const int ring_mask = a.ring_slots - 1;
    for (int i = threadIdx.x; i < 129 * 8; i += blockDim.x) {
        taps_s[i] = a.mmse_taps[i];
    }
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ch_raw = blockIdx.x * 32 + lane;
    const bool valid = ch_raw < a.n_ch;
    const int ch = valid ? ch_raw : a.n_ch - 1;  // idle lanes shadow the last channel and never write
    float *ring_lane = ring_s + lane;

    const sdrm_clock_state st = a.state[ch];
    int history = st.history;
    const int max_carry = a.ring_slots - 2 * kBlockRows - kGuard;
    if (history > max_carry) {  // cannot happen unless a previous call already raised the error flag
        history = max_carry;
    }
    if (warp == PROD) {
        for (int k = 0; k < history; k++) {
            ring_lane[(k & ring_mask) * 32] = a.carry[(size_t) k * a.delay_stride + ch];
        }
    }
    const int n_blocks = (a.n_rows + kBlockRows - 1) / kBlockRows;
    const int n_steps = n_blocks + PROD;
    const int tc_mask = a.ring_rows - 1;

    // producer state
    const bool has_dc = PROD == 4;
    float sum = 0.0f;
    float rcp = 0.0f;
    float *line = nullptr;
    const int len_x = a.dx_length;
    float *dx = nullptr;
    if (has_dc && warp < PROD) {
        sum = a.sums[(size_t) warp * a.delay_stride + ch];
        rcp = __frcp_rn((float) a.dc_length);
        line = a.delay + (size_t) warp * a.dc_length * a.delay_stride + ch;
        dx = a.delay + (size_t) 4 * a.dc_length * a.delay_stride + ch;
    }

    // clock state (warp PROD)
    float *soft = a.soft_out != nullptr ? a.soft_out + (size_t) ch * a.out_stride : nullptr;
    int8_t *hard = a.hard_out != nullptr ? a.hard_out + (size_t) ch * a.out_stride : nullptr;
    const int working_len = history + a.n_rows;
    // clock_recovery_mm.c:94-99: fewer than 8 samples are only buffered (idle lanes never run the loop)
    const bool run_clock = valid && working_len >= 8;
    int ii = 0;
    int oo = 0;
    int previous = 0;
    float mu = st.mu;
    float omega = st.omega;
    float last_sample = st.last_sample;
    bool overflow = false;

    FetchBuffers fb;
    fb.rows = fetch_s;
    fb.line = fetch_s + (size_t) (1 + (warp < PROD ? warp : 0)) * 2 * kBlockRows * 32;
    fb.dx = fetch_s + (size_t) (PROD + 1) * 2 * kBlockRows * 32;
    // a block's delay-line slots may be fetched one block ahead only if the previous block does not write them
    const bool lookahead = !has_dc || a.dc_length >= 2 * kBlockRows;

    __syncthreads();
    for (int t = 0; t < n_steps; t++) {
        if (warp < PROD) {
            const int b = t - warp;
            if (b >= 0 && b < n_blocks) {
                const int row0 = b * kBlockRows;
                const int nr = min(kBlockRows, a.n_rows - row0);
                if (b == 0 || !lookahead) {
                    fetch_block<PROD>(a, fb, warp, lane, ch, b, line, dx);
                }
                if (lookahead && b + 1 < n_blocks) {
                    fetch_block<PROD>(a, fb, warp, lane, ch, b + 1, line, dx);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                if (nr == kBlockRows) {
                    producer_block<PROD, true>(a, fb, warp, lane, valid, b, row0, nr, history, sum, rcp, line, dx, pipe_s, ring_lane);
                } else {
                    producer_block<PROD, false>(a, fb, warp, lane, valid, b, row0, nr, history, sum, rcp, line, dx, pipe_s, ring_lane);
                }
            }
        } else {
            // Mueller & Mueller loop over everything the last producer finished before this step
            const int done_blocks = t - PROD + 1;
            const int avail = history + (done_blocks <= 0 ? 0 : min(a.n_rows, done_blocks * kBlockRows));
            while (run_clock && ii >= 0 && ii + 7 < avail && oo < a.max_out) {
                if (avail + kBlockRows - (ii - 3) > a.ring_slots) {  // the lane fell behind its ring (pathological input)
                    overflow = true;
                    break;
                }
                const int imu = __float2int_rn(__fmul_rn(mu, 128.0f));
                const float *tp = taps_s + imu * 8;
                // aligned dot product of fir_filter_process_float_single: (ii & 3) earlier samples meet zero taps first
                const int lead = ii & 3;
                float acc = 0.0f;
                if (lead >= 3) acc = dot_step(a.fast, acc, ring_lane[((ii - 3) & ring_mask) * 32], 0.0f);
                if (lead >= 2) acc = dot_step(a.fast, acc, ring_lane[((ii - 2) & ring_mask) * 32], 0.0f);
                if (lead >= 1) acc = dot_step(a.fast, acc, ring_lane[((ii - 1) & ring_mask) * 32], 0.0f);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    acc = dot_step(a.fast, acc, ring_lane[((ii + k) & ring_mask) * 32], tp[7 - k]);
                }
                float out = acc;
                if (isnan(out)) {  // clock_recovery_mm.c:107-113
                    out = 0.0f;
                    if (valid && soft != nullptr) soft[oo] = out;
                    if (valid && hard != nullptr) hard[oo] = 0;
                    previous = ii;
                    ii += (int) floorf(omega);
                    oo++;
                    continue;
                }
                if (valid && soft != nullptr) soft[oo] = out;
                if (valid && hard != nullptr) {
                    const float scaled = __fmul_rn(out, 127.0f);
                    hard[oo] = scaled > 127.0f ? (int8_t) 127 : (scaled < -128.0f ? (int8_t) -128 : (int8_t) __float2int_rn(scaled));
                }
                const float mm_val = __fsub_rn(__fmul_rn(slice_pm1(last_sample), out), __fmul_rn(slice_pm1(out), last_sample));
                last_sample = out;
                previous = ii;
                omega = __fadd_rn(omega, __fmul_rn(a.gain_omega, mm_val));
                omega = __fadd_rn(a.omega_mid, branchless_clip(__fsub_rn(omega, a.omega_mid), a.omega_lim));
                mu = __fadd_rn(__fadd_rn(mu, omega), __fmul_rn(a.gain_mu, mm_val));
                const float whole = floorf(mu);
                ii += (int) whole;
                mu = __fsub_rn(mu, whole);
                oo++;
            }
        }
        __syncthreads();
    }

    if (!valid) {
        return;
    }
    if (warp < PROD) {
        if (has_dc) {
            a.sums[(size_t) warp * a.delay_stride + ch] = sum;
        }
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
        a.carry[(size_t) k * a.delay_stride + ch] = ring_lane[(((int) last_index + k) & ring_mask) * 32];
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
        bool redo = false;
        const float fast = div_by_length(sum, length_f, rcp, redo);
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

extern "C" int sdrm_cu_selftest_div(int length, uint32_t seed, int blocks, int per_thread, unsigned long long *h_mismatches) {
    unsigned long long *d = nullptr;
    if (cudaMalloc(&d, sizeof(*d)) != cudaSuccess) {
        return -12;
    }
    cudaMemset(d, 0, sizeof(*d));
    div_selftest_kernel<<<blocks, 256>>>(length, seed, per_thread, d);
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
        (has_dc && (args->dc_length < kBlockRows || args->dx_length < 2 * args->dc_length - 2 + 8 * kBlockRows))) {
        return -22;
    }
    cudaStream_t stream = (cudaStream_t) stream_ptr;
    const int blocks = (args->n_ch + 31) / 32;
    const int prod = has_dc ? 4 : 1;
    const size_t smem = (129 * 8 + 8 + (size_t) args->ring_slots * 32 + (size_t) (prod - 1) * 2 * kBlockRows * 32 +
                         (size_t) (prod + 2) * 2 * kBlockRows * 32) * sizeof(float);
    void (*kernel)(const sdrm_tail_args) = has_dc ? demod_tail_kernel<4> : demod_tail_kernel<1>;
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (err != cudaSuccess) {
        return -(int) err - 1000;
    }
    kernel<<<blocks, (prod + 1) * 32, smem, stream>>>(*args);
    err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}
