// Serial tail of the demod chain for sm_100a: DC blocker and Mueller & Mueller clock recovery (+ int8 conversion).
//
// Both stages are feedback loops whose float rounding is part of the result (running sums that are never
// reset, a timing loop that steers its own sample positions), so neither is split within a stream: one lane owns
// one channel for the whole call and walks it in time order. Parallelism is across channels only; data lives in
// the time-major TC ring (see sdrm_cuda.h) so that the 32 lanes of a warp touch 32 consecutive floats.
//
// Mirrors, operation by operation:
//   dc blocker  reference src/dsp/dc_blocker.c:52-64 (moving_average_process), :105-119 (dc_blocker_process)
//   clock loop  reference src/dsp/clock_recovery_mm.c:78-139, src/dsp/mmse_fir_interpolator.c:188-191,
//               src/dsp/fir_filter.c:116-121 (aligned 8-tap dot product with up to 3 leading zero taps)
//   int8        reference src/dsp/fsk_demod.c:106 (volk_32f_s32f_convert_8i generic: scale, saturate, rintf)

#include <cuda_runtime.h>
#include <stdint.h>

#include "sdrm_cuda.h"

namespace {

// y = x - x[n-L] + y_prev ; out = y / L        (true IEEE division, two separately rounded adds)
__device__ __forceinline__ float ma_step(float x, float delayed, float &sum, float length_f) {
    sum = __fadd_rn(__fsub_rn(x, delayed), sum);
    return __fdiv_rn(sum, length_f);
}

// Channel ch's column of the ring: ring[row][tc_stride] (per-block handles) or, grouped, the layout the decimating FIR writes
// for the batch, ring[ch / 32][row][32]. Returns the column's first element; rows are `stride` floats apart.
__device__ __forceinline__ float *ring_column(float *ring, size_t tc_stride, int ring_rows, int ch, int grouped, size_t *stride) {
    if (grouped) {
        *stride = 32;
        return ring + (size_t) (ch >> 5) * (size_t) ring_rows * 32 + (ch & 31);
    }
    *stride = tc_stride;
    return ring + ch;
}

__global__ void dc_blocker_kernel(float *ring, size_t tc_stride, int ring_mask, long long head, int n_rows, int n_ch,
                                  int length, float *__restrict__ delay, float *__restrict__ sums, int pos_l, int pos_x,
                                  int grouped) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= n_ch) {
        return;
    }
    const float length_f = (float) length;
    const int len_x = 2 * length - 2;
    // delay lines, each [slots][n_ch]: four moving averages (L slots) and the group-delay line (2L-2 slots)
    float *d0 = delay + ch;
    float *d1 = d0 + (size_t) length * n_ch;
    float *d2 = d1 + (size_t) length * n_ch;
    float *d3 = d2 + (size_t) length * n_ch;
    float *dx = d3 + (size_t) length * n_ch;
    float s0 = sums[ch];
    float s1 = sums[n_ch + ch];
    float s2 = sums[2 * n_ch + ch];
    float s3 = sums[3 * n_ch + ch];

    int pl = pos_l;
    int px = pos_x;
    size_t stride;
    float *column = ring_column(ring, tc_stride, ring_mask + 1, ch, grouped, &stride);
    for (int n = 0; n < n_rows; n++) {
        float *cell = column + (size_t) ((head + n) & ring_mask) * stride;
        const size_t ol = (size_t) pl * n_ch;
        const size_t ox = (size_t) px * n_ch;
        const float x = *cell;
        // slot pl holds the stage input of L samples ago; read it, then overwrite it with the current one
        const float o0 = ma_step(x, d0[ol], s0, length_f);
        d0[ol] = x;
        const float o1 = ma_step(o0, d1[ol], s1, length_f);
        d1[ol] = o0;
        const float o2 = ma_step(o1, d2[ol], s2, length_f);
        d2[ol] = o1;
        const float o3 = ma_step(o2, d3[ol], s3, length_f);
        d3[ol] = o2;
        const float delayed_x = dx[ox];  // x[n - (2L - 2)]
        dx[ox] = x;
        *cell = __fsub_rn(delayed_x, o3);
        pl = (pl + 1 == length) ? 0 : pl + 1;
        px = (px + 1 == len_x) ? 0 : px + 1;
    }
    sums[ch] = s0;
    sums[n_ch + ch] = s1;
    sums[2 * n_ch + ch] = s2;
    sums[3 * n_ch + ch] = s3;
}

__device__ __forceinline__ float slice_pm1(float x) { return x < 0.0f ? -1.0f : 1.0f; }

__device__ __forceinline__ float branchless_clip(float x, float clip) {
    return __fmul_rn(0.5f, __fsub_rn(fabsf(__fadd_rn(x, clip)), fabsf(__fsub_rn(x, clip))));
}

__global__ void clock_mm_kernel(const sdrm_clock_args a) {
    __shared__ float taps_s[129 * 8];
    for (int i = threadIdx.x; i < 129 * 8; i += blockDim.x) {
        taps_s[i] = a.mmse_taps[i];
    }
    __syncthreads();
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= a.n_ch) {
        return;
    }
    sdrm_clock_state st = a.state[ch];
    const long long working_len = (long long) st.history + a.n_rows;
    const long long base = a.head - st.history;  // absolute ring row of working_buffer[0]
    const int ring_mask = a.ring_rows - 1;
    if (working_len < 8) {
        st.history = (int) working_len;
        a.state[ch] = st;
        a.out_len[ch] = 0;
        return;
    }
    const unsigned long long max_index = (unsigned long long) (working_len - 7);
    size_t stride;
    const float *col = ring_column(const_cast<float *>(a.ring), a.tc_stride, a.ring_rows, ch, a.grouped, &stride);
    float *soft = a.soft_out != nullptr ? a.soft_out + (size_t) ch * a.out_stride : nullptr;
    int8_t *hard = a.hard_out != nullptr ? a.hard_out + (size_t) ch * a.out_stride : nullptr;

    int ii = 0;
    int oo = 0;
    int previous = 0;
    float mu = st.mu;
    float omega = st.omega;
    float last_sample = st.last_sample;
    // the reference compares the int index against size_t bounds: a negative index ends the loop
    while ((unsigned long long) (long long) ii < max_index && oo < a.max_out) {
        const int imu = __float2int_rn(__fmul_rn(mu, 128.0f));
        const float *t = taps_s + imu * 8;
        // fir_filter_process_float_single: dot product starts at the 16-byte aligned address at or below the
        // window, the (ii & 3) samples in front meet zero taps
        const int lead = ii & 3;
        float acc = 0.0f;
        for (int k = lead; k > 0; k--) {
            const float v = col[(size_t) ((base + ii - k) & ring_mask) * stride];
            acc = a.fast ? __fmaf_rn(v, 0.0f, acc) : __fadd_rn(acc, __fmul_rn(v, 0.0f));
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float v = col[(size_t) ((base + ii + k) & ring_mask) * stride];
            acc = a.fast ? __fmaf_rn(v, t[7 - k], acc) : __fadd_rn(acc, __fmul_rn(v, t[7 - k]));
        }
        float out = acc;
        if (isnan(out)) {
            out = 0.0f;
            if (soft != nullptr) soft[oo] = out;
            if (hard != nullptr) hard[oo] = 0;
            previous = ii;
            ii += (int) floorf(omega);
            oo++;
            continue;
        }
        if (soft != nullptr) soft[oo] = out;
        if (hard != nullptr) {
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
    const long long last_index = ((unsigned long long) (long long) ii > (unsigned long long) working_len) ? previous : ii;
    const long long history = working_len - last_index;
    if (history > a.max_history || history < 0) {
        atomicOr(a.error_flag, 1);
    }
    st.mu = mu;
    st.omega = omega;
    st.last_sample = last_sample;
    st.history = (int) history;
    a.state[ch] = st;
    a.out_len[ch] = (uint32_t) oo;
}

}  // namespace

extern "C" int sdrm_cu_dc_blocker(float *ring, size_t tc_stride, int ring_rows, long long head, int n_rows, int n_ch,
                                  int length, float *delay, float *sums, int pos_l, int pos_x, int grouped, void *stream_ptr) {
    if (n_rows <= 0 || n_ch <= 0) {
        return 0;
    }
    if (length < 2 || (ring_rows & (ring_rows - 1)) != 0) {
        return -22;
    }
    const int threads = 32;
    dc_blocker_kernel<<<(n_ch + threads - 1) / threads, threads, 0, (cudaStream_t) stream_ptr>>>(
        ring, tc_stride, ring_rows - 1, head, n_rows, n_ch, length, delay, sums, pos_l, pos_x, grouped);
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}

extern "C" int sdrm_cu_clock_mm(const sdrm_clock_args *args, void *stream_ptr) {
    if (args->n_ch <= 0) {
        return 0;
    }
    if ((args->ring_rows & (args->ring_rows - 1)) != 0) {
        return -22;
    }
    const int threads = 32;
    clock_mm_kernel<<<(args->n_ch + threads - 1) / threads, threads, 0, (cudaStream_t) stream_ptr>>>(*args);
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}
