// Microbenchmark: cycles per step of the frequency modulator's serial recurrence p = wrap(p + d) for several
// instruction selections. One warp, increments in shared memory. Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o wrap_chain wrap_chain.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr float kTwoPi = 6.283185307179586476925286766559f;
constexpr int N = 256;     // steps per pass over the shared tile
constexpr int PASSES = 64;

__device__ __forceinline__ float step_pred(float p, float d) {
    const float q = __fadd_rn(p, d);
    const float up = __fadd_rn(q, kTwoPi);
    const float down = __fsub_rn(q, kTwoPi);
    return q < -kTwoPi ? up : (q > kTwoPi ? down : q);
}

// sign-mask form: s = 2pi - |q| is negative exactly when |q| > 2pi, and then -s is the wrapped magnitude
__device__ __forceinline__ float step_mask(float p, float d) {
    const float q = __fadd_rn(p, d);
    const float s = __fsub_rn(kTwoPi, fabsf(q));
    const uint32_t sb = __float_as_uint(s);
    const uint32_t qb = __float_as_uint(q);
    const uint32_t w = (sb & 0x7fffffffu) | (qb & 0x80000000u);
    const uint32_t m = (uint32_t) ((int32_t) sb >> 31);
    return __uint_as_float((qb & ~m) | (w & m));
}

__device__ __forceinline__ float step_plain(float p, float d) { return __fadd_rn(p, d); }

// one predicate: |q| > 2pi selects the wrapped value built with integer ops
__device__ __forceinline__ float step_onepred(float p, float d) {
    const float q = __fadd_rn(p, d);
    const float s = __fsub_rn(fabsf(q), kTwoPi);
    const float w = copysignf(s, q);
    return fabsf(q) > kTwoPi ? w : q;
}

// speculative: the loop state is the pre-wrap sum q; both continuations q + d and wrapped(q) + d are formed before the
// comparison resolves. Returns the wrapped phase for the store through *out_p.
__device__ __forceinline__ float step_spec1(float q, float d, float *out_p) {
    const float s = __fsub_rn(fabsf(q), kTwoPi);
    const float w = copysignf(s, q);
    const bool wrap = fabsf(q) > kTwoPi;
    const float a = __fadd_rn(q, d);
    const float b = __fadd_rn(w, d);
    *out_p = wrap ? w : q;
    return wrap ? b : a;
}

__device__ __forceinline__ float step_spec2(float q, float d, float *out_p) {
    const float t = __fsub_rn(q, kTwoPi);
    const float u = __fadd_rn(q, kTwoPi);
    const float aq = __fadd_rn(q, d);
    const float at = __fadd_rn(t, d);
    const float au = __fadd_rn(u, d);
    const bool hi = q > kTwoPi;
    const bool lo = q < -kTwoPi;
    *out_p = lo ? u : (hi ? t : q);
    return lo ? au : (hi ? at : aq);
}

// predicated add: the offset -copysign(2pi, q) is one LOP3 on q, the comparison one FSETP beside it, and the wrap itself a
// predicated FADD into q's own register: FADD -> (LOP3 | FSETP) -> @P FADD
__device__ __forceinline__ float step_predadd(float p, float d) {
    float q = __fadd_rn(p, d);
    const float off = __uint_as_float((__float_as_uint(q) & 0x80000000u) ^ 0xC0C90FDBu);
    if (fabsf(q) > kTwoPi) {
        q = __fadd_rn(q, off);
    }
    return q;
}

// the same with the predication spelled out
__device__ __forceinline__ float step_predadd_ptx(float p, float d) {
    float q;
    asm("{\n"
        ".reg .pred w;\n"
        ".reg .b32 off, aq;\n"
        "add.rn.f32 %0, %1, %2;\n"
        "lop3.b32 off, %0, 0x80000000, 0xC0C90FDB, 0x6a;\n"  // (a & b) ^ c
        "abs.f32 aq, %0;\n"
        "setp.gt.f32 w, aq, 0f40C90FDB;\n"
        "@w add.rn.f32 %0, %0, off;\n"
        "}\n"
        : "=&f"(q)
        : "f"(p), "f"(d));
    return q;
}

// no predicate at all: u = 2pi - |q| is negative exactly when the phase wraps, and then |u| = |q| - 2pi < |q|. Flipping u's sign bit
// (+ 0x80000000) makes every non-wrapping u, including +0 for |q| == 2pi, an unsigned number above any |q|, so one
// add-and-unsigned-min picks the magnitude: FADD -> FADD -> VIADDMNMX -> LOP3
__device__ __forceinline__ float step_viaddmin(float p, float d) {
    const float q = __fadd_rn(p, d);
    const float u = __fsub_rn(kTwoPi, fabsf(q));
    const uint32_t qb = __float_as_uint(q);
    const uint32_t mag = __viaddmin_u32(__float_as_uint(u), 0x80000000u, qb & 0x7fffffffu);
    return __uint_as_float(mag | (qb & 0x80000000u));
}

template <int V>
__global__ void chain(const float *d_in, float *out, long long *cycles) {
    __shared__ float d[N][32];
    for (int i = threadIdx.x; i < N * 32; i += 32) {
        d[i / 32][i % 32] = d_in[i];
    }
    __syncwarp();
    float p = 0.0f;
    long long t0 = clock64();
    float *col = &d[0][threadIdx.x];
    if (V >= 4 && V < 6) {
        p = col[0];  // pre-wrap state: q_0 = 0 + d_0
    }
    for (int pass = 0; pass < PASSES; pass++) {
        for (int i0 = 0; i0 < N; i0 += 32) {
            float *c = col + i0 * 32;
#pragma unroll
            for (int i = 0; i < 32; i++) {
                if (V >= 6) {
                    const float x = c[i * 32];
                    p = V == 6 ? step_predadd(p, x) : V == 7 ? step_predadd_ptx(p, x) : step_viaddmin(p, x);
                    c[i * 32] = p;
                } else if (V < 4) {
                    const float x = c[i * 32];
                    p = V == 0 ? step_pred(p, x) : V == 1 ? step_mask(p, x) : V == 2 ? step_plain(p, x) : step_onepred(p, x);
                    c[i * 32] = p;
                } else {
                    // next increment (wraps to the start of the tile; the values only need to be consistent across variants)
                    const int nxt = (i0 + i + 1) % N;
                    const float x = col[nxt * 32];
                    float ph;
                    p = V == 4 ? step_spec1(p, x, &ph) : step_spec2(p, x, &ph);
                    c[i * 32] = ph;
                }
            }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = p + d[N / 2][threadIdx.x];
    if (threadIdx.x == 0) {
        *cycles = t1 - t0;
    }
}

// onepred over a quad layout: [step / 4][lane][4], so that one LDS.128 / STS.128 moves four steps of a lane. One warp pays
// about 5 cycles of issue per LDS.32 and 3 per STS.32 (lsu_issue.cu); measured on B200 this is NOT faster (23.75 cycles per step
// against 20.52): the wrap chain itself (add, add, sign merge, select) is the 20 cycles, the memory instructions fit its stalls.
__global__ void chain_quad(const float *d_in, float *out, long long *cycles) {
    __shared__ __align__(16) float d[N * 32];
    for (int i = threadIdx.x; i < N * 32; i += 32) {
        const int step = i / 32, lane = i % 32;
        d[((step >> 2) * 32 + lane) * 4 + (step & 3)] = d_in[i];
    }
    __syncwarp();
    float p = 0.0f;
    long long t0 = clock64();
    float4 *col = reinterpret_cast<float4 *>(d) + threadIdx.x;
    for (int pass = 0; pass < PASSES; pass++) {
        for (int i0 = 0; i0 < N / 4; i0 += 8) {
            float4 *c = col + i0 * 32;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float4 x = c[i * 32];
                x.x = p = step_onepred(p, x.x);
                x.y = p = step_onepred(p, x.y);
                x.z = p = step_onepred(p, x.z);
                x.w = p = step_onepred(p, x.w);
                c[i * 32] = x;
            }
        }
    }
    long long t1 = clock64();
    const int half = N / 2;
    out[threadIdx.x] = p + d[((half >> 2) * 32 + threadIdx.x) * 4 + (half & 3)];
    if (threadIdx.x == 0) {
        *cycles = t1 - t0;
    }
}

int main() {
    float *h = new float[N * 32];
    uint32_t r = 12345;
    for (int i = 0; i < N * 32; i++) {
        r = r * 1664525u + 1013904223u;
        h[i] = ((int) (r >> 8) % 20000 - 10000) * 1e-4f;
    }
    float *d_in, *out;
    long long *cyc;
    cudaMalloc(&d_in, N * 32 * 4);
    cudaMalloc(&out, 32 * 4 * 4);
    cudaMalloc(&cyc, 8);
    cudaMemcpy(d_in, h, N * 32 * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(chain<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 0);
    float res[9][32];
    long long c[9];
    for (int rep = 0; rep < 2; rep++) {
        chain<0><<<1, 32>>>(d_in, out, cyc);
        cudaMemcpy(&c[0], cyc, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(res[0], out, 128, cudaMemcpyDeviceToHost);
        chain<1><<<1, 32>>>(d_in, out, cyc);
        cudaMemcpy(&c[1], cyc, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(res[1], out, 128, cudaMemcpyDeviceToHost);
        chain<2><<<1, 32>>>(d_in, out, cyc);
        cudaMemcpy(&c[2], cyc, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(res[2], out, 128, cudaMemcpyDeviceToHost);
        chain<3><<<1, 32>>>(d_in, out, cyc);
        cudaMemcpy(&c[3], cyc, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(res[3], out, 128, cudaMemcpyDeviceToHost);
        chain<4><<<1, 32>>>(d_in, out, cyc);
        cudaMemcpy(&c[4], cyc, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(res[4], out, 128, cudaMemcpyDeviceToHost);
        chain<5><<<1, 32>>>(d_in, out, cyc);
        cudaMemcpy(&c[5], cyc, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(res[5], out, 128, cudaMemcpyDeviceToHost);
        chain<6><<<1, 32>>>(d_in, out, cyc);
        cudaMemcpy(&c[6], cyc, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(res[6], out, 128, cudaMemcpyDeviceToHost);
        chain<7><<<1, 32>>>(d_in, out, cyc);
        cudaMemcpy(&c[7], cyc, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(res[7], out, 128, cudaMemcpyDeviceToHost);
        chain<8><<<1, 32>>>(d_in, out, cyc);
        cudaMemcpy(&c[8], cyc, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(res[8], out, 128, cudaMemcpyDeviceToHost);
    }
    cudaError_t e = cudaDeviceSynchronize();
    int same1 = 1, same3 = 1;
    for (int i = 0; i < 32; i++) {
        same1 &= res[0][i] == res[1][i];
        same3 &= res[0][i] == res[3][i];
    }
    int same6 = 1, same7 = 1;
    for (int i = 0; i < 32; i++) {
        same6 &= res[0][i] == res[6][i];
        same7 &= res[0][i] == res[7][i];
    }
    int same8 = 1;
    for (int i = 0; i < 32; i++) {
        same8 &= res[0][i] == res[8][i];
    }
    {
        long long cq = 0;
        float rq[32];
        for (int rep = 0; rep < 2; rep++) {
            chain_quad<<<1, 32>>>(d_in, out, cyc);
            cudaMemcpy(&cq, cyc, 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(rq, out, 128, cudaMemcpyDeviceToHost);
        }
        int sameq = 1;
        for (int i = 0; i < 32; i++) {
            sameq &= rq[i] == res[3][i];
        }
        printf("onepred, quad layout (LDS.128 / STS.128 per 4 steps) %.2f (same=%d) cycles/step\n", cq / (double) (N * PASSES), sameq);
    }
    printf("viaddmin %.2f (same=%d) cycles/step\n", c[8] / (double) (N * PASSES), same8);
    printf("predadd %.2f (same=%d)  predadd_ptx %.2f (same=%d) cycles/step\n", c[6] / (double) (N * PASSES), same6,
           c[7] / (double) (N * PASSES), same7);
    printf("spec1 %.2f spec2 %.2f cycles/step (state differs by design; res %g %g %g)\n", c[4] / (double) (N * PASSES),
           c[5] / (double) (N * PASSES), res[0][0], res[4][0], res[5][0]);
    printf("err=%d pred %.2f  mask %.2f (same=%d)  plain %.2f  onepred %.2f (same=%d) cycles/step\n", (int) e, c[0] / (double) (N * PASSES),
           c[1] / (double) (N * PASSES), same1, c[2] / (double) (N * PASSES), c[3] / (double) (N * PASSES), same3);
    return 0;
}
