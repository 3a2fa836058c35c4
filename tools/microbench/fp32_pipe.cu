// Microbenchmark: FP32 pipe throughput on B200 for the instruction mixes the FIR kernel can use.
// Prints lane-ops/clk/SM for FFMA, FFMA2, FMUL+FADD (scalar, unfused) and FFMA2-pair (exact-mode emulation).
#include <cstdio>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) pipe_kernel(float2 *out, int iters, float2 one, float2 nz, float2 h) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float2 x = a[(i + 3) % 8];  // data dependent so that nothing is loop invariant
                if (MODE == 0) {  // scalar FFMA x2
                    a[i].x = __fmaf_rn(x.x, h.x, a[i].x);
                    a[i].y = __fmaf_rn(x.y, h.y, a[i].y);
                } else if (MODE == 1) {  // FFMA2
                    a[i] = __ffma2_rn(x, h, a[i]);
                } else if (MODE == 2) {  // scalar FMUL + FADD
                    a[i].x = __fadd_rn(a[i].x, __fmul_rn(x.x, h.x));
                    a[i].y = __fadd_rn(a[i].y, __fmul_rn(x.y, h.y));
                } else if (MODE == 3) {  // exact mode: two FFMA2
                    float2 p = __ffma2_rn(x, h, nz);
                    a[i] = __ffma2_rn(a[i], one, p);
                }
            }
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 8; i++) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
int run(const char *name, int flops_per_inner, int lane_ops_per_inner) {
    int dev = 0; cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, dev));
    int sms = prop.multiProcessorCount;
    float2 *out; CHECK(cudaMalloc(&out, sizeof(float2) * sms * 8 * 256));
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int blocks_per_sm = 2; blocks_per_sm <= 8; blocks_per_sm *= 2) {
        pipe_kernel<MODE><<<sms * blocks_per_sm, 256>>>(out, 100, make_float2(1, 1), make_float2(-0.f, -0.f), make_float2(0.999f, 1.001f));
        CHECK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        pipe_kernel<MODE><<<sms * blocks_per_sm, 256>>>(out, iters, make_float2(1, 1), make_float2(-0.f, -0.f), make_float2(0.999f, 1.001f));
        cudaEventRecord(e1);
        CHECK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double inner = (double) sms * blocks_per_sm * 256 * iters * 64.0;  // 64 (u,i) bodies per iteration per thread
        double tflops = inner * flops_per_inner / (ms * 1e-3) / 1e12;
        double lane_ops = inner * lane_ops_per_inner / (ms * 1e-3);
        printf("%-28s blocks/SM %d: %.3f ms  %.2f TFLOP/s  %.1f G lane-ops/s (%.1f per clk per SM at 1.965 GHz)\n", name, blocks_per_sm, ms, tflops,
               lane_ops / 1e9, lane_ops / 1.965e9 / sms);
    }
    cudaFree(out);
    return 0;
}

int main() {
    cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, 0));
    printf("%s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    // flops counted algorithmically: mul+add = 2 per float lane element
    run<0>("scalar FFMA x2", 4, 2);
    run<1>("FFMA2", 4, 2);
    run<2>("scalar FMUL+FADD x2", 4, 4);
    run<3>("FFMA2 pair (exact mode)", 4, 4);
    return 0;
}
