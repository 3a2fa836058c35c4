// Microbenchmark: what one warp pays per shared-memory instruction next to a dependent FADD chain (the shape of the
// serial stages: the modulator's phase walker, the moving averages and the clock loop of the demod tail).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o lsu_issue lsu_issue.cu
// V0: FADD chain over registers only          V1: + one LDS.32 and one STS.32 per step ([step][lane] tile)
// V2: + one LDS.128 and one STS.128 per 4 steps ([step/4][lane][4] tile)
// V3: LDS.32 only, independent sums           V4: LDS.128 only, independent sums
// V5: V1 with the stores dropped              V6: V2 with the stores dropped
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int N = 256;
constexpr int PASSES = 64;

template <int V>
__global__ void probe(const float *d_in, float *out, long long *cycles) {
    __shared__ __align__(16) float d[N * 32];
    for (int i = threadIdx.x; i < N * 32; i += 32) {
        d[i] = d_in[i];
    }
    __syncwarp();
    const int lane = threadIdx.x;
    float p = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
    long long t0 = clock64();
    for (int pass = 0; pass < PASSES; pass++) {
        for (int i0 = 0; i0 < N; i0 += 32) {
            if (V == 0) {
                float x = __int_as_float(0x3f800000 + pass + i0);
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    p = __fadd_rn(p, x);
                }
            } else if (V == 1 || V == 5) {
                float *c = d + i0 * 32 + lane;
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    p = __fadd_rn(p, c[i * 32]);
                    if (V == 1) c[i * 32] = p;
                }
            } else if (V == 2 || V == 6) {
                float4 *c = reinterpret_cast<float4 *>(d + i0 * 32) + lane;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    float4 x = c[i * 32];
                    x.x = p = __fadd_rn(p, x.x);
                    x.y = p = __fadd_rn(p, x.y);
                    x.z = p = __fadd_rn(p, x.z);
                    x.w = p = __fadd_rn(p, x.w);
                    if (V == 2) c[i * 32] = x;
                }
            } else if (V == 3) {
                const float *c = d + i0 * 32 + lane;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    p = __fadd_rn(p, c[i * 32]);
                    p1 = __fadd_rn(p1, c[(i + 1) * 32]);
                    p2 = __fadd_rn(p2, c[(i + 2) * 32]);
                    p3 = __fadd_rn(p3, c[(i + 3) * 32]);
                }
            } else {
                const float4 *c = reinterpret_cast<const float4 *>(d + i0 * 32) + lane;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float4 x = c[i * 32];
                    p = __fadd_rn(p, x.x);
                    p1 = __fadd_rn(p1, x.y);
                    p2 = __fadd_rn(p2, x.z);
                    p3 = __fadd_rn(p3, x.w);
                }
            }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = p + p1 + p2 + p3 + d[lane];
    if (threadIdx.x == 0) {
        *cycles = t1 - t0;
    }
}

template <int V>
double run(const float *d_in, float *out, long long *cyc) {
    long long c = 0;
    for (int rep = 0; rep < 2; rep++) {
        probe<V><<<1, 32>>>(d_in, out, cyc);
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    }
    return c / (double) (N * PASSES);
}

int main() {
    float *h = new float[N * 32];
    for (int i = 0; i < N * 32; i++) {
        h[i] = (i % 977) * 1e-3f;
    }
    float *d_in, *out;
    long long *cyc;
    cudaMalloc(&d_in, N * 32 * 4);
    cudaMalloc(&out, 128);
    cudaMalloc(&cyc, 8);
    cudaMemcpy(d_in, h, N * 32 * 4, cudaMemcpyHostToDevice);
    printf("cycles per step, one warp: chain only %.2f | +LDS.32+STS.32 %.2f | +LDS.128+STS.128 per 4 steps %.2f\n", run<0>(d_in, out, cyc),
           run<1>(d_in, out, cyc), run<2>(d_in, out, cyc));
    printf("loads only, independent sums: LDS.32 %.2f | LDS.128 %.2f ; chain + loads only: LDS.32 %.2f | LDS.128 %.2f\n", run<3>(d_in, out, cyc),
           run<4>(d_in, out, cyc), run<5>(d_in, out, cyc), run<6>(d_in, out, cyc));
    printf("err=%d\n", (int) cudaDeviceSynchronize());
    return 0;
}
