// Measures what the FP32 pipe of THIS device delivers for the two instruction mixes of the FIR kernels, so that roofline
// fractions are quoted against a measured ceiling instead of the nominal formula (SMs x 128 lanes x 2 x clock):
//   mode 0  one FFMA2 per tap                    (fast / FMA arithmetic)          -> the FMA peak
//   mode 1  FFMA2 pair, multiply then add        (exact arithmetic, fir.cu mac2)  -> half the flops per pipe slot
// Eight independent float2 accumulators per thread, operands from registers only, nothing loop invariant; flops are counted
// algorithmically (multiply + add = 2 per float lane and tap) in both modes. Synchronous; diagnostic only.
#include <cuda_runtime.h>

#include "sdrm_cuda.h"

namespace {

template <int MODE>
__global__ void __launch_bounds__(256) fp32_pipe_kernel(float2 *out, int iters, float2 one, float2 negzero, float2 h) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float2 x = a[(i + 3) % 8];
                if (MODE == 0) {
                    a[i] = __ffma2_rn(x, h, a[i]);
                } else {
                    const float2 p = __ffma2_rn(x, h, negzero);
                    a[i] = __ffma2_rn(a[i], one, p);
                }
            }
        }
    }
    float2 s = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        s.x += a[i].x;
        s.y += a[i].y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
int measure(int sms, float2 *out, double *tflops) {
    const int blocks = sms * 4;
    const int iters = 6000;
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
        return -5;
    }
    double best = 0.0;
    // the first launch also warms the clocks up; best of five timed launches of ~1 ms (mode 0) / ~2 ms (mode 1)
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0, 0);
        fp32_pipe_kernel<MODE><<<blocks, 256>>>(out, iters, make_float2(1.0f, 1.0f), make_float2(-0.0f, -0.0f),
                                                make_float2(0.999f, 1.001f));
        cudaEventRecord(e1, 0);
        if (cudaEventSynchronize(e1) != cudaSuccess) {
            return -5;
        }
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double taps = (double) blocks * 256 * iters * 64.0;  // 64 float2 taps per iteration and thread
        const double rate = taps * 4.0 / (ms * 1e-3) / 1e12;        // 2 lanes x (multiply + add)
        if (rep > 0 && rate > best) {
            best = rate;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return 0;
}

}  // namespace

extern "C" int sdrm_measure_fp32_peak(int device, double *fma_tflops, double *exact_pair_tflops) {
    if (fma_tflops == nullptr || exact_pair_tflops == nullptr) {
        return -1;
    }
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) {
        return -5;
    }
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        return -5;
    }
    float2 *out = nullptr;
    if (cudaMalloc(&out, sizeof(float2) * prop.multiProcessorCount * 4 * 256) != cudaSuccess) {
        return -12;
    }
    int code = measure<0>(prop.multiProcessorCount, out, fma_tflops);
    if (code == 0) {
        code = measure<1>(prop.multiProcessorCount, out, exact_pair_tflops);
    }
    cudaFree(out);
    return code;
}
