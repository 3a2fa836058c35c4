// SDR sample formats on the device (sm_100a): int16 IQ <-> cf32, so that the PCIe side of the chain carries 4 bytes per
// complex sample instead of 8. Reference: the PlutoSDR plugin converts on the host with VOLK,
//   RX  volk_16i_s32f_convert_32f(out, in, 2048.0F, n)      src/sdr/plutosdr.c:129   out = (float) in / scalar
//   TX  volk_32f_s32f_convert_16i(out, in, 32768, n)        src/sdr/plutosdr.c:83    saturate, round half to even
// Both are pure streaming passes bound by HBM (12 bytes per complex sample); rows are independent.

#include <cuda_runtime.h>
#include <stdint.h>

#include "sdrm_cuda.h"

namespace {

// in: int16 (I, Q) pairs, row r at in + r * in_stride pairs; out: float2 rows. One thread per complex sample: the 4-byte
// loads and 8-byte stores of a warp are contiguous.
__global__ void i16_to_cf32_kernel(const short2 *__restrict__ in, size_t in_stride, float2 *__restrict__ out, size_t out_stride,
                                   float scalar, int n, int rows) {
    for (int row = blockIdx.y; row < rows; row += gridDim.y) {
        const short2 *x = in + (size_t) row * in_stride;
        float2 *y = out + (size_t) row * out_stride;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const short2 v = x[i];
            y[i] = make_float2(__fdiv_rn((float) v.x, scalar), __fdiv_rn((float) v.y, scalar));
        }
    }
}

__device__ __forceinline__ short to_i16(float v, float scalar) {
    float r = __fmul_rn(v, scalar);
    if (r != r) {
        return 0;
    }
    r = fminf(fmaxf(r, -32768.0f), 32767.0f);
    return (short) __float2int_rn(r);
}

__global__ void cf32_to_i16_kernel(const float2 *__restrict__ in, size_t in_stride, short2 *__restrict__ out, size_t out_stride,
                                   float scalar, int n, int rows) {
    for (int row = blockIdx.y; row < rows; row += gridDim.y) {
        const float2 *x = in + (size_t) row * in_stride;
        short2 *y = out + (size_t) row * out_stride;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const float2 v = x[i];
            y[i] = make_short2(to_i16(v.x, scalar), to_i16(v.y, scalar));
        }
    }
}

dim3 convert_grid(int n, int rows) {
    long long bx = ((long long) n + 255) / 256;
    const long long want = (148LL * 16 + rows - 1) / rows;
    if (bx > want) {
        bx = want;
    }
    // gridDim.y is limited to 65535: larger batches loop over their rows inside the kernel
    return dim3((unsigned) bx, (unsigned) (rows < 65535 ? rows : 65535));
}

}  // namespace

extern "C" int sdrm_cu_i16_to_cf32(const void *in, size_t in_stride, void *out, size_t out_stride, float scalar, int n, int rows,
                                   void *stream) {
    if (n <= 0 || rows <= 0) {
        return 0;
    }
    if (((uintptr_t) in & 3) != 0 || ((uintptr_t) out & 7) != 0) {
        return -22;
    }
    i16_to_cf32_kernel<<<convert_grid(n, rows), 256, 0, (cudaStream_t) stream>>>((const short2 *) in, in_stride, (float2 *) out,
                                                                                out_stride, scalar, n, rows);
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}

extern "C" int sdrm_cu_cf32_to_i16(const void *in, size_t in_stride, void *out, size_t out_stride, float scalar, int n, int rows,
                                   void *stream) {
    if (n <= 0 || rows <= 0) {
        return 0;
    }
    if (((uintptr_t) in & 7) != 0 || ((uintptr_t) out & 3) != 0) {
        return -22;
    }
    cf32_to_i16_kernel<<<convert_grid(n, rows), 256, 0, (cudaStream_t) stream>>>((const float2 *) in, in_stride, (short2 *) out,
                                                                                out_stride, scalar, n, rows);
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}

// Real streams ride through the FIR two channels at a time (PAIR layout, sdrm_cuda.h): these two passes move between
// channel rows float [ch][stride] and pair rows float2 [ch / 2][stride] = (ch 2p, ch 2p + 1). An odd last channel is
// paired with zeros.
namespace {

__global__ void rows_to_pairs_kernel(const float *__restrict__ in, size_t in_stride, float2 *__restrict__ out, size_t out_stride,
                                     int n, int n_ch) {
    for (int p = blockIdx.y; 2 * p < n_ch; p += gridDim.y) {
        const float *a = in + (size_t) (2 * p) * in_stride;
        const float *b = 2 * p + 1 < n_ch ? in + (size_t) (2 * p + 1) * in_stride : nullptr;
        float2 *y = out + (size_t) p * out_stride;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            y[i] = make_float2(a[i], b != nullptr ? b[i] : 0.0f);
        }
    }
}

__global__ void pairs_to_rows_kernel(const float2 *__restrict__ in, size_t in_stride, float *__restrict__ out, size_t out_stride,
                                     int n, int n_ch) {
    for (int p = blockIdx.y; 2 * p < n_ch; p += gridDim.y) {
        const float2 *x = in + (size_t) p * in_stride;
        float *a = out + (size_t) (2 * p) * out_stride;
        float *b = 2 * p + 1 < n_ch ? out + (size_t) (2 * p + 1) * out_stride : nullptr;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const float2 v = x[i];
            a[i] = v.x;
            if (b != nullptr) {
                b[i] = v.y;
            }
        }
    }
}

}  // namespace

extern "C" int sdrm_cu_rows_to_pairs(const float *in, size_t in_stride, void *out, size_t out_stride, int n, int n_ch, void *stream) {
    if (n <= 0 || n_ch <= 0) {
        return 0;
    }
    rows_to_pairs_kernel<<<convert_grid(n, (n_ch + 1) / 2), 256, 0, (cudaStream_t) stream>>>(in, in_stride, (float2 *) out, out_stride,
                                                                                            n, n_ch);
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}

extern "C" int sdrm_cu_pairs_to_rows(const void *in, size_t in_stride, float *out, size_t out_stride, int n, int n_ch, void *stream) {
    if (n <= 0 || n_ch <= 0) {
        return 0;
    }
    pairs_to_rows_kernel<<<convert_grid(n, (n_ch + 1) / 2), 256, 0, (cudaStream_t) stream>>>((const float2 *) in, in_stride, out,
                                                                                            out_stride, n, n_ch);
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -(int) err - 1000;
}
