/*
 * Platform probe: what a plain pinned host -> device copy delivers on this box, right now, from the caller's buffer to the
 * caller's device memory. bench.py runs it on every rank at once with the same buffers and sizes as its end-to-end leg, so that
 * the end-to-end figure can be read as a fraction of what the host's PCIe / memory fabric gives that many GPUs at a time
 * (the 4- and 8-GPU end-to-end collapse had to be separated into platform limit and library overhead).
 * No kernel runs here and nothing of the reference corresponds to it.
 */
#include "../../include/sdrm/sdrm_batch.h"
#include "sdrm_internal.h"

int sdrm_probe_h2d(int device, const void *host, void *d_scratch, size_t bytes, int repeats, double *seconds) {
    if (host == NULL || d_scratch == NULL || seconds == NULL || repeats <= 0 || bytes == 0) {
        return -1;
    }
    if (device >= 0) {
        SDRM_CUDA_TRY(cudaSetDevice(device));
    }
    cudaStream_t stream = NULL;
    cudaEvent_t e0 = NULL;
    cudaEvent_t e1 = NULL;
    int code = sdrm_cuda_code(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "probe stream");
    if (code == 0) code = sdrm_cuda_code(cudaEventCreate(&e0), "probe event");
    if (code == 0) code = sdrm_cuda_code(cudaEventCreate(&e1), "probe event");
    if (code == 0) code = sdrm_cuda_code(cudaEventRecord(e0, stream), "probe record");
    for (int i = 0; i < repeats && code == 0; i++) {
        code = sdrm_cuda_code(cudaMemcpyAsync(d_scratch, host, bytes, cudaMemcpyHostToDevice, stream), "probe copy");
    }
    if (code == 0) code = sdrm_cuda_code(cudaEventRecord(e1, stream), "probe record");
    if (code == 0) code = sdrm_cuda_code(cudaEventSynchronize(e1), "probe sync");
    if (code == 0) {
        float ms = 0.0f;
        code = sdrm_cuda_code(cudaEventElapsedTime(&ms, e0, e1), "probe time");
        *seconds = (double) ms * 1e-3;
    }
    if (e0 != NULL) cudaEventDestroy(e0);
    if (e1 != NULL) cudaEventDestroy(e1);
    if (stream != NULL) cudaStreamDestroy(stream);
    return code;
}
