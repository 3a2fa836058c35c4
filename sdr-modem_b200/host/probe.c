/*
 * Platform probe: what a plain pinned host -> device copy delivers on this box, right now, from the caller's buffer to the
 * caller's device memory. bench.py runs it on every rank at once with the same buffers and sizes as its end-to-end leg, so that
 * the end-to-end figure can be read as a fraction of what the host's PCIe / memory fabric gives that many GPUs at a time
 * (the 4- and 8-GPU end-to-end collapse had to be separated into platform limit and library overhead).
 * No kernel runs here and nothing of the reference corresponds to it.
 */
#include <stdlib.h>

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

/*
 * Warm start. Creating the CUDA context and loading this library's kernels takes about two seconds on a B200 box, and with lazy
 * initialisation the first *_create of the process pays for it: the first client of a server then waits that long for its
 * response (the reference's own integration test gives a client two seconds, test/test_tcp_server.c:406). A server calls
 * sdrm_warm_start(device) while it starts up, or sets SDRM_WARM_START=<device> in its environment, in which case the library does
 * it when it is loaded. Off by default: a process per GPU (torchrun) must not all touch device 0.
 */
int sdrm_warm_start(int device) {
    if (device >= 0) {
        SDRM_CUDA_TRY(cudaSetDevice(device));
    }
    SDRM_CUDA_TRY(cudaFree(NULL));
    double fma = 0.0;
    double pair = 0.0;
    return sdrm_measure_fp32_peak(device, &fma, &pair); /* a first kernel launch: loads the module */
}

__attribute__((constructor)) static void sdrm_warm_start_from_environment(void) {
    const char *text = getenv("SDRM_WARM_START");
    if (text != NULL && text[0] != '\0') {
        (void) sdrm_warm_start(atoi(text));
    }
}
