/*
 * Batched NCO / mixer and the reference's sig_source handle on top of it (a batch of one).
 * Reference: src/dsp/sig_source.c:13-85; callers src/dsp/doppler.c:180, src/tcp_server.c:209,558,
 * src/sdr/file_source.c:52,122,145.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/sdrm/sdrm_batch.h"
#include "../../include/sdrm/sig_source.h"
#include "sdrm_internal.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define M_2PI ((float) (2 * M_PI))

struct sdrm_nco_batch_t {
    int device;
    uint32_t n_ch;
    uint32_t max_len;
    uint64_t fs;
    float *d_step;
    float *d_amp;
    float *d_phase;
    float *d_phases; /* scratch [n_ch][stride] */
    size_t stride;
    void *d_in;
    void *d_out;
    float *h_step;
    cudaStream_t stream;
};

int sdrm_nco_batch_create(uint32_t n_channels, float amplitude, uint64_t sampling_freq, uint32_t max_len, int device,
                          sdrm_nco_batch **batch) {
    if (n_channels == 0 || batch == NULL || sampling_freq == 0) {
        return -1;
    }
    sdrm_nco_batch *b = calloc(1, sizeof(*b));
    if (b == NULL) {
        return -ENOMEM;
    }
    int code = 0;
    if (device >= 0) {
        b->device = device;
    } else {
        code = sdrm_cuda_code(cudaGetDevice(&b->device), "cudaGetDevice");
    }
    if (code == 0) code = sdrm_cuda_code(cudaSetDevice(b->device), "cudaSetDevice");
    b->n_ch = n_channels;
    b->max_len = max_len;
    b->fs = sampling_freq;
    b->stride = sdrm_round_up((size_t) max_len, 2) + 2;
    if (code == 0) code = sdrm_dev_zalloc((void **) &b->d_step, n_channels * sizeof(float));
    if (code == 0) code = sdrm_dev_zalloc((void **) &b->d_amp, n_channels * sizeof(float));
    if (code == 0) code = sdrm_dev_zalloc((void **) &b->d_phase, n_channels * sizeof(float));
    if (code == 0) code = sdrm_dev_zalloc((void **) &b->d_phases, (size_t) n_channels * b->stride * sizeof(float));
    if (code == 0) code = sdrm_cuda_code(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking), "stream");
    if (code == 0) {
        b->h_step = malloc(n_channels * sizeof(float));
        if (b->h_step == NULL) {
            code = -ENOMEM;
        }
    }
    if (code == 0) {
        for (uint32_t c = 0; c < n_channels; c++) {
            b->h_step[c] = amplitude;
        }
        code = sdrm_cuda_code(cudaMemcpy(b->d_amp, b->h_step, n_channels * sizeof(float), cudaMemcpyHostToDevice), "amplitude upload");
    }
    if (code != 0) {
        sdrm_nco_batch_destroy(b);
        return code;
    }
    *batch = b;
    return 0;
}

static int nco_enqueue(sdrm_nco_batch *b, const int64_t *freq_hz, const void *d_in, size_t in_stride, size_t len, void *d_out,
                       size_t out_stride) {
    for (uint32_t c = 0; c < b->n_ch; c++) {
        /* float arithmetic, as sig_source.c:44 */
        b->h_step[c] = M_2PI * (float) freq_hz[c] / b->fs;
    }
    /* pageable source: the copy is staged before the call returns, so h_step can be reused next call */
    SDRM_CUDA_TRY(cudaMemcpyAsync(b->d_step, b->h_step, b->n_ch * sizeof(float), cudaMemcpyHostToDevice, b->stream));
    sdrm_nco_args a;
    memset(&a, 0, sizeof(a));
    a.in = d_in;
    a.in_stride = in_stride;
    a.out = d_out;
    a.out_stride = out_stride;
    a.step = b->d_step;
    a.amplitude = b->d_amp;
    a.phase_state = b->d_phase;
    a.phases = b->d_phases;
    a.phase_stride = b->stride;
    a.n = (int) len;
    a.n_ch = (int) b->n_ch;
    return sdrm_launch_code(sdrm_cu_nco(&a, b->stream), "nco");
}

static int nco_check(const sdrm_nco_batch *b, size_t len) {
    if (len > b->max_len) {
        SDRM_LOG_ERROR("requested buffer %zu is more than max: %u", len, b->max_len);
        return -1;
    }
    return 0;
}

int sdrm_nco_batch_process_device(sdrm_nco_batch *b, const int64_t *freq_hz, const void *d_input, size_t in_stride, size_t len,
                                  void *d_output, size_t out_stride) {
    if (b == NULL || freq_hz == NULL || d_output == NULL || nco_check(b, len) != 0) {
        return -1;
    }
    SDRM_CUDA_TRY(cudaSetDevice(b->device));
    return nco_enqueue(b, freq_hz, d_input, in_stride, len, d_output, out_stride);
}

int sdrm_nco_batch_process(sdrm_nco_batch *b, const int64_t *freq_hz, const float complex *input, size_t in_stride, size_t len,
                           float complex *output, size_t out_stride) {
    if (b == NULL || freq_hz == NULL || output == NULL || nco_check(b, len) != 0) {
        return -1;
    }
    SDRM_CUDA_TRY(cudaSetDevice(b->device));
    if (b->d_out == NULL) {
        int code = sdrm_dev_zalloc(&b->d_out, (size_t) b->n_ch * b->stride * 8);
        if (code != 0) return code;
    }
    if (input != NULL && b->d_in == NULL) {
        int code = sdrm_dev_zalloc(&b->d_in, (size_t) b->n_ch * b->stride * 8);
        if (code != 0) return code;
    }
    if (len == 0) {
        return 0;
    }
    if (input != NULL) {
        SDRM_CUDA_TRY(cudaMemcpy2DAsync(b->d_in, b->stride * 8, input, in_stride * 8, len * 8, b->n_ch, cudaMemcpyHostToDevice, b->stream));
    }
    int code = nco_enqueue(b, freq_hz, input != NULL ? b->d_in : NULL, b->stride, len, b->d_out, b->stride);
    if (code != 0) return code;
    SDRM_CUDA_TRY(cudaMemcpy2DAsync(output, out_stride * 8, b->d_out, b->stride * 8, len * 8, b->n_ch, cudaMemcpyDeviceToHost, b->stream));
    SDRM_CUDA_TRY(cudaStreamSynchronize(b->stream));
    return 0;
}

int sdrm_nco_batch_sync(sdrm_nco_batch *b) {
    if (b == NULL) {
        return -1;
    }
    SDRM_CUDA_TRY(cudaSetDevice(b->device));
    SDRM_CUDA_TRY(cudaStreamSynchronize(b->stream));
    return 0;
}

void *sdrm_nco_batch_stream(sdrm_nco_batch *b) { return b == NULL ? NULL : (void *) b->stream; }

void sdrm_nco_batch_destroy(sdrm_nco_batch *b) {
    if (b == NULL) {
        return;
    }
    cudaSetDevice(b->device);
    if (b->stream != NULL) {
        cudaStreamSynchronize(b->stream);
        cudaStreamDestroy(b->stream);
    }
    cudaFree(b->d_step);
    cudaFree(b->d_amp);
    cudaFree(b->d_phase);
    cudaFree(b->d_phases);
    cudaFree(b->d_in);
    cudaFree(b->d_out);
    free(b->h_step);
    free(b);
}

/* ---- the reference's single-stream handle ---------------------------------------------------------------------------- */

struct sig_source_t {
    sdrm_nco_batch *batch;
    float complex *output; /* owned by the handle (sig_source.c:17,56,73) */
    uint32_t output_len;
};

int sig_source_create(float amplitude, uint64_t rx_sampling_freq, uint32_t max_output_buffer_length, sig_source **source) {
    struct sig_source_t *result = calloc(1, sizeof(*result));
    if (result == NULL) {
        return -ENOMEM;
    }
    result->output_len = max_output_buffer_length;
    result->output = malloc(sizeof(float complex) * (max_output_buffer_length == 0 ? 1 : max_output_buffer_length));
    if (result->output == NULL) {
        sig_source_destroy(result);
        return -ENOMEM;
    }
    int code = sdrm_nco_batch_create(1, amplitude, rx_sampling_freq, max_output_buffer_length, -1, &result->batch);
    if (code != 0) {
        sig_source_destroy(result);
        return code;
    }
    *source = result;
    return 0;
}

void sig_source_process(int64_t freq, size_t expected_output_len, float complex **output, size_t *output_len, sig_source *source) {
    /* the reference does not bound-check here (sig_source.c:43-58); writing past the buffer is not reproduced */
    if (expected_output_len > source->output_len ||
        sdrm_nco_batch_process(source->batch, &freq, NULL, 0, expected_output_len, source->output, source->output_len) != 0) {
        *output = NULL;
        *output_len = 0;
        return;
    }
    *output = source->output;
    *output_len = expected_output_len;
}

void sig_source_multiply(int64_t freq, const float complex *input, size_t input_len, float complex **output, size_t *output_len,
                         sig_source *source) {
    if (input_len > source->output_len) {
        SDRM_LOG_ERROR("requested buffer %zu is more than max: %u", input_len, source->output_len);
        *output = NULL;
        *output_len = 0;
        return;
    }
    if (sdrm_nco_batch_process(source->batch, &freq, input, source->output_len, input_len, source->output, source->output_len) != 0) {
        *output = NULL;
        *output_len = 0;
        return;
    }
    *output = source->output;
    *output_len = input_len;
}

void sig_source_destroy(sig_source *source) {
    if (source == NULL) {
        return;
    }
    sdrm_nco_batch_destroy(source->batch);
    free(source->output);
    free(source);
}
