/*
 * N x lpf (reference src/dsp/lpf.c:12-51 over src/dsp/fir_filter.c:35-159): the same low-pass filter, with its own history,
 * over N independent streams in one launch. Complex streams are FIR rows as they are; real streams are interleaved two
 * channels per row (PAIR layout) so that the same complex-row kernel does two real filters per pass.
 * Typical use: a decimating channel filter between sdrm_doppler_batch and sdrm_fsk_demod_batch (device buffers).
 */
#include <stdlib.h>
#include <string.h>

#include "../../include/sdrm/sdrm_batch.h"
#include "sdrm_internal.h"

struct sdrm_lpf_batch_t {
    int device;
    uint32_t n_ch;
    uint32_t rows;      /* FIR rows: n_ch (complex) or ceil(n_ch / 2) (real) */
    uint8_t decimation;
    size_t num_bytes;
    uint32_t max_len;
    int n_taps;
    int hist_len;
    int phase;
    int cur;
    void *d_taps;
    float *h_taps; /* host copy of the (h, h) pairs */
    void *d_hist[2];
    size_t stride;      /* float2 per row of the internal buffers */
    void *d_pair_in;    /* real streams: interleaved input */
    void *d_pair_out;   /* real streams: interleaved output */
    void *d_in;         /* host entry point: staging */
    void *d_out;
    cudaStream_t stream;
    uint64_t launches;
};

int sdrm_lpf_batch_create(uint32_t n_channels, uint8_t decimation, uint64_t sampling_freq, uint64_t cutoff_freq,
                          uint32_t transition_width, uint32_t max_input_buffer_length, size_t num_bytes, int device,
                          sdrm_lpf_batch **batch) {
    if (n_channels == 0 || batch == NULL || decimation == 0 || (num_bytes != 8 && num_bytes != 4)) {
        return -1;
    }
    sdrm_lpf_batch *b = calloc(1, sizeof(*b));
    if (b == NULL) {
        return -ENOMEM;
    }
    float *taps = NULL;
    size_t taps_len = 0;
    int code = sdrm_design_low_pass(1.0F, sampling_freq, cutoff_freq, transition_width, &taps, &taps_len);
    if (code == 0) {
        if (device >= 0) {
            b->device = device;
        } else {
            code = sdrm_cuda_code(cudaGetDevice(&b->device), "cudaGetDevice");
        }
    }
    if (code == 0) code = sdrm_cuda_code(cudaSetDevice(b->device), "cudaSetDevice");
    b->n_ch = n_channels;
    b->rows = num_bytes == 8 ? n_channels : (n_channels + 1) / 2;
    b->decimation = decimation;
    b->num_bytes = num_bytes;
    b->max_len = max_input_buffer_length;
    b->n_taps = (int) taps_len;
    b->hist_len = (int) sdrm_round_up(taps_len > 0 ? taps_len - 1 : 0, 2);
    if (b->hist_len == 0) {
        b->hist_len = 2;
    }
    b->stride = sdrm_round_up((size_t) max_input_buffer_length, 2) + 2;
    if (code == 0) code = sdrm_upload_taps_dup(taps, taps_len, &b->d_taps);
    if (code == 0) b->h_taps = sdrm_host_taps_dup(taps, taps_len);
    free(taps);
    for (int i = 0; i < 2 && code == 0; i++) {
        code = sdrm_dev_zalloc(&b->d_hist[i], (size_t) b->rows * b->hist_len * 8);
    }
    if (code == 0 && num_bytes == 4) {
        code = sdrm_dev_zalloc(&b->d_pair_in, (size_t) b->rows * b->stride * 8);
        if (code == 0) code = sdrm_dev_zalloc(&b->d_pair_out, (size_t) b->rows * b->stride * 8);
    }
    if (code == 0) code = sdrm_cuda_code(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking), "stream");
    if (code != 0) {
        sdrm_lpf_batch_destroy(b);
        return code;
    }
    *batch = b;
    return 0;
}

static int lpf_check(const sdrm_lpf_batch *b, size_t len) {
    if (len > b->max_len) {
        SDRM_LOG_ERROR("requested buffer %zu is more than max: %u", len, b->max_len);
        return -1;
    }
    return 0;
}

int sdrm_lpf_batch_process_device(sdrm_lpf_batch *b, const void *d_input, size_t in_stride, size_t input_len, void *d_output,
                                  size_t out_stride, size_t *output_len) {
    if (b == NULL || d_output == NULL || (d_input == NULL && input_len > 0) || lpf_check(b, input_len) != 0) {
        return -1;
    }
    SDRM_CUDA_TRY(cudaSetDevice(b->device));
    const int dec = b->decimation;
    const int n_in = (int) input_len;
    const int n_out = n_in > b->phase ? (n_in - b->phase + dec - 1) / dec : 0;
    if (output_len != NULL) {
        *output_len = (size_t) n_out;
    }
    if (n_in == 0) {
        return 0;
    }
    const void *fir_in = d_input;
    size_t fir_in_stride = in_stride;
    void *fir_out = d_output;
    size_t fir_out_stride = out_stride;
    int code = 0;
    if (b->num_bytes == 4) {
        code = sdrm_launch_code(sdrm_cu_rows_to_pairs((const float *) d_input, in_stride, b->d_pair_in, b->stride, n_in, (int) b->n_ch,
                                                      b->stream),
                                "lpf batch interleave");
        if (code != 0) return code;
        b->launches++;
        fir_in = b->d_pair_in;
        fir_in_stride = b->stride;
        fir_out = b->d_pair_out;
        fir_out_stride = b->stride;
    }
    sdrm_fir_args a;
    memset(&a, 0, sizeof(a));
    a.in = fir_in;
    a.in_stride = fir_in_stride;
    a.hist = b->d_hist[b->cur];
    a.hist_len = b->hist_len;
    a.taps_dup = b->d_taps;
    a.h_taps_dup = b->h_taps;
    a.n_taps = b->n_taps;
    a.decimation = dec;
    a.phase = b->phase;
    a.n_in = n_in;
    a.n_out = n_out;
    a.rows = (int) b->rows;
    a.out_mode = SDRM_FIR_OUT_ROWS;
    a.out = fir_out;
    a.out_stride = fir_out_stride;
    code = sdrm_launch_code(sdrm_cu_fir(&a, b->stream), "lpf batch");
    if (code != 0) return code;
    code = sdrm_launch_code(sdrm_cu_hist_update(fir_in, fir_in_stride, b->d_hist[b->cur], b->d_hist[b->cur ^ 1], b->hist_len, n_in,
                                                (int) b->rows, b->stream),
                            "lpf batch history");
    if (code != 0) return code;
    b->launches += 2;
    b->cur ^= 1;
    b->phase = b->phase + n_out * dec - n_in;
    if (b->num_bytes == 4 && n_out > 0) {
        code = sdrm_launch_code(sdrm_cu_pairs_to_rows(b->d_pair_out, b->stride, (float *) d_output, out_stride, n_out, (int) b->n_ch,
                                                      b->stream),
                                "lpf batch de-interleave");
        if (code != 0) return code;
        b->launches++;
    }
    return 0;
}

int sdrm_lpf_batch_process(sdrm_lpf_batch *b, const void *input, size_t in_stride, size_t input_len, void *output,
                           size_t out_stride, size_t *output_len) {
    if (b == NULL || output == NULL || (input == NULL && input_len > 0) || lpf_check(b, input_len) != 0) {
        return -1;
    }
    SDRM_CUDA_TRY(cudaSetDevice(b->device));
    const size_t elem = b->num_bytes;
    /* staging rows hold `stride` float2, i.e. 2 * stride floats for real streams */
    const size_t dev_stride = elem == 8 ? b->stride : 2 * b->stride;
    if (b->d_in == NULL) {
        int code = sdrm_dev_zalloc(&b->d_in, (size_t) b->n_ch * dev_stride * elem);
        if (code == 0) code = sdrm_dev_zalloc(&b->d_out, (size_t) b->n_ch * dev_stride * elem);
        if (code != 0) return code;
    }
    if (input_len > 0) {
        SDRM_CUDA_TRY(cudaMemcpy2DAsync(b->d_in, dev_stride * elem, input, in_stride * elem, input_len * elem, b->n_ch,
                                        cudaMemcpyHostToDevice, b->stream));
    }
    size_t n_out = 0;
    int code = sdrm_lpf_batch_process_device(b, b->d_in, dev_stride, input_len, b->d_out, dev_stride, &n_out);
    if (code != 0) return code;
    if (n_out > 0) {
        SDRM_CUDA_TRY(cudaMemcpy2DAsync(output, out_stride * elem, b->d_out, dev_stride * elem, n_out * elem, b->n_ch,
                                        cudaMemcpyDeviceToHost, b->stream));
    }
    SDRM_CUDA_TRY(cudaStreamSynchronize(b->stream));
    if (output_len != NULL) {
        *output_len = n_out;
    }
    return 0;
}

int sdrm_lpf_batch_sync(sdrm_lpf_batch *b) {
    if (b == NULL) {
        return -1;
    }
    SDRM_CUDA_TRY(cudaSetDevice(b->device));
    SDRM_CUDA_TRY(cudaStreamSynchronize(b->stream));
    return 0;
}

void *sdrm_lpf_batch_stream(sdrm_lpf_batch *b) { return b == NULL ? NULL : (void *) b->stream; }

uint64_t sdrm_lpf_batch_launch_count(const sdrm_lpf_batch *b) { return b == NULL ? 0 : b->launches; }

void sdrm_lpf_batch_destroy(sdrm_lpf_batch *b) {
    if (b == NULL) {
        return;
    }
    cudaSetDevice(b->device);
    if (b->stream != NULL) {
        cudaStreamSynchronize(b->stream);
        cudaStreamDestroy(b->stream);
    }
    cudaFree(b->d_taps);
    free(b->h_taps);
    cudaFree(b->d_hist[0]);
    cudaFree(b->d_hist[1]);
    cudaFree(b->d_pair_in);
    cudaFree(b->d_pair_out);
    cudaFree(b->d_in);
    cudaFree(b->d_out);
    free(b);
}
