/*
 * Batched GFSK modulator and the reference's modulator-side handles on top of the same kernels:
 *   sdrm_gfsk_mod_batch_*   N x gfsk_mod           (reference src/dsp/gfsk_mod.c:43-148; caller src/tcp_server.c:196,529)
 *   gfsk_mod_*              batch of one
 *   interp_fir_filter_*     reference src/dsp/interp_fir_filter.c:75-173
 *   frequency_modulator_*   reference src/dsp/frequency_modulator.c:21-70
 *   gaussian_taps_create, gfsk_mod_convolve      host tap design (taps.c)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/sdrm/frequency_modulator.h"
#include "../../include/sdrm/gaussian_taps.h"
#include "../../include/sdrm/gfsk_mod.h"
#include "../../include/sdrm/interp_fir_filter.h"
#include "../../include/sdrm/sdrm_batch.h"
#include "sdrm_internal.h"

/* ---- shared: polyphase tap layout ------------------------------------------------------------------------------------- */

/* Zero-pads taps to a multiple of `interpolation` (interp_fir_filter.c:19-40), splits them into branches
 * h_p[k] = h[k * I + p] (:42-73), reverses each branch (fir_filter.c:27) and uploads [I][K]. */
static int upload_branches(const float *taps, size_t taps_len, int interpolation, float **d_taps_rev, int *branch_taps) {
    const size_t inter = (size_t) interpolation;
    const size_t padded = (taps_len + inter - 1) / inter * inter;
    const size_t k_len = padded / inter;
    if (k_len == 0 || k_len > 64 || k_len * inter > 2048) {
        SDRM_LOG_ERROR("unsupported interpolator shape: %zu taps, interpolation %d", taps_len, interpolation);
        return -1;
    }
    float *rev = calloc(padded, sizeof(float));
    if (rev == NULL) {
        return -ENOMEM;
    }
    for (size_t p = 0; p < inter; p++) {
        for (size_t k = 0; k < k_len; k++) {
            const size_t src = k * inter + p;
            const float h = src < taps_len ? taps[src] : 0.0f;
            rev[p * k_len + (k_len - 1 - k)] = h;
        }
    }
    int code = sdrm_dev_zalloc((void **) d_taps_rev, padded * sizeof(float));
    if (code == 0) {
        code = sdrm_cuda_code(cudaMemcpy(*d_taps_rev, rev, padded * sizeof(float), cudaMemcpyHostToDevice), "branch taps upload");
    }
    free(rev);
    *branch_taps = (int) k_len;
    return code;
}

int gaussian_taps_create(double gain, double samples_per_symbol, double bt, size_t taps_len, float **taps) {
    return sdrm_design_gaussian(gain, samples_per_symbol, bt, taps_len, taps);
}

int gfsk_mod_convolve(float *x, size_t x_len, float *y, size_t y_len, float **out, size_t *out_len) {
    return sdrm_convolve_full(x, x_len, y, y_len, out, out_len);
}

/* ---- batch ------------------------------------------------------------------------------------------------------------ */

struct sdrm_gfsk_mod_batch_t {
    int device;
    uint32_t n_ch;
    uint32_t max_bytes;
    int interpolation;
    int branch_taps;
    float sensitivity;
    float *d_taps_rev;
    float *d_history; /* [n_ch][branch_taps - 1] */
    float *d_phase;   /* [n_ch] */
    float *d_work[2]; /* GTC layout [groups][work_rows][32]: increments, then phases, in place; one per call in flight */
    size_t work_rows;
    uint32_t n_groups;
    void *d_in;
    size_t in_stride_dev;
    void *d_out;
    void *d_out16; /* int16 pairs, process_i16 only */
    size_t out_stride_dev;
    /* three stages on three streams, so that call k's trigonometry, call k+1's phase walk and call k+2's shaping overlap:
     * the walk is serial per channel and leaves most of the GPU idle */
    cudaStream_t s_shape; /* input copy, bits -> shaped increments */
    cudaStream_t s_walk;  /* float phase recurrence */
    cudaStream_t stream;  /* phases -> cf32; the public stream: results are complete on it */
    cudaEvent_t ev_shaped[2];
    cudaEvent_t ev_walked[2];
    cudaEvent_t ev_read[2]; /* the work buffer has been read out */
    uint64_t calls;
    uint64_t launches;
};

int sdrm_gfsk_mod_batch_create(uint32_t n_channels, float samples_per_symbol, float sensitivity, float bt,
                               uint32_t max_input_buffer_length, int device, sdrm_gfsk_mod_batch **batch) {
    if (n_channels == 0 || batch == NULL || !(samples_per_symbol >= 1.0f) || samples_per_symbol >= 256.0f) {
        return -1;
    }
    sdrm_gfsk_mod_batch *b = calloc(1, sizeof(*b));
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
    float *gauss = NULL;
    float *square = NULL;
    float *taps = NULL;
    size_t taps_len = 0;
    /* taps = gaussian(4 * sps) convolved with a one-symbol boxcar (gfsk_mod.c:57-77) */
    const size_t gauss_len = (size_t) (4 * samples_per_symbol);
    const size_t square_len = (size_t) (int) samples_per_symbol;
    if (code == 0) code = sdrm_design_gaussian(1.0F, samples_per_symbol, bt, gauss_len, &gauss);
    if (code == 0) {
        square = malloc(sizeof(float) * square_len);
        if (square == NULL) {
            code = -ENOMEM;
        } else {
            for (size_t i = 0; i < square_len; i++) {
                square[i] = 1.0F;
            }
            code = sdrm_convolve_full(gauss, gauss_len, square, square_len, &taps, &taps_len);
        }
    }
    b->n_ch = n_channels;
    b->max_bytes = max_input_buffer_length;
    b->interpolation = (int) samples_per_symbol;
    b->sensitivity = sensitivity;
    if (code == 0) code = upload_branches(taps, taps_len, b->interpolation, &b->d_taps_rev, &b->branch_taps);
    free(gauss);
    free(square);
    free(taps);
    const size_t max_out = (size_t) max_input_buffer_length * 8 * (size_t) b->interpolation;
    b->work_rows = sdrm_round_up(max_out, 32) + 32;
    b->n_groups = (n_channels + 31) / 32;
    if (code == 0) code = sdrm_dev_zalloc((void **) &b->d_history, (size_t) n_channels * (b->branch_taps > 1 ? b->branch_taps - 1 : 1) * sizeof(float));
    if (code == 0) code = sdrm_dev_zalloc((void **) &b->d_phase, n_channels * sizeof(float));
    for (int i = 0; i < 2 && code == 0; i++) {
        code = sdrm_dev_zalloc((void **) &b->d_work[i], (size_t) b->n_groups * b->work_rows * 32 * sizeof(float));
        if (code == 0) code = sdrm_cuda_code(cudaEventCreateWithFlags(&b->ev_shaped[i], cudaEventDisableTiming), "event");
        if (code == 0) code = sdrm_cuda_code(cudaEventCreateWithFlags(&b->ev_walked[i], cudaEventDisableTiming), "event");
        if (code == 0) code = sdrm_cuda_code(cudaEventCreateWithFlags(&b->ev_read[i], cudaEventDisableTiming), "event");
    }
    {
        /* the walk has few, long-running blocks and the next call's shaping feeds it: both go ahead of the wide trigonometry
         * kernel of the previous call, whose blocks would otherwise fill every SM first */
        int least = 0;
        int greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        if (code == 0) code = sdrm_cuda_code(cudaStreamCreateWithPriority(&b->s_shape, cudaStreamNonBlocking, greatest), "stream");
        if (code == 0) code = sdrm_cuda_code(cudaStreamCreateWithPriority(&b->s_walk, cudaStreamNonBlocking, greatest), "stream");
        if (code == 0) code = sdrm_cuda_code(cudaStreamCreateWithPriority(&b->stream, cudaStreamNonBlocking, least), "stream");
    }
    if (code != 0) {
        sdrm_gfsk_mod_batch_destroy(b);
        return code;
    }
    *batch = b;
    return 0;
}

static int mod_enqueue(sdrm_gfsk_mod_batch *b, const void *d_in, size_t in_stride, size_t n_bytes, void *d_out, size_t out_stride) {
    sdrm_interp_args a;
    memset(&a, 0, sizeof(a));
    a.in = d_in;
    a.in_stride = in_stride;
    a.in_is_bytes = 1;
    a.n_in = (int) (n_bytes * 8);
    a.n_ch = (int) b->n_ch;
    a.interpolation = b->interpolation;
    a.branch_taps = b->branch_taps;
    a.taps_rev = b->d_taps_rev;
    a.history = b->d_history;
    a.apply_scale = 1;
    a.scale = b->sensitivity;
    const int slot = (int) (b->calls & 1);
    float *work = b->d_work[slot];
    a.out = work;
    a.out_stride = b->work_rows;
    a.out_grouped = 1;
    /* the work buffer of two calls ago has been read out */
    SDRM_CUDA_TRY(cudaStreamWaitEvent(b->s_shape, b->ev_read[slot], 0));
    int code = sdrm_launch_code(sdrm_cu_interp_fir(&a, b->s_shape), "gfsk shaping");
    if (code != 0) return code;
    SDRM_CUDA_TRY(cudaEventRecord(b->ev_shaped[slot], b->s_shape));
    const long long n_out = (long long) n_bytes * 8 * b->interpolation;
    SDRM_CUDA_TRY(cudaStreamWaitEvent(b->s_walk, b->ev_shaped[slot], 0));
    code = sdrm_launch_code(sdrm_cu_phase_walk(work, b->work_rows, b->d_phase, n_out, (int) b->n_ch, b->s_walk), "phase walk");
    if (code != 0) return code;
    SDRM_CUDA_TRY(cudaEventRecord(b->ev_walked[slot], b->s_walk));
    SDRM_CUDA_TRY(cudaStreamWaitEvent(b->stream, b->ev_walked[slot], 0));
    code = sdrm_launch_code(sdrm_cu_phase_to_iq(work, b->work_rows, d_out, out_stride, n_out, (int) b->n_ch, b->stream),
                            "phase to iq");
    if (code != 0) return code;
    SDRM_CUDA_TRY(cudaEventRecord(b->ev_read[slot], b->stream));
    b->calls++;
    b->launches += 4;
    return 0;
}

static int mod_check(const sdrm_gfsk_mod_batch *b, size_t n_bytes) {
    if (n_bytes > b->max_bytes) {
        SDRM_LOG_ERROR("requested buffer %zu is more than max: %zu", n_bytes, (size_t) b->max_bytes);
        return -1;
    }
    return 0;
}

int sdrm_gfsk_mod_batch_process_device(sdrm_gfsk_mod_batch *b, const void *d_input, size_t in_stride, size_t input_len,
                                       void *d_output, size_t out_stride) {
    if (b == NULL || d_output == NULL || (d_input == NULL && input_len > 0) || mod_check(b, input_len) != 0) {
        return -1;
    }
    SDRM_CUDA_TRY(cudaSetDevice(b->device));
    return mod_enqueue(b, d_input, in_stride, input_len, d_output, out_stride);
}

int sdrm_gfsk_mod_batch_process(sdrm_gfsk_mod_batch *b, const uint8_t *input, size_t in_stride, size_t input_len,
                                float complex *output, size_t out_stride, size_t *output_len) {
    if (b == NULL || output == NULL || (input == NULL && input_len > 0) || mod_check(b, input_len) != 0) {
        return -1;
    }
    SDRM_CUDA_TRY(cudaSetDevice(b->device));
    const size_t n_out = input_len * 8 * (size_t) b->interpolation;
    if (b->d_in == NULL) {
        b->in_stride_dev = sdrm_round_up((size_t) b->max_bytes, 16) + 16;
        int code = sdrm_dev_zalloc(&b->d_in, (size_t) b->n_ch * b->in_stride_dev);
        if (code != 0) return code;
        b->out_stride_dev = sdrm_round_up((size_t) b->max_bytes * 8 * (size_t) b->interpolation, 2) + 2;
        code = sdrm_dev_zalloc(&b->d_out, (size_t) b->n_ch * b->out_stride_dev * 8);
        if (code != 0) return code;
    }
    if (input_len > 0) {
        SDRM_CUDA_TRY(cudaMemcpy2DAsync(b->d_in, b->in_stride_dev, input, in_stride, input_len, b->n_ch, cudaMemcpyHostToDevice, b->s_shape));
    }
    int code = mod_enqueue(b, b->d_in, b->in_stride_dev, input_len, b->d_out, b->out_stride_dev);
    if (code != 0) return code;
    if (n_out > 0) {
        SDRM_CUDA_TRY(cudaMemcpy2DAsync(output, out_stride * 8, b->d_out, b->out_stride_dev * 8, n_out * 8, b->n_ch, cudaMemcpyDeviceToHost,
                                        b->stream));
    }
    SDRM_CUDA_TRY(cudaStreamSynchronize(b->stream));
    if (output_len != NULL) {
        *output_len = n_out;
    }
    return 0;
}

/* as process, with the egress conversion of the PlutoSDR plugin (reference src/sdr/plutosdr.c:83) done on the device:
 * output int16 (I, Q) pairs [channels][out_stride pairs] */
int sdrm_gfsk_mod_batch_process_i16(sdrm_gfsk_mod_batch *b, const uint8_t *input, size_t in_stride, size_t input_len,
                                    int16_t *output, size_t out_stride, float scalar, size_t *output_len) {
    if (b == NULL || output == NULL || (input == NULL && input_len > 0) || mod_check(b, input_len) != 0) {
        return -1;
    }
    SDRM_CUDA_TRY(cudaSetDevice(b->device));
    const size_t n_out = input_len * 8 * (size_t) b->interpolation;
    if (b->d_in == NULL) {
        b->in_stride_dev = sdrm_round_up((size_t) b->max_bytes, 16) + 16;
        int code = sdrm_dev_zalloc(&b->d_in, (size_t) b->n_ch * b->in_stride_dev);
        if (code != 0) return code;
        b->out_stride_dev = sdrm_round_up((size_t) b->max_bytes * 8 * (size_t) b->interpolation, 2) + 2;
        code = sdrm_dev_zalloc(&b->d_out, (size_t) b->n_ch * b->out_stride_dev * 8);
        if (code != 0) return code;
    }
    if (b->d_out16 == NULL) {
        int code = sdrm_dev_zalloc(&b->d_out16, (size_t) b->n_ch * b->out_stride_dev * 4);
        if (code != 0) return code;
    }
    if (input_len > 0) {
        SDRM_CUDA_TRY(cudaMemcpy2DAsync(b->d_in, b->in_stride_dev, input, in_stride, input_len, b->n_ch, cudaMemcpyHostToDevice, b->s_shape));
    }
    int code = mod_enqueue(b, b->d_in, b->in_stride_dev, input_len, b->d_out, b->out_stride_dev);
    if (code != 0) return code;
    if (n_out > 0) {
        code = sdrm_launch_code(sdrm_cu_cf32_to_i16(b->d_out, b->out_stride_dev, b->d_out16, b->out_stride_dev, scalar, (int) n_out,
                                                    (int) b->n_ch, b->stream),
                                "int16 egress");
        if (code != 0) return code;
        b->launches++;
        SDRM_CUDA_TRY(cudaMemcpy2DAsync(output, out_stride * 4, b->d_out16, b->out_stride_dev * 4, n_out * 4, b->n_ch, cudaMemcpyDeviceToHost,
                                        b->stream));
    }
    SDRM_CUDA_TRY(cudaStreamSynchronize(b->stream));
    if (output_len != NULL) {
        *output_len = n_out;
    }
    return 0;
}

int sdrm_gfsk_mod_batch_sync(sdrm_gfsk_mod_batch *b) {
    if (b == NULL) {
        return -1;
    }
    SDRM_CUDA_TRY(cudaSetDevice(b->device));
    SDRM_CUDA_TRY(cudaStreamSynchronize(b->s_shape));
    SDRM_CUDA_TRY(cudaStreamSynchronize(b->s_walk));
    SDRM_CUDA_TRY(cudaStreamSynchronize(b->stream));
    return 0;
}

void *sdrm_gfsk_mod_batch_input_stream(sdrm_gfsk_mod_batch *b) { return b == NULL ? NULL : (void *) b->s_shape; }

void *sdrm_gfsk_mod_batch_stream(sdrm_gfsk_mod_batch *b) { return b == NULL ? NULL : (void *) b->stream; }

uint64_t sdrm_gfsk_mod_batch_launch_count(const sdrm_gfsk_mod_batch *b) { return b == NULL ? 0 : b->launches; }

void sdrm_gfsk_mod_batch_destroy(sdrm_gfsk_mod_batch *b) {
    if (b == NULL) {
        return;
    }
    cudaSetDevice(b->device);
    cudaStream_t streams[3] = {b->s_shape, b->s_walk, b->stream};
    for (int i = 0; i < 3; i++) {
        if (streams[i] != NULL) {
            cudaStreamSynchronize(streams[i]);
            cudaStreamDestroy(streams[i]);
        }
    }
    for (int i = 0; i < 2; i++) {
        cudaFree(b->d_work[i]);
        if (b->ev_shaped[i] != NULL) cudaEventDestroy(b->ev_shaped[i]);
        if (b->ev_walked[i] != NULL) cudaEventDestroy(b->ev_walked[i]);
        if (b->ev_read[i] != NULL) cudaEventDestroy(b->ev_read[i]);
    }
    cudaFree(b->d_taps_rev);
    cudaFree(b->d_history);
    cudaFree(b->d_phase);
    cudaFree(b->d_in);
    cudaFree(b->d_out);
    cudaFree(b->d_out16);
    free(b);
}

/* ---- gfsk_mod handle ----------------------------------------------------------------------------------------------------- */

struct gfsk_mod_t {
    sdrm_gfsk_mod_batch *batch;
    float complex *output;
    size_t output_cap;
    size_t max_bytes;
    size_t samples_per_byte;
    size_t locked_bytes; /* prefix of `output` that is page-locked (0: none) */
};

/* The handle's output buffer is sized for the largest call (258 MB for the reference's perf program, which then sends 2048
 * bytes at a time): only a prefix is page-locked, so that the results of ordinary calls arrive by DMA instead of through the
 * driver's staging buffers; a call that needs more gives the lock up. */
#define LOCKED_OUTPUT_PREFIX ((size_t) 16 << 20)

int gfsk_mod_create(float samples_per_symbol, float sensitivity, float bt, uint32_t max_input_buffer_length, gfsk_mod **mod) {
    struct gfsk_mod_t *result = calloc(1, sizeof(*result));
    if (result == NULL) {
        return -ENOMEM;
    }
    int code = sdrm_gfsk_mod_batch_create(1, samples_per_symbol, sensitivity, bt, max_input_buffer_length, -1, &result->batch);
    if (code != 0) {
        gfsk_mod_destroy(result);
        return code;
    }
    result->max_bytes = max_input_buffer_length;
    result->samples_per_byte = 8 * (size_t) (int) samples_per_symbol;
    result->output_cap = (size_t) max_input_buffer_length * result->samples_per_byte;
    const size_t output_bytes = sizeof(float complex) * (result->output_cap == 0 ? 1 : result->output_cap);
    void *buffer = NULL;
    if (posix_memalign(&buffer, 4096, output_bytes) != 0) {
        gfsk_mod_destroy(result);
        return -ENOMEM;
    }
    result->output = buffer;
    const size_t prefix = output_bytes < LOCKED_OUTPUT_PREFIX ? output_bytes : LOCKED_OUTPUT_PREFIX;
    if (cudaHostRegister(result->output, prefix, cudaHostRegisterPortable) == cudaSuccess) {
        result->locked_bytes = prefix;
    } else {
        (void) cudaGetLastError(); /* pageable it stays */
    }
    *mod = result;
    return 0;
}

void gfsk_mod_process(const uint8_t *input, size_t input_len, float complex **output, size_t *output_len, gfsk_mod *mod) {
    size_t produced = 0;
    if (mod->locked_bytes != 0 && input_len <= mod->max_bytes && input_len * mod->samples_per_byte * sizeof(float complex) > mod->locked_bytes) {
        cudaHostUnregister(mod->output); /* one copy must not straddle locked and pageable memory */
        mod->locked_bytes = 0;
    }
    if (sdrm_gfsk_mod_batch_process(mod->batch, input, mod->max_bytes, input_len, mod->output, mod->output_cap, &produced) != 0) {
        *output = NULL;
        *output_len = 0;
        return;
    }
    *output = mod->output;
    *output_len = produced;
}

void gfsk_mod_destroy(gfsk_mod *mod) {
    if (mod == NULL) {
        return;
    }
    sdrm_gfsk_mod_batch_destroy(mod->batch);
    if (mod->locked_bytes != 0) {
        cudaHostUnregister(mod->output);
    }
    free(mod->output);
    free(mod);
}

/* ---- interp_fir_filter handle ------------------------------------------------------------------------------------------------ */

struct interp_fir_filter_t {
    int interpolation;
    int branch_taps;
    uint32_t max_len;
    float *d_taps_rev;
    float *d_history;
    float *d_in;
    float *d_out;
    float *output;
    cudaStream_t stream;
};

int interp_fir_filter_create(float *taps, size_t taps_len, uint8_t interpolation, uint32_t max_input_buffer_length,
                             interp_fir_filter **filter) {
    if (interpolation == 0 || taps == NULL || taps_len == 0) {
        return -1;
    }
    /* the reference rejects a bad VOLK_ALIGNMENT here (fir_filter.c:44-51, test_interp_fir_filter.c:63-72); the variable has no
     * meaning on the GPU but callers may rely on the failure */
    const char *alignment = getenv("VOLK_ALIGNMENT");
    if (alignment != NULL && strtol(alignment, NULL, 10) == 0) {
        SDRM_LOG_ERROR("invalid VOLK_ALIGNMENT specified: %s", alignment);
        return -1;
    }
    struct interp_fir_filter_t *f = calloc(1, sizeof(*f));
    if (f == NULL) {
        return -ENOMEM;
    }
    f->interpolation = interpolation;
    f->max_len = max_input_buffer_length;
    int code = upload_branches(taps, taps_len, interpolation, &f->d_taps_rev, &f->branch_taps);
    const size_t out_cap = (size_t) max_input_buffer_length * interpolation + 4;
    if (code == 0) code = sdrm_dev_zalloc((void **) &f->d_history, (size_t) (f->branch_taps > 1 ? f->branch_taps - 1 : 1) * sizeof(float));
    if (code == 0) code = sdrm_dev_zalloc((void **) &f->d_in, ((size_t) max_input_buffer_length + 4) * sizeof(float));
    if (code == 0) code = sdrm_dev_zalloc((void **) &f->d_out, out_cap * sizeof(float));
    if (code == 0) code = sdrm_cuda_code(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking), "stream");
    if (code == 0) {
        f->output = malloc(out_cap * sizeof(float));
        if (f->output == NULL) {
            code = -ENOMEM;
        }
    }
    if (code != 0) {
        interp_fir_filter_destroy(f);
        return code;
    }
    free(taps); /* ownership passes to the filter on success (interp_fir_filter.c:134) */
    *filter = f;
    return 0;
}

void interp_fir_filter_process(float *input, size_t input_len, float **output, size_t *output_len, interp_fir_filter *f) {
    *output = NULL;
    *output_len = 0;
    if (input_len > f->max_len) {
        SDRM_LOG_ERROR("requested buffer %zu is more than max: %zu", input_len, (size_t) f->max_len);
        return;
    }
    if (input_len > 0 && cudaMemcpyAsync(f->d_in, input, input_len * sizeof(float), cudaMemcpyHostToDevice, f->stream) != cudaSuccess) {
        return;
    }
    sdrm_interp_args a;
    memset(&a, 0, sizeof(a));
    a.in = f->d_in;
    a.in_stride = (size_t) f->max_len + 4;
    a.in_is_bytes = 0;
    a.n_in = (int) input_len;
    a.n_ch = 1;
    a.interpolation = f->interpolation;
    a.branch_taps = f->branch_taps;
    a.taps_rev = f->d_taps_rev;
    a.history = f->d_history;
    a.apply_scale = 0;
    a.out = f->d_out;
    a.out_stride = (size_t) f->max_len * f->interpolation + 4;
    if (sdrm_launch_code(sdrm_cu_interp_fir(&a, f->stream), "interp fir") != 0) {
        return;
    }
    const size_t n_out = input_len * (size_t) f->interpolation;
    if (n_out > 0 && cudaMemcpyAsync(f->output, f->d_out, n_out * sizeof(float), cudaMemcpyDeviceToHost, f->stream) != cudaSuccess) {
        return;
    }
    if (sdrm_cuda_code(cudaStreamSynchronize(f->stream), "interp_fir_filter_process") != 0) {
        return;
    }
    *output = f->output;
    *output_len = n_out;
}

void interp_fir_filter_destroy(interp_fir_filter *f) {
    if (f == NULL) {
        return;
    }
    if (f->stream != NULL) {
        cudaStreamSynchronize(f->stream);
        cudaStreamDestroy(f->stream);
    }
    cudaFree(f->d_taps_rev);
    cudaFree(f->d_history);
    cudaFree(f->d_in);
    cudaFree(f->d_out);
    free(f->output);
    free(f);
}

/* ---- frequency_modulator handle -------------------------------------------------------------------------------------------------- */

struct frequency_modulator_t {
    float sensitivity;
    uint32_t max_len;
    float *d_work;
    float *d_phase;
    void *d_out;
    float *h_scaled;
    float complex *output;
    cudaStream_t stream;
};

int frequency_modulator_create(float sensitivity, uint32_t max_input_buffer_length, frequency_modulator **mod) {
    struct frequency_modulator_t *m = calloc(1, sizeof(*m));
    if (m == NULL) {
        return -ENOMEM;
    }
    m->sensitivity = sensitivity;
    m->max_len = max_input_buffer_length;
    const size_t cap = sdrm_round_up((size_t) max_input_buffer_length, 32) + 32;
    int code = sdrm_dev_zalloc((void **) &m->d_work, cap * 32 * sizeof(float));
    if (code == 0) code = sdrm_dev_zalloc((void **) &m->d_phase, sizeof(float));
    if (code == 0) code = sdrm_dev_zalloc(&m->d_out, cap * 8);
    if (code == 0) code = sdrm_cuda_code(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking), "stream");
    if (code == 0) {
        m->h_scaled = calloc(cap * 32, sizeof(float));
        m->output = malloc(cap * sizeof(float complex));
        if (m->h_scaled == NULL || m->output == NULL) {
            code = -ENOMEM;
        }
    }
    if (code != 0) {
        frequency_modulator_destroy(m);
        return code;
    }
    *mod = m;
    return 0;
}

void frequency_modulator_process(float *input, size_t input_len, float complex **output, size_t *output_len, frequency_modulator *m) {
    *output = NULL;
    *output_len = 0;
    if (input_len > m->max_len) {
        SDRM_LOG_ERROR("requested buffer %zu is more than max: %zu", input_len, (size_t) m->max_len);
        return;
    }
    if (input_len == 0) {
        *output = m->output;
        return;
    }
    /* sensitivity * input[i], rounded to float before it is added to the phase (frequency_modulator.c:49) */
    /* one stream rides in lane 0 of a channel group (GTC layout [time][32]) */
    for (size_t i = 0; i < input_len; i++) {
        m->h_scaled[i * 32] = m->sensitivity * input[i];
    }
    const size_t cap = sdrm_round_up((size_t) m->max_len, 32) + 32;
    int ok = cudaMemcpyAsync(m->d_work, m->h_scaled, input_len * 32 * sizeof(float), cudaMemcpyHostToDevice, m->stream) == cudaSuccess;
    ok = ok && sdrm_launch_code(sdrm_cu_freq_mod(m->d_work, cap, m->d_phase, m->d_out, cap, (long long) input_len, 1, m->stream),
                                "frequency modulator") == 0;
    ok = ok && cudaMemcpyAsync(m->output, m->d_out, input_len * 8, cudaMemcpyDeviceToHost, m->stream) == cudaSuccess;
    ok = ok && sdrm_cuda_code(cudaStreamSynchronize(m->stream), "frequency_modulator_process") == 0;
    if (!ok) {
        return;
    }
    *output = m->output;
    *output_len = input_len;
}

void frequency_modulator_destroy(frequency_modulator *m) {
    if (m == NULL) {
        return;
    }
    if (m->stream != NULL) {
        cudaStreamSynchronize(m->stream);
        cudaStreamDestroy(m->stream);
    }
    cudaFree(m->d_work);
    cudaFree(m->d_phase);
    cudaFree(m->d_out);
    free(m->h_scaled);
    free(m->output);
    free(m);
}
