/*
 * fsk_demod_create / fsk_demod_process / fsk_demod_destroy with the reference's signatures and error behaviour
 * (reference src/dsp/fsk_demod.h:11-15, src/dsp/fsk_demod.c:28-135): a batch of one channel on the GPU.
 * Callers in the reference: src/dsp_worker.c:75,140-143.
 */
#include <stdlib.h>

#include "../../include/sdrm/fsk_demod.h"
#include "../../include/sdrm/sdrm_batch.h"
#include "sdrm_internal.h"

struct fsk_demod_t {
    sdrm_fsk_demod_batch *batch;
    int8_t *output; /* owned by the handle, valid until the next call (fsk_demod.c:24,108) */
    uint32_t output_len;
    uint32_t max_input_buffer_length;
    int output_pinned;
};

/* results of up to this many symbols land in page-locked memory (a DMA straight from the device); larger handles keep a
 * pageable buffer, which the driver stages */
#define PINNED_OUTPUT_LIMIT ((size_t) 16 << 20)

int fsk_demod_create(uint64_t sampling_freq, uint32_t baud_rate, int64_t deviation, uint8_t decimation,
                     uint32_t transition_width, bool use_dc_block, uint32_t max_input_buffer_length, fsk_demod **demod) {
    struct fsk_demod_t *result = calloc(1, sizeof(struct fsk_demod_t));
    if (result == NULL) {
        return -ENOMEM;
    }
    sdrm_fsk_demod_batch_config config = {0};
    config.n_channels = 1;
    config.sampling_freq = sampling_freq;
    config.baud_rate = baud_rate;
    config.deviation = deviation;
    config.decimation = decimation;
    config.transition_width = transition_width;
    config.use_dc_block = use_dc_block;
    config.max_input_buffer_length = max_input_buffer_length;
    config.device = -1;
    int code = sdrm_fsk_demod_batch_create(&config, &result->batch);
    if (code != 0) {
        fsk_demod_destroy(result);
        return code;
    }
    result->max_input_buffer_length = max_input_buffer_length;
    result->output_len = max_input_buffer_length;
    const size_t output_bytes = sizeof(int8_t) * (result->output_len == 0 ? 1 : result->output_len);
    if (output_bytes <= PINNED_OUTPUT_LIMIT) {
        result->output = sdrm_pinned_alloc(output_bytes);
        result->output_pinned = result->output != NULL;
    }
    if (result->output == NULL) {
        result->output = malloc(output_bytes);
    }
    if (result->output == NULL) {
        fsk_demod_destroy(result);
        return -ENOMEM;
    }
    *demod = result;
    return 0;
}

void fsk_demod_process(const float complex *input, size_t input_len, int8_t **output, size_t *output_len, fsk_demod *demod) {
    uint32_t produced = 0;
    int code = sdrm_fsk_demod_batch_process(demod->batch, input, demod->max_input_buffer_length, input_len, demod->output, NULL,
                                            demod->output_len, &produced);
    if (code != 0) {
        *output = NULL;
        *output_len = 0;
        return;
    }
    *output = demod->output;
    *output_len = produced;
}

void fsk_demod_destroy(fsk_demod *demod) {
    if (demod == NULL) {
        return;
    }
    sdrm_fsk_demod_batch_destroy(demod->batch);
    if (demod->output_pinned) {
        sdrm_pinned_free(demod->output);
    } else {
        free(demod->output);
    }
    free(demod);
}
