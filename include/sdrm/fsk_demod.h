/* Drop-in for reference src/dsp/fsk_demod.h:11-15 — same names, arguments and error behaviour; the chain runs on
 * the GPU (a batch of one channel, see sdrm_batch.h). */
#ifndef SDRM_FSK_DEMOD_H
#define SDRM_FSK_DEMOD_H

#include <complex.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct fsk_demod_t fsk_demod;

int fsk_demod_create(uint64_t sampling_freq, uint32_t baud_rate, int64_t deviation, uint8_t decimation,
                     uint32_t transition_width, bool use_dc_block, uint32_t max_input_buffer_length, fsk_demod **demod);

void fsk_demod_process(const float complex *input, size_t input_len, int8_t **output, size_t *output_len, fsk_demod *demod);

void fsk_demod_destroy(fsk_demod *demod);

#endif
