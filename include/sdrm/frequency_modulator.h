/* Drop-in for reference src/dsp/frequency_modulator.h:10-14. */
#ifndef SDRM_FREQUENCY_MODULATOR_H
#define SDRM_FREQUENCY_MODULATOR_H

#include <complex.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct frequency_modulator_t frequency_modulator;

int frequency_modulator_create(float sensitivity, uint32_t max_input_buffer_length, frequency_modulator **mod);

void frequency_modulator_process(float *input, size_t input_len, float complex **output, size_t *output_len,
                                 frequency_modulator *mod);

void frequency_modulator_destroy(frequency_modulator *mod);

#endif
