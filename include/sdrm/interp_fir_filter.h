/* Drop-in for reference src/dsp/interp_fir_filter.h:9-13. Takes ownership of `taps` on success, as the reference. */
#ifndef SDRM_INTERP_FIR_FILTER_H
#define SDRM_INTERP_FIR_FILTER_H

#include <stdint.h>
#include <stdlib.h>

typedef struct interp_fir_filter_t interp_fir_filter;

int interp_fir_filter_create(float *taps, size_t taps_len, uint8_t interpolation, uint32_t max_input_buffer_length,
                             interp_fir_filter **filter);

void interp_fir_filter_process(float *input, size_t input_len, float **output, size_t *output_len, interp_fir_filter *filter);

void interp_fir_filter_destroy(interp_fir_filter *filter);

#endif
