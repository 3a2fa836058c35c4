/* Drop-in for reference src/dsp/fir_filter.h:29-35: streaming decimating FIR with real taps over complex (num_bytes 8)
 * or real (num_bytes 4) samples. `taps` is a malloc'ed array in design order that the filter owns after a successful
 * create (fir_filter.c:58,171-173). The reference declares the struct's fields (fir_filter.h:9-27); they describe its
 * host working buffers and nothing outside src/dsp reads them, so the type is opaque here. */
#ifndef SDRM_FIR_FILTER_H
#define SDRM_FIR_FILTER_H

#include <stdint.h>
#include <stdlib.h>

typedef struct fir_filter_t fir_filter;

int fir_filter_create(uint8_t decimation, float *taps, size_t taps_len, size_t output_len, size_t num_bytes, fir_filter **filter);

void fir_filter_process(const void *input, size_t input_len, void **output, size_t *output_len, fir_filter *filter);

/* one dot product over taps_len host floats starting at input, no history (used by the MMSE interpolator) */
float fir_filter_process_float_single(const float *input, fir_filter *filter);

void fir_filter_destroy(fir_filter *filter);

#endif
