/* Drop-in for reference src/dsp/lpf.h:10-14. num_bytes selects complex (8) or real (4) samples. */
#ifndef SDRM_LPF_H
#define SDRM_LPF_H

#include <complex.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct lpf_t lpf;

int lpf_create(uint8_t decimation, uint64_t sampling_freq, uint64_t cutoff_freq, uint32_t transition_width,
               size_t output_len, size_t num_bytes, lpf **filter);

void lpf_process(const void *input, size_t input_len, void **output, size_t *output_len, lpf *filter);

void lpf_destroy(lpf *filter);

#endif
