/* Drop-in for reference src/dsp/lpf_taps.h:6. Host only (runs once per create). */
#ifndef SDRM_LPF_TAPS_H
#define SDRM_LPF_TAPS_H

#include <stddef.h>
#include <stdint.h>

int create_low_pass_filter(float gain, uint64_t sampling_freq, uint64_t cutoff_freq, uint32_t transition_width,
                           float **taps, size_t *len);

#endif
