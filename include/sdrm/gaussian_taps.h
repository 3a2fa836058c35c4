/* Drop-in for reference src/dsp/gaussian_taps.h:6. Host only (runs once per create). */
#ifndef SDRM_GAUSSIAN_TAPS_H
#define SDRM_GAUSSIAN_TAPS_H

#include <stdlib.h>

int gaussian_taps_create(double gain, double samples_per_symbol, double bt, size_t taps_len, float **taps);

#endif
