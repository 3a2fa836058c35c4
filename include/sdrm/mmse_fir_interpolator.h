/* Drop-in for reference src/dsp/mmse_fir_interpolator.h:8-14: the 8-tap MMSE fractional interpolator, one output per call.
 * A scalar helper (the clock loop inside clock_mm / fsk_demod runs it on the GPU); evaluated on the host here. */
#ifndef SDRM_MMSE_FIR_INTERPOLATOR_H
#define SDRM_MMSE_FIR_INTERPOLATOR_H

#include <stdlib.h>

typedef struct mmse_fir_interpolator_t mmse_fir_interpolator;

int mmse_fir_interpolator_create(mmse_fir_interpolator **interp);

float mmse_fir_interpolator_process(const float *input, float mu, mmse_fir_interpolator *interp);

void mmse_fir_interpolator_destroy(mmse_fir_interpolator *interp);

int mmse_fir_interpolator_taps(mmse_fir_interpolator *interp);

#endif
