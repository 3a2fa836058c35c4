/* Drop-in for reference src/dsp/gfsk_mod.h:10-17. */
#ifndef SDRM_GFSK_MOD_H
#define SDRM_GFSK_MOD_H

#include <complex.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct gfsk_mod_t gfsk_mod;

int gfsk_mod_create(float samplesPerSymbol, float sensitivity, float bt, uint32_t max_input_buffer_length, gfsk_mod **mod);

void gfsk_mod_process(const uint8_t *input, size_t input_len, float complex **output, size_t *output_len, gfsk_mod *mod);

/* used in tests */
int gfsk_mod_convolve(float *x, size_t x_len, float *y, size_t y_len, float **out, size_t *out_len);

void gfsk_mod_destroy(gfsk_mod *mod);

#endif
