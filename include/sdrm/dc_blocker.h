/* Drop-in for reference src/dsp/dc_blocker.h:7-11. Processing is in place: *output == input. */
#ifndef SDRM_DC_BLOCKER_H
#define SDRM_DC_BLOCKER_H

#include <stdlib.h>

typedef struct dc_blocker_t dc_blocker;

int dc_blocker_create(int length, dc_blocker **blocker);

void dc_blocker_process(float *input, size_t input_len, float **output, size_t *output_len, dc_blocker *blocker);

void dc_blocker_destroy(dc_blocker *blocker);

#endif
