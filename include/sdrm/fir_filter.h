/* Drop-in for reference src/dsp/fir_filter.h:9-35: streaming decimating FIR with real taps over complex (num_bytes 8)
 * or real (num_bytes 4) samples. `taps` is a malloc'ed array in design order that the filter owns after a successful
 * create (fir_filter.c:58,171-173).
 *
 * The reference declares struct fir_filter_t in its public header (fir_filter.h:9-27), so the layout is part of the ABI:
 * the same fields, in the same order, are declared here and are the first bytes of every handle this library returns.
 * Fields that describe the reference's HOST working buffers have no counterpart (the history and the work buffers
 * live in device memory) and read as zero / NULL; the others hold what the reference would put there. */
#ifndef SDRM_FIR_FILTER_H
#define SDRM_FIR_FILTER_H

#include <stdint.h>
#include <stdlib.h>

typedef struct fir_filter_t fir_filter;

struct fir_filter_t {
    uint8_t decimation;

    float **taps;            /* one reversed copy (alignment offset 0): taps[0][j] = original_taps[taps_len - 1 - j] */
    size_t aligned_taps_len; /* 1 */
    size_t alignment;        /* 16, the shim's volk_get_alignment() */
    size_t taps_len;
    float *original_taps;    /* the caller's malloc, owned by the filter */

    void *working_buffer;    /* NULL: history + input live on the device */
    size_t history_offset;   /* taps_len - 1, as after fir_filter_create */
    size_t working_len_total;
    void *volk_output;       /* NULL */
    size_t max_input_buffer_length;

    void *output;            /* host buffer the process call returns */
    size_t output_len;       /* max_input_buffer_length / decimation + 1 */
    size_t num_bytes;
};

int fir_filter_create(uint8_t decimation, float *taps, size_t taps_len, size_t output_len, size_t num_bytes, fir_filter **filter);

void fir_filter_process(const void *input, size_t input_len, void **output, size_t *output_len, fir_filter *filter);

/* one dot product over taps_len host floats starting at input, no history (used by the MMSE interpolator) */
float fir_filter_process_float_single(const float *input, fir_filter *filter);

void fir_filter_destroy(fir_filter *filter);

#endif
