/*
 * TEST INFRASTRUCTURE ONLY. The helpers of the reference's test/utils.h that its DSP tests use (input ramps, array
 * asserts, file reads), written against oracle/shim/check.h. The reference's own test/utils.c cannot be built here
 * because its request builders need libprotobuf-c; semantics follow reference test/utils.c:104-187.
 */
#include <check.h>
#include <complex.h>
#include <errno.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <volk/volk.h>

void setup_input_data(float **input, size_t input_offset, size_t len) {
    float *result = malloc(sizeof(float) * len);
    ck_assert(result != NULL);
    for (size_t i = 0; i < len; i++) {
        result[i] = (float) (input_offset + i);
    }
    *input = result;
}

void setup_volk_input_data(float **input, size_t input_offset, size_t len) {
    float *result = volk_malloc(sizeof(float) * len, volk_get_alignment());
    ck_assert(result != NULL);
    for (size_t i = 0; i < len; i++) {
        result[i] = (float) (input_offset + i);
    }
    *input = result;
}

void setup_input_complex_data(float complex **input, size_t input_offset, size_t len) {
    float complex *result = malloc(sizeof(float complex) * len);
    ck_assert(result != NULL);
    for (size_t i = 0; i < len; i++) {
        result[i] = (float) (2 * input_offset + 2 * i) + (float) (2 * input_offset + 2 * i + 1) * I;
    }
    *input = result;
}

void assert_complex_array(const float expected[], size_t expected_size, float complex *actual, size_t actual_size) {
    ck_assert_int_eq(expected_size, actual_size);
    for (size_t i = 0, j = 0; i < expected_size * 2; i += 2, j++) {
        ck_assert(fabsl(expected[i] - crealf(actual[j])) < 0.01);
        ck_assert(fabsl(expected[i + 1] - cimagf(actual[j])) < 0.01);
    }
}

void assert_int16_array(const int16_t expected[], size_t expected_size, int16_t *actual, size_t actual_size) {
    ck_assert_int_eq(expected_size, actual_size);
    for (size_t i = 0; i < expected_size; i++) {
        ck_assert_int_eq(expected[i], actual[i]);
    }
}

void assert_float_array(const float expected[], size_t expected_size, float *actual, size_t actual_size) {
    ck_assert_int_eq(expected_size, actual_size);
    for (size_t i = 0; i < expected_size; i++) {
        ck_assert(fabsl(expected[i] - actual[i]) < 0.001);
    }
}

void assert_byte_array(const int8_t expected[], size_t expected_size, int8_t *actual, size_t actual_size, int tolerance) {
    ck_assert_int_eq(expected_size, actual_size);
    for (size_t i = 0; i < expected_size; i++) {
        ck_assert(abs((int8_t) expected[i] - actual[i]) <= tolerance);
    }
}

int read_data(uint8_t *output, size_t *output_len, size_t len, FILE *file) {
    size_t left = len;
    int result = 0;
    while (left > 0) {
        size_t received = fread(output + (len - left), sizeof(uint8_t), left, file);
        if (received == 0) {
            result = -1;
            break;
        }
        left -= received;
    }
    *output_len = len - left;
    return result;
}
