/*
 * Host-side filter design and table upload. Runs once per *_create, so it stays on the CPU: using the same double
 * formulas and the same libm as the reference gives bit-identical float taps, which the parity contract needs.
 *
 *   sdrm_design_low_pass   Hamming-windowed sinc, unity DC gain (GNU Radio firdes::low_pass as used by the reference,
 *                          src/dsp/lpf_taps.c:14-103)
 *   sdrm_design_gaussian   Gaussian pulse taps (reference src/dsp/gaussian_taps.c:10-33)
 *   sdrm_convolve_full     full linear convolution used to widen the Gaussian by one symbol
 *                          (reference src/dsp/gfsk_mod.c:17-41)
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "sdrm_internal.h"
#include "tables_data.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

int sdrm_dev_zalloc(void **p, size_t bytes) {
    *p = NULL;
    if (bytes == 0) {
        bytes = 16;
    }
    SDRM_CUDA_TRY(cudaMalloc(p, bytes));
    SDRM_CUDA_TRY(cudaMemset(*p, 0, bytes));
    /* cudaMemset of device memory is asynchronous and runs on the legacy default stream, which the library's non-blocking
     * streams do not wait for: a buffer allocated lazily (the staging buffers of the first host-buffer call) could be zeroed
     * AFTER the first copy into it had landed. Seen as one wrong symbol count on the first call of a shard when a second
     * device's context was being brought up at the same time; the fill is finished before the pointer is handed out. */
    SDRM_CUDA_TRY(cudaStreamSynchronize(0));
    return 0;
}

static int low_pass_tap_count(uint64_t sampling_freq, uint64_t transition_width) {
    /* Hamming window: 53 dB stop band -> 53 * fs / (22 * tw) taps, forced odd (lpf_taps.c:33-40) */
    int count = (int) (53.0 * (double) sampling_freq / (22.0 * (double) transition_width));
    return count | 1;
}

int sdrm_design_low_pass(float gain, uint64_t sampling_freq, uint64_t cutoff_freq, uint32_t transition_width,
                         float **taps, size_t *len) {
    /* parameter checks and messages as lpf_taps.c:14-31 */
    if (sampling_freq == 0) {
        SDRM_LOG_ERROR("sampling frequency should be positive");
        return -1;
    }
    if (cutoff_freq == 0 || (double) cutoff_freq > (double) sampling_freq / 2) {
        SDRM_LOG_ERROR("cutoff frequency should be positive and less than sampling freq / 2. got: %llu",
                       (unsigned long long) cutoff_freq);
        return -1;
    }
    if (transition_width == 0) {
        SDRM_LOG_ERROR("transition width should be positive");
        return -1;
    }
    const int count = low_pass_tap_count(sampling_freq, transition_width);
    const int half = (count - 1) / 2;
    float *h = calloc((size_t) count, sizeof(float));
    if (h == NULL) {
        return -ENOMEM;
    }
    const double omega_c = 2 * M_PI * (double) cutoff_freq / (double) sampling_freq;
    for (int k = 0; k < count; k++) {
        /* window value is rounded to float before it multiplies the sinc, as in the reference */
        const float window = (float) (0.54 - 0.46 * cos((2 * M_PI * k) / (count - 1)));
        const int n = k - half;
        if (n == 0) {
            h[k] = (float) (omega_c / M_PI * window);
        } else {
            h[k] = (float) (sin((double) n * omega_c) / (n * M_PI) * window);
        }
    }
    /* DC gain in float, centre tap plus twice the upper half, summed upwards */
    float dc = h[half];
    for (int n = 1; n <= half; n++) {
        dc += 2 * h[half + n];
    }
    const float scale = gain / dc;
    for (int k = 0; k < count; k++) {
        h[k] *= scale;
    }
    *taps = h;
    *len = (size_t) count;
    return 0;
}

int sdrm_design_gaussian(double gain, double samples_per_symbol, double bt, size_t taps_len, float **taps) {
    float *h = malloc(sizeof(float) * (taps_len == 0 ? 1 : taps_len));
    if (h == NULL) {
        return -ENOMEM;
    }
    const double dt = 1.0 / samples_per_symbol;
    const double s = 1.0 / (sqrt(log(2.0)) / (2 * M_PI * bt));
    double t = -0.5 * (double) taps_len;
    double total = 0;
    for (size_t i = 0; i < taps_len; i++) {
        t++;
        const double ts = s * dt * t;
        h[i] = (float) exp(-0.5 * ts * ts);
        total += h[i];
    }
    for (size_t i = 0; i < taps_len; i++) {
        h[i] = (float) (h[i] / total * gain);
    }
    *taps = h;
    return 0;
}

int sdrm_convolve_full(const float *x, size_t x_len, const float *y, size_t y_len, float **out, size_t *out_len) {
    const size_t n = x_len + y_len - 1;
    float *result = malloc(sizeof(float) * n);
    if (result == NULL) {
        return -ENOMEM;
    }
    for (size_t i = 0; i < n; i++) {
        /* sum_j y[j] * x[i - j], j ascending, x treated as zero outside [0, x_len); zero terms are added too */
        float sum = 0.0F;
        for (size_t j = 0; j < y_len && j <= i; j++) {
            const size_t k = i - j;
            const float xv = k < x_len ? x[k] : 0.0F;
            sum += y[j] * xv;
        }
        result[i] = sum;
    }
    *out = result;
    *out_len = n;
    return 0;
}

int sdrm_upload_taps_dup(const float *taps, size_t len, void **d_taps) {
    const size_t padded = ((len + 1) & ~(size_t) 1) + 2;
    float *dup = calloc(padded * 2, sizeof(float));
    if (dup == NULL) {
        return -ENOMEM;
    }
    for (size_t j = 0; j < len; j++) {
        const float h = taps[len - 1 - j]; /* reversed, as fir_filter.c:27 */
        dup[2 * j] = h;
        dup[2 * j + 1] = h;
    }
    int code = sdrm_dev_zalloc(d_taps, padded * 2 * sizeof(float));
    if (code == 0) {
        code = sdrm_cuda_code(cudaMemcpy(*d_taps, dup, padded * 2 * sizeof(float), cudaMemcpyHostToDevice), "taps upload");
    }
    free(dup);
    return code;
}

float *sdrm_host_taps_dup(const float *taps, size_t len) {
    float *dup = malloc((len == 0 ? 1 : len) * 2 * sizeof(float));
    if (dup == NULL) {
        return NULL;
    }
    for (size_t j = 0; j < len; j++) {
        const float h = taps[len - 1 - j]; /* reversed, as fir_filter.c:27 */
        dup[2 * j] = h;
        dup[2 * j + 1] = h;
    }
    return dup;
}

/*
 * How many Markstein corrections the tail kernel's division by `length` needs to round like an IEEE division:
 * q0 = RN(a * rcp), e = fma(-q, L, a), q' = fma(e, rcp, q). Inside the exponent range the kernel accepts nothing
 * underflows or overflows, so a and 2^k a behave alike and the 2^23 mantissas of one binade cover every operand (the
 * arithmetic is symmetric in sign). One correction is enough for every length tried (all that 32 * samples-per-symbol
 * gives for the reference's parameters); the check is what makes that a fact for the length at hand. Returns 1 or 2,
 * or 0 if even two corrections fail (the kernel then keeps __fdiv_rn, never seen).
 */
static int division_steps_uncached(int length) {
    const float L = (float) length;
    const float rcp = 1.0f / L;
    int one_ok = 1;
    int two_ok = 1;
    for (uint32_t m = 0; m < (1u << 23); m++) {
        union {
            uint32_t u;
            float f;
        } a;
        a.u = 0x3f800000u | m;
        const float exact = a.f / L;
        const float q0 = a.f * rcp;
        const float q1 = fmaf(fmaf(-q0, L, a.f), rcp, q0);
        if (q1 != exact) {
            one_ok = 0;
            const float q2 = fmaf(fmaf(-q1, L, a.f), rcp, q1);
            if (q2 != exact) {
                two_ok = 0;
                break;
            }
        }
    }
    return one_ok ? 1 : (two_ok ? 2 : 0);
}

int sdrm_division_steps(int length) {
    /* the check takes tens of milliseconds; handles are created by the hundred (one per session), lengths are few */
    static pthread_mutex_t lock = PTHREAD_MUTEX_INITIALIZER;
    static int known_length[16];
    static int known_steps[16];
    static int known = 0;
    if (length <= 0) {
        return 0;
    }
    pthread_mutex_lock(&lock);
    for (int i = 0; i < known; i++) {
        if (known_length[i] == length) {
            const int steps = known_steps[i];
            pthread_mutex_unlock(&lock);
            return steps;
        }
    }
    pthread_mutex_unlock(&lock);
    const int steps = division_steps_uncached(length);
    pthread_mutex_lock(&lock);
    if (known < 16) {
        known_length[known] = length;
        known_steps[known] = steps;
        known++;
    }
    pthread_mutex_unlock(&lock);
    return steps;
}

const float *sdrm_host_atan_table(void) { return (const float *) sdrm_atan_bits; }

const float *sdrm_host_mmse_table(void) { return (const float *) sdrm_mmse_bits; }

int sdrm_upload_atan_table(float **d_table) {
    int code = sdrm_dev_zalloc((void **) d_table, sizeof(sdrm_atan_bits));
    if (code != 0) {
        return code;
    }
    return sdrm_cuda_code(cudaMemcpy(*d_table, sdrm_atan_bits, sizeof(sdrm_atan_bits), cudaMemcpyHostToDevice), "atan table upload");
}

int sdrm_upload_mmse_table(float **d_table) {
    int code = sdrm_dev_zalloc((void **) d_table, sizeof(sdrm_mmse_bits));
    if (code != 0) {
        return code;
    }
    return sdrm_cuda_code(cudaMemcpy(*d_table, sdrm_mmse_bits, sizeof(sdrm_mmse_bits), cudaMemcpyHostToDevice), "mmse table upload");
}
