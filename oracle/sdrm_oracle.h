/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the reference's GMSK/FSK hot path in plain C99.
 * Never linked into libsdrmodem_b200.so; only tests/, __graft_entry__.smoke() and bench.py's CPU arm use it.
 *
 * Pinned: tests/test_oracle.py checks every function here against the reference's golden vectors
 * (tests/golden, copied from reference test/resources) and, when oracle/_ref is built, bit-for-bit against the
 * reference's own sources compiled in place.
 *
 * The port is written in stream semantics (what each output sample is a function of) instead of the reference's
 * working-buffer bookkeeping; the arithmetic — operation order, separate roundings, libm calls — is the reference's.
 * Build: gcc -std=c99 -O2 -ffp-contract=off (oracle/Makefile). No FMA contraction anywhere.
 */
#ifndef SDRM_ORACLE_H
#define SDRM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

/* --- tap design (reference src/dsp/lpf_taps.c:14-103, src/dsp/gaussian_taps.c:10-33, src/dsp/gfsk_mod.c:17-41) --- */
int orc_low_pass_taps(float gain, uint64_t fs, uint64_t cutoff, uint32_t transition_width, float **taps, size_t *len);
int orc_gaussian_taps(double gain, double sps, double bt, size_t len, float **taps);
int orc_convolve(const float *x, size_t x_len, const float *y, size_t y_len, float **out, size_t *out_len);

/* --- table arctangent (reference src/math/fast_atan2f.c:87-157) --- */
float orc_fast_atan2f(float y, float x);

/* --- streaming decimating FIR, real taps, real (width 1) or complex (width 2) samples
 *     (reference src/dsp/fir_filter.c:8-33,93-144 with VOLK generic dot products) --- */
typedef struct orc_fir_t orc_fir;
orc_fir *orc_fir_create(const float *taps, size_t taps_len, int decimation, int width);
/* out must hold n_in / decimation + 1 samples; returns the number of output samples */
size_t orc_fir_process(orc_fir *f, const float *in, size_t n_in, float *out);
void orc_fir_destroy(orc_fir *f);

/* --- quadrature demod (reference src/dsp/quadrature_demod.c:57-73) --- */
typedef struct {
    float gain;
    float prev_re;
    float prev_im;
} orc_quad_demod;
void orc_quad_demod_init(orc_quad_demod *q, float gain);
void orc_quad_demod_process(orc_quad_demod *q, const float *iq, size_t n, float *out);

/* --- DC blocker (reference src/dsp/dc_blocker.c:35-64,105-119) --- */
typedef struct orc_dc_blocker_t orc_dc_blocker;
orc_dc_blocker *orc_dc_blocker_create(int length);
void orc_dc_blocker_process(orc_dc_blocker *d, float *data, size_t n); /* in place */
void orc_dc_blocker_destroy(orc_dc_blocker *d);

/* --- Mueller & Mueller clock recovery (reference src/dsp/clock_recovery_mm.c:28-139,
 *     src/dsp/mmse_fir_interpolator.c:188-191, src/dsp/fir_filter.c:116-121) --- */
typedef struct orc_clock_mm_t orc_clock_mm;
orc_clock_mm *orc_clock_mm_create(float omega, float gain_omega, float mu, float gain_mu, float omega_relative_limit,
                                  size_t max_len);
size_t orc_clock_mm_process(orc_clock_mm *c, const float *in, size_t n_in, float *out);
void orc_clock_mm_destroy(orc_clock_mm *c);

/* --- float -> int8 (VOLK volk_32f_s32f_convert_8i generic, reference src/dsp/fsk_demod.c:106) --- */
void orc_convert_8i(const float *in, float scale, size_t n, int8_t *out);

/* --- SDR sample formats (VOLK generic 16i <-> 32f, reference src/sdr/plutosdr.c:83,129) --- */
void orc_convert_16i_32f(const int16_t *in, float scalar, size_t n, float *out);
void orc_convert_32f_16i(const float *in, float scalar, size_t n, int16_t *out);

/* --- the whole chain (reference src/dsp/fsk_demod.c:28-110) --- */
typedef struct orc_fsk_demod_t orc_fsk_demod;
orc_fsk_demod *orc_fsk_demod_create(uint64_t fs, uint32_t baud, int64_t deviation, uint8_t decimation, uint32_t tw,
                                    int use_dc, uint32_t max_len);
/* hard / soft must hold max_len entries (soft may be NULL); returns the number of symbols */
size_t orc_fsk_demod_process(orc_fsk_demod *d, const float *iq, size_t n, int8_t *hard, float *soft);
void orc_fsk_demod_destroy(orc_fsk_demod *d);

/* --- NCO / mixer (reference src/dsp/sig_source.c:43-75) --- */
typedef struct {
    float phase;
    float amplitude;
    uint64_t fs;
} orc_sig_source;
void orc_sig_source_init(orc_sig_source *s, float amplitude, uint64_t fs);
void orc_sig_source_generate(orc_sig_source *s, int64_t freq, size_t n, float *out_iq);
void orc_sig_source_multiply(orc_sig_source *s, int64_t freq, const float *in_iq, size_t n, float *out_iq);

/* --- frequency modulator (reference src/dsp/frequency_modulator.c:41-60) --- */
typedef struct {
    float phase;
    float sensitivity;
} orc_freq_mod;
void orc_freq_mod_process(orc_freq_mod *m, const float *in, size_t n, float *out_iq);

/* --- GFSK modulator (reference src/dsp/gfsk_mod.c:43-132, src/dsp/interp_fir_filter.c:19-154) --- */
typedef struct orc_gfsk_mod_t orc_gfsk_mod;
orc_gfsk_mod *orc_gfsk_mod_create(float sps, float sensitivity, float bt, uint32_t max_bytes);
/* out_iq must hold 8 * n_bytes * (int) sps complex samples; returns the number of complex samples */
size_t orc_gfsk_mod_process(orc_gfsk_mod *m, const uint8_t *bytes, size_t n_bytes, float *out_iq);
void orc_gfsk_mod_destroy(orc_gfsk_mod *m);

/* CPU throughput driver used by bench.py when oracle/_ref is unavailable: n_threads pthreads, channels split evenly,
 * each looping orc_fsk_demod_process over its channels' chunks. Returns wall seconds, negative on failure. */
double orc_bench_fsk_demod(uint64_t fs, uint32_t baud, int64_t deviation, uint8_t decimation, uint32_t tw, int use_dc,
                           uint32_t chunk, const float *iq, size_t stride_floats, size_t n_samples, int n_channels,
                           int n_threads, int passes, uint64_t *symbols_out);

/* libm sweep over float bit patterns [first_bits, first_bits + count): counts the (cos, sin) pairs in `got` that differ from
 * (float) cos((double) x), (float) sin((double) x) */
size_t orc_sincos_sweep(uint32_t first_bits, size_t count, const float *got, uint32_t *first_bad);

#endif
