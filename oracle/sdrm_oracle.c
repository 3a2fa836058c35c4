/*
 * TEST INFRASTRUCTURE ONLY — see sdrm_oracle.h. CPU restatement of the reference's GMSK/FSK hot path.
 * Each function cites the reference lines it follows. Compile with -ffp-contract=off: every `a * b + c` below is
 * two roundings, as in the reference built for VOLK_GENERIC=1.
 */
#define _POSIX_C_SOURCE 200809L

#include "sdrm_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../sdr-modem_b200/host/tables_data.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define ORC_2PI ((float) (2 * M_PI))

/* ------------------------------------------------------------------------------------------------ tap design */

/* reference src/dsp/lpf_taps.c:14-103: Hamming-windowed sinc, normalised to `gain` at DC */
int orc_low_pass_taps(float gain, uint64_t fs, uint64_t cutoff, uint32_t transition_width, float **taps, size_t *len) {
    if (fs == 0 || cutoff == 0 || (double) cutoff > (double) fs / 2 || transition_width == 0) {
        return -1; /* lpf_taps.c:14-31 */
    }
    int ntaps = (int) (53.0 * (double) fs / (22.0 * (double) transition_width)); /* lpf_taps.c:33-40 */
    if ((ntaps & 1) == 0) {
        ntaps++;
    }
    float *h = malloc(sizeof(float) * (size_t) ntaps);
    if (h == NULL) {
        return -12;
    }
    const int m = ntaps - 1;
    const int half = m / 2;
    const double w0 = 2 * M_PI * (double) cutoff / (double) fs;
    for (int i = 0; i < ntaps; i++) {
        const float win = (float) (0.54 - 0.46 * cos((2 * M_PI * i) / m)); /* lpf_taps.c:42-53 */
        const int n = i - half;
        h[i] = n == 0 ? (float) (w0 / M_PI * win) : (float) (sin((double) n * w0) / (n * M_PI) * win); /* :82-89 */
    }
    float sum = h[half]; /* lpf_taps.c:93-96 */
    for (int n = 1; n <= half; n++) {
        sum += 2 * h[half + n];
    }
    gain /= sum;
    for (int i = 0; i < ntaps; i++) {
        h[i] *= gain;
    }
    *taps = h;
    *len = (size_t) ntaps;
    return 0;
}

/* reference src/dsp/gaussian_taps.c:10-33 */
int orc_gaussian_taps(double gain, double sps, double bt, size_t len, float **taps) {
    float *h = malloc(sizeof(float) * (len == 0 ? 1 : len));
    if (h == NULL) {
        return -12;
    }
    double scale = 0;
    const double dt = 1.0 / sps;
    const double s = 1.0 / (sqrt(log(2.0)) / (2 * M_PI * bt));
    double t0 = -0.5 * (double) len;
    for (size_t i = 0; i < len; i++) {
        t0++;
        const double ts = s * dt * t0;
        h[i] = (float) exp(-0.5 * ts * ts);
        scale += h[i];
    }
    for (size_t i = 0; i < len; i++) {
        h[i] = (float) (h[i] / scale * gain);
    }
    *taps = h;
    return 0;
}

/* reference src/dsp/gfsk_mod.c:17-41: out[i] = sum_j y[j] * xpad[i - j], j ascending */
int orc_convolve(const float *x, size_t x_len, const float *y, size_t y_len, float **out, size_t *out_len) {
    const size_t n = x_len + y_len - 1;
    float *r = malloc(sizeof(float) * n);
    if (r == NULL) {
        return -12;
    }
    for (size_t i = 0; i < n; i++) {
        float sum = 0.0F;
        for (size_t j = 0; j < y_len && j <= i; j++) {
            const float xv = (i - j) < x_len ? x[i - j] : 0.0F;
            sum += y[j] * xv;
        }
        r[i] = sum;
    }
    *out = r;
    *out_len = n;
    return 0;
}

/* ------------------------------------------------------------------------------------------------ atan2 */

/* reference src/math/fast_atan2f.c:87-157 */
float orc_fast_atan2f(float y, float x) {
    const float *table = (const float *) sdrm_atan_bits;
    const float y_abs = fabsf(y);
    const float x_abs = fabsf(x);
    if (!((y_abs > 0.0f) || (x_abs > 0.0f))) {
        return 0.0F;
    }
    const float z = y_abs < x_abs ? y_abs / x_abs : x_abs / y_abs;
    float base;
    if (z < 0.003921569) { /* double comparison, as the reference's TAN_MAP_RES macro */
        base = z;
    } else {
        float alpha = z * 255.0f;
        const int index = ((int) alpha) & 0xff;
        alpha -= (float) index;
        base = table[index];
        base += (table[index + 1] - table[index]) * alpha;
    }
    float angle;
    if (x_abs > y_abs) {
        if (x >= 0.0) {
            angle = y >= 0.0 ? base : -base;
        } else {
            angle = 3.14159265358979323846F;
            angle = y >= 0.0 ? angle - base : base - angle;
        }
    } else {
        if (y >= 0.0) {
            angle = 1.57079632679489661923F;
            angle = x >= 0.0 ? angle - base : angle + base;
        } else {
            angle = -1.57079632679489661923F;
            angle = x >= 0.0 ? angle + base : angle - base;
        }
    }
    return angle;
}

/* ------------------------------------------------------------------------------------------------ FIR */

struct orc_fir_t {
    float *rev; /* taps reversed (fir_filter.c:27) */
    size_t taps_len;
    int decimation;
    int width;
    float *hist;  /* last taps_len - 1 samples of the stream, oldest first (zeros at start, fir_filter.c:74) */
    size_t skip;  /* input samples to pass before the next output's newest sample */
    float *work;
    size_t work_cap;
};

orc_fir *orc_fir_create(const float *taps, size_t taps_len, int decimation, int width) {
    orc_fir *f = calloc(1, sizeof(*f));
    if (f == NULL || taps_len == 0 || decimation < 1) {
        free(f);
        return NULL;
    }
    f->rev = malloc(sizeof(float) * taps_len);
    f->hist = calloc(taps_len * (size_t) width, sizeof(float));
    if (f->rev == NULL || f->hist == NULL) {
        orc_fir_destroy(f);
        return NULL;
    }
    for (size_t j = 0; j < taps_len; j++) {
        f->rev[j] = taps[taps_len - 1 - j];
    }
    f->taps_len = taps_len;
    f->decimation = decimation;
    f->width = width;
    return f;
}

/*
 * Output k of the stream has its newest sample at stream index k * decimation and is emitted as soon as that sample
 * has arrived (fir_filter.c:100-106: i advances by decimation over history + input). Accumulation is sequential from
 * the oldest sample, one accumulator per component (VOLK generic dot products, fir_filter.c:102,132).
 */
size_t orc_fir_process(orc_fir *f, const float *in, size_t n_in, float *out) {
    const size_t t = f->taps_len;
    const size_t w = (size_t) f->width;
    const size_t need = (t - 1 + n_in) * w;
    if (need > f->work_cap) {
        float *grown = realloc(f->work, sizeof(float) * (need == 0 ? 1 : need));
        if (grown == NULL) {
            return 0;
        }
        f->work = grown;
        f->work_cap = need;
    }
    memcpy(f->work, f->hist, sizeof(float) * (t - 1) * w);
    memcpy(f->work + (t - 1) * w, in, sizeof(float) * n_in * w);
    size_t produced = 0;
    size_t i = f->skip; /* index, within `in`, of the next output's newest sample */
    for (; i < n_in; i += (size_t) f->decimation, produced++) {
        const float *window = f->work + i * w; /* oldest sample of this output */
        if (w == 1) {
            float acc = 0.0f;
            for (size_t j = 0; j < t; j++) {
                acc += window[j] * f->rev[j];
            }
            out[produced] = acc;
        } else {
            float re = 0.0f;
            float im = 0.0f;
            for (size_t j = 0; j < t; j++) {
                re += window[2 * j] * f->rev[j];
                im += window[2 * j + 1] * f->rev[j];
            }
            out[2 * produced] = re;
            out[2 * produced + 1] = im;
        }
    }
    f->skip = i - n_in;
    /* keep the last t - 1 samples */
    memmove(f->hist, f->work + n_in * w, sizeof(float) * (t - 1) * w);
    return produced;
}

void orc_fir_destroy(orc_fir *f) {
    if (f == NULL) {
        return;
    }
    free(f->rev);
    free(f->hist);
    free(f->work);
    free(f);
}

/* ------------------------------------------------------------------------------------------------ quad demod */

void orc_quad_demod_init(orc_quad_demod *q, float gain) {
    q->gain = gain;
    q->prev_re = 0.0f; /* quadrature_demod.c:44: working buffer starts zeroed */
    q->prev_im = 0.0f;
}

/* quadrature_demod.c:64-69: t = x[i] * conj(x[i-1]) in C99 complex arithmetic, out = gain * fast_atan2f(Im t, Re t) */
void orc_quad_demod_process(orc_quad_demod *q, const float *iq, size_t n, float *out) {
    float pr = q->prev_re;
    float pi = q->prev_im;
    for (size_t i = 0; i < n; i++) {
        const float ar = iq[2 * i];
        const float ai = iq[2 * i + 1];
        const float re = ar * pr + ai * pi;
        const float im = ai * pr - ar * pi;
        out[i] = q->gain * orc_fast_atan2f(im, re);
        pr = ar;
        pi = ai;
    }
    q->prev_re = pr;
    q->prev_im = pi;
}

/* ------------------------------------------------------------------------------------------------ dc blocker */

struct orc_dc_blocker_t {
    int length;
    float *line[4]; /* circular: the last `length` inputs of each moving average */
    float *xline;   /* circular: the last 2 * length - 2 inputs of the blocker */
    float sum[4];
    int pos;
    int xpos;
};

orc_dc_blocker *orc_dc_blocker_create(int length) {
    if (length < 2) {
        return NULL;
    }
    orc_dc_blocker *d = calloc(1, sizeof(*d));
    if (d == NULL) {
        return NULL;
    }
    d->length = length;
    for (int s = 0; s < 4; s++) {
        d->line[s] = calloc((size_t) length, sizeof(float));
    }
    d->xline = calloc((size_t) (2 * length - 2), sizeof(float));
    if (d->line[0] == NULL || d->line[1] == NULL || d->line[2] == NULL || d->line[3] == NULL || d->xline == NULL) {
        orc_dc_blocker_destroy(d);
        return NULL;
    }
    return d;
}

/*
 * dc_blocker.c:52-64: each moving average computes y = x - x[n-L] + y_prev (left to right) and returns y / L;
 * dc_blocker.c:105-119: four of them in cascade, output = x[n - (2L - 2)] - y4. The reference shifts its delay lines
 * with memmove every sample; circular indexing reads the same values.
 */
void orc_dc_blocker_process(orc_dc_blocker *d, float *data, size_t n) {
    const float length_f = (float) d->length;
    for (size_t i = 0; i < n; i++) {
        const float x = data[i];
        float v = x;
        for (int s = 0; s < 4; s++) {
            const float delayed = d->line[s][d->pos];
            d->line[s][d->pos] = v;
            const float y = v - delayed + d->sum[s];
            d->sum[s] = y;
            v = y / length_f;
        }
        const float xd = d->xline[d->xpos];
        d->xline[d->xpos] = x;
        data[i] = xd - v;
        d->pos = d->pos + 1 == d->length ? 0 : d->pos + 1;
        d->xpos = d->xpos + 1 == 2 * d->length - 2 ? 0 : d->xpos + 1;
    }
}

void orc_dc_blocker_destroy(orc_dc_blocker *d) {
    if (d == NULL) {
        return;
    }
    for (int s = 0; s < 4; s++) {
        free(d->line[s]);
    }
    free(d->xline);
    free(d);
}

/* ------------------------------------------------------------------------------------------------ clock recovery */

struct orc_clock_mm_t {
    float omega;
    float omega_mid;
    float omega_lim;
    float gain_omega;
    float mu;
    float gain_mu;
    float last_sample;
    float *work; /* carried samples followed by the new ones; work[0] is 16-byte aligned in the reference */
    size_t history;
    size_t max_len;
    size_t work_cap;
};

orc_clock_mm *orc_clock_mm_create(float omega, float gain_omega, float mu, float gain_mu, float omega_relative_limit,
                                  size_t max_len) {
    orc_clock_mm *c = calloc(1, sizeof(*c));
    if (c == NULL) {
        return NULL;
    }
    c->mu = mu; /* clock_recovery_mm.c:38-46 */
    c->omega = omega;
    c->gain_omega = gain_omega;
    c->gain_mu = gain_mu;
    c->omega_mid = omega;
    c->omega_lim = c->omega_mid * omega_relative_limit;
    c->last_sample = 0.0F;
    c->max_len = max_len;
    c->work_cap = 2 * max_len + 64;
    c->work = calloc(c->work_cap, sizeof(float));
    if (c->work == NULL) {
        free(c);
        return NULL;
    }
    return c;
}

/* mmse_fir_interpolator.c:188-191 + fir_filter.c:116-121: the dot product starts at the 16-byte aligned address at or
 * below the window, so (index & 3) earlier samples are multiplied by zero taps first; then sum_j in[j] * row[7 - j]. */
static float orc_interpolate(const float *work, size_t index, float mu) {
    const int imu = (int) rint(mu * 128);
    const float *row = (const float *) sdrm_mmse_bits[imu];
    const size_t lead = index & 3;
    float acc = 0.0f;
    for (size_t k = lead; k > 0; k--) {
        acc += work[index - k] * 0.0f;
    }
    for (int j = 0; j < 8; j++) {
        acc += work[index + (size_t) j] * row[7 - j];
    }
    return acc;
}

static float orc_slice(float x) { return x < 0 ? -1.0F : 1.0F; }

static float orc_clip(float x, float clip) { return 0.5F * (fabsf(x + clip) - fabsf(x - clip)); }

/* clock_recovery_mm.c:78-139 */
size_t orc_clock_mm_process(orc_clock_mm *c, const float *in, size_t n_in, float *out) {
    if (n_in > c->max_len || c->history + n_in > c->work_cap) {
        return 0;
    }
    memcpy(c->work + c->history, in, sizeof(float) * n_in);
    const size_t working_len = c->history + n_in;
    if (working_len < 8) { /* :94-99 */
        c->history = working_len;
        return 0;
    }
    const size_t max_index = working_len - 7;
    int ii = 0;
    int oo = 0;
    int previous = 0;
    while ((size_t) ii < max_index && (size_t) oo < c->max_len) { /* int vs size_t comparison as in the reference */
        float o = orc_interpolate(c->work, (size_t) ii, c->mu);
        if (isnan(o)) { /* :107-113 */
            out[oo] = 0.0f;
            previous = ii;
            ii += (int) floorf(c->omega);
            oo++;
            continue;
        }
        out[oo] = o;
        const float mm_val = orc_slice(c->last_sample) * o - orc_slice(o) * c->last_sample;
        c->last_sample = o;
        previous = ii;
        c->omega = c->omega + c->gain_omega * mm_val;
        c->omega = c->omega_mid + orc_clip(c->omega - c->omega_mid, c->omega_lim);
        c->mu = c->mu + c->omega + c->gain_mu * mm_val;
        ii += (int) floorf(c->mu);
        c->mu = c->mu - floorf(c->mu);
        oo++;
    }
    const size_t last_index = (size_t) ii > working_len ? (size_t) previous : (size_t) ii; /* :127-133 */
    c->history = working_len - last_index;
    memmove(c->work, c->work + last_index, sizeof(float) * c->history);
    return (size_t) oo;
}

void orc_clock_mm_destroy(orc_clock_mm *c) {
    if (c == NULL) {
        return;
    }
    free(c->work);
    free(c);
}

/* VOLK volk_32f_s32f_convert_8i generic: scale, saturate, round half to even */
void orc_convert_8i(const float *in, float scale, size_t n, int8_t *out) {
    for (size_t i = 0; i < n; i++) {
        const float r = in[i] * scale;
        if (r > 127.0f) {
            out[i] = 127;
        } else if (r < -128.0f) {
            out[i] = -128;
        } else {
            out[i] = (int8_t) rintf(r);
        }
    }
}

/* VOLK volk_16i_s32f_convert_32f generic: out = (float) in / scalar. PlutoSDR ingest, reference src/sdr/plutosdr.c:129
 * (scalar 2048). "parity unpinned" beyond the reference's own KAT (test/test_plutosdr.c:149-154): plutosdr.c needs libiio
 * and is not part of oracle/_ref. */
void orc_convert_16i_32f(const int16_t *in, float scalar, size_t n, float *out) {
    for (size_t i = 0; i < n; i++) {
        out[i] = (float) in[i] / scalar;
    }
}

/* VOLK volk_32f_s32f_convert_16i generic: scale, saturate to [SHRT_MIN, SHRT_MAX], round half to even. PlutoSDR egress,
 * reference src/sdr/plutosdr.c:83 (scalar 32768); KAT test/test_plutosdr.c:192-194. NaN (unspecified in C) gives 0, which is
 * what the x86 conversion followed by the 16-bit truncation produces. */
void orc_convert_32f_16i(const float *in, float scalar, size_t n, int16_t *out) {
    for (size_t i = 0; i < n; i++) {
        float r = in[i] * scalar;
        if (r != r) {
            out[i] = 0;
            continue;
        }
        if (r > 32767.0f) {
            r = 32767.0f;
        } else if (r < -32768.0f) {
            r = -32768.0f;
        }
        out[i] = (int16_t) rintf(r);
    }
}

/* ------------------------------------------------------------------------------------------------ fsk demod */

struct orc_fsk_demod_t {
    orc_fir *lpf1;
    orc_fir *lpf2;
    orc_quad_demod qd;
    orc_dc_blocker *dc;
    orc_clock_mm *clock;
    float *a; /* lpf1 out */
    float *b; /* quad out */
    float *c; /* lpf2 / dc out */
    float *e; /* clock out */
    uint32_t max_len;
};

/* reference src/dsp/fsk_demod.c:28-78 */
orc_fsk_demod *orc_fsk_demod_create(uint64_t fs, uint32_t baud, int64_t deviation, uint8_t decimation, uint32_t tw,
                                    int use_dc, uint32_t max_len) {
    orc_fsk_demod *d = calloc(1, sizeof(*d));
    if (d == NULL || decimation == 0 || baud == 0) {
        free(d);
        return NULL;
    }
    float *taps = NULL;
    size_t len = 0;
    const double carson_cutoff = (double) llabs(deviation) + (double) baud / 2;
    if (orc_low_pass_taps(1.0F, fs, (uint64_t) carson_cutoff, (uint32_t) (0.1f * carson_cutoff), &taps, &len) != 0) goto fail;
    d->lpf1 = orc_fir_create(taps, len, 1, 2);
    free(taps);
    taps = NULL;
    orc_quad_demod_init(&d->qd, (float) ((double) fs / (2 * M_PI * (double) deviation)));
    if (orc_low_pass_taps(1.0F, fs, baud / 2, tw, &taps, &len) != 0) goto fail;
    d->lpf2 = orc_fir_create(taps, len, decimation, 1);
    free(taps);
    taps = NULL;
    const float sps = (float) ((double) fs / baud / decimation);
    if (use_dc) {
        d->dc = orc_dc_blocker_create((int) ceilf(sps * 32));
        if (d->dc == NULL) goto fail;
    }
    d->clock = orc_clock_mm_create(sps, (sps * (float) M_PI) / 100, 0.5f, 0.5f / 8.0f, 0.01f, max_len);
    d->max_len = max_len;
    d->a = malloc(sizeof(float) * 2 * (max_len + 1));
    d->b = malloc(sizeof(float) * (max_len + 1));
    d->c = malloc(sizeof(float) * (max_len + 1));
    d->e = malloc(sizeof(float) * (max_len + 1));
    if (d->lpf1 == NULL || d->lpf2 == NULL || d->clock == NULL || d->a == NULL || d->b == NULL || d->c == NULL || d->e == NULL) goto fail;
    return d;
fail:
    free(taps);
    orc_fsk_demod_destroy(d);
    return NULL;
}

/* reference src/dsp/fsk_demod.c:80-110 */
size_t orc_fsk_demod_process(orc_fsk_demod *d, const float *iq, size_t n, int8_t *hard, float *soft) {
    if (n > d->max_len) {
        return 0;
    }
    const size_t n1 = orc_fir_process(d->lpf1, iq, n, d->a);
    orc_quad_demod_process(&d->qd, d->a, n1, d->b);
    const size_t n2 = orc_fir_process(d->lpf2, d->b, n1, d->c);
    if (d->dc != NULL) {
        orc_dc_blocker_process(d->dc, d->c, n2);
    }
    const size_t n3 = orc_clock_mm_process(d->clock, d->c, n2, d->e);
    orc_convert_8i(d->e, 127.0f, n3, hard);
    if (soft != NULL) {
        memcpy(soft, d->e, sizeof(float) * n3);
    }
    return n3;
}

void orc_fsk_demod_destroy(orc_fsk_demod *d) {
    if (d == NULL) {
        return;
    }
    orc_fir_destroy(d->lpf1);
    orc_fir_destroy(d->lpf2);
    orc_dc_blocker_destroy(d->dc);
    orc_clock_mm_destroy(d->clock);
    free(d->a);
    free(d->b);
    free(d->c);
    free(d->e);
    free(d);
}

/* ------------------------------------------------------------------------------------------------ NCO */

void orc_sig_source_init(orc_sig_source *s, float amplitude, uint64_t fs) {
    s->phase = 0.0F;
    s->amplitude = amplitude;
    s->fs = fs;
}

/* sig_source.c:43-58: float phase accumulator, wrapped at +-2pi, trigonometry in double rounded to float */
void orc_sig_source_generate(orc_sig_source *s, int64_t freq, size_t n, float *out_iq) {
    const float step = ORC_2PI * (float) freq / s->fs;
    for (size_t i = 0; i < n; i++) {
        out_iq[2 * i] = (float) (cos(s->phase) * s->amplitude);
        out_iq[2 * i + 1] = (float) (sin(s->phase) * s->amplitude);
        s->phase += step;
        if (s->phase < -ORC_2PI) {
            s->phase += ORC_2PI;
        }
        if (s->phase > ORC_2PI) {
            s->phase -= ORC_2PI;
        }
    }
}

/* sig_source.c:60-75: out = in * nco, C99 complex product (volk_32fc_x2_multiply_32fc generic) */
void orc_sig_source_multiply(orc_sig_source *s, int64_t freq, const float *in_iq, size_t n, float *out_iq) {
    const float step = ORC_2PI * (float) freq / s->fs;
    for (size_t i = 0; i < n; i++) {
        const float cr = (float) (cos(s->phase) * s->amplitude);
        const float ci = (float) (sin(s->phase) * s->amplitude);
        s->phase += step;
        if (s->phase < -ORC_2PI) {
            s->phase += ORC_2PI;
        }
        if (s->phase > ORC_2PI) {
            s->phase -= ORC_2PI;
        }
        const float ar = in_iq[2 * i];
        const float ai = in_iq[2 * i + 1];
        out_iq[2 * i] = ar * cr - ai * ci;
        out_iq[2 * i + 1] = ar * ci + ai * cr;
    }
}

/* ------------------------------------------------------------------------------------------------ modulator */

/* frequency_modulator.c:41-60: the phase is advanced before the sample is produced */
void orc_freq_mod_process(orc_freq_mod *m, const float *in, size_t n, float *out_iq) {
    for (size_t i = 0; i < n; i++) {
        m->phase = m->phase + m->sensitivity * in[i];
        if (m->phase < -ORC_2PI) {
            m->phase += ORC_2PI;
        }
        if (m->phase > ORC_2PI) {
            m->phase -= ORC_2PI;
        }
        out_iq[2 * i] = (float) cos(m->phase);
        out_iq[2 * i + 1] = (float) sin(m->phase);
    }
}

struct orc_gfsk_mod_t {
    int interpolation;
    orc_fir **branch; /* polyphase branches h_p[k] = h[k * I + p] (interp_fir_filter.c:42-73) */
    orc_freq_mod fm;
    float *bits;
    float *shaped;
    float *branch_out;
    uint32_t max_bytes;
};

/* gfsk_mod.c:43-100, interp_fir_filter.c:75-137 */
orc_gfsk_mod *orc_gfsk_mod_create(float sps, float sensitivity, float bt, uint32_t max_bytes) {
    orc_gfsk_mod *m = calloc(1, sizeof(*m));
    if (m == NULL) {
        return NULL;
    }
    const size_t gauss_len = (size_t) (4 * sps);
    const size_t square_len = (size_t) (int) sps;
    float *gauss = NULL;
    float *square = malloc(sizeof(float) * (square_len == 0 ? 1 : square_len));
    float *taps = NULL;
    float *padded = NULL;
    size_t taps_len = 0;
    if (square == NULL || square_len == 0 || orc_gaussian_taps(1.0F, sps, bt, gauss_len, &gauss) != 0) goto fail;
    for (size_t i = 0; i < square_len; i++) {
        square[i] = 1.0F;
    }
    if (orc_convolve(gauss, gauss_len, square, square_len, &taps, &taps_len) != 0) goto fail;
    m->interpolation = (int) sps;
    /* zero-pad to a multiple of the interpolation (interp_fir_filter.c:19-40), then split into branches */
    const size_t inter = (size_t) m->interpolation;
    const size_t padded_len = (taps_len + inter - 1) / inter * inter;
    padded = calloc(padded_len, sizeof(float));
    m->branch = calloc(inter, sizeof(orc_fir *));
    if (padded == NULL || m->branch == NULL) goto fail;
    memcpy(padded, taps, sizeof(float) * taps_len);
    const size_t branch_len = padded_len / inter;
    for (size_t p = 0; p < inter; p++) {
        float *bt_taps = malloc(sizeof(float) * branch_len);
        if (bt_taps == NULL) goto fail;
        for (size_t k = 0; k < branch_len; k++) {
            bt_taps[k] = padded[k * inter + p];
        }
        m->branch[p] = orc_fir_create(bt_taps, branch_len, 1, 1);
        free(bt_taps);
        if (m->branch[p] == NULL) goto fail;
    }
    m->fm.phase = 0;
    m->fm.sensitivity = sensitivity;
    m->max_bytes = max_bytes;
    const size_t max_bits = (size_t) max_bytes * 8;
    m->bits = malloc(sizeof(float) * (max_bits + 1));
    m->branch_out = malloc(sizeof(float) * (max_bits + 1));
    m->shaped = malloc(sizeof(float) * (max_bits * inter + 1));
    if (m->bits == NULL || m->branch_out == NULL || m->shaped == NULL) goto fail;
    free(gauss);
    free(square);
    free(taps);
    free(padded);
    return m;
fail:
    free(gauss);
    free(square);
    free(taps);
    free(padded);
    orc_gfsk_mod_destroy(m);
    return NULL;
}

/* gfsk_mod.c:102-132: bytes -> +-1 MSB first, polyphase interpolation, frequency modulation */
size_t orc_gfsk_mod_process(orc_gfsk_mod *m, const uint8_t *bytes, size_t n_bytes, float *out_iq) {
    if (n_bytes > m->max_bytes) {
        return 0;
    }
    size_t n_bits = 0;
    for (size_t i = 0; i < n_bytes; i++) {
        for (int j = 0; j < 8; j++) {
            m->bits[n_bits++] = ((bytes[i] >> (7 - j)) & 1) ? 1.0F : -1.0F;
        }
    }
    const size_t inter = (size_t) m->interpolation;
    size_t total = 0;
    for (size_t p = 0; p < inter; p++) {
        const size_t got = orc_fir_process(m->branch[p], m->bits, n_bits, m->branch_out);
        total = inter * got;
        for (size_t k = 0; k < got; k++) {
            m->shaped[k * inter + p] = m->branch_out[k]; /* interp_fir_filter.c:147-149 */
        }
    }
    orc_freq_mod_process(&m->fm, m->shaped, total, out_iq);
    return total;
}

void orc_gfsk_mod_destroy(orc_gfsk_mod *m) {
    if (m == NULL) {
        return;
    }
    if (m->branch != NULL) {
        for (int p = 0; p < m->interpolation; p++) {
            orc_fir_destroy(m->branch[p]);
        }
        free(m->branch);
    }
    free(m->bits);
    free(m->shaped);
    free(m->branch_out);
    free(m);
}

/* ------------------------------------------------------------------------------------------------ CPU bench driver */

struct orc_slice_job {
    pthread_t thread;
    int n_channels;
    uint64_t fs;
    uint32_t baud;
    int64_t deviation;
    uint8_t decimation;
    uint32_t tw;
    int use_dc;
    uint32_t chunk;
    const float *iq;
    size_t stride;
    size_t n_samples;
    int passes;
    uint64_t symbols;
    int failed;
};

static void *orc_bench_thread(void *arg) {
    struct orc_slice_job *s = arg;
    orc_fsk_demod **demods = calloc((size_t) s->n_channels, sizeof(*demods));
    int8_t *hard = malloc(s->chunk + 1);
    if (demods == NULL || hard == NULL) {
        s->failed = 1;
        free(demods);
        free(hard);
        return NULL;
    }
    for (int c = 0; c < s->n_channels; c++) {
        demods[c] = orc_fsk_demod_create(s->fs, s->baud, s->deviation, s->decimation, s->tw, s->use_dc, s->chunk);
        if (demods[c] == NULL) {
            s->failed = 1;
        }
    }
    uint64_t symbols = 0;
    if (!s->failed) {
        for (int p = 0; p < s->passes; p++) {
            for (size_t off = 0; off < s->n_samples; off += s->chunk) {
                const size_t len = s->n_samples - off < s->chunk ? s->n_samples - off : s->chunk;
                for (int c = 0; c < s->n_channels; c++) {
                    symbols += orc_fsk_demod_process(demods[c], s->iq + (size_t) c * s->stride + 2 * off, len, hard, NULL);
                }
            }
        }
    }
    for (int c = 0; c < s->n_channels; c++) {
        orc_fsk_demod_destroy(demods[c]);
    }
    free(demods);
    free(hard);
    s->symbols = symbols;
    return NULL;
}

double orc_bench_fsk_demod(uint64_t fs, uint32_t baud, int64_t deviation, uint8_t decimation, uint32_t tw, int use_dc,
                           uint32_t chunk, const float *iq, size_t stride_floats, size_t n_samples, int n_channels,
                           int n_threads, int passes, uint64_t *symbols_out) {
    if (n_threads < 1 || n_channels < n_threads) {
        return -1.0;
    }
    struct orc_slice_job *jobs = calloc((size_t) n_threads, sizeof(*jobs));
    if (jobs == NULL) {
        return -1.0;
    }
    struct timespec t0;
    struct timespec t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int first = 0;
    for (int t = 0; t < n_threads; t++) {
        struct orc_slice_job *s = &jobs[t];
        s->n_channels = n_channels / n_threads + (t < n_channels % n_threads ? 1 : 0);
        s->fs = fs;
        s->baud = baud;
        s->deviation = deviation;
        s->decimation = decimation;
        s->tw = tw;
        s->use_dc = use_dc;
        s->chunk = chunk;
        s->iq = iq + (size_t) first * stride_floats;
        s->stride = stride_floats;
        s->n_samples = n_samples;
        s->passes = passes;
        first += s->n_channels;
        pthread_create(&s->thread, NULL, orc_bench_thread, s);
    }
    uint64_t symbols = 0;
    int failed = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(jobs[t].thread, NULL);
        symbols += jobs[t].symbols;
        failed |= jobs[t].failed;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(jobs);
    if (symbols_out != NULL) {
        *symbols_out = symbols;
    }
    return failed ? -1.0 : (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}

/* libm sweep (tests/test_gpu_sincos_sweep.py): (float) cos((double) x) and (float) sin((double) x) — what
 * frequency_modulator.c:56 and sig_source.c store for a float phase — for the floats with bit patterns
 * [first_bits, first_bits + count), compared bit for bit with `got` ((cos, sin) pairs computed elsewhere).
 * Returns the number of values that differ; *first_bad = the first pattern with a difference. */
size_t orc_sincos_sweep(uint32_t first_bits, size_t count, const float *got, uint32_t *first_bad) {
    size_t bad = 0;
    for (size_t i = 0; i < count; i++) {
        const uint32_t bits = first_bits + (uint32_t) i;
        float x;
        memcpy(&x, &bits, sizeof(x));
        const float c = (float) cos((double) x);
        const float s = (float) sin((double) x);
        const int differs = memcmp(&c, &got[2 * i], sizeof(float)) != 0 || memcmp(&s, &got[2 * i + 1], sizeof(float)) != 0;
        if (differs) {
            if (bad == 0 && first_bad != NULL) {
                *first_bad = bits;
            }
            bad++;
        }
    }
    return bad;
}
