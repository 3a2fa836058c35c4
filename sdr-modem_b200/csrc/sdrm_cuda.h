/*
 * Thin C ABI between the C host layer (sdr-modem_b200/host) and the sm_100a kernels (sdr-modem_b200/csrc).
 * Plain pointers and sizes only. Every launcher enqueues on `stream` (a cudaStream_t passed as void*)
 * and returns 0 or a negative errno-style code; none of them synchronises.
 *
 * Device data layouts used across stages ("rows" never share state):
 *   CF   complex stream rows       float2 (re, im)              row r at base + r * stride   (float2 units)
 *   PAIR two real channels per row float2 (ch 2p, ch 2p+1)      row p at base + p * stride   (float2 units)
 *   TC   time-major real           float  [time][n_ch_padded]   channel fastest (per-block handles)
 *   GTC  grouped time-major real   float  [group][time][32]     groups of 32 channels, so that 32 rows of one group
 *                                                               are 4 KB of contiguous memory (one TMA bulk copy)
 */
#ifndef SDRM_CUDA_H
#define SDRM_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum sdrm_fir_out_mode {
    SDRM_FIR_OUT_ROWS = 0,    /* float2 out[row * out_stride + m] */
    SDRM_FIR_OUT_QD_PAIR = 1, /* quadrature demod of the filter output, written into PAIR layout */
    SDRM_FIR_OUT_TC = 2       /* float2 (two channels) written into the GTC ring at row tc_head + m, channels 2 * row, 2 * row + 1 */
};

/*
 * Streaming decimating FIR with real taps over float2 rows (reference src/dsp/fir_filter.c:93-144).
 * The virtual input of a row is v[i], i in [-hist_len, n_in): `hist` for i < 0, `in` for i >= 0.
 * Output m (0 <= m < n_out) has newest sample i_m = phase + m * decimation and equals
 *   sum_{j < n_taps} taps_rev[j] * v[i_m - (n_taps - 1) + j]
 * accumulated sequentially from j = 0 with one accumulator per component (VOLK generic order).
 * taps_dup holds the reversed taps duplicated into float2 (h, h).
 */
typedef struct {
    const void *in;
    size_t in_stride;
    const void *hist;
    int hist_len; /* even, >= n_taps - 1 (+2 in QD_PAIR mode) */
    const void *taps_dup;
    int n_taps;
    int decimation;
    int phase;
    int n_in;
    int n_out;
    int rows;
    int fast; /* 0: separately rounded multiply and add (bit-exact vs the reference); 1: fused multiply-add */
    int out_mode;
    void *out;
    size_t out_stride;       /* ROWS: float2 per row; QD_PAIR: float2 per pair row; TC: unused */
    int tc_ring_rows;        /* TC only: power of two */
    long long tc_head;       /* TC only: absolute row of output 0 */
    float qd_gain;           /* QD_PAIR only */
    const float *atan_table; /* QD_PAIR only: 257 floats on the device */
    const void *h_taps_dup;  /* optional HOST copy of taps_dup (n_taps float2): short filters pass their taps as kernel parameters */
    void *acc_scratch;       /* optional device scratch of sdrm_cu_fir_scratch_bytes(): with it a filter of more than one tap block
                                runs in FMA mode as several launches with the taps in their parameters (decimation 1 or 2) */
} sdrm_fir_args;

#define SDRM_FIR_TAP_BLOCK 528 /* taps per block of the tile FIR (fir.cu kTapBlock) */

/* bytes of sdrm_fir_args.acc_scratch for `rows` rows of at most max_out outputs per call (quad-demod mode included) */
size_t sdrm_cu_fir_scratch_bytes(int rows, int max_out);

int sdrm_cu_fir(const sdrm_fir_args *args, void *stream);

/* hist_next[row][k] = v[n_in - hist_len + k]; hist and hist_next must not alias. */
int sdrm_cu_hist_update(const void *in, size_t in_stride, const void *hist, void *hist_next, int hist_len, int n_in,
                        int rows, void *stream);

/*
 * Quadrature demod of CF rows (reference src/dsp/quadrature_demod.c:57-73):
 * out[row][i] = gain * fast_atan2f(x[i] * conj(x[i-1])), prev[row] carries x[-1] and is updated.
 */
int sdrm_cu_quad_demod(const void *in, size_t in_stride, void *prev, float gain, const float *atan_table, float *out,
                       size_t out_stride, int n_in, int rows, void *stream);

/*
 * TC ring: the real-valued tail of the chain (lpf2 output -> dc blocker -> clock recovery) lives in a time-major
 * ring  float ring[ring_rows][tc_stride]  with ring_rows a power of two; absolute row t is stored at
 * t & (ring_rows - 1). Producers append at `head`, consumers may look back at rows they have not released yet,
 * so no history is ever copied.
 *
 * DC blocker over ring rows [head, head + n_rows), in place (reference src/dsp/dc_blocker.c:105-119).
 * One lane per channel, serial in time. delay: float [4 * length + (2 * length - 2)][n_ch] delay lines
 * (zeroed at create); sums: float [4][n_ch]; pos_l / pos_x: cursors into the L-slot and (2L-2)-slot delay lines,
 * identical for all channels and advanced by the caller (pos + n_rows modulo the line length).
 */
int sdrm_cu_dc_blocker(float *ring, size_t tc_stride, int ring_rows, long long head, int n_rows, int n_ch, int length,
                       float *delay, float *sums, int pos_l, int pos_x, int grouped, void *stream);

/*
 * Mueller & Mueller clock recovery + int8 conversion (reference src/dsp/clock_recovery_mm.c:78-139,
 * src/dsp/fsk_demod.c:106) over ring rows [head - state.history, head + n_rows). Outputs are channel-major.
 */
typedef struct {
    float mu;
    float omega;
    float last_sample;
    int history; /* samples carried from previous calls (reference history_offset) */
} sdrm_clock_state;

typedef struct {
    const float *ring;
    size_t tc_stride;
    int ring_rows;
    long long head;
    int n_rows;
    int n_ch;
    int max_history; /* rows guaranteed intact behind head */
    float omega_mid;
    float omega_lim;
    float gain_omega;
    float gain_mu;
    const float *mmse_taps; /* device: 129 * 8 floats, row-major, as published (not reversed) */
    sdrm_clock_state *state;
    float *soft_out;  /* [n_ch][out_stride] or NULL */
    int8_t *hard_out; /* [n_ch][out_stride] or NULL */
    size_t out_stride;
    uint32_t *out_len; /* [n_ch] */
    int max_out;       /* reference output_len: the loop stops after this many symbols */
    int *error_flag;   /* device int; bit 0 set if a channel's carried history exceeded max_history */
    int fast;          /* 1: the 8-tap interpolator dot product uses fused multiply-add */
    int grouped;       /* 0: ring[row][tc_stride]; 1: the grouped layout of the decimating FIR, ring[ch / 32][row][32] */
} sdrm_clock_args;

int sdrm_cu_clock_mm(const sdrm_clock_args *args, void *stream);

/*
 * Fused serial tail behind fsk_demod: dc blocker (optional) -> clock recovery -> int8 over ring rows
 * [head, head + n_rows), one lane per channel. The dc blocker's output stays in a per-lane shared-memory ring of
 * `ring_slots` samples; the samples the reference would carry in its working buffer are saved to / restored from
 * `carry` (float [ring_slots][delay_stride]) together with `state`.
 */
typedef struct {
    const float *rows; /* GTC ring written by the decimating FIR: float [n_groups][ring_rows][32] */
    int n_groups;      /* channel groups of 32 (delay_stride / 32) */
    int ring_rows;
    long long head;
    int n_rows;
    int n_ch;
    int dc_length;       /* 0: no dc blocker */
    int div_steps;       /* sdrm_division_steps(dc_length): corrections the branch-free division needs (1 or 2); 0 = use __fdiv_rn */
    float *delay;        /* float [4][n_groups][dc_length][32] (last L inputs of each moving average), then
                            float [n_groups][dx_length][32] (group delay line); zero-initialised */
    int dx_length;       /* >= 2 * dc_length - 2 + 8 pipeline steps (of 32 or 64 rows): the first stage writes ahead of the last */
    float *sums;         /* float [4][delay_stride] */
    size_t delay_stride; /* channels rounded up (row pitch of delay, sums and carry) */
    int pos_l;           /* rows processed so far modulo dc_length */
    int pos_x;           /* rows processed so far modulo dx_length */
    float omega_mid;
    float omega_lim;
    float gain_omega;
    float gain_mu;
    const float *mmse_taps;
    sdrm_clock_state *state;
    float *carry;
    int ring_slots; /* power of two >= 128; a lane may carry up to ring_slots - 80 samples between calls */
    float *soft_out;
    int8_t *hard_out;
    size_t out_stride;
    uint32_t *out_len;
    int max_out;
    int *error_flag; /* bit 0: carried samples exceeded the ring; bit 1: symbol capacity reached */
    int fast;
} sdrm_tail_args;

int sdrm_cu_demod_tail(const sdrm_tail_args *args, void *stream);

/* Self test (synchronous): counts sums for which the tail's branch-free division by `length` differs from an IEEE
 * division, over blocks * 256 * per_thread pseudo-random values. Used by tests only. */
/* (cos, sin) as the modulator stores them for the float bit patterns [first_bits, first_bits + count); h_cos_sin: 2 * count floats */
int sdrm_cu_selftest_sincos(uint32_t first_bits, size_t count, float *h_cos_sin);
int sdrm_cu_selftest_div(int length, int steps, uint32_t seed, int blocks, int per_thread, unsigned long long *mismatches);

/*
 * NCO / mixer (reference src/dsp/sig_source.c:43-75) over CF rows: out = in * amp * exp(j p), p advancing by
 * step[ch] per sample in float with the reference's +-2pi wrap. in == NULL generates the tone only.
 * phase_state[ch] carries p between calls; phases is scratch, float [n_ch][phase_stride].
 */
typedef struct {
    const void *in; /* float2 rows or NULL */
    size_t in_stride;
    void *out; /* float2 rows */
    size_t out_stride;
    const float *step;      /* [n_ch] phase increment per sample, computed in float on the host as the reference does */
    const float *amplitude; /* [n_ch] */
    float *phase_state;     /* [n_ch] */
    float *phases;          /* scratch */
    size_t phase_stride;
    int n;
    int n_ch;
} sdrm_nco_args;

int sdrm_cu_nco(const sdrm_nco_args *args, void *stream);

/*
 * Polyphase interpolating FIR (reference src/dsp/interp_fir_filter.c:139-154) over rows of bytes (unpacked to +-1 bits,
 * MSB first, src/dsp/gfsk_mod.c:109-120) or floats: out[ch][k * I + p] = scale * sum_j w[k + j] * taps_rev[p][j].
 * history: float [n_ch][branch_taps - 1], the last inputs of the previous calls (zeros at start); updated.
 */
typedef struct {
    const void *in; /* uint8 rows (in_is_bytes) or float rows */
    size_t in_stride;
    int in_is_bytes;
    int n_in; /* inputs per row: bits (8 * bytes) or floats */
    int n_ch;
    int interpolation;
    int branch_taps;
    const float *taps_rev; /* device, [interpolation][branch_taps], each branch reversed */
    float *history;
    int apply_scale;
    float scale;
    float *out;
    size_t out_stride; /* floats per row; with out_grouped: time steps per group (rows of the GTC layout) */
    int out_grouped;   /* 0: out[ch][m]; 1: GTC layout out[ch / 32][m][ch % 32], the input of sdrm_cu_freq_mod */
} sdrm_interp_args;

int sdrm_cu_interp_fir(const sdrm_interp_args *args, void *stream);

/*
 * Frequency modulator (reference src/dsp/frequency_modulator.c:41-60): phase = wrap(phase + work[m]) in float, in place,
 * then out[ch][m] = cos(phase) + j sin(phase) evaluated in double and rounded to float. `work` holds
 * sensitivity * input in the GTC layout float [ceil(n_ch / 32)][rows][32] (128-byte aligned, rows >= n) and is
 * overwritten with the phases. out: float2 rows.
 */
int sdrm_cu_freq_mod(float *work, size_t rows, float *phase_state, void *out, size_t out_stride, long long n, int n_ch,
                     void *stream);
/* the two halves of sdrm_cu_freq_mod, for callers that pipeline them on different streams */
int sdrm_cu_phase_walk(float *work, size_t rows, float *phase_state, long long n, int n_ch, void *stream);
int sdrm_cu_phase_to_iq(const float *work, size_t rows, void *out, size_t out_stride, long long n, int n_ch, void *stream);

/*
 * SDR sample formats (reference src/sdr/plutosdr.c:83,129, VOLK generic kernels): int16 (I, Q) pairs <-> float2 rows.
 * Strides in complex samples. i16 -> cf32: (float) v / scalar; cf32 -> i16: v * scalar, saturated, round half to even.
 */
int sdrm_cu_i16_to_cf32(const void *in, size_t in_stride, void *out, size_t out_stride, float scalar, int n, int rows,
                        void *stream);
int sdrm_cu_cf32_to_i16(const void *in, size_t in_stride, void *out, size_t out_stride, float scalar, int n, int rows,
                        void *stream);

/* channel rows float [n_ch][stride] <-> PAIR rows float2 [(n_ch + 1) / 2][stride]; an odd last channel pairs with zeros */
int sdrm_cu_rows_to_pairs(const float *in, size_t in_stride, void *out, size_t out_stride, int n, int n_ch, void *stream);
int sdrm_cu_pairs_to_rows(const void *in, size_t in_stride, float *out, size_t out_stride, int n, int n_ch, void *stream);

#ifdef __cplusplus
}
#endif

#endif
