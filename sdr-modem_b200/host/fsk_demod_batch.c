/*
 * Batched GMSK/FSK demodulator: N channel sessions, one parameter set, one set of kernel launches per call.
 *
 * Stands for N x { fsk_demod_create / fsk_demod_process / fsk_demod_destroy } of the reference
 * (src/dsp/fsk_demod.c:28-135) as driven per client by src/dsp_worker.c:44-106. Parameter derivation follows
 * fsk_demod_create line by line (Carson cutoff, quad gain, sps, dc length, clock gains); the stream state the
 * reference keeps inside its block structs lives here as flat per-channel device arrays:
 *
 *   lpf1      history  float2 [n_ch][T1+1]                 (fir_filter working_buffer head, fir_filter.c:107-110)
 *   quad      nothing: the previous lpf1 output is recomputed from the history (quadrature_demod.c:64-69)
 *   lpf2      history  float2 [n_pairs][T2-1 rounded up]   + decimation phase
 *   dc        four moving-average delay lines + group delay line + four running sums (dc_blocker.c:7-23)
 *   clock     mu, omega, last_sample, carried sample count (clock_recovery_mm.c:9-26), samples stay in the TC ring
 *
 * Two streams: `fir` (H2D, lpf1+quad, lpf2) and `tail` (dc blocker, clock recovery, D2H). The tail of call k runs
 * while the filters of call k+1 run; SLOTS calls may be in flight.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/sdrm/sdrm_batch.h"
#include "sdrm_internal.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define SLOTS 3
/* the filters may run this many calls ahead of the serial tail; the lpf2 output ring is RING_LAG + 1 calls deep */
#define RING_LAG 2

struct sdrm_fsk_demod_batch_t {
    sdrm_fsk_demod_batch_config cfg;
    int device;
    uint32_t n_ch;
    uint32_t n_pairs;
    uint32_t n_ch_pad;
    int fast;
    int want_soft;
    uint32_t aid; /* measurement aids (sdrm_internal.h), never reachable through the public flags */

    /* lpf1 + quad */
    void *d_taps1;
    float *h_taps1; /* host copies of the (h, h) pairs */
    float *h_taps2;
    int t1;
    int hist1_len;
    void *d_hist1[2];
    int hist1_cur;
    float qd_gain;
    float *d_atan;

    /* lpf2 */
    void *d_taps2;
    int t2;
    int hist2_len;
    void *d_hist2[2];
    int hist2_cur;
    int phase2;
    void *d_q; /* PAIR layout: quad demod output of the current call */
    size_t q_stride;
    void *d_acc; /* filters of more than one tap block: accumulators between the launches of one filter (fir.cu) */

    /* TC ring */
    float *d_ring;
    uint32_t ring_rows;
    size_t tc_stride;
    long long head;
    uint32_t max_rows;
    int max_history;

    /* dc blocker */
    int dc_len;
    int div_steps; /* sdrm_division_steps(dc_len) */
    int dx_len;
    float *d_delay;
    float *d_sums;
    int pos_l;
    int pos_x;

    /* clock */
    float omega_mid;
    float omega_lim;
    float gain_omega;
    float gain_mu;
    float *d_mmse;
    sdrm_clock_state *d_clock;
    float *d_carry;
    int ring_slots;
    int unfused; /* samples per symbol beyond what the fused tail's shared-memory ring holds: dc blocker and clock loop run as
                    the two plain kernels of tail.cu, straight on the lpf2 output ring */
    int *d_error;

    /* staging + results, one set per slot */
    void *d_in[SLOTS];
    void *d_in16[SLOTS]; /* int16 staging of submit_i16 */
    size_t in_stride_dev;
    int8_t *d_hard[SLOTS];
    float *d_soft[SLOTS];
    uint32_t *d_out_len[SLOTS];
    size_t out_stride;
    size_t slot_bound[SLOTS]; /* most symbols the clock loop can produce from the rows of the call in this slot */
    uint32_t *h_counts;       /* pinned, one per channel: the symbol counts of the call being fetched */
    size_t last_fetch_columns; /* columns per channel the last fetch moved */
    cudaEvent_t ev_fir[SLOTS];
    cudaEvent_t ev_tail[SLOTS];
    cudaEvent_t ev_copy[SLOTS];
    cudaEvent_t ev_done[SLOTS];
    cudaEvent_t ev_time[SLOTS][6]; /* profiling: lpf1 start/end, lpf2 end, dc start/end, clock end */
    int profiling;
    int last_slot;
    uint64_t submitted;
    uint64_t fetched;

    cudaStream_t s_copy;
    cudaStream_t s_conv; /* int16 -> cf32 conversion of submit_i16: off the copy stream, so that the next call's copy does not
                            wait behind it */
    cudaEvent_t ev_h2d[SLOTS];
    cudaStream_t s_fir;
    cudaStream_t s_tail;
    cudaStream_t s_out;
    uint64_t launches;
};

static int set_device(const sdrm_fsk_demod_batch *b) { return sdrm_cuda_code(cudaSetDevice(b->device), "cudaSetDevice"); }

int sdrm_fsk_demod_batch_create(const sdrm_fsk_demod_batch_config *config, sdrm_fsk_demod_batch **batch) {
    if (config == NULL || batch == NULL || config->n_channels == 0 || config->decimation == 0 ||
        config->max_input_buffer_length == 0 || config->baud_rate == 0 || config->deviation == 0 ||
        (config->flags & ~(SDRM_FLAG_FAST_FMA | SDRM_FLAG_SOFT_OUT)) != 0) {
        return -1;
    }
    sdrm_fsk_demod_batch *b = calloc(1, sizeof(*b));
    if (b == NULL) {
        return -ENOMEM;
    }
    b->cfg = *config;
    b->n_ch = config->n_channels;
    b->n_pairs = (b->n_ch + 1) / 2;
    b->n_ch_pad = (uint32_t) sdrm_round_up(b->n_ch, 32); /* the tail moves rows of 32 channels with TMA bulk copies */
    b->fast = (config->flags & SDRM_FLAG_FAST_FMA) != 0;
    b->want_soft = (config->flags & SDRM_FLAG_SOFT_OUT) != 0;
    int code = 0;
    if (config->device >= 0) {
        b->device = config->device;
    } else {
        code = sdrm_cuda_code(cudaGetDevice(&b->device), "cudaGetDevice");
    }
    if (code == 0) {
        code = set_device(b);
    }
    if (code != 0) {
        free(b);
        return code;
    }

    const uint64_t fs = config->sampling_freq;
    const uint32_t max_len = config->max_input_buffer_length;
    float *taps = NULL;
    size_t taps_len = 0;

    /* lpf1: Carson bandwidth, transition 10 % of it (fsk_demod.c:36-37) */
    const double carson_cutoff = (double) llabs(config->deviation) + (double) config->baud_rate / 2;
    code = sdrm_design_low_pass(1.0F, fs, (uint64_t) carson_cutoff, (uint32_t) (0.1f * carson_cutoff), &taps, &taps_len);
    if (code != 0) goto fail;
    b->t1 = (int) taps_len;
    code = sdrm_upload_taps_dup(taps, taps_len, &b->d_taps1);
    b->h_taps1 = sdrm_host_taps_dup(taps, taps_len);
    free(taps);
    taps = NULL;
    if (code != 0) goto fail;
    /* T1 - 1 samples of filter history + 2 so that the quad demod's previous output can be recomputed; even */
    b->hist1_len = (int) sdrm_round_up((size_t) b->t1 + 1, 2);
    for (int i = 0; i < 2 && code == 0; i++) {
        code = sdrm_dev_zalloc(&b->d_hist1[i], (size_t) b->n_ch * b->hist1_len * 8);
    }
    if (code != 0) goto fail;
    b->qd_gain = (float) ((double) fs / (2 * M_PI * (double) config->deviation)); /* fsk_demod.c:42 */
    code = sdrm_upload_atan_table(&b->d_atan);
    if (code != 0) goto fail;

    /* lpf2 (fsk_demod.c:47) */
    code = sdrm_design_low_pass(1.0F, fs, config->baud_rate / 2, config->transition_width, &taps, &taps_len);
    if (code != 0) goto fail;
    b->t2 = (int) taps_len;
    code = sdrm_upload_taps_dup(taps, taps_len, &b->d_taps2);
    b->h_taps2 = sdrm_host_taps_dup(taps, taps_len);
    free(taps);
    taps = NULL;
    if (code != 0) goto fail;
    b->hist2_len = (int) sdrm_round_up((size_t) b->t2 - 1, 2);
    if (b->hist2_len == 0) {
        b->hist2_len = 2;
    }
    for (int i = 0; i < 2 && code == 0; i++) {
        code = sdrm_dev_zalloc(&b->d_hist2[i], (size_t) b->n_pairs * b->hist2_len * 8);
    }
    if (code != 0) goto fail;
    b->q_stride = sdrm_round_up(max_len, 2) + 2;
    code = sdrm_dev_zalloc(&b->d_q, (size_t) b->n_pairs * b->q_stride * 8);
    if (code != 0) goto fail;
    if (b->t1 > SDRM_FIR_TAP_BLOCK || (b->t2 > SDRM_FIR_TAP_BLOCK && config->decimation <= 2)) {
        /* the two filters run one after the other on one stream and share the scratch */
        code = sdrm_dev_zalloc(&b->d_acc, sdrm_cu_fir_scratch_bytes((int) b->n_ch, (int) max_len));
        if (code != 0) goto fail;
    }

    /* clock + dc parameters (fsk_demod.c:53-66) */
    const float sps = (float) ((double) fs / config->baud_rate / config->decimation);
    b->omega_mid = sps;
    b->omega_lim = sps * 0.01f;
    b->gain_omega = (sps * (float) M_PI) / 100;
    b->gain_mu = 0.5f / 8.0f;
    b->max_rows = max_len / config->decimation + 1;
    b->max_history = (int) (4.0f * sps) + 64;
    b->ring_rows = sdrm_next_pow2((uint64_t) (RING_LAG + 1) * b->max_rows + (uint64_t) b->max_history);
    b->tc_stride = b->n_ch_pad;
    code = sdrm_dev_zalloc((void **) &b->d_ring, (size_t) b->ring_rows * b->tc_stride * sizeof(float));
    if (code != 0) goto fail;
    /* per-lane shared-memory ring of the fused tail: one symbol step + two pipeline steps of 64 rows + slack, power of two */
    b->ring_slots = (int) sdrm_next_pow2((uint64_t) ceilf(sps * 1.01f) + 8 + 2 * 64 + 16 + 8);
    if (b->ring_slots < 128) {
        b->ring_slots = 128;
    }
    if (b->ring_slots > 1024) {
        /* more than ~855 samples per symbol (the reference accepts any): no shared-memory ring, the two-kernel tail */
        b->unfused = 1;
        b->ring_slots = 128;
    }
    if (config->use_dc_block) {
        b->dc_len = (int) ceilf(sps * 32);
        if (b->dc_len < 2) {
            code = -1;
            goto fail;
        }
        if (b->dc_len < 32) {
            SDRM_LOG_ERROR("dc blocker length %d (fewer than one sample per symbol) is not supported", b->dc_len);
            code = -1;
            goto fail;
        }
        /* group delay line: 2L - 2 rows of delay + the rows the pipeline's first stage writes ahead of its last (4 steps of up
         * to 64 rows), with the same again as head room */
        b->dx_len = b->unfused ? 2 * b->dc_len - 2 : 2 * b->dc_len - 2 + 512;
        b->div_steps = sdrm_division_steps(b->dc_len);
        code = sdrm_dev_zalloc((void **) &b->d_delay, ((size_t) 4 * b->dc_len + b->dx_len) * b->n_ch_pad * sizeof(float));
        if (code != 0) goto fail;
        code = sdrm_dev_zalloc((void **) &b->d_sums, (size_t) 4 * b->n_ch_pad * sizeof(float));
        if (code != 0) goto fail;
    }
    code = sdrm_upload_mmse_table(&b->d_mmse);
    if (code != 0) goto fail;
    code = sdrm_dev_zalloc((void **) &b->d_clock, (size_t) b->n_ch_pad * sizeof(sdrm_clock_state));
    if (code != 0) goto fail;
    {
        sdrm_clock_state *init = calloc(b->n_ch_pad, sizeof(sdrm_clock_state));
        if (init == NULL) {
            code = -ENOMEM;
            goto fail;
        }
        for (uint32_t c = 0; c < b->n_ch_pad; c++) {
            init[c].mu = 0.5f; /* clock_mm_create(sps, gain_omega, 0.5, …), fsk_demod.c:61 */
            init[c].omega = sps;
            init[c].last_sample = 0.0f;
            init[c].history = 0;
        }
        code = sdrm_cuda_code(cudaMemcpy(b->d_clock, init, (size_t) b->n_ch_pad * sizeof(sdrm_clock_state), cudaMemcpyHostToDevice),
                              "clock state upload");
        free(init);
        if (code != 0) goto fail;
    }
    code = sdrm_dev_zalloc((void **) &b->d_carry, (size_t) b->ring_slots * b->n_ch_pad * sizeof(float));
    if (code != 0) goto fail;
    code = sdrm_dev_zalloc((void **) &b->d_error, sizeof(int));
    if (code != 0) goto fail;
    code = sdrm_cuda_code(cudaHostAlloc((void **) &b->h_counts, (size_t) b->n_ch_pad * sizeof(uint32_t), cudaHostAllocPortable), "pinned counts");
    if (code != 0) goto fail;

    b->out_stride = config->max_symbols_per_call != 0 ? config->max_symbols_per_call : max_len;
    b->out_stride = sdrm_round_up(b->out_stride, 16);
    b->in_stride_dev = sdrm_round_up(max_len, 2);
    for (int s = 0; s < SLOTS; s++) {
        code = sdrm_dev_zalloc((void **) &b->d_hard[s], (size_t) b->n_ch * b->out_stride);
        if (code != 0) goto fail;
        if (b->want_soft) {
            code = sdrm_dev_zalloc((void **) &b->d_soft[s], (size_t) b->n_ch * b->out_stride * sizeof(float));
            if (code != 0) goto fail;
        }
        code = sdrm_dev_zalloc((void **) &b->d_out_len[s], (size_t) b->n_ch_pad * sizeof(uint32_t));
        if (code != 0) goto fail;
        code = sdrm_cuda_code(cudaEventCreateWithFlags(&b->ev_fir[s], cudaEventDisableTiming), "event");
        if (code != 0) goto fail;
        code = sdrm_cuda_code(cudaEventCreateWithFlags(&b->ev_tail[s], cudaEventDisableTiming), "event");
        if (code != 0) goto fail;
        code = sdrm_cuda_code(cudaEventCreateWithFlags(&b->ev_copy[s], cudaEventDisableTiming), "event");
        if (code != 0) goto fail;
        code = sdrm_cuda_code(cudaEventCreateWithFlags(&b->ev_done[s], cudaEventDisableTiming), "event");
        if (code != 0) goto fail;
    }
    {
        /* the copy stream also runs the int16 -> cf32 conversion of the NEXT call while this call's filters fill the GPU:
         * its few blocks must not queue behind them */
        int least = 0;
        int greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        code = sdrm_cuda_code(cudaStreamCreateWithPriority(&b->s_copy, cudaStreamNonBlocking, greatest), "stream");
        if (code != 0) goto fail;
        code = sdrm_cuda_code(cudaStreamCreateWithPriority(&b->s_conv, cudaStreamNonBlocking, greatest), "stream");
        if (code != 0) goto fail;
        for (int s = 0; s < SLOTS; s++) {
            code = sdrm_cuda_code(cudaEventCreateWithFlags(&b->ev_h2d[s], cudaEventDisableTiming), "event");
            if (code != 0) goto fail;
        }
    }
    code = sdrm_cuda_code(cudaStreamCreateWithFlags(&b->s_fir, cudaStreamNonBlocking), "stream");
    if (code != 0) goto fail;
    code = sdrm_cuda_code(cudaStreamCreateWithFlags(&b->s_out, cudaStreamNonBlocking), "stream");
    if (code != 0) goto fail;
    {
        /* the tail has little parallelism and a long critical path: let its blocks go first */
        int least = 0;
        int greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        code = sdrm_cuda_code(cudaStreamCreateWithPriority(&b->s_tail, cudaStreamNonBlocking, greatest), "stream");
        if (code != 0) goto fail;
    }
    code = sdrm_cuda_code(cudaDeviceSynchronize(), "create sync");
    if (code != 0) goto fail;
    *batch = b;
    return 0;
fail:
    free(taps);
    sdrm_fsk_demod_batch_destroy(b);
    return code;
}

/* Enqueues one call. d_in: device cf32 [n_ch][in_stride]. */
static int enqueue(sdrm_fsk_demod_batch *b, const void *d_in, size_t in_stride, size_t n_in, int slot, int wait_copy) {
    int code;
    if (wait_copy) {
        SDRM_CUDA_TRY(cudaStreamWaitEvent(b->s_fir, b->ev_copy[slot], 0));
    }
    if (b->profiling) {
        SDRM_CUDA_TRY(cudaEventRecord(b->ev_time[slot][0], b->s_fir));
    }
    /* lpf1 (complex, decimation 1) fused with the quadrature demod -> PAIR layout */
    sdrm_fir_args f1;
    memset(&f1, 0, sizeof(f1));
    f1.in = d_in;
    f1.in_stride = in_stride;
    f1.hist = b->d_hist1[b->hist1_cur];
    f1.hist_len = b->hist1_len;
    f1.taps_dup = b->d_taps1;
    f1.h_taps_dup = b->h_taps1;
    f1.acc_scratch = b->d_acc;
    f1.n_taps = b->t1;
    f1.decimation = 1;
    f1.phase = 0;
    f1.n_in = (int) n_in;
    f1.n_out = (int) n_in;
    f1.rows = (int) b->n_ch;
    f1.fast = b->fast;
    f1.out_mode = SDRM_FIR_OUT_QD_PAIR;
    f1.out = b->d_q;
    f1.out_stride = b->q_stride;
    f1.qd_gain = b->qd_gain;
    f1.atan_table = b->d_atan;
    code = sdrm_launch_code(sdrm_cu_fir(&f1, b->s_fir), "lpf1+quad");
    if (code != 0) return code;
    if (b->profiling) {
        SDRM_CUDA_TRY(cudaEventRecord(b->ev_time[slot][1], b->s_fir));
    }
    code = sdrm_launch_code(sdrm_cu_hist_update(d_in, in_stride, b->d_hist1[b->hist1_cur], b->d_hist1[b->hist1_cur ^ 1],
                                                b->hist1_len, (int) n_in, (int) b->n_ch, b->s_fir),
                            "lpf1 history");
    if (code != 0) return code;
    b->hist1_cur ^= 1;
    b->launches += 2;

    /* lpf2 (real, decimating) over channel pairs -> TC ring */
    const int dec = b->cfg.decimation;
    const int n_q = (int) n_in;
    const int n_rows = n_q > b->phase2 ? (n_q - b->phase2 + dec - 1) / dec : 0;
    /* the ring rows written now were last needed by the tail of call k - RING_LAG (its look-back included) */
    if (b->submitted >= RING_LAG) {
        SDRM_CUDA_TRY(cudaStreamWaitEvent(b->s_fir, b->ev_tail[(b->submitted - RING_LAG) % SLOTS], 0));
    }
    sdrm_fir_args f2;
    memset(&f2, 0, sizeof(f2));
    f2.in = b->d_q;
    f2.in_stride = b->q_stride;
    f2.hist = b->d_hist2[b->hist2_cur];
    f2.hist_len = b->hist2_len;
    f2.taps_dup = b->d_taps2;
    f2.h_taps_dup = b->h_taps2;
    f2.acc_scratch = b->d_acc;
    f2.n_taps = b->t2;
    f2.decimation = dec;
    f2.phase = b->phase2;
    f2.n_in = n_q;
    f2.n_out = n_rows;
    f2.rows = (int) b->n_pairs;
    f2.fast = b->fast;
    f2.out_mode = SDRM_FIR_OUT_TC;
    f2.out = b->d_ring;
    f2.out_stride = b->tc_stride;
    f2.tc_ring_rows = (int) b->ring_rows;
    f2.tc_head = b->head;
    code = sdrm_launch_code(sdrm_cu_fir(&f2, b->s_fir), "lpf2");
    if (code != 0) return code;
    if (b->profiling) {
        SDRM_CUDA_TRY(cudaEventRecord(b->ev_time[slot][2], b->s_fir));
    }
    code = sdrm_launch_code(sdrm_cu_hist_update(b->d_q, b->q_stride, b->d_hist2[b->hist2_cur], b->d_hist2[b->hist2_cur ^ 1],
                                                b->hist2_len, n_q, (int) b->n_pairs, b->s_fir),
                            "lpf2 history");
    if (code != 0) return code;
    b->hist2_cur ^= 1;
    b->phase2 = b->phase2 + n_rows * dec - n_q;
    b->launches += (n_rows > 0 ? 1 : 0) + 1;
    SDRM_CUDA_TRY(cudaEventRecord(b->ev_fir[slot], b->s_fir));

    /* serial tail on its own stream */
    SDRM_CUDA_TRY(cudaStreamWaitEvent(b->s_tail, b->ev_fir[slot], 0));
    if (b->profiling) {
        SDRM_CUDA_TRY(cudaEventRecord(b->ev_time[slot][3], b->s_tail));
        SDRM_CUDA_TRY(cudaEventRecord(b->ev_time[slot][4], b->s_tail));
    }
    sdrm_tail_args ca;
    memset(&ca, 0, sizeof(ca));
    ca.rows = b->d_ring;
    ca.n_groups = (int) (b->n_ch_pad / 32);
    ca.ring_rows = (int) b->ring_rows;
    ca.head = b->head;
    ca.n_rows = n_rows;
    ca.n_ch = (int) b->n_ch;
    ca.dc_length = b->dc_len;
    ca.div_steps = b->div_steps;
    ca.delay = b->d_delay;
    ca.sums = b->d_sums;
    ca.delay_stride = b->n_ch_pad;
    ca.dx_length = b->dx_len;
    ca.pos_l = b->pos_l;
    ca.pos_x = b->pos_x;
    ca.omega_mid = b->omega_mid;
    ca.omega_lim = b->omega_lim;
    ca.gain_omega = b->gain_omega;
    ca.gain_mu = b->gain_mu;
    ca.mmse_taps = b->d_mmse;
    ca.state = b->d_clock;
    ca.carry = b->d_carry;
    ca.ring_slots = b->ring_slots;
    ca.soft_out = b->d_soft[slot];
    ca.hard_out = b->d_hard[slot];
    ca.out_stride = b->out_stride;
    ca.out_len = b->d_out_len[slot];
    ca.max_out = (int) (b->cfg.max_symbols_per_call != 0 ? b->cfg.max_symbols_per_call : b->cfg.max_input_buffer_length);
    if (b->aid & SDRM_AID_NO_CLOCK_LOOP) {
        ca.max_out = 0; /* measurement aid: the clock loop never runs (results are invalid) */
    }
    ca.error_flag = b->d_error;
    ca.fast = b->fast;
    if (b->dc_len > 0) {
        b->pos_l = (int) (((long long) b->pos_l + n_rows) % b->dc_len);
        b->pos_x = (int) (((long long) b->pos_x + n_rows) % b->dx_len);
    }
    if (b->aid & SDRM_AID_NO_TAIL) {
        code = 0; /* measurement aid: filters only, no tail (results are meaningless) */
    } else if (b->unfused) {
        code = 0;
        if (b->dc_len > 0) {
            code = sdrm_launch_code(sdrm_cu_dc_blocker(b->d_ring, 32, (int) b->ring_rows, b->head, n_rows, (int) b->n_ch, b->dc_len,
                                                       b->d_delay, b->d_sums, ca.pos_l, ca.pos_x, 1, b->s_tail),
                                    "dc blocker");
            b->launches += 1;
        }
        if (code == 0) {
            sdrm_clock_args k;
            memset(&k, 0, sizeof(k));
            k.ring = b->d_ring;
            k.tc_stride = 32;
            k.grouped = 1;
            k.ring_rows = (int) b->ring_rows;
            k.head = b->head;
            k.n_rows = n_rows;
            k.n_ch = (int) b->n_ch;
            k.max_history = b->max_history;
            k.omega_mid = b->omega_mid;
            k.omega_lim = b->omega_lim;
            k.gain_omega = b->gain_omega;
            k.gain_mu = b->gain_mu;
            k.mmse_taps = b->d_mmse;
            k.state = b->d_clock;
            k.soft_out = b->d_soft[slot];
            k.hard_out = b->d_hard[slot];
            k.out_stride = b->out_stride;
            k.out_len = b->d_out_len[slot];
            k.max_out = ca.max_out;
            k.error_flag = b->d_error;
            k.fast = b->fast;
            code = sdrm_launch_code(sdrm_cu_clock_mm(&k, b->s_tail), "clock recovery");
        }
    } else {
        code = sdrm_launch_code(sdrm_cu_demod_tail(&ca, b->s_tail), "dc blocker + clock recovery");
    }
    if (code != 0) return code;
    b->launches += 1;
    if (b->profiling) {
        SDRM_CUDA_TRY(cudaEventRecord(b->ev_time[slot][5], b->s_tail));
    }
    b->last_slot = slot;
    /* k symbols move the loop's cursor by sum floor(mu + omega) >= k * (omega_mid - omega_lim) - 1 rows
     * (clock_recovery_mm.c:101-125), and it has this call's rows plus what it carried (less than 8 + 2 omega) to move over:
     * the fetch copies this many columns and looks at the counts before it trusts the bound */
    b->slot_bound[slot] = (size_t) (((double) n_rows + 16.0 + 2.0 * (b->omega_mid + b->omega_lim)) / (b->omega_mid - b->omega_lim)) + 4;
    if (b->aid & SDRM_AID_FETCH_BOUND_1) {
        b->slot_bound[slot] = 1;
    }
    SDRM_CUDA_TRY(cudaEventRecord(b->ev_tail[slot], b->s_tail));
    b->head += n_rows;
    b->submitted++;
    return 0;
}

static int check_len(const sdrm_fsk_demod_batch *b, size_t input_len) {
    if (input_len > b->cfg.max_input_buffer_length) {
        /* same message as the reference blocks (e.g. src/dsp/fir_filter.c:148) */
        SDRM_LOG_ERROR("requested buffer %zu is more than max: %u", input_len, b->cfg.max_input_buffer_length);
        return -1;
    }
    return 0;
}

int sdrm_fsk_demod_batch_process_device(sdrm_fsk_demod_batch *b, const void *d_input, size_t in_stride, size_t input_len) {
    if (b == NULL || d_input == NULL || check_len(b, input_len) != 0) {
        return -1;
    }
    if (b->submitted - b->fetched >= SLOTS) {
        SDRM_LOG_ERROR("too many calls in flight: fetch results first");
        return -EBUSY;
    }
    int code = set_device(b);
    if (code != 0) return code;
    return enqueue(b, d_input, in_stride, input_len, (int) (b->submitted % SLOTS), 0);
}

static int ensure_staging(sdrm_fsk_demod_batch *b, int slot) {
    if (b->d_in[slot] != NULL) {
        return 0;
    }
    return sdrm_dev_zalloc(&b->d_in[slot], (size_t) b->n_ch * b->in_stride_dev * 8);
}

/* host rows that can move with one linear copy: even stride (16-byte aligned device rows), little padding, fits staging */
static int packed_rows(const sdrm_fsk_demod_batch *b, size_t in_stride, size_t input_len) {
    return (in_stride & 1) == 0 && in_stride <= b->in_stride_dev && in_stride >= input_len &&
           in_stride - input_len <= input_len / 16;
}

int sdrm_fsk_demod_batch_submit(sdrm_fsk_demod_batch *b, const float complex *input, size_t in_stride, size_t input_len) {
    if (b == NULL || (input == NULL && input_len > 0) || check_len(b, input_len) != 0) {
        return -1;
    }
    if (b->submitted - b->fetched >= SLOTS) {
        SDRM_LOG_ERROR("too many calls in flight: fetch results first");
        return -EBUSY;
    }
    int code = set_device(b);
    if (code != 0) return code;
    const int slot = (int) (b->submitted % SLOTS);
    code = ensure_staging(b, slot);
    if (code != 0) return code;
    size_t dev_stride = b->in_stride_dev;
    if (input_len > 0) {
        /* staging buffer of call k - SLOTS was last read by its filters */
        if (b->submitted >= SLOTS) {
            SDRM_CUDA_TRY(cudaStreamWaitEvent(b->s_copy, b->ev_fir[slot], 0));
        }
        if (packed_rows(b, in_stride, input_len)) {
            /* rows (nearly) back to back: one linear copy keeps the host layout and runs at full PCIe rate (55.6 GB/s
             * against 49.9 GB/s for the pitched copy, tools/microbench/h2d_probe.py) */
            SDRM_CUDA_TRY(cudaMemcpyAsync(b->d_in[slot], input, ((size_t) (b->n_ch - 1) * in_stride + input_len) * 8,
                                          cudaMemcpyHostToDevice, b->s_copy));
            dev_stride = in_stride;
        } else {
            SDRM_CUDA_TRY(cudaMemcpy2DAsync(b->d_in[slot], b->in_stride_dev * 8, input, in_stride * 8, input_len * 8, b->n_ch,
                                            cudaMemcpyHostToDevice, b->s_copy));
        }
    }
    SDRM_CUDA_TRY(cudaEventRecord(b->ev_copy[slot], b->s_copy));
    return enqueue(b, b->d_in[slot], dev_stride, input_len, slot, 1);
}

/* int16 IQ from the SDR (reference src/sdr/plutosdr.c:129 converts on the host): half the PCIe bytes, converted on the device */
int sdrm_fsk_demod_batch_submit_i16(sdrm_fsk_demod_batch *b, const int16_t *input, size_t in_stride, size_t input_len, float scalar) {
    if (b == NULL || (input == NULL && input_len > 0) || check_len(b, input_len) != 0) {
        return -1;
    }
    if (b->submitted - b->fetched >= SLOTS) {
        SDRM_LOG_ERROR("too many calls in flight: fetch results first");
        return -EBUSY;
    }
    int code = set_device(b);
    if (code != 0) return code;
    const int slot = (int) (b->submitted % SLOTS);
    code = ensure_staging(b, slot);
    if (code == 0 && b->d_in16[slot] == NULL) {
        code = sdrm_dev_zalloc(&b->d_in16[slot], (size_t) b->n_ch * b->in_stride_dev * 4);
    }
    if (code != 0) return code;
    if (input_len > 0) {
        if (b->submitted >= SLOTS) {
            /* the int16 staging buffer of call k - SLOTS was last read by its conversion (ev_copy is recorded behind it) */
            SDRM_CUDA_TRY(cudaStreamWaitEvent(b->s_copy, b->ev_copy[slot], 0));
        }
        size_t stride16 = b->in_stride_dev;
        if (packed_rows(b, in_stride, input_len)) {
            SDRM_CUDA_TRY(cudaMemcpyAsync(b->d_in16[slot], input, ((size_t) (b->n_ch - 1) * in_stride + input_len) * 4,
                                          cudaMemcpyHostToDevice, b->s_copy));
            stride16 = in_stride;
        } else {
            SDRM_CUDA_TRY(cudaMemcpy2DAsync(b->d_in16[slot], b->in_stride_dev * 4, input, in_stride * 4, input_len * 4, b->n_ch,
                                            cudaMemcpyHostToDevice, b->s_copy));
        }
        /* The conversion runs on its own stream: on the copy stream the next call's copy would wait behind it, and behind the
         * filter CTAs it has to wait for, with the copy engine idle (int16 ingest is bound by the copy: 10.1 ms per 0.5 GiB
         * against 8.9 ms of kernels). */
        SDRM_CUDA_TRY(cudaEventRecord(b->ev_h2d[slot], b->s_copy));
        SDRM_CUDA_TRY(cudaStreamWaitEvent(b->s_conv, b->ev_h2d[slot], 0));
        if (b->submitted >= SLOTS) {
            /* the cf32 buffer it writes was last read by the filters of call k - SLOTS */
            SDRM_CUDA_TRY(cudaStreamWaitEvent(b->s_conv, b->ev_fir[slot], 0));
        }
        code = sdrm_launch_code(sdrm_cu_i16_to_cf32(b->d_in16[slot], stride16, b->d_in[slot], b->in_stride_dev, scalar,
                                                    (int) input_len, (int) b->n_ch, b->s_conv),
                                "int16 ingest");
        if (code != 0) return code;
        b->launches++;
        SDRM_CUDA_TRY(cudaEventRecord(b->ev_copy[slot], b->s_conv));
    } else {
        SDRM_CUDA_TRY(cudaEventRecord(b->ev_copy[slot], b->s_copy));
    }
    return enqueue(b, b->d_in[slot], b->in_stride_dev, input_len, slot, 1);
}

static int copy_columns(sdrm_fsk_demod_batch *b, int slot, int8_t *output, float *soft, size_t out_stride, size_t first, size_t width) {
    if (width == 0) {
        return 0;
    }
    if (output != NULL) {
        SDRM_CUDA_TRY(cudaMemcpy2DAsync(output + first, out_stride, b->d_hard[slot] + first, b->out_stride, width, b->n_ch,
                                        cudaMemcpyDeviceToHost, b->s_out));
    }
    if (soft != NULL) {
        SDRM_CUDA_TRY(cudaMemcpy2DAsync(soft + first, out_stride * sizeof(float), b->d_soft[slot] + first, b->out_stride * sizeof(float),
                                        width * sizeof(float), b->n_ch, cudaMemcpyDeviceToHost, b->s_out));
    }
    return 0;
}

int sdrm_fsk_demod_batch_fetch(sdrm_fsk_demod_batch *b, int8_t *output, float *soft, size_t out_stride, uint32_t *output_len) {
    if (b == NULL || b->fetched >= b->submitted || (soft != NULL && !b->want_soft)) {
        return -1; /* nothing is enqueued and the call stays un-fetched */
    }
    int code = set_device(b);
    if (code != 0) return code;
    const int slot = (int) (b->fetched % SLOTS);
    const size_t width = out_stride < b->out_stride ? out_stride : b->out_stride;
    /* Only the columns the call can have filled travel: a handle created for 2016000-sample buffers (the reference's perf
     * program) and fed 4096 samples would otherwise move 2 MB for 400 symbols. */
    const size_t first_pass = b->slot_bound[slot] < width ? b->slot_bound[slot] : width;
    /* results leave on their own stream so that a younger call's tail does not delay them */
    SDRM_CUDA_TRY(cudaStreamWaitEvent(b->s_out, b->ev_tail[slot], 0));
    code = copy_columns(b, slot, output, soft, out_stride, 0, first_pass);
    if (code != 0) return code;
    SDRM_CUDA_TRY(cudaMemcpyAsync(b->h_counts, b->d_out_len[slot], (size_t) b->n_ch * sizeof(uint32_t), cudaMemcpyDeviceToHost, b->s_out));
    SDRM_CUDA_TRY(cudaEventRecord(b->ev_done[slot], b->s_out));
    SDRM_CUDA_TRY(cudaEventSynchronize(b->ev_done[slot]));
    uint32_t most = 0;
    for (uint32_t c = 0; c < b->n_ch; c++) {
        most = b->h_counts[c] > most ? b->h_counts[c] : most;
    }
    if ((size_t) most > first_pass && first_pass < width) {
        /* never seen; kept so that the result does not rest on the bound */
        const size_t upto = (size_t) most < width ? (size_t) most : width;
        code = copy_columns(b, slot, output, soft, out_stride, first_pass, upto - first_pass);
        if (code != 0) return code;
        SDRM_CUDA_TRY(cudaStreamSynchronize(b->s_out));
    }
    if (output_len != NULL) {
        memcpy(output_len, b->h_counts, (size_t) b->n_ch * sizeof(uint32_t));
    }
    b->last_fetch_columns = (size_t) most > first_pass && first_pass < width ? ((size_t) most < width ? (size_t) most : width) : first_pass;
    b->fetched++;
    return 0;
}

int sdrm_fsk_demod_batch_process(sdrm_fsk_demod_batch *b, const float complex *input, size_t in_stride, size_t input_len,
                                 int8_t *output, float *soft, size_t out_stride, uint32_t *output_len) {
    int code = sdrm_fsk_demod_batch_submit(b, input, in_stride, input_len);
    if (code != 0) {
        return code;
    }
    return sdrm_fsk_demod_batch_fetch(b, output, soft, out_stride, output_len);
}

int sdrm_fsk_demod_batch_device_outputs(sdrm_fsk_demod_batch *b, const int8_t **d_output, const uint32_t **d_output_len,
                                        size_t *out_stride) {
    if (b == NULL || b->submitted == 0) {
        return -1;
    }
    const int slot = (int) ((b->submitted - 1) % SLOTS);
    if (d_output != NULL) *d_output = b->d_hard[slot];
    if (d_output_len != NULL) *d_output_len = b->d_out_len[slot];
    if (out_stride != NULL) *out_stride = b->out_stride;
    return 0;
}

/* Drops the results of the oldest un-fetched call without copying them (device-resident pipelines). */
int sdrm_fsk_demod_batch_release(sdrm_fsk_demod_batch *b) {
    if (b == NULL || b->fetched >= b->submitted) {
        return -1;
    }
    b->fetched++;
    return 0;
}

int sdrm_fsk_demod_batch_sync(sdrm_fsk_demod_batch *b) {
    if (b == NULL) {
        return -1;
    }
    int code = set_device(b);
    if (code != 0) return code;
    SDRM_CUDA_TRY(cudaStreamSynchronize(b->s_copy));
    SDRM_CUDA_TRY(cudaStreamSynchronize(b->s_conv));
    SDRM_CUDA_TRY(cudaStreamSynchronize(b->s_fir));
    SDRM_CUDA_TRY(cudaStreamSynchronize(b->s_tail));
    SDRM_CUDA_TRY(cudaStreamSynchronize(b->s_out));
    return 0;
}

int sdrm_fsk_demod_batch_set_profiling(sdrm_fsk_demod_batch *b, int enabled) {
    if (b == NULL) {
        return -1;
    }
    int code = set_device(b);
    if (code != 0) return code;
    if (enabled && b->ev_time[0][0] == NULL) {
        for (int s = 0; s < SLOTS; s++) {
            for (int k = 0; k < 6; k++) {
                SDRM_CUDA_TRY(cudaEventCreate(&b->ev_time[s][k]));
            }
        }
    }
    b->profiling = enabled != 0;
    return 0;
}

int sdrm_fsk_demod_batch_stage_times(sdrm_fsk_demod_batch *b, float ms[4]) {
    if (b == NULL || ms == NULL || !b->profiling || b->submitted == 0) {
        return -1;
    }
    int code = sdrm_fsk_demod_batch_sync(b);
    if (code != 0) return code;
    cudaEvent_t *e = b->ev_time[b->last_slot];
    SDRM_CUDA_TRY(cudaEventElapsedTime(&ms[0], e[0], e[1]));
    SDRM_CUDA_TRY(cudaEventElapsedTime(&ms[1], e[1], e[2]));
    SDRM_CUDA_TRY(cudaEventElapsedTime(&ms[2], e[4], e[5])); /* fused dc blocker + clock recovery kernel */
    SDRM_CUDA_TRY(cudaEventElapsedTime(&ms[3], e[0], e[5])); /* whole call, first kernel start to tail end */
    return 0;
}

void *sdrm_fsk_demod_batch_stream(sdrm_fsk_demod_batch *b) { return b == NULL ? NULL : (void *) b->s_fir; }

void *sdrm_fsk_demod_batch_tail_stream(sdrm_fsk_demod_batch *b) { return b == NULL ? NULL : (void *) b->s_tail; }

void *sdrm_fsk_demod_batch_out_stream(sdrm_fsk_demod_batch *b) { return b == NULL ? NULL : (void *) b->s_out; }

/* Device-resident consumers: orders `stream` behind the tail of the most recently enqueued call, whose results
 * sdrm_fsk_demod_batch_device_outputs points at. */
int sdrm_fsk_demod_batch_wait_outputs(sdrm_fsk_demod_batch *b, void *stream) {
    if (b == NULL || b->submitted == 0) {
        return -1;
    }
    int code = set_device(b);
    if (code != 0) return code;
    SDRM_CUDA_TRY(cudaStreamWaitEvent((cudaStream_t) stream, b->ev_tail[(b->submitted - 1) % SLOTS], 0));
    return 0;
}

int sdrm_debug_set_measurement_aid(sdrm_fsk_demod_batch *b, uint32_t mask) {
    if (b == NULL) {
        return -1;
    }
    b->aid = mask;
    return 0;
}

uint64_t sdrm_fsk_demod_batch_launch_count(const sdrm_fsk_demod_batch *b) { return b == NULL ? 0 : b->launches; }

size_t sdrm_fsk_demod_batch_last_fetch_columns(const sdrm_fsk_demod_batch *b) { return b == NULL ? 0 : b->last_fetch_columns; }

int sdrm_fsk_demod_batch_error_flags(sdrm_fsk_demod_batch *b) {
    if (b == NULL) {
        return -1;
    }
    int flags = 0;
    if (set_device(b) != 0 || sdrm_fsk_demod_batch_sync(b) != 0 ||
        cudaMemcpy(&flags, b->d_error, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) {
        return -EIO;
    }
    return flags;
}

void sdrm_fsk_demod_batch_destroy(sdrm_fsk_demod_batch *b) {
    if (b == NULL) {
        return;
    }
    cudaSetDevice(b->device);
    cudaDeviceSynchronize();
    cudaFree(b->d_taps1);
    free(b->h_taps1);
    free(b->h_taps2);
    cudaFree(b->d_taps2);
    cudaFree(b->d_atan);
    cudaFree(b->d_mmse);
    cudaFree(b->d_q);
    cudaFree(b->d_acc);
    cudaFree(b->d_ring);
    cudaFree(b->d_delay);
    cudaFree(b->d_sums);
    cudaFree(b->d_clock);
    cudaFree(b->d_carry);
    cudaFree(b->d_error);
    if (b->h_counts != NULL) cudaFreeHost(b->h_counts);
    for (int i = 0; i < 2; i++) {
        cudaFree(b->d_hist1[i]);
        cudaFree(b->d_hist2[i]);
    }
    for (int s = 0; s < SLOTS; s++) {
        cudaFree(b->d_in[s]);
        cudaFree(b->d_in16[s]);
        cudaFree(b->d_hard[s]);
        cudaFree(b->d_soft[s]);
        cudaFree(b->d_out_len[s]);
        if (b->ev_fir[s] != NULL) cudaEventDestroy(b->ev_fir[s]);
        if (b->ev_tail[s] != NULL) cudaEventDestroy(b->ev_tail[s]);
        if (b->ev_copy[s] != NULL) cudaEventDestroy(b->ev_copy[s]);
        if (b->ev_done[s] != NULL) cudaEventDestroy(b->ev_done[s]);
        for (int k = 0; k < 6; k++) {
            if (b->ev_time[s][k] != NULL) cudaEventDestroy(b->ev_time[s][k]);
        }
    }
    if (b->s_copy != NULL) cudaStreamDestroy(b->s_copy);
    if (b->s_conv != NULL) cudaStreamDestroy(b->s_conv);
    for (int s = 0; s < SLOTS; s++) {
        if (b->ev_h2d[s] != NULL) cudaEventDestroy(b->ev_h2d[s]);
    }
    if (b->s_fir != NULL) cudaStreamDestroy(b->s_fir);
    if (b->s_tail != NULL) cudaStreamDestroy(b->s_tail);
    if (b->s_out != NULL) cudaStreamDestroy(b->s_out);
    free(b);
}

void *sdrm_pinned_alloc(size_t bytes) { return sdrm_pinned_alloc_near_device(bytes, -1); }

int sdrm_pinned_region_release(void *p); /* affinity.c */

void sdrm_pinned_free(void *p) {
    if (p != NULL && !sdrm_pinned_region_release(p)) {
        cudaFreeHost(p);
    }
}

const char *sdrm_version(void) {
    static char text[96];
    int runtime = 0;
    cudaRuntimeGetVersion(&runtime);
    snprintf(text, sizeof(text), "sdr-modem_b200 0.2.0; sm_100a; CUDA runtime %d", runtime);
    return text;
}

/* ---- SDR sample formats on device buffers (include/sdrm/sdrm_batch.h) --------------------------------------------------- */

int sdrm_samples_i16_to_cf32_device(const void *d_input, size_t in_stride, void *d_output, size_t out_stride, float scalar,
                                    size_t len, uint32_t rows, void *stream) {
    if (d_input == NULL || d_output == NULL || len > 0x7fffffffu) {
        return -1;
    }
    return sdrm_launch_code(sdrm_cu_i16_to_cf32(d_input, in_stride, d_output, out_stride, scalar, (int) len, (int) rows, stream),
                            "int16 -> cf32");
}

int sdrm_samples_cf32_to_i16_device(const void *d_input, size_t in_stride, void *d_output, size_t out_stride, float scalar,
                                    size_t len, uint32_t rows, void *stream) {
    if (d_input == NULL || d_output == NULL || len > 0x7fffffffu) {
        return -1;
    }
    return sdrm_launch_code(sdrm_cu_cf32_to_i16(d_input, in_stride, d_output, out_stride, scalar, (int) len, (int) rows, stream),
                            "cf32 -> int16");
}
