/*
 * Batch extension of the reference's src/dsp API: N independent channel sessions that share one parameter set run as
 * one set of kernel launches per call. A reference handle (fsk_demod_create …) is a batch of one.
 *
 * What each entry point stands for in the reference:
 *   sdrm_fsk_demod_batch_create   N x fsk_demod_create            (src/dsp/fsk_demod.c:28-78, called from src/dsp_worker.c:140)
 *   sdrm_fsk_demod_batch_process  N x fsk_demod_process           (src/dsp/fsk_demod.c:80-110, called from src/dsp_worker.c:75)
 *   sdrm_fsk_demod_batch_destroy  N x fsk_demod_destroy           (src/dsp/fsk_demod.c:112-135)
 *
 * All channels of a batch receive the same number of samples per call (they hang off the same SDR block, reference
 * src/sdr_worker.c:46). Channel c of a [channels][stride] buffer starts at base + c * stride elements.
 * Return codes follow the reference: 0, -ENOMEM, -1 for invalid parameters; CUDA failures map to -EIO.
 * A handle is used by one thread at a time; different handles may be used concurrently.
 */
#ifndef SDRM_BATCH_H
#define SDRM_BATCH_H

#include <complex.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDRM_FLAG_FAST_FMA 1u  /* FIR dot products use fused multiply-add (not bit-identical to the reference) */
#define SDRM_FLAG_SOFT_OUT 2u  /* also keep the float soft symbols (clock_mm output) */

typedef struct sdrm_fsk_demod_batch_t sdrm_fsk_demod_batch;

typedef struct {
    uint32_t n_channels;
    uint64_t sampling_freq;
    uint32_t baud_rate;
    int64_t deviation;
    uint8_t decimation;
    uint32_t transition_width;
    bool use_dc_block;
    uint32_t max_input_buffer_length; /* samples per channel per call, as in the reference */
    uint32_t max_symbols_per_call;    /* output capacity per channel; 0 = max_input_buffer_length (reference) */
    uint32_t flags;                   /* SDRM_FLAG_*; unknown bits are rejected with -1 */
    int device;                       /* CUDA device ordinal, -1 = current */
} sdrm_fsk_demod_batch_config;

/*
 * Parameter ranges against the reference's fsk_demod_create. With sps = sampling_freq / baud_rate / decimation (samples per
 * symbol after decimation):
 *   sps <= 855   the fused serial tail (one kernel: dc blocker + clock recovery, one symbol step of every channel in shared
 *                memory);
 *   sps >  855   accepted like the reference does, served by two plain kernels (dc blocker, clock loop) on the lpf2 output
 *                ring: same results, a slower tail;
 *   with use_dc_block the dc blocker length ceil(32 * sps) must be >= 32, i.e. sps >= 1: create logs the reason and returns
 *                -1 below that (the reference accepts fractional sps < 1, which no FSK receiver can use).
 * Batches of more than 65535 channels are accepted (kernels that index rows with gridDim.y loop over the remainder).
 */
int sdrm_fsk_demod_batch_create(const sdrm_fsk_demod_batch_config *config, sdrm_fsk_demod_batch **batch);

/*
 * Host buffers in, host buffers out; returns when the outputs are in place.
 * input: cf32 [channels][in_stride], output: int8 [channels][out_stride], output_len: [channels].
 * soft (optional, needs SDRM_FLAG_SOFT_OUT): float [channels][out_stride].
 */
int sdrm_fsk_demod_batch_process(sdrm_fsk_demod_batch *batch, const float complex *input, size_t in_stride,
                                 size_t input_len, int8_t *output, float *soft, size_t out_stride, uint32_t *output_len);

/*
 * The same call split in two so that a caller can keep SDRM_MAX_IN_FLIGHT calls in flight: submit copies the input
 * host->device asynchronously (pinned memory recommended) and enqueues the chain, fetch (below) collects the oldest
 * call. process == submit + fetch.
 */
#define SDRM_MAX_IN_FLIGHT 3
int sdrm_fsk_demod_batch_submit(sdrm_fsk_demod_batch *batch, const float complex *input, size_t in_stride, size_t input_len);

/*
 * submit for int16 (I, Q) samples as the SDR delivers them: [channels][in_stride pairs]; converted on the device as the
 * reference's PlutoSDR plugin does on the host, (float) v / scalar with scalar = 2048 (src/sdr/plutosdr.c:129). Half the
 * host->device bytes of the cf32 call, same results as converting first.
 */
int sdrm_fsk_demod_batch_submit_i16(sdrm_fsk_demod_batch *batch, const int16_t *input, size_t in_stride, size_t input_len,
                                    float scalar);

/*
 * Device-resident variant: d_input is cf32 [channels][in_stride] in device memory (16-byte aligned, even stride).
 * Work is enqueued on the batch's streams; results stay on the device until sdrm_fsk_demod_batch_fetch.
 * Calls may be issued back to back: the serial tail of call k overlaps the filters of call k+1.
 */
int sdrm_fsk_demod_batch_process_device(sdrm_fsk_demod_batch *batch, const void *d_input, size_t in_stride,
                                        size_t input_len);

/* Waits for the oldest un-fetched call and copies its results out (any pointer may be NULL). */
int sdrm_fsk_demod_batch_fetch(sdrm_fsk_demod_batch *batch, int8_t *output, float *soft, size_t out_stride,
                               uint32_t *output_len);

/* Forgets the oldest un-fetched call without copying its results (device-resident pipelines). */
int sdrm_fsk_demod_batch_release(sdrm_fsk_demod_batch *batch);

/* Device pointers of the most recently enqueued call's results: int8 [channels][*out_stride], uint32 [channels].
 * They are complete once the call's tail has run: order the consumer with sdrm_fsk_demod_batch_wait_outputs. The buffers are
 * reused SDRM_MAX_IN_FLIGHT calls later; a device-resident consumer that releases calls without fetching them must have
 * its own stream wait (below) before the batch is given the call that reuses the slot. */
int sdrm_fsk_demod_batch_device_outputs(sdrm_fsk_demod_batch *batch, const int8_t **d_output, const uint32_t **d_output_len,
                                        size_t *out_stride);
/* Makes `stream` (a cudaStream_t) wait for the tail of the most recently enqueued call. */
int sdrm_fsk_demod_batch_wait_outputs(sdrm_fsk_demod_batch *batch, void *stream);

/* Blocks until everything enqueued so far has finished. */
int sdrm_fsk_demod_batch_sync(sdrm_fsk_demod_batch *batch);

/* cudaStream_t of the filter stage (for callers that order their own device work against the batch). */
void *sdrm_fsk_demod_batch_stream(sdrm_fsk_demod_batch *batch);
/* cudaStream_t of the serial tail (dc blocker, clock recovery). */
void *sdrm_fsk_demod_batch_tail_stream(sdrm_fsk_demod_batch *batch);
/* cudaStream_t the result copies of sdrm_fsk_demod_batch_fetch are issued on. */
void *sdrm_fsk_demod_batch_out_stream(sdrm_fsk_demod_batch *batch);

/* Number of kernels launched by this batch since creation (bench.py reports it as gpu_launches). */
uint64_t sdrm_fsk_demod_batch_launch_count(const sdrm_fsk_demod_batch *batch);
/* Columns (symbols) per channel the last fetch copied to the host: the most a call of that many samples can produce, not the
 * whole capacity of the result rows. For byte accounting. */
size_t sdrm_fsk_demod_batch_last_fetch_columns(const sdrm_fsk_demod_batch *batch);

/*
 * Optional per-stage timing with CUDA events on the launching streams. stage_times synchronises and returns, for the
 * most recent call, milliseconds of { lpf1+quad-demod kernel, lpf1 history + lpf2 kernel, fused dc blocker + clock
 * recovery kernel, whole call from the first kernel's start to the tail's end }.
 */
int sdrm_fsk_demod_batch_set_profiling(sdrm_fsk_demod_batch *batch, int enabled);
int sdrm_fsk_demod_batch_stage_times(sdrm_fsk_demod_batch *batch, float ms[4]);

/* Sticky error bits raised on the device (bit 0: clock history overflow, bit 1: symbol capacity reached). */
int sdrm_fsk_demod_batch_error_flags(sdrm_fsk_demod_batch *batch);

void sdrm_fsk_demod_batch_destroy(sdrm_fsk_demod_batch *batch);

/*
 * N x sig_source (reference src/dsp/sig_source.c): per-channel NCO with carried float phase, optional mix with the input.
 * freq_hz[c] is the integer frequency the reference passes to sig_source_multiply for channel c on this call
 * (doppler_process hands it (int64_t) current_freq_difference per call segment, src/dsp/doppler.c:180).
 */
typedef struct sdrm_nco_batch_t sdrm_nco_batch;

int sdrm_nco_batch_create(uint32_t n_channels, float amplitude, uint64_t sampling_freq, uint32_t max_output_buffer_length,
                          int device, sdrm_nco_batch **batch);
/* host buffers; input == NULL generates the tone (sig_source_process) */
int sdrm_nco_batch_process(sdrm_nco_batch *batch, const int64_t *freq_hz, const float complex *input, size_t in_stride,
                           size_t len, float complex *output, size_t out_stride);
/* device buffers (cf32 rows), asynchronous on the batch's stream */
int sdrm_nco_batch_process_device(sdrm_nco_batch *batch, const int64_t *freq_hz, const void *d_input, size_t in_stride,
                                  size_t len, void *d_output, size_t out_stride);
int sdrm_nco_batch_sync(sdrm_nco_batch *batch);
void *sdrm_nco_batch_stream(sdrm_nco_batch *batch);
void sdrm_nco_batch_destroy(sdrm_nco_batch *batch);

/*
 * N x doppler (reference src/dsp/doppler.c:44-190, called from src/dsp_worker.c:68,130 and src/tcp_server.c:202,549):
 * every channel has its own satellite (TLE), start time and ground station; all share the sample rate, centre frequency
 * and the number of samples per call. The orbit model runs on the host (twice per channel per second of signal), the
 * per-sample rotation on the GPU. direction: +1 = rx (doppler_process_rx), -1 = tx (doppler_process_tx).
 */
typedef struct sdrm_doppler_batch_t sdrm_doppler_batch;

typedef struct {
    double latitude;  /* degrees */
    double longitude; /* degrees */
    double altitude;  /* km */
    int64_t constant_offset;
    int64_t start_time_seconds; /* unix time; 0 = wall clock at the first call, as the reference */
    char tle[3][80];
} sdrm_doppler_channel;

int sdrm_doppler_batch_create(uint32_t n_channels, const sdrm_doppler_channel *channels, uint64_t sampling_freq,
                              uint64_t center_freq, uint32_t max_output_buffer_length, int device, sdrm_doppler_batch **batch);
/* host buffers, cf32 [channels][stride] */
int sdrm_doppler_batch_process(sdrm_doppler_batch *batch, int direction, const float complex *input, size_t in_stride,
                               size_t len, float complex *output, size_t out_stride);
/* device buffers, asynchronous on the batch's stream */
int sdrm_doppler_batch_process_device(sdrm_doppler_batch *batch, int direction, const void *d_input, size_t in_stride,
                                      size_t len, void *d_output, size_t out_stride);
int sdrm_doppler_batch_sync(sdrm_doppler_batch *batch);
void *sdrm_doppler_batch_stream(sdrm_doppler_batch *batch);
void sdrm_doppler_batch_destroy(sdrm_doppler_batch *batch);

/*
 * N x gfsk_mod (reference src/dsp/gfsk_mod.c:43-148, called from src/tcp_server.c:196,529): bytes in, cf32 out,
 * 8 * (int) samples_per_symbol output samples per input byte. All channels receive the same number of bytes per call.
 */
typedef struct sdrm_gfsk_mod_batch_t sdrm_gfsk_mod_batch;

int sdrm_gfsk_mod_batch_create(uint32_t n_channels, float samples_per_symbol, float sensitivity, float bt,
                               uint32_t max_input_buffer_length, int device, sdrm_gfsk_mod_batch **batch);
/* host buffers: input uint8 [channels][in_stride], output cf32 [channels][out_stride] */
int sdrm_gfsk_mod_batch_process(sdrm_gfsk_mod_batch *batch, const uint8_t *input, size_t in_stride, size_t input_len,
                                float complex *output, size_t out_stride, size_t *output_len);
/* as process, followed on the device by the PlutoSDR egress conversion (src/sdr/plutosdr.c:83): int16 (I, Q) pairs
 * [channels][out_stride pairs] = saturate(rint(v * scalar)), scalar = 32768 */
int sdrm_gfsk_mod_batch_process_i16(sdrm_gfsk_mod_batch *batch, const uint8_t *input, size_t in_stride, size_t input_len,
                                    int16_t *output, size_t out_stride, float scalar, size_t *output_len);
/*
 * Device buffers, asynchronous; d_output rows must be 8-byte aligned. A call is three stages on three streams (shaping,
 * phase walk, trigonometry) so that consecutive calls overlap: d_input is consumed on sdrm_gfsk_mod_batch_input_stream
 * (order its producer before that stream, or have it ready when the call is made), d_output is complete on
 * sdrm_gfsk_mod_batch_stream. Two calls may be in flight; give them different output buffers.
 */
int sdrm_gfsk_mod_batch_process_device(sdrm_gfsk_mod_batch *batch, const void *d_input, size_t in_stride, size_t input_len,
                                       void *d_output, size_t out_stride);
int sdrm_gfsk_mod_batch_sync(sdrm_gfsk_mod_batch *batch);
void *sdrm_gfsk_mod_batch_stream(sdrm_gfsk_mod_batch *batch);
void *sdrm_gfsk_mod_batch_input_stream(sdrm_gfsk_mod_batch *batch);
uint64_t sdrm_gfsk_mod_batch_launch_count(const sdrm_gfsk_mod_batch *batch);
void sdrm_gfsk_mod_batch_destroy(sdrm_gfsk_mod_batch *batch);

/*
 * N x lpf (reference src/dsp/lpf.c:12-51): one low-pass filter design, N streams with their own history, complex
 * (num_bytes 8) or real (num_bytes 4) samples, decimation as in lpf_create. All streams receive the same number of samples
 * per call, so they all produce the same number of outputs (*output_len).
 * Strides are in samples of the stream's type.
 */
typedef struct sdrm_lpf_batch_t sdrm_lpf_batch;

int sdrm_lpf_batch_create(uint32_t n_channels, uint8_t decimation, uint64_t sampling_freq, uint64_t cutoff_freq,
                          uint32_t transition_width, uint32_t max_input_buffer_length, size_t num_bytes, int device,
                          sdrm_lpf_batch **batch);
/* host buffers [channels][stride] */
int sdrm_lpf_batch_process(sdrm_lpf_batch *batch, const void *input, size_t in_stride, size_t input_len, void *output,
                           size_t out_stride, size_t *output_len);
/* device buffers, asynchronous on the batch's stream; complex rows 16-byte aligned with an even stride */
int sdrm_lpf_batch_process_device(sdrm_lpf_batch *batch, const void *d_input, size_t in_stride, size_t input_len,
                                  void *d_output, size_t out_stride, size_t *output_len);
int sdrm_lpf_batch_sync(sdrm_lpf_batch *batch);
void *sdrm_lpf_batch_stream(sdrm_lpf_batch *batch);
uint64_t sdrm_lpf_batch_launch_count(const sdrm_lpf_batch *batch);
void sdrm_lpf_batch_destroy(sdrm_lpf_batch *batch);

/*
 * SDR sample formats on device buffers (reference src/sdr/plutosdr.c:83,129; VOLK generic 16i <-> 32f kernels), rows of
 * `len` complex samples, strides in complex samples, asynchronous on `stream` (a cudaStream_t, NULL = default stream).
 */
int sdrm_samples_i16_to_cf32_device(const void *d_input, size_t in_stride, void *d_output, size_t out_stride, float scalar,
                                    size_t len, uint32_t rows, void *stream);
int sdrm_samples_cf32_to_i16_device(const void *d_input, size_t in_stride, void *d_output, size_t out_stride, float scalar,
                                    size_t len, uint32_t rows, void *stream);

/* Pinned host memory for the host-buffer entry points (pageable memory works too, but copies serialise). */
void *sdrm_pinned_alloc(size_t bytes);
/* The same on the NUMA node of CUDA device `device`: anonymous 2 MB-aligned mapping bound to the node the kernel reports for
 * the device's PCI function (/sys/bus/pci/devices/<id>/numa_node), transparent huge pages requested, registered with the
 * driver. Where no node is reported (-1: single-node hosts, most virtual machines) it is sdrm_pinned_alloc. */
void *sdrm_pinned_alloc_near_device(size_t bytes, int device);
void sdrm_pinned_free(void *p);
/* NUMA node of a CUDA device's PCI function, -1 when the platform does not say. */
int sdrm_device_numa_node(int device);

/*
 * Moves the calling thread onto the CPUs that are local to CUDA device `device` (sysfs local_cpulist of its PCI function,
 * intersected with the CPUs the process may use), so that pinned memory it allocates next lands on the device's socket.
 * Returns the number of CPUs in the new mask, 0 when nothing was changed (no topology information, or no local CPU
 * allowed), a negative errno on failure. Call it before sdrm_pinned_alloc / *_batch_create in a thread that feeds one GPU.
 */
int sdrm_bind_thread_near_device(int device);
/* Parser behind it, exported for tests: number of CPUs in a sysfs cpu list such as "0-23,48-71", -1 if malformed. */
int sdrm_cpulist_parse_count(const char *text);

/*
 * Measures the FP32 pipe of `device` (-1 = current) with the two instruction mixes of the FIR kernels, in algorithmic TFLOP/s
 * (multiply + add = 2 flop per float lane and tap): *fma_tflops one FFMA2 per tap (the FMA peak roofline fractions are quoted
 * against), *exact_pair_tflops the separately rounded multiply and add of exact mode (two FFMA2 per tap). Synchronous, ~15 ms.
 */
int sdrm_measure_fp32_peak(int device, double *fma_tflops, double *exact_pair_tflops);

/*
 * Times `repeats` plain cudaMemcpyAsync copies of `bytes` from `host` (pinned) to `d_scratch` on a stream of its own and
 * returns the elapsed device time in *seconds: the platform's host -> device ceiling for that buffer at that moment (run it
 * on all GPUs at once to see what the host gives N of them together).
 */
int sdrm_probe_h2d(int device, const void *host, void *d_scratch, size_t bytes, int repeats, double *seconds);

/*
 * Creates the CUDA context of `device` (-1 = current) and loads the library's kernels now instead of inside the first *_create
 * (about two seconds on a B200 box). A server calls it while it starts up; SDRM_WARM_START=<device> in the environment makes the
 * library do it when it is loaded. 0, or -EIO without a usable device.
 */
int sdrm_warm_start(int device);

/* Library/runtime identification: "sdr-modem_b200 <version>; sm_100a; CUDA runtime <n>". */
const char *sdrm_version(void);

#ifdef __cplusplus
}
#endif

#endif
