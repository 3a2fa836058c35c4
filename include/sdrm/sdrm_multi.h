/*
 * Multi-GPU entry points: N channel sessions partitioned over a list of CUDA devices by channel range, device g of G owning
 * channels [g * C / G, (g + 1) * C / G). Sessions are independent, so there is no collective and no
 * peer traffic: every device gets its own host thread, stream set and pinned ingest ring, and the calls below only hand each
 * thread its slice.
 *
 * What this stands for in the reference: one sdr_worker thread handing every SDR block to all of its dsp_worker threads
 * (src/sdr_worker.c:31-55), each of which runs the chain for one client on one CPU core (src/dsp_worker.c:44-106).
 * Results are bit-identical to a single-device batch of the same channels (tests/test_gpu_multi_device.py).
 */
#ifndef SDRM_MULTI_H
#define SDRM_MULTI_H

#include "rx_group.h"
#include "sdrm_batch.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- N x fsk_demod over several devices ------------------------------------------------------------------------------------ */

typedef struct sdrm_fsk_demod_multi_t sdrm_fsk_demod_multi;

/*
 * config->n_channels is the TOTAL channel count, config->device is ignored; `devices` lists the CUDA ordinals to use (an
 * ordinal may appear more than once: two shards on one GPU). Fails like sdrm_fsk_demod_batch_create; additionally -1 when
 * n_devices is 0 or larger than the channel count.
 */
int sdrm_fsk_demod_multi_create(const sdrm_fsk_demod_batch_config *config, const int *devices, uint32_t n_devices,
                                sdrm_fsk_demod_multi **multi);

/* As sdrm_fsk_demod_batch_submit / _fetch / _process for all channels: host buffers [channels][stride]. Every device's
 * thread copies and enqueues its own channel range; pageable input is first staged through that thread's pinned ring (the
 * threads run in parallel, each on the CPUs next to its GPU), pinned input is copied from where it lies. Up to
 * SDRM_MAX_IN_FLIGHT calls may be in flight; the input of a submit must stay valid until the matching fetch has returned
 * (or sdrm_fsk_demod_multi_sync). */
int sdrm_fsk_demod_multi_submit(sdrm_fsk_demod_multi *multi, const float complex *input, size_t in_stride, size_t input_len);
int sdrm_fsk_demod_multi_submit_i16(sdrm_fsk_demod_multi *multi, const int16_t *input, size_t in_stride, size_t input_len,
                                    float scalar);
int sdrm_fsk_demod_multi_fetch(sdrm_fsk_demod_multi *multi, int8_t *output, float *soft, size_t out_stride, uint32_t *output_len);
int sdrm_fsk_demod_multi_process(sdrm_fsk_demod_multi *multi, const float complex *input, size_t in_stride, size_t input_len,
                                 int8_t *output, float *soft, size_t out_stride, uint32_t *output_len);
int sdrm_fsk_demod_multi_sync(sdrm_fsk_demod_multi *multi);

uint32_t sdrm_fsk_demod_multi_device_count(const sdrm_fsk_demod_multi *multi);
/* channel range [*first, *first + *count) and CUDA ordinal of shard g */
int sdrm_fsk_demod_multi_shard(const sdrm_fsk_demod_multi *multi, uint32_t g, uint32_t *first, uint32_t *count, int *device);
/* the per-device batch of shard g (for its streams, launch count, profiling); owned by the multi handle */
sdrm_fsk_demod_batch *sdrm_fsk_demod_multi_batch(sdrm_fsk_demod_multi *multi, uint32_t g);
uint64_t sdrm_fsk_demod_multi_launch_count(const sdrm_fsk_demod_multi *multi);
/* OR of the shards' sticky error bits */
int sdrm_fsk_demod_multi_error_flags(sdrm_fsk_demod_multi *multi);
void sdrm_fsk_demod_multi_destroy(sdrm_fsk_demod_multi *multi);

/* ---- the sdr_worker fan-out over several devices ------------------------------------------------------------------------- */

typedef struct sdrm_rx_multi_t sdrm_rx_multi;

/* Sessions [g * n / G, (g + 1) * n / G) run as one sdrm_rx_group on devices[g] (config->device is ignored): its own
 * queue of pinned blocks, its own thread, one host->device copy of every block per device. */
int sdrm_rx_multi_create(const sdrm_rx_group_config *config, const sdrm_rx_session *sessions, uint32_t n_sessions,
                         const int *devices, uint32_t n_devices, sdrm_rx_multi **multi);
/* the sdr thread's side: the block goes into every device's queue (src/sdr_worker.c:46-53) */
void sdrm_rx_multi_put(float complex *block, size_t len, sdrm_rx_multi *multi);
void sdrm_rx_multi_shutdown(sdrm_rx_multi *multi);
/* blocks delivered by the slowest device */
uint64_t sdrm_rx_multi_blocks_done(const sdrm_rx_multi *multi);
int sdrm_rx_multi_failed(const sdrm_rx_multi *multi);
void sdrm_rx_multi_destroy(sdrm_rx_multi *multi);

#ifdef __cplusplus
}
#endif

#endif
