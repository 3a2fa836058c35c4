/*
 * Batched RX dispatcher: what the reference does with one sdr_worker thread fanning every SDR block out to N dsp_worker
 * threads (src/sdr_worker.c:25-55 -> src/dsp_worker.c:44-106), done as ONE hand-off and ONE set of kernel launches per
 * block for all sessions that hang off the same SDR stream and share demodulator parameters:
 *
 *     put(block) -> queue (pinned, same drop / blocking / poison-pill rules as src/queue.c)
 *                -> one host->device copy of the block
 *                -> doppler_process_rx for every session that asked for it (own TLE, ground station, start time)
 *                -> fsk_demod_process for every session
 *                -> per session: sink callback, or the client socket as dsp_worker does
 *
 * Up to two blocks are in flight, so that the serial tail of block k overlaps the filters of block k + 1; when the queue
 * runs empty everything in flight is drained at once, so a slow source sees no added latency.
 * Per-session results are bit-identical to N independent dsp_workers.
 */
#ifndef SDRM_RX_GROUP_H
#define SDRM_RX_GROUP_H

#include <complex.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

typedef struct sdrm_rx_group_t sdrm_rx_group;

/* called on the group's thread, once per block and session that produced symbols; `symbols` is valid during the call */
typedef void (*sdrm_rx_sink)(void *ctx, uint32_t session_id, const int8_t *symbols, size_t len);

typedef struct {
    uint32_t id;
    int client_socket;   /* used when sink == NULL: symbols are written to it (src/dsp_worker.c:93); -1 = discard */
    sdrm_rx_sink sink;
    void *sink_ctx;
    /* struct RxRequest doppler settings, same units as api.proto (degrees * 10E6, km * 10E3) */
    bool has_doppler;
    char doppler_tle[3][80];
    int32_t doppler_latitude;
    int32_t doppler_longitude;
    int32_t doppler_altitude;
    int64_t file_start_time_seconds; /* 0 = wall clock at the first block */
} sdrm_rx_session;

typedef struct {
    uint64_t rx_center_freq;
    uint64_t rx_sampling_freq;
    uint32_t demod_baud_rate;
    uint32_t demod_decimation;
    int64_t demod_fsk_deviation;
    uint32_t demod_fsk_transition_width;
    bool demod_fsk_use_dc_block;
    uint32_t buffer_size; /* server_config: samples per block at most */
    uint16_t queue_size;
    bool blocking_queue;  /* true: put waits for a free slot (file source); false: the oldest block is dropped */
    int device;           /* CUDA device ordinal, -1 = current */
} sdrm_rx_group_config;

int sdrm_rx_group_create(const sdrm_rx_group_config *config, const sdrm_rx_session *sessions, uint32_t n_sessions,
                         sdrm_rx_group **group);

/* the sdr thread's side: copies the block into the queue and returns (src/sdr_worker.c:25-29 for all sessions at once) */
void sdrm_rx_group_put(float complex *block, size_t len, sdrm_rx_group *group);

/* poison pill: the thread finishes the queued blocks, delivers what is in flight and stops */
void sdrm_rx_group_shutdown(sdrm_rx_group *group);

/* number of blocks fully processed and delivered so far */
uint64_t sdrm_rx_group_blocks_done(const sdrm_rx_group *group);

/* 1 once the group's thread has stopped on an error (a block that could not be enqueued, a failed result copy):
 * the reference's worker thread ends in the same situations; 0 while it is healthy */
int sdrm_rx_group_failed(const sdrm_rx_group *group);

/* sessions whose client socket could not be written to: they are no longer served (the reference's dsp_worker ends
 * on the first failed write, src/dsp_worker.c:93-103), the other sessions of the group carry on */
uint32_t sdrm_rx_group_sessions_failed(const sdrm_rx_group *group);

void sdrm_rx_group_destroy(sdrm_rx_group *group);

#endif
