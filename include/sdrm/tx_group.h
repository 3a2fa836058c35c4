/*
 * Batched TX dispatcher: the transmit side of the reference's tcp_worker (chain built in handle_tx_client,
 * src/tcp_server.c:491-570, driven by handle_tx_data, src/tcp_server.c:175-241) for N sessions that share the modulator
 * parameters, as ONE host->device copy, ONE set of kernel launches and ONE device->host copy per batch of bytes:
 *
 *     process(bytes) -> loop over batches of at most buffer_size bytes (src/tcp_server.c:186-192)
 *                    -> gfsk_mod_process                                   (:196)
 *                    -> doppler_process_tx for the sessions that asked for it, tx_offset inside (:202, :549)
 *                       or sig_source_multiply(tx_offset) for the others with an offset (:209, :558)
 *                    -> [<base_path>/tx.mod2sdr.<id>.cf32]                 (:214, :568)
 *                    -> the session's sink = tx_device->sdr_process_tx     (:223, src/sdr/sdr_device.h:22),
 *                       optionally as int16 (I, Q) pairs, the PlutoSDR plugin's format (src/sdr/plutosdr.c:83)
 *
 * The samples stay on the device between the stages; two batches are in flight, so the sinks and dump files of batch k
 * run while batch k + 1 is being modulated. Every session's samples are bit-identical to the reference's chain on the
 * same bytes and batch sizes.
 *
 * One deliberate difference: the reference sizes its doppler / sig_source buffers as samples_per_symbol * buffer_size
 * (src/tcp_server.c:537-538) although a batch of buffer_size bytes modulates to 8 times as many samples, so any batch
 * above buffer_size / 8 bytes is rejected by doppler_process_tx and silently transmits nothing. Buffers here hold a full
 * batch.
 */
#ifndef SDRM_TX_GROUP_H
#define SDRM_TX_GROUP_H

#include <complex.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

typedef struct sdrm_tx_group_t sdrm_tx_group;

/*
 * The SDR side. `samples` holds `len` complex samples: float complex, or int16_t (I, Q) pairs when the group was created
 * with output_int16; valid during the call. Called on the thread that runs sdrm_tx_group_process. A non-zero return is
 * the reference's "unable to transmit request fully": the session receives nothing more from this process call.
 */
typedef int (*sdrm_tx_sink)(void *ctx, uint32_t session_id, const void *samples, size_t len);

typedef struct {
    uint32_t id;
    sdrm_tx_sink sink; /* NULL: no device (samples are only dumped) */
    void *sink_ctx;
    /* struct TxRequest, same units as api.proto */
    bool tx_dump_file;
    int64_t tx_offset;
    bool has_doppler;
    char doppler_tle[3][80];
    int32_t doppler_latitude;  /* degrees * 10E6 */
    int32_t doppler_longitude; /* degrees * 10E6 */
    int32_t doppler_altitude;  /* km * 10E3 */
    int64_t file_start_time_seconds; /* 0 = wall clock at the first batch */
} sdrm_tx_session;

typedef struct {
    uint64_t tx_center_freq;
    uint64_t tx_sampling_freq;
    uint32_t mod_baud_rate;
    int64_t mod_fsk_deviation;
    uint32_t buffer_size;  /* server_config: bytes per batch at most */
    const char *base_path; /* directory of the dump files; may be NULL when no session dumps */
    bool output_int16;     /* sinks receive int16 pairs = saturate(rint(v * int16_scalar)) */
    float int16_scalar;    /* 0 = 32768 (src/sdr/plutosdr.c:83) */
    int device;            /* CUDA device ordinal, -1 = current */
} sdrm_tx_group_config;

/* 0, -1 for invalid parameters (as gfsk_mod_create / doppler_create), -ENOMEM, -EIO */
int sdrm_tx_group_create(const sdrm_tx_group_config *config, const sdrm_tx_session *sessions, uint32_t n_sessions,
                         sdrm_tx_group **group);

/*
 * handle_tx_data for every session at once: data holds `len` bytes per session, session i at data + i * stride, in the
 * order the sessions were given to create. session_status (may be NULL) receives per session 0 or
 * RESPONSE_DETAILS_INTERNAL_ERROR (3) when its sink failed. Returns 0, or a negative code when the GPU chain itself failed.
 */
int sdrm_tx_group_process(sdrm_tx_group *group, const uint8_t *data, size_t stride, size_t len, int *session_status);

/* complex samples produced per input byte: 8 * (int) (tx_sampling_freq / mod_baud_rate) */
size_t sdrm_tx_group_samples_per_byte(const sdrm_tx_group *group);

void sdrm_tx_group_destroy(sdrm_tx_group *group);

#endif
