/*
 * The dsp_worker hand-off of the reference (src/dsp_worker.h:14-22, src/dsp_worker.c:44-227) with the GPU chain behind
 * it: one thread per RX session that takes blocks from its queue and runs
 *     [rx dump file] -> doppler_process_rx (optional) -> fsk_demod_process -> [demod file] -> [client socket].
 *
 * dsp_worker_create has the reference's signature (src/dsp_worker.h:22): struct RxRequest is the protobuf message of
 * api_messages.h, struct server_config the layout of server_config.h. It is a thin layer over sdrm_dsp_worker_create, whose
 * plain config struct carries exactly the fields the reference reads from those two objects (src/dsp_worker.c:118-180), in
 * the same units, for hosts that have neither protobuf nor libconfig. put / shutdown / find / destroy keep the reference's
 * names and meaning.
 */
#ifndef SDRM_DSP_WORKER_H
#define SDRM_DSP_WORKER_H

#include <complex.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#include "api_messages.h"
#include "server_config.h"

typedef struct dsp_worker_t dsp_worker;

enum sdrm_demod_destination {
    SDRM_DEMOD_DESTINATION_FILE = 0,   /* api.proto DemodDestination */
    SDRM_DEMOD_DESTINATION_SOCKET = 1,
    SDRM_DEMOD_DESTINATION_BOTH = 2
};

typedef struct {
    /* struct RxRequest */
    uint64_t rx_center_freq;
    uint64_t rx_sampling_freq;
    bool rx_dump_file;
    bool demod_gmsk;            /* demod_type == MODEM_TYPE__GMSK; false: no demodulator (blocks are only dumped) */
    uint32_t demod_baud_rate;
    uint32_t demod_decimation;
    int64_t demod_fsk_deviation;
    uint32_t demod_fsk_transition_width;
    bool demod_fsk_use_dc_block;
    int demod_destination;      /* enum sdrm_demod_destination */
    bool has_doppler;
    char doppler_tle[3][80];
    int32_t doppler_latitude;   /* degrees * 10E6, as api.proto */
    int32_t doppler_longitude;  /* degrees * 10E6 */
    int32_t doppler_altitude;   /* km * 10E3 */
    int64_t file_start_time_seconds; /* file_settings->start_time_seconds, 0 = live */
    /* struct server_config */
    uint32_t buffer_size;
    uint16_t queue_size;
    bool blocking_queue;        /* rx_sdr_type == RX_SDR_TYPE_FILE */
    const char *base_path;
    bool doppler_scaled_unsigned; /* the three scaled doppler values are the uint32 wire fields of struct DopplerSettings and
                                     are divided as unsigned numbers, as the reference does (set by dsp_worker_create) */
} sdrm_dsp_worker_config;

int sdrm_dsp_worker_create(uint32_t id, int client_socket, const sdrm_dsp_worker_config *config, dsp_worker **result);

/* the reference's entry point (src/dsp_worker.c:108-197): 0, -ENOMEM, -1 / the failing block's code; a GMSK request without
 * fsk_settings is rejected with -1 (the reference dereferences the NULL pointer) */
int dsp_worker_create(uint32_t id, int client_socket, struct server_config *config, struct RxRequest *req, dsp_worker **result);

void dsp_worker_put(float complex *output, size_t output_len, dsp_worker *worker);

void dsp_worker_shutdown(void *arg, void *data);

bool dsp_worker_find_by_id(void *id, void *data);

void dsp_worker_destroy(void *data);

#endif
