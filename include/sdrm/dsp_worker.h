/*
 * The dsp_worker hand-off of the reference (src/dsp_worker.h:14-22, src/dsp_worker.c:44-227) with the GPU chain behind
 * it: one thread per RX session that takes blocks from its queue and runs
 *     [rx dump file] -> doppler_process_rx (optional) -> fsk_demod_process -> [demod file] -> [client socket].
 *
 * The reference's dsp_worker_create takes its parameters from protobuf (struct RxRequest) and libconfig
 * (struct server_config) objects, which belong to the control plane and are out of scope here; sdrm_dsp_worker_config
 * carries exactly the fields dsp_worker_create reads from them (src/dsp_worker.c:118-180), in the same units.
 * put / shutdown / destroy keep the reference's names and meaning.
 */
#ifndef SDRM_DSP_WORKER_H
#define SDRM_DSP_WORKER_H

#include <complex.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

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
} sdrm_dsp_worker_config;

int sdrm_dsp_worker_create(uint32_t id, int client_socket, const sdrm_dsp_worker_config *config, dsp_worker **result);

void dsp_worker_put(float complex *output, size_t output_len, dsp_worker *worker);

void dsp_worker_shutdown(void *arg, void *data);

bool dsp_worker_find_by_id(void *id, void *data);

void dsp_worker_destroy(void *data);

#endif
