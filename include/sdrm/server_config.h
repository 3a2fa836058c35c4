/*
 * Layout of the reference's struct server_config (src/server_config.h:17-44): dsp_worker_create and the reference's TCP
 * server pass it by pointer, so the field order and types are ABI. This library reads four fields of it (buffer_size,
 * queue_size, rx_sdr_type, base_path; src/dsp_worker.c:131-180); loading it from a libconfig file stays with the host
 * (server_config_create is control plane and not provided here).
 */
#ifndef SDRM_SERVER_CONFIG_H
#define SDRM_SERVER_CONFIG_H

#include <stdbool.h>
#include <stdint.h>

#define RX_SDR_TYPE_SDR_SERVER 0
#define RX_SDR_TYPE_PLUTOSDR 1
#define RX_SDR_TYPE_FILE 2

#define TX_SDR_TYPE_NONE 0
#define TX_SDR_TYPE_PLUTOSDR 1
#define TX_SDR_TYPE_FILE 2

struct server_config {
    /* socket settings */
    char *bind_address;
    uint16_t port;
    int read_timeout_seconds;

    uint32_t buffer_size; /* samples per block: max_input_buffer_length of every block of the chain */
    uint16_t queue_size;

    uint8_t rx_sdr_type; /* RX_SDR_TYPE_FILE selects the blocking queue (src/dsp_worker.c:172) */
    char *rx_sdr_server_address;
    int rx_sdr_server_port;

    /* output settings */
    char *base_path; /* directory of the dump files rx.sdr2demod.<id>.cf32, rx.demod2client.<id>.s8, tx.mod2sdr.<id>.cf32 */

    char *rx_file_base_path;
    char *tx_file_base_path;

    uint8_t tx_sdr_type;
    double tx_plutosdr_gain;
    double rx_plutosdr_gain;
    unsigned int tx_plutosdr_timeout_millis;
    void *iio; /* the reference's iio_lib *: libiio entry points loaded at run time, never touched here */
};

#endif
