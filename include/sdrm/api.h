/*
 * Framing of the reference's TCP API: the 6-byte message header (src/api.h:6-27), the helpers that read a header and a
 * message body from a socket and write a Response (src/api_utils.h:7-17), and the blocking socket helpers under them
 * (src/tcp_utils.h:7-11). Same names, signatures, return codes and limits (32 KiB per message body, src/api_utils.c:8) as
 * the reference, so that its callers (src/tcp_server.c:400-832, test/sdr_modem_client.c) link against this library
 * unchanged. The message bodies are the proto2 messages of api_messages.h.
 *
 *   byte 0      protocol_version (PROTOCOL_VERSION = 0)
 *   byte 1      type (TYPE_*)
 *   bytes 2-5   message_length, big endian: bytes of the protobuf body that follows
 */
#ifndef SDRM_API_H
#define SDRM_API_H

#include <stddef.h>
#include <stdint.h>

#include "api_messages.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PROTOCOL_VERSION 0

/* client to server */
#define TYPE_RX_REQUEST 0
#define TYPE_SHUTDOWN 1
#define TYPE_PING 3
#define TYPE_TX_DATA 4
#define TYPE_TX_REQUEST 5
/* server to client */
#define TYPE_RESPONSE 2

#define RESPONSE_NO_DETAILS 0
#define RESPONSE_DETAILS_INVALID_REQUEST 1
#define RESPONSE_DETAILS_INTERNAL_ERROR 3
#define RESPONSE_DETAILS_TX_IS_BEING_USED 4
#define RESPONSE_DETAILS_RX_IS_BEING_USED 5

struct message_header {
    uint8_t protocol_version;
    uint8_t type;
    uint32_t message_length;
} __attribute__((packed));

/* reads the 6 header bytes; message_length is returned in host order. 0, or the tcp_utils_read_data code */
int api_utils_read_header(int socket, struct message_header *header);

/* read header->message_length bytes and unpack them; -1 when the body is longer than 32 KiB, cannot be read or does not
 * parse, -ENOMEM; *request is released with the message's __free_unpacked */
int api_utils_read_rx_request(int socket, const struct message_header *header, struct RxRequest **request);
int api_utils_read_tx_request(int socket, const struct message_header *header, struct TxRequest **request);
int api_utils_read_tx_data(int socket, const struct message_header *header, struct TxData **request);

/* header (TYPE_RESPONSE) + Response{status, details} in one write */
int api_utils_write_response(int socket, ResponseStatus status, uint32_t details);

/* the three TLE lines of a DopplerSettings into the char[3][80] the doppler block takes (strncpy semantics) */
void api_utils_convert_tle(char **tle, char (*output)[80]);

/* blocking write of the whole buffer: 0, -1 on a failed write */
int tcp_utils_write_data(uint8_t *buffer, size_t total_len_bytes, int client_socket);
/* blocking read of exactly len_bytes: 0; -1 when the peer closed or on error; -EWOULDBLOCK / -EAGAIN on a receive timeout */
int tcp_utils_read_data(void *result, size_t len_bytes, int client_socket);
int tcp_utils_read_data_partially(void *result, size_t len_bytes, size_t *actually_read, int client_socket);

#ifdef __cplusplus
}
#endif

#endif
