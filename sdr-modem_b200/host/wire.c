/*
 * The reference's wire format without libprotobuf-c: a small table-driven proto2 codec for the eight messages of api.proto
 * behind the entry points protoc-c generates for them (include/sdrm/api_messages.h), and the header / socket helpers of
 * src/api_utils.c and src/tcp_utils.c (include/sdrm/api.h).
 *
 * Wire format (proto2): a message is a sequence of fields, each a varint key (field_number << 3 | wire_type) followed by
 *   wire type 0  varint           uint32, uint64, int64 (two's complement, ten bytes when negative), bool, enum
 *   wire type 2  length-delimited varint length + bytes: string, bytes, embedded message
 * wire types 1 (64-bit) and 5 (32-bit) do not occur in api.proto and are skipped like any unknown field.
 * Packing writes the fields in field-number order, required scalars always, optional messages when the pointer is set,
 * repeated strings one key per element. Unpacking accepts any order, keeps the last value of a repeated scalar key, merges
 * nothing (a second occurrence of an embedded message replaces the first), and fails when a required field is missing.
 */
#include <arpa/inet.h>
#include <errno.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>

#include "../../include/sdrm/api.h"

enum field_type { F_UINT32, F_UINT64, F_INT64, F_BOOL, F_ENUM, F_STRING, F_BYTES, F_MESSAGE };
enum field_label { L_REQUIRED, L_OPTIONAL, L_REPEATED };

struct message_layout;

struct field_layout {
    uint32_t number;
    enum field_type type;
    enum field_label label;
    size_t offset;       /* of the value (or of the pointer for strings / messages / repeated arrays) */
    size_t count_offset; /* repeated only: the size_t element count */
    const struct message_layout *message; /* F_MESSAGE only */
};

/* Starts with the magic number of a protobuf-c message descriptor: callers only ever take the address of a descriptor (the
 * __INIT macros), the rest of the layout is this codec's own. */
struct message_layout {
    uint32_t magic;
    const char *name;
    size_t size;
    const void *defaults; /* a message of this type in its initial state */
    size_t n_fields;
    const struct field_layout *fields;
};

#define MAGIC 0x28aaeef9u
#define FIELD(msg, num, type, label, member) { num, type, label, offsetof(msg, member), 0, NULL }
#define FIELD_MESSAGE(msg, num, member, layout) { num, F_MESSAGE, L_OPTIONAL, offsetof(msg, member), 0, layout }

/* The descriptors are declared `extern const ProtobufCMessageDescriptor x__descriptor` in the public header (an incomplete
 * type there, protobuf-c's own type in a host that has it). Callers only take their address, so they are DEFINED here with
 * this codec's layout and given the public symbol names with assembler labels. */
extern const struct message_layout layout_doppler_settings __asm__("doppler_settings__descriptor");
extern const struct message_layout layout_fsk_demodulation_settings __asm__("fsk_demodulation_settings__descriptor");
extern const struct message_layout layout_fsk_modulation_settings __asm__("fsk_modulation_settings__descriptor");
extern const struct message_layout layout_file_settings __asm__("file_settings__descriptor");
extern const struct message_layout layout_rx_request __asm__("rx_request__descriptor");
extern const struct message_layout layout_tx_request __asm__("tx_request__descriptor");
extern const struct message_layout layout_response __asm__("response__descriptor");
extern const struct message_layout layout_tx_data __asm__("tx_data__descriptor");

#define BASE(layout) { (const ProtobufCMessageDescriptor *) (const void *) &layout, 0, NULL }

static const DopplerSettings doppler_settings_defaults = { BASE(layout_doppler_settings), 0, NULL, 0, 0, 0 };
static const FskDemodulationSettings fsk_demodulation_settings_defaults = { BASE(layout_fsk_demodulation_settings), 0, 0, 0 };
static const FskModulationSettings fsk_modulation_settings_defaults = { BASE(layout_fsk_modulation_settings), 0 };
static const FileSettings file_settings_defaults = { BASE(layout_file_settings), NULL, 0 };
static const RxRequest rx_request_defaults = { BASE(layout_rx_request), 0, 0, 0, 0, MODEM_TYPE__GMSK, 0, 0, DEMOD_DESTINATION__FILE,
                                               NULL, NULL, NULL };
static const TxRequest tx_request_defaults = { BASE(layout_tx_request), 0, 0, 0, 0, MODEM_TYPE__GMSK, 0, NULL, NULL, NULL };
static const Response response_defaults = { BASE(layout_response), RESPONSE_STATUS__SUCCESS, 0 };
static const TxData tx_data_defaults = { BASE(layout_tx_data), { 0, NULL } };

static const struct field_layout doppler_settings_fields[] = {
    { 1, F_STRING, L_REPEATED, offsetof(DopplerSettings, tle), offsetof(DopplerSettings, n_tle), NULL },
    FIELD(DopplerSettings, 2, F_UINT32, L_REQUIRED, latitude),
    FIELD(DopplerSettings, 3, F_UINT32, L_REQUIRED, longitude),
    FIELD(DopplerSettings, 4, F_UINT32, L_REQUIRED, altitude),
};
static const struct field_layout fsk_demodulation_settings_fields[] = {
    FIELD(FskDemodulationSettings, 1, F_INT64, L_REQUIRED, demod_fsk_deviation),
    FIELD(FskDemodulationSettings, 2, F_UINT32, L_REQUIRED, demod_fsk_transition_width),
    FIELD(FskDemodulationSettings, 3, F_BOOL, L_REQUIRED, demod_fsk_use_dc_block),
};
static const struct field_layout fsk_modulation_settings_fields[] = {
    FIELD(FskModulationSettings, 1, F_INT64, L_REQUIRED, mod_fsk_deviation),
};
static const struct field_layout file_settings_fields[] = {
    FIELD(FileSettings, 1, F_STRING, L_REQUIRED, filename),
    FIELD(FileSettings, 2, F_UINT64, L_REQUIRED, start_time_seconds),
};
static const struct field_layout rx_request_fields[] = {
    FIELD(RxRequest, 1, F_UINT64, L_REQUIRED, rx_center_freq),
    FIELD(RxRequest, 2, F_UINT64, L_REQUIRED, rx_sampling_freq),
    FIELD(RxRequest, 3, F_BOOL, L_REQUIRED, rx_dump_file),
    FIELD(RxRequest, 4, F_INT64, L_REQUIRED, rx_offset),
    FIELD(RxRequest, 5, F_ENUM, L_REQUIRED, demod_type),
    FIELD(RxRequest, 6, F_UINT32, L_REQUIRED, demod_baud_rate),
    FIELD(RxRequest, 7, F_UINT32, L_REQUIRED, demod_decimation),
    FIELD(RxRequest, 8, F_ENUM, L_REQUIRED, demod_destination),
    FIELD_MESSAGE(RxRequest, 9, doppler, &layout_doppler_settings),
    FIELD_MESSAGE(RxRequest, 10, fsk_settings, &layout_fsk_demodulation_settings),
    FIELD_MESSAGE(RxRequest, 11, file_settings, &layout_file_settings),
};
static const struct field_layout tx_request_fields[] = {
    FIELD(TxRequest, 1, F_UINT64, L_REQUIRED, tx_center_freq),
    FIELD(TxRequest, 2, F_UINT64, L_REQUIRED, tx_sampling_freq),
    FIELD(TxRequest, 3, F_BOOL, L_REQUIRED, tx_dump_file),
    FIELD(TxRequest, 4, F_INT64, L_REQUIRED, tx_offset),
    FIELD(TxRequest, 5, F_ENUM, L_REQUIRED, mod_type),
    FIELD(TxRequest, 6, F_UINT32, L_REQUIRED, mod_baud_rate),
    FIELD_MESSAGE(TxRequest, 7, doppler, &layout_doppler_settings),
    FIELD_MESSAGE(TxRequest, 8, fsk_settings, &layout_fsk_modulation_settings),
    FIELD_MESSAGE(TxRequest, 9, file_settings, &layout_file_settings),
};
static const struct field_layout response_fields[] = {
    FIELD(Response, 1, F_ENUM, L_REQUIRED, status),
    FIELD(Response, 2, F_UINT32, L_REQUIRED, details),
};
static const struct field_layout tx_data_fields[] = {
    FIELD(TxData, 1, F_BYTES, L_REQUIRED, data),
};

#define LAYOUT(symbol, text, type, table) \
    const struct message_layout symbol = { MAGIC, text, sizeof(type), &table##_defaults, sizeof(table##_fields) / sizeof(table##_fields[0]), table##_fields }

LAYOUT(layout_doppler_settings, "doppler_settings", DopplerSettings, doppler_settings);
LAYOUT(layout_fsk_demodulation_settings, "fsk_demodulation_settings", FskDemodulationSettings, fsk_demodulation_settings);
LAYOUT(layout_fsk_modulation_settings, "fsk_modulation_settings", FskModulationSettings, fsk_modulation_settings);
LAYOUT(layout_file_settings, "file_settings", FileSettings, file_settings);
LAYOUT(layout_rx_request, "RxRequest", RxRequest, rx_request);
LAYOUT(layout_tx_request, "TxRequest", TxRequest, tx_request);
LAYOUT(layout_response, "Response", Response, response);
LAYOUT(layout_tx_data, "TxData", TxData, tx_data);

/* ---- enums ------------------------------------------------------------------------------------------------------------------ */

struct enum_layout {
    uint32_t magic;
    const char *name;
    size_t n_values;
    const ProtobufCEnumValue *values;
};

static const ProtobufCEnumValue modem_type_values[] = { { "GMSK", "MODEM_TYPE__GMSK", 1 } };
static const ProtobufCEnumValue demod_destination_values[] = { { "FILE", "DEMOD_DESTINATION__FILE", 0 },
                                                               { "SOCKET", "DEMOD_DESTINATION__SOCKET", 1 },
                                                               { "BOTH", "DEMOD_DESTINATION__BOTH", 2 } };
static const ProtobufCEnumValue response_status_values[] = { { "SUCCESS", "RESPONSE_STATUS__SUCCESS", 0 },
                                                             { "FAILURE", "RESPONSE_STATUS__FAILURE", 1 } };

#define ENUM_MAGIC 0x114315afu
const struct enum_layout layout_modem_type __asm__("modem_type__descriptor") = { ENUM_MAGIC, "modem_type", 1, modem_type_values };
const struct enum_layout layout_demod_destination __asm__("demod_destination__descriptor") = { ENUM_MAGIC, "demod_destination", 3,
                                                                                               demod_destination_values };
const struct enum_layout layout_response_status __asm__("response_status__descriptor") = { ENUM_MAGIC, "response_status", 2,
                                                                                           response_status_values };

const ProtobufCEnumValue *protobuf_c_enum_descriptor_get_value(const ProtobufCEnumDescriptor *desc, int value) {
    const struct enum_layout *layout = (const struct enum_layout *) (const void *) desc;
    if (layout == NULL || layout->magic != ENUM_MAGIC) {
        return NULL;
    }
    for (size_t i = 0; i < layout->n_values; i++) {
        if (layout->values[i].value == value) {
            return &layout->values[i];
        }
    }
    return NULL;
}

/* ---- allocation ------------------------------------------------------------------------------------------------------------ */

static void *wire_alloc(ProtobufCAllocator *allocator, size_t size) {
    if (size == 0) {
        size = 1;
    }
    return allocator != NULL ? allocator->alloc(allocator->allocator_data, size) : malloc(size);
}

static void wire_free(ProtobufCAllocator *allocator, void *p) {
    if (p == NULL) {
        return;
    }
    if (allocator != NULL) {
        allocator->free(allocator->allocator_data, p);
    } else {
        free(p);
    }
}

/* ---- packing ------------------------------------------------------------------------------------------------------------------ */

static size_t put_varint(uint64_t v, uint8_t *out) {
    size_t n = 0;
    while (v >= 0x80) {
        out[n++] = (uint8_t) (v | 0x80);
        v >>= 7;
    }
    out[n++] = (uint8_t) v;
    return n;
}

static uint64_t scalar_value(const struct field_layout *f, const void *member) {
    switch (f->type) {
    case F_UINT32:
        return *(const uint32_t *) member;
    case F_UINT64:
        return *(const uint64_t *) member;
    case F_INT64:
        return (uint64_t) *(const int64_t *) member;
    case F_BOOL:
        return *(const protobuf_c_boolean *) member ? 1 : 0;
    default: /* F_ENUM: negative values are sign-extended to 64 bits, as int32 is */
        return (uint64_t) (int64_t) *(const int *) member;
    }
}

static size_t message_size(const struct message_layout *layout, const void *message);

/* out == NULL: size only */
static size_t pack_message(const struct message_layout *layout, const void *message, uint8_t *out) {
    size_t n = 0;
    const char *base = message;
    uint8_t scratch[10];
    for (size_t i = 0; i < layout->n_fields; i++) {
        const struct field_layout *f = &layout->fields[i];
        const void *member = base + f->offset;
        const uint64_t key_varint = (uint64_t) f->number << 3;
        if (f->type == F_STRING || f->type == F_BYTES || f->type == F_MESSAGE) {
            size_t count = 1;
            char *const *strings = NULL;
            if (f->label == L_REPEATED) {
                count = *(const size_t *) (base + f->count_offset);
                strings = *(char *const *const *) member;
            }
            for (size_t k = 0; k < count; k++) {
                const uint8_t *payload = NULL;
                size_t len = 0;
                const void *sub = NULL;
                if (f->type == F_STRING) {
                    const char *text = f->label == L_REPEATED ? strings[k] : *(const char *const *) member;
                    if (text == NULL) {
                        if (f->label == L_OPTIONAL) {
                            continue;
                        }
                        text = "";
                    }
                    payload = (const uint8_t *) text;
                    len = strlen(text);
                } else if (f->type == F_BYTES) {
                    const ProtobufCBinaryData *data = member;
                    payload = data->data;
                    len = data->len;
                } else {
                    sub = *(const void *const *) member;
                    if (sub == NULL) {
                        continue; /* optional message not present */
                    }
                    len = message_size(f->message, sub);
                }
                n += put_varint(key_varint | 2, out != NULL ? out + n : scratch);
                n += put_varint(len, out != NULL ? out + n : scratch);
                if (out != NULL) {
                    if (sub != NULL) {
                        pack_message(f->message, sub, out + n);
                    } else if (len > 0) {
                        memcpy(out + n, payload, len);
                    }
                }
                n += len;
            }
        } else {
            n += put_varint(key_varint, out != NULL ? out + n : scratch);
            n += put_varint(scalar_value(f, member), out != NULL ? out + n : scratch);
        }
    }
    return n;
}

static size_t message_size(const struct message_layout *layout, const void *message) { return pack_message(layout, message, NULL); }

static size_t pack_to_buffer(const struct message_layout *layout, const void *message, ProtobufCBuffer *buffer) {
    const size_t len = message_size(layout, message);
    uint8_t *tmp = malloc(len == 0 ? 1 : len);
    if (tmp == NULL) {
        return 0;
    }
    pack_message(layout, message, tmp);
    buffer->append(buffer, len, tmp);
    free(tmp);
    return len;
}

/* ---- unpacking ---------------------------------------------------------------------------------------------------------------- */

static void free_message(const struct message_layout *layout, void *message, ProtobufCAllocator *allocator);

/* reads a varint at data[*pos]; 0 on success */
static int get_varint(const uint8_t *data, size_t len, size_t *pos, uint64_t *value) {
    uint64_t v = 0;
    for (unsigned shift = 0; shift < 70; shift += 7) {
        if (*pos >= len) {
            return -1;
        }
        const uint8_t byte = data[(*pos)++];
        if (shift < 64) {
            v |= (uint64_t) (byte & 0x7f) << shift;
        }
        if ((byte & 0x80) == 0) {
            *value = v;
            return 0;
        }
    }
    return -1; /* more than ten bytes */
}

static const struct field_layout *find_field(const struct message_layout *layout, uint64_t number) {
    for (size_t i = 0; i < layout->n_fields; i++) {
        if (layout->fields[i].number == number) {
            return &layout->fields[i];
        }
    }
    return NULL;
}

static void *unpack_message(const struct message_layout *layout, ProtobufCAllocator *allocator, size_t len, const uint8_t *data) {
    char *message = wire_alloc(allocator, layout->size);
    if (message == NULL) {
        return NULL;
    }
    memcpy(message, layout->defaults, layout->size);
    uint64_t seen = 0; /* bit i: field i of the layout was present */
    size_t pos = 0;
    int ok = 1;
    while (ok && pos < len) {
        uint64_t key = 0;
        if (get_varint(data, len, &pos, &key) != 0) {
            ok = 0;
            break;
        }
        const unsigned wire_type = (unsigned) (key & 7);
        const struct field_layout *f = find_field(layout, key >> 3);
        if ((key >> 3) == 0) {
            ok = 0;
            break;
        }
        uint64_t value = 0;
        const uint8_t *payload = NULL;
        size_t payload_len = 0;
        switch (wire_type) {
        case 0:
            ok = get_varint(data, len, &pos, &value) == 0;
            break;
        case 1:
            ok = len - pos >= 8;
            pos += 8;
            break;
        case 2:
            ok = get_varint(data, len, &pos, &value) == 0 && value <= len - pos;
            if (ok) {
                payload = data + pos;
                payload_len = (size_t) value;
                pos += payload_len;
            }
            break;
        case 5:
            ok = len - pos >= 4;
            pos += 4;
            break;
        default:
            ok = 0; /* groups are not part of this API */
            break;
        }
        if (!ok || f == NULL) {
            continue; /* unknown field: skipped */
        }
        const int is_delimited = f->type == F_STRING || f->type == F_BYTES || f->type == F_MESSAGE;
        if ((is_delimited && wire_type != 2) || (!is_delimited && wire_type != 0)) {
            ok = 0; /* a known field with the wrong wire type */
            break;
        }
        void *member = message + f->offset;
        seen |= 1ull << (f - layout->fields);
        switch (f->type) {
        case F_UINT32:
            *(uint32_t *) member = (uint32_t) value;
            break;
        case F_UINT64:
            *(uint64_t *) member = value;
            break;
        case F_INT64:
            *(int64_t *) member = (int64_t) value;
            break;
        case F_BOOL:
            *(protobuf_c_boolean *) member = value != 0;
            break;
        case F_ENUM:
            *(int *) member = (int) (int64_t) value;
            break;
        case F_STRING: {
            char *copy = wire_alloc(allocator, payload_len + 1);
            if (copy == NULL) {
                ok = 0;
                break;
            }
            memcpy(copy, payload, payload_len);
            copy[payload_len] = '\0';
            if (f->label == L_REPEATED) {
                size_t *count = (size_t *) (message + f->count_offset);
                char ***array = member;
                char **grown = wire_alloc(allocator, (*count + 1) * sizeof(char *));
                if (grown == NULL) {
                    wire_free(allocator, copy);
                    ok = 0;
                    break;
                }
                if (*count > 0) {
                    memcpy(grown, *array, *count * sizeof(char *));
                }
                wire_free(allocator, *array);
                grown[*count] = copy;
                *array = grown;
                (*count)++;
            } else {
                wire_free(allocator, *(char **) member);
                *(char **) member = copy;
            }
            break;
        }
        case F_BYTES: {
            ProtobufCBinaryData *bytes = member;
            uint8_t *copy = wire_alloc(allocator, payload_len);
            if (copy == NULL) {
                ok = 0;
                break;
            }
            if (payload_len > 0) {
                memcpy(copy, payload, payload_len);
            }
            wire_free(allocator, bytes->data);
            bytes->data = copy;
            bytes->len = payload_len;
            break;
        }
        case F_MESSAGE: {
            void *sub = unpack_message(f->message, allocator, payload_len, payload);
            if (sub == NULL) {
                ok = 0;
                break;
            }
            if (*(void **) member != NULL) {
                free_message(f->message, *(void **) member, allocator);
            }
            *(void **) member = sub;
            break;
        }
        }
    }
    for (size_t i = 0; ok && i < layout->n_fields; i++) {
        if (layout->fields[i].label == L_REQUIRED && (seen & (1ull << i)) == 0) {
            ok = 0; /* proto2: a message without one of its required fields does not parse */
        }
    }
    if (!ok) {
        free_message(layout, message, allocator);
        return NULL;
    }
    return message;
}

static void free_message(const struct message_layout *layout, void *message, ProtobufCAllocator *allocator) {
    if (message == NULL) {
        return;
    }
    char *base = message;
    for (size_t i = 0; i < layout->n_fields; i++) {
        const struct field_layout *f = &layout->fields[i];
        void *member = base + f->offset;
        if (f->type == F_STRING && f->label == L_REPEATED) {
            const size_t count = *(const size_t *) (base + f->count_offset);
            char **array = *(char ***) member;
            for (size_t k = 0; k < count; k++) {
                wire_free(allocator, array[k]);
            }
            wire_free(allocator, array);
        } else if (f->type == F_STRING) {
            wire_free(allocator, *(char **) member);
        } else if (f->type == F_BYTES) {
            wire_free(allocator, ((ProtobufCBinaryData *) member)->data);
        } else if (f->type == F_MESSAGE) {
            free_message(f->message, *(void **) member, allocator);
        }
    }
    wire_free(allocator, message);
}

/* ---- the generated entry points -------------------------------------------------------------------------------------------- */

#define MESSAGE_API(Type, prefix, layout)                                                                        \
    void prefix##__init(Type *message) { memcpy(message, (layout).defaults, sizeof(Type)); }                     \
    size_t prefix##__get_packed_size(const Type *message) { return message_size(&(layout), message); }           \
    size_t prefix##__pack(const Type *message, uint8_t *out) { return pack_message(&(layout), message, out); }   \
    size_t prefix##__pack_to_buffer(const Type *message, ProtobufCBuffer *buffer) {                              \
        return pack_to_buffer(&(layout), message, buffer);                                                       \
    }                                                                                                            \
    Type *prefix##__unpack(ProtobufCAllocator *allocator, size_t len, const uint8_t *data) {                     \
        return unpack_message(&(layout), allocator, len, data);                                                  \
    }                                                                                                            \
    void prefix##__free_unpacked(Type *message, ProtobufCAllocator *allocator) {                                 \
        free_message(&(layout), message, allocator);                                                             \
    }

MESSAGE_API(DopplerSettings, doppler_settings, layout_doppler_settings)
MESSAGE_API(FskDemodulationSettings, fsk_demodulation_settings, layout_fsk_demodulation_settings)
MESSAGE_API(FskModulationSettings, fsk_modulation_settings, layout_fsk_modulation_settings)
MESSAGE_API(FileSettings, file_settings, layout_file_settings)
MESSAGE_API(RxRequest, rx_request, layout_rx_request)
MESSAGE_API(TxRequest, tx_request, layout_tx_request)
MESSAGE_API(Response, response, layout_response)
MESSAGE_API(TxData, tx_data, layout_tx_data)

/* ---- sockets (src/tcp_utils.c) ---------------------------------------------------------------------------------------------- */

int tcp_utils_write_data(uint8_t *buffer, size_t total_len_bytes, int client_socket) {
    size_t done = 0;
    while (done < total_len_bytes) {
        const ssize_t written = write(client_socket, buffer + done, total_len_bytes - done);
        if (written < 0) {
            return -1;
        }
        done += (size_t) written;
    }
    return 0;
}

int tcp_utils_read_data_partially(void *result, size_t len_bytes, size_t *actually_read, int client_socket) {
    size_t done = 0;
    int code = 0;
    while (done < len_bytes) {
        const ssize_t received = recv(client_socket, (char *) result + done, len_bytes - done, 0);
        if (received > 0) {
            done += (size_t) received;
            continue;
        }
        if (received == 0) {
            code = -1; /* the peer closed the connection */
            break;
        }
        if (errno == EINTR) {
            continue;
        }
        /* a receive timeout is reported as such so that the caller can poll its shutdown flag (tcp_server.c:415-430) */
        code = (errno == EWOULDBLOCK || errno == EAGAIN) ? -errno : -1;
        break;
    }
    *actually_read = done;
    return code;
}

int tcp_utils_read_data(void *result, size_t len_bytes, int client_socket) {
    size_t actually_read = 0;
    return tcp_utils_read_data_partially(result, len_bytes, &actually_read, client_socket);
}

/* ---- framing (src/api_utils.c) ----------------------------------------------------------------------------------------------- */

static const uint32_t MAX_MESSAGE_LENGTH = 32 * 1024;

int api_utils_read_header(int socket, struct message_header *header) {
    const int code = tcp_utils_read_data(header, sizeof(*header), socket);
    if (code == 0) {
        header->message_length = ntohl(header->message_length);
    }
    return code;
}

/* reads the body announced by the header into a malloc'ed buffer */
static int read_body(int socket, const struct message_header *header, uint8_t **body) {
    if (header->message_length > MAX_MESSAGE_LENGTH) {
        return -1;
    }
    uint8_t *buffer = malloc(header->message_length == 0 ? 1 : header->message_length);
    if (buffer == NULL) {
        return -ENOMEM;
    }
    if (tcp_utils_read_data(buffer, header->message_length, socket) != 0) {
        free(buffer);
        return -1;
    }
    *body = buffer;
    return 0;
}

int api_utils_read_rx_request(int socket, const struct message_header *header, struct RxRequest **request) {
    uint8_t *body = NULL;
    const int code = read_body(socket, header, &body);
    if (code != 0) {
        return code;
    }
    RxRequest *result = rx_request__unpack(NULL, header->message_length, body);
    free(body);
    if (result == NULL) {
        return -1;
    }
    *request = result;
    return 0;
}

int api_utils_read_tx_request(int socket, const struct message_header *header, struct TxRequest **request) {
    uint8_t *body = NULL;
    const int code = read_body(socket, header, &body);
    if (code != 0) {
        return code;
    }
    TxRequest *result = tx_request__unpack(NULL, header->message_length, body);
    free(body);
    if (result == NULL) {
        return -1;
    }
    *request = result;
    return 0;
}

int api_utils_read_tx_data(int socket, const struct message_header *header, struct TxData **request) {
    uint8_t *body = NULL;
    const int code = read_body(socket, header, &body);
    if (code != 0) {
        return code;
    }
    TxData *result = tx_data__unpack(NULL, header->message_length, body);
    free(body);
    if (result == NULL) {
        return -1;
    }
    *request = result;
    return 0;
}

int api_utils_write_response(int socket, ResponseStatus status, uint32_t details) {
    Response response;
    response__init(&response);
    response.status = status;
    response.details = details;
    const size_t len = response__get_packed_size(&response);
    uint8_t frame[sizeof(struct message_header) + 32];
    if (len > 32) {
        return -1;
    }
    struct message_header header;
    header.protocol_version = PROTOCOL_VERSION;
    header.type = TYPE_RESPONSE;
    header.message_length = htonl((uint32_t) len);
    memcpy(frame, &header, sizeof(header));
    response__pack(&response, frame + sizeof(header));
    return tcp_utils_write_data(frame, sizeof(header) + len, socket);
}

void api_utils_convert_tle(char **tle, char (*output)[80]) {
    for (int i = 0; i < 3; i++) {
        strncpy(output[i], tle[i], 80);
    }
}
