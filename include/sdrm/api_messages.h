/*
 * Wire messages of the reference's TCP API (api.proto, proto2) without a protobuf dependency: the message structs and the
 * pack / unpack entry points that protoc-c generates for the reference (src/api.pb-c.h), implemented by hand in
 * sdr-modem_b200/host/wire.c on the proto2 wire format (varints, length-delimited fields).
 *
 * Names, field order and field types are those of the generated header because they are the ABI the reference's callers
 * are compiled against (src/tcp_server.c:455-570, src/api_utils.c:19-112, src/dsp_worker.c:108-197, test/utils.c:6-103):
 * a host that has protobuf-c keeps including its own api.pb-c.h and simply links this library instead of api.pb-c.c and
 * libprotobuf-c; a host that has not includes this header. Message <-> field-number mapping (api.proto):
 *
 *   doppler_settings           1 tle (repeated string)  2 latitude  3 longitude  4 altitude (uint32, degrees / km scaled)
 *   fsk_demodulation_settings  1 demod_fsk_deviation (int64)  2 demod_fsk_transition_width (uint32)  3 demod_fsk_use_dc_block
 *   fsk_modulation_settings    1 mod_fsk_deviation (int64)
 *   file_settings              1 filename (string)  2 start_time_seconds (uint64)
 *   RxRequest                  1 rx_center_freq  2 rx_sampling_freq (uint64)  3 rx_dump_file (bool)  4 rx_offset (int64)
 *                              5 demod_type (enum)  6 demod_baud_rate  7 demod_decimation (uint32)  8 demod_destination (enum)
 *                              9 doppler  10 fsk_settings  11 file_settings (optional messages)
 *   TxRequest                  1 tx_center_freq  2 tx_sampling_freq  3 tx_dump_file  4 tx_offset  5 mod_type  6 mod_baud_rate
 *                              7 doppler  8 fsk_settings  9 file_settings
 *   Response                   1 status (enum)  2 details (uint32)
 *   TxData                     1 data (bytes)
 *
 * Semantics follow protobuf-c: unpack returns NULL for malformed input or a missing required field, unknown fields are
 * skipped, everything an unpacked message points to is released by X__free_unpacked; allocator == NULL means malloc / free.
 */
#ifndef SDRM_API_MESSAGES_H
#define SDRM_API_MESSAGES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the few protobuf-c base types the generated structs embed; skipped when the real (or a stand-in) header is already there */
#if !defined(PROTOBUF_C_H) && !defined(PROTOBUF_C_PROTOBUF_C_H) && !defined(SDRM_PROTOBUF_C_SHIM_H)
#define SDRM_OWN_PROTOBUF_C_TYPES 1
typedef int protobuf_c_boolean;
typedef struct ProtobufCMessageDescriptor ProtobufCMessageDescriptor;
typedef struct ProtobufCEnumDescriptor ProtobufCEnumDescriptor;
typedef struct ProtobufCMessageUnknownField ProtobufCMessageUnknownField;
typedef struct ProtobufCEnumValue {
    const char *name;   /* as in api.proto: "GMSK", "SOCKET", ... */
    const char *c_name; /* the C enumerator: "MODEM_TYPE__GMSK", ... */
    int value;
} ProtobufCEnumValue;
typedef struct ProtobufCAllocator {
    void *(*alloc)(void *allocator_data, size_t size);
    void (*free)(void *allocator_data, void *pointer);
    void *allocator_data;
} ProtobufCAllocator;
typedef struct ProtobufCBuffer {
    void (*append)(struct ProtobufCBuffer *buffer, size_t len, const uint8_t *data);
} ProtobufCBuffer;
typedef struct ProtobufCBinaryData {
    size_t len;
    uint8_t *data;
} ProtobufCBinaryData;
typedef struct ProtobufCMessage {
    const ProtobufCMessageDescriptor *descriptor;
    unsigned n_unknown_fields;
    ProtobufCMessageUnknownField *unknown_fields;
} ProtobufCMessage;
#define PROTOBUF_C_MESSAGE_INIT(descriptor) { descriptor, 0, NULL }
#endif

typedef struct DopplerSettings DopplerSettings;
typedef struct FskDemodulationSettings FskDemodulationSettings;
typedef struct FskModulationSettings FskModulationSettings;
typedef struct FileSettings FileSettings;
typedef struct RxRequest RxRequest;
typedef struct TxRequest TxRequest;
typedef struct Response Response;
typedef struct TxData TxData;

typedef enum _ModemType { MODEM_TYPE__GMSK = 1, _MODEM_TYPE_IS_INT_SIZE = INT32_MAX } ModemType;
typedef enum _DemodDestination {
    DEMOD_DESTINATION__FILE = 0,
    DEMOD_DESTINATION__SOCKET = 1,
    DEMOD_DESTINATION__BOTH = 2,
    _DEMOD_DESTINATION_IS_INT_SIZE = INT32_MAX
} DemodDestination;
typedef enum _ResponseStatus {
    RESPONSE_STATUS__SUCCESS = 0,
    RESPONSE_STATUS__FAILURE = 1,
    _RESPONSE_STATUS_IS_INT_SIZE = INT32_MAX
} ResponseStatus;

struct DopplerSettings {
    ProtobufCMessage base;
    size_t n_tle;
    char **tle;
    uint32_t latitude;  /* degrees times 10^6 (the reference divides by 10E6, src/dsp_worker.c:130) */
    uint32_t longitude;
    uint32_t altitude;
};
struct FskDemodulationSettings {
    ProtobufCMessage base;
    int64_t demod_fsk_deviation;
    uint32_t demod_fsk_transition_width;
    protobuf_c_boolean demod_fsk_use_dc_block;
};
struct FskModulationSettings {
    ProtobufCMessage base;
    int64_t mod_fsk_deviation;
};
struct FileSettings {
    ProtobufCMessage base;
    char *filename;
    uint64_t start_time_seconds;
};
struct RxRequest {
    ProtobufCMessage base;
    uint64_t rx_center_freq;
    uint64_t rx_sampling_freq;
    protobuf_c_boolean rx_dump_file;
    int64_t rx_offset;
    ModemType demod_type;
    uint32_t demod_baud_rate;
    uint32_t demod_decimation; /* the actual is uint8 */
    DemodDestination demod_destination;
    DopplerSettings *doppler;
    FskDemodulationSettings *fsk_settings;
    FileSettings *file_settings;
};
struct TxRequest {
    ProtobufCMessage base;
    uint64_t tx_center_freq;
    uint64_t tx_sampling_freq;
    protobuf_c_boolean tx_dump_file;
    int64_t tx_offset;
    ModemType mod_type;
    uint32_t mod_baud_rate;
    DopplerSettings *doppler;
    FskModulationSettings *fsk_settings;
    FileSettings *file_settings;
};
struct Response {
    ProtobufCMessage base;
    ResponseStatus status;
    uint32_t details;
};
struct TxData {
    ProtobufCMessage base;
    ProtobufCBinaryData data;
};

extern const ProtobufCMessageDescriptor doppler_settings__descriptor;
extern const ProtobufCMessageDescriptor fsk_demodulation_settings__descriptor;
extern const ProtobufCMessageDescriptor fsk_modulation_settings__descriptor;
extern const ProtobufCMessageDescriptor file_settings__descriptor;
extern const ProtobufCMessageDescriptor rx_request__descriptor;
extern const ProtobufCMessageDescriptor tx_request__descriptor;
extern const ProtobufCMessageDescriptor response__descriptor;
extern const ProtobufCMessageDescriptor tx_data__descriptor;

extern const ProtobufCEnumDescriptor modem_type__descriptor;
extern const ProtobufCEnumDescriptor demod_destination__descriptor;
extern const ProtobufCEnumDescriptor response_status__descriptor;
/* the one protobuf-c run time call the reference's server makes on them (names for its log lines, src/tcp_server.c:607,679):
 * the value's entry, NULL when the enum has no such value */
const ProtobufCEnumValue *protobuf_c_enum_descriptor_get_value(const ProtobufCEnumDescriptor *desc, int value);

#define DOPPLER_SETTINGS__INIT { PROTOBUF_C_MESSAGE_INIT(&doppler_settings__descriptor), 0, NULL, 0, 0, 0 }
#define FSK_DEMODULATION_SETTINGS__INIT { PROTOBUF_C_MESSAGE_INIT(&fsk_demodulation_settings__descriptor), 0, 0, 0 }
#define FSK_MODULATION_SETTINGS__INIT { PROTOBUF_C_MESSAGE_INIT(&fsk_modulation_settings__descriptor), 0 }
#define FILE_SETTINGS__INIT { PROTOBUF_C_MESSAGE_INIT(&file_settings__descriptor), NULL, 0 }
#define RX_REQUEST__INIT \
    { PROTOBUF_C_MESSAGE_INIT(&rx_request__descriptor), 0, 0, 0, 0, MODEM_TYPE__GMSK, 0, 0, DEMOD_DESTINATION__FILE, NULL, NULL, NULL }
#define TX_REQUEST__INIT { PROTOBUF_C_MESSAGE_INIT(&tx_request__descriptor), 0, 0, 0, 0, MODEM_TYPE__GMSK, 0, NULL, NULL, NULL }
#define RESPONSE__INIT { PROTOBUF_C_MESSAGE_INIT(&response__descriptor), RESPONSE_STATUS__SUCCESS, 0 }
#define TX_DATA__INIT { PROTOBUF_C_MESSAGE_INIT(&tx_data__descriptor), { 0, NULL } }

/* the six entry points protoc-c generates per message: init, packed size, pack to memory / to a buffer, unpack, free */
void doppler_settings__init(DopplerSettings *message);
size_t doppler_settings__get_packed_size(const DopplerSettings *message);
size_t doppler_settings__pack(const DopplerSettings *message, uint8_t *out);
size_t doppler_settings__pack_to_buffer(const DopplerSettings *message, ProtobufCBuffer *buffer);
DopplerSettings *doppler_settings__unpack(ProtobufCAllocator *allocator, size_t len, const uint8_t *data);
void doppler_settings__free_unpacked(DopplerSettings *message, ProtobufCAllocator *allocator);

void fsk_demodulation_settings__init(FskDemodulationSettings *message);
size_t fsk_demodulation_settings__get_packed_size(const FskDemodulationSettings *message);
size_t fsk_demodulation_settings__pack(const FskDemodulationSettings *message, uint8_t *out);
size_t fsk_demodulation_settings__pack_to_buffer(const FskDemodulationSettings *message, ProtobufCBuffer *buffer);
FskDemodulationSettings *fsk_demodulation_settings__unpack(ProtobufCAllocator *allocator, size_t len, const uint8_t *data);
void fsk_demodulation_settings__free_unpacked(FskDemodulationSettings *message, ProtobufCAllocator *allocator);

void fsk_modulation_settings__init(FskModulationSettings *message);
size_t fsk_modulation_settings__get_packed_size(const FskModulationSettings *message);
size_t fsk_modulation_settings__pack(const FskModulationSettings *message, uint8_t *out);
size_t fsk_modulation_settings__pack_to_buffer(const FskModulationSettings *message, ProtobufCBuffer *buffer);
FskModulationSettings *fsk_modulation_settings__unpack(ProtobufCAllocator *allocator, size_t len, const uint8_t *data);
void fsk_modulation_settings__free_unpacked(FskModulationSettings *message, ProtobufCAllocator *allocator);

void file_settings__init(FileSettings *message);
size_t file_settings__get_packed_size(const FileSettings *message);
size_t file_settings__pack(const FileSettings *message, uint8_t *out);
size_t file_settings__pack_to_buffer(const FileSettings *message, ProtobufCBuffer *buffer);
FileSettings *file_settings__unpack(ProtobufCAllocator *allocator, size_t len, const uint8_t *data);
void file_settings__free_unpacked(FileSettings *message, ProtobufCAllocator *allocator);

void rx_request__init(RxRequest *message);
size_t rx_request__get_packed_size(const RxRequest *message);
size_t rx_request__pack(const RxRequest *message, uint8_t *out);
size_t rx_request__pack_to_buffer(const RxRequest *message, ProtobufCBuffer *buffer);
RxRequest *rx_request__unpack(ProtobufCAllocator *allocator, size_t len, const uint8_t *data);
void rx_request__free_unpacked(RxRequest *message, ProtobufCAllocator *allocator);

void tx_request__init(TxRequest *message);
size_t tx_request__get_packed_size(const TxRequest *message);
size_t tx_request__pack(const TxRequest *message, uint8_t *out);
size_t tx_request__pack_to_buffer(const TxRequest *message, ProtobufCBuffer *buffer);
TxRequest *tx_request__unpack(ProtobufCAllocator *allocator, size_t len, const uint8_t *data);
void tx_request__free_unpacked(TxRequest *message, ProtobufCAllocator *allocator);

void response__init(Response *message);
size_t response__get_packed_size(const Response *message);
size_t response__pack(const Response *message, uint8_t *out);
size_t response__pack_to_buffer(const Response *message, ProtobufCBuffer *buffer);
Response *response__unpack(ProtobufCAllocator *allocator, size_t len, const uint8_t *data);
void response__free_unpacked(Response *message, ProtobufCAllocator *allocator);

void tx_data__init(TxData *message);
size_t tx_data__get_packed_size(const TxData *message);
size_t tx_data__pack(const TxData *message, uint8_t *out);
size_t tx_data__pack_to_buffer(const TxData *message, ProtobufCBuffer *buffer);
TxData *tx_data__unpack(ProtobufCAllocator *allocator, size_t len, const uint8_t *data);
void tx_data__free_unpacked(TxData *message, ProtobufCAllocator *allocator);

#ifdef __cplusplus
}
#endif

#endif
