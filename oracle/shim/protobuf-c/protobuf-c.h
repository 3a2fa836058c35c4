/*
 * TEST INFRASTRUCTURE ONLY. The reference's test/utils.h includes src/api.pb-c.h, which includes
 * <protobuf-c/protobuf-c.h> (not installed here). The DSP tests never call protobuf; this header only lets the generated
 * declarations parse.
 */
#ifndef SDRM_PROTOBUF_C_SHIM_H
#define SDRM_PROTOBUF_C_SHIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define PROTOBUF_C__BEGIN_DECLS extern "C" {
#define PROTOBUF_C__END_DECLS }
#else
#define PROTOBUF_C__BEGIN_DECLS
#define PROTOBUF_C__END_DECLS
#endif

#define PROTOBUF_C_VERSION_NUMBER 1004000
#define PROTOBUF_C_MIN_COMPILER_VERSION 1000000
#define PROTOBUF_C__FORCE_ENUM_TO_BE_INT_SIZE(enum_name) , _##enum_name##_IS_INT_SIZE = INT32_MAX
#define PROTOBUF_C_MESSAGE_INIT(descriptor) { descriptor, 0, NULL }

typedef int protobuf_c_boolean;

typedef struct ProtobufCAllocator ProtobufCAllocator;
typedef struct ProtobufCBuffer ProtobufCBuffer;
typedef struct ProtobufCEnumDescriptor ProtobufCEnumDescriptor;
typedef struct ProtobufCMessageDescriptor ProtobufCMessageDescriptor;
typedef struct ProtobufCMessageUnknownField ProtobufCMessageUnknownField;

struct ProtobufCEnumDescriptor {
    uint32_t magic;
};

struct ProtobufCMessageDescriptor {
    uint32_t magic;
};

typedef struct ProtobufCBinaryData {
    size_t len;
    uint8_t *data;
} ProtobufCBinaryData;

typedef struct ProtobufCMessage {
    const ProtobufCMessageDescriptor *descriptor;
    unsigned n_unknown_fields;
    ProtobufCMessageUnknownField *unknown_fields;
} ProtobufCMessage;

typedef void (*ProtobufCClosure)(const ProtobufCMessage *, void *closure_data);

/* protobuf-c run time call the reference's TCP server makes to print enum names (src/tcp_server.c:607,679); the product
 * library implements it for the three enums of api.proto */
typedef struct ProtobufCEnumValue {
    const char *name;
    const char *c_name;
    int value;
} ProtobufCEnumValue;
const ProtobufCEnumValue *protobuf_c_enum_descriptor_get_value(const ProtobufCEnumDescriptor *desc, int value);

#endif
