#!/usr/bin/env python3
"""Generates tests/golden/wire_vectors.json: known-answer bytes for the reference's TCP API messages (api.proto, proto2),
serialised by an independent protobuf implementation (the `protobuf` Python package, google.protobuf), so that the
hand-written codec of sdr-modem_b200/host/wire.c is pinned to real protobuf bytes and not only to its own round trip.

The message schema is built programmatically from a FileDescriptorProto that restates /root/reference/api.proto field by field
(no protoc here). Run in the build container:  python tests/golden/make_wire_fixtures.py
"""
import json
import os

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

F = descriptor_pb2.FieldDescriptorProto
REQ, OPT, REP = F.LABEL_REQUIRED, F.LABEL_OPTIONAL, F.LABEL_REPEATED


def schema():
    fd = descriptor_pb2.FileDescriptorProto(name="api.proto", syntax="proto2")
    e = fd.enum_type.add(name="modem_type")
    e.value.add(name="GMSK", number=1)
    e = fd.enum_type.add(name="demod_destination")
    for i, n in enumerate(("FILE", "SOCKET", "BOTH")):
        e.value.add(name=n, number=i)
    e = fd.enum_type.add(name="response_status")
    for i, n in enumerate(("SUCCESS", "FAILURE")):
        e.value.add(name=n, number=i)

    def message(name, fields):
        m = fd.message_type.add(name=name)
        for number, (fname, ftype, label, type_name) in enumerate(fields, start=1):
            f = m.field.add(name=fname, number=number, type=ftype, label=label)
            if type_name:
                f.type_name = type_name

    message("doppler_settings", [("tle", F.TYPE_STRING, REP, None), ("latitude", F.TYPE_UINT32, REQ, None),
                                 ("longitude", F.TYPE_UINT32, REQ, None), ("altitude", F.TYPE_UINT32, REQ, None)])
    message("fsk_demodulation_settings", [("demod_fsk_deviation", F.TYPE_INT64, REQ, None),
                                          ("demod_fsk_transition_width", F.TYPE_UINT32, REQ, None),
                                          ("demod_fsk_use_dc_block", F.TYPE_BOOL, REQ, None)])
    message("fsk_modulation_settings", [("mod_fsk_deviation", F.TYPE_INT64, REQ, None)])
    message("file_settings", [("filename", F.TYPE_STRING, REQ, None), ("start_time_seconds", F.TYPE_UINT64, REQ, None)])
    message("RxRequest", [("rx_center_freq", F.TYPE_UINT64, REQ, None), ("rx_sampling_freq", F.TYPE_UINT64, REQ, None),
                          ("rx_dump_file", F.TYPE_BOOL, REQ, None), ("rx_offset", F.TYPE_INT64, REQ, None),
                          ("demod_type", F.TYPE_ENUM, REQ, ".modem_type"), ("demod_baud_rate", F.TYPE_UINT32, REQ, None),
                          ("demod_decimation", F.TYPE_UINT32, REQ, None),
                          ("demod_destination", F.TYPE_ENUM, REQ, ".demod_destination"),
                          ("doppler", F.TYPE_MESSAGE, OPT, ".doppler_settings"),
                          ("fsk_settings", F.TYPE_MESSAGE, OPT, ".fsk_demodulation_settings"),
                          ("file_settings", F.TYPE_MESSAGE, OPT, ".file_settings")])
    message("TxRequest", [("tx_center_freq", F.TYPE_UINT64, REQ, None), ("tx_sampling_freq", F.TYPE_UINT64, REQ, None),
                          ("tx_dump_file", F.TYPE_BOOL, REQ, None), ("tx_offset", F.TYPE_INT64, REQ, None),
                          ("mod_type", F.TYPE_ENUM, REQ, ".modem_type"), ("mod_baud_rate", F.TYPE_UINT32, REQ, None),
                          ("doppler", F.TYPE_MESSAGE, OPT, ".doppler_settings"),
                          ("fsk_settings", F.TYPE_MESSAGE, OPT, ".fsk_modulation_settings"),
                          ("file_settings", F.TYPE_MESSAGE, OPT, ".file_settings")])
    message("Response", [("status", F.TYPE_ENUM, REQ, ".response_status"), ("details", F.TYPE_UINT32, REQ, None)])
    message("TxData", [("data", F.TYPE_BYTES, REQ, None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return {m.name: message_factory.GetMessageClass(pool.FindMessageTypeByName(m.name)) for m in fd.message_type}


TLE = ["LUCKY-7", "1 44406U 19038W   20069.88080907  .00000505  00000-0  32890-4 0  9992",
       "2 44406  97.5270  32.5584 0026284 107.4758 252.9348 15.12089395 37524"]


def main():
    cls = schema()
    vectors = []

    def add(name, message, fields):
        vectors.append({"name": name, "type": type(message).DESCRIPTOR.name, "fields": fields,
                        "hex": message.SerializeToString(deterministic=True).hex()})

    # the request of reference test/utils.c:6-55 (create_rx_request)
    rx = cls["RxRequest"](rx_center_freq=437525000, rx_sampling_freq=48000, rx_dump_file=False, rx_offset=0, demod_type=1,
                          demod_baud_rate=4800, demod_decimation=2, demod_destination=1)
    rx.doppler.tle.extend(TLE)
    rx.doppler.latitude, rx.doppler.longitude, rx.doppler.altitude = 537200000, 475700000, 0
    rx.fsk_settings.demod_fsk_deviation, rx.fsk_settings.demod_fsk_transition_width = 5000, 2000
    rx.fsk_settings.demod_fsk_use_dc_block = True
    rx.file_settings.filename, rx.file_settings.start_time_seconds = "/tmp/tx.cf32", 0
    add("rx_request_test_utils", rx, {
        "rx_center_freq": 437525000, "rx_sampling_freq": 48000, "rx_dump_file": 0, "rx_offset": 0, "demod_type": 1,
        "demod_baud_rate": 4800, "demod_decimation": 2, "demod_destination": 1,
        "doppler": {"tle": TLE, "latitude": 537200000, "longitude": 475700000, "altitude": 0},
        "fsk_settings": {"demod_fsk_deviation": 5000, "demod_fsk_transition_width": 2000, "demod_fsk_use_dc_block": 1},
        "file_settings": {"filename": "/tmp/tx.cf32", "start_time_seconds": 0}})
    # negative int64s (ten-byte varints), large values, no optional parts
    rx2 = cls["RxRequest"](rx_center_freq=2 ** 40 + 7, rx_sampling_freq=2400000, rx_dump_file=True, rx_offset=-12500, demod_type=1,
                           demod_baud_rate=2400, demod_decimation=100, demod_destination=2)
    add("rx_request_negative_offset_no_optionals", rx2, {
        "rx_center_freq": 2 ** 40 + 7, "rx_sampling_freq": 2400000, "rx_dump_file": 1, "rx_offset": -12500, "demod_type": 1,
        "demod_baud_rate": 2400, "demod_decimation": 100, "demod_destination": 2, "doppler": None, "fsk_settings": None,
        "file_settings": None})
    # the request of reference test/utils.c:57-103 (create_tx_request)
    tx = cls["TxRequest"](tx_center_freq=437525000, tx_sampling_freq=580000, tx_dump_file=False, tx_offset=0, mod_type=1,
                          mod_baud_rate=4800)
    tx.doppler.tle.extend(TLE)
    tx.doppler.latitude, tx.doppler.longitude, tx.doppler.altitude = 537200000, 475700000, 0
    tx.fsk_settings.mod_fsk_deviation = 5000
    tx.file_settings.filename, tx.file_settings.start_time_seconds = "/tmp/tx.cf32", 1583840449
    add("tx_request_test_utils", tx, {
        "tx_center_freq": 437525000, "tx_sampling_freq": 580000, "tx_dump_file": 0, "tx_offset": 0, "mod_type": 1,
        "mod_baud_rate": 4800, "doppler": {"tle": TLE, "latitude": 537200000, "longitude": 475700000, "altitude": 0},
        "fsk_settings": {"mod_fsk_deviation": 5000}, "file_settings": {"filename": "/tmp/tx.cf32", "start_time_seconds": 1583840449}})
    tx2 = cls["TxRequest"](tx_center_freq=1, tx_sampling_freq=2, tx_dump_file=True, tx_offset=-1, mod_type=1, mod_baud_rate=9600)
    tx2.fsk_settings.mod_fsk_deviation = -5000
    add("tx_request_negative_deviation", tx2, {
        "tx_center_freq": 1, "tx_sampling_freq": 2, "tx_dump_file": 1, "tx_offset": -1, "mod_type": 1, "mod_baud_rate": 9600,
        "doppler": None, "fsk_settings": {"mod_fsk_deviation": -5000}, "file_settings": None})
    for status, details in ((0, 0), (1, 1), (1, 5), (0, 4294967295)):
        add("response_%d_%d" % (status, details), cls["Response"](status=status, details=details),
            {"status": status, "details": details})
    for n in (0, 1, 127, 128, 300, 2048):
        data = bytes((i * 7 + 3) & 0xFF for i in range(n))
        add("tx_data_%d" % n, cls["TxData"](data=data), {"data_hex": data.hex()})
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wire_vectors.json")
    with open(out, "w") as f:
        json.dump({"generator": "tests/golden/make_wire_fixtures.py, google.protobuf %s" % __import__("google.protobuf").protobuf.__version__,
                   "vectors": vectors}, f, indent=1)
    print("wrote %d vectors to %s" % (len(vectors), out))


if __name__ == "__main__":
    main()
