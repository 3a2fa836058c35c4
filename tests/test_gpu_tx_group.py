"""Batched TX dispatcher (include/sdrm/tx_group.h) against the reference's own transmit chain, session by session:
gfsk_mod_process -> doppler_process_tx | sig_source_multiply(tx_offset) in batches of buffer_size bytes, as handle_tx_data
drives it (src/tcp_server.c:175-241). The checker is the reference build (oracle/_ref)."""
import ctypes as C
import math

import numpy as np
import pytest

from conftest import LUCKY7_TLE

pytestmark = pytest.mark.gpu

VP, SZ = C.c_void_p, C.c_size_t
SINK = C.CFUNCTYPE(C.c_int, VP, C.c_uint32, VP, SZ)
FS, FC, BAUD, DEVIATION = 48000, 437525000, 4800, 5000
LAT_E7, LON_E7 = 537200000, 475700000


class TxSession(C.Structure):
    _fields_ = [("id", C.c_uint32), ("sink", SINK), ("sink_ctx", VP), ("tx_dump_file", C.c_bool), ("tx_offset", C.c_int64),
                ("has_doppler", C.c_bool), ("doppler_tle", (C.c_char * 80) * 3), ("doppler_latitude", C.c_int32),
                ("doppler_longitude", C.c_int32), ("doppler_altitude", C.c_int32), ("file_start_time_seconds", C.c_int64)]


class TxGroupConfig(C.Structure):
    _fields_ = [("tx_center_freq", C.c_uint64), ("tx_sampling_freq", C.c_uint64), ("mod_baud_rate", C.c_uint32),
                ("mod_fsk_deviation", C.c_int64), ("buffer_size", C.c_uint32), ("base_path", C.c_char_p),
                ("output_int16", C.c_bool), ("int16_scalar", C.c_float), ("device", C.c_int)]


def setup(lib):
    lib.sdrm_tx_group_create.argtypes = [C.POINTER(TxGroupConfig), C.POINTER(TxSession), C.c_uint32, C.POINTER(VP)]
    lib.sdrm_tx_group_process.argtypes = [VP, VP, SZ, SZ, C.POINTER(C.c_int)]
    lib.sdrm_tx_group_samples_per_byte.argtypes = [VP]
    lib.sdrm_tx_group_samples_per_byte.restype = SZ
    lib.sdrm_tx_group_destroy.argtypes = [VP]
    lib.sdrm_tx_group_destroy.restype = None


def config(buffer_size, base_path=None, int16=False):
    cfg = TxGroupConfig()
    cfg.tx_center_freq, cfg.tx_sampling_freq, cfg.mod_baud_rate, cfg.mod_fsk_deviation = FC, FS, BAUD, DEVIATION
    cfg.buffer_size, cfg.base_path, cfg.output_int16, cfg.int16_scalar, cfg.device = buffer_size, base_path, int16, 0.0, -1
    return cfg


def make_sessions(specs, sink, dump=()):
    """specs: per session (tx_offset, doppler start time or None)"""
    sessions = (TxSession * len(specs))()
    for i, (offset, start) in enumerate(specs):
        s = sessions[i]
        s.id, s.sink, s.sink_ctx, s.tx_dump_file, s.tx_offset = 100 + i, sink, None, i in dump, offset
        s.has_doppler = start is not None
        for k, line in enumerate(LUCKY7_TLE):
            raw = line.encode("ascii")
            C.memmove(C.addressof(s.doppler_tle[k]), raw + b"\0", len(raw) + 1)
        s.doppler_latitude, s.doppler_longitude, s.doppler_altitude = LAT_E7, LON_E7, 0
        s.file_start_time_seconds = start or 0
    return sessions


def reference_chain(ref, data, offset, start, buffer_size):
    """one session through the reference's own blocks, batch by batch (src/tcp_server.c:186-211)"""
    sps = FS // BAUD
    mod = ref.gfsk_mod(float(FS / BAUD), 2 * math.pi * DEVIATION / FS, 0.5, buffer_size)
    max_out = 8 * sps * buffer_size
    dop = ref.doppler(LAT_E7 / 10E6, LON_E7 / 10E6, 0.0, FS, FC, offset, start, max_out, LUCKY7_TLE, tx=True) if start else None
    sig = ref.sig_source(1.0, FS, max_out) if (start is None and offset != 0) else None
    out = []
    for o in range(0, len(data), buffer_size):
        y = mod.process(data[o:o + buffer_size])
        if dop is not None:
            y = dop.process(y)
        elif sig is not None:
            y = sig.multiply(offset, y)
        out.append(y.copy())
    return np.concatenate(out)


SPECS = [(0, None), (1200, 1583840449), (-3000, None), (0, 1583840449 + 40), (0, None), (2500, None), (-700, 1583840449 + 90)]


def close_enough(got, want):
    # cos/sin are double precision rounded to float on both sides; CUDA's and glibc's doubles may differ in the last place,
    # which moves the float only next to a rounding boundary (about one sample in 1e8)
    same = got.view(np.uint32) == want.view(np.uint32)
    return same.mean() > 0.99999 and np.abs(got - want).max() < 1e-6


def test_tx_group_matches_reference_chain(sdrm, ref, tmp_path):
    lib = sdrm.lib
    setup(lib)
    buffer_size, n_bytes = 512, 1300  # batches of 512, 512, 276 bytes
    rng = np.random.default_rng(4)
    data = rng.integers(0, 256, (len(SPECS), n_bytes + 20), dtype=np.uint8)  # row stride larger than the payload
    collected = {100 + i: [] for i in range(len(SPECS))}

    def on_samples(ctx, session_id, samples, n):
        collected[session_id].append(np.ctypeslib.as_array(C.cast(samples, C.POINTER(C.c_float)), shape=(2 * n,)).copy())
        return 0

    sink = SINK(on_samples)
    sessions = make_sessions(SPECS, sink, dump=(1, 4))
    g = VP()
    cfg = config(buffer_size, str(tmp_path).encode())
    assert lib.sdrm_tx_group_create(C.byref(cfg), sessions, len(SPECS), C.byref(g)) == 0
    assert lib.sdrm_tx_group_samples_per_byte(g) == 8 * (FS // BAUD)
    status = (C.c_int * len(SPECS))(*([7] * len(SPECS)))
    for part in (data[:, :700], data[:, 700:n_bytes]):  # two TxData messages; modulator and NCO state carry over
        part = np.ascontiguousarray(part)
        assert lib.sdrm_tx_group_process(g, part.ctypes.data_as(VP), part.shape[1], part.shape[1], status) == 0
        assert list(status) == [0] * len(SPECS)
    lib.sdrm_tx_group_destroy(g)
    for i, (offset, start) in enumerate(SPECS):
        want = reference_chain_two_messages(ref, data[i, :n_bytes], offset, start, buffer_size, 700)
        got = np.concatenate(collected[100 + i]).view(np.complex64)
        assert len(got) == len(want) == n_bytes * 8 * (FS // BAUD), "session %d" % i
        assert close_enough(got, want), "session %d" % i
        path = tmp_path / ("tx.mod2sdr.%d.cf32" % (100 + i))
        if i in (1, 4):
            assert np.array_equal(np.fromfile(path, dtype=np.complex64).view(np.uint32), got.view(np.uint32))
        else:
            assert not path.exists()


def reference_chain_two_messages(ref, data, offset, start, buffer_size, split):
    """handle_tx_data is called once per TxData message and batches inside each message"""
    sps = FS // BAUD
    mod = ref.gfsk_mod(float(FS / BAUD), 2 * math.pi * DEVIATION / FS, 0.5, buffer_size)
    max_out = 8 * sps * buffer_size
    dop = ref.doppler(LAT_E7 / 10E6, LON_E7 / 10E6, 0.0, FS, FC, offset, start, max_out, LUCKY7_TLE, tx=True) if start else None
    sig = ref.sig_source(1.0, FS, max_out) if (start is None and offset != 0) else None
    out = []
    for message in (data[:split], data[split:]):
        for o in range(0, len(message), buffer_size):
            y = mod.process(message[o:o + buffer_size])
            if dop is not None:
                y = dop.process(y)
            elif sig is not None:
                y = sig.multiply(offset, y)
            out.append(np.array(y, copy=True))
    return np.concatenate(out)


def test_tx_group_int16_egress_and_failing_sink(sdrm, ref, port):
    """PlutoSDR egress format on the device (src/sdr/plutosdr.c:83) and the reference's `unable to transmit request fully`:
    a session whose sink fails gets nothing more from that call, the others are not disturbed."""
    lib = sdrm.lib
    setup(lib)
    buffer_size, n_bytes = 256, 900
    specs = SPECS[:4]
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, (len(specs), n_bytes), dtype=np.uint8)
    collected = {100 + i: [] for i in range(len(specs))}

    def on_samples(ctx, session_id, samples, n):
        if session_id == 102 and len(collected[102]) == 1:
            return -1  # the second batch of session 2 cannot be transmitted
        collected[session_id].append(np.ctypeslib.as_array(C.cast(samples, C.POINTER(C.c_int16)), shape=(2 * n,)).copy())
        return 0

    sink = SINK(on_samples)
    sessions = make_sessions(specs, sink)
    g = VP()
    cfg = config(buffer_size, None, int16=True)
    assert lib.sdrm_tx_group_create(C.byref(cfg), sessions, len(specs), C.byref(g)) == 0
    status = (C.c_int * len(specs))()
    assert lib.sdrm_tx_group_process(g, data.ctypes.data_as(VP), n_bytes, n_bytes, status) == 0
    assert list(status) == [0, 0, 3, 0]
    lib.sdrm_tx_group_destroy(g)
    for i, (offset, start) in enumerate(specs):
        cf = reference_chain(ref, data[i], offset, start, buffer_size)
        want = port.convert_32f_16i(cf.view(np.float32), 32768.0)
        got = np.concatenate(collected[100 + i])
        if i == 2:
            assert len(got) == 2 * buffer_size * 8 * (FS // BAUD)
            want = want[:len(got)]
        assert len(got) == len(want)
        assert np.mean(got == want) > 0.99999 and np.abs(got.astype(int) - want.astype(int)).max() <= 1, "session %d" % i


def test_tx_group_create_failures(sdrm, tmp_path):
    lib = sdrm.lib
    setup(lib)
    sink = SINK(lambda ctx, sid, samples, n: 0)
    g = VP()
    cfg = config(256)
    assert lib.sdrm_tx_group_create(C.byref(cfg), make_sessions(SPECS[:1], sink), 0, C.byref(g)) == -1
    cfg = config(0)
    assert lib.sdrm_tx_group_create(C.byref(cfg), make_sessions(SPECS[:1], sink), 1, C.byref(g)) == -1
    cfg = config(256)
    cfg.mod_baud_rate = 100  # 480 samples per symbol: interp_fir_filter's interpolation is a uint8_t
    assert lib.sdrm_tx_group_create(C.byref(cfg), make_sessions(SPECS[:1], sink), 1, C.byref(g)) == -1
    # a dump file in a directory that does not exist (src/tcp_server.c:569-575)
    cfg = config(256, str(tmp_path / "missing").encode())
    assert lib.sdrm_tx_group_create(C.byref(cfg), make_sessions(SPECS[:1], sink, dump=(0,)), 1, C.byref(g)) == -1
    # broken TLE checksum -> doppler_create fails (src/dsp/doppler.c:104-108)
    sessions = make_sessions(SPECS[1:2], sink)
    bad = (LUCKY7_TLE[1][:-1] + "0").encode("ascii")
    C.memmove(C.addressof(sessions[0].doppler_tle[1]), bad + b"\0", len(bad) + 1)
    cfg = config(256)
    assert lib.sdrm_tx_group_create(C.byref(cfg), sessions, 1, C.byref(g)) == -1
    assert not g.value
