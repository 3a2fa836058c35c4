"""GPU parity tests for the Doppler correction (reference src/dsp/doppler.c): the reference's golden pair
lucky7.cf32 <-> lucky7.expected.cf32 (test/test_doppler.c:39-115) and, when the reference build travelled with the
repo, a long low-rate run that crosses many one-second schedule updates (each one an SGP4 evaluation on the host whose
result is truncated to integer Hz: a single differing Hz would rotate the rest of the stream away)."""
import ctypes as C

import numpy as np
import pytest

from conftest import LUCKY7_TLE, golden_array
from test_gpu_blocks import SZ, VP
from test_gpu_mod_nco import close_trig

pytestmark = pytest.mark.gpu


class DopplerHandle:
    def __init__(self, lib, lat, lon, alt, fs, fc, offset, start, max_len, tle):
        self.lib = lib
        lib.doppler_create.argtypes = [C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, C.c_uint32,
                                       VP, C.POINTER(VP)]
        for name in ("doppler_process_rx", "doppler_process_tx"):
            getattr(lib, name).argtypes = [VP, SZ, C.POINTER(VP), C.POINTER(SZ), VP]
            getattr(lib, name).restype = None
        lib.doppler_destroy.argtypes = [VP]
        buf = C.create_string_buffer(240)
        for i, line in enumerate(tle):
            raw = line.encode("ascii")[:79]
            buf[i * 80:i * 80 + len(raw)] = raw
        self.h = VP()
        code = lib.doppler_create(lat, lon, alt, fs, fc, offset, start, max_len, buf, C.byref(self.h))
        if code != 0:
            raise RuntimeError("doppler_create failed with %d" % code)

    def run(self, x, chunk, tx=False):
        fn = self.lib.doppler_process_tx if tx else self.lib.doppler_process_rx
        parts = []
        for o in range(0, len(x), chunk):
            part = np.ascontiguousarray(x[o:o + chunk], dtype=np.complex64)
            out, n = VP(), SZ()
            fn(part.ctypes.data_as(VP), len(part), C.byref(out), C.byref(n), self.h)
            parts.append(np.frombuffer((C.c_char * (n.value * 8)).from_address(out.value), dtype=np.complex64).copy())
        return np.concatenate(parts)

    def close(self):
        self.lib.doppler_destroy(self.h)


LAT, LON = float(np.float32(53.72)), float(np.float32(47.57))


def test_doppler_rx_tx_goldens(sdrm):
    """test/test_doppler.c:39-63 (rx) and :87-115 (tx), 2000-sample calls, reference tolerance 1e-2"""
    raw = golden_array("lucky7.cf32", np.complex64)
    corrected = golden_array("lucky7.expected.cf32", np.complex64)
    d = DopplerHandle(sdrm.lib, LAT, LON, 0.0, 48000, 437525000, 0, 1583840449, 2000, LUCKY7_TLE)
    y = d.run(raw, 2000)
    d.close()
    assert len(y) == len(corrected) and np.abs(y - corrected).max() < 1e-2
    d = DopplerHandle(sdrm.lib, LAT, LON, 0.0, 48000, 437525000, 0, 1583840449, 2000, LUCKY7_TLE)
    back = d.run(corrected, 2000, tx=True)
    d.close()
    assert np.abs(back - raw).max() < 1e-2


def test_doppler_create_errors(sdrm):
    bad = [LUCKY7_TLE[0], LUCKY7_TLE[1][:-1] + "0", LUCKY7_TLE[2]]  # broken checksum (reference test_dsp_worker.c:38-50)
    with pytest.raises(RuntimeError):
        DopplerHandle(sdrm.lib, LAT, LON, 0.0, 48000, 437525000, 0, 1583840449, 2000, bad)
    d = DopplerHandle(sdrm.lib, LAT, LON, 0.0, 48000, 437525000, 0, 1583840449, 100, LUCKY7_TLE)
    out, n = VP(), SZ()
    x = np.zeros(101, np.complex64)
    sdrm.lib.doppler_process_rx(x.ctypes.data_as(VP), 101, C.byref(out), C.byref(n), d.h)  # oversize -> NULL / 0
    assert n.value == 0 and not out.value
    sdrm.lib.doppler_process_rx(None, 0, C.byref(out), C.byref(n), d.h)  # NULL / 0 input -> NULL / 0
    assert n.value == 0 and not out.value
    d.close()


def test_doppler_vs_reference_build_long_run(sdrm, ref):
    """300 schedule updates (fs = 2 kHz, 300 s), ragged call sizes that straddle the one-second boundaries."""
    fs, n = 2000, 600000
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    for start, offset, tx in ((1583840449, 0, False), (1583840449 + 400, 1250, True)):
        want = ref.doppler(LAT, LON, 0.1, fs, 437525000, offset, start, 3000, LUCKY7_TLE, tx=tx).run(x, 2777)
        d = DopplerHandle(sdrm.lib, LAT, LON, 0.1, fs, 437525000, offset, start, 3000, LUCKY7_TLE)
        got = d.run(x, 2777, tx=tx)
        d.close()
        assert close_trig(got, want)


def test_doppler_batch_channels_with_their_own_clocks(sdrm, ref):
    fs, n, n_ch = 48000, 100000, 4
    rng = np.random.default_rng(2)
    x = (rng.standard_normal((n_ch, n)) + 1j * rng.standard_normal((n_ch, n))).astype(np.complex64)
    starts = [1583840449 + 37 * c for c in range(n_ch)]
    channels = [sdrm.doppler_channel(LAT, LON, 0.0, 100 * c, starts[c], LUCKY7_TLE) for c in range(n_ch)]
    b = sdrm.DopplerBatch(channels, fs, 437525000, 4096)
    got = np.concatenate([b.process(x[:, o:o + 4096]) for o in range(0, n, 4096)], axis=1)
    b.close()
    for c in range(n_ch):
        want = ref.doppler(LAT, LON, 0.0, fs, 437525000, 100 * c, starts[c], 4096, LUCKY7_TLE).run(x[c], 4096)
        assert close_trig(got[c], want)


# Spacetrack Report #3 SDP4 test case (reference test/resources/test-002.tle): period 10.5 h, e = 0.73, 12 h resonance
MOLNIYA_LIKE_TLE = ["TEST SAT SDP 001",
                    "1 11801U          80230.29629788  .01431103  00000-0  14311-1 0     2",
                    "2 11801  46.7916 230.4354 7318036  47.4722  10.4117  2.28537848     2"]


def test_doppler_deep_space_satellite(sdrm, ref):
    """a deep-space element set goes through SDP4 (reference src/dsp/doppler.c:33-34): 120 one-second schedule updates"""
    fs, n = 2000, 240000
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    start = 335000000 + 86400 * 3  # a few days after the element set's epoch (1980 day 230)
    want = ref.doppler(LAT, LON, 0.0, fs, 145900000, 0, start, 3000, MOLNIYA_LIKE_TLE).run(x, 2500)
    d = DopplerHandle(sdrm.lib, LAT, LON, 0.0, fs, 145900000, 0, start, 3000, MOLNIYA_LIKE_TLE)
    got = d.run(x, 2500)
    d.close()
    assert close_trig(got, want)
