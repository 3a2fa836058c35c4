"""GPU parity tests for the modulator side (gfsk_mod, interp_fir_filter, frequency_modulator) and the NCO / mixer
(sig_source), batch entry points and the reference's handles, through the C ABI.

Float outputs here pass through double-precision cos/sin rounded to float. The GPU's and glibc's double routines are different
code, but tests/test_gpu_sincos_sweep.py shows that they round to the same float for every float phase in [-2 pi, 2 pi] (all
2.17e9 of them), so these tests demand bit equality of the outputs like everything before the trigonometry (phase recurrences,
shaping filter). Phases outside that range can only reach the sig_source / frequency_modulator handles through inputs the
reference's callers never produce.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ramp, same_bits
from test_gpu_blocks import SZ, VP, Block

pytestmark = pytest.mark.gpu


def close_trig(a, b):
    """bit equality (the name is from when a 1e-6 fraction of one-ulp differences was tolerated; the sweep made that unnecessary)"""
    a = np.ascontiguousarray(a).view(np.float32)
    b = np.ascontiguousarray(b).view(np.float32)
    return a.shape == b.shape and same_bits(a, b)


def test_gfsk_mod_reference_kat(sdrm, kats):
    """reference test/test_gfsk_mod.c test_normal and test_exceeded_input"""
    lib = sdrm.lib
    lib.gfsk_mod_create.argtypes = [C.c_float, C.c_float, C.c_float, C.c_uint32, C.POINTER(VP)]
    sps = np.float32(19200) / np.float32(9600)
    sens = np.float32(2 * np.pi * 5000 / np.float32(19200))
    m = Block(lib, "gfsk_mod", (sps, sens, 0.5, 1000), np.uint8, np.complex64)
    y = m.process(np.arange(10, dtype=np.uint8))
    m.close()
    e = kats["test_gfsk_mod.c:test_normal:expected"].view(np.complex64)
    assert len(y) == 160 and np.abs(y - e).max() < 1e-2
    m = Block(lib, "gfsk_mod", (sps, sens, 0.5, 10), np.uint8, np.complex64)
    assert len(m.process(np.arange(11, dtype=np.uint8))) == 0
    m.close()


@pytest.mark.parametrize("sps", [2.0, 20.0, 5.0])
def test_gfsk_mod_batch_vs_oracle(sdrm, port, sps):
    rng = np.random.default_rng(int(sps))
    n_ch = 5
    data = rng.integers(0, 256, (n_ch, 700), dtype=np.uint8)
    sens = float(np.float32(2 * np.pi * 5000 / (9600 * sps)))
    b = sdrm.GfskModBatch(n_ch, sps, sens, 0.5, 400)
    oracles = [port.GfskMod(sps, sens, 0.5, 400) for _ in range(n_ch)]
    for lo, hi in ((0, 400), (400, 401), (401, 401), (401, 700)):  # packets with carried phase and filter state
        got = b.process(data[:, lo:hi])
        for c in range(n_ch):
            want = oracles[c].process(data[c, lo:hi])
            assert got[c].shape == want.shape
            assert close_trig(got[c], want)
    b.close()


@pytest.mark.parametrize("seed", range(16))
def test_gfsk_mod_random_parameter_sets(sdrm, port, seed):
    """gfsk_mod_create's parameter space, sampled: 2..40 samples per symbol (fractional values truncate as in
    src/dsp/gfsk_mod.c:85), BT 0.3 / 0.5 / 1.0, sensitivities from small to more than pi / 2 per sample (frequent phase wraps),
    1..70 channels (partial and several 32-channel groups), packets of random lengths incl. empty ones with carried phase and
    filter history. Sets the reference rejects must be rejected."""
    rng = np.random.default_rng(900 + seed)
    sps = float(rng.choice([2.0, 2.0, 3.0, 4.0, 5.5, 8.0, 10.0, 20.0, 33.0, 40.0]))
    bt = float(rng.choice([0.3, 0.5, 1.0]))
    sens = float(np.float32(rng.uniform(0.01, 2.2)))
    n_ch = int(rng.choice([1, 2, 31, 32, 33, 70]))
    max_bytes = 300
    try:
        oracles = [port.GfskMod(sps, sens, bt, max_bytes) for _ in range(n_ch)]
    except ValueError:
        with pytest.raises(sdrm.SdrmError):
            sdrm.GfskModBatch(n_ch, sps, sens, bt, max_bytes)
        return
    b = sdrm.GfskModBatch(n_ch, sps, sens, bt, max_bytes)
    try:
        for _ in range(6):
            n = int(rng.choice([0, 1, 2, 7, 64, 255, 300]))
            data = rng.integers(0, 256, (n_ch, n), dtype=np.uint8)
            got = b.process(data)
            for c in range(n_ch):
                want = oracles[c].process(data[c])
                assert got[c].shape == want.shape and close_trig(got[c], want), (sps, bt, sens, n_ch, n, c)
    finally:
        b.close()


def test_gfsk_mod_perf_shape_packets(sdrm, port):
    """BASELINE config 4 shape: sps 2, 2048-byte packets, bytes (uint8) i (reference test/perf_fsk_modem.c:23-38)"""
    sps, sens = 2.0, float(np.float32(2 * np.pi * 5000 / 19200))
    packet = (np.arange(2048) & 0xFF).astype(np.uint8)
    b = sdrm.GfskModBatch(3, sps, sens, 0.5, 2048)
    o = port.GfskMod(sps, sens, 0.5, 2048)
    for _ in range(3):
        got = b.process(np.stack([packet] * 3))
        want = o.process(packet)
        for c in range(3):
            assert close_trig(got[c], want)
    b.close()


def test_interp_fir_filter_handle(sdrm, port, kats):
    """reference test/test_interp_fir_filter.c:14-61 (incl. invalid VOLK_ALIGNMENT) + bit-exactness on noise"""
    lib = sdrm.lib
    lib.interp_fir_filter_create.argtypes = [VP, SZ, C.c_uint8, C.c_uint32, C.POINTER(VP)]
    libc = C.CDLL(None)
    libc.malloc.restype = VP
    libc.malloc.argtypes = [SZ]

    def malloc_taps(taps):
        p = libc.malloc(taps.nbytes)
        C.memmove(p, taps.ctypes.data, taps.nbytes)
        return p

    taps = port.gaussian_taps(1.5, 2 * float(np.float32(32000.0) / np.float32(1200)), 0.5, 12)
    f = Block(lib, "interp_fir_filter", (malloc_taps(taps), 12, 2, 1000), np.float32, np.float32)
    y = f.process(ramp(200))
    f.close()
    assert len(y) == 400 and np.abs(y - kats["test_interp_fir_filter.c:test_normal:expected"]).max() < 1e-3
    os.environ["VOLK_ALIGNMENT"] = "invalid"
    try:
        with pytest.raises(RuntimeError):
            Block(lib, "interp_fir_filter", (malloc_taps(taps), 12, 2, 1000), np.float32, np.float32)
    finally:
        del os.environ["VOLK_ALIGNMENT"]
    x = np.random.default_rng(1).standard_normal(3000).astype(np.float32)
    t2 = np.random.default_rng(2).standard_normal(23).astype(np.float32)
    f = Block(lib, "interp_fir_filter", (malloc_taps(t2), 23, 5, 1100), np.float32, np.float32)
    y = f.run(x, 1000)
    f.close()
    # oracle: the same polyphase split as interp_fir_filter.c, through the port's FIR
    padded = np.zeros(25, np.float32)
    padded[:23] = t2
    branches = [port.Fir(padded[p::5], 1, False) for p in range(5)]
    want = np.zeros(len(x) * 5, np.float32)
    for p in range(5):
        want[p::5] = branches[p].run(x, 1000)
    assert same_bits(y, want)


def test_frequency_modulator_handle(sdrm, port, kats):
    lib = sdrm.lib
    lib.frequency_modulator_create.argtypes = [C.c_float, C.c_uint32, C.POINTER(VP)]
    m = Block(lib, "frequency_modulator", (1.2, 1000), np.float32, np.complex64)
    y = m.process(ramp(100))
    e = kats["test_frequency_modulator.c:test_normal:expected"][:200].view(np.complex64)
    assert np.abs(y - e).max() < 1e-2
    m.close()
    m = Block(lib, "frequency_modulator", (1.2, 100), np.float32, np.complex64)
    assert len(m.process(ramp(200))) == 0  # test_input_exceeded
    m.close()
    x = np.random.default_rng(3).standard_normal(20000).astype(np.float32)
    m = Block(lib, "frequency_modulator", (0.8, 4096), np.float32, np.complex64)
    y = m.run(x, 4096)
    m.close()
    assert close_trig(y, port.FreqMod(0.8).process(x))


def test_sig_source_handle(sdrm, port, kats):
    """reference test/test_sig_source.c:8-18 + mixer vs oracle with changing frequency and carried phase"""
    lib = sdrm.lib
    lib.sig_source_create.argtypes = [C.c_float, C.c_uint64, C.c_uint32, C.POINTER(VP)]
    lib.sig_source_process.argtypes = [C.c_int64, SZ, C.POINTER(VP), C.POINTER(SZ), VP]
    lib.sig_source_process.restype = None
    lib.sig_source_multiply.argtypes = [C.c_int64, VP, SZ, C.POINTER(VP), C.POINTER(SZ), VP]
    lib.sig_source_multiply.restype = None
    lib.sig_source_destroy.argtypes = [VP]
    h = VP()
    assert lib.sig_source_create(1.0, 4, 4, C.byref(h)) == 0
    out, n = VP(), SZ()
    lib.sig_source_process(1, 4, C.byref(out), C.byref(n), h)
    y = np.frombuffer((C.c_char * (n.value * 8)).from_address(out.value), dtype=np.complex64).copy()
    lib.sig_source_destroy(h)
    assert np.abs(y - kats["test_sig_source.c:test_success:buffer"].view(np.complex64)).max() < 1e-2

    rng = np.random.default_rng(4)
    x = (rng.standard_normal(9000) + 1j * rng.standard_normal(9000)).astype(np.complex64)
    assert lib.sig_source_create(1.0, 48000, 3000, C.byref(h)) == 0
    o = port.SigSource(1.0, 48000)
    for k, freq in enumerate((4321, -9000, 0)):
        part = np.ascontiguousarray(x[3000 * k:3000 * (k + 1)])
        lib.sig_source_multiply(freq, part.ctypes.data_as(VP), len(part), C.byref(out), C.byref(n), h)
        y = np.frombuffer((C.c_char * (n.value * 8)).from_address(out.value), dtype=np.complex64).copy()
        assert close_trig(y, o.multiply(freq, part))
    lib.sig_source_multiply(1, x.ctypes.data_as(VP), 3001, C.byref(out), C.byref(n), h)  # oversize
    assert n.value == 0 and not out.value
    lib.sig_source_destroy(h)


def test_nco_batch_vs_oracle(sdrm, port):
    n_ch, n = 6, 5000
    rng = np.random.default_rng(6)
    x = (rng.standard_normal((n_ch, 2 * n)) + 1j * rng.standard_normal((n_ch, 2 * n))).astype(np.complex64)
    b = sdrm.NcoBatch(n_ch, 1.0, 2400000, n)
    oracles = [port.SigSource(1.0, 2400000) for _ in range(n_ch)]
    for call in range(2):
        freq = rng.integers(-12000, 12000, n_ch)
        got = b.multiply(freq, x[:, call * n:(call + 1) * n])
        for c in range(n_ch):
            assert close_trig(got[c], oracles[c].multiply(int(freq[c]), x[c, call * n:(call + 1) * n]))
    b.close()
