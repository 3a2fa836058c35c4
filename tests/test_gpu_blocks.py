"""GPU parity tests for the reference's per-block handles (lpf, quadrature_demod, dc_blocker, clock_mm) through the
C ABI: the reference's own known-answer arrays, and bit-exactness against the oracle on random streams."""
import ctypes as C

import numpy as np
import pytest

from conftest import complex_ramp, ramp, same_bits

pytestmark = pytest.mark.gpu

VP, SZ = C.c_void_p, C.c_size_t


class Block:
    def __init__(self, lib, prefix, create_args, in_dtype, out_dtype):
        self.lib, self.prefix, self.in_dtype, self.out_dtype = lib, prefix, in_dtype, out_dtype
        self.h = VP()
        code = getattr(lib, prefix + "_create")(*create_args, C.byref(self.h))
        if code != 0:
            raise RuntimeError("%s_create failed with %d" % (prefix, code))
        self.proc = getattr(lib, prefix + "_process")
        self.proc.restype = None
        self.proc.argtypes = [VP, SZ, C.POINTER(VP), C.POINTER(SZ), VP]
        getattr(lib, prefix + "_destroy").argtypes = [VP]

    def process(self, x):
        x = np.ascontiguousarray(x, dtype=self.in_dtype).copy()
        out, n = VP(), SZ()
        self.proc(x.ctypes.data_as(VP), len(x), C.byref(out), C.byref(n), self.h)
        if n.value == 0 or not out.value:
            return np.zeros(0, self.out_dtype)
        size = n.value * np.dtype(self.out_dtype).itemsize
        return np.frombuffer((C.c_char * size).from_address(out.value), dtype=self.out_dtype).copy()

    def run(self, x, chunk):
        return np.concatenate([self.process(x[o:o + chunk]) for o in range(0, len(x), chunk)])

    def close(self):
        getattr(self.lib, self.prefix + "_destroy")(self.h)


def make_lpf(lib, dec, fs, cutoff, tw, max_len, cplx):
    lib.lpf_create.argtypes = [C.c_uint8, C.c_uint64, C.c_uint64, C.c_uint32, SZ, SZ, C.POINTER(VP)]
    dt = np.complex64 if cplx else np.float32
    return Block(lib, "lpf", (dec, fs, cutoff, tw, max_len, 8 if cplx else 4), dt, dt)


def make_quad(lib, gain, max_len):
    lib.quadrature_demod_create.argtypes = [C.c_float, C.c_uint32, C.POINTER(VP)]
    return Block(lib, "quadrature_demod", (gain, max_len), np.complex64, np.float32)


def make_dc(lib, length):
    lib.dc_blocker_create.argtypes = [C.c_int, C.POINTER(VP)]
    return Block(lib, "dc_blocker", (length,), np.float32, np.float32)


def make_clock(lib, omega, gain_omega, mu, gain_mu, lim, max_len):
    lib.clock_mm_create.argtypes = [C.c_float] * 5 + [SZ, C.POINTER(VP)]
    return Block(lib, "clock_mm", (omega, gain_omega, mu, gain_mu, lim, max_len), np.float32, np.float32)


def noise(n, seed, cplx=False):
    rng = np.random.default_rng(seed)
    if cplx:
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    return rng.standard_normal(n).astype(np.float32)


def test_lpf_reference_kats(sdrm, kats):
    """reference test/test_lpf.c: complex dec 1, real dec 2, small and oversize buffers"""
    f = make_lpf(sdrm.lib, 1, 48000, 4800, 2000, 2000, True)
    x = complex_ramp(500)
    for part, key in ((x[:250], "expected"), (x[250:], "expected2")):
        e = kats["test_lpf.c:test_complex:" + key].view(np.complex64)
        y = f.process(part)
        assert len(y) == len(e) and np.abs(y - e).max() < 1e-2
    f.close()
    f = make_lpf(sdrm.lib, 2, 48000, 4800, 2000, 2000, False)
    x = ramp(1000)
    for part, key in ((x[:500], "expected"), (x[500:], "expected2")):
        e = kats["test_lpf.c:test_normal:" + key]
        y = f.process(part)
        assert len(y) == len(e) and np.abs(y - e).max() < 1e-3
    f.close()
    f = make_lpf(sdrm.lib, 2, 48000, 4800, 2000, 2000, True)
    x = complex_ramp(500)
    assert len(f.process(x[:0])) == 0
    assert len(f.process(x[0:1])) == 1
    assert len(f.process(x[1:2])) == 0
    y = f.process(x[2:3])
    assert len(y) == 1 and abs(-0.005327 - y[0].real) < 1e-3 and abs(-0.007783 - y[0].imag) < 1e-3
    f.close()
    f = make_lpf(sdrm.lib, 2, 48000, 4800, 2000, 100, True)
    assert len(f.process(complex_ramp(101))) == 0  # test_big_buffer
    f.close()


@pytest.mark.parametrize("dec,cplx,chunk", [(1, True, 777), (2, False, 1000), (2, True, 333), (3, False, 1001), (5, True, 64), (1, False, 4096)])
def test_lpf_bit_exact(sdrm, port, dec, cplx, chunk):
    x = noise(20000, 1, cplx)
    f = make_lpf(sdrm.lib, dec, 48000, 4800, 2000, 4096, cplx)
    y = f.run(x, chunk)
    f.close()
    assert same_bits(y, port.Fir(port.low_pass_taps(1.0, 48000, 4800, 2000), dec, cplx).run(x, chunk))


def make_fir(lib, dec, taps, max_len, cplx):
    """fir_filter_create takes ownership of a malloc'ed taps array"""
    libc = C.CDLL(None)
    libc.malloc.restype = VP
    libc.malloc.argtypes = [SZ]
    taps = np.ascontiguousarray(taps, np.float32)
    mem = libc.malloc(taps.nbytes)
    C.memmove(mem, taps.ctypes.data, taps.nbytes)
    lib.fir_filter_create.argtypes = [C.c_uint8, VP, SZ, SZ, SZ, C.POINTER(VP)]
    dt = np.complex64 if cplx else np.float32
    return Block(lib, "fir_filter", (dec, mem, len(taps), max_len, 8 if cplx else 4), dt, dt)


@pytest.mark.parametrize("dec,cplx,ntaps,chunk", [(1, True, 8, 500), (2, False, 33, 999), (4, True, 129, 1024), (1, False, 1, 100),
                                                  (3, False, 600, 2048)])
def test_fir_filter_bit_exact(sdrm, port, dec, cplx, ntaps, chunk):
    """fir_filter handle with arbitrary taps (reference src/dsp/fir_filter.c), chunked, against the oracle"""
    taps = noise(ntaps, 77)
    x = noise(12000, 3, cplx)
    f = make_fir(sdrm.lib, dec, taps, 2048, cplx)
    y = f.run(x, chunk)
    assert len(f.process(x[:4000])) == 0  # over max: NULL / 0, state untouched
    f.close()
    assert same_bits(y, port.Fir(taps, dec, cplx).run(x, chunk))


def test_fir_filter_float_single(sdrm):
    """fir_filter_process_float_single: one sequential dot product with the reversed taps (fir_filter.c:116-121)"""
    taps = noise(8, 5)
    f = make_fir(sdrm.lib, 1, taps, 64, False)
    sdrm.lib.fir_filter_process_float_single.restype = C.c_float
    sdrm.lib.fir_filter_process_float_single.argtypes = [VP, VP]
    x = noise(40, 6)
    for off in range(0, 12):
        got = sdrm.lib.fir_filter_process_float_single(x[off:].ctypes.data_as(VP), f.h)
        acc = np.float32(0)
        for j in range(8):
            acc = np.float32(acc + np.float32(x[off + j] * taps[7 - j]))
        assert np.float32(got) == acc
    f.close()


@pytest.mark.parametrize("n_ch,dec,cplx,chunk", [(5, 1, True, 777), (8, 2, True, 1000), (3, 5, True, 4096), (7, 2, False, 999),
                                                 (4, 3, False, 4096), (1, 1, False, 50)])
def test_lpf_batch_bit_exact(sdrm, port, n_ch, dec, cplx, chunk):
    """sdrm_lpf_batch: N streams through one filter design, each equal to its own lpf handle of the reference"""
    x = np.stack([noise(9000, 100 + c, cplx) for c in range(n_ch)])
    b = sdrm.LpfBatch(n_ch, dec, 48000, 4800, 2000, 4096, cplx)
    parts = [b.process(x[:, o:o + chunk]) for o in range(0, x.shape[1], chunk)]
    assert b.process(x[:, :0]).shape[1] == 0
    with pytest.raises(sdrm.SdrmError):
        b.process(x[:, :5000])
    b.close()
    y = np.concatenate(parts, axis=1)
    taps = port.low_pass_taps(1.0, 48000, 4800, 2000)
    for c in range(n_ch):
        assert same_bits(y[c], port.Fir(taps, dec, cplx).run(x[c], chunk))


def test_quadrature_demod(sdrm, port, kats):
    q = make_quad(sdrm.lib, 25.4, 2000)
    x = complex_ramp(200)
    assert np.abs(q.process(x[:2]) - kats["test_quadrature_demod.c:test_normal:expected"]).max() < 1e-3
    assert np.abs(q.process(x[2:]) - kats["test_quadrature_demod.c:test_normal:expected2"]).max() < 1e-3
    q.close()
    x = noise(30000, 2, True)
    x[100] = 0
    x[101] = 0
    q = make_quad(sdrm.lib, 6.1, 4096)
    y = q.run(x, 999)
    q.close()
    assert same_bits(y, port.QuadDemod(6.1).run(x, 999))


def test_dc_blocker(sdrm, port, kats):
    d = make_dc(sdrm.lib, 32)
    assert np.abs(d.process(ramp(200)) - kats["test_dc_blocker.c:test_normal:expected"]).max() < 1e-3
    d.close()
    x = noise(50000, 3)
    d = make_dc(sdrm.lib, 160)
    y = d.run(x, 4096)
    d.close()
    assert same_bits(y, port.DcBlocker(160).run(x, 4096))


def test_clock_recovery(sdrm, port, kats):
    args = (2.0, np.float32(0.25) * np.float32(0.175) * np.float32(0.175), 0.005, 0.175, 0.005)
    x = ramp(100)
    c = make_clock(sdrm.lib, *args, 100)
    assert len(c.process(x[:0])) == 0 and len(c.process(x[:4])) == 0 and len(c.process(x[4:7])) == 0
    y = c.process(x[7:8])
    assert len(y) == 1 and abs(3.007791 - y[0]) < 1e-3
    c.close()
    c = make_clock(sdrm.lib, *args, 100)
    assert np.abs(c.process(x[:42]) - kats["test_clock_recovery_mm.c:test_normal:expected"]).max() < 1e-3
    assert np.abs(c.process(x[42:78]) - kats["test_clock_recovery_mm.c:test_normal:expected2"]).max() < 1e-3
    c.close()
    c = make_clock(sdrm.lib, *args, 10)
    assert len(c.process(ramp(11))) == 0  # test_big_buffers
    c.close()
    sig = np.sin(np.arange(60000) * 2 * np.pi / 10.03).astype(np.float32) + 0.1 * noise(60000, 5)
    cargs = (10.0, np.float32(10.0) * np.float32(np.pi) / 100, 0.5, 0.0625, 0.01)
    for chunk in (2048, 37, 5000):
        c = make_clock(sdrm.lib, *cargs, 5000)
        y = c.run(sig, chunk)
        c.close()
        assert same_bits(y, port.ClockMm(*cargs, 5000).run(sig, chunk))
