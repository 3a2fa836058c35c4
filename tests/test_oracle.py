"""CPU-only: pins the oracle. (1) our C restatement (oracle/sdrm_oracle.c) against the reference's golden files and
inline known-answer arrays; (2) the restatement bit-for-bit against the reference's own sources compiled in place
(oracle/_ref), when that build is present."""
import numpy as np
import pytest

from conftest import FSK_GOLDENS, LUCKY7_TLE, complex_ramp, golden_array, ramp, same_bits


# ---- end-to-end goldens (reference test/test_fsk_demod.c:22-81, tolerance abs(diff) <= 2 at :47) -------------------
@pytest.mark.parametrize("name", sorted(FSK_GOLDENS))
def test_port_fsk_demod_goldens(port, name):
    inp, exp, args = FSK_GOLDENS[name]
    iq = golden_array(inp, np.complex64)
    expected = golden_array(exp, np.int8)
    hard, _ = port.FskDemod(*args, 4096).run(iq, 4096)  # 4096-sample calls are part of the contract (:14-20)
    assert len(hard) == len(expected)
    assert np.abs(hard.astype(int) - expected.astype(int)).max() <= 2


@pytest.mark.parametrize("name", sorted(FSK_GOLDENS))
def test_ref_fsk_demod_goldens(ref, name):
    inp, exp, args = FSK_GOLDENS[name]
    iq = golden_array(inp, np.complex64)
    expected = golden_array(exp, np.int8)
    hard = ref.fsk_demod_run(*args, iq, 4096)
    assert len(hard) == len(expected)
    assert np.abs(hard.astype(int) - expected.astype(int)).max() <= 2


@pytest.mark.parametrize("name", sorted(FSK_GOLDENS))
@pytest.mark.parametrize("chunk", [4096, 1000, 4097])
def test_port_equals_ref_fsk_chain(port, ref, name, chunk):
    inp, _, args = FSK_GOLDENS[name]
    iq = golden_array(inp, np.complex64)
    r = ref.fsk_chain(*args, iq, chunk)
    hard, soft = port.FskDemod(*args, chunk).run(iq, chunk)
    assert same_bits(hard, r["hard"])
    assert same_bits(soft, r["soft"])


def test_ref_doppler_golden(ref):
    """reference test/test_doppler.c:39-63: lucky7.cf32 -> lucky7.expected.cf32, 2000-sample calls, tolerance 1e-2"""
    x = golden_array("lucky7.cf32", np.complex64)
    e = golden_array("lucky7.expected.cf32", np.complex64)
    d = ref.doppler(np.float32(53.72), np.float32(47.57), 0.0, 48000, 437525000, 0, 1583840449, 2000, LUCKY7_TLE)
    y = d.run(x, 2000)
    assert len(y) == len(e)
    assert np.abs(y - e).max() < 1e-2


# ---- inline known-answer arrays of the reference unit tests -----------------------------------------------------------
def test_kat_lpf_taps(port, kats):
    """test/test_lpf_taps.c:28-40"""
    taps = port.low_pass_taps(1.0, 8000, 1750, 500)
    e = kats["test_lpf_taps.c:test_lowpassTaps:expected_taps"]
    assert len(taps) == 39
    assert np.array_equal((e * 10000).astype(np.int32), (taps * 10000).astype(np.int32))
    for bad in ((0, 1750, 500), (8000, 5000, 500), (8000, 1750, 0)):  # test_bounds1..3
        with pytest.raises(ValueError):
            port.low_pass_taps(1.0, *bad)


def test_kat_lpf_complex(port, kats):
    """test/test_lpf.c:49-73: lpf_create(1, 48000, 4800, 2000), complex ramp, two calls of 250"""
    f = port.Fir(port.low_pass_taps(1.0, 48000, 4800, 2000), 1, True)
    x = complex_ramp(500)
    for part, key in ((x[:250], "expected"), (x[250:], "expected2")):
        y = f.process(part)
        e = kats["test_lpf.c:test_complex:" + key].view(np.complex64)
        assert len(y) == len(e)
        assert np.abs(y - e).max() < 1e-2


def test_kat_lpf_real_decim2(port, kats):
    """test/test_lpf.c test_normal: lpf_create(2, 48000, 4800, 2000, real), ramp, two calls of 500"""
    f = port.Fir(port.low_pass_taps(1.0, 48000, 4800, 2000), 2, False)
    x = ramp(1000)
    for part, key in ((x[:500], "expected"), (x[500:], "expected2")):
        y = f.process(part)
        e = kats["test_lpf.c:test_normal:" + key]
        assert len(y) == len(e)
        assert np.abs(y - e).max() < 1e-3


def test_kat_lpf_small_buffers(port):
    """test/test_lpf.c:25-45: 0, 1, 1, 1 samples through a decimate-by-2 complex filter"""
    f = port.Fir(port.low_pass_taps(1.0, 48000, 4800, 2000), 2, True)
    x = complex_ramp(500)
    assert len(f.process(x[:0])) == 0
    assert len(f.process(x[0:1])) == 1
    assert len(f.process(x[1:2])) == 0
    y = f.process(x[2:3])
    assert len(y) == 1
    assert abs(-0.005327 - y[0].real) < 1e-3 and abs(-0.007783 - y[0].imag) < 1e-3


def test_kat_quadrature_demod(port, kats):
    """test/test_quadrature_demod.c:10-27"""
    q = port.QuadDemod(25.4)
    x = complex_ramp(200)
    assert np.abs(q.process(x[:2]) - kats["test_quadrature_demod.c:test_normal:expected"]).max() < 1e-3
    assert np.abs(q.process(x[2:]) - kats["test_quadrature_demod.c:test_normal:expected2"]).max() < 1e-3


def test_kat_dc_blocker(port, kats):
    """test/test_dc_blocker.c:10-21"""
    y = port.DcBlocker(32).process(ramp(200))
    assert np.abs(y - kats["test_dc_blocker.c:test_normal:expected"]).max() < 1e-3


def test_kat_clock_recovery(port, kats):
    """test/test_clock_recovery_mm.c:11-66"""
    make = lambda n: port.ClockMm(2.0, np.float32(0.25) * np.float32(0.175) * np.float32(0.175), 0.005, 0.175, 0.005, n)
    x = ramp(100)
    c = make(100)
    assert len(c.process(x[:0])) == 0
    assert len(c.process(x[:4])) == 0
    assert len(c.process(x[4:7])) == 0
    y = c.process(x[7:8])
    assert len(y) == 1 and abs(3.007791 - y[0]) < 1e-3
    c = make(100)
    assert np.abs(c.process(x[:42]) - kats["test_clock_recovery_mm.c:test_normal:expected"]).max() < 1e-3
    assert np.abs(c.process(x[42:78]) - kats["test_clock_recovery_mm.c:test_normal:expected2"]).max() < 1e-3
    assert len(make(10).process(ramp(11))) == 0  # test_big_buffers


def test_kat_sig_source(port, kats):
    """test/test_sig_source.c:8-18"""
    y = port.SigSource(1.0, 4).generate(1, 4)
    assert np.abs(y - kats["test_sig_source.c:test_success:buffer"].view(np.complex64)).max() < 1e-2


def test_kat_gaussian_taps(port, kats):
    """test/test_gaussian_taps.c:7-15"""
    taps = port.gaussian_taps(1.5, 2 * float(np.float32(48000.0) / np.float32(9600)), 0.5, 12)
    e = kats["test_gaussian_taps.c:test_normal:expected_taps"]
    assert np.array_equal((e * 10000).astype(np.int32), (taps * 10000).astype(np.int32))


def test_kat_convolve(port, kats):
    """test/test_gfsk_mod.c test_convolve"""
    y = port.convolve([0, 1, 0.5], [1, 2, 3])
    assert np.abs(y - kats["test_gfsk_mod.c:test_convolve:expected"]).max() < 1e-3


def test_kat_frequency_modulator(port, kats):
    """test/test_frequency_modulator.c:24-53: sensitivity 1.2, ramp input, first 100 samples"""
    y = port.FreqMod(1.2).process(ramp(100))
    e = kats["test_frequency_modulator.c:test_normal:expected"][:200].view(np.complex64)
    assert np.abs(y - e).max() < 1e-2


def test_kat_gfsk_mod(port, kats):
    """test/test_gfsk_mod.c test_normal: 19200/9600, deviation 5000, BT 0.5, bytes 0..9"""
    sps = np.float32(19200) / np.float32(9600)
    m = port.GfskMod(sps, np.float32(2 * np.pi * 5000 / np.float32(19200)), 0.5, 1000)
    y = m.process(np.arange(10, dtype=np.uint8))
    e = kats["test_gfsk_mod.c:test_normal:expected"].view(np.complex64)
    assert len(y) == 160
    assert np.abs(y - e).max() < 1e-2
    assert len(port.GfskMod(sps, 0.1, 0.5, 10).process(np.arange(11, dtype=np.uint8))) == 0  # test_exceeded_input


# ---- restatement == reference build, stage by stage, bit for bit --------------------------------------------------------
def _noise(n, seed, complex_=False):
    rng = np.random.default_rng(seed)
    if complex_:
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    return rng.standard_normal(n).astype(np.float32)


def test_port_equals_ref_taps(port, ref):
    for args in ((192000, 9800, 980), (192000, 4800, 2000), (48000, 7400, 740), (2400000, 6200, 620), (8000, 1750, 500)):
        assert same_bits(port.low_pass_taps(1.0, *args), ref.lpf_taps(1.0, *args))
    for sps, n in ((2.0, 8), (20.0, 80), (10.4166, 12)):
        assert same_bits(port.gaussian_taps(1.0, sps, 0.5, n), ref.gaussian_taps(1.0, sps, 0.5, n))


@pytest.mark.parametrize("dec,cplx,chunk", [(1, True, 777), (2, False, 1000), (2, True, 333), (3, False, 1001), (5, True, 64)])
def test_port_equals_ref_lpf(port, ref, dec, cplx, chunk):
    x = _noise(5000, 1, cplx)
    r = ref.lpf(dec, 48000, 4800, 2000, 1100, cplx).run(x, chunk)
    p = port.Fir(port.low_pass_taps(1.0, 48000, 4800, 2000), dec, cplx).run(x, chunk)
    assert same_bits(p, r)


def test_port_equals_ref_atan(port, ref):
    rng = np.random.default_rng(2)
    y = np.concatenate([rng.standard_normal(4000), [0, 0, 1, -1, 0.0039, 1e-30, np.inf, np.nan, 3, -0.0]]).astype(np.float32)
    x = np.concatenate([rng.standard_normal(4000), [0, 1, 0, 0, 1.0, 1e-30, 1, 1, np.inf, -0.0]]).astype(np.float32)
    assert same_bits(port.fast_atan2f(y, x), ref.fast_atan2f(y, x))


def test_port_equals_ref_quad_dc_clock(port, ref):
    x = _noise(6000, 3, True)
    assert same_bits(port.QuadDemod(6.1).run(x, 999), ref.quadrature_demod(6.1, 1000).run(x, 999))
    f = _noise(20000, 4)
    assert same_bits(port.DcBlocker(160).run(f, 4096), ref.dc_blocker(160).run(f, 4096))
    sig = np.sin(np.arange(30000) * 2 * np.pi / 10.03).astype(np.float32) + 0.1 * _noise(30000, 5)
    args = (10.0, np.float32(10.0) * np.float32(np.pi) / 100, 0.5, 0.0625, 0.01)
    for chunk in (2048, 37, 5000):
        assert same_bits(port.ClockMm(*args, 5000).run(sig, chunk), ref.clock_mm(*args, 5000).run(sig, chunk))


def test_port_equals_ref_nco_and_mod(port, ref):
    x = _noise(9000, 6, True)
    p, r = port.SigSource(1.0, 48000), ref.sig_source(1.0, 48000, 3000)
    for k, freq in enumerate((4321, -9000, 0)):
        part = x[3000 * k:3000 * (k + 1)]
        assert same_bits(p.multiply(freq, part), r.multiply(freq, part))
    f = _noise(5000, 7)
    assert same_bits(port.FreqMod(0.8).process(f), ref.frequency_modulator(0.8, 5000).process(f))
    data = np.random.default_rng(8).integers(0, 256, 300, dtype=np.uint8)
    for sps in (2.0, 20.0, 5.0):
        sens = float(np.float32(2 * np.pi * 5000 / (9600 * sps)))
        pm, rm = port.GfskMod(sps, sens, 0.5, 200), ref.gfsk_mod(sps, sens, 0.5, 200)
        for part in (data[:200], data[200:]):
            assert same_bits(pm.process(part), rm.process(part))


def test_sample_format_converts_match_reference_kats(port):
    """orc_convert_16i_32f / orc_convert_32f_16i against the reference's PlutoSDR known answers (test/test_plutosdr.c:149-154,
    192-194): rx int16 i -> i / 2048 printed to 6 decimals, tx x * 32768 -> 0, 16, 32 ..."""
    rx = port.convert_16i_32f(np.arange(50, dtype=np.int16), 2048.0)
    assert np.array_equal(rx, (np.arange(50) / 2048.0).astype(np.float32))
    printed = np.array([float("%.6f" % v) for v in rx], dtype=np.float32)
    tx = port.convert_32f_16i(printed, 32768.0)
    assert np.array_equal(tx, np.arange(50, dtype=np.int16) * 16)
    edge = np.array([1.0, -1.0, 2.0, -2.0, 0.5 / 32768, 1.5 / 32768, 2.5 / 32768, -0.5 / 32768, np.nan], dtype=np.float32)
    assert list(port.convert_32f_16i(edge, 32768.0)) == [32767, -32768, 32767, -32768, 0, 2, 2, 0, 0]


def test_libm_sweep_counts_one_ulp_differences(port):
    """the checker behind tests/test_gpu_sincos_sweep.py: agrees with numpy's double cos / sin on a block of float phases and
    counts exactly the values that were moved by one ulp"""
    from oracle import port as port_module
    first, n = 0x40000000, 1 << 18   # floats from 2.0 upwards
    x = (np.arange(n, dtype=np.uint32) + first).view(np.float32).astype(np.float64)
    got = np.stack([np.cos(x).astype(np.float32), np.sin(x).astype(np.float32)], axis=1)
    clean, _ = port_module.sincos_sweep(first, got, threads=4)
    assert clean <= 2  # numpy's SIMD routines may differ from libm in the last place of a rare value; libm is the reference
    moved = got.copy()
    moved[12345, 1] = np.nextafter(moved[12345, 1], np.float32(2))
    moved[200000, 0] = np.nextafter(moved[200000, 0], np.float32(2))
    bad, where = port_module.sincos_sweep(first, moved, threads=4)
    assert clean + 1 <= bad <= clean + 2 and where is not None
