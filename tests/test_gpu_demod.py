"""GPU parity tests for the demod hot path, all through the C ABI (libsdrmodem_b200.so via ctypes).

Bar: exact mode is BIT-IDENTICAL to the oracle (int8 hard symbols and float soft symbols) for identical call sizes.
The oracle is the reference's own sources built in place (oracle/_ref) when that prebuilt library travelled with the
repo, and always our C restatement (oracle/sdrm_oracle.c), which test_oracle.py pins to the reference.
"""
import ctypes as C

import numpy as np
import pytest

import workloads
from conftest import FSK_GOLDENS, complex_ramp, golden_array, ramp, same_bits

pytestmark = pytest.mark.gpu


def oracle_chain(port, args, iq, chunk):
    """The expected result: our C restatement (oracle/sdrm_oracle.c) and, wherever the reference build travelled with the
    repo (oracle/_ref), the reference's own sources on the same input, which must agree with it bit for bit: every
    comparison of this file is therefore a comparison with the reference itself, on the machine the GPU tests run on."""
    hard, soft = port.FskDemod(*args, chunk).run(iq, chunk)
    from oracle import ref
    if ref.available():
        r = ref.fsk_chain(*args, iq, chunk)
        assert same_bits(hard, r["hard"]) and same_bits(soft, r["soft"]), "oracle port differs from the reference build"
    return hard, soft


def gpu_chain(sdrm, args, iq_channels, chunk, **kw):
    b = sdrm.FskDemodBatch(iq_channels.shape[0], *args, chunk, soft=True, **kw)
    try:
        hard, soft = b.run_stream(iq_channels, chunk)
        assert b.error_flags() == 0
    finally:
        b.close()
    return hard, soft


# ---- the reference's own goldens ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(FSK_GOLDENS))
def test_goldens_bit_exact_vs_oracle_and_within_reference_tolerance(sdrm, port, name):
    inp, exp, args = FSK_GOLDENS[name]
    iq = golden_array(inp, np.complex64)
    expected = golden_array(exp, np.int8)
    hard, soft = gpu_chain(sdrm, args, iq[None, :], 4096)
    o_hard, o_soft = oracle_chain(port, args, iq, 4096)
    assert same_bits(hard[0], o_hard)
    assert same_bits(soft[0], o_soft)
    assert len(hard[0]) == len(expected)
    assert np.abs(hard[0].astype(int) - expected.astype(int)).max() <= 2  # reference test/test_fsk_demod.c:47


@pytest.mark.parametrize("name", ["lucky7", "nan"])
def test_goldens_vs_reference_build(sdrm, ref, name):
    inp, _, args = FSK_GOLDENS[name]
    iq = golden_array(inp, np.complex64)
    hard, soft = gpu_chain(sdrm, args, iq[None, :], 4096)
    r = ref.fsk_chain(*args, iq, 4096)
    assert same_bits(hard[0], r["hard"])
    assert same_bits(soft[0], r["soft"])


@pytest.mark.parametrize("chunk", [4096, 4001, 1, 7, 50000, 96000])
def test_call_size_is_part_of_the_semantics(sdrm, port, chunk):
    """Ragged / odd / tiny / whole-file calls: same call sizes on both sides, bit-identical results."""
    _, _, args = FSK_GOLDENS["lucky7"]
    iq = golden_array("lucky7.expected.cf32", np.complex64)
    if chunk < 100:
        iq = iq[:3000]
    hard, soft = gpu_chain(sdrm, args, iq[None, :], chunk)
    o_hard, o_soft = oracle_chain(port, args, iq, chunk)
    assert same_bits(hard[0], o_hard) and same_bits(soft[0], o_soft)


def test_mixed_call_sizes_and_empty_calls(sdrm, port):
    _, _, args = FSK_GOLDENS["lucky7"]
    iq = golden_array("lucky7.expected.cf32", np.complex64)[:40000]
    sizes = [0, 1, 0, 5, 4096, 3, 1001, 0, 8192, 4095]
    b = sdrm.FskDemodBatch(2, *args, 8192, soft=True)
    o = port.FskDemod(*args, 8192)
    off = 0
    for n in sizes * 2:
        part = iq[off:off + n]
        off += n
        hard, lens, soft = b.process(np.stack([part, part]))
        oh, os_ = o.process(part)
        for c in range(2):
            assert same_bits(hard[c, :lens[c]], oh) and same_bits(soft[c, :lens[c]], os_)
    assert b.error_flags() == 0
    b.close()


# ---- synthetic workloads of BASELINE.json ---------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [workloads.C2_PARITY, workloads.PERF_SHAPE], ids=lambda s: s.name)
def test_synthetic_channels_bit_exact(sdrm, port, shape):
    n_ch, n = 6, 40000
    iq = workloads.gfsk_channels(n_ch, n, shape, seed=11).numpy()
    hard, soft = gpu_chain(sdrm, shape.create_args, iq, shape.chunk)
    for c in range(n_ch):
        oh, os_ = oracle_chain(port, shape.create_args, iq[c], shape.chunk)
        assert same_bits(hard[c], oh), "channel %d hard symbols differ" % c
        assert same_bits(soft[c], os_), "channel %d soft symbols differ" % c


def test_odd_channel_count_and_no_dc(sdrm, port):
    shape = workloads.C2_PARITY
    args = shape.create_args[:5] + (False,)
    iq = workloads.gfsk_channels(5, 20000, shape, seed=12).numpy()
    hard, soft = gpu_chain(sdrm, args, iq, 4096)
    for c in range(5):
        oh, os_ = oracle_chain(port, args, iq[c], 4096)
        assert same_bits(hard[c], oh) and same_bits(soft[c], os_)


@pytest.mark.parametrize("dec", [1, 3, 4])
def test_other_decimations(sdrm, port, dec):
    args = (48000, 2400, 5000, dec, 2000, True)
    iq = golden_array("lucky7.expected.cf32", np.complex64)[:50000]
    hard, soft = gpu_chain(sdrm, args, iq[None, :], 4096)
    oh, os_ = oracle_chain(port, args, iq, 4096)
    assert same_bits(hard[0], oh) and same_bits(soft[0], os_)


def test_long_filter_multi_block_taps(sdrm, port):
    """T1 > 528 taps exercises the tap-blocked, double-buffered TMA path (fs 480k: T1 = 1179, T2 = 577)."""
    args = (480000, 9600, 5000, 2, 2000, True)
    shape = workloads.DemodShape("x", 480000, 9600, 5000, 2, 2000, True, 8192)
    iq = workloads.gfsk_channels(2, 30000, shape, seed=13).numpy()
    hard, soft = gpu_chain(sdrm, args, iq, 8192)
    for c in range(2):
        oh, os_ = oracle_chain(port, args, iq[c], 8192)
        assert same_bits(hard[c], oh) and same_bits(soft[c], os_)


def test_non_finite_samples(sdrm, port):
    """NaN / Inf in the input must poison exactly the symbols they poison in the reference (clock NaN guard,
    reference src/dsp/clock_recovery_mm.c:107-113; leading zero taps of the aligned dot product)."""
    shape = workloads.PERF_SHAPE
    iq = workloads.gfsk_channels(1, 30000, shape, seed=14).numpy()[0].copy()
    iq[5000] = np.nan
    iq[12000] = np.inf + 0j
    iq[12001] = -np.inf * 1j
    iq[20000:20003] = np.nan + 1j * np.nan
    hard, soft = gpu_chain(sdrm, shape.create_args, iq[None, :], 4096)
    oh, os_ = oracle_chain(port, shape.create_args, iq, 4096)
    assert same_bits(hard[0], oh) and same_bits(soft[0], os_)


def test_batch_position_invariance_at_scale(sdrm, port):
    """Full-size property: 256 channels made of 4 distinct signals; every copy must equal the oracle's answer for its
    signal wherever it sits in the batch (pairing, tiles and warps differ by position)."""
    shape = workloads.C2_PARITY
    base = workloads.gfsk_channels(4, 3 * 16384, shape, seed=15).numpy()
    order = np.random.default_rng(0).integers(0, 4, 256)
    iq = base[order]
    hard, soft = gpu_chain(sdrm, shape.create_args, iq, 16384)
    want = [oracle_chain(port, shape.create_args, base[k], 16384) for k in range(4)]
    for c in range(256):
        assert same_bits(hard[c], want[order[c]][0]) and same_bits(soft[c], want[order[c]][1]), c


def test_device_resident_pipeline_matches_host_path(sdrm, port):
    import torch
    shape = workloads.C2_PARITY
    n_ch, chunk, calls = 8, 8192, 5
    iq = workloads.gfsk_channels(n_ch, chunk * calls, shape, seed=16)
    d_iq = iq.cuda()
    b = sdrm.FskDemodBatch(n_ch, *shape.create_args, chunk, max_symbols_per_call=1024, soft=True)
    got = [[] for _ in range(n_ch)]
    pending = 0
    bufs = [d_iq[:, k * chunk:(k + 1) * chunk].contiguous() for k in range(calls)]
    for k in range(calls):
        b.process_device(bufs[k].data_ptr(), chunk, chunk)
        pending += 1
        if pending == sdrm.MAX_IN_FLIGHT:
            hard, lens, soft = b.fetch()
            pending -= 1
            for c in range(n_ch):
                got[c].append(soft[c, :lens[c]].copy())
    while pending:
        hard, lens, soft = b.fetch()
        pending -= 1
        for c in range(n_ch):
            got[c].append(soft[c, :lens[c]].copy())
    assert b.error_flags() == 0
    b.close()
    for c in range(n_ch):
        _, os_ = oracle_chain(port, shape.create_args, iq[c].numpy(), chunk)
        assert same_bits(np.concatenate(got[c]), os_)


# ---- fast (FMA) mode: checked against the FMA-order reference build, reported against the strict one ----------------------
def test_fast_mode_equals_fma_order_reference(sdrm, ref):
    from oracle import ref as ref_module
    if not ref_module.available(fma=True):
        pytest.skip("oracle/_ref/libsdrmodem_ref_fma.so not built")
    _, _, args = FSK_GOLDENS["lucky7"]
    iq = golden_array("lucky7.expected.cf32", np.complex64)
    hard, soft = gpu_chain(sdrm, args, iq[None, :], 4096, fast=True)
    r = ref_module.fsk_chain(*args, iq, 4096, fma=True)
    assert same_bits(hard[0], r["hard"]) and same_bits(soft[0], r["soft"])


def test_fast_mode_epsilon_rule_vs_strict_oracle(sdrm, port):
    """north_star rule: identical hard bits except where |soft| < eps; soft within 1e-4 relative. Fast mode does NOT
    guarantee it (feedback loops amplify the fused roundings); this test only bounds how far it strays on lucky7."""
    _, _, args = FSK_GOLDENS["lucky7"]
    iq = golden_array("lucky7.expected.cf32", np.complex64)
    hard, soft = gpu_chain(sdrm, args, iq[None, :], 4096, fast=True)
    oh, os_ = oracle_chain(port, args, iq, 4096)
    assert len(hard[0]) == len(oh)
    eps = 5e-3
    strong = np.abs(os_) >= eps
    assert np.array_equal((soft[0] < 0)[strong], (os_ < 0)[strong])
    assert np.abs(hard[0].astype(int) - oh.astype(int)).max() <= 2


# ---- the reference's single-channel entry points -----------------------------------------------------------------------------
def test_dropin_fsk_demod_handle(sdrm, port):
    _, _, args = FSK_GOLDENS["lucky7"]
    iq = golden_array("lucky7.expected.cf32", np.complex64)
    d = sdrm.FskDemod(*args, 4096)
    out = d.run(iq, 4096)
    assert len(d.process(np.zeros(4097, np.complex64))) == 0  # oversize call: NULL / 0 as the reference
    d.close()
    oh, _ = oracle_chain(port, args, iq, 4096)
    assert same_bits(out, oh)


def test_dropin_create_errors(sdrm):
    with pytest.raises(sdrm.SdrmError):
        sdrm.FskDemod(48000, 48000, 5000, 2, 2000, True, 4096)  # cutoff above fs/2 (reference test_dsp_worker.c:66-78)
    with pytest.raises(sdrm.SdrmError):
        sdrm.FskDemod(0, 4800, 5000, 2, 2000, True, 4096)


@pytest.mark.parametrize("steps", [1, 2, 0])
@pytest.mark.parametrize("length", [320, 160, 32, 3200, 17, 1000, 4096, 33333])
def test_tail_division_by_length_is_ieee(sdrm, length, steps):
    """The fused tail divides running sums by the dc blocker length without a generic division; it must round like one
    (reference src/dsp/dc_blocker.c:63 is a true float division). steps = Markstein corrections (the host proves per
    length, over all 2^23 mantissas, that one is enough before the kernel is allowed to use one); 0 = __fdiv_rn."""
    lib = sdrm.lib
    lib.sdrm_cu_selftest_div.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_ulonglong)]
    assert lib.sdrm_division_steps(length) == 1
    bad = C.c_ulonglong(123)
    assert lib.sdrm_cu_selftest_div(length, steps, 12345 + length, 1184, 4096, C.byref(bad)) == 0
    assert bad.value == 0


# ---- random parameter sets ------------------------------------------------------------------------------------------------
def random_demod_config(seed):
    rng = np.random.default_rng(5000 + seed)
    dec = int(rng.choice([1, 2, 2, 3, 4, 5, 6, 8]))
    sps_out = float(rng.uniform(2.0, 24.0))
    baud = int(rng.choice([1000, 1200, 2400, 4800, 7777, 9600, 19200]))
    fs = int(round(baud * sps_out * dec)) + int(rng.integers(0, 7))
    deviation = int(baud * rng.uniform(0.25, 1.5)) * (1 if rng.random() < 0.8 else -1)
    tw = max(50, int(baud * rng.uniform(0.05, 0.5)))
    use_dc = bool(rng.random() < 0.7)
    return rng, (fs, baud, deviation, dec, tw, use_dc)


@pytest.mark.parametrize("seed", range(40))
def test_random_parameter_sets_bit_exact(sdrm, port, seed):
    """fsk_demod_create's whole parameter space, sampled: decimation 1..8, 2..24 samples per symbol (non-integer), either sign of
    the deviation, filters from a few dozen to several thousand taps (one tap block, several, and the multi-launch path), dc
    blocker on and off; even seeds run a fixed call size (compared with the restatement AND the reference build), odd seeds a
    ragged sequence of call sizes with empty and tiny calls in it. A parameter set the reference rejects must be rejected."""
    rng, args = random_demod_config(seed)
    fs, baud, deviation, dec, tw, use_dc = args
    n = 30000
    shape = workloads.DemodShape("random", fs, baud, abs(deviation), dec, tw, use_dc, 4096)
    iq = workloads.gfsk_channels(3, n, shape, seed=300 + seed).numpy()
    try:
        probe = port.FskDemod(*args, 4096)
    except Exception:
        with pytest.raises(sdrm.SdrmError):
            sdrm.FskDemodBatch(3, *args, 4096)
        return
    del probe
    if seed % 2 == 0:
        chunk = int(rng.choice([1024, 4096, 5000, 30000]))
        hard, soft = gpu_chain(sdrm, args, iq, chunk)
        for c in range(3):
            oh, os_ = oracle_chain(port, args, iq[c], chunk)
            assert len(oh) > 0.8 * n / (fs / baud) - 20
            assert same_bits(hard[c], oh) and same_bits(soft[c], os_), (args, chunk, c)
        return
    sizes = []
    while sum(sizes) < n:
        sizes.append(int(rng.choice([0, 1, 3, 17, 500, 2048, 4095, 6000])))
    sizes[-1] -= sum(sizes) - n
    b = sdrm.FskDemodBatch(3, *args, 6000, soft=True)
    oracles = [port.FskDemod(*args, 6000) for _ in range(3)]
    off = 0
    try:
        for size in sizes:
            part = iq[:, off:off + size]
            off += size
            hard, lens, soft = b.process(np.ascontiguousarray(part))
            for c in range(3):
                oh, os_ = oracles[c].process(part[c])
                assert same_bits(hard[c, :lens[c]], oh) and same_bits(soft[c, :lens[c]], os_), (args, sizes, c)
        assert b.error_flags() == 0
    finally:
        b.close()


@pytest.mark.parametrize("use_dc", [True, False])
def test_samples_per_symbol_beyond_the_fused_tail(sdrm, port, use_dc):
    """100 baud at 96 ksps without decimation: 960 samples per symbol, a dc blocker of 30720 rows. The fused tail keeps one symbol
    step of every channel in shared memory and stops at ~855 samples per symbol; beyond that the batch runs the dc blocker and the
    clock loop as the two plain kernels of tail.cu on the grouped lpf2 ring (the reference accepts any value,
    src/dsp/fsk_demod.c:53). 34 channels = two channel groups; calls of 50000 samples and a ragged tail."""
    args = (96000, 100, 5000, 1, 2000, use_dc)
    shape = workloads.DemodShape("slow", 96000, 100, 5000, 1, 2000, use_dc, 50000)
    n_ch, n = 34, 230000
    iq = workloads.gfsk_channels(n_ch, n, shape, seed=77).numpy()
    b = sdrm.FskDemodBatch(n_ch, *args, 50000, soft=True)
    checked = (0, 1, 31, 32, 33)
    oracles = {c: port.FskDemod(*args, 50000) for c in checked}
    total = 0
    try:
        for lo, hi in ((0, 50000), (50000, 100000), (100000, 100003), (100003, 150003), (150003, 150003), (150003, 200003), (200003, n)):
            hard, lens, soft = b.process(np.ascontiguousarray(iq[:, lo:hi]))
            for c in checked:
                oh, os_ = oracles[c].process(iq[c, lo:hi])
                assert same_bits(hard[c, :lens[c]], oh) and same_bits(soft[c, :lens[c]], os_), (use_dc, lo, c)
            total += int(lens[0])
        assert b.error_flags() == 0
    finally:
        b.close()
    assert 200 < total < 260
