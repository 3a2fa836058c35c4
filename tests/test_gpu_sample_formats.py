"""GPU: SDR sample formats on the device (int16 IQ <-> cf32), SURVEY §8 f3. Reference: the PlutoSDR plugin's host-side VOLK
conversions (src/sdr/plutosdr.c:83,129); known answers from test/test_plutosdr.c:149-154,192-194."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import same_bits

pytestmark = pytest.mark.gpu

# test/test_plutosdr.c:149-151 (rx: int16 i -> i / 2048) and :192 (tx: x * 32768 rounded)
PLUTO_FLOATS = np.array([0.000000, 0.000488, 0.000977, 0.001465, 0.001953, 0.002441, 0.002930, 0.003418, 0.003906, 0.004395,
                         0.004883, 0.005371, 0.005859, 0.006348, 0.006836, 0.007324, 0.007812, 0.008301, 0.008789, 0.009277,
                         0.009766, 0.010254, 0.010742, 0.011230, 0.011719, 0.012207, 0.012695, 0.013184, 0.013672, 0.014160,
                         0.014648, 0.015137, 0.015625, 0.016113, 0.016602, 0.017090, 0.017578, 0.018066, 0.018555, 0.019043,
                         0.019531, 0.020020, 0.020508, 0.020996, 0.021484, 0.021973, 0.022461, 0.022949, 0.023438, 0.023926],
                        dtype=np.float32)


def to_cf32(sdrm, x16, scalar):
    rows, n = x16.shape[0], x16.shape[1]
    d_in = torch.from_numpy(x16).cuda()
    d_out = torch.empty((rows, n, 2), dtype=torch.float32, device="cuda")
    assert sdrm.lib.sdrm_samples_i16_to_cf32_device(d_in.data_ptr(), n, d_out.data_ptr(), n, scalar, n, rows, None) == 0
    torch.cuda.synchronize()
    return d_out.cpu().numpy()


def to_i16(sdrm, x, scalar):
    rows, n = x.shape[0], x.shape[1]
    d_in = torch.from_numpy(x).cuda()
    d_out = torch.empty((rows, n, 2), dtype=torch.int16, device="cuda")
    assert sdrm.lib.sdrm_samples_cf32_to_i16_device(d_in.data_ptr(), n, d_out.data_ptr(), n, scalar, n, rows, None) == 0
    torch.cuda.synchronize()
    return d_out.cpu().numpy()


def test_reference_known_answers(sdrm):
    rx = to_cf32(sdrm, np.arange(50, dtype=np.int16).reshape(1, 25, 2), 2048.0)
    assert np.abs(rx.ravel() - PLUTO_FLOATS).max() < 1e-6  # the reference prints 6 decimals and compares at 1e-2
    tx = to_i16(sdrm, PLUTO_FLOATS.reshape(1, 25, 2), 32768.0)
    assert np.array_equal(tx.ravel(), np.arange(50, dtype=np.int16) * 16)


def test_conversions_bit_exact_against_oracle(sdrm, port):
    rng = np.random.default_rng(5)
    x16 = rng.integers(-32768, 32768, (7, 4099, 2), dtype=np.int16)
    for scalar in (2048.0, 32768.0, 3.0):
        assert same_bits(to_cf32(sdrm, x16, scalar), port.convert_16i_32f(x16, scalar))
    x = rng.standard_normal((5, 3001, 2)).astype(np.float32)
    x[0, :8] = [[1.0, -1.0], [0.99999, -0.99999], [2.0, -2.0], [1.5 / 32768, 2.5 / 32768], [-1.5 / 32768, -2.5 / 32768],
                [np.inf, -np.inf], [np.nan, 0.0], [-0.0, 1e-30]]
    for scalar in (32768.0, 127.0):
        assert np.array_equal(to_i16(sdrm, x, scalar), port.convert_32f_16i(x, scalar))


def test_demod_from_int16_equals_converting_first(sdrm, port):
    """submit_i16 == the host conversion of plutosdr.c:129 followed by the cf32 call, bit for bit"""
    import workloads
    shape = workloads.C2_PARITY
    n_ch, n = 6, 3 * shape.chunk
    iq = workloads.gfsk_channels(n_ch, n, shape, seed=77, device="cpu").numpy()
    iq16 = np.clip(np.rint(iq.view(np.float32).reshape(n_ch, n, 2) * 1500.0), -2048, 2047).astype(np.int16)
    as_float = port.convert_16i_32f(iq16, 2048.0).reshape(n_ch, n, 2).copy().view(np.complex64)[:, :, 0]
    batch = sdrm.FskDemodBatch(n_ch, *shape.create_args, shape.chunk)
    got = [[] for _ in range(n_ch)]
    for o in range(0, n, shape.chunk):
        batch.submit_i16(iq16[:, o:o + shape.chunk])
        hard, lens, _ = batch.fetch()
        for c in range(n_ch):
            got[c].append(hard[c, :lens[c]].copy())
    for c in range(n_ch):
        want, _ = port.FskDemod(*shape.create_args, shape.chunk).run(as_float[c], shape.chunk)
        assert same_bits(np.concatenate(got[c]), want)


def test_modulator_int16_egress(sdrm, port):
    rng = np.random.default_rng(9)
    data = rng.integers(0, 256, (4, 300), dtype=np.uint8)
    sens = float(np.float32(2 * np.pi * 5000 / 19200))
    a = sdrm.GfskModBatch(4, 2.0, sens, 0.5, 512)
    b = sdrm.GfskModBatch(4, 2.0, sens, 0.5, 512)
    for _ in range(3):  # carried phase / filter state across calls
        cf = a.process(data)
        i16 = b.process_i16(data)
        want = port.convert_32f_16i(cf.view(np.float32).reshape(4, -1, 2), 32768.0)
        assert np.array_equal(i16, want)
