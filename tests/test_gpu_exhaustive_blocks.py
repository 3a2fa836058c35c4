"""Sweeps over whole input domains, where a block's input space is small enough to try every value (companion of
tests/test_gpu_sincos_sweep.py). All through public entry points, against the oracle on the host.

* fast_atan2f (reference src/math/fast_atan2f.c:87-157) is a function of the ratio z = min(|y|,|x|) / max(|y|,|x|) in [0, 1]
  — a division, a 255-step table lookup with linear interpolation, a small-angle shortcut — followed by an octant fix-up.
  The quadrature_demod handle computes gain * fast_atan2f(im, re) of cur * conj(prev); with prev = (1, 0) and gain 1 that is
  fast_atan2f(cur.im, cur.re) exactly, so a stream (1,0), c1, (1,0), c2, ... evaluates it at any chosen points, and at
  (-c.im, c.re) in between. EVERY float ratio in [0, 1] (1 065 353 217 values) goes through the first octant (y = z, x = 1);
  the other seven octants and the swapped form (y = 1, x = z) share that code path up to the final add / subtract and are
  swept with a stride.
* volk_32f_s32f_convert_16i (reference src/sdr/plutosdr.c:83: saturate, round half to even, NaN -> 0) over ALL 2^32 floats for
  the PlutoSDR's scalar, and volk_16i_s32f_convert_32f (plutosdr.c:129) over all 65536 int16 values for several scalars.

SDRM_SWEEP=quick runs every 31st block only."""
import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest
import torch

from conftest import same_bits
from test_gpu_blocks import make_quad

pytestmark = pytest.mark.gpu

QUICK = os.environ.get("SDRM_SWEEP", "") == "quick"
ONE_BITS = 0x3F800000
BLOCK = 1 << 22


def bounded_map(pool, fn, items, window=12):
    """pool.map that keeps at most `window` results ahead of the consumer (a block's arrays are ~100 MB)"""
    pending = []
    for item in items:
        pending.append(pool.submit(fn, item))
        if len(pending) >= window:
            yield pending.pop(0).result()
    for future in pending:
        yield future.result()


def atan_points(first_bits, count, octant, swapped):
    """complex64 stream (1,0), c_0, (1,0), c_1, ... with c_k built from the ratio z_k = float(first_bits + k)"""
    z = (np.arange(count, dtype=np.uint32) + np.uint32(first_bits)).view(np.float32)
    one = np.ones(count, np.float32)
    y, x = (one, z) if swapped else (z, one)
    if octant & 1:
        y = -y
    if octant & 2:
        x = -x
    stream = np.zeros(2 * count, np.complex64)
    parts = stream.view(np.float32).reshape(-1, 2)
    parts[0::2, 0] = 1.0
    parts[1::2, 0] = x
    parts[1::2, 1] = y
    return stream


def sweep_atan(sdrm, port, octant, swapped, stride):
    lib = sdrm.lib
    q = make_quad(lib, 1.0, 2 * BLOCK)
    checked = 0
    firsts = [f for i, f in enumerate(range(0, ONE_BITS + 1, BLOCK)) if stride == 1 or i % stride == 0 or f + BLOCK > ONE_BITS or f == 0]

    def oracle_block(first):
        count = min(BLOCK, ONE_BITS + 1 - first)
        stream = atan_points(first, count, octant, swapped)
        o = port.QuadDemod(1.0)
        return stream, o.process(stream)

    with ThreadPoolExecutor(8) as pool:
        for first, (stream, want) in zip(firsts, bounded_map(pool, oracle_block, firsts)):
            # a fresh handle state per block on both sides: re-create is cheap next to 8M samples, but a (1,0) sample ends
            # every block's predecessor anyway, so only the very first output of a block depends on the previous block
            got = q.process(stream)
            assert len(got) == len(want)
            assert same_bits(got[1:], want[1:]), "octant %d swapped %d block at ratio bits 0x%08x" % (octant, swapped, first)
            checked += len(stream) // 2
    q.close()
    return checked


def test_fast_atan2f_every_ratio_in_the_first_octant(sdrm, port):
    checked = sweep_atan(sdrm, port, 0, False, 31 if QUICK else 1)
    print("fast_atan2f: %d ratios through quadrature_demod, all equal to the oracle" % checked)
    assert checked > (ONE_BITS // 40 if QUICK else ONE_BITS)


@pytest.mark.parametrize("octant,swapped", [(o, s) for o in range(4) for s in (False, True) if (o, s) != (0, False)])
def test_fast_atan2f_other_octants_strided(sdrm, port, octant, swapped):
    checked = sweep_atan(sdrm, port, octant, swapped, 127 if QUICK else 16)
    assert checked > ONE_BITS // 200


def test_float_to_int16_every_float(sdrm, port):
    """2^32 floats x the PlutoSDR scalar (32768): blocks of 2^24 bit patterns made on the device, converted, compared on the host"""
    lib = sdrm.lib
    block = 1 << 24
    scalar = 32768.0
    checked = 0
    firsts = [f for i, f in enumerate(range(0, 1 << 32, block)) if not QUICK or i % 31 == 0]

    def oracle_block(first):
        x = (np.arange(block, dtype=np.uint32) + np.uint32(first)).view(np.float32)
        return np.asarray(port.convert_32f_16i(x, scalar)).reshape(-1)

    with ThreadPoolExecutor(8) as pool:
        for first, want in zip(firsts, bounded_map(pool, oracle_block, firsts)):
            bits = torch.arange(first, first + block, dtype=torch.int64, device="cuda").to(torch.int32)  # wraps into the sign bit
            d_in = bits.view(torch.float32).view(1, -1)
            d_out = torch.zeros(1, block, dtype=torch.int16, device="cuda")
            n = block // 2  # complex samples
            assert lib.sdrm_samples_cf32_to_i16_device(d_in.data_ptr(), n, d_out.data_ptr(), n, scalar, n, 1, None) == 0
            torch.cuda.synchronize()
            assert np.array_equal(d_out.cpu().numpy().reshape(-1), want), "block at bits 0x%08x" % first
            checked += block
    print("float -> int16: %d floats, all equal to the oracle" % checked)
    assert checked >= ((1 << 32) // 40 if QUICK else 1 << 32)


def test_int16_to_float_every_value(sdrm, port):
    lib = sdrm.lib
    x16 = np.arange(-32768, 32768, dtype=np.int16)
    for scalar in (2048.0, 32768.0, 1.0, 3.0, 1e-3, 2047.5):
        d_in = torch.from_numpy(x16.copy()).cuda().view(1, -1)
        n = len(x16) // 2
        d_out = torch.zeros(1, len(x16), dtype=torch.float32, device="cuda")
        assert lib.sdrm_samples_i16_to_cf32_device(d_in.data_ptr(), n, d_out.data_ptr(), n, scalar, n, 1, None) == 0
        torch.cuda.synchronize()
        assert same_bits(d_out.cpu().numpy().reshape(-1), np.asarray(port.convert_16i_32f(x16, scalar)).reshape(-1)), scalar
