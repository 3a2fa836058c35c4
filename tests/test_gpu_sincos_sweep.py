"""The modulator's and the NCO's (float) cos / sin of a float phase, over EVERY float phase they can see.

The reference stores cos(phase) + I sin(phase) computed by libm in double and rounded to float (frequency_modulator.c:56,
sig_source.c). The kernels compute the same through one device function (csrc/device_math.cuh sdrm_phase_sincos). Two double
routines that are each accurate to an ulp or so agree after rounding to float except when the true value sits within a few
double ulps of a float rounding boundary — about one value in 10^8 — so "it matched on the test signals" is not a proof. The
phase is a float that the wrap keeps inside [-2 pi, 2 pi] (frequency_modulator.c:50-55, sig_source.c), which is few enough
values to try them all: 2 x 1 088 425 985 floats (up to 7.0). This test sweeps all of them against libm on the host (the oracle's C side);
SDRM_SINCOS_SWEEP=quick limits it to every 61st block for a fast run."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TWO_PI_BITS = struct.unpack("<I", struct.pack("<f", 6.283185307179586))[0]  # 0x40C90FDB
# the wrap compares with > and <, so +-2 pi itself is a phase; the sweep runs on to 7.0 for margin
LAST_BITS = struct.unpack("<I", struct.pack("<f", 7.0))[0] + 4096
BLOCK = 1 << 24


def sweep(sdrm, sincos_sweep, sign_bit, stride_blocks):
    lib = sdrm.lib
    lib.sdrm_cu_selftest_sincos.argtypes = [C.c_uint32, C.c_size_t, C.c_void_p]
    lib.sdrm_cu_selftest_sincos.restype = C.c_int
    last = LAST_BITS
    out = np.empty((BLOCK, 2), np.float32)
    checked = 0
    bad_total = 0
    first_bad = None
    block_index = 0
    for first in range(0, last + 1, BLOCK):
        block_index += 1
        edge = first == 0 or first + BLOCK > last  # the blocks with zero / subnormals and with 2 pi always run
        if not edge and stride_blocks > 1 and block_index % stride_blocks != 0:
            continue
        count = min(BLOCK, last + 1 - first)
        assert lib.sdrm_cu_selftest_sincos(sign_bit | first, count, out.ctypes.data_as(C.c_void_p)) == 0
        bad, where = sincos_sweep(sign_bit | first, out[:count])
        checked += count
        bad_total += bad
        if bad and first_bad is None:
            first_bad = where
    return checked, bad_total, first_bad


@pytest.mark.parametrize("sign", ["positive", "negative"])
def test_every_float_phase_gives_libms_cos_and_sin(sdrm, port, sign):
    from oracle import port as port_module
    quick = os.environ.get("SDRM_SINCOS_SWEEP", "") == "quick"
    checked, bad, first_bad = sweep(sdrm, port_module.sincos_sweep, 0x80000000 if sign == "negative" else 0, 61 if quick else 1)
    print("sincos sweep (%s phases): %d floats checked, %d differ from libm%s" %
          (sign, checked, bad, "" if first_bad is None else ", first at bits 0x%08x" % first_bad))
    assert checked >= (TWO_PI_BITS if not quick else TWO_PI_BITS // 100)
    assert bad == 0
