import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sdr-modem_b200"))
GOLDEN = os.path.join(ROOT, "tests", "golden")

LUCKY7_TLE = ["LUCKY-7",
              "1 44406U 19038W   20069.88080907  .00000505  00000-0  32890-4 0  9992",
              "2 44406  97.5270  32.5584 0026284 107.4758 252.9348 15.12089395 37524"]

# the four end-to-end goldens of reference test/test_fsk_demod.c:52-81: (input, expected, fsk_demod_create args)
FSK_GOLDENS = {
    "nusat": ("nusat.cf32", "processed.s8", (192000, 40000, 5000, 1, 2000, True)),
    "nan": ("inputnan.cf32", "nan.s8", (240000, 9600, 5000, 1, 2000, True)),
    "lucky7": ("lucky7.expected.cf32", "lucky7.expected.s8", (48000, 4800, 5000, 2, 2000, True)),
    "lucky7_nodc": ("lucky7.expected.cf32", "lucky7.expected.nodc.s8", (48000, 4800, 5000, 2, 2000, False)),
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_array(name, dtype):
    return np.fromfile(os.path.join(GOLDEN, name), dtype=dtype)


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(GOLDEN, "reference_kats.json")) as f:
        return {k: np.array(v, dtype=np.float32) for k, v in json.load(f).items()}


@pytest.fixture(scope="session")
def port():
    from oracle import port as module
    module.load()
    return module


@pytest.fixture(scope="session")
def ref():
    """The reference's own sources compiled in place (oracle/_ref). Present in the build container and, prebuilt, on
    the GPU box; tests that need it skip when it was never built."""
    from oracle import ref as module
    if not module.available():
        pytest.skip("oracle/_ref/libsdrmodem_ref.so not built")
    module.load()
    return module


@pytest.fixture(scope="session")
def sdrm():
    import sdrm as module
    return module


def ramp(n, offset=0):
    """reference test/utils.c:104-113 setup_input_data"""
    return (np.arange(n) + offset).astype(np.float32)


def complex_ramp(n, offset=0):
    """reference test/utils.c:126-134 setup_input_complex_data: (2i) + (2i+1)j"""
    i = np.arange(n) + offset
    return ((2 * i).astype(np.float32) + 1j * (2 * i + 1).astype(np.float32)).astype(np.complex64)


def same_bits(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype == np.complex64:
        a, b = a.view(np.float32), b.view(np.float32)
    if a.dtype == np.float32:
        ua, ub = a.view(np.uint32), b.view(np.uint32)
        both_nan = np.isnan(a) & np.isnan(b)  # NaN payloads differ between x86 and the GPU; NaN-ness must not
        return bool(np.all((ua == ub) | both_nan))
    return bool(np.array_equal(a, b))
