"""CPU-only: the C-ABI shared library loads here (no GPU) and exports every function include/sdrm/*.h declares;
without a GPU its entry points fail loudly instead of falling back to a CPU path."""
import ctypes as C
import glob
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, same_bits

HEADERS = sorted(glob.glob(os.path.join(ROOT, "include", "sdrm", "*.h")))
DECL = re.compile(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?[\s\*](\w+)\s*\([^;{]*\)\s*;", re.M | re.S)


def declared_functions(path):
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    text = text.replace("\\\n", " ")  # continuation lines belong to their #define
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    text = re.sub(r"__attribute__\s*\(\(.*?\)\)", "", text)
    # struct and enum bodies (members may be function pointers) hold no function declarations
    text = re.sub(r"(?:typedef\s+)?(?:struct|enum)\s*\w*\s*\{[^{}]*\}\s*\w*\s*;", "", text, flags=re.S)
    names = []
    for stmt in text.split(";"):
        m = re.search(r"(\w+)\s*\(", stmt)
        if m and "typedef" not in stmt and m.group(1) not in ("defined",):
            names.append(m.group(1))
    return names


def test_headers_exist():
    names = {os.path.basename(h) for h in HEADERS}
    for needed in ("fsk_demod.h", "lpf.h", "quadrature_demod.h", "dc_blocker.h", "clock_recovery_mm.h", "sdrm_batch.h"):
        assert needed in names


@pytest.mark.parametrize("header", HEADERS, ids=[os.path.basename(h) for h in HEADERS])
def test_library_exports_every_declared_symbol(sdrm, header):
    names = declared_functions(header)
    if os.path.basename(header) in ("server_config.h",):
        assert not names  # a struct layout only (the reference's src/server_config.h:17-44), nothing to export
        return
    assert names, "no declarations parsed from %s" % header
    for name in names:
        assert hasattr(sdrm.lib, name), "%s declared in %s but not exported" % (name, os.path.basename(header))


def test_headers_compile_as_c99():
    for header in HEADERS:
        src = '#include "%s"\nint main(void) { return 0; }\n' % header
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", "-"], input=src.encode(),
                       check=True)


def test_no_cpu_fallback_without_gpu(sdrm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sdrm.SdrmError):
        sdrm.FskDemodBatch(1, 48000, 4800, 5000, 2, 2000, True, 4096)
    with pytest.raises(sdrm.SdrmError):
        sdrm.FskDemod(48000, 4800, 5000, 2, 2000, True, 4096)


def test_host_tap_design_matches_oracle(sdrm, port):
    """create_low_pass_filter runs on the host (it is part of *_create); it must give the oracle's taps bit for bit."""
    lib = sdrm.lib
    lib.create_low_pass_filter.argtypes = [C.c_float, C.c_uint64, C.c_uint64, C.c_uint32,
                                           C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_size_t)]
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    for args in ((192000, 9800, 980), (192000, 4800, 2000), (48000, 7400, 740), (2400000, 6200, 620), (8000, 1750, 500)):
        p, n = C.POINTER(C.c_float)(), C.c_size_t()
        assert lib.create_low_pass_filter(1.0, *args, C.byref(p), C.byref(n)) == 0
        taps = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
        libc.free(p)
        assert np.array_equal(taps.view(np.uint32), port.low_pass_taps(1.0, *args).view(np.uint32))
    p, n = C.POINTER(C.c_float)(), C.c_size_t()
    for bad in ((0, 1750, 500), (8000, 5000, 500), (8000, 1750, 0)):  # reference test/test_lpf_taps.c bounds
        assert lib.create_low_pass_filter(1.0, *bad, C.byref(p), C.byref(n)) == -1


def test_host_tap_design_matches_reference_over_random_parameters(sdrm, port):
    """300 random (sampling rate, cutoff, transition width) triples across the range fsk_demod_create and dsp_worker derive
    (8 kHz ... 10 MHz, 3 ... 20000 taps), and Gaussian taps for random samples-per-symbol: the product's host code against
    the reference build where it is present, else the restatement (which tests/test_oracle.py pins to the reference)."""
    from oracle import ref
    lib = sdrm.lib
    lib.create_low_pass_filter.argtypes = [C.c_float, C.c_uint64, C.c_uint64, C.c_uint32,
                                           C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_size_t)]
    lib.gaussian_taps_create.argtypes = [C.c_double, C.c_double, C.c_double, C.c_size_t, C.POINTER(C.POINTER(C.c_float))]
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    want_lpf = ref.lpf_taps if ref.available() else port.low_pass_taps
    want_gauss = ref.gaussian_taps if ref.available() else port.gaussian_taps
    rng = np.random.default_rng(2024)
    lengths = []
    for _ in range(300):
        fs = int(10 ** rng.uniform(np.log10(8000), 7))
        cutoff = int(fs * rng.uniform(0.002, 0.45))
        tw = max(1, int(fs * 10 ** rng.uniform(-3.6, -0.7)))
        gain = float(np.float32(rng.choice([1.0, 0.5, 2.25])))
        p, n = C.POINTER(C.c_float)(), C.c_size_t()
        code = lib.create_low_pass_filter(gain, fs, cutoff, tw, C.byref(p), C.byref(n))
        if code != 0:
            with pytest.raises(Exception):
                want_lpf(gain, fs, cutoff, tw)
            continue
        taps = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
        libc.free(p)
        assert same_bits(taps, want_lpf(gain, fs, cutoff, tw)), (gain, fs, cutoff, tw)
        lengths.append(len(taps))
    assert len(lengths) > 250 and min(lengths) < 20 and max(lengths) > 5000
    for _ in range(100):
        sps = float(rng.uniform(1.5, 40.0))
        bt = float(rng.choice([0.3, 0.5, 1.0]))
        ntaps = int(rng.integers(2, 200))
        p = C.POINTER(C.c_float)()
        assert lib.gaussian_taps_create(1.0, sps, bt, ntaps, C.byref(p)) == 0
        taps = np.ctypeslib.as_array(p, shape=(ntaps,)).copy()
        libc.free(p)
        assert same_bits(taps, want_gauss(1.0, sps, bt, ntaps)), (sps, bt, ntaps)


def test_host_fast_atan2f_matches_oracle(sdrm, port):
    lib = sdrm.lib
    lib.fast_atan2f.restype = C.c_float
    lib.fast_atan2f.argtypes = [C.c_float, C.c_float]
    rng = np.random.default_rng(0)
    y = np.concatenate([rng.standard_normal(2000), [0, 0, 1, -1, 0.0039, np.inf]]).astype(np.float32)
    x = np.concatenate([rng.standard_normal(2000), [0, 1, 0, 0, 1.0, 1]]).astype(np.float32)
    got = np.array([lib.fast_atan2f(float(a), float(b)) for a, b in zip(y, x)], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), port.fast_atan2f(y, x).view(np.uint32))


def test_cpulist_parser(sdrm):
    """sysfs cpu lists behind sdrm_bind_thread_near_device (host/affinity.c)."""
    count = sdrm.lib.sdrm_cpulist_parse_count
    assert count(b"0-23\n") == 24
    assert count(b"0-23,48-71\n") == 48
    assert count(b"5") == 1
    assert count(b"") == 0
    assert count(b"3-1") == -1
    assert count(b"x") == -1


def test_bind_thread_without_gpu_is_an_error_not_a_crash(sdrm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    assert sdrm.lib.sdrm_bind_thread_near_device(0) < 0


def test_division_steps_check_runs_on_the_host(sdrm):
    """host/taps.c: exhaustive proof, per dc-blocker length, of how many corrections the tail's branch-free division needs."""
    steps = sdrm.lib.sdrm_division_steps
    for length in (320, 160, 107, 32, 3, 33333):
        assert steps(length) == 1
    assert steps(0) == 0


def test_every_public_header_compiles_on_its_own_as_c99(tmp_path):
    """a host that includes one header must not need another one first (the reference's headers have the same property), and
    nothing in them goes beyond C99"""
    headers = sorted(glob.glob(os.path.join(ROOT, "include", "sdrm", "*.h")))
    assert len(headers) >= 20
    for header in headers:
        source = tmp_path / "one.c"
        source.write_text('#include "%s"\nint main(void) { return 0; }\n' % header)
        proc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-c", str(source), "-o", str(tmp_path / "one.o")],
                              capture_output=True, text=True)
        assert proc.returncode == 0, "%s:\n%s" % (os.path.basename(header), proc.stderr[:2000])
