"""TEST INFRASTRUCTURE ONLY: ctypes access to oracle/liboracle_port.so (oracle/sdrm_oracle.c, our C restatement).

Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module. The product library never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_port.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    vp, sz, fp = C.c_void_p, C.c_size_t, C.POINTER(C.c_float)
    lib.orc_low_pass_taps.argtypes = [C.c_float, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(fp), C.POINTER(sz)]
    lib.orc_gaussian_taps.argtypes = [C.c_double, C.c_double, C.c_double, sz, C.POINTER(fp)]
    lib.orc_convolve.argtypes = [vp, sz, vp, sz, C.POINTER(fp), C.POINTER(sz)]
    lib.orc_fast_atan2f.restype = C.c_float
    lib.orc_fast_atan2f.argtypes = [C.c_float, C.c_float]
    lib.orc_fir_create.restype = vp
    lib.orc_fir_create.argtypes = [vp, sz, C.c_int, C.c_int]
    lib.orc_fir_process.restype = sz
    lib.orc_fir_process.argtypes = [vp, vp, sz, vp]
    lib.orc_fir_destroy.argtypes = [vp]
    lib.orc_quad_demod_init.argtypes = [vp, C.c_float]
    lib.orc_quad_demod_process.argtypes = [vp, vp, sz, vp]
    lib.orc_dc_blocker_create.restype = vp
    lib.orc_dc_blocker_create.argtypes = [C.c_int]
    lib.orc_dc_blocker_process.argtypes = [vp, vp, sz]
    lib.orc_dc_blocker_destroy.argtypes = [vp]
    lib.orc_clock_mm_create.restype = vp
    lib.orc_clock_mm_create.argtypes = [C.c_float] * 5 + [sz]
    lib.orc_clock_mm_process.restype = sz
    lib.orc_clock_mm_process.argtypes = [vp, vp, sz, vp]
    lib.orc_clock_mm_destroy.argtypes = [vp]
    lib.orc_convert_8i.argtypes = [vp, C.c_float, sz, vp]
    lib.orc_convert_16i_32f.argtypes = [vp, C.c_float, sz, vp]
    lib.orc_convert_32f_16i.argtypes = [vp, C.c_float, sz, vp]
    lib.orc_fsk_demod_create.restype = vp
    lib.orc_fsk_demod_create.argtypes = [C.c_uint64, C.c_uint32, C.c_int64, C.c_uint8, C.c_uint32, C.c_int, C.c_uint32]
    lib.orc_fsk_demod_process.restype = sz
    lib.orc_fsk_demod_process.argtypes = [vp, vp, sz, vp, vp]
    lib.orc_fsk_demod_destroy.argtypes = [vp]
    lib.orc_sig_source_init.argtypes = [vp, C.c_float, C.c_uint64]
    lib.orc_sig_source_generate.argtypes = [vp, C.c_int64, sz, vp]
    lib.orc_sig_source_multiply.argtypes = [vp, C.c_int64, vp, sz, vp]
    lib.orc_freq_mod_process.argtypes = [vp, vp, sz, vp]
    lib.orc_gfsk_mod_create.restype = vp
    lib.orc_gfsk_mod_create.argtypes = [C.c_float, C.c_float, C.c_float, C.c_uint32]
    lib.orc_gfsk_mod_process.restype = sz
    lib.orc_gfsk_mod_process.argtypes = [vp, vp, sz, vp]
    lib.orc_gfsk_mod_destroy.argtypes = [vp]
    lib.orc_bench_fsk_demod.restype = C.c_double
    lib.orc_bench_fsk_demod.argtypes = [C.c_uint64, C.c_uint32, C.c_int64, C.c_uint8, C.c_uint32, C.c_int, C.c_uint32,
                                        vp, sz, sz, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    lib.orc_sincos_sweep.restype = sz
    lib.orc_sincos_sweep.argtypes = [C.c_uint32, sz, vp, C.POINTER(C.c_uint32)]
    lib.free = C.CDLL(None).free
    lib.free.argtypes = [vp]
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def low_pass_taps(gain, fs, cutoff, tw):
    lib = load()
    p, n = C.POINTER(C.c_float)(), C.c_size_t()
    code = lib.orc_low_pass_taps(gain, fs, cutoff, tw, C.byref(p), C.byref(n))
    if code != 0:
        raise ValueError("orc_low_pass_taps failed with %d" % code)
    out = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
    lib.free(p)
    return out


def gaussian_taps(gain, sps, bt, n):
    lib = load()
    p = C.POINTER(C.c_float)()
    lib.orc_gaussian_taps(gain, sps, bt, n, C.byref(p))
    out = np.ctypeslib.as_array(p, shape=(n,)).copy()
    lib.free(p)
    return out


def convolve(x, y):
    lib = load()
    x, y = _f32(x), _f32(y)
    p, n = C.POINTER(C.c_float)(), C.c_size_t()
    lib.orc_convolve(_p(x), len(x), _p(y), len(y), C.byref(p), C.byref(n))
    out = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
    lib.free(p)
    return out


def convert_16i_32f(x, scalar):
    x = np.ascontiguousarray(x, dtype=np.int16)
    out = np.empty(x.shape, np.float32)
    load().orc_convert_16i_32f(_p(x), scalar, x.size, _p(out))
    return out


def convert_32f_16i(x, scalar):
    x = _f32(x)
    out = np.empty(x.shape, np.int16)
    load().orc_convert_32f_16i(_p(x), scalar, x.size, _p(out))
    return out


def fast_atan2f(y, x):
    lib = load()
    return np.array([lib.orc_fast_atan2f(float(a), float(b)) for a, b in zip(_f32(y).ravel(), _f32(x).ravel())],
                    dtype=np.float32)


class Fir:
    def __init__(self, taps, decimation, complex_input):
        self.lib = load()
        taps = _f32(taps)
        self.width = 2 if complex_input else 1
        self.decimation = decimation
        self.h = self.lib.orc_fir_create(_p(taps), len(taps), decimation, self.width)

    def process(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex64 if self.width == 2 else np.float32)
        out = np.zeros((len(x) // self.decimation + 2) * self.width, dtype=np.float32)
        n = self.lib.orc_fir_process(self.h, _p(x), len(x), _p(out))
        out = out[:n * self.width].copy()
        return out.view(np.complex64) if self.width == 2 else out

    def run(self, x, chunk):
        parts = [self.process(x[o:o + chunk]) for o in range(0, len(x), chunk)]
        return np.concatenate(parts)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_fir_destroy(self.h)
            self.h = None


class QuadDemod:
    def __init__(self, gain):
        self.lib = load()
        self.state = (C.c_float * 3)()
        self.lib.orc_quad_demod_init(self.state, gain)

    def process(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        out = np.zeros(len(x), dtype=np.float32)
        self.lib.orc_quad_demod_process(self.state, _p(x), len(x), _p(out))
        return out

    def run(self, x, chunk):
        return np.concatenate([self.process(x[o:o + chunk]) for o in range(0, len(x), chunk)])


class DcBlocker:
    def __init__(self, length):
        self.lib = load()
        self.h = self.lib.orc_dc_blocker_create(length)

    def process(self, x):
        x = _f32(x).copy()
        self.lib.orc_dc_blocker_process(self.h, _p(x), len(x))
        return x

    def run(self, x, chunk):
        return np.concatenate([self.process(x[o:o + chunk]) for o in range(0, len(x), chunk)])

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_dc_blocker_destroy(self.h)
            self.h = None


class ClockMm:
    def __init__(self, omega, gain_omega, mu, gain_mu, omega_relative_limit, max_len):
        self.lib = load()
        self.max_len = max_len
        self.h = self.lib.orc_clock_mm_create(omega, gain_omega, mu, gain_mu, omega_relative_limit, max_len)

    def process(self, x):
        x = _f32(x)
        out = np.zeros(self.max_len + 8, dtype=np.float32)
        n = self.lib.orc_clock_mm_process(self.h, _p(x), len(x), _p(out))
        return out[:n].copy()

    def run(self, x, chunk):
        return np.concatenate([self.process(x[o:o + chunk]) for o in range(0, len(x), chunk)])

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_clock_mm_destroy(self.h)
            self.h = None


class FskDemod:
    def __init__(self, fs, baud, deviation, decimation, tw, use_dc, max_len):
        self.lib = load()
        self.max_len = max_len
        self.h = self.lib.orc_fsk_demod_create(fs, baud, deviation, decimation, tw, int(bool(use_dc)), max_len)
        if not self.h:
            raise ValueError("orc_fsk_demod_create failed")

    def process(self, iq):
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        hard = np.zeros(self.max_len + 8, dtype=np.int8)
        soft = np.zeros(self.max_len + 8, dtype=np.float32)
        n = self.lib.orc_fsk_demod_process(self.h, _p(iq), len(iq), _p(hard), _p(soft))
        return hard[:n].copy(), soft[:n].copy()

    def run(self, iq, chunk):
        parts = [self.process(iq[o:o + chunk]) for o in range(0, len(iq), chunk)]
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_fsk_demod_destroy(self.h)
            self.h = None


class SigSource:
    def __init__(self, amplitude, fs):
        self.lib = load()

        class _S(C.Structure):
            _fields_ = [("phase", C.c_float), ("amplitude", C.c_float), ("fs", C.c_uint64)]
        self.state = _S()
        self.lib.orc_sig_source_init(C.byref(self.state), amplitude, fs)

    def generate(self, freq, n):
        out = np.zeros(n, dtype=np.complex64)
        self.lib.orc_sig_source_generate(C.byref(self.state), int(freq), n, _p(out))
        return out

    def multiply(self, freq, x):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        out = np.zeros(len(x), dtype=np.complex64)
        self.lib.orc_sig_source_multiply(C.byref(self.state), int(freq), _p(x), len(x), _p(out))
        return out


class FreqMod:
    def __init__(self, sensitivity):
        self.lib = load()
        self.state = (C.c_float * 2)(0.0, sensitivity)

    def process(self, x):
        x = _f32(x)
        out = np.zeros(len(x), dtype=np.complex64)
        self.lib.orc_freq_mod_process(self.state, _p(x), len(x), _p(out))
        return out


class GfskMod:
    def __init__(self, sps, sensitivity, bt, max_bytes):
        self.lib = load()
        self.sps = int(sps)
        self.h = self.lib.orc_gfsk_mod_create(sps, sensitivity, bt, max_bytes)
        if not self.h:
            raise ValueError("orc_gfsk_mod_create failed")

    def process(self, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(len(data) * 8 * self.sps + 8, dtype=np.complex64)
        n = self.lib.orc_gfsk_mod_process(self.h, _p(data), len(data), _p(out))
        return out[:n].copy()

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_gfsk_mod_destroy(self.h)
            self.h = None


def bench_fsk_demod(fs, baud, deviation, decimation, tw, use_dc, chunk, iq_channels, n_threads, passes=1):
    lib = load()
    a = np.ascontiguousarray(iq_channels, dtype=np.complex64)
    nch, n = a.shape
    sym = C.c_uint64()
    sec = lib.orc_bench_fsk_demod(fs, baud, deviation, decimation, tw, int(bool(use_dc)), chunk, _p(a), 2 * n, n, nch,
                                  n_threads, passes, C.byref(sym))
    if sec < 0:
        raise ValueError("orc_bench_fsk_demod failed")
    return sec, sym.value


def sincos_sweep(first_bits, got, threads=16):
    """got: float32 array [count][2] of (cos, sin) for the float bit patterns first_bits ...; returns (number of values that
    differ from libm's double cos / sin rounded to float, first differing pattern or None). Split over threads (ctypes
    releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    lib = load()
    got = np.ascontiguousarray(got, dtype=np.float32)
    count = got.shape[0]
    edges = np.linspace(0, count, threads + 1).astype(np.int64)

    def part(k):
        lo, hi = int(edges[k]), int(edges[k + 1])
        bad_at = C.c_uint32(0)
        bad = lib.orc_sincos_sweep(first_bits + lo, hi - lo, got[lo:].ctypes.data_as(C.c_void_p), C.byref(bad_at))
        return bad, bad_at.value
    with ThreadPoolExecutor(threads) as pool:
        results = list(pool.map(part, range(threads)))
    total = sum(r[0] for r in results)
    first = next((r[1] for r in results if r[0]), None)
    return total, first
