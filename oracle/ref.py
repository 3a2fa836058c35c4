"""TEST INFRASTRUCTURE ONLY: ctypes access to oracle/_ref/libsdrmodem_ref*.so.

That library is the reference's own src/dsp + src/math + src/sgpsdp compiled in place by oracle/Makefile
(strict build: -ffp-contract=off, VOLK generic semantics from oracle/shim/volk/volk.h). It is the
parity-defining checker. Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

c_float_p = C.POINTER(C.c_float)


def tuned_level():
    """x86-64 micro-architecture level of the tuned CPU build this machine can run: 4 (AVX-512), 3 (AVX2 + FMA) or 0."""
    try:
        with open("/proc/cpuinfo") as f:
            flags = set()
            for line in f:
                if line.startswith("flags"):
                    flags = set(line.split(":", 1)[1].split())
                    break
    except OSError:
        return 0
    if {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= flags:
        return 4
    if {"avx2", "fma", "bmi2", "movbe"} <= flags:
        return 3
    return 0


def lib_path(fma=False):
    if fma == "tuned":
        return os.path.join(_HERE, "_ref", "libsdrmodem_ref_tuned_v%d.so" % tuned_level())
    return os.path.join(_HERE, "_ref", "libsdrmodem_ref_fma.so" if fma else "libsdrmodem_ref.so")


def available(fma=False):
    return os.path.exists(lib_path(fma))


def load(fma=False):
    """fma: False = strict build (the parity-defining checker), True = FMA-order build (checker of fast mode),
    "tuned" = -O3 vectorised build (CPU baseline only, never a checker)."""
    key = fma if fma == "tuned" else bool(fma)
    if key in _LIBS:
        return _LIBS[key]
    lib = C.CDLL(lib_path(fma), mode=os.RTLD_LOCAL | os.RTLD_NOW)
    lib.ref_fsk_chain_run.restype = C.c_long
    lib.ref_fsk_chain_run.argtypes = [C.c_uint64, C.c_uint32, C.c_int64, C.c_uint8, C.c_uint32, C.c_int, C.c_uint32,
                                      C.c_void_p, C.c_size_t, C.c_size_t,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_size_t]
    lib.ref_fsk_demod_run.restype = C.c_long
    lib.ref_fsk_demod_run.argtypes = [C.c_uint64, C.c_uint32, C.c_int64, C.c_uint8, C.c_uint32, C.c_int, C.c_uint32,
                                      C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t]
    lib.ref_bench_fsk_demod.restype = C.c_double
    lib.ref_bench_fsk_demod.argtypes = [C.c_uint64, C.c_uint32, C.c_int64, C.c_uint8, C.c_uint32, C.c_int, C.c_uint32,
                                        C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_uint64)]
    lib.ref_bench_gfsk_mod.restype = C.c_double
    lib.ref_bench_gfsk_mod.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_size_t, C.c_size_t,
                                       C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    lib.fast_atan2f.restype = C.c_float
    lib.fast_atan2f.argtypes = [C.c_float, C.c_float]
    lib.create_low_pass_filter.restype = C.c_int
    lib.create_low_pass_filter.argtypes = [C.c_float, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(c_float_p),
                                           C.POINTER(C.c_size_t)]
    lib.gaussian_taps_create.restype = C.c_int
    lib.gaussian_taps_create.argtypes = [C.c_double, C.c_double, C.c_double, C.c_size_t, C.POINTER(c_float_p)]
    lib.gfsk_mod_convolve.restype = C.c_int
    lib.gfsk_mod_convolve.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(c_float_p),
                                      C.POINTER(C.c_size_t)]
    lib.free = C.CDLL(None).free
    lib.free.argtypes = [C.c_void_p]
    # generic block API: X_create(..., X**) / X_process(in, len, &out, &out_len, X*) / X_destroy(X*)
    for name in ("lpf", "fsk_demod", "quadrature_demod", "dc_blocker", "clock_mm", "sig_source", "doppler",
                 "gfsk_mod", "interp_fir_filter", "frequency_modulator", "fir_filter"):
        getattr(lib, name + "_destroy").argtypes = [C.c_void_p]
        getattr(lib, name + "_destroy").restype = None
    lib.lpf_create.argtypes = [C.c_uint8, C.c_uint64, C.c_uint64, C.c_uint32, C.c_size_t, C.c_size_t,
                               C.POINTER(C.c_void_p)]
    lib.lpf_process.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p]
    lib.lpf_process.restype = None
    lib.fir_filter_create.argtypes = [C.c_uint8, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.fir_filter_process.argtypes = lib.lpf_process.argtypes
    lib.fir_filter_process.restype = None
    lib.fsk_demod_create.argtypes = [C.c_uint64, C.c_uint32, C.c_int64, C.c_uint8, C.c_uint32, C.c_bool, C.c_uint32,
                                     C.POINTER(C.c_void_p)]
    lib.fsk_demod_process.argtypes = lib.lpf_process.argtypes
    lib.fsk_demod_process.restype = None
    lib.quadrature_demod_create.argtypes = [C.c_float, C.c_uint32, C.POINTER(C.c_void_p)]
    lib.quadrature_demod_process.argtypes = lib.lpf_process.argtypes
    lib.quadrature_demod_process.restype = None
    lib.dc_blocker_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.dc_blocker_process.argtypes = lib.lpf_process.argtypes
    lib.dc_blocker_process.restype = None
    lib.clock_mm_create.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_size_t,
                                    C.POINTER(C.c_void_p)]
    lib.clock_mm_process.argtypes = lib.lpf_process.argtypes
    lib.clock_mm_process.restype = None
    lib.sig_source_create.argtypes = [C.c_float, C.c_uint64, C.c_uint32, C.POINTER(C.c_void_p)]
    lib.sig_source_process.argtypes = [C.c_int64, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p]
    lib.sig_source_process.restype = None
    lib.sig_source_multiply.argtypes = [C.c_int64, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_size_t), C.c_void_p]
    lib.sig_source_multiply.restype = None
    lib.doppler_create.argtypes = [C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64,
                                   C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.doppler_process_rx.argtypes = lib.lpf_process.argtypes
    lib.doppler_process_rx.restype = None
    lib.doppler_process_tx.argtypes = lib.lpf_process.argtypes
    lib.doppler_process_tx.restype = None
    lib.gfsk_mod_create.argtypes = [C.c_float, C.c_float, C.c_float, C.c_uint32, C.POINTER(C.c_void_p)]
    lib.gfsk_mod_process.argtypes = lib.lpf_process.argtypes
    lib.gfsk_mod_process.restype = None
    lib.interp_fir_filter_create.argtypes = [C.c_void_p, C.c_size_t, C.c_uint8, C.c_uint32, C.POINTER(C.c_void_p)]
    lib.interp_fir_filter_process.argtypes = lib.lpf_process.argtypes
    lib.interp_fir_filter_process.restype = None
    lib.frequency_modulator_create.argtypes = [C.c_float, C.c_uint32, C.POINTER(C.c_void_p)]
    lib.frequency_modulator_process.argtypes = lib.lpf_process.argtypes
    lib.frequency_modulator_process.restype = None
    lib.malloc = C.CDLL(None).malloc
    lib.malloc.restype = C.c_void_p
    lib.malloc.argtypes = [C.c_size_t]
    _LIBS[key] = lib
    return lib


def _as_f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _iq_f32(iq):
    """complex64 or interleaved float32 -> contiguous float32 view (re, im, re, im …)."""
    a = np.ascontiguousarray(iq)
    if a.dtype == np.complex64:
        return a.view(np.float32)
    return _as_f32(a)


class _Block:
    """A reference block handle driven in caller-chosen chunks, copying each call's output out."""

    def __init__(self, lib, prefix, handle, process, in_dtype, out_dtype):
        self.lib, self.prefix, self.handle = lib, prefix, handle
        self._process, self.in_dtype, self.out_dtype = process, in_dtype, out_dtype

    def process(self, data):
        data = np.ascontiguousarray(data, dtype=self.in_dtype)
        out = C.c_void_p()
        out_len = C.c_size_t()
        self._process(data.ctypes.data_as(C.c_void_p), data.shape[0], C.byref(out), C.byref(out_len), self.handle)
        n = out_len.value
        if n == 0 or not out.value:
            return np.zeros(0, dtype=self.out_dtype)
        buf = (C.c_char * (n * np.dtype(self.out_dtype).itemsize)).from_address(out.value)
        return np.frombuffer(buf, dtype=self.out_dtype).copy()

    def run(self, data, chunk):
        data = np.ascontiguousarray(data, dtype=self.in_dtype)
        outs = [self.process(data[o:o + chunk]) for o in range(0, data.shape[0], chunk)]
        return np.concatenate(outs) if outs else np.zeros(0, dtype=self.out_dtype)

    def close(self):
        if self.handle is not None:
            getattr(self.lib, self.prefix + "_destroy")(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _check(code, what):
    if code != 0:
        raise ValueError("%s failed with %d" % (what, code))


def lpf(decimation, fs, cutoff, tw, max_len, complex_input, fma=False):
    lib = load(fma)
    h = C.c_void_p()
    _check(lib.lpf_create(decimation, fs, cutoff, tw, max_len, 8 if complex_input else 4, C.byref(h)), "lpf_create")
    dt = np.complex64 if complex_input else np.float32
    return _Block(lib, "lpf", h, lib.lpf_process, dt, dt)


def fir_filter(decimation, taps, max_len, complex_input, fma=False):
    """fir_filter_create takes ownership of a malloc'ed taps array (reference src/dsp/fir_filter.c:58,171)."""
    lib = load(fma)
    taps = _as_f32(taps)
    p = lib.malloc(taps.nbytes)
    C.memmove(p, taps.ctypes.data, taps.nbytes)
    h = C.c_void_p()
    _check(lib.fir_filter_create(decimation, p, taps.shape[0], max_len, 8 if complex_input else 4, C.byref(h)),
           "fir_filter_create")
    dt = np.complex64 if complex_input else np.float32
    return _Block(lib, "fir_filter", h, lib.fir_filter_process, dt, dt)


def fsk_demod(fs, baud, deviation, decimation, tw, use_dc, max_len, fma=False):
    lib = load(fma)
    h = C.c_void_p()
    _check(lib.fsk_demod_create(fs, baud, deviation, decimation, tw, bool(use_dc), max_len, C.byref(h)),
           "fsk_demod_create")
    return _Block(lib, "fsk_demod", h, lib.fsk_demod_process, np.complex64, np.int8)


def quadrature_demod(gain, max_len, fma=False):
    lib = load(fma)
    h = C.c_void_p()
    _check(lib.quadrature_demod_create(gain, max_len, C.byref(h)), "quadrature_demod_create")
    return _Block(lib, "quadrature_demod", h, lib.quadrature_demod_process, np.complex64, np.float32)


def dc_blocker(length, fma=False):
    lib = load(fma)
    h = C.c_void_p()
    _check(lib.dc_blocker_create(length, C.byref(h)), "dc_blocker_create")
    return _Block(lib, "dc_blocker", h, lib.dc_blocker_process, np.float32, np.float32)


def clock_mm(omega, gain_omega, mu, gain_mu, omega_relative_limit, max_len, fma=False):
    lib = load(fma)
    h = C.c_void_p()
    _check(lib.clock_mm_create(omega, gain_omega, mu, gain_mu, omega_relative_limit, max_len, C.byref(h)),
           "clock_mm_create")
    return _Block(lib, "clock_mm", h, lib.clock_mm_process, np.float32, np.float32)


class _SigSource(_Block):
    def multiply(self, freq, data):
        data = np.ascontiguousarray(data, dtype=np.complex64)
        out = C.c_void_p()
        out_len = C.c_size_t()
        self.lib.sig_source_multiply(int(freq), data.ctypes.data_as(C.c_void_p), data.shape[0], C.byref(out),
                                     C.byref(out_len), self.handle)
        n = out_len.value
        if n == 0 or not out.value:
            return np.zeros(0, dtype=np.complex64)
        return np.frombuffer((C.c_char * (n * 8)).from_address(out.value), dtype=np.complex64).copy()

    def generate(self, freq, n):
        out = C.c_void_p()
        out_len = C.c_size_t()
        self.lib.sig_source_process(int(freq), n, C.byref(out), C.byref(out_len), self.handle)
        n = out_len.value
        return np.frombuffer((C.c_char * (n * 8)).from_address(out.value), dtype=np.complex64).copy()


def sig_source(amplitude, fs, max_len, fma=False):
    lib = load(fma)
    h = C.c_void_p()
    _check(lib.sig_source_create(amplitude, fs, max_len, C.byref(h)), "sig_source_create")
    return _SigSource(lib, "sig_source", h, None, np.complex64, np.complex64)


def tle_buffer(tle_lines):
    buf = C.create_string_buffer(3 * 80)
    for i, line in enumerate(tle_lines):
        raw = line.encode("ascii")[:79]
        buf[i * 80:i * 80 + len(raw)] = raw
    return buf


def doppler(lat, lon, alt, fs, fc, const_offset, start_time, max_len, tle_lines, tx=False, fma=False):
    lib = load(fma)
    h = C.c_void_p()
    buf = tle_buffer(tle_lines)
    _check(lib.doppler_create(lat, lon, alt, fs, fc, const_offset, start_time, max_len, buf, C.byref(h)),
           "doppler_create")
    return _Block(lib, "doppler", h, lib.doppler_process_tx if tx else lib.doppler_process_rx, np.complex64,
                  np.complex64)


def gfsk_mod(sps, sensitivity, bt, max_len, fma=False):
    lib = load(fma)
    h = C.c_void_p()
    _check(lib.gfsk_mod_create(sps, sensitivity, bt, max_len, C.byref(h)), "gfsk_mod_create")
    return _Block(lib, "gfsk_mod", h, lib.gfsk_mod_process, np.uint8, np.complex64)


def interp_fir_filter(taps, interpolation, max_len, fma=False):
    lib = load(fma)
    taps = _as_f32(taps)
    p = lib.malloc(taps.nbytes)
    C.memmove(p, taps.ctypes.data, taps.nbytes)
    h = C.c_void_p()
    _check(lib.interp_fir_filter_create(p, taps.shape[0], interpolation, max_len, C.byref(h)),
           "interp_fir_filter_create")
    return _Block(lib, "interp_fir_filter", h, lib.interp_fir_filter_process, np.float32, np.float32)


def frequency_modulator(sensitivity, max_len, fma=False):
    lib = load(fma)
    h = C.c_void_p()
    _check(lib.frequency_modulator_create(sensitivity, max_len, C.byref(h)), "frequency_modulator_create")
    return _Block(lib, "frequency_modulator", h, lib.frequency_modulator_process, np.float32, np.complex64)


def lpf_taps(gain, fs, cutoff, tw, fma=False):
    lib = load(fma)
    p = c_float_p()
    n = C.c_size_t()
    _check(lib.create_low_pass_filter(gain, fs, cutoff, tw, C.byref(p), C.byref(n)), "create_low_pass_filter")
    out = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
    lib.free(p)
    return out


def gaussian_taps(gain, sps, bt, n, fma=False):
    lib = load(fma)
    p = c_float_p()
    _check(lib.gaussian_taps_create(gain, sps, bt, n, C.byref(p)), "gaussian_taps_create")
    out = np.ctypeslib.as_array(p, shape=(n,)).copy()
    lib.free(p)
    return out


def fast_atan2f(y, x, fma=False):
    lib = load(fma)
    y = _as_f32(y).ravel()
    x = _as_f32(x).ravel()
    return np.array([lib.fast_atan2f(float(a), float(b)) for a, b in zip(y, x)], dtype=np.float32)


def fsk_chain(fs, baud, deviation, decimation, tw, use_dc, iq, chunk, max_len=None, stages=False, fma=False):
    """Demod chain composed block by block (reference src/dsp/fsk_demod.c:80-110); returns dict of arrays."""
    lib = load(fma)
    x = _iq_f32(iq)
    n = x.shape[0] // 2
    max_len = max_len or chunk
    soft = np.zeros(n + 16, dtype=np.float32)
    hard = np.zeros(n + 16, dtype=np.int8)
    res = {}
    ptrs = [None, None, None, None]
    if stages:
        res["lpf1"] = np.zeros(2 * n, dtype=np.float32)
        res["qd"] = np.zeros(n, dtype=np.float32)
        res["lpf2"] = np.zeros(n, dtype=np.float32)
        res["dc"] = np.zeros(n, dtype=np.float32)
        ptrs = [res[k].ctypes.data_as(C.c_void_p) for k in ("lpf1", "qd", "lpf2", "dc")]
    cnt = C.c_size_t()
    got = lib.ref_fsk_chain_run(fs, baud, deviation, decimation, tw, int(bool(use_dc)), max_len,
                                x.ctypes.data_as(C.c_void_p), n, chunk,
                                ptrs[0], ptrs[1], ptrs[2], C.byref(cnt), ptrs[3],
                                soft.ctypes.data_as(C.c_void_p), hard.ctypes.data_as(C.c_void_p), soft.shape[0])
    if got < 0:
        raise ValueError("ref_fsk_chain_run failed")
    res["soft"] = soft[:got].copy()
    res["hard"] = hard[:got].copy()
    if stages:
        res["lpf1"] = res["lpf1"].view(np.complex64)
        res["lpf2"] = res["lpf2"][:cnt.value].copy()
        res["dc"] = res["dc"][:cnt.value].copy()
    return res


def fsk_demod_run(fs, baud, deviation, decimation, tw, use_dc, iq, chunk, max_len=None, fma=False):
    lib = load(fma)
    x = _iq_f32(iq)
    n = x.shape[0] // 2
    out = np.zeros(n + 16, dtype=np.int8)
    got = lib.ref_fsk_demod_run(fs, baud, deviation, decimation, tw, int(bool(use_dc)), max_len or chunk,
                                x.ctypes.data_as(C.c_void_p), n, chunk, out.ctypes.data_as(C.c_void_p), out.shape[0])
    if got < 0:
        raise ValueError("ref_fsk_demod_run failed")
    return out[:got].copy()


def bench_fsk_demod(fs, baud, deviation, decimation, tw, use_dc, chunk, iq_channels, n_threads, passes=1, fma=False):
    """iq_channels: complex64 [channels, n]. Returns (seconds, symbols)."""
    lib = load(fma)
    a = np.ascontiguousarray(iq_channels, dtype=np.complex64)
    nch, n = a.shape
    sym = C.c_uint64()
    sec = lib.ref_bench_fsk_demod(fs, baud, deviation, decimation, tw, int(bool(use_dc)), chunk,
                                  a.ctypes.data_as(C.c_void_p), 2 * n, n, nch, n_threads, passes, C.byref(sym))
    if sec < 0:
        raise ValueError("ref_bench_fsk_demod failed")
    return sec, sym.value


def bench_gfsk_mod(sps, sensitivity, bt, byte_channels, n_threads, packets=1, fma=False):
    lib = load(fma)
    a = np.ascontiguousarray(byte_channels, dtype=np.uint8)
    nch, n = a.shape
    cnt = C.c_uint64()
    sec = lib.ref_bench_gfsk_mod(sps, sensitivity, bt, a.ctypes.data_as(C.c_void_p), n, n, nch, n_threads, packets,
                                 C.byref(cnt))
    if sec < 0:
        raise ValueError("ref_bench_gfsk_mod failed")
    return sec, cnt.value
