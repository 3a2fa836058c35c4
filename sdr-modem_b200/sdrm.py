"""ctypes binding of libsdrmodem_b200.so — the same C ABI a C host links against (include/sdrm/*.h).

There is no CPU implementation behind this module: if the shared library (and with it the CUDA kernels) is
missing, importing fails; if no GPU is present, every create call fails with the CUDA error.
Used by tests/, bench.py and __graft_entry__.py; it mirrors the reference's block API name for name.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsdrmodem_b200.so")

FLAG_FAST_FMA = 1
FLAG_SOFT_OUT = 2
MAX_IN_FLIGHT = 2
# measurement aids (host/sdrm_internal.h): results become invalid; tools/probe_*.py and bench.py --debug-no-tail only
AID_NO_CLOCK_LOOP = 1
AID_NO_TAIL = 2
AID_FETCH_BOUND_1 = 4


class FskDemodBatchConfig(C.Structure):
    _fields_ = [("n_channels", C.c_uint32),
                ("sampling_freq", C.c_uint64),
                ("baud_rate", C.c_uint32),
                ("deviation", C.c_int64),
                ("decimation", C.c_uint8),
                ("transition_width", C.c_uint32),
                ("use_dc_block", C.c_bool),
                ("max_input_buffer_length", C.c_uint32),
                ("max_symbols_per_call", C.c_uint32),
                ("flags", C.c_uint32),
                ("device", C.c_int)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("libsdrmodem_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C sdr-modem_b200`); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    lib.sdrm_version.restype = C.c_char_p
    lib.sdrm_pinned_alloc.restype = vp
    lib.sdrm_pinned_alloc.argtypes = [sz]
    lib.sdrm_pinned_free.argtypes = [vp]
    lib.sdrm_pinned_alloc_near_device.restype = vp
    lib.sdrm_pinned_alloc_near_device.argtypes = [sz, i32]
    lib.sdrm_device_numa_node.argtypes = [i32]
    lib.sdrm_measure_fp32_peak.argtypes = [i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.sdrm_probe_h2d.argtypes = [i32, vp, vp, sz, i32, C.POINTER(C.c_double)]
    lib.sdrm_bind_thread_near_device.argtypes = [i32]
    lib.sdrm_cpulist_parse_count.argtypes = [C.c_char_p]
    lib.sdrm_fsk_demod_batch_create.argtypes = [C.POINTER(FskDemodBatchConfig), C.POINTER(vp)]
    lib.sdrm_fsk_demod_batch_process.argtypes = [vp, vp, sz, sz, vp, vp, sz, vp]
    lib.sdrm_fsk_demod_batch_submit.argtypes = [vp, vp, sz, sz]
    lib.sdrm_fsk_demod_batch_submit_i16.argtypes = [vp, vp, sz, sz, C.c_float]
    lib.sdrm_fsk_demod_batch_process_device.argtypes = [vp, vp, sz, sz]
    lib.sdrm_samples_i16_to_cf32_device.argtypes = [vp, sz, vp, sz, C.c_float, sz, C.c_uint32, vp]
    lib.sdrm_samples_cf32_to_i16_device.argtypes = [vp, sz, vp, sz, C.c_float, sz, C.c_uint32, vp]
    lib.sdrm_fsk_demod_batch_fetch.argtypes = [vp, vp, vp, sz, vp]
    lib.sdrm_fsk_demod_batch_release.argtypes = [vp]
    lib.sdrm_fsk_demod_batch_device_outputs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(sz)]
    lib.sdrm_fsk_demod_batch_sync.argtypes = [vp]
    lib.sdrm_fsk_demod_batch_stream.restype = vp
    lib.sdrm_fsk_demod_batch_stream.argtypes = [vp]
    lib.sdrm_fsk_demod_batch_tail_stream.restype = vp
    lib.sdrm_fsk_demod_batch_tail_stream.argtypes = [vp]
    lib.sdrm_fsk_demod_batch_out_stream.restype = vp
    lib.sdrm_fsk_demod_batch_out_stream.argtypes = [vp]
    lib.sdrm_fsk_demod_batch_wait_outputs.argtypes = [vp, vp]
    lib.sdrm_debug_set_measurement_aid.argtypes = [vp, C.c_uint32]
    lib.sdrm_fsk_demod_batch_launch_count.restype = C.c_uint64
    lib.sdrm_fsk_demod_batch_launch_count.argtypes = [vp]
    lib.sdrm_fsk_demod_batch_last_fetch_columns.restype = C.c_size_t
    lib.sdrm_fsk_demod_batch_last_fetch_columns.argtypes = [vp]
    lib.sdrm_fsk_demod_batch_error_flags.argtypes = [vp]
    lib.sdrm_fsk_demod_batch_set_profiling.argtypes = [vp, i32]
    lib.sdrm_fsk_demod_batch_stage_times.argtypes = [vp, C.POINTER(C.c_float)]
    lib.sdrm_fsk_demod_batch_destroy.argtypes = [vp]
    lib.sdrm_fsk_demod_batch_destroy.restype = None
    lib.sdrm_nco_batch_create.argtypes = [C.c_uint32, C.c_float, C.c_uint64, C.c_uint32, i32, C.POINTER(vp)]
    lib.sdrm_nco_batch_process.argtypes = [vp, vp, vp, sz, sz, vp, sz]
    lib.sdrm_nco_batch_process_device.argtypes = [vp, vp, vp, sz, sz, vp, sz]
    lib.sdrm_nco_batch_sync.argtypes = [vp]
    lib.sdrm_nco_batch_stream.restype = vp
    lib.sdrm_nco_batch_stream.argtypes = [vp]
    lib.sdrm_nco_batch_destroy.argtypes = [vp]
    lib.sdrm_nco_batch_destroy.restype = None
    lib.sdrm_gfsk_mod_batch_create.argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_uint32, i32, C.POINTER(vp)]
    lib.sdrm_gfsk_mod_batch_process.argtypes = [vp, vp, sz, sz, vp, sz, C.POINTER(sz)]
    lib.sdrm_gfsk_mod_batch_process_device.argtypes = [vp, vp, sz, sz, vp, sz]
    lib.sdrm_gfsk_mod_batch_process_i16.argtypes = [vp, vp, sz, sz, vp, sz, C.c_float, C.POINTER(sz)]
    lib.sdrm_gfsk_mod_batch_sync.argtypes = [vp]
    lib.sdrm_gfsk_mod_batch_stream.restype = vp
    lib.sdrm_gfsk_mod_batch_stream.argtypes = [vp]
    lib.sdrm_gfsk_mod_batch_input_stream.restype = vp
    lib.sdrm_gfsk_mod_batch_input_stream.argtypes = [vp]
    lib.sdrm_gfsk_mod_batch_launch_count.restype = C.c_uint64
    lib.sdrm_gfsk_mod_batch_launch_count.argtypes = [vp]
    lib.sdrm_gfsk_mod_batch_destroy.argtypes = [vp]
    lib.sdrm_gfsk_mod_batch_destroy.restype = None
    lib.sdrm_lpf_batch_create.argtypes = [C.c_uint32, C.c_uint8, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, sz, i32,
                                          C.POINTER(vp)]
    lib.sdrm_lpf_batch_process.argtypes = [vp, vp, sz, sz, vp, sz, C.POINTER(sz)]
    lib.sdrm_lpf_batch_process_device.argtypes = [vp, vp, sz, sz, vp, sz, C.POINTER(sz)]
    lib.sdrm_lpf_batch_sync.argtypes = [vp]
    lib.sdrm_lpf_batch_stream.restype = vp
    lib.sdrm_lpf_batch_stream.argtypes = [vp]
    lib.sdrm_lpf_batch_launch_count.restype = C.c_uint64
    lib.sdrm_lpf_batch_launch_count.argtypes = [vp]
    lib.sdrm_lpf_batch_destroy.argtypes = [vp]
    lib.sdrm_lpf_batch_destroy.restype = None
    lib.sdrm_doppler_batch_create.argtypes = [C.c_uint32, vp, C.c_uint64, C.c_uint64, C.c_uint32, i32, C.POINTER(vp)]
    lib.sdrm_doppler_batch_process.argtypes = [vp, i32, vp, sz, sz, vp, sz]
    lib.sdrm_doppler_batch_process_device.argtypes = [vp, i32, vp, sz, sz, vp, sz]
    lib.sdrm_doppler_batch_sync.argtypes = [vp]
    lib.sdrm_doppler_batch_stream.restype = vp
    lib.sdrm_doppler_batch_stream.argtypes = [vp]
    lib.sdrm_doppler_batch_destroy.argtypes = [vp]
    lib.sdrm_doppler_batch_destroy.restype = None
    # multi-device entry points (include/sdrm/sdrm_multi.h)
    lib.sdrm_fsk_demod_multi_create.argtypes = [C.POINTER(FskDemodBatchConfig), C.POINTER(i32), C.c_uint32, C.POINTER(vp)]
    lib.sdrm_fsk_demod_multi_submit.argtypes = [vp, vp, sz, sz]
    lib.sdrm_fsk_demod_multi_submit_i16.argtypes = [vp, vp, sz, sz, C.c_float]
    lib.sdrm_fsk_demod_multi_fetch.argtypes = [vp, vp, vp, sz, vp]
    lib.sdrm_fsk_demod_multi_process.argtypes = [vp, vp, sz, sz, vp, vp, sz, vp]
    lib.sdrm_fsk_demod_multi_sync.argtypes = [vp]
    lib.sdrm_fsk_demod_multi_device_count.restype = C.c_uint32
    lib.sdrm_fsk_demod_multi_device_count.argtypes = [vp]
    lib.sdrm_fsk_demod_multi_shard.argtypes = [vp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(i32)]
    lib.sdrm_fsk_demod_multi_batch.restype = vp
    lib.sdrm_fsk_demod_multi_batch.argtypes = [vp, C.c_uint32]
    lib.sdrm_fsk_demod_multi_launch_count.restype = C.c_uint64
    lib.sdrm_fsk_demod_multi_launch_count.argtypes = [vp]
    lib.sdrm_fsk_demod_multi_error_flags.argtypes = [vp]
    lib.sdrm_fsk_demod_multi_destroy.argtypes = [vp]
    lib.sdrm_fsk_demod_multi_destroy.restype = None
    # reference-named single-channel API
    lib.fsk_demod_create.argtypes = [C.c_uint64, C.c_uint32, C.c_int64, C.c_uint8, C.c_uint32, C.c_bool, C.c_uint32,
                                     C.POINTER(vp)]
    lib.fsk_demod_process.argtypes = [vp, sz, C.POINTER(vp), C.POINTER(sz), vp]
    lib.fsk_demod_process.restype = None
    lib.fsk_demod_destroy.argtypes = [vp]
    lib.fsk_demod_destroy.restype = None
    return lib


lib = _load()


def version():
    return lib.sdrm_version().decode()


class SdrmError(RuntimeError):
    pass


def _check(code, what):
    if code != 0:
        raise SdrmError("%s failed with %d" % (what, code))


def bind_thread_near_device(device):
    """Move the calling thread onto the CPUs local to CUDA device `device`, so that pinned buffers allocated next are on its
    socket. Returns the size of the new CPU mask, 0 if nothing changed."""
    n = lib.sdrm_bind_thread_near_device(int(device))
    if n < 0:
        raise SdrmError("sdrm_bind_thread_near_device(%d) failed with %d" % (device, n))
    return n


def measure_fp32_peak(device=-1):
    """(FMA peak, exact-mode FFMA2-pair rate) of the device's FP32 pipe in algorithmic TFLOP/s, measured now."""
    fma, pair = C.c_double(), C.c_double()
    _check(lib.sdrm_measure_fp32_peak(int(device), C.byref(fma), C.byref(pair)), "sdrm_measure_fp32_peak")
    return fma.value, pair.value


def probe_h2d(device, host_ptr, d_ptr, nbytes, repeats):
    """seconds of device time for `repeats` plain pinned host -> device copies of nbytes"""
    sec = C.c_double()
    _check(lib.sdrm_probe_h2d(int(device), C.c_void_p(host_ptr), C.c_void_p(d_ptr), nbytes, repeats, C.byref(sec)),
           "sdrm_probe_h2d")
    return sec.value


class PinnedArray:
    """numpy view over pinned host memory from sdrm_pinned_alloc (device given: on that device's NUMA node)."""

    def __init__(self, shape, dtype, device=-1):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = lib.sdrm_pinned_alloc_near_device(self.nbytes, int(device))
        if not self.ptr:
            raise SdrmError("sdrm_pinned_alloc(%d) failed" % self.nbytes)
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def close(self):
        if self.ptr:
            self.array = None
            lib.sdrm_pinned_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FskDemodBatch:
    """N x fsk_demod (reference src/dsp/fsk_demod.c) as one batched GPU session."""

    def __init__(self, n_channels, sampling_freq, baud_rate, deviation, decimation, transition_width, use_dc_block,
                 max_input_buffer_length, max_symbols_per_call=0, fast=False, soft=False, device=-1, measurement_aid=0):
        cfg = FskDemodBatchConfig(n_channels, sampling_freq, baud_rate, deviation, decimation, transition_width,
                                  bool(use_dc_block), max_input_buffer_length, max_symbols_per_call,
                                  (FLAG_FAST_FMA if fast else 0) | (FLAG_SOFT_OUT if soft else 0), device)
        self.handle = C.c_void_p()
        self.n_channels = n_channels
        self.max_len = max_input_buffer_length
        self.capacity = max_symbols_per_call or max_input_buffer_length
        self.soft = soft
        _check(lib.sdrm_fsk_demod_batch_create(C.byref(cfg), C.byref(self.handle)), "sdrm_fsk_demod_batch_create")
        if measurement_aid:
            _check(lib.sdrm_debug_set_measurement_aid(self.handle, measurement_aid), "sdrm_debug_set_measurement_aid")

    # -- host buffers -------------------------------------------------------------------------------------------
    def process(self, iq):
        """iq: complex64 [channels, n]. Returns (hard int8 [channels, cap], lens uint32 [channels], soft or None)."""
        self.submit(iq)
        return self.fetch()

    def submit(self, iq):
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        assert iq.ndim == 2 and iq.shape[0] == self.n_channels
        self._keep = iq
        _check(lib.sdrm_fsk_demod_batch_submit(self.handle, iq.ctypes.data_as(C.c_void_p), iq.shape[1], iq.shape[1]),
               "sdrm_fsk_demod_batch_submit")

    def submit_ptr(self, host_ptr, in_stride, n):
        _check(lib.sdrm_fsk_demod_batch_submit(self.handle, host_ptr, in_stride, n), "sdrm_fsk_demod_batch_submit")

    def submit_i16(self, iq16, scalar=2048.0):
        """iq16: int16 [channels, n, 2] (I, Q) as the SDR delivers it; converted on the device (plutosdr.c:129)."""
        iq16 = np.ascontiguousarray(iq16, dtype=np.int16)
        assert iq16.ndim == 3 and iq16.shape[0] == self.n_channels and iq16.shape[2] == 2
        self._keep = iq16
        _check(lib.sdrm_fsk_demod_batch_submit_i16(self.handle, iq16.ctypes.data_as(C.c_void_p), iq16.shape[1], iq16.shape[1],
                                                   scalar), "sdrm_fsk_demod_batch_submit_i16")

    def submit_i16_ptr(self, host_ptr, in_stride, n, scalar=2048.0):
        _check(lib.sdrm_fsk_demod_batch_submit_i16(self.handle, host_ptr, in_stride, n, scalar),
               "sdrm_fsk_demod_batch_submit_i16")

    def fetch(self, hard=None, lens=None, soft=None):
        cap = self.capacity
        if hard is None:
            hard = np.zeros((self.n_channels, cap), dtype=np.int8)
        if lens is None:
            lens = np.zeros(self.n_channels, dtype=np.uint32)
        if soft is None and self.soft:
            soft = np.zeros((self.n_channels, cap), dtype=np.float32)
        _check(lib.sdrm_fsk_demod_batch_fetch(self.handle, hard.ctypes.data_as(C.c_void_p),
                                              soft.ctypes.data_as(C.c_void_p) if soft is not None else None,
                                              hard.shape[1], lens.ctypes.data_as(C.c_void_p)),
               "sdrm_fsk_demod_batch_fetch")
        return hard, lens, soft

    def fetch_ptr(self, hard_ptr, out_stride, lens_ptr):
        _check(lib.sdrm_fsk_demod_batch_fetch(self.handle, hard_ptr, None, out_stride, lens_ptr),
               "sdrm_fsk_demod_batch_fetch")

    # -- device-resident ----------------------------------------------------------------------------------------
    def process_device(self, d_ptr, in_stride, n):
        _check(lib.sdrm_fsk_demod_batch_process_device(self.handle, C.c_void_p(d_ptr), in_stride, n),
               "sdrm_fsk_demod_batch_process_device")

    def release(self):
        _check(lib.sdrm_fsk_demod_batch_release(self.handle), "sdrm_fsk_demod_batch_release")

    def device_outputs(self):
        d_out, d_len, stride = C.c_void_p(), C.c_void_p(), C.c_size_t()
        _check(lib.sdrm_fsk_demod_batch_device_outputs(self.handle, C.byref(d_out), C.byref(d_len), C.byref(stride)),
               "sdrm_fsk_demod_batch_device_outputs")
        return d_out.value, d_len.value, stride.value

    def sync(self):
        _check(lib.sdrm_fsk_demod_batch_sync(self.handle), "sdrm_fsk_demod_batch_sync")

    @property
    def stream(self):
        return lib.sdrm_fsk_demod_batch_stream(self.handle)

    @property
    def tail_stream(self):
        return lib.sdrm_fsk_demod_batch_tail_stream(self.handle)

    @property
    def out_stream(self):
        return lib.sdrm_fsk_demod_batch_out_stream(self.handle)

    def wait_outputs(self, stream):
        _check(lib.sdrm_fsk_demod_batch_wait_outputs(self.handle, C.c_void_p(stream)), "sdrm_fsk_demod_batch_wait_outputs")

    @property
    def launch_count(self):
        return lib.sdrm_fsk_demod_batch_launch_count(self.handle)

    def last_fetch_columns(self):
        return lib.sdrm_fsk_demod_batch_last_fetch_columns(self.handle)

    def set_profiling(self, enabled=True):
        _check(lib.sdrm_fsk_demod_batch_set_profiling(self.handle, int(enabled)), "sdrm_fsk_demod_batch_set_profiling")

    def stage_times(self):
        """ms of (lpf1+quad kernel, lpf1 history + lpf2, fused dc+clock tail kernel, whole call) for the latest call."""
        ms = (C.c_float * 4)()
        _check(lib.sdrm_fsk_demod_batch_stage_times(self.handle, ms), "sdrm_fsk_demod_batch_stage_times")
        return list(ms)

    def error_flags(self):
        return lib.sdrm_fsk_demod_batch_error_flags(self.handle)

    def run_stream(self, iq, chunk):
        """Feeds complex64 [channels, n] in `chunk`-sized calls; returns per-channel (hard, soft) concatenations."""
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        hard_parts = [[] for _ in range(self.n_channels)]
        soft_parts = [[] for _ in range(self.n_channels)]
        for off in range(0, iq.shape[1], chunk):
            hard, lens, soft = self.process(iq[:, off:off + chunk])
            for c in range(self.n_channels):
                hard_parts[c].append(hard[c, :lens[c]].copy())
                if soft is not None:
                    soft_parts[c].append(soft[c, :lens[c]].copy())
        hard_out = [np.concatenate(p) if p else np.zeros(0, np.int8) for p in hard_parts]
        soft_out = [np.concatenate(p) if p else np.zeros(0, np.float32) for p in soft_parts] if self.soft else None
        return hard_out, soft_out

    def close(self):
        if self.handle:
            lib.sdrm_fsk_demod_batch_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FskDemodMulti:
    """N x fsk_demod partitioned over a list of CUDA devices by channel range (include/sdrm/sdrm_multi.h)."""

    def __init__(self, devices, n_channels, sampling_freq, baud_rate, deviation, decimation, transition_width, use_dc_block,
                 max_input_buffer_length, max_symbols_per_call=0, fast=False, soft=False):
        cfg = FskDemodBatchConfig(n_channels, sampling_freq, baud_rate, deviation, decimation, transition_width,
                                  bool(use_dc_block), max_input_buffer_length, max_symbols_per_call,
                                  (FLAG_FAST_FMA if fast else 0) | (FLAG_SOFT_OUT if soft else 0), -1)
        self.handle = C.c_void_p()
        self.n_channels = n_channels
        self.capacity = max_symbols_per_call or max_input_buffer_length
        self.soft = soft
        dev = (C.c_int * len(devices))(*devices)
        _check(lib.sdrm_fsk_demod_multi_create(C.byref(cfg), dev, len(devices), C.byref(self.handle)),
               "sdrm_fsk_demod_multi_create")

    def submit(self, iq):
        assert iq.ndim == 2 and iq.shape[0] == self.n_channels and iq.dtype == np.complex64 and iq.flags.c_contiguous
        self._keep = iq
        _check(lib.sdrm_fsk_demod_multi_submit(self.handle, iq.ctypes.data_as(C.c_void_p), iq.shape[1], iq.shape[1]),
               "sdrm_fsk_demod_multi_submit")

    def submit_ptr(self, host_ptr, in_stride, n):
        _check(lib.sdrm_fsk_demod_multi_submit(self.handle, host_ptr, in_stride, n), "sdrm_fsk_demod_multi_submit")

    def submit_i16_ptr(self, host_ptr, in_stride, n, scalar=2048.0):
        _check(lib.sdrm_fsk_demod_multi_submit_i16(self.handle, host_ptr, in_stride, n, scalar),
               "sdrm_fsk_demod_multi_submit_i16")

    def fetch(self):
        hard = np.zeros((self.n_channels, self.capacity), dtype=np.int8)
        lens = np.zeros(self.n_channels, dtype=np.uint32)
        soft = np.zeros((self.n_channels, self.capacity), dtype=np.float32) if self.soft else None
        _check(lib.sdrm_fsk_demod_multi_fetch(self.handle, hard.ctypes.data_as(C.c_void_p),
                                              soft.ctypes.data_as(C.c_void_p) if soft is not None else None,
                                              hard.shape[1], lens.ctypes.data_as(C.c_void_p)), "sdrm_fsk_demod_multi_fetch")
        return hard, lens, soft

    def fetch_ptr(self, hard_ptr, out_stride, lens_ptr):
        _check(lib.sdrm_fsk_demod_multi_fetch(self.handle, hard_ptr, None, out_stride, lens_ptr), "sdrm_fsk_demod_multi_fetch")

    def process(self, iq):
        self.submit(np.ascontiguousarray(iq, dtype=np.complex64))
        return self.fetch()

    def shards(self):
        out = []
        for g in range(lib.sdrm_fsk_demod_multi_device_count(self.handle)):
            first, count, dev = C.c_uint32(), C.c_uint32(), C.c_int()
            _check(lib.sdrm_fsk_demod_multi_shard(self.handle, g, C.byref(first), C.byref(count), C.byref(dev)), "shard")
            out.append((first.value, count.value, dev.value))
        return out

    def sync(self):
        _check(lib.sdrm_fsk_demod_multi_sync(self.handle), "sdrm_fsk_demod_multi_sync")

    @property
    def launch_count(self):
        return lib.sdrm_fsk_demod_multi_launch_count(self.handle)

    def error_flags(self):
        return lib.sdrm_fsk_demod_multi_error_flags(self.handle)

    def close(self):
        if self.handle:
            lib.sdrm_fsk_demod_multi_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FskDemod:
    """fsk_demod_create / _process / _destroy, the reference's single-channel entry points."""

    def __init__(self, sampling_freq, baud_rate, deviation, decimation, transition_width, use_dc_block, max_len):
        self.handle = C.c_void_p()
        code = lib.fsk_demod_create(sampling_freq, baud_rate, deviation, decimation, transition_width,
                                    bool(use_dc_block), max_len, C.byref(self.handle))
        if code != 0:
            self.handle = C.c_void_p()
            raise SdrmError("fsk_demod_create failed with %d" % code)

    def process(self, iq):
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        out, n = C.c_void_p(), C.c_size_t()
        lib.fsk_demod_process(iq.ctypes.data_as(C.c_void_p), iq.shape[0], C.byref(out), C.byref(n), self.handle)
        if not out.value or n.value == 0:
            return np.zeros(0, dtype=np.int8)
        return np.frombuffer((C.c_char * n.value).from_address(out.value), dtype=np.int8).copy()

    def run(self, iq, chunk):
        parts = [self.process(iq[o:o + chunk]) for o in range(0, len(iq), chunk)]
        return np.concatenate(parts) if parts else np.zeros(0, np.int8)

    def close(self):
        if self.handle:
            lib.fsk_demod_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DopplerChannel(C.Structure):
    _fields_ = [("latitude", C.c_double), ("longitude", C.c_double), ("altitude", C.c_double),
                ("constant_offset", C.c_int64), ("start_time_seconds", C.c_int64), ("tle", (C.c_char * 80) * 3)]


def doppler_channel(lat, lon, alt, constant_offset, start_time, tle_lines):
    ch = DopplerChannel(lat, lon, alt, constant_offset, start_time)
    for i, line in enumerate(tle_lines):
        raw = line.encode("ascii")[:79]
        C.memmove(C.addressof(ch.tle[i]), raw + b"\0", len(raw) + 1)
    return ch


class DopplerBatch:
    """N x doppler (reference src/dsp/doppler.c): host SGP4 schedule + GPU mixer. direction +1 rx, -1 tx."""

    def __init__(self, channels, sampling_freq, center_freq, max_len, device=-1):
        self.handle = C.c_void_p()
        self.n_channels = len(channels)
        arr = (DopplerChannel * len(channels))(*channels)
        _check(lib.sdrm_doppler_batch_create(len(channels), arr, sampling_freq, center_freq, max_len, device,
                                             C.byref(self.handle)), "sdrm_doppler_batch_create")

    def process(self, iq, direction=1):
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        out = np.zeros_like(iq)
        _check(lib.sdrm_doppler_batch_process(self.handle, direction, iq.ctypes.data_as(C.c_void_p), iq.shape[1], iq.shape[1],
                                              out.ctypes.data_as(C.c_void_p), iq.shape[1]), "sdrm_doppler_batch_process")
        return out

    def process_device(self, d_in, in_stride, n, d_out, out_stride, direction=1):
        _check(lib.sdrm_doppler_batch_process_device(self.handle, direction, C.c_void_p(d_in), in_stride, n, C.c_void_p(d_out),
                                                     out_stride), "sdrm_doppler_batch_process_device")

    def sync(self):
        _check(lib.sdrm_doppler_batch_sync(self.handle), "sdrm_doppler_batch_sync")

    @property
    def stream(self):
        return lib.sdrm_doppler_batch_stream(self.handle)

    def close(self):
        if self.handle:
            lib.sdrm_doppler_batch_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NcoBatch:
    """N x sig_source (reference src/dsp/sig_source.c): per-channel NCO with carried float phase."""

    def __init__(self, n_channels, amplitude, sampling_freq, max_len, device=-1):
        self.handle = C.c_void_p()
        self.n_channels = n_channels
        _check(lib.sdrm_nco_batch_create(n_channels, amplitude, sampling_freq, max_len, device, C.byref(self.handle)),
               "sdrm_nco_batch_create")

    def multiply(self, freq_hz, iq):
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        freq = np.ascontiguousarray(freq_hz, dtype=np.int64)
        out = np.zeros_like(iq)
        _check(lib.sdrm_nco_batch_process(self.handle, freq.ctypes.data_as(C.c_void_p), iq.ctypes.data_as(C.c_void_p),
                                          iq.shape[1], iq.shape[1], out.ctypes.data_as(C.c_void_p), iq.shape[1]),
               "sdrm_nco_batch_process")
        return out

    def generate(self, freq_hz, n):
        freq = np.ascontiguousarray(freq_hz, dtype=np.int64)
        out = np.zeros((self.n_channels, n), dtype=np.complex64)
        _check(lib.sdrm_nco_batch_process(self.handle, freq.ctypes.data_as(C.c_void_p), None, 0, n,
                                          out.ctypes.data_as(C.c_void_p), n), "sdrm_nco_batch_process")
        return out

    def process_device(self, freq_hz, d_in, in_stride, n, d_out, out_stride):
        freq = np.ascontiguousarray(freq_hz, dtype=np.int64)
        _check(lib.sdrm_nco_batch_process_device(self.handle, freq.ctypes.data_as(C.c_void_p), C.c_void_p(d_in), in_stride, n,
                                                 C.c_void_p(d_out), out_stride), "sdrm_nco_batch_process_device")

    def sync(self):
        _check(lib.sdrm_nco_batch_sync(self.handle), "sdrm_nco_batch_sync")

    @property
    def stream(self):
        return lib.sdrm_nco_batch_stream(self.handle)

    def close(self):
        if self.handle:
            lib.sdrm_nco_batch_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LpfBatch:
    """N x lpf (reference src/dsp/lpf.c): one design, N streams, complex or real samples."""

    def __init__(self, n_channels, decimation, sampling_freq, cutoff_freq, transition_width, max_len, complex_samples=True,
                 device=-1):
        self.handle = C.c_void_p()
        self.n_channels = n_channels
        self.dtype = np.complex64 if complex_samples else np.float32
        _check(lib.sdrm_lpf_batch_create(n_channels, decimation, sampling_freq, cutoff_freq, transition_width, max_len,
                                         8 if complex_samples else 4, device, C.byref(self.handle)), "sdrm_lpf_batch_create")

    def process(self, x):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        assert x.ndim == 2 and x.shape[0] == self.n_channels
        out = np.zeros((self.n_channels, max(x.shape[1], 1)), dtype=self.dtype)
        produced = C.c_size_t()
        _check(lib.sdrm_lpf_batch_process(self.handle, x.ctypes.data_as(C.c_void_p), x.shape[1], x.shape[1],
                                          out.ctypes.data_as(C.c_void_p), out.shape[1], C.byref(produced)),
               "sdrm_lpf_batch_process")
        return out[:, :produced.value]

    def process_device(self, d_in, in_stride, n, d_out, out_stride):
        produced = C.c_size_t()
        _check(lib.sdrm_lpf_batch_process_device(self.handle, C.c_void_p(d_in), in_stride, n, C.c_void_p(d_out), out_stride,
                                                 C.byref(produced)), "sdrm_lpf_batch_process_device")
        return produced.value

    def sync(self):
        _check(lib.sdrm_lpf_batch_sync(self.handle), "sdrm_lpf_batch_sync")

    @property
    def stream(self):
        return lib.sdrm_lpf_batch_stream(self.handle)

    def close(self):
        if self.handle:
            lib.sdrm_lpf_batch_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GfskModBatch:
    """N x gfsk_mod (reference src/dsp/gfsk_mod.c): bytes in, cf32 out."""

    def __init__(self, n_channels, samples_per_symbol, sensitivity, bt, max_bytes, device=-1):
        self.handle = C.c_void_p()
        self.n_channels = n_channels
        self.interpolation = int(samples_per_symbol)
        _check(lib.sdrm_gfsk_mod_batch_create(n_channels, samples_per_symbol, sensitivity, bt, max_bytes, device,
                                              C.byref(self.handle)), "sdrm_gfsk_mod_batch_create")

    def process(self, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        assert data.ndim == 2 and data.shape[0] == self.n_channels
        n_out = data.shape[1] * 8 * self.interpolation
        out = np.zeros((self.n_channels, max(n_out, 1)), dtype=np.complex64)
        produced = C.c_size_t()
        _check(lib.sdrm_gfsk_mod_batch_process(self.handle, data.ctypes.data_as(C.c_void_p), data.shape[1], data.shape[1],
                                               out.ctypes.data_as(C.c_void_p), out.shape[1], C.byref(produced)),
               "sdrm_gfsk_mod_batch_process")
        return out[:, :produced.value]

    def process_i16(self, data, scalar=32768.0):
        """bytes in, int16 (I, Q) pairs out [channels, n, 2]: the PlutoSDR egress format (plutosdr.c:83)"""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        assert data.ndim == 2 and data.shape[0] == self.n_channels
        n_out = data.shape[1] * 8 * self.interpolation
        out = np.zeros((self.n_channels, max(n_out, 1), 2), dtype=np.int16)
        produced = C.c_size_t()
        _check(lib.sdrm_gfsk_mod_batch_process_i16(self.handle, data.ctypes.data_as(C.c_void_p), data.shape[1], data.shape[1],
                                                   out.ctypes.data_as(C.c_void_p), out.shape[1], scalar, C.byref(produced)),
               "sdrm_gfsk_mod_batch_process_i16")
        return out[:, :produced.value]

    def process_device(self, d_in, in_stride, n_bytes, d_out, out_stride):
        _check(lib.sdrm_gfsk_mod_batch_process_device(self.handle, C.c_void_p(d_in), in_stride, n_bytes, C.c_void_p(d_out),
                                                      out_stride), "sdrm_gfsk_mod_batch_process_device")

    def sync(self):
        _check(lib.sdrm_gfsk_mod_batch_sync(self.handle), "sdrm_gfsk_mod_batch_sync")

    @property
    def stream(self):
        """results are complete on this stream"""
        return lib.sdrm_gfsk_mod_batch_stream(self.handle)

    @property
    def input_stream(self):
        """inputs are consumed on this stream"""
        return lib.sdrm_gfsk_mod_batch_input_stream(self.handle)

    @property
    def launch_count(self):
        return lib.sdrm_gfsk_mod_batch_launch_count(self.handle)

    def close(self):
        if self.handle:
            lib.sdrm_gfsk_mod_batch_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
