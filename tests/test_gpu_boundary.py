"""GPU tests of boundary details settled in round 2: return conventions, the public struct
layout of fir_filter, flag validation, device-resident consumers, very wide batches, failing clients of the batched dispatcher."""
import ctypes as C
import socket

import numpy as np
import pytest

import workloads
from conftest import FSK_GOLDENS, golden_array, same_bits
from test_gpu_blocks import SZ, VP, make_clock

pytestmark = pytest.mark.gpu


def test_clock_mm_returns_null_while_collecting_its_first_eight_samples(sdrm, port):
    """reference src/dsp/clock_recovery_mm.c:94-99: history + input shorter than the interpolator's 8 taps => *output = NULL,
    *output_len = 0; from then on the handle's buffer, also for calls that produce nothing"""
    lib = sdrm.lib
    c = make_clock(lib, 10.0, 0.3, 0.5, 0.0625, 0.01, 4096)
    x = np.sin(np.arange(300) * 0.31).astype(np.float32)

    def raw(part):
        part = np.ascontiguousarray(part, dtype=np.float32)
        out, n = VP(), SZ()
        c.proc(part.ctypes.data_as(VP), len(part), C.byref(out), C.byref(n), c.h)
        return out.value, n.value

    assert raw(x[:3]) == (None, 0)
    assert raw(x[3:7]) == (None, 0)            # 7 samples buffered: still collecting
    ptr, n = raw(x[7:9])                         # 9 samples: the loop runs
    assert ptr is not None
    ptr2, n2 = raw(x[9:9])                       # an empty call afterwards still returns the handle's buffer
    assert ptr2 is not None
    c.close()
    # and the symbols are the oracle's for the same call sizes, including the reference's end-of-call rollback: the step of the
    # one symbol of call 3 went past the end of the buffer, so nothing was consumed and the empty call 4 emits it again
    # (clock_recovery_mm.c:127-133)
    c = make_clock(lib, 10.0, 0.3, 0.5, 0.0625, 0.01, 4096)
    o = port.ClockMm(10.0, 0.3, 0.5, 0.0625, 0.01, 4096)
    counts = []
    for lo, hi in ((0, 3), (3, 7), (7, 9), (9, 9), (9, 300)):
        got, want = c.process(x[lo:hi]), o.process(x[lo:hi])
        assert same_bits(got, want)
        counts.append(len(got))
    assert counts[:4] == [0, 0, 1, 1]
    c.close()


class FirFilterPublic(C.Structure):
    """reference src/dsp/fir_filter.h:9-27"""
    _fields_ = [("decimation", C.c_uint8), ("taps", C.POINTER(C.POINTER(C.c_float))), ("aligned_taps_len", SZ), ("alignment", SZ),
                ("taps_len", SZ), ("original_taps", C.POINTER(C.c_float)), ("working_buffer", VP), ("history_offset", SZ),
                ("working_len_total", SZ), ("volk_output", VP), ("max_input_buffer_length", SZ), ("output", VP),
                ("output_len", SZ), ("num_bytes", SZ)]


def test_fir_filter_handle_has_the_reference_public_layout(sdrm):
    lib = sdrm.lib
    libc = C.CDLL(None)
    libc.malloc.restype = VP
    libc.malloc.argtypes = [SZ]
    taps = np.array([0.1, 0.2, 0.3, 0.4, 0.5], dtype=np.float32)
    p = libc.malloc(taps.nbytes)
    C.memmove(p, taps.ctypes.data, taps.nbytes)
    lib.fir_filter_create.argtypes = [C.c_uint8, VP, SZ, SZ, SZ, C.POINTER(VP)]
    h = VP()
    assert lib.fir_filter_create(2, p, 5, 1000, 4, C.byref(h)) == 0
    f = C.cast(h, C.POINTER(FirFilterPublic)).contents
    assert f.decimation == 2 and f.taps_len == 5 and f.num_bytes == 4 and f.max_input_buffer_length == 1000
    assert f.history_offset == 4 and f.working_len_total == 1004 and f.output_len == 501 and f.alignment == 16
    assert C.cast(f.original_taps, VP).value == p
    assert f.aligned_taps_len == 1 and [f.taps[0][j] for j in range(5)] == [np.float32(v) for v in taps[::-1]]
    assert bool(f.output)
    lib.fir_filter_process.argtypes = [VP, SZ, C.POINTER(VP), C.POINTER(SZ), VP]
    lib.fir_filter_process.restype = None
    x = np.arange(10, dtype=np.float32)
    out, n = VP(), SZ()
    lib.fir_filter_process(x.ctypes.data_as(VP), 10, C.byref(out), C.byref(n), h)
    assert out.value == f.output and n.value == 5  # the process call returns the buffer the public field names
    lib.fir_filter_destroy.argtypes = [VP]
    lib.fir_filter_destroy(h)


def test_unknown_flag_bits_are_rejected(sdrm):
    cfg = sdrm.FskDemodBatchConfig(1, 48000, 4800, 5000, 2, 2000, True, 4096, 0, 0x40000000, -1)
    h = C.c_void_p()
    assert sdrm.lib.sdrm_fsk_demod_batch_create(C.byref(cfg), C.byref(h)) == -1
    cfg.flags = sdrm.FLAG_SOFT_OUT | 4
    assert sdrm.lib.sdrm_fsk_demod_batch_create(C.byref(cfg), C.byref(h)) == -1


def test_fetch_with_soft_pointer_but_no_soft_flag_leaves_the_call_unfetched(sdrm):
    _, _, args = FSK_GOLDENS["lucky7"]
    iq = golden_array("lucky7.expected.cf32", np.complex64)[None, :4096]
    b = sdrm.FskDemodBatch(1, *args, 4096)  # no soft output
    b.submit(np.ascontiguousarray(iq))
    hard = np.zeros((1, 4096), np.int8)
    soft = np.zeros((1, 4096), np.float32)
    lens = np.zeros(1, np.uint32)
    code = sdrm.lib.sdrm_fsk_demod_batch_fetch(b.handle, hard.ctypes.data_as(C.c_void_p), soft.ctypes.data_as(C.c_void_p), 4096,
                                              lens.ctypes.data_as(C.c_void_p))
    assert code == -1 and not hard.any()      # nothing was copied ...
    got, lens, _ = b.fetch()                  # ... and the call can still be fetched properly
    assert lens[0] > 300 and got[0, :lens[0]].any()
    b.close()


def test_fetch_moves_only_what_a_call_can_have_produced_and_survives_a_wrong_bound(sdrm, port):
    """A handle created for 2016000-sample buffers (what test/perf_fsk_modem.c does) and fed 4096 samples copies a bounded
    number of columns back, not its whole capacity; the counts that come back with them are checked against the bound, and when
    they exceed it (forced here through the test aid) the rest follows in a second pass. Both ways the symbols are the oracle's,
    and nothing outside the produced prefix of the caller's buffer is touched."""
    _, _, args = FSK_GOLDENS["lucky7"]
    iq = golden_array("lucky7.expected.cf32", np.complex64)
    chunk = 4096
    o_hard, o_soft = port.FskDemod(*args, 2016000).run(iq, chunk)
    for aid in (0, sdrm.AID_FETCH_BOUND_1):
        b = sdrm.FskDemodBatch(3, *args, 2016000, soft=True, measurement_aid=aid)
        hard = np.full((3, 2016000), 77, np.int8)
        soft = np.full((3, 2016000), 7.0, np.float32)
        lens = np.zeros(3, np.uint32)
        parts = [[] for _ in range(3)]
        soft_parts = [[] for _ in range(3)]
        most = 0
        for o in range(0, len(iq), chunk):
            block = np.ascontiguousarray(np.broadcast_to(iq[o:o + chunk], (3, len(iq[o:o + chunk]))))
            b.submit(block)
            b.fetch(hard=hard, lens=lens, soft=soft)
            for c in range(3):
                parts[c].append(hard[c, :lens[c]].copy())
                soft_parts[c].append(soft[c, :lens[c]].copy())
            most = max(most, int(lens.max()))
        assert b.error_flags() == 0
        b.close()
        for c in range(3):
            assert same_bits(np.concatenate(parts[c]), o_hard) and same_bits(np.concatenate(soft_parts[c]), o_soft), (aid, c)
        # 4096 samples / decimation 2 / 5 samples per symbol: about 410 symbols; the bound adds a few columns, not megabytes
        assert 300 < most < 600
        assert (hard[:, 1024:] == 77).all() and (soft[:, 1024:] == 7.0).all()


class RawCuda:
    """a device pointer as something torch.as_tensor can view without copying"""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


def test_device_resident_consumer_orders_its_stream_behind_the_results(sdrm, port):
    """sdrm_fsk_demod_batch_device_outputs + sdrm_fsk_demod_batch_wait_outputs: a consumer that never leaves the device reads
    the results on its own stream, ordered behind the call's tail by the library's event"""
    import torch
    shape = workloads.C2_PARITY
    n_ch, chunk, cap = 4, 8192, 1024
    iq = workloads.gfsk_channels(n_ch, chunk, shape, seed=51)
    d_iq = iq.cuda()
    b = sdrm.FskDemodBatch(n_ch, *shape.create_args, chunk, max_symbols_per_call=cap)
    consumer = torch.cuda.Stream()
    b.process_device(d_iq.data_ptr(), chunk, chunk)
    b.wait_outputs(consumer.cuda_stream)
    d_out, d_len, stride = b.device_outputs()
    assert stride >= cap
    with torch.cuda.stream(consumer):
        lens = torch.as_tensor(RawCuda(d_len, (n_ch,), "<u4"), device="cuda").clone()
        hard = torch.as_tensor(RawCuda(d_out, (n_ch, stride), "|i1"), device="cuda").clone()
    consumer.synchronize()
    b.release()
    lens = lens.cpu().numpy().astype(np.int64)
    hard = hard.cpu().numpy()
    assert b.error_flags() == 0
    b.close()
    for c in range(n_ch):
        want, _ = port.FskDemod(*shape.create_args, chunk).run(iq[c].numpy(), chunk)
        assert lens[c] == len(want) and same_bits(hard[c, :lens[c]], want)


def test_more_than_65535_channels(sdrm, port):
    """kernels that index rows with gridDim.y (history update, int16 conversion ...) loop over the rows beyond its limit"""
    shape = workloads.PERF_SHAPE
    n_ch, chunk = 65600, 512
    base = workloads.gfsk_channels(4, 2 * chunk, shape, seed=61).numpy()
    order = np.arange(n_ch) % 4
    b = sdrm.FskDemodBatch(n_ch, *shape.create_args, chunk, max_symbols_per_call=128)
    got = [[] for _ in range(4)]
    probe = [0, 1, 2, 3, 65532, 65533, 65534, 65535, 65536, 65537, 65598, 65599]
    res = {c: [] for c in probe}
    for k in range(2):
        # int16 ingest exercises the conversion kernel's row loop as well
        part = np.ascontiguousarray(base[order, k * chunk:(k + 1) * chunk])
        q = np.clip(np.rint(part.view(np.float32).reshape(n_ch, chunk, 2) * 1500.0), -2048, 2047).astype(np.int16)
        b.submit_i16(q)
        hard, lens, _ = b.fetch()
        for c in probe:
            res[c].append(hard[c, :lens[c]].copy())
    assert b.error_flags() == 0
    b.close()
    for c in probe:
        x = (np.clip(np.rint(base[order[c]].view(np.float32) * 1500.0), -2048, 2047).astype(np.int16).astype(np.float32)
             / np.float32(2048.0)).view(np.complex64)
        want, _ = port.FskDemod(*shape.create_args, chunk).run(x, chunk)
        assert same_bits(np.concatenate(res[c]), want), "channel %d" % c


def test_rx_group_stops_serving_a_client_whose_socket_fails(sdrm, port):
    """the reference's dsp_worker ends on the first failed write (src/dsp_worker.c:93-103); in a group the other sessions go on"""
    from test_gpu_worker import RxSession, group_config, setup_group
    lib = sdrm.lib
    setup_group(lib)
    lib.sdrm_rx_group_sessions_failed.argtypes = [VP]
    lib.sdrm_rx_group_sessions_failed.restype = C.c_uint32
    lib.sdrm_rx_group_failed.argtypes = [VP]
    _, _, args = FSK_GOLDENS["lucky7"]
    iq = golden_array("lucky7.expected.cf32", np.complex64)[:40960]
    good_a, good_b = socket.socketpair()
    dead_a, dead_b = socket.socketpair()
    dead_b.close()  # writes to dead_a now fail with EPIPE
    import signal
    old = signal.signal(signal.SIGPIPE, signal.SIG_IGN)
    try:
        sessions = (RxSession * 2)()
        sessions[0].id, sessions[0].client_socket, sessions[0].has_doppler = 0, dead_a.fileno(), False
        sessions[1].id, sessions[1].client_socket, sessions[1].has_doppler = 1, good_a.fileno(), False
        cfg = group_config(4096)
        g = VP()
        assert lib.sdrm_rx_group_create(C.byref(cfg), sessions, 2, C.byref(g)) == 0
        chunks = []
        good_b.settimeout(30)
        for o in range(0, len(iq), 4096):
            part = np.ascontiguousarray(iq[o:o + 4096])
            lib.sdrm_rx_group_put(part.ctypes.data_as(VP), len(part), g)
        lib.sdrm_rx_group_shutdown(g)
        import time
        t0 = time.time()
        while lib.sdrm_rx_group_blocks_done(g) < len(iq) // 4096 and time.time() - t0 < 30:
            time.sleep(0.01)
        assert lib.sdrm_rx_group_sessions_failed(g) == 1 and lib.sdrm_rx_group_failed(g) == 0
        lib.sdrm_rx_group_destroy(g)
        good_a.close()
        while True:
            data = good_b.recv(65536)
            if not data:
                break
            chunks.append(data)
    finally:
        signal.signal(signal.SIGPIPE, old)
        good_b.close()
        dead_a.close()
    got = np.frombuffer(b"".join(chunks), dtype=np.int8)
    want, _ = port.FskDemod(*args, 4096).run(iq, 4096)
    assert same_bits(got, want)


def test_create_destroy_cycles_release_device_and_pinned_memory(sdrm):
    """Every handle type created, used once and destroyed 25 times over: the device's free memory comes back to where it
    was (cudaMemGetInfo; a leaked staging buffer of these sizes would show as tens of MB per cycle), and so does the
    process's pinned host memory as far as the resident set shows it."""
    import resource

    import torch
    from conftest import LUCKY7_TLE
    from test_gpu_doppler import LAT, LON, DopplerHandle

    rng = np.random.default_rng(3)
    iq = (rng.standard_normal((8, 20000)) + 1j * rng.standard_normal((8, 20000))).astype(np.complex64)
    data = rng.integers(0, 256, (8, 256), dtype=np.uint8)

    def cycle():
        b = sdrm.FskDemodBatch(8, 192000, 9600, 5000, 2, 2000, True, 20000, soft=True)
        b.process(iq)
        b.submit_i16(np.zeros((8, 20000, 2), np.int16))
        b.fetch()
        b.close()
        h = sdrm.FskDemod(48000, 4800, 5000, 2, 2000, True, 20000)
        h.process(iq[0])
        h.close()
        m = sdrm.GfskModBatch(8, 2.0, 1.6, 0.5, 256)
        m.process(data)
        m.close()
        lp = sdrm.LpfBatch(8, 4, 192000, 10000, 2000, 20000, True)
        lp.process(iq)
        lp.close()
        nco = sdrm.NcoBatch(8, 1.0, 192000, 20000)
        nco.multiply(np.arange(8) * 100, iq)
        nco.close()
        d = DopplerHandle(sdrm.lib, LAT, LON, 0.0, 48000, 437525000, 0, 1583840449, 20000, LUCKY7_TLE)
        d.run(iq[0], 20000)
        d.close()
        multi = sdrm.FskDemodMulti([0], 8, 192000, 9600, 5000, 2, 2000, True, 20000)
        multi.process(iq)
        multi.close()

    for _ in range(3):
        cycle()  # first-use allocations of the runtime itself (module load, stream pools, the allocator's own pools)
    torch.cuda.synchronize()
    free_before, _ = torch.cuda.mem_get_info()
    rss_before = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    for _ in range(25):
        cycle()
    torch.cuda.synchronize()
    free_after, _ = torch.cuda.mem_get_info()
    rss_after = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    assert free_before - free_after < 8 << 20, "device memory went down by %.1f MB over 25 cycles" % ((free_before - free_after) / 2 ** 20)
    assert rss_after - rss_before < 64 * 1024, "peak resident set grew by %d KB over 25 cycles" % (rss_after - rss_before)
