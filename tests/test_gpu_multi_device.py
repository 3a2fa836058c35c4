"""Multi-device entry points (include/sdrm/sdrm_multi.h): a job partitioned over a device list by channel range must give,
bit for bit, what one batch on one device gives. The partition, the per-device threads and the pinned ingest rings are
exercised on a single GPU by listing device 0 more than once; with two or more GPUs the same job also runs across them.
Reference analogue: the sdr_worker -> N x dsp_worker fan-out (src/sdr_worker.c:31-55)."""
import ctypes as C

import numpy as np
import pytest

import workloads
from conftest import LUCKY7_TLE, golden_array, same_bits

pytestmark = pytest.mark.gpu


def device_count():
    import torch
    return torch.cuda.device_count()


def single_device(sdrm, shape, iq, chunk, cap):
    b = sdrm.FskDemodBatch(iq.shape[0], *shape.create_args, chunk, max_symbols_per_call=cap, soft=True, device=0)
    out = []
    for o in range(0, iq.shape[1], chunk):
        hard, lens, soft = b.process(iq[:, o:o + chunk])
        out.append((hard.copy(), lens.copy(), soft.copy()))
    assert b.error_flags() == 0
    b.close()
    return out


@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0], "all"], ids=["2_shards_on_gpu0", "3_shards_on_gpu0", "all_gpus"])
@pytest.mark.parametrize("pinned", [False, True], ids=["pageable_input_staged", "pinned_input"])
def test_multi_device_equals_single_device(sdrm, devices, pinned):
    if devices == "all":
        if device_count() < 2:
            pytest.skip("needs two GPUs")
        devices = list(range(device_count()))
    shape = workloads.C2_PARITY
    n_ch, chunk, calls, cap = 7, 4096, 5, 256  # 7 channels over 2 / 3 shards: uneven slices, odd counts inside the slices
    if len(devices) > n_ch:
        devices = devices[:n_ch]
    iq = workloads.gfsk_channels(n_ch, chunk * calls, shape, seed=21).numpy()
    want = single_device(sdrm, shape, iq, chunk, cap)
    m = sdrm.FskDemodMulti(devices, n_ch, *shape.create_args, chunk, max_symbols_per_call=cap, soft=True)
    shards = m.shards()
    assert [s[2] for s in shards] == devices
    assert shards[0][0] == 0 and sum(s[1] for s in shards) == n_ch
    assert all(shards[g][0] + shards[g][1] == shards[g + 1][0] for g in range(len(shards) - 1))
    keep = []
    try:
        # three calls in flight, then steady state: submit k, fetch k - 2
        got = []
        for k in range(calls):
            part = np.ascontiguousarray(iq[:, k * chunk:(k + 1) * chunk])
            if pinned:
                pin = sdrm.PinnedArray(part.shape, np.complex64)
                pin.array[:] = part
                keep.append(pin)
                part = pin.array
            m.submit(part)
            keep.append(part)
            if k >= 2:
                got.append(m.fetch())
        while len(got) < calls:
            got.append(m.fetch())
        assert m.error_flags() == 0
        assert m.launch_count >= 5 * calls * len(devices)
    finally:
        m.close()
    for k in range(calls):
        hard, lens, soft = got[k]
        w_hard, w_lens, w_soft = want[k]
        assert np.array_equal(lens, w_lens)
        for c in range(n_ch):
            assert same_bits(hard[c, :lens[c]], w_hard[c, :lens[c]]), (k, c)
            assert same_bits(soft[c, :lens[c]], w_soft[c, :lens[c]]), (k, c)


def test_multi_device_create_errors(sdrm):
    shape = workloads.C2_PARITY
    with pytest.raises(sdrm.SdrmError):
        sdrm.FskDemodMulti([], 4, *shape.create_args, 4096)
    with pytest.raises(sdrm.SdrmError):
        sdrm.FskDemodMulti([0, 0, 0], 2, *shape.create_args, 4096)  # more shards than channels
    with pytest.raises(sdrm.SdrmError):
        sdrm.FskDemodMulti([0, 99], 4, *shape.create_args, 4096)  # no such device


def test_rx_multi_fan_out_equals_one_group(sdrm):
    """The same SDR stream and sessions through sdrm_rx_multi on shards [0, 0] (and on every GPU when there are several)
    and through one sdrm_rx_group: identical symbols per session."""
    from test_gpu_worker import SINK, RxSession, VP, group_config, setup_group
    lib = sdrm.lib
    setup_group(lib)
    lib.sdrm_rx_multi_create.argtypes = [VP, VP, C.c_uint32, C.POINTER(C.c_int), C.c_uint32, C.POINTER(VP)]
    lib.sdrm_rx_multi_put.argtypes = [VP, C.c_size_t, VP]
    lib.sdrm_rx_multi_put.restype = None
    lib.sdrm_rx_multi_shutdown.argtypes = [VP]
    lib.sdrm_rx_multi_shutdown.restype = None
    lib.sdrm_rx_multi_destroy.argtypes = [VP]
    lib.sdrm_rx_multi_destroy.restype = None
    lib.sdrm_rx_multi_failed.argtypes = [VP]
    raw = golden_array("lucky7.cf32", np.complex64)[:40000]
    chunk = 2000
    starts = [1583840449, None, 1583840449 + 61, None, 1583840449 + 95]

    def run(create, put, shutdown, destroy, failed=None):
        collected = {i: [] for i in range(len(starts))}

        def on_symbols(ctx, session_id, symbols, n):
            collected[session_id].append(np.ctypeslib.as_array(symbols, shape=(n,)).copy())

        sink = SINK(on_symbols)
        sessions = (RxSession * len(starts))()
        for i, start in enumerate(starts):
            s = sessions[i]
            s.id, s.client_socket, s.sink, s.sink_ctx = i, -1, sink, None
            s.has_doppler = start is not None
            for k, line in enumerate(LUCKY7_TLE):
                rawline = line.encode("ascii")
                C.memmove(C.addressof(s.doppler_tle[k]), rawline + b"\0", len(rawline) + 1)
            s.doppler_latitude, s.doppler_longitude, s.doppler_altitude = 537200000, 475700000, 0
            s.file_start_time_seconds = start or 0
        cfg = group_config(chunk)
        h = create(cfg, sessions)
        for o in range(0, len(raw), chunk):
            part = np.ascontiguousarray(raw[o:o + chunk])
            put(part.ctypes.data_as(VP), len(part), h)
        shutdown(h)
        if failed is not None:
            assert failed(h) in (0, 1)
        destroy(h)
        return {i: (np.concatenate(v) if v else np.zeros(0, np.int8)) for i, v in collected.items()}

    def create_group(cfg, sessions):
        g = VP()
        assert lib.sdrm_rx_group_create(C.byref(cfg), sessions, len(starts), C.byref(g)) == 0
        return g

    want = run(create_group, lib.sdrm_rx_group_put, lib.sdrm_rx_group_shutdown, lib.sdrm_rx_group_destroy)
    device_lists = [[0, 0]]
    if device_count() >= 2:
        device_lists.append(list(range(min(device_count(), len(starts)))))
    for devices in device_lists:
        def create_multi(cfg, sessions):
            m = VP()
            dev = (C.c_int * len(devices))(*devices)
            assert lib.sdrm_rx_multi_create(C.byref(cfg), sessions, len(starts), dev, len(devices), C.byref(m)) == 0
            return m

        got = run(create_multi, lib.sdrm_rx_multi_put, lib.sdrm_rx_multi_shutdown, lib.sdrm_rx_multi_destroy)
        for i in range(len(starts)):
            assert len(want[i]) > 300 and same_bits(got[i], want[i]), "session %d on devices %s" % (i, devices)
