#!/usr/bin/env python3
"""Secondary benchmark lines (not the driver's contract; that is bench.py = BASELINE configs[1]):

  python tools/bench_configs.py --config c4 [--channels 1024]   gfsk_mod batch, 2048-byte packets, sps 2 (configs[3])
  python tools/bench_configs.py --config c1                      one channel on one host core, reference CPU chain (configs[0])
  python tools/bench_configs.py --config c3alt                   configs[2] recomposed as doppler -> decimating lpf -> fsk_demod
  python tools/bench_configs.py --config perf                    the reference's perf_fsk_modem shape (48 ksps, 4800 baud) as a batch
  python tools/bench_configs.py --config c3 [--channels 4096]   doppler + GMSK 2400 baud from 2.4 Msps, decim 100 (configs[2])

Each prints one JSON line with the device-resident throughput, the roofline that bounds the stage, and the reference's
CPU chain (oracle/_ref) timed on the host cores for a bounded sample.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sdr-modem_b200"))

LUCKY7_TLE = ["LUCKY-7",
              "1 44406U 19038W   20069.88080907  .00000505  00000-0  32890-4 0  9992",
              "2 44406  97.5270  32.5584 0026284 107.4758 252.9348 15.12089395 37524"]


def xorshift32_channels(n_bytes, seeds):
    """xorshift32 payload bytes for many channels at once: uint8 [channels][n_bytes]"""
    x = np.array(seeds, dtype=np.uint32)
    out = np.empty((len(seeds), n_bytes), dtype=np.uint8)
    for i in range(n_bytes):
        x ^= x << np.uint32(13)
        x ^= x >> np.uint32(17)
        x ^= x << np.uint32(5)
        out[:, i] = x & np.uint32(0xFF)
    return out


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def bench_c4(args):
    import torch
    import sdrm
    from oracle import ref
    import workloads
    n_ch, packet = args.channels, 2048
    device = getattr(args, "device", 0)
    sps, sens = 2.0, float(np.float32(2 * np.pi * 5000 / 19200))
    out_per_packet = packet * 8 * int(sps)
    mod = sdrm.GfskModBatch(n_ch, sps, sens, 0.5, packet, device=device)
    rng = np.random.default_rng(2000)
    # payload: xorshift32(seed = 2000 + c) per channel, two packets per channel, rotating
    data = torch.from_numpy(xorshift32_channels(2 * packet, [2000 + c for c in range(n_ch)]).reshape(n_ch, 2, packet)
                            .transpose(1, 0, 2).copy()).cuda()
    out = torch.empty((2, n_ch, out_per_packet), dtype=torch.complex64, device="cuda")
    stream = torch.cuda.ExternalStream(mod.stream)

    def step(k):
        mod.process_device(data[k % 2].data_ptr(), packet, packet, out[k % 2].data_ptr(), out_per_packet)

    for k in range(args.warmup):
        step(k)
    mod.sync()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(stream)
    for k in range(args.steps):
        step(k)
    end.record(stream)
    mod.sync()
    ms = start.elapsed_time(end) / args.steps
    samples = n_ch * out_per_packet
    value = samples / (ms * 1e-3) / 1e6
    pk, kind = peaks()
    bytes_per_sample = 8 + 1.0 / (8 * sps)
    achieved = value * 1e6 * bytes_per_sample / 1e9
    cores = os.cpu_count() or 1
    cpu = None
    if not getattr(args, "no_cpu", False):
        cpu_data = rng.integers(0, 256, (cores, packet), dtype=np.uint8)
        sec, cnt = ref.bench_gfsk_mod(sps, sens, 0.5, cpu_data, cores, packets=200)
        cpu = {"value": cnt / sec / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "reference",
               "sample": "%d channels x 200 packets of %d bytes" % (cores, packet)}
    launches = int(mod.launch_count)
    mod.close()
    del data, out
    return ({
        "metric": "modulated Msamples/s (output)", "value": value, "unit": "Msamples/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%d channels x %d-byte packets, gfsk_mod sps 2, BT 0.5 (BASELINE configs[3])" % (n_ch, packet),
                   "l2": "2 x %.0f MB output buffers rotating" % (samples * 8 / 1e6)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                     "note": "8.06 algorithmic B per output sample; bounded by the serial float phase recurrence (one lane per "
                             "channel, 20.5 cycles per sample: 0.34 ms per 32768-sample packet whatever the channel count, up to "
                             "148 SMs x 4 warps x 32 lanes = 18944 channels) and, past that, by the 24 B per sample the three passes (shaping, walk, cos / sin) move through HBM"},
        "cpu_baseline": cpu, "gpu_launches": launches})


def bench_c3(args):
    import torch
    import sdrm
    from oracle import ref
    import workloads
    n_ch, chunk = args.channels, 131072
    device = getattr(args, "device", 0)
    fast = getattr(args, "mode", "exact") == "fast"
    shape = workloads.DemodShape("gmsk2400@2.4M/chunk131072", 2400000, 2400, 5000, 100, 2000, True, chunk)
    flops, t1, t2 = workloads.demod_flops_per_sample(shape)
    lat, lon = float(np.float32(53.72)), float(np.float32(47.57))
    channels = [sdrm.doppler_channel(lat, lon, 0.0, 0, 1583840449 + c, LUCKY7_TLE) for c in range(n_ch)]
    dop = sdrm.DopplerBatch(channels, shape.sampling_freq, 437525000, chunk, device=device)
    cap = int(chunk / 100 / 10 * 1.2) + 64
    demod = sdrm.FskDemodBatch(n_ch, *shape.create_args, chunk, max_symbols_per_call=cap, device=device, fast=fast)
    t0 = time.time()
    iq = workloads.gfsk_channels(n_ch, 2 * chunk, shape, seed=3000, device="cuda", max_offset_hz=4000.0)
    bufs = [iq[:, i * chunk:(i + 1) * chunk].contiguous() for i in range(2)]
    del iq
    corrected = torch.empty((n_ch, chunk), dtype=torch.complex64, device="cuda")
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    dop_stream = torch.cuda.ExternalStream(dop.stream)
    fir_stream = torch.cuda.ExternalStream(demod.stream)
    tail_stream = torch.cuda.ExternalStream(demod.tail_stream)
    ev = torch.cuda.Event()

    def step(k):
        # the doppler output buffer is reused every step: wait until the previous call's filters have consumed it
        dop_stream.wait_stream(fir_stream)
        dop.process_device(bufs[k % 2].data_ptr(), chunk, chunk, corrected.data_ptr(), chunk, direction=1)
        ev.record(dop_stream)
        fir_stream.wait_event(ev)
        demod.process_device(corrected.data_ptr(), chunk, chunk)
        demod.release()

    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(dop_stream)
    for k in range(args.steps):
        step(k)
    end.record(tail_stream)
    torch.cuda.synchronize()
    ms = start.elapsed_time(end) / args.steps
    value = n_ch * chunk / (ms * 1e-3) / 1e6
    demod.set_profiling(True)
    step(0)
    stage = demod.stage_times()
    pk, kind = peaks()
    peak = getattr(args, "fp32_peak", None) or 148 * 128 * 2 * float(pk.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
    achieved = value * 1e6 * flops / 1e12
    cores = os.cpu_count() or 1
    cpu = None
    flags = demod.error_flags()
    demod.close()
    dop.close()
    del bufs, corrected
    if not args.no_cpu:
        x = workloads.gfsk_channels(cores, chunk, shape, seed=3000, device="cpu", max_offset_hz=4000.0).numpy()
        t0 = time.time()
        # the reference's dsp_worker chain: doppler_process_rx -> fsk_demod_process, one channel per core (threads via processes)
        import concurrent.futures as cf

        def one(c):
            d = ref.doppler(lat, lon, 0.0, shape.sampling_freq, 437525000, 0, 1583840449 + c, chunk, LUCKY7_TLE)
            f = ref.fsk_demod(*shape.create_args, chunk)
            return len(f.process(d.process(x[c])))
        with cf.ThreadPoolExecutor(cores) as pool:  # ctypes releases the GIL inside the reference's C code
            list(pool.map(one, range(cores)))
        sec = time.time() - t0
        cpu = {"value": cores * chunk / sec / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "reference",
               "sample": "%d channels x %d samples (one call each), doppler_process_rx + fsk_demod_process" % (cores, chunk)}
    return ({
        "metric": "demodulated Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%d channels x doppler + %s, decim 100, dc on (BASELINE configs[2]); T1 = %d, T2 = %d"
                               % (n_ch, shape.name, t1, t2), "mode": "fast" if fast else "exact", "input_gen_s": gen_s,
                   "flop_per_sample": flops},
        "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "kernel_ms": stage[0], "kernel_frac": 4.0 * t1 * n_ch * chunk / (stage[0] * 1e-3) / 1e12 / peak,
                     "stage_ms": {"lpf1_quad": stage[0], "lpf2": stage[1], "dc_clock_tail": stage[2]}},
        "cpu_baseline": cpu, "error_flags": flags})


def bench_perf_shape(args):
    """The reference's own perf_fsk_modem demodulator shape (test/perf_fsk_modem.c:72: 48 ksps, 4800 baud, decimation 2: T1 = 157,
    T2 = 57, 5 samples per symbol after decimation) as a batch of 1024 channels: the filters are light here and the serial tail
    (dc blocker + clock recovery) is what bounds the step."""
    import torch
    import sdrm
    from oracle import ref
    import workloads
    n_ch, chunk = args.channels or 1024, 131072
    shape = workloads.DemodShape("perf_fsk_modem 4800@48k/chunk131072", 48000, 4800, 5000, 2, 2000, True, chunk)
    flops, t1, t2 = workloads.demod_flops_per_sample(shape)
    iq = workloads.gfsk_channels(n_ch, 2 * chunk, shape, seed=1000, device="cuda")
    bufs = [iq[:, :chunk].contiguous(), iq[:, chunk:].contiguous()]
    demod = sdrm.FskDemodBatch(n_ch, *shape.create_args, chunk, max_symbols_per_call=int(chunk / 10 * 1.2) + 64)
    fir, tail = torch.cuda.ExternalStream(demod.stream), torch.cuda.ExternalStream(demod.tail_stream)

    def step(k):
        demod.process_device(bufs[k % 2].data_ptr(), chunk, chunk)
        demod.release()

    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(fir)
    for k in range(args.steps):
        step(k)
    end.record(tail)
    torch.cuda.synchronize()
    ms = start.elapsed_time(end) / args.steps
    value = n_ch * chunk / (ms * 1e-3) / 1e6
    demod.set_profiling(True)
    step(0)
    stage = demod.stage_times()
    cpu = None
    if not args.no_cpu:
        cores = os.cpu_count() or 1
        x = workloads.gfsk_channels(cores, 2 * chunk, shape, seed=1000, device="cpu").numpy()
        sec, _ = ref.bench_fsk_demod(*shape.create_args, chunk, x, cores, passes=4)
        cpu = {"value": cores * 2 * chunk * 4 / sec / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "reference",
               "sample": "%d channels x %d samples x 4 passes, one thread per channel" % (cores, 2 * chunk)}
    return ({
        "metric": "demodulated Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%d channels x %s, dc on (the reference's perf_fsk_modem shape); T1 = %d, T2 = %d"
                               % (n_ch, shape.name, t1, t2), "mode": "exact", "flop_per_sample": flops},
        "roofline": {"bound": "serial tail (dependent-issue latency), not a throughput roofline",
                     "stage_ms": {"lpf1_quad": stage[0], "lpf2": stage[1], "dc_clock_tail": stage[2]}},
        "cpu_baseline": cpu, "error_flags": demod.error_flags()})


def bench_c1(args):
    """BASELINE configs[0]: ONE channel on ONE host core through the reference's own CPU chain (oracle/_ref), 10 s of signal in
    4096-sample calls; plus the reference's published perf shape (test/perf_fsk_modem.c: 48 kHz, 4800 baud, input (uint8) i)."""
    from oracle import ref
    import workloads
    out = {}
    for name, shape, n in (("c1_192k_9600", workloads.C2_PARITY, 1920000), ("perf_fsk_modem_48k_4800", workloads.PERF_SHAPE, 480000)):
        if name.startswith("perf"):
            iq = ((np.arange(n) % 256).astype(np.float32) + 0j).astype(np.complex64)[None, :]
        else:
            iq = workloads.gfsk_channels(1, n, shape, seed=1000, device="cpu").numpy()
        sec, symbols = ref.bench_fsk_demod(*shape.create_args, shape.chunk, iq, 1, passes=1)
        out[name] = {"msamples_per_s": n / sec / 1e6, "seconds": sec, "samples": n, "symbols": int(symbols), "chunk": shape.chunk}
        try:
            # the same stream through the library's single-handle drop-in (fsk_demod_create / fsk_demod_process, one synchronous
            # call per 4096 samples: copy in, five launches, copy out): latency-bound, the path perf_fsk_modem.c measures
            import torch
            if torch.cuda.is_available():
                import time
                import sdrm
                warm = sdrm.FskDemod(*shape.create_args, shape.chunk)  # context, module load, first-launch costs
                warm.process(iq[0, :shape.chunk])
                warm.close()
                handle = sdrm.FskDemod(*shape.create_args, shape.chunk)
                t0 = time.perf_counter()
                got = 0
                for o in range(0, n, shape.chunk):
                    got += len(handle.process(iq[0, o:o + shape.chunk]))
                gpu_sec = time.perf_counter() - t0
                handle.close()
                out[name]["gpu_single_handle_msamples_per_s"] = n / gpu_sec / 1e6
                out[name]["gpu_single_handle_symbols"] = int(got)
        except Exception as ex:  # the CPU figure stands on its own
            out[name]["gpu_single_handle_error"] = str(ex)
    return ({"metric": "demodulated Msamples/s", "value": out["c1_192k_9600"]["msamples_per_s"], "unit": "Msamples/s",
                      "n_gpus": 0, "impl": "reference", "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "1 channel, 1 core, oracle/_ref strict build (BASELINE configs[0])"},
                      "cpu_baseline": {"cores": 1, "kind": "reference"}, "shapes": out,
                      "gpu_single_handle_value": out["c1_192k_9600"].get("gpu_single_handle_msamples_per_s")})


def bench_c3alt(args):
    """The same sessions as C3 composed the other legal way: doppler -> complex decimating lpf (dec 25, 2.4 Msps ->
    96 ksps) -> fsk_demod at the low rate (decim 4). Not the parity-defining chain; shows what the public API allows."""
    import torch
    import sdrm
    import workloads
    n_ch, chunk, dec1 = args.channels, 131072, 25
    shape = workloads.DemodShape("gmsk2400@2.4M/chunk131072", 2400000, 2400, 5000, 100, 2000, True, chunk)
    lat, lon = float(np.float32(53.72)), float(np.float32(47.57))
    channels = [sdrm.doppler_channel(lat, lon, 0.0, 0, 1583840449 + c, LUCKY7_TLE) for c in range(n_ch)]
    dop = sdrm.DopplerBatch(channels, shape.sampling_freq, 437525000, chunk, device=0)
    lpf = sdrm.LpfBatch(n_ch, dec1, shape.sampling_freq, 20000, 10000, chunk, True, device=0)
    low_len = chunk // dec1 + 1
    demod = sdrm.FskDemodBatch(n_ch, shape.sampling_freq // dec1, 2400, 5000, 4, 2000, True, low_len,
                               max_symbols_per_call=int(low_len / 40 * 1.2) + 64, device=0)
    iq = workloads.gfsk_channels(n_ch, 2 * chunk, shape, seed=3000, device="cuda", max_offset_hz=4000.0)
    bufs = [iq[:, i * chunk:(i + 1) * chunk].contiguous() for i in range(2)]
    del iq
    corrected = torch.empty((n_ch, chunk), dtype=torch.complex64, device="cuda")
    low_stride = low_len + (low_len & 1)
    low = [torch.empty((n_ch, low_stride), dtype=torch.complex64, device="cuda") for _ in range(2)]
    torch.cuda.synchronize()
    dop_stream = torch.cuda.ExternalStream(dop.stream)
    lpf_stream = torch.cuda.ExternalStream(lpf.stream)
    fir_stream = torch.cuda.ExternalStream(demod.stream)
    tail_stream = torch.cuda.ExternalStream(demod.tail_stream)
    ev_dop, ev_lpf, ev_low_free = torch.cuda.Event(), torch.cuda.Event(), [torch.cuda.Event(), torch.cuda.Event()]

    def step(k):
        dop_stream.wait_stream(lpf_stream)  # `corrected` is reused every step
        dop.process_device(bufs[k % 2].data_ptr(), chunk, chunk, corrected.data_ptr(), chunk, direction=1)
        ev_dop.record(dop_stream)
        lpf_stream.wait_event(ev_dop)
        lpf_stream.wait_event(ev_low_free[k % 2])
        n_low = lpf.process_device(corrected.data_ptr(), chunk, chunk, low[k % 2].data_ptr(), low_stride)
        ev_lpf.record(lpf_stream)
        fir_stream.wait_event(ev_lpf)
        demod.process_device(low[k % 2].data_ptr(), low_stride, n_low)
        ev_low_free[k % 2].record(fir_stream)
        demod.release()

    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(dop_stream)
    for k in range(args.steps):
        step(k)
    end.record(tail_stream)
    torch.cuda.synchronize()
    ms = start.elapsed_time(end) / args.steps
    value = n_ch * chunk / (ms * 1e-3) / 1e6
    pk, kind = peaks()
    return ({
        "metric": "demodulated Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%d channels x doppler -> lpf(dec %d, cutoff 20 kHz, tw 10 kHz) -> fsk_demod(96 ksps, 2400 baud, "
                               "decim 4, dc on): BASELINE configs[2] recomposed, NOT the parity-defining chain" % (n_ch, dec1)},
        "roofline": {"bound": "hbm", "achieved": value * 1e6 * 24 / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": value * 1e6 * 24 / 1e9 / pk["hbm_gbs"],
                     "note": "24 algorithmic B per input sample: doppler 8 in + 8 out, lpf 8 in (+ 8/25 out)"},
        "error_flags": demod.error_flags()})


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--config", required=True, choices=["c1", "c3", "c3alt", "c4", "perf"])
    p.add_argument("--channels", type=int, default=None)
    p.add_argument("--steps", type=int, default=None)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--mode", default="exact", choices=["exact", "fast"], help="c3: arithmetic mode of the demodulator")
    p.add_argument("--device", type=int, default=0)
    p.add_argument("--fp32-peak", type=float, default=None, dest="fp32_peak",
                   help="measured FMA peak in TFLOP/s to quote the fp32 roofline against (bench.py passes its own measurement)")
    args = p.parse_args()
    if args.config not in ("c1",):
        import torch
        torch.cuda.set_device(args.device)
    if args.config == "c4":
        args.channels = args.channels or 1024
        args.steps = args.steps or 50
        line = bench_c4(args)
    elif args.config == "c1":
        line = bench_c1(args)
    elif args.config == "c3alt":
        args.channels = args.channels or 1024
        args.steps = args.steps or 10
        line = bench_c3alt(args)
    elif args.config == "perf":
        args.steps = args.steps or 20
        line = bench_perf_shape(args)
    else:
        args.channels = args.channels or 4096
        args.steps = args.steps or 3
        line = bench_c3(args)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
