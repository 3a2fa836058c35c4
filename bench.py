#!/usr/bin/env python3
"""Headline benchmark: batched GMSK 9600-baud demodulation (BASELINE.json configs[1]) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--channels C] [--mode exact|fast]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      the reference's own CPU chain (oracle/_ref) on the host cores

One step = one call of the demod chain (lpf1 -> quadrature demod -> lpf2 -> dc blocker -> clock recovery -> int8)
over `chunk` new samples of every channel. `value` counts input complex samples per second with the inputs resident
in HBM and the int8 results copied back to pinned host memory inside the timed loop; `e2e` is the same through the
host-buffer C-ABI call (pinned host input, H2D and D2H inside the timed region), beside a plain-copy probe of what the
host gives that many GPUs at once. Channels are sharded across ranks (weak scaling: `--channels` per GPU), there is no
collective on the data path. Prints ONE JSON line on rank 0.

The line also carries, outside every timed region (rank 0): the FMA arithmetic mode with its measured parity counts against
the strict reference build (`roofline.fma_*`), the reference's CPU chain on the host cores (`cpu_baseline`, strict and
tuned builds), and the other BASELINE.json configs (`configs`: modulator, Doppler + 2.4 Msps chain, one CPU channel, and on 8
GPUs the 32768-channel job), each with its own roofline fraction and CPU baseline.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sdr-modem_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

METRIC = "demodulated Msamples/s"
UNIT = "Msamples/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--channels", type=int, default=1024, help="channels per GPU")
    p.add_argument("--chunk", type=int, default=131072, help="samples per channel per step")
    p.add_argument("--mode", default="exact", choices=["exact", "fast"])
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-fma-line", action="store_true", help="skip the extra measurement in the FMA arithmetic mode")
    p.add_argument("--no-parity", action="store_true", help="skip the fast-mode parity counts (CPU reference runs)")
    p.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs (C1, C3 subset, C4, C5)")
    p.add_argument("--cpu-seconds", type=float, default=10.0)
    p.add_argument("--debug-no-tail", action="store_true", help="measurement aid: skip the serial tail (invalid result)")
    p.add_argument("--full-line", default=None, help="also write the JSON line to this file")
    p.add_argument("--single-process", action="store_true",
                   help="one process drives --gpus N devices through the C multi-device entry points (sdrm_fsk_demod_multi_*): "
                        "host buffers in, int8 out; not the driver's launch, which is one rank per GPU")
    return p.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def wait_first_sample(self, timeout=10.0):
        """nvidia-smi takes a while to start; the timed region should not begin before it delivers"""
        t0 = time.time()
        while self.proc is not None and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, t_begin=None, t_end=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, sm_max, power, reasons = [], [], [], set()
        for stamp, line in self.lines:
            if t_begin is not None and not (t_begin <= stamp <= t_end):
                continue  # only samples taken while the timed region was running
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                sm_max.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(sm_max)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def k1_traffic(samples_per_launch):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture, scaled to this launch."""
    for name in ("r2_k1_traffic.json", "r1_k1_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            with open(path) as f:
                t = json.load(f)
            per_sample = (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["samples_per_launch"]
            return per_sample * samples_per_launch
    return None


def c2_shape(chunk):
    import workloads
    return workloads.DemodShape("gmsk9600@192k/chunk%d" % chunk, 192000, 9600, 5000, 2, 2000, True, chunk)


def c2_config(args, shape, mode):
    """The `config` of the line: identical in the b200 arm and the reference arm (the reference arm times a bounded sample of
    this workload and says so in cpu_baseline.sample)."""
    import workloads
    flops, t1, t2 = workloads.demod_flops_per_sample(shape)
    return {"workload": "%d channels/GPU x %s, GMSK BT 0.5, Eb/N0 12 dB, dev 5 kHz, decim 2, dc on (BASELINE configs[1])"
                        % (args.channels, shape.name),
            "channels_per_gpu": args.channels, "chunk": args.chunk, "mode": mode,
            "parallelism": "channels sharded, no collective",
            "l2": "inputs larger than L2 (2 x %.2f GiB rotating)" % (args.channels * args.chunk * 8 / 2 ** 30),
            "flop_per_sample": flops, "t1": t1, "t2": t2,
            "generator": "workloads.gfsk_channels: torch RNG payload, Gaussian pulse by conv1d, per-channel carrier/timing offset, "
                         "AWGN (same statistics as SURVEY 8d's xorshift + oracle gfsk_mod recipe, not its bytes; both arms and the "
                         "parity checks read the same samples)"}


def cpu_reference_run(shape, seconds, n_threads=None, build=False):
    """The reference's own CPU chain (oracle/_ref) driven like dsp_worker: one pthread per channel. build: False = strict
    (the parity-defining build), "tuned" = -O3 + lane-partial dot products (SIMD VOLK stand-in, baseline only)."""
    from oracle import ref
    import workloads
    if not ref.available(build):
        return None
    cores = n_threads or os.cpu_count() or 1
    n = shape.chunk * 2
    iq = workloads.gfsk_channels(cores, n, shape, seed=1000, device="cpu").numpy()
    # calibrate with one pass, then size the timed run
    sec, _ = ref.bench_fsk_demod(*shape.create_args, shape.chunk, iq, cores, passes=1, fma=build)
    passes = max(1, int(seconds / max(sec, 1e-3)))
    sec, symbols = ref.bench_fsk_demod(*shape.create_args, shape.chunk, iq, cores, passes=passes, fma=build)
    samples = cores * n * passes
    flavour = ("oracle/_ref tuned build v%d (-O3 -march=x86-64-v%d -ffp-contract=fast, lane-partial SIMD dot products: stand-in for "
               "SIMD VOLK, not bit-identical)" % (ref.tuned_level(), ref.tuned_level())) if build == "tuned" else \
        "oracle/_ref strict build (-O2 -ffp-contract=off, VOLK generic shim)"
    return {"value": samples / sec / 1e6, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "%d ch (1/thread) x %d samples x %d passes, chunk %d, %.1f s; per-sample rate, channel count does not enter; %s"
                      % (cores, n, passes, shape.chunk, sec, flavour),
            "seconds": sec, "symbols": int(symbols)}


_RESULT_FD = None


def claim_stdout():
    """The driver parses stdout for ONE JSON line. Libraries write there too (NCCL prints its version banner on stdout when
    NCCL_DEBUG is set in the environment), so file descriptor 1 is pointed at stderr for everything except emit_line()."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit_line(line, also=None):
    data = (json.dumps(line) + "\n").encode()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, data)
    if also:
        with open(also, "wb") as f:
            f.write(data)


def run_reference_arm(args, shape):
    """bench.py --impl reference: the reference's own CPU chain (oracle/_ref, built from the reference sources in place) on all
    host cores, same metric and `config` as the b200 arm; a step is a bounded sample of that workload (one channel per core,
    2 chunks, a few passes), so that K steps + W warm-ups end within a few minutes. Both CPU builds are timed, the strict one
    (what the reference's tests pin, VOLK_GENERIC=1) and the tuned one (stand-in for SIMD VOLK); `value` is the FASTER of
    the two, so that the driver's ratio is against the best CPU figure available here."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref
    import workloads
    if not ref.available():
        emit_line({"impl": "reference", "unavailable": "oracle/_ref/libsdrmodem_ref.so not built"})
        return
    cores = os.cpu_count() or 1
    n = shape.chunk * 2
    iq = workloads.gfsk_channels(cores, n, shape, seed=1000, device="cpu").numpy()
    builds = [False] + (["tuned"] if ref.available("tuned") and ref.tuned_level() else [])
    results = {}
    for build in builds:
        sec, _ = ref.bench_fsk_demod(*shape.create_args, shape.chunk, iq, cores, passes=1, fma=build)
        target = max(0.2, args.cpu_seconds / len(builds) / max(1, args.steps))
        passes = max(1, int(round(target / max(sec, 1e-3))))
        for _ in range(args.warmup):
            ref.bench_fsk_demod(*shape.create_args, shape.chunk, iq, cores, passes=1, fma=build)
        total_sec, total_samples = 0.0, 0
        for _ in range(args.steps):
            sec, _ = ref.bench_fsk_demod(*shape.create_args, shape.chunk, iq, cores, passes=passes, fma=build)
            total_sec += sec
            total_samples += cores * n * passes
        results["tuned" if build else "strict"] = (total_samples / total_sec / 1e6, total_sec, passes)
    best = max(results, key=lambda k: results[k][0])
    value, total_sec, passes = results[best]
    sample = ("faster of 2 CPU builds = %s; %d ch (1/thread) x %d samples x %d passes/step, chunk %d; per-sample rate, channel count "
              "does not enter; strict = oracle/_ref -O2 -ffp-contract=off VOLK generic shim (parity-defining), tuned = -O3 "
              "x86-64-v%d lane-partial SIMD dot products (SIMD VOLK stand-in)" % (best, cores, n, passes, shape.chunk, ref.tuned_level()))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_sec / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": c2_config(args, shape, args.mode),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                             "build": best, "strict_value": results["strict"][0],
                             "tuned_value": results["tuned"][0] if "tuned" in results else None},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_line(line, args.full_line)


def fast_mode_parity(sdrm, shape, bufs, n_channels, cap, device):
    """Counts of the north-star epsilon rule for the FMA mode against the STRICT reference build (checker only, outside every
    timed region): the first n_channels channels of the resident bench input (2 calls of `chunk` samples = 1.37 s of signal per
    channel at the defaults) and the four golden files of the reference's test_fsk_demod.c."""
    import torch
    from oracle import parity, ref
    if not ref.available():
        return None
    chunk = shape.chunk
    sub = [b[:n_channels].contiguous() for b in bufs]
    iq = torch.cat(sub, dim=1).cpu().numpy()
    batch = sdrm.FskDemodBatch(n_channels, *shape.create_args, chunk, max_symbols_per_call=cap, fast=True, soft=True, device=device)
    got_hard = [[] for _ in range(n_channels)]
    got_soft = [[] for _ in range(n_channels)]
    for b in sub:
        batch.process_device(b.data_ptr(), chunk, chunk)
        hard, lens, soft = batch.fetch()
        for c in range(n_channels):
            got_hard[c].append(hard[c, :lens[c]].copy())
            got_soft[c].append(soft[c, :lens[c]].copy())
    flags = batch.error_flags()
    batch.close()
    got = [(np.concatenate(h), np.concatenate(s)) for h, s in zip(got_hard, got_soft)]
    report = {"c2": parity.fast_mode_report(got, shape.create_args, iq, chunk), "error_flags": flags}
    report["c2"]["signal"] = "%d channels x %d samples (%.2f s each), calls of %d" % (n_channels, iq.shape[1], iq.shape[1] /
                                                                                  shape.sampling_freq, chunk)
    golden_dir = os.path.join(ROOT, "tests", "golden")
    goldens = {"nusat": ("nusat.cf32", (192000, 40000, 5000, 1, 2000, True)),
               "nan": ("inputnan.cf32", (240000, 9600, 5000, 1, 2000, True)),
               "lucky7": ("lucky7.expected.cf32", (48000, 4800, 5000, 2, 2000, True)),
               "lucky7_nodc": ("lucky7.expected.cf32", (48000, 4800, 5000, 2, 2000, False))}
    reports = []
    for name, (fname, gargs) in sorted(goldens.items()):
        path = os.path.join(golden_dir, fname)
        if not os.path.exists(path):
            continue
        x = np.fromfile(path, dtype=np.complex64)[None, :]
        b = sdrm.FskDemodBatch(1, *gargs, 4096, fast=True, soft=True, device=device)
        hard, soft = b.run_stream(x, 4096)
        b.close()
        r = parity.fast_mode_report([(hard[0], soft[0])], gargs, x, 4096)
        report["golden_" + name] = r
        reports.append(r)
    if reports:
        merged = parity.merge(reports)
        merged["channels_bit_identical_to_fma_order_reference"] = sum(
            r["channels_bit_identical_to_fma_order_reference"] or 0 for r in reports)
        report["goldens_all"] = merged
    return report


def run_other_configs(args, world, rank, local_rank, fp32_peak):
    """BASELINE.json configs other than the headline one, each as {name, value, unit, ms_per_step, roofline_frac, ...}.
    N = 1: configs[3] (modulator, 1024 channels and at the channel count where the phase walkers fill the GPU), a 256-channel
    subset of configs[2] (Doppler + 2.4 Msps chain; the full 4096 channels take 0.6 s per step) and configs[0] (one CPU channel).
    Each runs tools/bench_configs.py in a process of its own (a clean CUDA context, and a failure there cannot cost the headline
    line). N = 8: configs[4], 32768 channels = 4096 per GPU (run_c5)."""
    out = []
    if world != 1:
        return out
    tool = os.path.join(ROOT, "tools", "bench_configs.py")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", ""))
    if not env["CUDA_VISIBLE_DEVICES"]:
        env.pop("CUDA_VISIBLE_DEVICES")

    def run(config, extra):
        cmd = [sys.executable, tool, "--config", config, "--device", str(local_rank)] + extra + (["--no-cpu"] if args.no_cpu else [])
        proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        lines = [x for x in proc.stdout.splitlines() if x.startswith("{")]
        if proc.returncode != 0 or not lines:
            raise RuntimeError("%s: rc %d: %s" % (" ".join(cmd[1:]), proc.returncode, proc.stderr[-300:]))
        return json.loads(lines[-1])

    for n_ch, steps, name in ((1024, 50, "configs[3] gfsk_mod 1024 channels"),
                              (16384, 10, "configs[3] shape at 16384 channels (walkers saturated)")):
        try:
            r = run("c4", ["--channels", str(n_ch), "--steps", str(steps)] + (["--no-cpu"] if n_ch != 1024 and not args.no_cpu else []))
            out.append({"name": name, "metric": r["metric"], "value": r["value"], "unit": r["unit"], "ms_per_step": r["ms_per_step"],
                        "steps": steps, "roofline_bound": "hbm", "roofline_achieved_gbs": r["roofline"]["achieved"],
                        "roofline_peak_gbs": r["roofline"]["peak"], "roofline_frac": r["roofline"]["frac"],
                        "cpu_value": r["cpu_baseline"]["value"] if r["cpu_baseline"] else None,
                        "cpu_cores": r["cpu_baseline"]["cores"] if r["cpu_baseline"] else None,
                        "gpu_launches": r["gpu_launches"], "note": r["roofline"]["note"]})
        except Exception as e:  # a secondary line must not cost the headline
            out.append({"name": name, "error": repr(e)[:300]})
    try:
        r = run("c3", ["--channels", "256", "--steps", "2", "--fp32-peak", "%.4f" % fp32_peak])
        out.append({"name": "configs[2] doppler + GMSK 2400 baud from 2.4 Msps, 256-channel subset of the 4096", "metric": r["metric"],
                    "value": r["value"], "unit": r["unit"], "ms_per_step": r["ms_per_step"], "steps": 2,
                    "roofline_bound": "fp32", "roofline_achieved_tflops": r["roofline"]["achieved"],
                    "roofline_peak_tflops": r["roofline"]["peak"], "roofline_frac": r["roofline"]["frac"],
                    "kernel_ms": r["roofline"]["kernel_ms"], "kernel_frac": r["roofline"]["kernel_frac"],
                    "flop_per_sample": r["config"]["flop_per_sample"],
                    "cpu_value": r["cpu_baseline"]["value"] if r["cpu_baseline"] else None,
                    "cpu_cores": r["cpu_baseline"]["cores"] if r["cpu_baseline"] else None,
                    "error_flags": r["error_flags"], "workload": r["config"]["workload"]})
    except Exception as e:
        out.append({"name": "configs[2]", "error": repr(e)[:300]})
    if not args.no_cpu:
        try:
            r = run("c1", [])
            out.append({"name": "configs[0] one channel on one host core (reference CPU chain, strict build)", "metric": r["metric"],
                        "value": r["value"], "unit": r["unit"], "cpu_cores": 1,
                        "perf_fsk_modem_shape_value": r["shapes"]["perf_fsk_modem_48k_4800"]["msamples_per_s"],
                        "seconds": r["shapes"]["c1_192k_9600"]["seconds"],
                        # the same single stream through this library's one-handle drop-in (one synchronous call per 4096
                        # samples): the latency-bound path, what the reference's perf_fsk_modem.c times
                        "gpu_single_handle_value": r["shapes"]["c1_192k_9600"].get("gpu_single_handle_msamples_per_s"),
                        "gpu_single_handle_perf_shape_value":
                            r["shapes"]["perf_fsk_modem_48k_4800"].get("gpu_single_handle_msamples_per_s")})
        except Exception as e:
            out.append({"name": "configs[0]", "error": repr(e)[:300]})
    return out


def run_c5(args, world, rank, local_rank, device, shape, fp32_peak, dist):
    """BASELINE configs[4]: 32768 channels of the headline shape, 4096 per GPU on 8 GPUs (every rank takes part)."""
    import torch
    import sdrm
    import workloads
    n_ch, chunk, steps = 4096, args.chunk, 5
    cap = int(chunk / 20 * 1.1) + 64
    batch = sdrm.FskDemodBatch(n_ch, *shape.create_args, chunk, max_symbols_per_call=cap, device=local_rank)
    bufs = []
    for i in range(2):
        x = workloads.gfsk_channels(n_ch, chunk, shape, seed=1000 + rank * n_ch + 7919 * i, device=device)
        bufs.append(x)
    pin_out = sdrm.PinnedArray((n_ch, cap), np.int8, device=local_rank)
    pin_len = sdrm.PinnedArray((n_ch,), np.uint32, device=local_rank)
    fir = torch.cuda.ExternalStream(batch.stream, device=device)
    out = torch.cuda.ExternalStream(batch.out_stream, device=device)

    def loop(n):
        fetched = 0
        for k in range(n):
            batch.process_device(bufs[k % 2].data_ptr(), chunk, chunk)
            if k + 1 >= 3:
                batch.fetch_ptr(pin_out.ptr, cap, pin_len.ptr)
                fetched += 1
        while fetched < n:
            batch.fetch_ptr(pin_out.ptr, cap, pin_len.ptr)
            fetched += 1

    loop(3)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(fir)
    loop(steps)
    e1.record(out)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    flags = batch.error_flags()
    batch.close()
    pin_out.close()
    pin_len.close()
    flops, _, _ = workloads.demod_flops_per_sample(shape)
    value = world * n_ch * chunk / (ms * 1e-3) / 1e6
    return {"name": "configs[4] 32768 channels of the headline shape, 4096 per GPU x %d GPUs" % world, "metric": METRIC, "value": value,
            "unit": UNIT, "ms_per_step": ms, "steps": steps, "channels_total": world * n_ch, "roofline_bound": "fp32",
            "roofline_frac": value / world * 1e6 * flops / 1e12 / fp32_peak, "error_flags": flags}


def run_single_process(args, shape):
    """One process, N GPUs, through include/sdrm/sdrm_multi.h: the job's channels are split over the device list by channel range,
    every device has its own host thread, stream set and pinned ring inside the library. The interface is host buffers, so the
    figure is end to end by construction (pinned input, H2D, chain, D2H of the int8 symbols)."""
    import torch
    import sdrm
    import workloads
    n_dev = args.gpus
    assert torch.cuda.device_count() >= n_dev, "needs %d visible GPUs" % n_dev
    n_ch, chunk = args.channels * n_dev, args.chunk
    cap = int(chunk / 20 * 1.1) + 64
    multi = sdrm.FskDemodMulti(list(range(n_dev)), n_ch, *shape.create_args, chunk, max_symbols_per_call=cap)
    pin_in = [sdrm.PinnedArray((n_ch, chunk), np.complex64) for _ in range(2)]
    for g in range(n_dev):
        dev = torch.device("cuda", g)
        x = workloads.gfsk_channels(args.channels, 2 * chunk, shape, seed=1000 + g * args.channels, device=dev)
        for i in range(2):
            torch.from_numpy(pin_in[i].array[g * args.channels:(g + 1) * args.channels].view(np.float32)).copy_(
                torch.view_as_real(x[:, i * chunk:(i + 1) * chunk].contiguous()).reshape(args.channels, 2 * chunk))
        del x
    for g in range(n_dev):
        torch.cuda.synchronize(g)
    pin_out = sdrm.PinnedArray((n_ch, cap), np.int8)
    pin_len = sdrm.PinnedArray((n_ch,), np.uint32)

    def loop(n):
        fetched = 0
        for k in range(n):
            multi.submit_ptr(pin_in[k % 2].ptr, chunk, chunk)
            if k + 1 >= 3:
                multi.fetch_ptr(pin_out.ptr, cap, pin_len.ptr)
                fetched += 1
        while fetched < n:
            multi.fetch_ptr(pin_out.ptr, cap, pin_len.ptr)
            fetched += 1

    steps = max(3, min(args.steps, 12))
    loop(max(3, args.warmup))
    multi.sync()
    launches_before = multi.launch_count
    t0 = time.perf_counter()
    loop(steps)
    multi.sync()
    sec = time.perf_counter() - t0
    launches = int(multi.launch_count - launches_before)
    value = n_ch * chunk * steps / sec / 1e6
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_dev, "steps": steps, "warmup": max(3, args.warmup),
            "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": dict(c2_config(args, shape, "exact"), launch="one process, sdrm_fsk_demod_multi_* over %d "
                                                "devices, one library thread per GPU" % n_dev),
            "gpu_launches": launches,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": n_ch * chunk * 8, "d2h_bytes_per_step": n_ch * cap + n_ch * 4,
                    "symbols_last_step": int(pin_len.array.sum())},
            "note": "value IS the end-to-end figure here: the multi-device entry points take host buffers",
            "shards": multi.shards(), "error_flags": multi.error_flags(), "lib": sdrm.version()}
    multi.close()
    emit_line(line, args.full_line)


def main():
    args = parse_args()
    claim_stdout()
    import workloads
    shape = c2_shape(args.chunk)
    if args.impl == "reference":
        run_reference_arm(args, shape)
        return
    if args.single_process:
        run_single_process(args, shape)
        return

    import torch
    import torch.distributed as dist
    import sdrm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    # host threads and the pinned buffers they allocate stay on the socket of this rank's GPU
    all_cpus = os.sched_getaffinity(0)
    local_cpus = sdrm.bind_thread_near_device(local_rank)
    numa_node = sdrm.lib.sdrm_device_numa_node(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # the FP32 pipe of this device, measured now with the FIR kernels' own instruction mixes (roofline denominator)
    fma_peak, pair_peak = sdrm.measure_fp32_peak(local_rank)

    n_ch, chunk = args.channels, args.chunk
    first_channel = rank * n_ch  # channel c of the job is the same signal on any sharding
    cap = int(chunk / shape.decimation / (shape.sampling_freq / shape.baud_rate / shape.decimation) * 1.1) + 64
    batch = sdrm.FskDemodBatch(n_ch, *shape.create_args, chunk, max_symbols_per_call=cap, fast=(args.mode == "fast"),
                               device=local_rank, measurement_aid=(sdrm.AID_NO_TAIL if args.debug_no_tail else 0))

    # two resident input buffers of 8 * n_ch * chunk bytes each (1 GiB at the defaults) >> 126 MB L2
    n_buf = 2
    t0 = time.time()
    stream_all = workloads.gfsk_channels(n_ch, n_buf * chunk, shape, seed=1000 + first_channel, device=device)
    bufs = [stream_all[:, i * chunk:(i + 1) * chunk].contiguous() for i in range(n_buf)]
    del stream_all
    torch.cuda.synchronize()
    gen_s = time.time() - t0

    fir_stream = torch.cuda.ExternalStream(batch.stream, device=device)
    out_stream = torch.cuda.ExternalStream(batch.out_stream, device=device)
    pin_out = [sdrm.PinnedArray((n_ch, cap), np.int8, device=local_rank) for _ in range(2)]
    pin_len = [sdrm.PinnedArray((n_ch,), np.uint32, device=local_rank) for _ in range(2)]
    depth = 3  # SDRM_MAX_IN_FLIGHT

    def device_loop(b, n_steps, first=0):
        """n_steps calls on resident input, results (int8 symbols + counts) fetched into pinned host memory: the serial tail of
        call k and the result copy of call k - 1 run under the filters of call k + 1"""
        fetched = 0
        for k in range(n_steps):
            b.process_device(bufs[(first + k) % n_buf].data_ptr(), chunk, chunk)
            if k + 1 >= depth:
                b.fetch_ptr(pin_out[fetched % 2].ptr, cap, pin_len[fetched % 2].ptr)
                fetched += 1
        while fetched < n_steps:
            b.fetch_ptr(pin_out[fetched % 2].ptr, cap, pin_len[fetched % 2].ptr)
            fetched += 1

    sampler = ClockSampler(local_rank)
    sampler.start()
    device_loop(batch, args.warmup)
    sampler.wait_first_sample()
    barrier()
    launches_before = batch.launch_count
    start = torch.cuda.Event(enable_timing=True)
    end = torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    start.record(fir_stream)
    device_loop(batch, args.steps, first=args.warmup)
    end.record(out_stream)  # the last fetch has returned: every kernel and result copy of the K steps is done
    barrier()
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end)
    ms_total = max_over_ranks(start.elapsed_time(end))
    launches = int(batch.launch_count - launches_before)
    symbols_last = int(pin_len[0].array.sum())
    fetched_columns = int(batch.last_fetch_columns())

    # per-kernel times of the dominant kernel, CUDA events on its own stream, inside a (second) timed loop
    batch.set_profiling(True)
    k1_ms, k3_ms, tail_ms, call_ms = [], [], [], []
    for k in range(min(args.steps, 5)):
        batch.process_device(bufs[k % n_buf].data_ptr(), chunk, chunk)
        batch.release()
        t = batch.stage_times()
        k1_ms.append(t[0]); k3_ms.append(t[1]); tail_ms.append(t[2]); call_ms.append(t[3])
    batch.set_profiling(False)
    flags = batch.error_flags()

    samples_per_step = world * n_ch * chunk
    value = samples_per_step * args.steps / (ms_total * 1e-3) / 1e6

    # ---- the same workload in the optional FMA arithmetic (one FFMA2 per tap, not the parity-defining mode) -----------------
    fma_mode = None
    parity_report = None
    if args.mode == "exact" and world == 1 and not args.no_fma_line:
        fast = sdrm.FskDemodBatch(n_ch, *shape.create_args, chunk, max_symbols_per_call=cap, fast=True, device=local_rank)
        fast_fir = torch.cuda.ExternalStream(fast.stream, device=device)
        fast_out = torch.cuda.ExternalStream(fast.out_stream, device=device)
        fast_steps = max(3, min(args.steps, 20))
        device_loop(fast, max(3, args.warmup))
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(fast_fir)
        device_loop(fast, fast_steps)
        f1.record(fast_out)
        torch.cuda.synchronize()
        fast_ms = f0.elapsed_time(f1) / fast_steps
        fast.set_profiling(True)
        fast_k1 = []
        for k in range(3):
            fast.process_device(bufs[k % n_buf].data_ptr(), chunk, chunk)
            fast.release()
            fast_k1.append(fast.stage_times()[0])
        fma_mode = {"value": n_ch * chunk / (fast_ms * 1e-3) / 1e6, "ms_per_step": fast_ms, "steps": fast_steps,
                    "kernel_ms": float(np.mean(fast_k1)), "error_flags": fast.error_flags()}
        fast.close()
        if not args.no_parity and not args.no_cpu:
            parity_report = fast_mode_parity(sdrm, shape, bufs, min(64, n_ch), cap, local_rank)

    # ---- end to end through the host-buffer entry points (pinned input, H2D + D2H in the timed region) ----------
    e2e = None
    if not args.no_e2e:
        # a finite run pays one un-overlapped call at its end (kernels + tail + result copy of the last call, ~10 ms): enough steps
        # that the figure is the sustained rate and not that edge (12 steps: 4 % of a cf32 run, 8 % of an int16 run)
        e2e_steps = max(3, args.steps) if args.steps < 5 else max(40, args.steps)
        pin_in = [sdrm.PinnedArray((n_ch, chunk), np.complex64, device=local_rank) for _ in range(2)]
        for i in range(2):
            torch.from_numpy(pin_in[i].array.view(np.float32).reshape(n_ch, 2 * chunk)).copy_(
                torch.view_as_real(bufs[i]).reshape(n_ch, 2 * chunk))
        torch.cuda.synchronize()

        def pipelined(submit, n_steps):
            """keep `depth` calls in flight (SDRM_MAX_IN_FLIGHT): the copy of call k+2 runs under the filters of call k+1"""
            fetched = 0
            for k in range(n_steps):
                submit(k)
                if k + 1 >= depth:
                    batch.fetch_ptr(pin_out[fetched % 2].ptr, cap, pin_len[fetched % 2].ptr)
                    fetched += 1
            while fetched < n_steps:
                batch.fetch_ptr(pin_out[fetched % 2].ptr, cap, pin_len[fetched % 2].ptr)
                fetched += 1

        def timed(loop):
            loop(3)
            barrier()
            t_start = time.perf_counter()
            loop(e2e_steps)
            torch.cuda.synchronize()
            return max_over_ranks(time.perf_counter() - t_start)

        e2e_s = timed(lambda n: pipelined(lambda k: batch.submit_ptr(pin_in[k % 2].ptr, chunk, chunk), n))
        symbols = int(pin_len[0].array.sum())
        # what the platform gives these N GPUs at once for the same buffers: plain pinned copies, nothing else running
        scratch = torch.empty(n_ch * chunk, dtype=torch.complex64, device=device)
        sdrm.probe_h2d(local_rank, pin_in[0].ptr, scratch.data_ptr(), n_ch * chunk * 8, 2)
        barrier()
        probe_reps = 6
        probe_s = sdrm.probe_h2d(local_rank, pin_in[0].ptr, scratch.data_ptr(), n_ch * chunk * 8, probe_reps)
        my_gbs = n_ch * chunk * 8 * probe_reps / probe_s / 1e9
        ceiling_gbs = world * n_ch * chunk * 8 * probe_reps / max_over_ranks(probe_s) / 1e9
        sum_gbs = sum_over_ranks(my_gbs)
        del scratch
        e2e_value = samples_per_step * e2e_steps / e2e_s / 1e6
        e2e = {"value": e2e_value, "unit": UNIT,
               "h2d_bytes_per_step": world * n_ch * chunk * 8,
               # what a fetch moves: the columns a call of this size can fill (the library's bound), not the rows' capacity
               "d2h_bytes_per_step": world * (n_ch * batch.last_fetch_columns() + n_ch * 4),
               "steps": e2e_steps, "symbols_last_step": symbols,
               "h2d_gbs": e2e_value * 8e6 / 1e9,
               "h2d_ceiling_gbs": ceiling_gbs, "h2d_ceiling_sum_of_ranks_gbs": sum_gbs,
               "frac_of_ceiling": e2e_value * 8e6 / 1e9 / ceiling_gbs,
               "h2d_ceiling_how": "plain cudaMemcpyAsync of the same pinned buffers (%d x %.2f GiB per rank), all %d ranks at once, "
                                  "no kernels; total bytes / slowest rank" % (probe_reps, n_ch * chunk * 8 / 2 ** 30, world),
               "pinned_numa_node": numa_node}
        # the same stream as 12-bit int16 IQ (what the PlutoSDR hands over, plutosdr.c:129), converted on the device:
        # half the host->device bytes. Reported beside the cf32 figure, which stays the headline (the reference API is cf32).
        pin_in16 = [sdrm.PinnedArray((n_ch, chunk, 2), np.int16, device=local_rank) for _ in range(2)]
        for i in range(2):
            q = torch.clamp(torch.round(torch.view_as_real(bufs[i]) * 1500.0), -2048, 2047).to(torch.int16)
            torch.from_numpy(pin_in16[i].array).copy_(q)
        torch.cuda.synchronize()
        e2e16_s = timed(lambda n: pipelined(lambda k: batch.submit_i16_ptr(pin_in16[k % 2].ptr, chunk, chunk), n))
        e2e16_value = samples_per_step * e2e_steps / e2e16_s / 1e6
        e2e["int16_value"] = e2e16_value
        e2e["int16_h2d_bytes_per_step"] = world * n_ch * chunk * 4
        e2e["int16_frac_of_ceiling"] = e2e16_value * 4e6 / 1e9 / ceiling_gbs
        e2e["int16_symbols_last_step"] = int(pin_len[0].array.sum())
        for a in pin_in + pin_in16:
            a.close()
    flags |= batch.error_flags()
    batch.close()
    del bufs
    torch.cuda.empty_cache()

    # ---- BASELINE configs[4] on 8 GPUs (all ranks), the other configs on rank 0 at N = 1 ----------------------------------------
    configs = []
    if not args.no_configs and world == 8:
        c5 = run_c5(args, world, rank, local_rank, device, shape, fma_peak, dist)
        configs.append(c5)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    if not args.no_configs and world == 1:
        os.sched_setaffinity(0, all_cpus)
        configs += run_other_configs(args, world, rank, local_rank, fma_peak)

    peaks, peaks_kind = measured_peaks()
    flops_per_sample, t1, t2 = workloads.demod_flops_per_sample(shape)
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    nominal_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    k1 = float(np.mean(k1_ms))
    k1_flops = 4.0 * t1 * n_ch * chunk  # 2 mul + 2 add per tap per complex sample
    achieved = k1_flops / (k1 * 1e-3) / 1e12
    in_bytes = 8.0 * n_ch * chunk
    step_ms = ms_total / args.steps
    roofline = {
        "bound": "fp32",
        "kernel": "fir_tile_kernel<1> (lpf1 + quadrature demod)",
        "achieved": achieved, "peak": fma_peak, "unit": "TFLOP/s", "frac": achieved / fma_peak,
        "peak_source": "measured",
        "peak_how": "sdrm_measure_fp32_peak on this device at start: FFMA2 stream, best of 5; nominal 148 SMs x 128 lanes x 2 x "
                    "%.0f MHz = %.2f" % (sm_max, nominal_peak),
        "peak_nominal": nominal_peak,
        "exact_mode_ceiling": pair_peak,
        "frac_of_exact_mode_ceiling": achieved / pair_peak if args.mode == "exact" else None,
        "exact_mode_ceiling_how": "same microbenchmark with the separately rounded multiply and add of exact mode (two FFMA2 per "
                                  "tap): what bit-exact arithmetic allows on this pipe",
        "chain_achieved": value / world * 1e6 * flops_per_sample / 1e12,
        "chain_frac": value / world * 1e6 * flops_per_sample / 1e12 / fma_peak,
        "kernel_ms": k1, "kernel_share_of_step": k1 / step_ms,
        "traffic": k1_traffic(n_ch * chunk),
        "algorithmic_bytes": in_bytes + in_bytes / 2,
        "hbm_achieved_gbs": (in_bytes + in_bytes / 2) / (k1 * 1e-3) / 1e9, "hbm_peak_gbs": peaks.get("hbm_gbs"),
        "lpf1_quad_ms": k1, "lpf2_ms": float(np.mean(k3_ms)), "dc_clock_tail_ms": float(np.mean(tail_ms)),
        "call_unpipelined_ms": float(np.mean(call_ms)),
    }
    if fma_mode is not None:
        fa = k1_flops / (fma_mode["kernel_ms"] * 1e-3) / 1e12
        roofline.update({
            "fma_note": "SDRM_FLAG_FAST_FMA (one FFMA2 per tap, same summation order): bit-identical to the FMA-order build of the "
                        "reference, NOT parity-defining; fma_parity_* count its deviation from the strict reference build",
            "fma_value": fma_mode["value"], "fma_ms_per_step": fma_mode["ms_per_step"], "fma_kernel_ms": fma_mode["kernel_ms"],
            "fma_achieved": fa, "fma_frac": fa / fma_peak, "fma_error_flags": fma_mode["error_flags"]})
        if parity_report is not None:
            c2 = parity_report["c2"]
            g = parity_report.get("goldens_all")
            roofline.update({
                "fma_parity_rule": c2["rule"],
                "fma_parity_c2_signal": c2["signal"],
                "fma_parity_c2_symbols": c2["symbols"],
                "fma_parity_c2_soft_over_1e-4": c2["soft_rel_over_1e-4"],
                "fma_parity_c2_soft_over_1e-4_strong": c2["soft_rel_over_1e-4_strong"],
                "fma_parity_c2_hard_flips": c2["hard_flips"],
                "fma_parity_c2_hard_flips_strong": c2["hard_flips_strong"],
                "fma_parity_c2_max_int8_delta": c2["max_int8_delta"],
                "fma_parity_c2_max_abs_dsoft": c2["max_abs_dsoft"],
                "fma_parity_c2_channels_bit_identical_to_fma_reference": c2["channels_bit_identical_to_fma_order_reference"],
                "fma_parity_c2_channels": c2["channels"]})
            if g is not None:
                roofline.update({
                    "fma_parity_goldens_symbols": g["symbols"],
                    "fma_parity_goldens_soft_over_1e-4_strong": g["soft_rel_over_1e-4_strong"],
                    "fma_parity_goldens_hard_flips_strong": g["hard_flips_strong"],
                    "fma_parity_goldens_max_int8_delta": g["max_int8_delta"],
                    "fma_parity_goldens_bit_identical_to_fma_reference": g["channels_bit_identical_to_fma_order_reference"],
                    "fma_parity_nodc_hard_flips_strong": parity_report["golden_lucky7_nodc"]["hard_flips_strong"],
                    "fma_parity_nodc_max_int8_delta": parity_report["golden_lucky7_nodc"]["max_int8_delta"]})
            roofline["fma_parity"] = parity_report
    cpu = None
    if not args.no_cpu and world == 1:  # the CPU baseline is reported at N = 1 only
        os.sched_setaffinity(0, all_cpus)  # the reference gets every host core, not only the GPU's socket
        from oracle import ref
        strict = cpu_reference_run(shape, args.cpu_seconds)
        if strict is not None:
            cpu = {k: strict[k] for k in ("value", "unit", "cores", "kind", "sample")}
            cpu["build"] = "strict (parity-defining)"
            if ref.available("tuned") and ref.tuned_level():
                tuned = cpu_reference_run(shape, args.cpu_seconds / 2, build="tuned")
                cpu["tuned_value"] = tuned["value"]
                cpu["tuned_sample"] = tuned["sample"]
    config = c2_config(args, shape, args.mode)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config,
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "value_includes": "int8 symbols + counts copied to pinned host memory every step (%d B per step)"
                          % (n_ch * fetched_columns + n_ch * 4),
        "symbols_last_step": symbols_last,
        "realtime_channels_per_gpu": value / world * 1e6 / shape.sampling_freq,
        "input_gen_s": gen_s, "host_cpus_near_gpu": local_cpus,
        "configs": configs, "error_flags": flags, "lib": sdrm.version(),
    }
    emit_line(line, args.full_line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
