#!/usr/bin/env python3
"""Headline benchmark: batched GMSK 9600-baud demodulation (BASELINE.json configs[1]) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--channels C] [--mode exact|fast]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      the reference's own CPU chain (oracle/_ref) on the host cores

One step = one call of the demod chain (lpf1 -> quadrature demod -> lpf2 -> dc blocker -> clock recovery -> int8)
over `chunk` new samples of every channel. `value` counts input complex samples per second with the inputs resident
in HBM; `e2e` is the same through the host-buffer C-ABI call (pinned host input, H2D and D2H inside the timed region).
Channels are sharded across ranks (weak scaling: `--channels` per GPU), there is no collective on the data path.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
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
    p.add_argument("--cpu-seconds", type=float, default=12.0)
    p.add_argument("--debug-no-tail", action="store_true", help="measurement aid: skip the serial tail (invalid result)")
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
    path = os.path.join(ROOT, "profiles", "r1_k1_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f)
    per_sample = (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["samples_per_launch"]
    return per_sample * samples_per_launch


def cpu_reference_run(shape, seconds, n_threads=None):
    """The reference's own CPU chain (oracle/_ref, strict build) driven like dsp_worker: one pthread per channel."""
    import torch
    from oracle import ref
    import workloads
    if not ref.available():
        return None
    cores = n_threads or os.cpu_count() or 1
    n = shape.chunk * 2
    iq = workloads.gfsk_channels(cores, n, shape, seed=1000, device="cpu").numpy()
    # calibrate with one pass, then size the timed run
    sec, _ = ref.bench_fsk_demod(*shape.create_args, shape.chunk, iq, cores, passes=1)
    passes = max(1, int(seconds / max(sec, 1e-3)))
    sec, symbols = ref.bench_fsk_demod(*shape.create_args, shape.chunk, iq, cores, passes=passes)
    samples = cores * n * passes
    return {"value": samples / sec / 1e6, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "%d channels x %d samples x %d passes, chunk %d, %.1f s, oracle/_ref strict build (-O2 -ffp-contract=off, VOLK generic shim)"
                      % (cores, n, passes, shape.chunk, sec),
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


def emit_line(line):
    data = (json.dumps(line) + "\n").encode()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, data)


def run_reference_arm(args, shape):
    """bench.py --impl reference: the reference's own CPU chain (oracle/_ref, built from the reference sources in place) on all
    host cores, same metric and workload shape; a step is a bounded sample (a few passes of 16 x 2 chunks), so that
    K steps + W warm-ups end within a few minutes."""
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
    sec, _ = ref.bench_fsk_demod(*shape.create_args, shape.chunk, iq, cores, passes=1)
    target = max(0.2, args.cpu_seconds / max(1, args.steps))
    passes = max(1, int(round(target / max(sec, 1e-3))))
    for _ in range(args.warmup):
        ref.bench_fsk_demod(*shape.create_args, shape.chunk, iq, cores, passes=1)
    total_sec, total_samples = 0.0, 0
    for _ in range(args.steps):
        sec, _ = ref.bench_fsk_demod(*shape.create_args, shape.chunk, iq, cores, passes=passes)
        total_sec += sec
        total_samples += cores * n * passes
    value = total_samples / total_sec / 1e6
    sample = ("%d channels x %d samples x %d passes per step, chunk %d, oracle/_ref strict build (-O2 -ffp-contract=off, VOLK "
              "generic shim), one thread per channel" % (cores, n, passes, shape.chunk))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_sec / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%d channels x %s, GMSK BT 0.5, Eb/N0 12 dB, dev 5 kHz, decim 2, dc on (BASELINE configs[1], "
                                   "bounded sample)" % (cores, shape.name), "channels": cores, "chunk": shape.chunk},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_line(line)


def main():
    args = parse_args()
    claim_stdout()
    import workloads
    shape = workloads.DemodShape("gmsk9600@192k/chunk%d" % args.chunk, 192000, 9600, 5000, 2, 2000, True, args.chunk)
    if args.impl == "reference":
        run_reference_arm(args, shape)
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

    n_ch, chunk = args.channels, args.chunk
    first_channel = rank * n_ch  # channel c of the job is the same signal on any sharding
    cap = int(chunk / shape.decimation / (shape.sampling_freq / shape.baud_rate / shape.decimation) * 1.1) + 64
    batch = sdrm.FskDemodBatch(n_ch, *shape.create_args, chunk, max_symbols_per_call=cap, fast=(args.mode == "fast"),
                               device=local_rank, debug_flags=(0x80000000 if args.debug_no_tail else 0))

    # two resident input buffers of 8 * n_ch * chunk bytes each (1 GiB at the defaults) >> 126 MB L2
    n_buf = 2
    t0 = time.time()
    stream_all = workloads.gfsk_channels(n_ch, n_buf * chunk, shape, seed=1000 + first_channel, device=device)
    bufs = [stream_all[:, i * chunk:(i + 1) * chunk].contiguous() for i in range(n_buf)]
    del stream_all
    torch.cuda.synchronize()
    gen_s = time.time() - t0

    fir_stream = torch.cuda.ExternalStream(batch.stream, device=device)
    tail_stream = torch.cuda.ExternalStream(batch.tail_stream, device=device)

    def step(k):
        batch.process_device(bufs[k % n_buf].data_ptr(), chunk, chunk)
        batch.release()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for k in range(args.warmup):
        step(k)
    sampler.wait_first_sample()
    barrier()
    launches_before = batch.launch_count
    start = torch.cuda.Event(enable_timing=True)
    end = torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    start.record(fir_stream)
    for k in range(args.steps):
        step(args.warmup + k)
    end.record(tail_stream)
    barrier()
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end)
    ms_total = start.elapsed_time(end)
    launches = int(batch.launch_count - launches_before)

    # per-kernel times of the dominant kernel, CUDA events on its own stream, inside a (second) timed loop
    batch.set_profiling(True)
    k1_ms, k3_ms, tail_ms, call_ms = [], [], [], []
    for k in range(min(args.steps, 5)):
        step(k)
        t = batch.stage_times()
        k1_ms.append(t[0]); k3_ms.append(t[1]); tail_ms.append(t[2]); call_ms.append(t[3])
    batch.set_profiling(False)
    flags = batch.error_flags()

    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    samples_per_step = world * n_ch * chunk
    value = samples_per_step * args.steps / (ms_total * 1e-3) / 1e6

    # ---- the same workload in the optional FMA arithmetic (one FFMA2 per tap, not the parity-defining mode): shows what the
    # kernels reach once the bit-exact multiply-then-add no longer halves the pipe's useful rate -------------------------------
    fma_mode = None
    if args.mode == "exact" and world == 1 and not args.no_fma_line:
        fast = sdrm.FskDemodBatch(n_ch, *shape.create_args, chunk, max_symbols_per_call=cap, fast=True, device=local_rank)
        fast_fir = torch.cuda.ExternalStream(fast.stream, device=device)
        fast_tail = torch.cuda.ExternalStream(fast.tail_stream, device=device)
        fast_steps = max(3, min(args.steps, 20))
        for k in range(max(3, args.warmup)):
            fast.process_device(bufs[k % n_buf].data_ptr(), chunk, chunk)
            fast.release()
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(fast_fir)
        for k in range(fast_steps):
            fast.process_device(bufs[k % n_buf].data_ptr(), chunk, chunk)
            fast.release()
        f1.record(fast_tail)
        torch.cuda.synchronize()
        fast_ms = f0.elapsed_time(f1) / fast_steps
        fast.set_profiling(True)
        fast_k1 = []
        for k in range(3):
            fast.process_device(bufs[k % n_buf].data_ptr(), chunk, chunk)
            fast.release()
            fast_k1.append(fast.stage_times()[0])
        fma_mode = {"value": n_ch * chunk / (fast_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": fast_ms, "steps": fast_steps,
                    "kernel_ms": float(np.mean(fast_k1)), "error_flags": fast.error_flags(),
                    "note": "SDRM_FLAG_FAST_FMA: same summation order, fused rounding; bit-identical to the FMA-order build of the "
                            "reference, not to its strict build"}
        fast.close()

    # ---- end to end through the host-buffer entry points (pinned input, H2D + D2H in the timed region) ----------
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 12))
        pin_in = [sdrm.PinnedArray((n_ch, chunk), np.complex64) for _ in range(2)]
        pin_out = [sdrm.PinnedArray((n_ch, cap), np.int8) for _ in range(2)]
        pin_len = [sdrm.PinnedArray((n_ch,), np.uint32) for _ in range(2)]
        for i in range(2):
            torch.from_numpy(pin_in[i].array.view(np.float32).reshape(n_ch, 2 * chunk)).copy_(
                torch.view_as_real(bufs[i]).reshape(n_ch, 2 * chunk))
        torch.cuda.synchronize()

        def pipelined(submit, n_steps, depth=3):
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

        def e2e_loop(n_steps):
            pipelined(lambda k: batch.submit_ptr(pin_in[k % 2].ptr, chunk, chunk), n_steps)

        e2e_loop(3)
        barrier()
        t_start = time.perf_counter()
        e2e_loop(e2e_steps)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t_start
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        symbols = int(pin_len[0].array.sum())
        e2e = {"value": samples_per_step * e2e_steps / e2e_s / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": world * n_ch * chunk * 8, "d2h_bytes_per_step": world * (n_ch * cap + n_ch * 4),
               "steps": e2e_steps, "symbols_last_step": symbols}
        # the same stream as 12-bit int16 IQ (what the PlutoSDR hands over, plutosdr.c:129), converted on the device:
        # half the host->device bytes. Reported beside the cf32 figure, which stays the headline (the reference API is cf32).
        pin_in16 = [sdrm.PinnedArray((n_ch, chunk, 2), np.int16) for _ in range(2)]
        for i in range(2):
            q = torch.clamp(torch.round(torch.view_as_real(bufs[i]) * 1500.0), -2048, 2047).to(torch.int16)
            torch.from_numpy(pin_in16[i].array).copy_(q)
        torch.cuda.synchronize()

        def e2e16_loop(n_steps):
            pipelined(lambda k: batch.submit_i16_ptr(pin_in16[k % 2].ptr, chunk, chunk), n_steps)

        e2e16_loop(3)
        barrier()
        t_start = time.perf_counter()
        e2e16_loop(e2e_steps)
        torch.cuda.synchronize()
        e2e16_s = time.perf_counter() - t_start
        if world > 1:
            t = torch.tensor([e2e16_s], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e16_s = float(t.item())
        e2e["int16_ingest"] = {"value": samples_per_step * e2e_steps / e2e16_s / 1e6, "unit": UNIT,
                               "h2d_bytes_per_step": world * n_ch * chunk * 4, "symbols_last_step": int(pin_len[0].array.sum())}
        for a in pin_in + pin_out + pin_len + pin_in16:
            a.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peaks_kind = measured_peaks()
    flops_per_sample, t1, t2 = workloads.demod_flops_per_sample(shape)
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    fp32_peak_tflops = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    k1 = float(np.mean(k1_ms))
    k1_flops = 4.0 * t1 * n_ch * chunk  # 2 mul + 2 add per tap per complex sample
    achieved = k1_flops / (k1 * 1e-3) / 1e12
    in_bytes = 8.0 * n_ch * chunk
    roofline = {
        "bound": "fp32",
        "kernel": "fir_tile_kernel<1> (lpf1 + quadrature demod)",
        "achieved": achieved, "peak": fp32_peak_tflops, "unit": "TFLOP/s", "frac": achieved / fp32_peak_tflops,
        "peak_source": "148 SMs x 128 FP32 lanes x 2 flop x %.0f MHz (sm_max_mhz, %s); FMA peak — exact mode issues a separately "
                       "rounded multiply and add per tap, so its own ceiling is 0.5" % (sm_max, peaks_kind),
        "pipe_frac": achieved / fp32_peak_tflops * (2.0 if args.mode == "exact" else 1.0),
        "kernel_ms": k1, "kernel_share_of_step": k1 / (ms_total / args.steps),
        "traffic": k1_traffic(n_ch * chunk),
        "hbm": {"achieved": (in_bytes + in_bytes / 2) / (k1 * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s"},
        "fma_mode": None if fma_mode is None else dict(
            fma_mode, achieved=k1_flops / (fma_mode["kernel_ms"] * 1e-3) / 1e12,
            frac=k1_flops / (fma_mode["kernel_ms"] * 1e-3) / 1e12 / fp32_peak_tflops),
        "stage_ms": {"lpf1_quad": k1, "lpf2": float(np.mean(k3_ms)), "dc_clock_tail": float(np.mean(tail_ms)),
                     "call_unpipelined": float(np.mean(call_ms))},
    }
    cpu = None
    if not args.no_cpu and world == 1:  # the CPU baseline is reported at N = 1 only
        os.sched_setaffinity(0, all_cpus)  # the reference gets every host core, not only the GPU's socket
        cpu = cpu_reference_run(shape, args.cpu_seconds)
        if cpu is not None:
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%d channels/GPU x %s, GMSK BT 0.5, Eb/N0 12 dB, dev 5 kHz, decim 2, dc on (BASELINE configs[1])"
                               % (n_ch, shape.name),
                   "channels_per_gpu": n_ch, "chunk": chunk, "mode": args.mode, "parallelism": "channels sharded, no collective",
                   "l2": "inputs larger than L2 (2 x %.2f GiB rotating)" % (n_ch * chunk * 8 / 2 ** 30),
                   "realtime_channels_per_gpu": value / world * 1e6 / shape.sampling_freq,
                   "flop_per_sample": flops_per_sample, "t1": t1, "t2": t2, "input_gen_s": gen_s,
                   "host_cpus_near_gpu": local_cpus},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "error_flags": flags, "lib": sdrm.version(),
    }
    emit_line(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
