#!/usr/bin/env python3
"""Measurement aid: steady-state ms per call of the C2 demod with the serial tail (dc blocker on / off) overlapping the next
call's filters, against the filters alone: what the tail costs the filters it runs beside."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sdr-modem_b200"))
import sdrm  # noqa: E402
import workloads  # noqa: E402

n_ch, chunk, steps = 1024, 131072, 30
shape = workloads.C2_THROUGHPUT
iq = workloads.gfsk_channels(n_ch, 2 * chunk, shape, seed=1000, device="cuda")
bufs = [iq[:, :chunk].contiguous(), iq[:, chunk:].contiguous()]
for label, use_dc, flags in (("dc on", True, 0), ("dc off", False, 0), ("no tail", True, sdrm.AID_NO_TAIL)):
    b = sdrm.FskDemodBatch(n_ch, 192000, 9600, 5000, 2, 2000, use_dc, chunk, max_symbols_per_call=int(chunk / 20 * 1.2) + 64,
                           measurement_aid=flags)
    fir = torch.cuda.ExternalStream(b.stream)
    tail = torch.cuda.ExternalStream(b.tail_stream)
    for k in range(3):
        b.process_device(bufs[k % 2].data_ptr(), chunk, chunk)
        b.release()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(fir)
    for k in range(steps):
        b.process_device(bufs[k % 2].data_ptr(), chunk, chunk)
        b.release()
    end.record(tail)
    torch.cuda.synchronize()
    print(label, "ms per call %.3f" % (start.elapsed_time(end) / steps))
    b.close()

# K1 / K3 of a call whose filters run beside the previous call's tail, against the same kernels alone
b = sdrm.FskDemodBatch(n_ch, 192000, 9600, 5000, 2, 2000, True, chunk, max_symbols_per_call=int(chunk / 20 * 1.2) + 64)
b.set_profiling(True)
for k in range(3):
    b.process_device(bufs[k % 2].data_ptr(), chunk, chunk)
    b.release()
    alone = b.stage_times()
for k in range(6):
    b.process_device(bufs[k % 2].data_ptr(), chunk, chunk)
    b.release()
beside = b.stage_times()
print("alone   K1 %.3f K3+hist %.3f tail %.3f" % (alone[0], alone[1], alone[2]))
print("beside  K1 %.3f K3+hist %.3f tail %.3f" % (beside[0], beside[1], beside[2]))
b.close()
