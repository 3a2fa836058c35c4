#!/usr/bin/env python3
"""Measurement aid: a few calls of the C2 demodulator so that ncu can capture the fused tail kernel (tools: ncu -k regex:demod_tail)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sdr-modem_b200"))
import sdrm  # noqa: E402
import workloads  # noqa: E402

fs, baud = (192000, 9600) if len(sys.argv) < 2 or sys.argv[1] != "sps5" else (96000, 9600)
n_ch, chunk = 1024, 131072
shape = workloads.DemodShape("probe", fs, baud, 5000, 2, 2000, True, chunk)
iq = workloads.gfsk_channels(n_ch, chunk, shape, seed=1000, device="cuda")
b = sdrm.FskDemodBatch(n_ch, fs, baud, 5000, 2, 2000, True, chunk, max_symbols_per_call=int(chunk / (fs // baud) * 1.2) + 64)
for k in range(3):
    b.process_device(iq.data_ptr(), chunk, chunk)
    b.release()
b.sync()
b.close()
