#!/usr/bin/env python3
"""Measurement aid: time of the fused tail kernel with and without the dc blocker (C2 shape), CUDA events via stage_times."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sdr-modem_b200"))
import sdrm  # noqa: E402
import workloads  # noqa: E402

n_ch, chunk = 1024, 131072
for name, fs, baud in (("C2 (sps 10)", 192000, 9600), ("sps 5", 96000, 9600)):
    shape = workloads.DemodShape(name, fs, baud, 5000, 2, 2000, True, chunk)
    iq = workloads.gfsk_channels(n_ch, chunk, shape, seed=1000, device="cuda")
    for use_dc, flags in ((True, 0), (False, 0), (True, sdrm.AID_NO_CLOCK_LOOP)):
        b = sdrm.FskDemodBatch(n_ch, fs, baud, 5000, 2, 2000, use_dc, chunk, max_symbols_per_call=int(chunk / (fs // baud) * 1.2) + 64,
                               measurement_aid=flags)
        b.set_profiling(True)
        times = []
        for k in range(4):
            b.process_device(iq.data_ptr(), chunk, chunk)
            b.release()
            times.append(b.stage_times())
        print(name, "dc", use_dc, "clock loop off" if flags else "", "tail ms", [round(t[2], 3) for t in times], "K1", round(times[-1][0], 3),
              "K3", round(times[-1][1], 3))
        b.close()
