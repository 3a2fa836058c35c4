#!/usr/bin/env python3
"""Measurement aid (GPU box, best with two or more GPUs): the multi-device entry points against the oracle in a loop, first with
shards on one device, then across all devices, in one process. This sequence exposed the asynchronous cudaMemset race that
sdrm_dev_zalloc now closes (a shard's first call after a second device's context came up).  usage: tools/stress_multi.py [iterations]"""
import sys, os
ROOT="/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+"/sdr-modem_b200")
import numpy as np
import sdrm, workloads
from oracle import port
port.load()
shape = workloads.C2_PARITY
n_ch, chunk, calls, cap = 7, 4096, 5, 256
iq = workloads.gfsk_channels(n_ch, chunk * calls, shape, seed=21).numpy()
want = [port.FskDemod(*shape.create_args, chunk) for _ in range(n_ch)]
want_out = [[want[c].process(iq[c, k*chunk:(k+1)*chunk]) for k in range(calls)] for c in range(n_ch)]
import torch
devs = list(range(torch.cuda.device_count()))
bad_single = bad_multi = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    # single
    b = sdrm.FskDemodBatch(n_ch, *shape.create_args, chunk, max_symbols_per_call=cap, soft=True, device=0)
    for k in range(calls):
        hard, lens, soft = b.process(iq[:, k*chunk:(k+1)*chunk])
        for c in range(n_ch):
            if lens[c] != len(want_out[c][k][0]) or not np.array_equal(hard[c,:lens[c]], want_out[c][k][0]):
                bad_single += 1; print("single bad it", it, "call", k, "ch", c, lens[c], len(want_out[c][k][0]))
    b.close()
    for dl in ([0, 0], [0, 0, 0], devs if len(devs) > 1 else [0, 0]):
        m = sdrm.FskDemodMulti(dl, n_ch, *shape.create_args, chunk, max_symbols_per_call=cap, soft=True)
        got = []
        keep = []
        for k in range(calls):
            part = np.ascontiguousarray(iq[:, k*chunk:(k+1)*chunk]); keep.append(part)
            m.submit(part)
            if k >= 2: got.append(m.fetch())
        while len(got) < calls: got.append(m.fetch())
        m.close()
        for k in range(calls):
            hard, lens, soft = got[k]
            for c in range(n_ch):
                if lens[c] != len(want_out[c][k][0]) or not np.array_equal(hard[c,:lens[c]], want_out[c][k][0]):
                    bad_multi += 1; print("multi bad it", it, "devices", dl, "call", k, "ch", c, lens[c], len(want_out[c][k][0]))
print("bad_single", bad_single, "bad_multi", bad_multi)
