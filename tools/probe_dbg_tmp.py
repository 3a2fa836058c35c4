import os, sys, ctypes as C
import numpy as np, torch
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sdr-modem_b200"))
import sdrm, workloads
n_ch, chunk = 1024, 131072
shape = workloads.C2_THROUGHPUT
iq = workloads.gfsk_channels(n_ch, chunk, shape, seed=1000, device="cuda")
b = sdrm.FskDemodBatch(n_ch, 192000, 9600, 5000, 2, 2000, True, chunk, max_symbols_per_call=int(chunk / 20 * 1.2) + 64)
b.set_profiling(True)
for k in range(3):
    b.process_device(iq.data_ptr(), chunk, chunk); b.release(); t = b.stage_times()
print("tail ms", t[2])
out = (C.c_longlong * 64)()
sdrm.lib.sdrm_cu_tail_debug_read(out)
names = ["loop top/other", "fetch admin", "mbar wait", "producer block", "bulk wait", "clock/idle", "barrier"]
arr = np.array(list(out)).reshape(8, 8)
for w in range(5):
    print("warp", w, {names[i]: int(arr[w, i] / 2048) for i in range(7)}, "sum", int(arr[w, :7].sum() / 2048))
