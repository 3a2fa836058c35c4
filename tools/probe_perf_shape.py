import os, sys
import torch
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sdr-modem_b200"))
import sdrm, workloads
n_ch, chunk, steps = 1024, 131072, 20
shape = workloads.DemodShape("perf", 48000, 4800, 5000, 2, 2000, True, chunk)
iq = workloads.gfsk_channels(n_ch, 2 * chunk, shape, seed=1000, device="cuda")
bufs = [iq[:, :chunk].contiguous(), iq[:, chunk:].contiguous()]
b = sdrm.FskDemodBatch(n_ch, *shape.create_args, chunk, max_symbols_per_call=int(chunk / 10 * 1.2) + 64)
fir = torch.cuda.ExternalStream(b.stream); tail = torch.cuda.ExternalStream(b.tail_stream)
for k in range(3):
    b.process_device(bufs[k % 2].data_ptr(), chunk, chunk); b.release()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(fir)
for k in range(steps):
    b.process_device(bufs[k % 2].data_ptr(), chunk, chunk); b.release()
e.record(tail); torch.cuda.synchronize()
ms = s.elapsed_time(e) / steps
b.set_profiling(True)
b.process_device(bufs[0].data_ptr(), chunk, chunk); b.release()
print("perf shape ms/call %.3f -> %.1f Gsamples/s; stages" % (ms, n_ch * chunk / ms / 1e6), [round(x, 3) for x in b.stage_times()], "flags", b.error_flags())
