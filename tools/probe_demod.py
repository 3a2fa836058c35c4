"""GPU probe: fsk demod chain vs the reference-built oracle on the golden inputs. Prints diagnostics."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sdr-modem_b200"))
import sdrm
from oracle import ref
G = os.path.join(ROOT, "tests", "golden")
print(sdrm.version())

def compare(name, iq, params, chunk, n_ch=1, fast=False):
    fs, baud, dev, dec, tw, dc = params
    o = ref.fsk_chain(fs, baud, dev, dec, tw, dc, iq, chunk, fma=fast)
    b = sdrm.FskDemodBatch(n_ch, fs, baud, dev, dec, tw, dc, chunk, soft=True, fast=fast)
    x = np.tile(iq[None, :], (n_ch, 1))
    t = time.time()
    hard, soft = b.run_stream(x, chunk)
    dt = time.time() - t
    flags = b.error_flags()
    ok = True
    for c in range(n_ch):
        same_len = len(hard[c]) == len(o["hard"])
        n = min(len(hard[c]), len(o["hard"]))
        hd = np.abs(hard[c][:n].astype(int) - o["hard"][:n].astype(int))
        sb = soft[c][:n].view(np.uint32) != o["soft"][:n].view(np.uint32)
        if c == 0 or not same_len or hd.max(initial=0) or sb.any():
            print(f"{name} ch{c}: gpu {len(hard[c])} ref {len(o['hard'])} hard maxdiff {hd.max(initial=0)} ndiff {(hd>0).sum()} soft bitdiff {sb.sum()} first {np.argmax(sb) if sb.any() else -1} flags {flags} t {dt:.3f}s")
        ok &= same_len and hd.max(initial=0) == 0 and not sb.any()
        if c > 2 and not ok: break
    b.close()
    return ok

lucky = np.fromfile(os.path.join(G, "lucky7.expected.cf32"), dtype=np.complex64)
nusat = np.fromfile(os.path.join(G, "nusat.cf32"), dtype=np.complex64)
nan = np.fromfile(os.path.join(G, "inputnan.cf32"), dtype=np.complex64)
res = {}
res["lucky_dc"] = compare("lucky7 dc", lucky, (48000, 4800, 5000, 2, 2000, True), 4096)
res["lucky_nodc"] = compare("lucky7 nodc", lucky, (48000, 4800, 5000, 2, 2000, False), 4096)
res["nusat"] = compare("nusat", nusat, (192000, 40000, 5000, 1, 2000, True), 4096)
res["nan"] = compare("nan", nan, (240000, 9600, 5000, 1, 2000, True), 4096)
res["lucky_dc_3ch"] = compare("lucky7 dc x3", lucky, (48000, 4800, 5000, 2, 2000, True), 4096, n_ch=3)
res["lucky_odd_chunks"] = compare("lucky7 chunk 4001", lucky, (48000, 4800, 5000, 2, 2000, True), 4001)
res["lucky_big_chunk"] = compare("lucky7 chunk 50000", lucky, (48000, 4800, 5000, 2, 2000, True), 50000)
res["c2_shape"] = compare("lucky7 as 192k/9600", lucky, (192000, 9600, 5000, 2, 2000, True), 8192)
res["lucky_fast"] = compare("lucky7 dc fast-vs-fma-oracle", lucky, (48000, 4800, 5000, 2, 2000, True), 4096, fast=True)
res["dec3"] = compare("lucky7 dec 3", lucky, (48000, 2400, 5000, 3, 2000, True), 4096)
# drop-in single handle
d = sdrm.FskDemod(48000, 4800, 5000, 2, 2000, True, 4096)
out = d.run(lucky, 4096); d.close()
o = ref.fsk_demod_run(48000, 4800, 5000, 2, 2000, True, lucky, 4096)
res["dropin"] = bool(len(out) == len(o) and np.array_equal(out, o))
print(res)
print("ALL OK" if all(res.values()) else "MISMATCH")
