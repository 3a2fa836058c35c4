"""TEST INFRASTRUCTURE ONLY: the accounting of the north-star epsilon rule for the library's optional FMA ("fast") mode.

The parity-defining mode of the library is exact mode (bit-identical to the strict reference build, tests/test_gpu_demod.py).
Fast mode fuses the multiply and the add of every FIR tap; it is bit-identical to the reference sources compiled with
fmaf dot products (oracle/_ref/libsdrmodem_ref_fma.so) but NOT to the strict build: the rounding differences (1e-7 relative in
the filter outputs) pass through the dc blocker's running sums and the Mueller & Mueller feedback loop, which amplify them.
SURVEY.md section 8(d) "Parity rule" asks for the violations of the rule to be counted and reported, not hidden:

    soft symbols within 1e-4 relative of the strict reference            abs(d soft) <= REL * abs(soft_ref)
    hard bits identical except where abs(soft_ref) < EPS                 EPS = 5e-3 (0.6 int8 LSB)

`compare` produces those counts for one channel, `merge` adds reports up, `reference_soft_symbols` runs the strict (or the
FMA-order) reference build over a set of channels on all host cores. Callers: tests/test_gpu_fast_mode_parity.py and the
`roofline.fma_mode.parity` block of bench.py (rank 0, N = 1, outside every timed region; checker only).
"""
import concurrent.futures as cf
import os

import numpy as np

from . import ref

REL = 1e-4
EPS = 5e-3


def reference_soft_symbols(args, iq_channels, chunk, fma=False, threads=None):
    """[(hard int8, soft float32)] per channel from oracle/_ref (strict build, or the FMA-order build), one thread per channel
    (the reference's C code runs outside the GIL)."""
    iq_channels = np.ascontiguousarray(iq_channels, dtype=np.complex64)

    def one(c):
        r = ref.fsk_chain(*args, iq_channels[c], chunk, fma=fma)
        return r["hard"], r["soft"]

    with cf.ThreadPoolExecutor(threads or os.cpu_count() or 1) as pool:
        return list(pool.map(one, range(iq_channels.shape[0])))


def compare(hard, soft, ref_hard, ref_soft):
    """Counts for one channel: `hard`/`soft` from the mode under test, `ref_*` from the strict reference build."""
    n = min(len(soft), len(ref_soft))
    report = {"channels": 1, "symbols": int(n), "length_mismatch": int(len(soft) != len(ref_soft))}
    a = np.asarray(soft[:n], dtype=np.float64)
    b = np.asarray(ref_soft[:n], dtype=np.float64)
    finite = np.isfinite(a) & np.isfinite(b)
    report["non_finite_mismatch"] = int(np.count_nonzero(np.isfinite(a) != np.isfinite(b)))
    d = np.where(finite, np.abs(a - b), 0.0)
    mag = np.where(finite, np.abs(b), 0.0)
    over = d > REL * mag
    strong = mag >= EPS
    flips = finite & ((a < 0) != (b < 0))
    report["bit_identical_soft"] = int(np.count_nonzero(np.asarray(soft[:n]).view(np.uint32) == np.asarray(ref_soft[:n]).view(np.uint32)))
    report["soft_rel_over_1e-4"] = int(np.count_nonzero(over))
    report["soft_rel_over_1e-4_strong"] = int(np.count_nonzero(over & strong))
    report["hard_flips"] = int(np.count_nonzero(flips))
    report["hard_flips_strong"] = int(np.count_nonzero(flips & strong))
    report["max_abs_dsoft"] = float(d.max()) if n else 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(strong, d / mag, 0.0)
    report["max_rel_dsoft_strong"] = float(rel.max()) if n else 0.0
    h = np.abs(np.asarray(hard[:n], dtype=np.int64) - np.asarray(ref_hard[:n], dtype=np.int64))
    report["int8_differ"] = int(np.count_nonzero(h))
    report["max_int8_delta"] = int(h.max()) if n else 0
    return report


_SUM = ("channels", "symbols", "length_mismatch", "non_finite_mismatch", "bit_identical_soft", "soft_rel_over_1e-4",
        "soft_rel_over_1e-4_strong", "hard_flips", "hard_flips_strong", "int8_differ")
_MAX = ("max_abs_dsoft", "max_rel_dsoft_strong", "max_int8_delta")


def merge(reports):
    out = {k: 0 for k in _SUM}
    out.update({k: 0 for k in _MAX})
    for r in reports:
        for k in _SUM:
            out[k] += r[k]
        for k in _MAX:
            out[k] = max(out[k], r[k])
    return out


def same_bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype == np.float32:
        return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))
    return bool(np.array_equal(a, b))


def fast_mode_report(gpu_results, args, iq_channels, chunk, threads=None):
    """gpu_results: [(hard, soft)] per channel from the library in fast mode on `iq_channels` in `chunk`-sized calls.
    Returns the merged counts against the strict reference build plus how many channels are bit-identical to the
    FMA-order reference build (None when that build is absent)."""
    strict = reference_soft_symbols(args, iq_channels, chunk, fma=False, threads=threads)
    report = merge([compare(g[0], g[1], s[0], s[1]) for g, s in zip(gpu_results, strict)])
    report["rule"] = "abs(dsoft) <= %g * abs(soft_ref); hard bits equal where abs(soft_ref) >= %g" % (REL, EPS)
    report["checker"] = "oracle/_ref strict build (-O2 -ffp-contract=off)"
    if ref.available(fma=True):
        fused = reference_soft_symbols(args, iq_channels, chunk, fma=True, threads=threads)
        report["channels_bit_identical_to_fma_order_reference"] = int(sum(
            same_bits(g[0], f[0]) and same_bits(g[1], f[1]) for g, f in zip(gpu_results, fused)))
    else:
        report["channels_bit_identical_to_fma_order_reference"] = None
    return report
