#!/usr/bin/env python3
"""Extracts the inline known-answer arrays of the reference's own DSP unit tests into reference_kats.json.

Run in the build container (needs /root/reference); the JSON is committed because the reference tree does not exist
on the GPU box. Keys are "<test file>:<START_TEST name>:<array name>"; values are the float literals as written.
The scenarios that produce these arrays (inputs, parameters, call sequence) are restated in tests/test_reference_kats.py
with file:line citations. The binary golden files next to this script are byte copies of reference test/resources/*.
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ["test_lpf.c", "test_lpf_taps.c", "test_quadrature_demod.c", "test_dc_blocker.c", "test_clock_recovery_mm.c",
         "test_sig_source.c", "test_gaussian_taps.c", "test_interp_fir_filter.c", "test_frequency_modulator.c",
         "test_gfsk_mod.c"]


def main():
    out = {}
    for name in FILES:
        src = open(os.path.join(REF, "test", name)).read()
        for m in re.finditer(r"START_TEST\s*\(\s*(\w+)\s*\)(.*?)END_TEST", src, re.S):
            test, body = m.group(1), m.group(2)
            for a in re.finditer(r"(?:const\s+)?float\s+(\w+)\s*\[\s*\d*\s*\]\s*=\s*\{([^}]*)\}", body):
                vals = [float(t.strip().rstrip("fF")) for t in a.group(2).split(",") if t.strip()]
                out["%s:%s:%s" % (name, test, a.group(1))] = vals
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    for k, v in out.items():
        print(k, len(v))


if __name__ == "__main__":
    main()
