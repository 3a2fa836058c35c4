#!/usr/bin/env python3
"""Measurement aid: what the first *_create / *_process of a process costs (CUDA context, module load, tables), and the second."""
import ctypes as C
import os
import sys
import time

t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sdr-modem_b200"))
import numpy as np  # noqa: E402
import sdrm  # noqa: E402

print("import sdrm %.3f s" % (time.time() - t0))
for k in range(3):
    t = time.time()
    d = sdrm.FskDemod(48000, 4800, 5000, 2, 2000, True, 2048)
    t_create = time.time() - t
    t = time.time()
    d.process(np.zeros(2048, np.complex64))
    t_first = time.time() - t
    t = time.time()
    d.process(np.zeros(2048, np.complex64))
    t_second = time.time() - t
    d.close()
    print("handle %d: create %.3f s, first process %.3f s, second process %.4f s" % (k, t_create, t_first, t_second))
