"""Synthetic workloads for the demod / mod hot path, shared by tests/ and bench.py.

Only input synthesis lives here (torch is used as an array library so that the same code fills HBM directly for
the bench and runs on the CPU for small parity cases). Nothing in this file is on the measured path.
"""
import math
from dataclasses import dataclass

import numpy as np
import torch


@dataclass(frozen=True)
class DemodShape:
    """Parameters of fsk_demod_create (reference src/dsp/fsk_demod.c:28) plus the call size."""
    name: str
    sampling_freq: int
    baud_rate: int
    deviation: int
    decimation: int
    transition_width: int
    use_dc_block: bool
    chunk: int

    @property
    def create_args(self):
        return (self.sampling_freq, self.baud_rate, self.deviation, self.decimation, self.transition_width,
                self.use_dc_block)

    @property
    def samples_per_symbol_in(self):
        return self.sampling_freq // self.baud_rate


# BASELINE.json configs[0..1,4]: GMSK 9600 baud at 192 ksps (T1 = 471, T2 = 231 taps)
C2_PARITY = DemodShape("gmsk9600@192k/chunk4096", 192000, 9600, 5000, 2, 2000, True, 4096)
C2_THROUGHPUT = DemodShape("gmsk9600@192k/chunk131072", 192000, 9600, 5000, 2, 2000, True, 131072)
# the reference's own perf_fsk_modem shape (test/perf_fsk_modem.c:72,76)
PERF_SHAPE = DemodShape("perf_fsk_modem 4800@48k/chunk4096", 48000, 4800, 5000, 2, 2000, True, 4096)


def lowpass_taps_count(fs, tw):
    return int(53.0 * fs / (22.0 * tw)) | 1


def demod_flops_per_sample(shape):
    """Algorithmic flop per input sample, SURVEY.md §8d: 4*T1 + 2*T2/D + 15 + 14/D + 31/(D*sps)."""
    carson = abs(shape.deviation) + shape.baud_rate / 2
    t1 = lowpass_taps_count(shape.sampling_freq, int(np.float32(0.1) * carson))
    t2 = lowpass_taps_count(shape.sampling_freq, shape.transition_width)
    d = shape.decimation
    sps = shape.sampling_freq / shape.baud_rate / d
    return 4 * t1 + 2 * t2 / d + 15 + 14 / d + 31 / (d * sps), t1, t2


def gaussian_pulse(sps, bt=0.5, span=4):
    n = span * sps
    t = (np.arange(n) - (n - 1) / 2) / sps
    s = 2 * math.pi * bt / math.sqrt(math.log(2.0))
    h = np.exp(-0.5 * (s * t) ** 2)
    return np.convolve(h / h.sum(), np.ones(sps))  # an impulse train of +-1 every sps samples settles at +-1


def gfsk_channels(n_channels, n_samples, shape, seed, device="cpu", eb_n0_db=12.0, max_offset_hz=1500.0,
                  dtype=torch.complex64):
    """cf32 [n_channels, n_samples]: per-channel random payload, GFSK BT=0.5, carrier and timing offsets, AWGN.

    Channel c depends only on (seed, c) and the device type, whatever n_channels is, so shards of a bigger job
    see the same data.
    """
    sps = shape.samples_per_symbol_in
    dev = torch.device(device)
    n_sym = (n_samples + sps - 1) // sps + 8
    pulse = torch.tensor(gaussian_pulse(sps), dtype=torch.float32, device=dev)
    out = torch.empty((n_channels, n_samples), dtype=dtype, device=dev)
    block = 64 if dev.type == "cuda" else 8
    sens = 2 * math.pi * shape.deviation / shape.sampling_freq
    sigma = math.sqrt(sps / (2 * 10 ** (eb_n0_db / 10)))  # per real component, unit signal amplitude
    n_idx = torch.arange(n_samples, device=dev, dtype=torch.float64)
    for c0 in range(0, n_channels, block):
        c1 = min(n_channels, c0 + block)
        rows = []
        for c in range(c0, c1):
            g = torch.Generator(device="cpu")
            g.manual_seed(int(seed) * 1000003 + c)
            bits = torch.randint(0, 2, (n_sym,), generator=g, dtype=torch.int64)
            extras = torch.rand(2, generator=g, dtype=torch.float64)
            rows.append((bits, extras))
        nrz = torch.stack([r[0] for r in rows]).to(dev).to(torch.float32) * 2 - 1
        up = torch.zeros((c1 - c0, n_sym * sps), dtype=torch.float32, device=dev)
        up[:, ::sps] = nrz
        shaped = torch.nn.functional.conv1d(up[:, None, :], pulse.flip(0)[None, None, :], padding=pulse.numel() - 1)[:, 0, :]
        phase = torch.cumsum(shaped.to(torch.float64) * sens, dim=1)
        f_off = torch.stack([(r[1][0] * 2 - 1) * max_offset_hz for r in rows]).to(dev)
        t_off = torch.stack([torch.floor(r[1][1] * sps) for r in rows]).to(dev).to(torch.int64)
        idx = torch.arange(n_samples, device=dev)[None, :] + t_off[:, None] + 2 * sps
        ph = torch.gather(phase, 1, idx) + 2 * math.pi * f_off[:, None] / shape.sampling_freq * n_idx[None, :]
        gn = torch.Generator(device=dev)
        noise = torch.empty((c1 - c0, n_samples, 2), device=dev, dtype=torch.float32)
        for c in range(c0, c1):
            gn.manual_seed(int(seed) * 7919 + c)
            noise[c - c0] = torch.randn((n_samples, 2), generator=gn, device=dev, dtype=torch.float32)
        noise *= sigma
        sig = torch.stack([torch.cos(ph), torch.sin(ph)], dim=-1).to(torch.float32) + noise
        out[c0:c1] = torch.view_as_complex(sig.contiguous())
    return out


def xorshift_bytes(n, seed):
    """Deterministic payload bytes (xorshift32), used for the modulator workloads."""
    out = np.empty(n, dtype=np.uint8)
    x = np.uint32(seed if seed != 0 else 1)
    with np.errstate(over="ignore"):
        for i in range(n):
            x ^= np.uint32(x << np.uint32(13))
            x ^= np.uint32(x >> np.uint32(17))
            x ^= np.uint32(x << np.uint32(5))
            out[i] = np.uint8(x & np.uint32(0xFF))
    return out
