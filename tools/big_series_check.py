#!/usr/bin/env python3
"""One series longer than 2^31 samples (8.6 GB of float32): 64-bit offsets inside a series, every kernel family, checked
against the oracle on the head, the 2^31 crossing and the tail.  (development aid; needs ~20 GB of HBM)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import windows as OW, spectral as OS
from pymhealth_b200 import engine, spectral as SP
from pymhealth_b200.generic import stats, timedom

W, S, fs = 500, 250, 50.0
n = (1 << 31) + 250 * 4001 + 77
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(3)
x = torch.empty(n, dtype=torch.float32, device=dev)
step = 1 << 28
for a in range(0, n, step):
    b = min(n, a + step)
    x[a:b] = torch.randn(b - a, generator=g, device=dev) * 0.3 + torch.sin(torch.arange(a, b, device=dev, dtype=torch.float64) * 0.05).float()
nw = engine.n_windows(n, W, S)
feats = [stats.mean.feature(), stats.std.feature(), stats.dmax.feature(), stats.kurtosis.feature(),
         timedom.zero_crossing_count.feature(), timedom.line_length.feature(), stats.median.feature(),
         stats.percentile.feature(90.0), SP.total_power(fs).feature(), SP.band_power(fs, 0.5, 3.0).feature(),
         SP.peak_bin(fs, 0.3, 12.0).feature(), SP.spectral_entropy(fs).feature()]
names = ["mean", "std", "max", "kurtosis", "zero_crossing_count", "line_length", "median", "percentile"]
tab = engine.window_table(x.unsqueeze(0), W, S, feats, fs=fs, out_dtype=torch.float64)[0]
torch.cuda.synchronize()
print("n = %d samples, %d windows" % (n, nw))
bad = 0
cross = ((1 << 31) // S) - 3
for w0 in (0, cross, nw - 40):
    cnt = 40 if w0 + 40 <= nw else nw - w0
    seg = x[w0 * S:(w0 + cnt - 1) * S + W].cpu().numpy()
    got = tab[w0:w0 + cnt].cpu().numpy()
    for j, nm in enumerate(names):
        want = OW.rolling(nm, seg, W, S, 90.0 if nm == "percentile" else (0.0 if nm == "zero_crossing_count" else None))
        ok = np.allclose(got[:, j], want, rtol=1e-6 if nm == "line_length" else 1e-9, atol=1e-12)
        if not ok:
            bad += 1
            print("MISMATCH at window %d: %s" % (w0, nm), np.abs(got[:, j] - want).max())
    sp = OS.spectral_table(seg, W, S, fs, [(0.5, 3.0)], 0.3, 12.0)
    tot = sp["total_power"]
    if not (np.all(np.abs(got[:, 8] - tot) <= 1e-5 * tot) and
            np.all(np.abs(got[:, 9] - sp["band_power_0"]) <= 1e-5 * np.maximum(sp["band_power_0"], 1e-3 * tot)) and
            np.allclose(got[:, 11], sp["spectral_entropy"], rtol=1e-5, atol=1e-9) and
            np.array_equal(got[:, 10].astype(np.int64), sp["peak_bin"])):
        bad += 1
        print("SPECTRAL MISMATCH at window %d" % w0)
print("big series: %d mismatches" % bad)
sys.exit(1 if bad else 0)
