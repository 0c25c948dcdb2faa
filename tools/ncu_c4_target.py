#!/usr/bin/env python3
"""ncu target: config-4 kernels (W = 1920 / S = 64) -- python tools/ncu_c4_target.py [stats|spec]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymhealth_b200 import engine, synth, spectral as SP
from pymhealth_b200.generic import stats, timedom

what = sys.argv[1] if len(sys.argv) > 1 else "spec"
dev = torch.device("cuda:0")
x = synth.device_ppg(int(os.environ.get("NSUB", "32")), 5_529_600, dev)
fs = 64.0
if what == "stats":
    f = [stats.mean.feature(), stats.std.feature(), stats.var.feature(), stats.dmin.feature(), stats.dmax.feature(),
         stats.drange.feature(), stats.skewness.feature(), stats.kurtosis.feature(),
         timedom.zero_crossing_count.feature(0.0), timedom.line_length.feature()]
else:
    f = [SP.total_power(fs).feature(), SP.band_power(fs, 0.5, 4.0).feature(), SP.band_power(fs, 4.0, 8.0).feature(),
         SP.relative_band_power(fs, 0.5, 4.0).feature(), SP.peak_frequency(fs, 0.5, 4.0).feature(), SP.spectral_entropy(fs).feature()]
out = torch.empty((x.shape[0], engine.n_windows(x.shape[1], 1920, 64), len(f)), dtype=torch.float32, device=dev)
for _ in range(3):
    engine.window_table(x, 1920, 64, f, fs=fs, out=out)
torch.cuda.synchronize()
print("ok", float(out[0, 0, 0]))
