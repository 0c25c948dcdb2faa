import os, sys
sys.path.insert(0, '/root/repo')
import torch
from pymhealth_b200 import engine, synth, spectral as SP
from tools.perf_stats import timeit
dev = torch.device("cuda:0")
nsub = 32
x = synth.device_accelerometer(nsub, 30_240_000, dev).view(nsub * 3, -1)
fs = 50.0
sets = {
 "total only": [SP.total_power(fs).feature()],
 "total+entropy": [SP.total_power(fs).feature(), SP.spectral_entropy(fs).feature()],
 "total+band": [SP.total_power(fs).feature(), SP.band_power(fs, 0.5, 3.0).feature()],
 "total+peak": [SP.total_power(fs).feature(), SP.peak_frequency(fs, 0.3, 12.0).feature()],
 "all six": [SP.total_power(fs).feature(), SP.band_power(fs, 0.5, 3.0).feature(), SP.band_power(fs, 3.0, 8.0).feature(),
             SP.relative_band_power(fs, 0.5, 3.0).feature(), SP.peak_frequency(fs, 0.3, 12.0).feature(), SP.spectral_entropy(fs).feature()],
 "twelve": [SP.total_power(fs).feature(), SP.band_power(fs, 0.5, 3.0).feature(), SP.band_power(fs, 3.0, 8.0).feature(),
             SP.relative_band_power(fs, 0.5, 3.0).feature(), SP.peak_frequency(fs, 0.3, 12.0).feature(), SP.spectral_entropy(fs).feature()] * 2,
}
for name, feats in sets.items():
    ns, n = x.shape
    nw = engine.n_windows(n, 500, 250)
    out = torch.empty((ns, nw, len(feats)), dtype=torch.float32, device=dev)
    best, med = timeit(lambda: engine.window_table(x, 500, 250, feats, fs=fs, out=out), iters=5, warm=2)
    print("%-16s %d cols: %.3f ms  %.3f Gwin/s" % (name, len(feats), best, ns * nw / best / 1e6), flush=True)
