#!/usr/bin/env python3
"""Device-resident timing of the config-3 step (development aid): fused call vs statistics-only vs spectral-only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymhealth_b200 import engine, synth, spectral as SP
from pymhealth_b200.generic import stats, timedom
from tools.perf_stats import timeit

def main():
    dev = torch.device("cuda:0")
    nsub = int(os.environ.get("NSUB", "32"))
    fs = 50.0
    x = synth.device_accelerometer(nsub, 30_240_000, dev).view(nsub * 3, -1)
    sf = [stats.mean.feature(), stats.std.feature(), stats.var.feature(), stats.dmin.feature(), stats.dmax.feature(),
          stats.drange.feature(), stats.skewness.feature(), stats.kurtosis.feature(),
          timedom.zero_crossing_count.feature(0.0), timedom.line_length.feature()]
    pf = [SP.total_power(fs).feature(), SP.band_power(fs, 0.5, 3.0).feature(), SP.band_power(fs, 3.0, 8.0).feature(),
          SP.relative_band_power(fs, 0.5, 3.0).feature(), SP.peak_frequency(fs, 0.3, 12.0).feature(),
          SP.spectral_entropy(fs).feature()]
    ns, n = x.shape
    nw = engine.n_windows(n, 500, 250)
    out = torch.empty((ns, nw, 16), dtype=torch.float32, device=dev)
    for label, feats, o in (("fused(16)", sf + pf, out), ("stats(10)", sf, out[:, :, :10]), ("spectral(6)", pf, out[:, :, 10:])):
        best, med = timeit(lambda: engine.window_table(x, 500, 250, feats, fs=fs, out=o), iters=5, warm=2)
        alg = x.numel() * 4 + ns * nw * len(feats) * 4
        print("%-12s ns=%d nw=%d best %.3f ms med %.3f ms  %.3f Gwin/s  %.1f GB/s (%.3f of 6537)" % (
            label, ns, nw, best, med, ns * nw / best / 1e6, alg / best / 1e6, alg / best / 1e6 / 6537.3), flush=True)

if __name__ == "__main__":
    main()
