#!/usr/bin/env python3
"""Device-resident timing of the spectral kernel (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymhealth_b200 import engine, synth, spectral as SP
from tools.perf_stats import timeit

def main():
    dev = torch.device("cuda:0")
    nsub = int(os.environ.get("NSUB", "8"))
    for (label, x, W, S, fs) in (("C3", synth.device_accelerometer(nsub, 30_240_000, dev).view(nsub * 3, -1), 500, 250, 50.0),
                                 ("C4", synth.device_ppg(16, 5_529_600, dev), 1920, 64, 64.0)):
        feats = [SP.total_power(fs).feature(), SP.band_power(fs, 0.5, 3.0).feature(), SP.band_power(fs, 3.0, 8.0).feature(),
                 SP.relative_band_power(fs, 0.5, 3.0).feature(), SP.peak_frequency(fs, 0.3, 12.0).feature(),
                 SP.spectral_entropy(fs).feature()]
        ns, n = x.shape
        nw = engine.n_windows(n, W, S)
        out = torch.empty((ns, nw, len(feats)), dtype=torch.float32, device=dev)
        best, med = timeit(lambda: engine.window_table(x, W, S, feats, fs=fs, out=out), iters=5, warm=2)
        print("%s spectral(6): ns=%d nw=%d best %.3f ms med %.3f ms  %.3f Gwin/s  %.1f GB/s" % (
            label, ns, nw, best, med, ns * nw / best / 1e6, (x.numel() * 4 + out.numel() * 4) / best / 1e6), flush=True)
        del x, out

if __name__ == "__main__":
    main()
