#!/usr/bin/env python3
"""Small single-purpose launcher for ncu captures: python tools/ncu_target.py <what> [iters]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymhealth_b200 import engine, synth
from pymhealth_b200.generic import stats, timedom


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "c3_full"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda:0")
    lvl0 = [stats.mean.feature(), stats.var.feature(), stats.dmin.feature(), stats.dmax.feature()]
    full = [stats.mean.feature(), stats.std.feature(), stats.var.feature(), stats.dmin.feature(), stats.dmax.feature(),
            stats.drange.feature(), stats.skewness.feature(), stats.kurtosis.feature(),
            timedom.zero_crossing_count.feature(0.0), timedom.line_length.feature()]
    if what.startswith("c3"):
        nsub = int(os.environ.get("NSUB", "2"))
        x = synth.device_accelerometer(nsub, 30_240_000, dev).view(nsub * 3, -1)
        W, S = 500, 250
    elif what.startswith("c4"):
        x = synth.device_ppg(32, 5_529_600, dev)
        W, S = 1920, 64
    if what == "c3_mag":          # kernel 1a in magnitude mode: the three axis planes of the same tensor
        nsub = x.shape[0] // 3
        x3 = x.view(nsub, 3, -1)
        out = torch.empty((nsub, engine.n_windows(x.shape[1], W, S), len(full)), dtype=torch.float32, device=dev)
        for _ in range(iters):
            engine.magnitude_window_table(x3[:, 0], x3[:, 1], x3[:, 2], W, S, full, out=out)
        torch.cuda.synchronize()
        print("ok", what, float(out[0, 0, 0]))
        return
    feats = lvl0 if what.endswith("lvl0") else full
    if what.endswith("spec"):
        from pymhealth_b200 import spectral as SP
        fs = 50.0 if what.startswith("c3") else 64.0
        feats = [SP.total_power(fs).feature(), SP.band_power(fs, 0.5, 3.0).feature(), SP.band_power(fs, 3.0, 8.0).feature(),
                 SP.relative_band_power(fs, 0.5, 3.0).feature(), SP.peak_frequency(fs, 0.3, 12.0).feature(),
                 SP.spectral_entropy(fs).feature()]
    out = torch.empty((x.shape[0], engine.n_windows(x.shape[1], W, S), len(feats)), dtype=torch.float32, device=dev)
    for _ in range(iters):
        engine.window_table(x, W, S, feats, out=out, fs=50.0 if what.startswith("c3") else 64.0)
    torch.cuda.synchronize()
    print("ok", what, float(out[0, 0, 0]))


if __name__ == "__main__":
    main()
