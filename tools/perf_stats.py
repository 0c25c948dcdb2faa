#!/usr/bin/env python3
"""Quick device-resident timing of the window kernels (development aid, not the bench contract)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymhealth_b200 import engine, synth
from pymhealth_b200.generic import stats, timedom

PEAK = 6537.3


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[0], ts[len(ts) // 2]


def run(label, x, W, S, feats, out_dtype=torch.float32):
    ns, n = x.shape
    nw = engine.n_windows(n, W, S)
    out = torch.empty((ns, nw, len(feats)), dtype=out_dtype, device=x.device)
    best, med = timeit(lambda: engine.window_table(x, W, S, feats, out=out))
    bytes_alg = x.numel() * x.element_size() + out.numel() * out.element_size()
    print("%-34s ns=%4d n=%9d nw=%7d F=%2d  best %8.3f ms  med %8.3f ms  %8.1f GB/s (%.3f of %g)  %.3f Gwin/s"
          % (label, ns, n, nw, len(feats), best, med, bytes_alg / best / 1e6, bytes_alg / best / 1e6 / PEAK, PEAK,
             ns * nw / best / 1e6), flush=True)


def main():
    dev = torch.device("cuda:0")
    lvl0 = [stats.mean.feature(), stats.var.feature(), stats.dmin.feature(), stats.dmax.feature()]
    td = lvl0 + [timedom.zero_crossing_count.feature(0.0), timedom.line_length.feature()]
    m4 = lvl0 + [stats.skewness.feature(), stats.kurtosis.feature()]
    full = [stats.mean.feature(), stats.std.feature(), stats.var.feature(), stats.dmin.feature(), stats.dmax.feature(),
            stats.drange.feature(), stats.skewness.feature(), stats.kurtosis.feature(),
            timedom.zero_crossing_count.feature(0.0), timedom.line_length.feature()]
    nsub = int(os.environ.get("NSUB", "8"))
    x = synth.device_accelerometer(nsub, 30_240_000, dev).view(nsub * 3, -1)
    for label, f in (("C3 acc lvl0(4)", lvl0), ("C3 acc +zc,ll(6)", td), ("C3 acc +skew,kurt(6)", m4), ("C3 acc full(10)", full)):
        run(label, x, 500, 250, f)
    del x
    p = synth.device_ppg(64, 5_529_600, dev)
    for label, f in (("C4 ppg lvl0(4)", lvl0), ("C4 ppg full(10)", full)):
        run(label, p, 1920, 64, f)
    x2 = synth.device_accelerometer(1, 4_320_000, dev).view(3, -1)
    run("C2 acc full(10)", x2, 500, 250, full)


if __name__ == "__main__":
    main()
