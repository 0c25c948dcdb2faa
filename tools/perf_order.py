#!/usr/bin/env python3
"""Device-resident timing of kernel 1b (order statistics / Hjorth) at config 3 / 4 window geometries."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymhealth_b200 import engine, synth
from pymhealth_b200.generic import stats, timedom
from tools.perf_stats import timeit

dev = torch.device("cuda:0")
for (label, x, W, S) in (("C3", synth.device_accelerometer(2, 30_240_000, dev).view(6, -1), 500, 250),
                         ("C4", synth.device_ppg(4, 5_529_600, dev), 1920, 64)):
    for name, feats in (("median", [stats.median.feature()]),
                        ("median+p10+p90+iqr", [stats.median.feature(), stats.percentile.feature(10.0), stats.percentile.feature(90.0),
                                                stats.interquartile_range.feature()]),
                        ("hjorth mobility+complexity", [timedom.hjorth_mobility.feature(), timedom.hjorth_complexity.feature()])):
        ns, n = x.shape
        nw = engine.n_windows(n, W, S)
        out = torch.empty((ns, nw, len(feats)), dtype=torch.float32, device=dev)
        best, med = timeit(lambda: engine.window_table(x, W, S, feats, out=out), iters=3, warm=1)
        print("%s %-28s ns=%d nw=%d best %.3f ms  %.4f Gwin/s  %.1f GB/s of window bytes" % (
            label, name, ns, nw, best, ns * nw / best / 1e6, ns * nw * W * 4 / best / 1e6), flush=True)
