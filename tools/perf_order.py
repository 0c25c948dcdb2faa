#!/usr/bin/env python3
"""Timing of kernel 1b (median + p90) at the config-3 and config-4 geometries.  python tools/perf_order.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymhealth_b200 import engine, synth
from pymhealth_b200.generic import stats

dev = torch.device("cuda:0")


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


nsub = int(os.environ.get("NSUB", "32"))
x = synth.device_accelerometer(nsub, 30_240_000, dev).view(nsub * 3, -1)
for name, f in (("median+p90", [stats.median.feature(), stats.percentile.feature(90.0)]),
                ("median", [stats.median.feature()]),
                ("median+p10+p90+iqr", [stats.median.feature(), stats.percentile.feature(10.0), stats.percentile.feature(90.0),
                                        stats.interquartile_range.feature()])):
    nw = engine.n_windows(x.shape[1], 500, 250)
    out = torch.empty((x.shape[0], nw, len(f)), dtype=torch.float32, device=dev)
    ms = timeit(lambda: engine.window_table(x, 500, 250, f, out=out))
    print("c3 %-20s %8.3f ms  %.3f G windows/s" % (name, ms, x.shape[0] * nw / ms / 1e6))
del x
x = synth.device_ppg(32, 5_529_600, dev)
for name, f in (("median", [stats.median.feature()]),
                ("median+p10+p90+iqr", [stats.median.feature(), stats.percentile.feature(10.0), stats.percentile.feature(90.0),
                                        stats.interquartile_range.feature()])):
    nw = engine.n_windows(x.shape[1], 1920, 64)
    out = torch.empty((x.shape[0], nw, len(f)), dtype=torch.float32, device=dev)
    ms = timeit(lambda: engine.window_table(x, 1920, 64, f, out=out))
    print("c4 %-20s %8.3f ms  %.3f G windows/s" % (name, ms, x.shape[0] * nw / ms / 1e6))
