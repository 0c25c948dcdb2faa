#!/usr/bin/env python3
"""Cross-check CUDA-event, wall-clock and (when run under ncu) gpu__time_duration timings of kernel 1a at bench size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymhealth_b200 import engine, synth
from pymhealth_b200.generic import stats, timedom

dev = torch.device("cuda:0")
nsub = int(os.environ.get("NSUB", "125"))
x = synth.device_accelerometer(nsub, 30_240_000, dev).view(nsub * 3, -1)
full = [stats.mean.feature(), stats.std.feature(), stats.var.feature(), stats.dmin.feature(), stats.dmax.feature(),
        stats.drange.feature(), stats.skewness.feature(), stats.kurtosis.feature(),
        timedom.zero_crossing_count.feature(0.0), timedom.line_length.feature()]
out = torch.empty((x.shape[0], engine.n_windows(x.shape[1], 500, 250), len(full)), dtype=torch.float32, device=dev)
for _ in range(3):
    engine.window_table(x, 500, 250, full, out=out)
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    engine.window_table(x, 500, 250, full, out=out)
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("single launch: events %.3f ms  wall %.3f ms" % (e0.elapsed_time(e1), (t1 - t0) * 1e3), flush=True)
    time.sleep(0.5)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(10):
    engine.window_table(x, 500, 250, full, out=out)
e1.record()
torch.cuda.synchronize()
t1 = time.perf_counter()
print("10 back-to-back: events %.3f ms/launch  wall %.3f ms/launch" % (e0.elapsed_time(e1) / 10, (t1 - t0) * 100), flush=True)
