#!/usr/bin/env python3
"""Does running kernel 1a and kernel 2 on two streams beat running them back to back?  (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymhealth_b200 import engine, synth
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

dev = torch.device("cuda:0")
nsub = int(os.environ.get("NSUB", "32"))
n = 30_240_000
x = synth.device_accelerometer(nsub, n, dev).view(nsub * 3, n)
stream_f, spec_f = bench.feature_list()
nw = engine.n_windows(n, 500, 250)
table = torch.empty((nsub * 3, nw, 16), dtype=torch.float32, device=dev)
t_stats, t_spec = table[:, :, :10], table[:, :, 10:]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def serial():
    engine.window_table(x, 500, 250, stream_f, out=t_stats)
    engine.window_table(x, 500, 250, spec_f, fs=50.0, out=t_spec)


def overlapped():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur)
    s2.wait_stream(cur)
    with torch.cuda.stream(s2):
        engine.window_table(x, 500, 250, spec_f, fs=50.0, out=t_spec)
    with torch.cuda.stream(s1):
        engine.window_table(x, 500, 250, stream_f, out=t_stats)
    cur.wait_stream(s1)
    cur.wait_stream(s2)


for name, fn in (("serial", serial), ("two streams", overlapped), ("serial", serial), ("two streams", overlapped)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%-12s %.3f ms  %.3f Gwin/s" % (name, ms, nsub * 3 * nw / ms / 1e6), flush=True)
