#!/usr/bin/env python3
"""Fused magnitude + kernel 1a against the two-step form (magnitude kernel, then kernel 1a).  (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymhealth_b200 import engine, synth, _lib as L
import bench

dev = torch.device("cuda:0")
nsub = int(os.environ.get("NSUB", "32"))
n = 30_240_000
x3 = synth.device_accelerometer(nsub, n, dev)          # [nsub, 3, n]
x, y, z = (x3[:, a, :].contiguous() for a in range(3))
del x3
stream_f, _ = bench.feature_list()
nw = engine.n_windows(n, 500, 250)
table = torch.empty((nsub, nw, len(stream_f)), dtype=torch.float32, device=dev)
mag = torch.empty_like(x)
lib = L.load()


def two_step():
    L.check(lib.mhb_accel_elementwise(0, 0, x.data_ptr(), y.data_ptr(), z.data_ptr(), nsub * n, mag.data_ptr(),
                                      engine._stream_ptr(torch)), "magnitude")
    engine.window_table(mag, 500, 250, stream_f, out=table)


def fused():
    engine.magnitude_window_table(x, y, z, 500, 250, stream_f, out=table)


def plain():
    engine.window_table(mag, 500, 250, stream_f, out=table)


def timeit(name, fn):
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
    print("%-58s %.3f ms  %.3f G magnitude-windows/s  %.2f TB/s of axis samples" %
          (name, ms, nsub * nw / ms / 1e6, 12.0 * nsub * n / ms / 1e9), flush=True)


two_step()
ref = table.clone()
timeit("kernel 1a alone (on a materialised magnitude)", plain)
timeit("two-step", two_step)
configs = [dict(MHB_MAG_COPY="1"), dict()]
for cfg in configs:
    for k in ("MHB_MAG_COPY",):
        os.environ.pop(k, None)
    os.environ.update(cfg)
    try:
        fused()
        torch.cuda.synchronize()
        same = bool(torch.equal(ref, table))
        timeit("fused %s same=%s" % (" ".join("%s=%s" % (k[8:], v) for k, v in cfg.items()) or "(defaults)", same), fused)
    except Exception as e:           # noqa: BLE001
        print("fused", cfg, "failed:", str(e)[:100])
