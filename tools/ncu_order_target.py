#!/usr/bin/env python3
"""ncu target: kernel 1b (median + p90) at config-3 geometry.  python tools/ncu_order_target.py [c3|c4]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymhealth_b200 import engine, synth
from pymhealth_b200.generic import stats

what = sys.argv[1] if len(sys.argv) > 1 else "c3"
dev = torch.device("cuda:0")
if what == "c3":
    nsub = int(os.environ.get("NSUB", "8"))
    x = synth.device_accelerometer(nsub, 30_240_000, dev).view(nsub * 3, -1)
    W, S = 500, 250
else:
    x = synth.device_ppg(16, 5_529_600, dev)
    W, S = 1920, 64
f = [stats.median.feature(), stats.percentile.feature(90.0)]
out = torch.empty((x.shape[0], engine.n_windows(x.shape[1], W, S), 2), dtype=torch.float32, device=dev)
for _ in range(3):
    engine.window_table(x, W, S, f, out=out)
torch.cuda.synchronize()
print("ok", float(out[0, 0, 0]))
