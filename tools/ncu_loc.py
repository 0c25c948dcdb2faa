#!/usr/bin/env python3
"""ncu target: the per-segment location kernel on 1024 subject-days of 1 Hz GPS."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymhealth_b200 import synth
from pymhealth_b200.location import features
dev = torch.device("cuda:0")
nseg, day = 1024, 86400
lat0, lon0, t0, home0 = synth.gps(1, day, 1)
lat = torch.from_numpy(lat0).to(dev).repeat(nseg)
lon = torch.from_numpy(lon0).to(dev).repeat(nseg)
t = torch.from_numpy(t0).to(dev).repeat(nseg)
offs = torch.arange(nseg + 1, device=dev, dtype=torch.int64) * day
home = torch.tensor([home0] * nseg, dtype=torch.float64, device=dev)
for _ in range(2):
    rows = features.segment_rows(lat, lon, t, offs, home)
torch.cuda.synchronize()
print("ok", float(rows[0, 1]))
