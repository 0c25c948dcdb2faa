#!/usr/bin/env python3
"""Config 5 at one node: subjects x 30 days of 1 Hz GPS sharded over the ranks, per-day feature rows on every rank,
one NCCL gather of the [subject-days, 11] tables.  Launch with torchrun (or plain python for one GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/perf_config5.py

Synthetic traces: a few host-generated subject-months (pymhealth_b200.synth.gps) tiled with per-subject offsets on the
device; SUBJECTS_PER_GPU (default 16) subjects per rank."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from pymhealth_b200 import sharded, synth
from pymhealth_b200.location import features


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    nsub = int(os.environ.get("SUBJECTS_PER_GPU", "16"))
    day, ndays = 86400, 30
    n_month = day * ndays
    base = [synth.gps(1000 * rank + k, n_month, 1) for k in range(2)]          # two generated subject-months per rank
    lat = torch.empty(nsub * n_month, dtype=torch.float64, device=dev)
    lon = torch.empty_like(lat)
    t = torch.empty(nsub * n_month, dtype=torch.int64, device=dev)
    home = torch.empty((nsub * ndays, 2), dtype=torch.float64, device=dev)
    for s in range(nsub):
        la, lo, tt, hm = base[s % 2]
        sl = slice(s * n_month, (s + 1) * n_month)
        lat[sl] = torch.from_numpy(la).to(dev) + 1e-3 * (s // 2)
        lon[sl] = torch.from_numpy(lo).to(dev)
        t[sl] = torch.from_numpy(tt).to(dev)
        home[s * ndays:(s + 1) * ndays, 0] = hm[0] + 1e-3 * (s // 2)
        home[s * ndays:(s + 1) * ndays, 1] = hm[1]
    offs = torch.arange(nsub * ndays + 1, device=dev, dtype=torch.int64) * day
    for _ in range(2):
        rows = features.segment_rows(lat, lon, t, offs, home)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    rows = features.segment_rows(lat, lon, t, offs, home)
    e1.record()
    full = sharded.gather_tables(rows, nsub * ndays * world) if world > 1 else rows
    e2.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        pts = world * nsub * n_month
        print(json.dumps({"config": "BASELINE configs[4] shape: %d subjects x 30 d x 86400 points, %d per GPU" % (world * nsub, nsub),
                          "n_gpus": world, "points": pts, "subject_days": world * nsub * ndays,
                          "rows_ms": float(ms[0]), "gather_ms": float(ms[1]),
                          "points_per_s": pts / (float(ms[0]) * 1e-3), "gb_per_s_per_gpu": nsub * n_month * 24 / (float(ms[0]) * 1e-3) / 1e9,
                          "table_shape": list(full.shape), "row0": [float(v) for v in full[0][:4]]}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
