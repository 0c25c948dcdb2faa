#!/usr/bin/env python3
"""Device-resident timing of the location kernels at config-5 shape (1 Hz GPS, one segment per subject-day)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pymhealth_b200 import synth
from pymhealth_b200.location import features, distance
from tools.perf_stats import timeit


def main():
    dev = torch.device("cuda:0")
    nseg = int(os.environ.get("NSEG", "512"))            # subject-days
    day = 86400
    lat0, lon0, t0, home0 = synth.gps(1, day, 1)
    lat = torch.from_numpy(lat0).to(dev).repeat(nseg) + torch.arange(nseg, device=dev).repeat_interleave(day) * 1e-4
    lon = torch.from_numpy(lon0).to(dev).repeat(nseg)
    t = torch.from_numpy(t0).to(dev).repeat(nseg)
    offs = torch.arange(nseg + 1, device=dev, dtype=torch.int64) * day
    home = torch.tensor([home0] * nseg, dtype=torch.float64, device=dev)
    n = nseg * day
    best, med = timeit(lambda: features.segment_rows(lat, lon, t, offs, home), iters=5, warm=2)
    print("C5 per-day rows: %d segments x %d pts  best %.3f ms  %.2f Gpts/s  %.1f GB/s (24 B/pt)  %.0f segments/s" % (
        nseg, day, best, n / best / 1e6, n * 24 / best / 1e6, nseg / best * 1e3), flush=True)
    best, med = timeit(lambda: features.arr_successive_distance(lat, lon), iters=5, warm=2)
    print("successive distance: %d pts  best %.3f ms  %.2f Gpts/s  %.1f GB/s (24 B/pt)" % (n, best, n / best / 1e6, n * 24 / best / 1e6), flush=True)
    best, med = timeit(lambda: distance.haversine_vector(home0[0], home0[1], lat, lon), iters=5, warm=2)
    print("distance from home: %d pts  best %.3f ms  %.2f Gpts/s  %.1f GB/s (24 B/pt)" % (n, best, n / best / 1e6, n * 24 / best / 1e6), flush=True)


if __name__ == "__main__":
    main()
