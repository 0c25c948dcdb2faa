#!/usr/bin/env python3
"""Randomised differential test of the GPU paths against the oracle (development aid; the committed tests hold the
fixed cases).  python tools/fuzz_gpu.py [n_cases] [seed]"""
import functools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import windows as OW, spectral as OS
from pymhealth_b200 import spectral as SP
from pymhealth_b200.generic import stats, timedom
from pymhealth_b200.util import rolling_apply
from pymhealth_b200.util.windows import nonuniform_rolling_apply


def main():
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad = 0
    for case in range(ncases):
        kind = case % 4
        if kind == 0:          # order statistics: geometries that hit the block paths (k <= 2, k >= 8) and the full sort
            g = int(rng.choice([16, 24, 32, 50, 64, 100, 125, 250, 256, 300]))
            k = int(rng.choice([1, 2, 2, 2, 3, 5, 8, 12, 30]))
            hop = int(rng.integers(1, max(2, k + 1)))
            W, S = g * k, g * hop
            n = W + S * int(rng.integers(0, 40)) + int(rng.integers(0, S))
            x = np.round(rng.standard_normal(n) * 30).astype(np.float32) / 4 + rng.choice([0.0, 100.0])
            qs = [float(q) for q in rng.choice([0, 5, 12.5, 25, 50, 75, 90, 99.9, 100], size=2, replace=False)]
            funcs = [np.median, functools.partial(np.percentile, q=qs[0]), functools.partial(np.percentile, q=qs[1]),
                     stats.interquartile_range]
            got = rolling_apply(funcs)(x, W, S)
            names = ["median", "percentile", "percentile", "iqr"]
            pars = [None, qs[0], qs[1], None]
            for g_, nm, p in zip(got, names, pars):
                want = OW.rolling(nm, x, W, S, p)
                if not np.allclose(g_, want, rtol=1e-12, atol=0):
                    bad += 1
                    print("ORDER MISMATCH W=%d S=%d n=%d %s %s max|d|=%g" % (W, S, n, nm, p, np.abs(g_ - want).max()))
        elif kind == 1:        # streaming statistics, random geometry
            W = int(rng.integers(2, 700))
            S = int(rng.integers(1, 2 * W))
            n = W + S * int(rng.integers(0, 60)) + int(rng.integers(0, S))
            x = (rng.standard_normal(n) * rng.choice([0.01, 1.0, 50.0]) + rng.choice([0.0, 1.0, -300.0])).astype(np.float32)
            th = float(rng.choice([0.0, 0.5]))
            funcs = [np.mean, np.var, np.min, np.max, stats.skewness, stats.kurtosis, timedom.line_length,
                     functools.partial(timedom.zero_crossing_count, th=th)]
            names = ["mean", "var", "min", "max", "skewness", "kurtosis", "line_length", "zero_crossing_count"]
            got = rolling_apply(funcs)(x, W, S)
            for g_, nm in zip(got, names):
                want = OW.rolling(nm, x, W, S, th)
                scale = np.maximum(np.abs(want), 1e-3 * (np.abs(x).max() ** (2 if nm == "var" else 1)) + 1e-12)
                tol = 0 if nm in ("min", "max", "zero_crossing_count") else (1e-6 if nm == "line_length" else 1e-7)
                if nm in ("skewness", "kurtosis"):
                    scale = np.maximum(np.abs(want), 1.0)
                if np.any(np.abs(g_ - want) > tol * scale):
                    bad += 1
                    print("STATS MISMATCH W=%d S=%d n=%d %s worst=%g" % (W, S, n, nm, np.max(np.abs(g_ - want) / scale)))
        elif kind == 2:        # spectral: the two fast paths with ragged lengths, and random even / odd windows
            W, S = [(500, 250), (1920, 64), (int(rng.integers(8, 400)) * 2, int(rng.integers(1, 200))),
                    (int(rng.integers(8, 300)) * 2 + 1, int(rng.integers(1, 200)))][int(rng.integers(0, 4))]
            n = W + S * int(rng.integers(0, 50)) + int(rng.integers(0, S))
            fs = float(rng.choice([50.0, 64.0, 100.0]))
            t = np.arange(n) / fs
            x = (rng.choice([0.0, 1.0]) + 0.3 * np.sin(2 * np.pi * rng.uniform(0.5, 5) * t) + 0.05 * rng.standard_normal(n)).astype(np.float32)
            lo, hi = 0.5, float(rng.uniform(2.0, 12.0))
            try:
                got = rolling_apply([SP.total_power(fs), SP.band_power(fs, lo, hi), SP.spectral_entropy(fs)])(x, W, S)
            except NotImplementedError:
                continue          # a prime factor > 31
            tab = OS.spectral_table(x, W, S, fs, [(lo, hi)], 0.3, 12.0)
            tot = tab["total_power"]
            ok = (np.all(np.abs(got[0] - tot) <= 1e-5 * tot) and
                  np.all(np.abs(got[1] - tab["band_power_0"]) <= 1e-5 * np.maximum(tab["band_power_0"], 1e-3 * tot)) and
                  np.allclose(got[2], tab["spectral_entropy"], rtol=1e-5))
            if not ok:
                bad += 1
                print("SPECTRAL MISMATCH W=%d S=%d n=%d fs=%g" % (W, S, n, fs))
        else:                  # non-uniform windows
            n = int(rng.integers(50, 5000))
            idx = np.cumsum(rng.integers(1, 5, n)).astype(np.int64)
            x = rng.standard_normal(n)
            wsize, wstep = int(rng.integers(5, 400)), int(rng.integers(1, 200))
            minlen = int(rng.integers(1, 6))
            got = nonuniform_rolling_apply([np.mean, np.std, np.max, np.median], minlen)(idx, x, wsize, wstep)
            for g_, nm in zip(got, ["mean", "std", "max", "median"]):
                want = OW.nonuniform_rolling(nm, idx, x, wsize, wstep, minlen)
                if not np.allclose(g_, want, rtol=1e-9, atol=1e-12, equal_nan=True):
                    bad += 1
                    print("NONUNIFORM MISMATCH n=%d wsize=%d wstep=%d %s" % (n, wsize, wstep, nm))
    print("fuzz: %d cases, %d mismatches" % (ncases, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
