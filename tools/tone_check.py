"""Entropy of noiseless tones (one bin holds nearly all the power) on the three spectral kernels vs the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import spectral as OS           # noqa: E402
from pymhealth_b200 import spectral as SP   # noqa: E402
from pymhealth_b200 import engine   # noqa: E402
import torch   # noqa: E402

rng = np.random.default_rng(5)
for W, S, fs in [(500, 250, 50.0), (1920, 64, 64.0), (256, 128, 50.0), (75, 25, 50.0)]:
    worst = {}
    for trial in range(40):
        n = W + S * int(rng.integers(3, 40)) + int(rng.integers(0, S))
        f0 = rng.uniform(0.3, fs / 2 - 0.5)
        if trial % 4 == 0:
            f0 = round(f0 * W / fs) * fs / W            # exactly on a bin
        off = float(rng.choice([0.0, 0.0, 1.0, -0.02]))
        x = (off + 0.4 * np.sin(2 * np.pi * f0 * np.arange(n) / fs)).astype(np.float32)
        for dt in (np.float64, np.float32):
            tab = engine.window_table(torch.from_numpy(x[None]).cuda(), W, S,
                                      [SP.spectral_entropy(fs).feature(), SP.total_power(fs).feature()], fs=fs,
                                      out_dtype=torch.float64 if dt is np.float64 else torch.float32)
            got = tab[0, :, 0].cpu().numpy().astype(np.float64)
            want = OS.spectral_table(x, W, S, fs, [], None, None)["spectral_entropy"]
            rel = np.abs(got - want) / want
            i = int(np.argmax(rel))
            key = (dt.__name__, "on-bin" if trial % 4 == 0 else "off-bin", off)
            if rel[i] > worst.get(key, (0, 0, 0))[0]:
                worst[key] = (float(rel[i]), float(want[i]), float(got[i]))
    for k in sorted(worst):
        print(W, S, k, "max rel err %.3g at H=%.6g (got %.6g)" % worst[k])
