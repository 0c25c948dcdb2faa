"""Seeded synthetic sensor streams for the five BASELINE.json configurations.

numpy generators (host) used by the parity tests, the golden-fixture script and the host
side of ``bench.py``; ``device_accelerometer`` / ``device_ppg`` build the same kind of signal
directly in HBM (torch ops: plumbing only) for workloads too big to stage through the host.

There is no network in the build environment, so all data is synthetic by construction.
"""
import numpy as np

ACC_FS = 50.0       # Hz, configs 2/3
ACC_WSIZE, ACC_WSTEP = 500, 250      # 10 s windows, 50 % overlap
PPG_FS = 64.0       # Hz, config 4
PPG_WSIZE, PPG_WSTEP = 1920, 64      # 30 s windows, 1 s hop
DAY_S = 86400


def accelerometer(subject_id, n, fs=ACC_FS, dtype=np.float32):
    """Triaxial accelerometer, layout [axis][sample] (SoA): gravity vector rotated by a slow
    random walk + 2-3 activity tones + 0.05 g white noise.  seed = 1000 + subject_id."""
    rng = np.random.default_rng(1000 + int(subject_id))
    t = np.arange(n, dtype=np.float64) / fs
    # slow orientation walk: piecewise-linear pitch / roll with knots every ~10 min
    nk = max(2, int(n / (fs * 600)) + 2)
    knots = np.linspace(0, n - 1, nk)
    pitch = np.interp(np.arange(n), knots, np.cumsum(rng.normal(0, 0.15, nk)))
    roll = np.interp(np.arange(n), knots, np.cumsum(rng.normal(0, 0.15, nk)))
    g = np.stack([np.sin(pitch), -np.cos(pitch) * np.sin(roll), np.cos(pitch) * np.cos(roll)])
    out = g + 0.05 * rng.standard_normal((3, n))
    for _ in range(int(rng.integers(2, 4))):
        f = rng.uniform(0.5, 5.0)
        for ax in range(3):
            out[ax] += rng.uniform(0.05, 0.5) * np.sin(2 * np.pi * f * t + rng.uniform(0, 2 * np.pi))
    return np.ascontiguousarray(out, dtype=dtype)


def ppg(subject_id, n, fs=PPG_FS, dtype=np.float32):
    """PPG-like 1-D stream: heart-rate random walk in [50, 120] bpm, fundamental + 2nd harmonic,
    0.2 Hz baseline wander, 0.02 noise.  seed = 2000 + subject_id."""
    rng = np.random.default_rng(2000 + int(subject_id))
    nk = max(2, int(n / (fs * 30)) + 2)
    hr_k = np.clip(75 + np.cumsum(rng.normal(0, 2.0, nk)), 50, 120)
    hr = np.interp(np.arange(n), np.linspace(0, n - 1, nk), hr_k)
    phase = 2 * np.pi * np.cumsum(hr / 60.0) / fs
    t = np.arange(n, dtype=np.float64) / fs
    x = np.sin(phase) + 0.35 * np.sin(2 * phase + 0.7) + 0.3 * np.sin(2 * np.pi * 0.2 * t)
    x += 0.02 * rng.standard_normal(n)
    return np.ascontiguousarray(x, dtype=dtype)


def gps(subject_id, n, period_s=60, t0=1_600_000_000):
    """GPS trace: dwell at 3-6 places within ~20 km of a home point / straight-line travel at
    5-50 km/h, + N(0, 10 m) jitter.  Returns (lat f64[n], lon f64[n], t int64[n], home (lat, lon)).
    seed = 3000 + subject_id."""
    rng = np.random.default_rng(3000 + int(subject_id))
    home = (rng.uniform(-60, 60), rng.uniform(-180, 180))
    kml = 111.19                                      # km per degree latitude
    kmo = kml * np.cos(np.radians(home[0]))
    npl = int(rng.integers(3, 7))
    places = [(0.0, 0.0)] + [tuple(rng.uniform(-20, 20, 2)) for _ in range(npl - 1)]   # km offsets
    xs = np.empty(n)
    ys = np.empty(n)
    i = 0
    cur = 0
    while i < n:
        dwell = int(min(n - i, max(1, rng.lognormal(np.log(7200.0), 0.8) / period_s)))
        xs[i:i + dwell] = places[cur][0]
        ys[i:i + dwell] = places[cur][1]
        i += dwell
        if i >= n:
            break
        nxt = int(rng.integers(0, npl))
        if nxt == cur:
            nxt = (cur + 1) % npl
        dist = np.hypot(places[nxt][0] - places[cur][0], places[nxt][1] - places[cur][1])
        steps = int(min(n - i, max(1, dist / rng.uniform(5, 50) * 3600 / period_s)))
        f = (np.arange(steps) + 1) / (steps + 1)
        xs[i:i + steps] = places[cur][0] + f * (places[nxt][0] - places[cur][0])
        ys[i:i + steps] = places[cur][1] + f * (places[nxt][1] - places[cur][1])
        i += steps
        cur = nxt
    xs += rng.normal(0, 0.010, n)
    ys += rng.normal(0, 0.010, n)
    lat = home[0] + ys / kml
    lon = home[1] + xs / kmo
    t = t0 + period_s * np.arange(n, dtype=np.int64)
    return np.ascontiguousarray(lat), np.ascontiguousarray(lon), t, home


# ------------------------------------------------------------------ device-side generators
def device_accelerometer(n_subjects, n, device, first_subject=0, fs=ACC_FS, chunk=8):
    """float32 [n_subjects, 3, n] built in HBM with torch ops (gravity offset + tones + noise).
    Same signal family as ``accelerometer`` but NOT the same values (different RNG); parity
    tests use the numpy generator, the big bench workloads use this one."""
    import torch
    out = torch.empty((n_subjects, 3, n), dtype=torch.float32, device=device)
    t = torch.arange(n, device=device, dtype=torch.float64) / fs     # fp64: 7 days of phase
    for s0 in range(0, n_subjects, chunk):
        s1 = min(n_subjects, s0 + chunk)
        gen = torch.Generator(device=device)
        gen.manual_seed(1000 + first_subject + s0)
        blk = out[s0:s1]
        blk.normal_(0.0, 0.05, generator=gen)
        g = torch.tensor([0.1, -0.2, 0.97], device=device).view(1, 3, 1)
        blk += g
        for _ in range(3):
            f = torch.empty((s1 - s0, 1, 1), device=device).uniform_(0.5, 5.0, generator=gen)
            a = torch.empty((s1 - s0, 3, 1), device=device).uniform_(0.05, 0.5, generator=gen)
            ph = torch.empty((s1 - s0, 3, 1), device=device).uniform_(0, 6.2831853, generator=gen)
            for k in range(s1 - s0):            # one subject at a time: bounded temporaries
                cyc = torch.frac(f[k].double().view(1) * t).float().view(1, n)
                blk[k] += a[k] * torch.sin(6.2831853 * cyc + ph[k])
    return out


def device_ppg(n_subjects, n, device, first_subject=0, fs=PPG_FS):
    """float32 [n_subjects, n] PPG-like stream built in HBM."""
    import torch
    out = torch.empty((n_subjects, n), dtype=torch.float32, device=device)
    t = torch.arange(n, device=device, dtype=torch.float64) / fs
    wander = torch.sin(6.2831853 * torch.frac(0.2 * t).float())
    gen = torch.Generator(device=device)
    gen.manual_seed(2000 + first_subject)
    for k in range(n_subjects):
        hr = 50 + 70 * torch.rand((), device=device, generator=gen)
        ph = 6.2831853 * torch.frac((hr.double() / 60.0) * t).float()
        row = out[k]
        row.normal_(0.0, 0.02, generator=gen)
        row += torch.sin(ph) + 0.35 * torch.sin(2 * ph + 0.7) + 0.3 * wander
    return out
