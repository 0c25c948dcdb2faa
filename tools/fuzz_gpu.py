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
        kind = case % 18
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
        elif kind == 4:        # location: random segment boundaries (empty ones included), random thresholds
            from oracle import location_ext as OX
            from pymhealth_b200 import synth, _lib as L
            from pymhealth_b200.location import features
            n = int(rng.integers(200, 6000))
            lat, lon, t, home = synth.gps(int(rng.integers(0, 10**6)), n, int(rng.choice([1, 10, 60])))
            cuts = np.sort(rng.integers(0, n + 1, int(rng.integers(0, 8))))
            offs = np.concatenate([[0], cuts, [n]]).astype(np.int64)
            homes = np.tile(np.array(home), (len(offs) - 1, 1)) + rng.normal(0, 1e-3, (len(offs) - 1, 2))
            limit, sd, mins = float(rng.choice([0.05, 0.1, 1.0])), float(rng.choice([0.05, 0.2, 0.5])), int(rng.choice([60, 600, 1800]))
            rows, labels = features.segment_rows(lat, lon, t, offs, homes, limit, sd, mins, labels=True)
            want, wlab = OX.segment_features(lat, lon, t, offs, homes, limit, sd, mins)
            ok = np.array_equal(labels, wlab)
            for c, name in enumerate(L.SEG_COLUMNS):
                if name in ("n_points", "home_stay_count", "n_stay_points", "n_labels"):
                    ok = ok and np.array_equal(rows[:, c], want[:, c])
                else:
                    ok = ok and np.allclose(rows[:, c], want[:, c], rtol=1e-9, atol=1e-12, equal_nan=True)
            if not ok:
                bad += 1
                print("LOCATION MISMATCH n=%d segments=%d limit=%g sd=%g mins=%d" % (n, len(offs) - 1, limit, sd, mins))
        elif kind == 5:        # float64 multi-series device tensors through the engine
            import torch
            from pymhealth_b200 import engine
            ns, W, S = int(rng.integers(1, 5)), int(rng.integers(2, 300)), int(rng.integers(1, 300))
            n = W + S * int(rng.integers(0, 40)) + int(rng.integers(0, S))
            x = rng.standard_normal((ns, n)) * 3 + 10
            feats = [stats.mean.feature(), stats.std.feature(), stats.dmax.feature(), stats.median.feature(),
                     timedom.hjorth_mobility.feature(), stats.mode.feature()]
            x[:, ::3] = np.round(x[:, ::3])          # ties for the mode
            if W < 2:
                continue
            tab = engine.window_table(torch.from_numpy(x).cuda(), W, S, feats, out_dtype=torch.float64).cpu().numpy()
            for si in range(ns):
                for c, nm in enumerate(["mean", "std", "max", "median", "hjorth_mobility", "mode"]):
                    want = OW.rolling(nm, x[si], W, S)
                    if not np.allclose(tab[si, :, c], want, rtol=1e-9, atol=1e-12, equal_nan=True):
                        bad += 1
                        print("F64 MISMATCH ns=%d W=%d S=%d n=%d %s" % (ns, W, S, n, nm))
        elif kind == 6:        # spectral fast path, many columns with random (possibly empty / full) ranges
            fs = 50.0
            n = 500 + 250 * int(rng.integers(0, 70)) + int(rng.integers(0, 250))
            tt = np.arange(n) / fs
            x = (rng.choice([0.0, 1.0, -2.0]) + 0.4 * np.sin(2 * np.pi * rng.uniform(0.3, 20) * tt) + rng.choice([0.0, 0.05]) *
                 rng.standard_normal(n)).astype(np.float32)
            bands = [(float(a), float(a + b)) for a, b in zip(rng.uniform(0, 20, 5), rng.uniform(0.05, 10, 5))]
            plo, phi = float(rng.uniform(0, 5)), float(rng.uniform(6, 25))
            funcs = [SP.total_power(fs)] + [SP.band_power(fs, lo_, hi_) for lo_, hi_ in bands] + \
                    [SP.relative_band_power(fs, *bands[0]), SP.peak_bin(fs, plo, phi), SP.peak_bin(fs), SP.spectral_entropy(fs)]
            got = rolling_apply(funcs)(x, 500, 250)
            tab = OS.spectral_table(x, 500, 250, fs, bands, plo, phi)
            tot = tab["total_power"]
            why = []
            ok = np.all(np.abs(got[0] - tot) <= 1e-5 * np.maximum(tot, 1e-30))
            if not ok:
                why.append("total")
            for j in range(5):
                okj = np.all(np.abs(got[1 + j] - tab["band_power_%d" % j]) <= 1e-5 * np.maximum(tab["band_power_%d" % j], 1e-3 * tot) + 1e-30)
                if not okj:
                    e = np.abs(got[1 + j] - tab["band_power_%d" % j])
                    i = int(np.argmax(e))
                    why.append("band%d win %d got %.9g want %.9g tot %.9g" % (j, i, got[1 + j][i], tab["band_power_%d" % j][i], tot[i]))
                ok = ok and okj
            psd_ref, freqs = OS.window_psd(x, 500, 250, fs)
            lidx, uidx = OS.first_index(freqs, plo), OS.first_index(freqs, phi)
            pb = got[7].astype(np.int64)
            wantb = lidx + np.argmax(psd_ref[:, lidx:uidx], axis=1)
            for i in np.nonzero(pb != wantb)[0]:
                a_, b_ = psd_ref[i, pb[i]], psd_ref[i, wantb[i]]
                okp = abs(a_ - b_) <= 1e-5 * max(b_, 1e-3 * tot[i])
                if not okp:
                    why.append("peak win %d got bin %d (%.9g) want bin %d (%.9g)" % (i, pb[i], a_, wantb[i], b_))
                ok = ok and okp
            if np.all(tot > 0):
                oke = np.allclose(got[9], tab["spectral_entropy"], rtol=1e-5, atol=1e-9)
                if not oke:
                    e = np.abs(got[9] - tab["spectral_entropy"])
                    i = int(np.argmax(e))
                    why.append("entropy win %d got %.9g want %.9g" % (i, got[9][i], tab["spectral_entropy"][i]))
                ok = ok and oke
            if not ok:
                bad += 1
                print("SPECTRAL-FAST MISMATCH n=%d bands=%s peak=(%g,%g) %s" % (n, bands, plo, phi, "; ".join(why)))
        elif kind == 7:        # generic spectral kernel: random smooth window lengths (even and odd), tones and noise
            fs = float(rng.choice([25.0, 50.0, 64.0, 100.0]))
            W = int(rng.choice([36, 45, 48, 60, 64, 75, 80, 90, 96, 100, 120, 125, 128, 150, 160, 200, 243, 256, 320, 375, 384,
                                512, 600, 625, 750, 1000, 1024]))
            S = int(rng.integers(1, W + 1))
            n = W + S * int(rng.integers(0, 30)) + int(rng.integers(0, S))
            tt = np.arange(n) / fs
            x = (rng.choice([0.0, 1.0]) + 0.4 * np.sin(2 * np.pi * rng.uniform(0.3, fs / 2 - 0.5) * tt) +
                 rng.choice([0.0, 0.0, 0.05, 0.5]) * rng.standard_normal(n)).astype(np.float32)
            lo_, hi_ = float(rng.uniform(0, fs / 4)), float(rng.uniform(fs / 4, fs / 2))
            got = rolling_apply([SP.total_power(fs), SP.band_power(fs, lo_, hi_), SP.peak_bin(fs, lo_, hi_),
                                 SP.spectral_entropy(fs)])(x, W, S)
            tab = OS.spectral_table(x, W, S, fs, [(lo_, hi_)], lo_, hi_)
            tot = tab["total_power"]
            why = []
            if not np.all(np.abs(got[0] - tot) <= 1e-5 * tot):
                why.append("total")
            if not np.all(np.abs(got[1] - tab["band_power_0"]) <= 1e-5 * np.maximum(tab["band_power_0"], 1e-3 * tot)):
                why.append("band")
            if not np.all(np.abs(got[3] - tab["spectral_entropy"]) <= 1e-5 * tab["spectral_entropy"] + 1e-9):
                e = np.abs(got[3] - tab["spectral_entropy"])
                i = int(np.argmax(e))
                why.append("entropy win %d got %.9g want %.9g" % (i, got[3][i], tab["spectral_entropy"][i]))
            psd_ref, freqs = OS.window_psd(x, W, S, fs)
            lidx, uidx = OS.first_index(freqs, lo_), OS.first_index(freqs, hi_)
            if uidx > lidx:
                pb = got[2].astype(np.int64)
                wantb = lidx + np.argmax(psd_ref[:, lidx:uidx], axis=1)
                for i in np.nonzero(pb != wantb)[0]:
                    a_, b_ = psd_ref[i, pb[i]], psd_ref[i, wantb[i]]
                    if not abs(a_ - b_) <= 1e-5 * max(b_, 1e-3 * tot[i]):
                        why.append("peak win %d got %d want %d" % (i, pb[i], wantb[i]))
            if why:
                bad += 1
                print("SPECTRAL-GENERIC MISMATCH W=%d S=%d n=%d fs=%g band=(%g,%g): %s" % (W, S, n, fs, lo_, hi_, "; ".join(why[:4])))
        elif kind == 8:        # HRV time-domain metrics
            from oracle import hrv as OH
            from pymhealth_b200.heart import hrv
            n = int(rng.integers(3, 200000))
            nn = (800 + 60 * rng.standard_normal(n)).clip(300, 2000)
            if rng.random() < 0.3:
                nn = np.round(nn)
            pairs = [("sdnn", hrv.sdnn(nn), OH.sdnn(nn)), ("rmssd", hrv.rmssd(nn), OH.rmssd(nn)), ("ssd", hrv.ssd(nn), OH.ssd(nn)),
                     ("sdsd", hrv.sdsd(nn), OH.sdsd(nn)), ("pnn50", hrv.pnnx(nn, "ms", 50.0), OH.pnnx(nn, "ms", 50.0)),
                     ("pnn20", hrv.pnnx(nn, "ms", 20.0), OH.pnnx(nn, "ms", 20.0))]
            if n > 700:
                pairs += [("sdann", hrv.sdann(nn, unit="ms"), OH.sdann(nn, unit="ms")),
                          ("sdnni", hrv.sdnni(nn, unit="ms"), OH.sdnni(nn, unit="ms"))]
            for nm, g_, w_ in pairs:
                # ssd telescopes to nn[-1] - nn[0]: its rounding error scales with sum |diff|, not with the result
                atol = 1e-12 * float(np.abs(np.diff(nn)).sum()) if nm == "ssd" else 1e-12
                if not np.isclose(g_, w_, rtol=1e-9, atol=atol, equal_nan=True):
                    bad += 1
                    print("HRV MISMATCH n=%d %s got %.15g want %.15g" % (n, nm, g_, w_))
        elif kind == 9:        # accelerometer elementwise
            from oracle import accel as OA
            from pymhealth_b200.inertial import accelerometer as acc
            n = int(rng.integers(1, 300000))
            dt = rng.choice([np.float32, np.float64])
            x, y, z = (rng.standard_normal(n).astype(dt) for _ in range(3))
            for nm, g_, w_ in [("roll", acc.roll(y, z), OA.roll(y, z)), ("pitch", acc.pitch(x, y, z), OA.pitch(x, y, z)),
                               ("magnitude", acc.magnitude(x, y, z), OA.magnitude(x, y, z)),
                               ("magnitude_dot", acc.magnitude_dot(x, y, z), OA.magnitude_dot(x, y, z))]:
                tol = 1e-5 if dt is np.float32 else 1e-12
                if np.shape(g_) != np.shape(w_) or not np.allclose(g_, w_, rtol=tol, atol=tol):
                    bad += 1
                    print("ACCEL MISMATCH n=%d %s %s" % (n, np.dtype(dt).name, nm))
        elif kind == 10:       # haversine forms
            from oracle import location as OLc
            from pymhealth_b200.location import distance as D
            n, m_ = int(rng.integers(1, 50000)), int(rng.integers(1, 40))
            la1, la2 = rng.uniform(-89, 89, n), rng.uniform(-89, 89, n)
            lo1, lo2 = rng.uniform(-180, 180, n), rng.uniform(-180, 180, n)
            if rng.random() < 0.5:              # neighbouring fixes (1 m .. 1 km apart)
                la2 = la1 + rng.normal(0, 1e-4, n)
                lo2 = lo1 + rng.normal(0, 1e-4, n)
            checks = [("elementwise", D.haversine_elementwise(la1, lo1, la2, lo2), OLc.haversine_elementwise(la1, lo1, la2, lo2)),
                      ("vector", D.haversine_vector(la1[0], lo1[0], la2, lo2), OLc.haversine_vector(la1[0], lo1[0], la2, lo2)),
                      ("outer", D.haversine_outer_product(la1[:m_], lo1[:m_], la2[:97], lo2[:97]),
                       OLc.haversine_outer_product(la1[:m_], lo1[:m_], la2[:97], lo2[:97]))]
            for nm, g_, w_ in checks:
                if g_.shape != w_.shape or not np.allclose(g_, w_, rtol=1e-9, atol=1e-12):
                    bad += 1
                    print("HAVERSINE MISMATCH n=%d %s maxrel %.3g" % (n, nm, np.max(np.abs(g_ - w_) / np.maximum(w_, 1e-300))))
        elif kind == 11:       # complex128 FFT drop-in: lengths with prime factors <= 31, real and complex rows
            from pymhealth_b200 import fft as F
            n = int(np.prod(rng.choice([2, 2, 2, 3, 3, 5, 5, 7, 11, 13, 17, 19, 23, 29, 31], int(rng.integers(1, 6)))))
            if n > 4096:
                n = int(rng.integers(1, 700))
                while max((p_ for p_ in range(2, n + 1) if n % p_ == 0 and all(p_ % q for q in range(2, p_))), default=1) > 31:
                    n += 1
            rows = int(rng.integers(1, 40))
            a = rng.standard_normal((rows, n))
            if rng.random() < 0.5:
                a = a + 1j * rng.standard_normal((rows, n))
            fwd, want = F.fft(a), np.fft.fft(a, axis=-1)
            inv, wanti = F.ifft(a), np.fft.ifft(a, axis=-1)
            scale = np.abs(want).max() + 1e-300
            if np.abs(fwd - want).max() > 1e-12 * scale * max(1, np.log2(n)) or np.abs(inv - wanti).max() > 1e-12 * max(1, np.log2(n)) * (np.abs(wanti).max() + 1e-300):
                bad += 1
                print("FFT MISMATCH n=%d rows=%d cplx=%s err %.3g" % (n, rows, np.iscomplexobj(a), np.abs(fwd - want).max() / scale))
        elif kind == 12:       # cluster label statistics
            from oracle import location as OLc
            from pymhealth_b200.location import distribution as DB
            n = int(rng.integers(1, 200000))
            lab = rng.integers(-1, int(rng.integers(0, 300)) + 1, n).astype(np.int64)
            if rng.random() < 0.3:
                lab = np.sort(lab)
            g1, w1 = DB.num_clusters(lab), OLc.num_clusters(lab)
            g2, w2 = DB.cluster_totals(lab), OLc.cluster_totals(lab)
            g3, w3 = DB.cluster_entropy(lab), OLc.cluster_entropy(lab)
            g4, w4 = DB.normalized_cluster_entropy(lab), OLc.normalized_cluster_entropy(lab)
            same_tot = dict(g2) == dict(w2) if isinstance(w2, dict) else np.array_equal(np.asarray(g2), np.asarray(w2))
            if g1 != w1 or not same_tot or not np.isclose(g3, w3, rtol=1e-9, atol=1e-12) or \
                    not np.isclose(g4, w4, rtol=1e-9, atol=1e-12, equal_nan=True):
                bad += 1
                print("LABEL MISMATCH n=%d: %r %r | %r %r | %r %r" % (n, g1, w1, g3, w3, g4, w4))
        elif kind == 13:       # whole-series forms: gradient, zero_crossings, slope_sum
            from oracle import reducers as ORd
            from pymhealth_b200.heart import ppg
            n = int(rng.integers(2, 300000))
            x = rng.standard_normal(n) * rng.choice([1.0, 100.0])
            if rng.random() < 0.5:
                x = x.astype(np.float32)
            th = float(rng.choice([0.0, 0.1, 0.5]))
            g_ = timedom.gradient(x)
            w_ = np.gradient(x.astype(np.float64)) if x.dtype == np.float64 else ORd.gradient(x)
            if not np.allclose(g_, w_, rtol=1e-6 if x.dtype == np.float32 else 1e-12, atol=1e-6 if x.dtype == np.float32 else 1e-12):
                bad += 1
                print("GRADIENT MISMATCH n=%d %s" % (n, x.dtype))
            xz = x.copy()
            xz[np.abs(xz) <= th] = 0
            pos = xz > 0
            if not np.array_equal(timedom.zero_crossings(x, th), np.logical_xor(pos[:-1], pos[1:])):
                bad += 1
                print("ZERO-CROSSINGS MISMATCH n=%d th=%g" % (n, th))
            w = int(rng.integers(1, 64))
            xs = x[:20000].astype(np.float64)
            # ppg.py:28-42 restated: out[i] = sum(diff(x)[i-w:i]) for w <= i < len(x) - 1, zero elsewhere
            dx = np.diff(xs)
            cs = np.concatenate([[0.0], np.cumsum(dx)])
            want = np.zeros(len(xs))
            ii = np.arange(w, len(xs) - 1)
            want[ii] = cs[ii] - cs[ii - w]
            got = ppg.slope_sum(xs, w)
            if got.shape != want.shape or not np.allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(xs).max()):
                bad += 1
                print("SLOPE-SUM MISMATCH n=%d w=%d" % (len(xs), w))
        elif kind == 14:       # get_indices / indices_rolling_apply on an irregular integer index
            from pymhealth_b200.util.windows import get_indices, indices_rolling_apply
            n = int(rng.integers(2, 50000))
            idx = np.cumsum(rng.integers(1, 1000, n)).astype(np.int64) + int(rng.integers(-10**6, 10**6))
            wsize, wstep = int(rng.integers(1, 20000)), int(rng.integers(1, 20000))
            gi, wi = get_indices(idx, wsize, wstep), OW.get_indices(idx, wsize, wstep)
            if not np.array_equal(gi, wi):
                bad += 1
                print("GET-INDICES MISMATCH n=%d wsize=%d wstep=%d" % (n, wsize, wstep))
            else:
                x = rng.standard_normal(n)
                minlen = int(rng.integers(1, 4))
                got = indices_rolling_apply(np.var, minlen)(gi, x, minlen)
                want = OW.indices_rolling("var", wi, x, minlen)
                if not np.allclose(got, want, rtol=1e-9, atol=1e-12, equal_nan=True):
                    bad += 1
                    print("INDICES-ROLLING MISMATCH n=%d wsize=%d wstep=%d" % (n, wsize, wstep))
        elif kind == 15:       # spectral W = 1920 / S = 64 kernel: ragged batches, tones (redo pass), several columns
            fs = 64.0
            n = 1920 + 64 * int(rng.integers(0, 60)) + int(rng.integers(0, 64))
            tt = np.arange(n) / fs
            x = (rng.choice([0.0, 2.0]) + 0.4 * np.sin(2 * np.pi * rng.uniform(0.3, 30) * tt) +
                 rng.choice([0.0, 0.0, 0.05, 0.5]) * rng.standard_normal(n)).astype(np.float32)
            lo_, hi_ = float(rng.uniform(0, 8)), float(rng.uniform(8, 32))
            got = rolling_apply([SP.total_power(fs), SP.band_power(fs, lo_, hi_), SP.peak_bin(fs, lo_, hi_),
                                 SP.spectral_entropy(fs), SP.relative_band_power(fs, 0.0, lo_)])(x, 1920, 64)
            tab = OS.spectral_table(x, 1920, 64, fs, [(lo_, hi_), (0.0, lo_)], lo_, hi_)
            tot = tab["total_power"]
            ok = np.all(np.abs(got[0] - tot) <= 1e-5 * tot)
            ok = ok and np.all(np.abs(got[1] - tab["band_power_0"]) <= 1e-5 * np.maximum(tab["band_power_0"], 1e-3 * tot))
            ok = ok and np.all(np.abs(got[4] - tab["rel_band_power_1"]) <= 1e-5 * np.maximum(tab["rel_band_power_1"], 1e-3))
            ok = ok and np.all(np.abs(got[3] - tab["spectral_entropy"]) <= 1e-5 * tab["spectral_entropy"] + 1e-9)
            psd_ref, freqs = OS.window_psd(x, 1920, 64, fs)
            lidx, uidx = OS.first_index(freqs, lo_), OS.first_index(freqs, hi_)
            pb = got[2].astype(np.int64)
            wantb = lidx + np.argmax(psd_ref[:, lidx:uidx], axis=1)
            for i in np.nonzero(pb != wantb)[0]:
                a_, b_ = psd_ref[i, pb[i]], psd_ref[i, wantb[i]]
                ok = ok and abs(a_ - b_) <= 1e-5 * max(b_, 1e-3 * tot[i])
            if not ok:
                bad += 1
                print("SPECTRAL-W1920 MISMATCH n=%d band=(%g,%g)" % (n, lo_, hi_))
        elif kind == 16:       # fused magnitude statistics against the two-step form and the oracle
            from oracle import accel as OA
            from pymhealth_b200.inertial import accelerometer as acc
            g = int(rng.choice([10, 25, 32, 50, 64, 100, 125, 250]))
            k, hop = int(rng.integers(1, 12)), int(rng.integers(1, 6))
            W, S = g * k, g * hop
            n = W + S * int(rng.integers(0, 60)) + int(rng.integers(0, S))
            dt = rng.choice([np.float32, np.float64])
            x, y, z = ((rng.standard_normal(n) + o).astype(dt) for o in (0.0, 0.3, 1.0))
            got = acc.rolling_magnitude([np.mean, np.std, np.min, stats.kurtosis, timedom.line_length], W, S)(x, y, z)
            mag = OA.magnitude(x, y, z)
            for g_, nm in zip(got, ["mean", "std", "min", "kurtosis", "line_length"]):
                want = OW.rolling(nm, mag, W, S)
                if not np.allclose(g_, want, rtol=1e-6 if nm == "line_length" else 1e-9, atol=1e-12):
                    bad += 1
                    print("MAGNITUDE MISMATCH W=%d S=%d n=%d %s %s" % (W, S, n, np.dtype(dt).name, nm))
        elif kind == 17:       # engine plumbing: strided multi-series views (aligned or not), families interleaved in the
            import torch       # column order, float32 / float64 tables
            from pymhealth_b200 import engine
            fs = 50.0
            g = int(rng.choice([16, 25, 50, 64, 125, 250]))
            k, hop = int(rng.integers(1, 5)), int(rng.integers(1, 4))
            W, S = g * k, g * hop
            if W % 2 or W < 8:
                W, S = 500, 250
            ns = int(rng.integers(1, 6))
            n = W + S * int(rng.integers(0, 50)) + int(rng.integers(0, S))
            off = int(rng.integers(0, 7))
            big = torch.from_numpy((rng.standard_normal((ns, n + 11)) + 1.0).astype(np.float32)).cuda()
            xv = big[:, off:off + n]
            pool = [("mean", stats.mean.feature()), ("median", stats.median.feature()), ("S:total", SP.total_power(fs).feature()),
                    ("kurtosis", stats.kurtosis.feature()), ("S:entropy", SP.spectral_entropy(fs).feature()),
                    ("max", stats.dmax.feature()), ("percentile", stats.percentile.feature(25.0)),
                    ("line_length", timedom.line_length.feature()), ("S:band", SP.band_power(fs, 1.0, 9.0).feature())]
            order = rng.permutation(len(pool))[:int(rng.integers(1, len(pool) + 1))]
            feats = [pool[i] for i in order]
            odt = torch.float64 if rng.random() < 0.5 else torch.float32
            tab = engine.window_table(xv, W, S, [f for _, f in feats], fs=fs, out_dtype=odt).cpu().numpy().astype(np.float64)
            xh = xv.cpu().numpy()
            ftol = 1e-9 if odt is torch.float64 else 3e-7
            for si in range(ns):
                sp = OS.spectral_table(xh[si], W, S, fs, [(1.0, 9.0)], None, None)
                for c, (nm, _) in enumerate(feats):
                    if nm.startswith("S:"):
                        want = {"S:total": sp["total_power"], "S:entropy": sp["spectral_entropy"], "S:band": sp["band_power_0"]}[nm]
                        okc = np.all(np.abs(tab[si, :, c] - want) <= (1e-5 + ftol) * np.maximum(np.abs(want), 1e-3 * sp["total_power"] if nm != "S:entropy" else 0) + 1e-9)
                    else:
                        want = OW.rolling(nm, xh[si], W, S, 25.0 if nm == "percentile" else None)
                        okc = np.allclose(tab[si, :, c], want, rtol=max(ftol, 1e-6 if nm == "line_length" else 1e-9), atol=1e-12 if odt is torch.float64 else 1e-6)
                    if not okc:
                        bad += 1
                        print("ENGINE MISMATCH W=%d S=%d ns=%d n=%d off=%d %s col %d (%s) of %s" % (W, S, ns, n, off, odt, c, nm, [a for a, _ in feats]))
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
