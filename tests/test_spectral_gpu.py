"""Kernel 2 (window FFT + PSD reducers, batched DFT) vs fixtures and the oracle.

Tolerances.  The window FFT computes in float32 with float64 sums: a spectral column c of a window is
accepted when |c - ref| <= 1e-5 * max(|ref|, total_power_of_that_window * 1e-3) for powers (a band
that holds < 0.1 % of the window's power is judged against that floor), 1e-5 relative for entropy,
and exactly for peak bins unless the reference's own two largest candidates are within 1e-5
(a tie no float32 pipeline can order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check_power(got, want, total, what):
    tol = 1e-5 * np.maximum(np.abs(want), 1e-3 * total)
    bad = np.abs(got - want) > tol
    assert not bad.any(), "%s: %d/%d off, worst rel %g" % (what, bad.sum(), bad.size,
                                                           np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300)))


def _check_peaks(got_bin, psd_ref, lidx, uidx, what):
    want = lidx + np.argmax(psd_ref[:, lidx:uidx], axis=1)
    for i in np.nonzero(got_bin != want)[0]:
        a, b = psd_ref[i, int(got_bin[i])], psd_ref[i, want[i]]
        assert abs(a - b) <= 1e-5 * b, "%s: window %d picked bin %d (%.6g) over %d (%.6g)" % (
            what, i, got_bin[i], a, want[i], b)


@pytest.mark.parametrize("case", ["acc", "ppg", "odd", "acc-generic"])
def test_golden_spectral(ref_spectral, case, monkeypatch):
    if case.endswith("-generic"):           # W=500/S=250 normally takes the batched kernel: cover the generic one too
        monkeypatch.setenv("MHB_SPECTRAL_GENERIC", "1")
        case = case.split("-")[0]
    from pymhealth_b200 import spectral as SP
    from pymhealth_b200.util import rolling_apply
    from oracle import spectral as OS
    x = ref_spectral[case + "/x"]
    W, S, fs = ref_spectral[case + "/wsf"]
    W, S = int(W), int(S)
    bands = [tuple(b) for b in ref_spectral[case + "/bands"]]
    funcs = {"total": SP.total_power(fs), "entropy": SP.spectral_entropy(fs), "peak": SP.peak_frequency(fs, 0.3, 12.0),
             "peak_bin": SP.peak_bin(fs, 0.3, 12.0), "peak_all": SP.peak_frequency(fs)}
    for j, (lo, hi) in enumerate(bands):
        funcs["bp%d" % j] = SP.band_power(fs, lo, hi)
        funcs["rbp%d" % j] = SP.relative_band_power(fs, lo, hi)
    got = rolling_apply(funcs)(x, W, S)
    tot = ref_spectral[case + "/total_power"]
    _check_power(got["total"], tot, tot, "total")
    for j in range(len(bands)):
        _check_power(got["bp%d" % j], ref_spectral[case + "/band_power"][:, j], tot, "band %d" % j)
        _check_power(got["rbp%d" % j], ref_spectral[case + "/rel_band_power"][:, j], np.ones_like(tot), "rel band %d" % j)
    np.testing.assert_allclose(got["entropy"], ref_spectral[case + "/entropy"], rtol=1e-5)
    psd_ref, freqs = OS.window_psd(x, W, S, fs)
    lidx, uidx = OS.first_index(freqs, 0.3), OS.first_index(freqs, 12.0)
    _check_peaks(got["peak_bin"].astype(np.int64), psd_ref, lidx, uidx, "peak bin")
    # the frequency column against the reference's own values: equal, except where the bin check above found a tie
    ties = got["peak_bin"].astype(np.int64) != (lidx + np.argmax(psd_ref[:, lidx:uidx], axis=1))
    np.testing.assert_array_equal(got["peak"][~ties], ref_spectral[case + "/peak_frequency_0.3_12"][~ties])
    assert ties.sum() <= max(1, len(ties) // 100), "%d of %d windows are arg-max ties" % (ties.sum(), len(ties))
    np.testing.assert_array_equal(got["peak"], freqs[got["peak_bin"].astype(np.int64)])
    _check_peaks(np.round(got["peak_all"] / freqs[1]).astype(np.int64), psd_ref, 0, len(freqs), "peak all")
    # raw PSD rows
    psd, fr = SP.window_psd(x, W, S, fs)
    np.testing.assert_array_equal(fr, freqs)
    assert psd.shape == psd_ref.shape
    err = np.abs(psd - psd_ref).max(axis=1)
    assert np.all(err <= 1e-5 * psd_ref.max(axis=1))


@pytest.mark.parametrize("n", [500, 749, 4250, 4251, 8137, 12500 + 250 * 16 * 3])
def test_batched_kernel_ragged_batches(n):
    """The W=500/S=250 fast path works on batches of 16 windows: cover short / ragged last batches, several series,
    an unaligned base pointer and a row stride that defeats TMA, against the generic kernel and the oracle."""
    import torch
    from oracle import spectral as OS
    from pymhealth_b200 import engine, synth, spectral as SP
    x = np.stack([synth.accelerometer(40 + s, n)[s % 3] for s in range(5)])
    feats = [SP.total_power(50.0).feature(), SP.band_power(50.0, 0.5, 3.0).feature(), SP.peak_bin(50.0, 0.3, 12.0).feature(),
             SP.spectral_entropy(50.0).feature()]
    xt = torch.from_numpy(x).cuda()
    got = engine.window_table(xt, 500, 250, feats, fs=50.0, out_dtype=torch.float64).cpu().numpy()
    wide = torch.zeros((5, n + 3), dtype=torch.float32, device="cuda")        # stride not a multiple of 4 -> no TMA
    wide[:, :n] = xt
    got2 = engine.window_table(wide[:, :n], 500, 250, feats, fs=50.0, out_dtype=torch.float64).cpu().numpy()
    np.testing.assert_array_equal(got, got2)
    for s in range(5):
        tab = OS.spectral_table(x[s], 500, 250, 50.0, [(0.5, 3.0)], 0.3, 12.0)
        _check_power(got[s, :, 0], tab["total_power"], tab["total_power"], "total")
        _check_power(got[s, :, 1], tab["band_power_0"], tab["total_power"], "band")
        np.testing.assert_allclose(got[s, :, 3], tab["spectral_entropy"], rtol=1e-5)
        psd_ref, freqs = OS.window_psd(x[s], 500, 250, 50.0)
        _check_peaks(got[s, :, 2].astype(np.int64), psd_ref, OS.first_index(freqs, 0.3), OS.first_index(freqs, 12.0), "peak")


@pytest.mark.parametrize("n", [1920, 1920 + 64 * 3, 1920 + 64 * 4 + 13, 9000, 1920 + 64 * 41])
def test_w1920_kernel_ragged_batches(n, monkeypatch):
    """The W=1920/S=64 fast path works on batches of 4 windows: short / ragged last batches, several series, a row
    stride that defeats TMA, many columns -- against the generic kernel and the oracle."""
    import torch
    from oracle import spectral as OS
    from pymhealth_b200 import engine, synth, spectral as SP
    x = np.stack([synth.ppg(70 + s, n) + (3.0 if s == 1 else 0.0) for s in range(3)]).astype(np.float32)
    fs = 64.0
    feats = [SP.total_power(fs).feature(), SP.band_power(fs, 0.5, 3.0).feature(), SP.band_power(fs, 0.0, 40.0).feature(),
             SP.relative_band_power(fs, 3.0, 8.0).feature(), SP.peak_bin(fs, 0.3, 12.0).feature(),
             SP.peak_bin(fs).feature(), SP.spectral_entropy(fs).feature()]
    xt = torch.from_numpy(x).cuda()
    got = engine.window_table(xt, 1920, 64, feats, fs=fs, out_dtype=torch.float64).cpu().numpy()
    wide = torch.zeros((3, n + 3), dtype=torch.float32, device="cuda")        # stride not a multiple of 4 -> no TMA
    wide[:, :n] = xt
    got2 = engine.window_table(wide[:, :n], 1920, 64, feats, fs=fs, out_dtype=torch.float64).cpu().numpy()
    np.testing.assert_array_equal(got, got2)
    monkeypatch.setenv("MHB_SPECTRAL_GENERIC", "1")
    gen = engine.window_table(xt, 1920, 64, feats, fs=fs, out_dtype=torch.float64).cpu().numpy()
    monkeypatch.delenv("MHB_SPECTRAL_GENERIC")
    for s in range(3):
        tab = OS.spectral_table(x[s], 1920, 64, fs, [(0.5, 3.0), (0.0, 40.0), (3.0, 8.0)], 0.3, 12.0)
        tot = tab["total_power"]
        _check_power(got[s, :, 0], tot, tot, "total")
        _check_power(got[s, :, 1], tab["band_power_0"], tot, "band 0")
        _check_power(got[s, :, 2], tab["band_power_1"], tot, "band 1 (includes bin 0)")
        _check_power(got[s, :, 3], tab["rel_band_power_2"], np.ones_like(tot), "rel band 2")
        np.testing.assert_allclose(got[s, :, 6], tab["spectral_entropy"], rtol=1e-5)
        psd_ref, freqs = OS.window_psd(x[s], 1920, 64, fs)
        _check_peaks(got[s, :, 4].astype(np.int64), psd_ref, OS.first_index(freqs, 0.3), OS.first_index(freqs, 12.0), "peak")
        _check_peaks(got[s, :, 5].astype(np.int64), psd_ref, 0, len(freqs), "peak all")
        _check_power(gen[s, :, 0], tot, tot, "generic total")
        np.testing.assert_allclose(gen[s, :, 6], tab["spectral_entropy"], rtol=1e-5)


@pytest.mark.parametrize("W,S,fs", [(500, 250, 50.0), (1920, 64, 64.0), (256, 128, 50.0), (75, 25, 50.0)])
def test_noiseless_tone_entropy(W, S, fs):
    """One bin (not bin 0) holds nearly all the power: H ~ 1e-5 .. 1e-1.  The float32 evaluation of p ln p next to p = 1
    would be off by ~1e-7 absolute (1e-3 relative here); the kernels switch to the log1p form of psd_entropy.cuh -- the
    W=1920 kernel through redo markers and a second pass of the generic kernel.  Tolerance: 1e-5 relative plus 1e-9
    absolute (an exactly periodic tone has H ~ 1e-14 in float64, which is rounding noise of either transform)."""
    import torch
    from oracle import spectral as OS
    from pymhealth_b200 import engine, spectral as SP
    rng = np.random.default_rng(W)
    small = 0
    for trial in range(12):
        n = W + S * int(rng.integers(3, 40)) + int(rng.integers(0, S))
        f0 = rng.uniform(0.3, fs / 2 - 0.5)
        if trial % 4 == 0:
            f0 = round(f0 * W / fs) * fs / W            # exactly on a bin
        off = [0.0, 0.0, 1.0, -0.02][trial % 4 if trial >= 4 else 0]
        x = (off + 0.4 * np.sin(2 * np.pi * f0 * np.arange(n) / fs)).astype(np.float32)
        want = OS.spectral_table(x, W, S, fs, [], None, None)["spectral_entropy"]
        small += int(np.sum(want < 0.05))
        for dt in (torch.float64, torch.float32):
            tab = engine.window_table(torch.from_numpy(x[None]).cuda(), W, S,
                                      [SP.total_power(fs).feature(), SP.spectral_entropy(fs).feature()], fs=fs, out_dtype=dt)
            got = tab[0, :, 1].cpu().numpy().astype(np.float64)
            tol = 1e-5 if dt is torch.float64 else 2e-7 + 1e-5          # + the rounding of a float32 cell
            assert np.all(np.abs(got - want) <= tol * want + 1e-9), (trial, f0, off, np.abs(got - want).max())
    assert small > 0            # the case above really occurred


@pytest.mark.parametrize("W,S", [(36, 9), (64, 16), (90, 30), (96, 32), (100, 50), (120, 40), (128, 64), (256, 128), (384, 96),
                                 (1000, 500), (1024, 512), (75, 25)])
def test_generic_kernel_window_lengths(W, S):
    """The generic warp-per-window kernel over window lengths that exercise every register radix of the planner
    (16 / 12 / 10 / 8 / 6 composites, 5 / 4 / 3 / 2, odd W) against the oracle."""
    from oracle import spectral as OS
    from pymhealth_b200 import synth, spectral as SP
    from pymhealth_b200.util import rolling_apply
    fs = 50.0
    x = synth.accelerometer(90 + W, 6 * W + 17)[2]
    got = rolling_apply({"total": SP.total_power(fs), "bp": SP.band_power(fs, 1.0, 6.0), "pk": SP.peak_bin(fs, 0.5, 12.0),
                         "h": SP.spectral_entropy(fs)})(x, W, S)
    tab = OS.spectral_table(x, W, S, fs, [(1.0, 6.0)], 0.5, 12.0)
    _check_power(got["total"], tab["total_power"], tab["total_power"], "total W=%d" % W)
    _check_power(got["bp"], tab["band_power_0"], tab["total_power"], "band W=%d" % W)
    np.testing.assert_allclose(got["h"], tab["spectral_entropy"], rtol=1e-5)
    psd_ref, freqs = OS.window_psd(x, W, S, fs)
    _check_peaks(got["pk"].astype(np.int64), psd_ref, OS.first_index(freqs, 0.5), OS.first_index(freqs, 12.0), "peak W=%d" % W)


def test_fft_dropin(ref_spectral):
    from pymhealth_b200 import fft as F
    for case in ("acc", "ppg", "odd"):
        x = ref_spectral[case + "/x"]
        W = int(ref_spectral[case + "/wsf"][0])
        spec = F.fft(x[:W].astype(np.float64))
        want = ref_spectral[case + "/fft0"]
        assert spec.dtype == np.complex128
        np.testing.assert_allclose(spec, want, rtol=0, atol=1e-12 * np.abs(want).max())
        back = F.ifft(want)
        np.testing.assert_allclose(back, ref_spectral[case + "/ifft0"], rtol=0, atol=1e-13 * np.abs(x[:W]).max() + 1e-15)
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 7, 31, 60, 250, 500, 961, 1000, 1920, 2 * 3 * 5 * 7 * 11):
        a = rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))
        np.testing.assert_allclose(F.fft(a), np.fft.fft(a, axis=1), rtol=0, atol=1e-12 * n)
        np.testing.assert_allclose(F.ifft(F.fft(a)), a, rtol=0, atol=1e-13 * n)           # round trip
    with pytest.raises(NotImplementedError):
        F.fft(np.zeros(37 * 4))


def test_psd_reducers_literal_signatures(ref_spectral):
    from pymhealth_b200.heart import hrv
    from pymhealth_b200.generic.frequency import density
    from pymhealth_b200.generic import information
    from oracle import spectral as OS
    x = ref_spectral["acc/x"]
    psd, freqs = OS.window_psd(x, 500, 250, 50.0)
    for i in (0, 7, psd.shape[0] - 1):
        assert hrv.power_band(psd[i], freqs, 0.5, 3.0) == pytest.approx(ref_spectral["acc/band_power"][i, 0], rel=1e-13)
        assert hrv.relative_power_band(psd[i], freqs, 3.0, 8.0) == pytest.approx(ref_spectral["acc/rel_band_power"][i, 1], rel=1e-13)
        assert hrv.power_band(psd[i], freqs) == pytest.approx(ref_spectral["acc/total_power"][i], rel=1e-13)
        assert density.peak_frequency(psd[i], freqs, 0.3, 12.0) == ref_spectral["acc/peak_frequency_0.3_12"][i]
        assert density.peak_frequency(psd[i], freqs) == ref_spectral["acc/peak_frequency_all"][i]
        assert information.entropy(psd[i]) == pytest.approx(ref_spectral["acc/entropy"][i], rel=1e-12)
    assert density.first_index(freqs, 0.3) == OS.first_index(freqs, 0.3)


def test_linearity_and_parseval_full_config2():
    """Config-2 size properties: Parseval (sum of the two-sided PSD = W * sum x^2) and scaling."""
    import torch
    from pymhealth_b200 import synth, engine, spectral as SP
    x = torch.from_numpy(synth.accelerometer(1, 1_000_000)).cuda()
    W, S = 500, 250
    psd, _ = SP.window_psd(x, W, S, 50.0, out_dtype=torch.float64)
    two_sided = psd[..., 0] + 2 * psd[..., 1:-1].sum(-1) + psd[..., -1]
    win = x.double().unfold(1, W, S)
    torch.testing.assert_close(two_sided, W * (win * win).sum(-1), rtol=2e-6, atol=0)
    psd2, _ = SP.window_psd(x * 3.0, W, S, 50.0, out_dtype=torch.float64)
    err = (psd2 - psd * 9.0).abs().amax(dim=-1)
    assert bool((err <= 1e-5 * 9.0 * psd.amax(dim=-1)).all())      # float32 FFT: error scales with the row's peak


def test_hrv_peak_frequency_intent():
    """heart.hrv.peak_frequency: the reference (heart/hrv.py:182-189) returns freqs[argmax(psd[mask])] -- the UNMASKED
    frequency vector indexed with the masked arg-max, correct only while the mask starts at bin 0.  The drop-in
    implements the evident intent, freqs[mask][argmax(psd[mask])] (both bounds inclusive, first maximum), and agrees
    with the reference's own expression wherever that expression is right (lower = None / lower <= min f)."""
    from pymhealth_b200.heart import hrv
    rng = np.random.default_rng(11)
    freqs = np.fft.rfftfreq(500, 1 / 50.0)
    for trial in range(12):
        psd = rng.gamma(2.0, 1.0, freqs.size)
        psd[rng.integers(0, freqs.size)] += 25.0                       # a clear peak somewhere
        for lower, upper in [(None, None), (None, 7.5), (0.0, 12.0), (0.3, 12.0), (3.0, 8.0), (8.0, 8.0), (2.05, 2.35)]:
            lo = freqs.min() if lower is None else lower
            hi = freqs.max() if upper is None else upper
            mask = np.logical_and(freqs >= lo, freqs <= hi)
            if not mask.any():
                continue
            want = freqs[mask][np.argmax(psd[mask])]
            got = hrv.peak_frequency(psd, freqs, lower, upper)
            assert got == want, (trial, lower, upper, got, want)
            if lo <= freqs.min():                                      # where the reference's formula is correct
                assert got == freqs[np.argmax(psd[mask])]
    # ties: the first maximum wins, as np.argmax does
    psd = np.ones(freqs.size)
    assert hrv.peak_frequency(psd, freqs, 1.0, 2.0) == freqs[np.nonzero(freqs >= 1.0)[0][0]]
