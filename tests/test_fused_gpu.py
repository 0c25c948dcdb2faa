"""Statistical + spectral columns of ONE call (mhb_window_features_f32: kernels 1a and 2 back to back), through the
drop-in API and through the C-ABI table call, against the oracle, the reference-generated fixtures and the separate
calls; the host-buffer pipeline on float32 samples and on raw int16 counts.  Integer-valued columns bit-exact; statistics rtol 1e-9 (contract 1e-5); spectral columns to the
tolerances of test_spectral_gpu.py."""
import numpy as np
import pytest

from test_window_stats_gpu import assert_feature_close, _reducers

pytestmark = pytest.mark.gpu

STREAM = ["mean", "var", "std", "min", "max", "drange", "skewness", "kurtosis", "kurtosis_excess",
          "coeff_var", "zero_crossing_count", "line_length", "hjorth_activity"]
FS = 50.0


def _spectral_reducers():
    from pymhealth_b200 import spectral as SP
    return [SP.total_power(FS), SP.band_power(FS, 0.5, 3.0), SP.band_power(FS, 3.0, 8.0), SP.relative_band_power(FS, 0.5, 3.0),
            SP.peak_bin(FS, 0.3, 12.0), SP.spectral_entropy(FS)]


def _check_spectral(got, x):
    from oracle import spectral as OS
    tab = OS.spectral_table(x, 500, 250, FS, [(0.5, 3.0), (3.0, 8.0)], 0.3, 12.0)
    tot = tab["total_power"]
    floor = 1e-3 * tot
    for g, k in zip(got[:4], ["total_power", "band_power_0", "band_power_1", "rel_band_power_0"]):
        want = tab[k]
        fl = floor if k != "rel_band_power_0" else 1e-3
        assert np.all(np.abs(g - want) <= 1e-5 * np.maximum(np.abs(want), fl)), k
    peak = got[4]
    bad = peak != tab["peak_bin"]
    if bad.any():
        psd, _ = OS.window_psd(x, 500, 250, FS)
        rows = np.nonzero(bad)[0]
        a = psd[rows, peak[rows].astype(int)]
        b = psd[rows, tab["peak_bin"][rows].astype(int)]
        assert np.all(np.abs(a - b) <= 1e-5 * np.maximum(a, b)), "peak bins differ beyond a 1e-5 tie"
    assert np.all(np.abs(got[5] - tab["spectral_entropy"]) <= 1e-5 * tab["spectral_entropy"] + 1e-9), "entropy"


@pytest.mark.parametrize("n", [500, 749, 750, 4250, 4251, 4500, 8249, 12345, 40000, 200003])
def test_fused_vs_oracle(n):
    """Every batch shape: one window, a ragged last batch (1..16 windows), several CTAs' worth of batches."""
    from oracle import windows as OW
    from pymhealth_b200.util import rolling_apply
    R = _reducers()
    rng = np.random.default_rng(n)
    t = np.arange(n)
    x = (0.9 + 0.2 * rng.standard_normal(n) + 0.4 * np.sin(t * 0.21) + 0.1 * np.sin(t * 0.013)).astype(np.float32)
    funcs = [R[k] for k in STREAM] + _spectral_reducers()
    got = rolling_apply(funcs)(x, 500, 250)
    for name, g in zip(STREAM, got):
        want = OW.rolling(name, x, 500, 250, 0.0 if name == "zero_crossing_count" else None)
        assert g.shape == want.shape
        assert_feature_close(name, g, want, x, 1e-9)
    _check_spectral(got[len(STREAM):], x)


def test_fused_golden(ref_windows):
    """Reference-generated fixtures (tests/golden/make_golden.py) for the two W = 500 / S = 250 cases."""
    from pymhealth_b200.util import rolling_apply
    R = _reducers()
    for case in ("acc_z_500_250", "acc_x_500_250"):
        x = ref_windows[case + "/x"]
        names = [k for k in STREAM if k != "zero_crossing_count"]
        got = rolling_apply([R[k] for k in names] + [R["zero_crossing_count"]] + _spectral_reducers())(x, 500, 250)
        for name, g in zip(names, got):
            assert_feature_close(name, g, ref_windows[case + "/" + name], x, 1e-9)
        np.testing.assert_array_equal(got[len(names)], ref_windows[case + "/zero_crossing_count:0"])
        _check_spectral(got[len(names) + 1:], x)


def test_fused_equals_separate_kernels():
    """Many series, a threshold for the crossings, float32 table: the fused call against kernel 1a and the spectral-only
    call on the same device buffer (identical selections; sums agree to float32 rounding of the table)."""
    import torch
    from pymhealth_b200 import engine, synth, spectral as SP
    from pymhealth_b200.generic import stats, timedom
    dev = torch.device("cuda:0")
    x = synth.device_accelerometer(3, 123_457, dev).view(9, -1)
    sf = [stats.mean.feature(), stats.std.feature(), stats.var.feature(), stats.dmin.feature(), stats.dmax.feature(),
          stats.drange.feature(), stats.skewness.feature(), stats.kurtosis.feature(),
          timedom.zero_crossing_count.feature(0.03), timedom.line_length.feature()]
    pf = [r.feature() for r in _spectral_reducers()]
    fused = engine.window_table(x, 500, 250, sf + pf, zc_threshold=0.03, fs=FS, out_dtype=torch.float64)
    a = engine.window_table(x, 500, 250, sf, zc_threshold=0.03, out_dtype=torch.float64)
    b = engine.window_table(x, 500, 250, pf, fs=FS, out_dtype=torch.float64)
    fa, fb = fused[:, :, :len(sf)].cpu().numpy(), fused[:, :, len(sf):].cpu().numpy()
    a, b = a.cpu().numpy(), b.cpu().numpy()
    for j in (3, 4, 5, 8):                               # min, max, drange, zero crossings: bit-exact
        np.testing.assert_array_equal(fa[:, :, j], a[:, :, j])
    for j in range(len(sf)):                             # line length (j = 9): float32 partial sums in both kernels
        np.testing.assert_allclose(fa[:, :, j], a[:, :, j], rtol=1e-6 if j == 9 else 1e-9, atol=1e-12)
    np.testing.assert_array_equal(fb, b)                 # the same spectral code path
    # a second run gives the same bits (fixed reduction orders; no atomics)
    again = engine.window_table(x, 500, 250, sf + pf, zc_threshold=0.03, fs=FS, out_dtype=torch.float64)
    assert torch.equal(fused, again)


def test_fused_unaligned_rows_and_strided_table():
    """Rows that start off a 16-byte boundary take the guarded-copy path; the table is a column range of a wider array."""
    import torch
    from pymhealth_b200 import engine, synth, spectral as SP
    from pymhealth_b200.generic import stats, timedom
    dev = torch.device("cuda:0")
    base = synth.device_accelerometer(1, 30_001 * 3 + 7, dev).view(-1)
    x = torch.as_strided(base[1:], (3, 30_000), (30_001, 1))          # unaligned base, odd row stride
    sf = [stats.mean.feature(), stats.var.feature(), timedom.zero_crossing_count.feature(0.0), timedom.line_length.feature()]
    pf = [SP.total_power(FS).feature(), SP.spectral_entropy(FS).feature()]
    nw = engine.n_windows(30_000, 500, 250)
    wide = torch.full((3, nw, 9), -7.0, dtype=torch.float32, device=dev)
    engine.window_table(x, 500, 250, sf + pf, fs=FS, out=wide[:, :, 2:8])
    ref = engine.window_table(x.contiguous(), 500, 250, sf + pf, fs=FS)
    assert torch.equal(wide[:, :, 2:8], ref)
    assert bool((wide[:, :, :2] == -7.0).all()) and bool((wide[:, :, 8] == -7.0).all())


def test_pipeline_host_buffers_float_and_raw_counts():
    """FeaturePipeline.run: the returned HOST table is complete when the call returns (no external synchronize), and
    int16 raw counts (2 bytes per sample over PCIe, widened on the device) give the table of the float32 samples
    count * scale bit for bit."""
    import torch
    from pymhealth_b200 import engine, synth, spectral as SP
    from pymhealth_b200.generic import stats, timedom
    from pymhealth_b200.pipeline import FeaturePipeline
    feats = [stats.mean.feature(), stats.std.feature(), stats.dmin.feature(), timedom.zero_crossing_count.feature(0.0),
             SP.total_power(FS).feature(), SP.peak_bin(FS, 0.3, 12.0).feature(), SP.spectral_entropy(FS).feature()]
    x = np.concatenate([synth.accelerometer(70 + s, 60_000) for s in range(3)])        # [9, 60000]
    counts = np.round(x * 4096.0).astype(np.int16)
    xq = (counts.astype(np.float32) * np.float32(1.0 / 4096.0))
    pipe_f = FeaturePipeline(feats, 500, 250, fs=FS, chunk_series=4)
    pipe_c = FeaturePipeline(feats, 500, 250, fs=FS, chunk_series=4, count_scale=1.0 / 4096.0)
    for _ in range(3):                                                                 # repeated runs reuse the staging buffers
        host_f = pipe_f.run(xq).numpy().copy()                                         # read at once: must be complete
        host_c = pipe_c.run(counts).numpy().copy()
        want = engine.window_table(torch.from_numpy(xq).cuda(), 500, 250, feats, fs=FS).cpu().numpy()
        np.testing.assert_array_equal(host_f, want)
        np.testing.assert_array_equal(host_c, want)
    assert pipe_f.launches_per_chunk() == 2


def test_tables_do_not_depend_on_the_batch_composition():
    """The rows of a series are the same bits whether it is processed alone or as one of many series (kernel 1a pivots sit
    on fixed segments of the series, kernel 2 works on batches aligned to the series start): what makes the per-rank
    tables of a sharded run identical to a single-GPU run whatever the shard sizes."""
    import torch
    from pymhealth_b200 import engine, synth, spectral as SP
    from pymhealth_b200.generic import stats, timedom
    dev = torch.device("cuda:0")
    x = synth.device_accelerometer(8, 1_000_003, dev).view(24, -1)
    sf = [stats.mean.feature(), stats.std.feature(), stats.var.feature(), stats.dmin.feature(), stats.dmax.feature(),
          stats.skewness.feature(), stats.kurtosis.feature(), timedom.zero_crossing_count.feature(0.0), timedom.line_length.feature()]
    pf = [r.feature() for r in _spectral_reducers()]
    big = engine.window_table(x, 500, 250, sf + pf, fs=FS)
    for rows in (slice(0, 1), slice(5, 8), slice(23, 24)):
        alone = engine.window_table(x[rows].clone(), 500, 250, sf + pf, fs=FS)
        assert torch.equal(alone, big[rows]), rows
    if torch.cuda.device_count() >= 2:                     # ... and on another device of the box
        other = engine.window_table(x[5:8].to("cuda:1"), 500, 250, sf + pf, fs=FS)
        assert torch.equal(other.cpu(), big[5:8].cpu())
