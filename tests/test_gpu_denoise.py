"""GPU parity tests of the upstream denoise stage (denoise/denoise.py:34-65) through the C ABI.

The checker is oracle/oracle_nlm.c -- PARITY UNPINNED (skimage / PyWavelets are absent from the image; the header of
that file says what is restated).  Criterion, the same two-part rule as for the degrade path (tests/parity_util.py):
  (1) |ours - exact| <= 1e-5 x band range + flip_5e-5     exact = float64 value of the formula, flip = what the hard
      `distance > 5 -> skip` cut-off can change at that pixel when a distance moves by 5e-5 (a weight of e^-5
      appears or disappears: a discontinuity of the algorithm, not an inaccuracy of either evaluation);
  (2) |ours - reference| <= 1e-5 x range + |reference - exact| + flip_2e-3     reference = the integral-image
      algorithm in float32 as skimage runs it for a float32 image (its own rounding noise at radiance level 80 is
      ~1e-4 x range).
sigma: <= 1e-5 relative to the float64 evaluation of the estimator, and <= 1e-5 + |reference - float64| relative to
the float32 evaluation (float32 pywt at radiance level 80 is itself ~6e-5 off on a sigma of 0.05).  The NLM check
is run with the sigma the GPU estimated (already held to 1e-5), so it isolates the NLM arithmetic.
"""
import os

import numpy as np
import pytest
import torch

from oracle import kmsr_oracle as orc
from oracle import oracle_c

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def K():
    from kmsr_b200 import _lib, ops
    from kmsr_b200 import D_build_noise_pool as D
    from kmsr_b200 import denoise

    class NS:
        pass
    ns = NS()
    ns.lib, ns.ops, ns.den, ns.D = _lib, ops, denoise, D
    assert torch.cuda.is_available()
    return ns


def _bands(h, w, seed, kind):
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    if kind == "textured":
        clean = 50.0 + 6.0 * np.sin(xx / 9.0) + 4.0 * np.cos(yy / 7.0) + 5.0 * (xx > 0.6 * w) + 3.0 * (yy > 0.3 * h)
        sig = 0.5
    else:                                   # water: level 80, range ~2
        clean = 80.0 + 0.4 * np.sin(xx / 23.0) + 0.3 * np.cos(yy / 17.0)
        sig = 0.05
    return (clean + sig * rs.randn(h, w)).astype(np.float32)


def _filled(band):
    """denoise.py:42-43 with the fill value correctly rounded: the reference's np.nanmean of a float32 band is a
    float32 pairwise sum (~1e-6 relative noise, 5e-5 absolute at level 80); the CUDA path fills with the float64
    mean rounded once.  The checker gets the same filled band so the comparison isolates the transform / NLM."""
    return np.where(np.isnan(band), np.float32(np.nanmean(band, dtype=np.float64)), band).astype(np.float32)


def _check_sigma(got, band, name=""):
    filled = _filled(band)
    ref, ex = oracle_c.estimate_sigma(filled), oracle_c.estimate_sigma(filled, f64=True)
    assert abs(got - ex) <= TOL * ex, (name, "sigma vs float64", got, ex)
    assert abs(got - ref) <= TOL * ex + abs(ref - ex), (name, "sigma vs reference", got, ref)
    return ex


def _check_band(out, img, h_factor, d=11, name="", sigma=None):
    filled = _filled(img)
    if sigma is None:
        sigma = oracle_c.estimate_sigma(filled)
    else:
        _check_sigma(sigma, img, name)
    h = h_factor * sigma
    exact, flip_t = oracle_c.nlm_exact_f64(filled, h, sigma, 7, d, eps=5e-5)
    _, flip_l = oracle_c.nlm_exact_f64(filled, h, sigma, 7, d, eps=2e-3)
    ref = oracle_c.nlm_fast_f32(filled, h, sigma, 7, d)
    rng = float(np.nanmax(img) - np.nanmin(img))
    valid = ~np.isnan(img)
    assert np.array_equal(np.isnan(out), ~valid), f"{name}: NaN pattern"
    o = out.astype(np.float64)
    e1 = np.where(valid, np.abs(o - exact), 0.0)
    assert (e1 <= TOL * rng + flip_t).all(), (name, "vs exact", float((e1 - flip_t).max() / rng))
    e2 = np.where(valid, np.abs(o - ref), 0.0)
    r_ex = np.abs(ref.astype(np.float64) - exact)
    assert (e2 <= TOL * rng + r_ex + flip_l).all(), (name, "vs reference", float(e2.max() / rng))
    return float(e1.max() / rng), float(e2.max() / rng), sigma


def test_estimate_sigma_matches_the_oracle(K):
    shapes = [(256, 256), (64, 64), (33, 47), (130, 17)]
    for h, w in shapes:
        x = np.stack([_bands(h, w, 10 + c, "textured" if c % 2 == 0 else "water") for c in range(5)])[None]
        x[0, 1, 2:5, 3:9] = np.nan                  # filled with the band's nanmean before the transform
        x[0, 3] = np.nan                            # all-NaN band: sigma 0.0 (denoise.py:40-41)
        x[0, 4] = 0.0                               # every coefficient exactly zero -> median of nothing -> NaN
        got = K.ops.estimate_sigma(torch.from_numpy(x).cuda()).cpu().numpy()[0]
        for c in range(3):
            _check_sigma(got[c], x[0, c], name=f"{h}x{w} band {c}")
        assert got[3] == 0.0 and np.isnan(got[4])


@pytest.mark.parametrize("shape,kind,hf", [((64, 64), "textured", 1.15), ((96, 80), "water", 1.8),
                                           ((70, 131), "textured", 1.0), ((256, 256), "water", 1.0)])
def test_nlm_matches_the_oracle(K, shape, kind, hf):
    img = _bands(shape[0], shape[1], 21, kind)
    out, sigma = K.ops.denoise_nlm(torch.from_numpy(img[None, None]).cuda(), hf)
    e1, e2, s = _check_band(out.cpu().numpy()[0, 0], img, hf, name=f"{shape} {kind}", sigma=float(sigma[0, 0]))
    print(f"nlm {shape} {kind}: |ours-exact| {e1:.2e} x range, |ours-ref| {e2:.2e} x range, sigma {s:.4f}")


def test_nlm_batch_nan_and_small_distance(K):
    x = np.stack([np.stack([_bands(72, 72, 30 + 5 * n + c, "textured" if (n + c) % 2 else "water") for c in range(5)])
                  for n in range(2)])
    x[0, 2, 10:14, 20:31] = np.nan
    x[1, 4] = np.nan
    t = torch.from_numpy(x).cuda()
    out, sigma = K.ops.denoise_nlm(t, 1.8)
    out = out.cpu().numpy()
    for n in range(2):
        for c in range(5):
            if n == 1 and c == 4:
                assert np.isnan(out[n, c]).all() and float(sigma[n, c]) == 0.0
                continue
            _check_band(out[n, c], x[n, c], 1.8, name=f"batch {n},{c}", sigma=float(sigma[n, c]))
    # non-contiguous patch stride (a view into a larger batch) and patch_distance 5
    big = torch.zeros((3, 5, 72, 72), device="cuda")
    big[::2] = t
    out5, s5 = K.ops.denoise_nlm(big[::2], 1.0, 7, 5)
    _check_band(out5.cpu().numpy()[0, 0], x[0, 0], 1.0, d=5, name="d=5 strided", sigma=float(s5[0, 0]))
    # patch_distance 0: only the zero shift -> the (NaN-filled, NaN-restored) input
    out0, _ = K.ops.denoise_nlm(t, 1.0, 7, 0)
    assert np.array_equal(out0.cpu().numpy(), x, equal_nan=True)


def test_dropin_signatures(K, tmp_path, capsys):
    img = _bands(64, 80, 7, "textured")
    img[5, 5] = np.nan
    den, sigma = K.den.denoise_band_float_nlm(img, h_factor=1.15, verbose=True)
    assert "Sigma" in capsys.readouterr().out
    rden, rsigma = orc.denoise_band_float_nlm(img, 1.15)
    assert isinstance(den, np.ndarray) and den.dtype == np.float32 and den.shape == img.shape
    assert isinstance(sigma, float) and sigma == pytest.approx(rsigma, rel=1e-4)
    _check_band(den, img, 1.15, name="drop-in", sigma=sigma)
    allnan = np.full((16, 16), np.nan, dtype=np.float32)
    out, s = K.den.denoise_band_float_nlm(allnan, verbose=False)
    assert out is allnan and s == 0.0
    # folder-level contract over the .npz group container: zeros are NaN (denoise.py:31), a `denoised` group and the
    # per-band attributes appear in <stem>_denoised, failures are reported, not raised
    from kmsr_b200 import BAND_NAMES, patch_io
    geo = np.stack([_bands(64, 64, 40 + c, "textured") for c in range(5)])
    geo[0, :4, :4] = 0.0
    src = str(tmp_path / "goci_000_001.npz")
    patch_io.write_groups(src, {"geophysical_data": {b: geo[c] for c, b in enumerate(BAND_NAMES)},
                                "navigation_data": {"latitude": np.zeros((64, 64)), "longitude": np.zeros((64, 64))}})
    ok, path, err = K.den.process_nc_file(src, str(tmp_path / "out"), h_factor=1.8, verbose=False)
    assert ok and err is None and path.endswith("goci_000_001_denoised.npz")
    den5 = patch_io.read_group_bands(path, "denoised")
    assert np.isnan(den5[0, :4, :4]).all() and np.isfinite(den5[1:]).all()
    z = np.load(path)
    assert float(z["__attrs__/denoised/h_factor"]) == 1.8 and "__attrs__/denoised/L_TOA_865_sigma" in z.files
    assert np.array_equal(patch_io.read_group_bands(path, "geophysical_data"), geo)
    ok, path, err = K.den.process_nc_file(str(tmp_path / "missing.npz"), str(tmp_path / "out"), verbose=False)
    assert not ok and path is None and err.startswith("Error:")
    # batch_denoise.py's loop: two good files and one that cannot be read
    indir = tmp_path / "goci"
    indir.mkdir()
    for j in range(2):
        patch_io.write_groups(str(indir / f"p_{j}.npz"), {"geophysical_data": {b: geo[c] + j for c, b in enumerate(BAND_NAMES)}})
    (indir / "broken.npz").write_bytes(b"not a zip")
    okc, failed = K.den.batch_denoise(str(indir), h_factor=1.0)
    assert okc == 2 and [f for f, _ in failed] == ["broken.npz"]
    assert sorted(os.listdir(str(tmp_path / "goci_denoised"))) == ["p_0_denoised.npz", "p_1_denoised.npz"]
    # the denoised group feeds D_build_noise_pool's `geophysical_data - denoised` (D:84-88)
    g0 = np.where(geo != 0, geo, np.nan)
    noise = K.ops.crop_sub(torch.from_numpy(g0).cuda(), torch.from_numpy(den5).cuda(), [8], [8], 32).cpu().numpy()[0]
    assert np.array_equal(noise, (g0 - den5)[:, 8:40, 8:40], equal_nan=True)
    assert 0.2 < np.nanstd(noise[1]) < 0.7


def test_denoise_error_codes(K):
    x = torch.zeros((1, 1, 32, 32), device="cuda")
    for ps, pd in ((5, 11), (9, 11), (7, 12), (7, -1)):
        with pytest.raises(K.lib.KmsrError) as ei:
            K.ops.denoise_nlm(x, 1.0, ps, pd)
        assert ei.value.code == K.lib.E_UNSUPPORTED
    with pytest.raises(K.lib.KmsrError) as ei:
        K.ops.denoise_nlm(x, 1.0, out=x)
    assert ei.value.code == K.lib.E_INVALID
    with pytest.raises(ValueError):
        K.ops.denoise_nlm(x[0], 1.0)
    empty, s = K.ops.denoise_nlm(torch.zeros((0, 5, 32, 32), device="cuda"))
    assert empty.shape == (0, 5, 32, 32) and s.shape == (0, 5)
