"""CPU checks of the denoise-stage oracle (oracle/oracle_nlm.c; PARITY UNPINNED: skimage / PyWavelets are not in
this image, see that file's header).  What CAN be checked without them: the restated estimator recovers a known
noise level, the db2 high-pass filter annihilates what a db2 detail filter must annihilate, the float32
integral-image evaluation and the float64 gather evaluation of the non-local-means formula agree up to float32
rounding and the cut-off sensitivity, and the algorithm's invariances hold."""
import numpy as np
import pytest

from oracle import kmsr_oracle as orc
from oracle import oracle_c


def _scene(h, w, seed, sigma=0.5, level=80.0):
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    clean = level + 3.0 * np.sin(xx / 9.0) + 2.0 * np.cos(yy / 7.0) + 2.0 * (xx > 0.6 * w)
    return clean.astype(np.float32), (clean + sigma * rs.randn(h, w)).astype(np.float32)


def test_estimate_sigma_recovers_the_noise_level():
    for sigma in (0.05, 0.5, 2.0):
        _, noisy = _scene(256, 256, 3, sigma)
        est = oracle_c.estimate_sigma(noisy)
        assert abs(est - sigma) / sigma < 0.08, (sigma, est)


def test_db2_detail_filter_properties():
    # two vanishing moments: constants and linear ramps have zero diagonal detail away from the borders;
    # the symmetric extension keeps constants at zero on the border too
    const = np.full((32, 48), 7.25, dtype=np.float32)
    s, dd = oracle_c.estimate_sigma(const, return_dd=True)
    assert dd.shape == (17, 25) and np.abs(dd).max() < 1e-5
    yy, xx = np.mgrid[0:40, 0:40]
    ramp = (0.5 * xx + 0.25 * yy).astype(np.float32)
    _, dd = oracle_c.estimate_sigma(ramp, return_dd=True)
    assert np.abs(dd[2:-2, 2:-2]).max() < 1e-5
    # all coefficients exactly zero -> np.median of an empty array -> NaN
    assert np.isnan(oracle_c.estimate_sigma(np.zeros((16, 16), dtype=np.float32)))
    # median convention: mean of the two middle values for an even count
    rs = np.random.RandomState(0)
    img = rs.randn(30, 30).astype(np.float32)
    s, dd = oracle_c.estimate_sigma(img, return_dd=True)
    v = dd[dd != 0]
    assert s == pytest.approx(float(np.median(v)) / 0.6744897501960817, rel=1e-7)


@pytest.mark.parametrize("shape,d", [((48, 40), 11), ((33, 61), 5), ((70, 70), 11)])
def test_integral_image_form_equals_gather_form(shape, d):
    _, noisy = _scene(shape[0], shape[1], 11)
    sigma = oracle_c.estimate_sigma(noisy)
    h = 1.15 * sigma
    fast = oracle_c.nlm_fast_f32(noisy, h, sigma, 7, d)
    exact, flip = oracle_c.nlm_exact_f64(noisy, h, sigma, 7, d, eps=2e-3)
    rng = float(noisy.max() - noisy.min())
    # float32 integral images at radiance level 80: the reference's own rounding noise
    assert (np.abs(fast - exact) <= 2e-4 * rng + flip).all(), float(np.abs(fast - exact).max() / rng)
    # and it denoises
    clean, _ = _scene(shape[0], shape[1], 11)
    assert np.sqrt(np.mean((exact - clean) ** 2)) < 0.6 * np.sqrt(np.mean((noisy - clean) ** 2))


def test_nlm_invariances():
    _, noisy = _scene(40, 44, 5)
    sigma = oracle_c.estimate_sigma(noisy)
    h = 1.8 * sigma
    base, _ = oracle_c.nlm_exact_f64(noisy, h, sigma)
    # transposition and flips commute with the algorithm EXCEPT for the patch window, which is the asymmetric
    # 6 x 6 box [-2, +3]^2 of the integral-image form: transposing keeps it, flipping does not
    tr, _ = oracle_c.nlm_exact_f64(np.ascontiguousarray(noisy.T), h, sigma)
    assert np.allclose(tr.T, base, rtol=0, atol=1e-9)
    fl, _ = oracle_c.nlm_exact_f64(np.ascontiguousarray(noisy[::-1, ::-1]), h, sigma)
    assert not np.allclose(fl[::-1, ::-1], base, rtol=0, atol=1e-9)
    # offsetting the radiance level changes nothing but the level (differences only)
    off, _ = oracle_c.nlm_exact_f64(noisy + np.float32(16.0), h, sigma)
    assert np.allclose(off - 16.0, base, rtol=0, atol=1e-6)
    # patch_distance 0: only the zero shift, output = input
    same, _ = oracle_c.nlm_exact_f64(noisy, h, sigma, 7, 0)
    assert np.array_equal(same, noisy.astype(np.float64))
    # even patch sizes are rounded up
    p6, _ = oracle_c.nlm_exact_f64(noisy, h, sigma, 6, 11)
    assert np.array_equal(p6, base)


def test_call_site_port_nan_handling():
    _, noisy = _scene(36, 36, 9)
    noisy[3:6, 10:14] = np.nan
    den, sigma = orc.denoise_band_float_nlm(noisy, 1.8)
    assert np.array_equal(np.isnan(den), np.isnan(noisy)) and sigma > 0
    filled = np.where(np.isnan(noisy), np.nanmean(noisy), noisy).astype(np.float32)
    assert sigma == oracle_c.estimate_sigma(filled)
    allnan = np.full((20, 20), np.nan, dtype=np.float32)
    out, s = orc.denoise_band_float_nlm(allnan)
    assert out is allnan and s == 0.0
