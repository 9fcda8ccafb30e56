"""GPU parity tests: the CUDA path (through the C ABI of libkmsr.so) against the oracle and against
the golden vectors produced by the real reference (tests/golden/make_golden.py).

Tolerances (BASELINE.json north_star / SURVEY.md 8c):
  * LR pixels: max|d| <= 1e-5 x (band max - band min) of the HR patch band;
  * indices, offsets, gathers, adds, masks: bit exact;
  * statistics: <= 1e-6 relative to fp64 truth.
"""
import hashlib

import numpy as np
import pytest
import torch

from oracle import kmsr_oracle as orc
from oracle import oracle_c

from parity_util import PIX_TOL, check_pixels, exact_degrade

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    import kmsr_b200
    from kmsr_b200 import _lib, ops, rng
    from kmsr_b200 import C_30apply_kernel_to_landsat as C30
    from kmsr_b200 import C_31apply_muti_kernel_to_landsat as C31
    from kmsr_b200 import D_build_noise_pool as D
    from kmsr_b200 import E_make_train_data as E
    from kmsr_b200 import data_mean_std as S
    from kmsr_b200 import A_00_patch_cutter_universal as CUT

    class NS:
        pass
    ns = NS()
    ns.lib, ns.ops, ns.rng, ns.C30, ns.C31, ns.D, ns.E, ns.S, ns.CUT = _lib, ops, rng, C30, C31, D, E, S, CUT
    assert torch.cuda.is_available()
    return ns


def _case_inputs(z, name, synth):
    if f"{name}__img" in z.files:
        img = z[f"{name}__img"]
    else:
        img = synth.make_hr(1, int(z[f"{name}__seed"]), "textured")[0]
        assert hashlib.sha256(img.tobytes()).hexdigest() == str(z[f"{name}__sha256"])
    return img, z[f"{name}__kernel"], int(z[f"{name}__factor"])


@pytest.mark.parametrize("algo", ["tiled", "auto"])
def test_golden_degrade_cases(K, golden, synth, algo):
    """Every reference-generated case (edge shapes, odd dims, 2-D kernels, zero taps, unnormalised and
    non-positive-sum kernels, k=11/21, factors 1/2/4/6/8, 3 bands, water stress case)."""
    z = golden("golden_degrade.npz")
    worst = {}
    for name in z["cases"]:
        img, kern, f = _case_inputs(z, name, synth)
        ref = z[f"{name}__out"]
        k = torch.from_numpy(kern)
        if k.ndim == 2:
            k = k.unsqueeze(0).repeat(img.shape[0], 1, 1)
        out = K.ops.degrade_batch(torch.from_numpy(img).cuda().unsqueeze(0), k.cuda(), factor=f,
                                  algo=algo)[0].cpu().numpy()
        assert out.shape == ref.shape, name
        worst[name] = check_pixels(out, ref, img, exact_degrade(img, kern, f), name=name, detail=True)
    # The plain north-star bar (max |ours - ref| <= 1e-5 x range) is asserted wherever the reference's own MEASURED
    # deviation from the exact value of its formula leaves room for it (< half the bar): no exemption by name.  Where the
    # reference itself sits further than that from exact, |ours - ref| may exceed the bar by at most that deviation
    # (check_pixels asserted it per pixel) and ours must still be within 5e-6 of exact.
    bad = {k: v for k, v in worst.items() if v["pure"] > PIX_TOL and v["ref_vs_exact"] < 0.5 * PIX_TOL}
    assert not bad, (bad, worst)
    noisy = {k: v for k, v in worst.items() if v["ref_vs_exact"] >= 0.5 * PIX_TOL}
    assert all(v["pure"] <= PIX_TOL + v["ref_vs_exact"] and v["ours_vs_exact"] <= 5e-6 for v in noisy.values()), noisy
    # the audit trail: which cases needed the discount, and how many of their pixels are over the plain bar
    print({k: (round(v["pure"] * 1e5, 3), round(v["ref_vs_exact"] * 1e5, 3), v["frac_over_bar"]) for k, v in noisy.items()})


def test_dropin_signatures_match_reference_outputs(K, golden, synth):
    """C_30 / C_31 apply_kernel_degradation with CPU tensors in, CPU tensors out."""
    z = golden("golden_degrade.npz")
    for name in ("p64_k13_s8", "p64_kernel2d", "p256_water", "odd_70x52_s8", "c3_bands"):
        img, kern, f = _case_inputs(z, name, synth)
        for mod in (K.C30, K.C31):
            out = mod.apply_kernel_degradation(torch.from_numpy(img), torch.from_numpy(kern), f)
            assert isinstance(out, torch.Tensor) and not out.is_cuda and out.dtype == torch.float32
            check_pixels(out.numpy(), z[f"{name}__out"], img, exact_degrade(img, kern, f), name=name)
    img, kern, _ = _case_inputs(z, "p64_k13_s8", synth)
    out = K.C30.apply_kernel_degradation(torch.from_numpy(img).cuda(), torch.from_numpy(kern).cuda())
    assert out.is_cuda and out.shape == (5, 8, 8)                       # default factor 8


def test_config1_single_kernel_64_patches(K, synth, bank):
    """BASELINE config 1: kernel_0 on 64 synthetic 256x256 patches, all 64 checked against the oracle."""
    kb, _ = bank
    hr = np.concatenate([synth.make_hr(32, 1235, "textured"), synth.make_hr(32, 1236, "water")])
    lr = K.C30.degrade_patches(torch.from_numpy(hr), torch.from_numpy(kb[0]), 8).numpy()
    assert lr.shape == (64, 5, 32, 32)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ref = np.stack([orc.apply_kernel_degradation(torch.from_numpy(hr[i]), torch.from_numpy(kb[0]), 8).numpy()
                    for i in range(64)])
    pure = [check_pixels(lr[i], ref[i], hr[i], exact_degrade(hr[i], kb[0], 8) if i >= 32 else None, name=f"patch{i}")
            for i in range(64)]
    assert max(pure[:32]) <= PIX_TOL          # textured regime: the pure north-star bar


@pytest.mark.parametrize("algo", ["tiled", "auto"])
def test_config2_multi_kernel_sigma_noise(K, synth, bank, golden, algo):
    """BASELINE config 2 composition on a 192-patch subset: indices exact, pixels within tolerance."""
    kb, sb = bank
    n = 128
    g = golden("golden_rng.npz")
    kidx, nidx = K.rng.draw_multi_kernel_indices(4096, 10, 4096, 42)
    assert np.array_equal(kidx, g["cfg2_kidx"]) and np.array_equal(nidx, g["cfg2_nidx"])
    kidx, nidx = kidx[:n], nidx[:n]
    pool = synth.make_noise_pool(4096, 42)
    hr = np.concatenate([synth.make_hr(n // 2, 1236, "textured"), synth.make_hr(n // 2, 1237, "water")])
    lr = K.ops.degrade_batch(torch.from_numpy(hr).cuda(), torch.from_numpy(kb).cuda(), kidx=kidx,
                             sigma=torch.from_numpy(sb), pool=torch.from_numpy(pool).cuda(), nidx=nidx,
                             factor=8, noise_mode="sigma", algo=algo).cpu().numpy()
    ref = orc.multi_kernel_pairs(hr, kb, sb, pool, kidx, nidx, 8)
    nz = sb[kidx][:, :, None, None].astype(np.float64) * pool[nidx]
    pure = [check_pixels(lr[i], ref[i], hr[i], exact_degrade(hr[i], kb[kidx[i]], 8) if i >= n // 2 else None,
                         noise=nz[i] if i >= n // 2 else None, name=f"patch{i}") for i in range(n)]
    assert max(pure[:n // 2]) <= PIX_TOL      # textured regime: the pure north-star bar
    # noise really is sigma-scaled pool noise: removing it leaves the plain degrade
    plain = K.ops.degrade_batch(torch.from_numpy(hr).cuda(), torch.from_numpy(kb).cuda(), kidx=kidx,
                                factor=8, algo=algo).cpu().numpy()
    assert np.abs((lr - plain) - nz).max() <= 2e-5


def test_e_pair_assembly_matches_reference_chain(K, synth, bank):
    """C_30 -> E.add_noise chain fused in one launch: same indices as E's global stream, scale-1 add."""
    kb, _ = bank
    n = 24
    hr = synth.make_hr(n, 1240, "textured")
    pool = synth.make_noise_pool(512, 42)
    _, lr, nidx = K.E.make_pairs(hr, kb[3], pool, seed=42)
    assert np.array_equal(nidx, orc.draw_noise_indices(n, 512, 42))
    blurred = [orc.apply_kernel_degradation(torch.from_numpy(hr[i]), torch.from_numpy(kb[3]), 8).numpy() for i in range(n)]
    pairs = orc.make_pairs(list(hr), blurred, pool, seed=42)
    ref = np.stack([p[1] for p in pairs])
    for i in range(n):
        check_pixels(lr[i].numpy(), ref[i], hr[i], name=f"pair{i}")       # textured regime: the pure bar


def test_gem_mode_zero_pad_decimate(K, synth, bank):
    """train_gemini.py:124-134 variant: zero padding + [::4] decimation, per-sample kernels."""
    kb, _ = bank
    hr = synth.make_hr(6, 1250, "textured", size=128)
    kidx = np.array([0, 3, 9, 1, 7, 7], dtype=np.int32)
    lr = K.ops.degrade_batch(torch.from_numpy(hr).cuda(), torch.from_numpy(kb).cuda(), kidx=kidx, factor=4,
                             pad_mode="zero", down_mode="decimate").cpu().numpy()
    ref = orc.multi_kernel_pairs(hr, kb, None, None, kidx, None, 4, "zero", "decimate", "none")
    assert lr.shape == ref.shape == (6, 5, 32, 32)
    for i in range(6):
        check_pixels(lr[i], ref[i], hr[i], name=f"gem{i}")


def test_add_noise_bit_exact(K, golden, synth):
    z = golden("golden_rng.npz")
    pool = synth.make_noise_pool(64, int(z["pool64_seed"]))
    np.random.seed(42)
    for b, want in zip(z["add_noise_blurred"], z["add_noise_out"]):
        got = K.E.add_noise(b, pool)                              # consumes the global numpy stream
        assert got.dtype == np.float32 and np.array_equal(got, want)
    # batched gather + sigma scaling equals the fp32 fma of the oracle's C restatement
    idx = np.array([5, 0, 63], dtype=np.int32)
    sig = np.array([[0.75, 0.8, 0.85, 0.9, 0.95]], dtype=np.float32)
    out = K.ops.add_noise_batch(torch.from_numpy(z["add_noise_blurred"]).cuda(), torch.from_numpy(pool).cuda(),
                                idx, sigma=sig).cpu().numpy()
    for i in range(3):
        assert np.array_equal(out[i], oracle_c.add_noise(z["add_noise_blurred"][i], pool[idx[i]], sig[0]))


def test_noise_pool_bit_exact(K, golden):
    z = golden("golden_noise_pool.npz")
    shapes = [tuple(s) for s in z["shapes"]]
    geos = [z[f"geo{i}"] for i in range(len(shapes))]
    dens = [z[f"den{i}"] for i in range(len(shapes))]
    pool, offs = K.D.build_noise_pool_arrays(geos, dens, 2, 32, 42, return_offsets=True)
    assert np.array_equal(offs, z["offsets"])
    assert pool.dtype == np.float32 and np.array_equal(pool, z["pool"])
    import random
    random.seed(42)
    crops = K.D.random_crop(geos[0] - dens[0], 32, 2)
    assert all(np.array_equal(c, w) for c, w in zip(crops, z["pool"][:2]))
    with pytest.raises(ValueError):
        K.D.random_crop(np.zeros((5, 16, 64), np.float32), 32, 1)


def test_band_stats(K, golden, synth, capsys):
    z = golden("golden_stats.npz")
    patches = z["patches"]
    r = K.S.radiance_stats(patches)
    m64, s64, am, asd = orc.radiance_stats_f64(patches)
    assert np.allclose(r["mean"], m64, rtol=1e-6, atol=0) and np.allclose(r["std"], s64, rtol=1e-6, atol=0)
    assert np.allclose(r["avg_mean"], am, rtol=1e-6) and np.allclose(r["avg_std"], asd, rtol=1e-6)
    big = np.concatenate([synth.make_hr(4, 5, "textured"), synth.make_hr(4, 6, "water")])
    big[2, 1, 100:140, 17] = np.nan
    big[5, 4] = np.nan                                           # a fully-NaN band -> NaN mean/std
    r = K.S.radiance_stats(big)
    m64, s64, _, _ = orc.radiance_stats_f64(big)
    ok = ~np.isnan(m64)
    assert np.array_equal(np.isnan(r["mean"]), ~ok) and np.array_equal(np.isnan(r["std"]), ~ok)
    assert np.allclose(r["mean"][ok], m64[ok], rtol=1e-6) and np.allclose(r["std"][ok], s64[ok], rtol=1e-6)


def test_fused_pair_statistics(K, synth, bank):
    """kmsr_degrade_stats_prepared: LR bits identical to the plain call; mean / std of the HR patches
    within 1e-6 of fp64 (data_mean_std.py:32-33), NaN bands through the exact NaN-skipping kernel,
    `sums` accumulated like band_stats; the unfused route (tiled kernel) gives the same contract."""
    kb, _ = bank
    hr = np.concatenate([synth.make_hr(6, 1290, "textured"), synth.make_hr(6, 1291, "water")])
    hr[3, 2, 17, 100] = np.nan
    hr[8, 0, 255, 255] = np.nan
    pool = synth.make_noise_pool(64, 42)
    nidx = K.rng.draw_noise_indices(12, 64, 42)
    hd = torch.from_numpy(hr).cuda()
    m64, s64, _, _ = orc.radiance_stats_f64(hr)
    for algo in ("auto", "tiled"):
        sums = torch.zeros(11, dtype=torch.float64, device="cuda")
        lr, m, s = K.ops.degrade_batch_stats(hd, torch.from_numpy(kb[4]).cuda(), pool=torch.from_numpy(pool).cuda(),
                                             nidx=nidx, factor=8, noise_mode="add", sums=sums, algo=algo)
        assert K.lib.last_algo() == ("tma" if algo == "auto" else "tiled")
        plain = K.ops.degrade_batch(hd, torch.from_numpy(kb[4]).cuda(), pool=torch.from_numpy(pool).cuda(), nidx=nidx,
                                    factor=8, noise_mode="add", algo=algo)
        assert torch.equal(torch.nan_to_num(lr, nan=-1.0), torch.nan_to_num(plain, nan=-1.0))
        m, s = m.cpu().numpy(), s.cpu().numpy()
        assert np.isfinite(m).all() and np.isfinite(s).all()
        assert np.abs(m - m64).max() <= 1e-6 * np.abs(m64).max() and (np.abs(m - m64) <= 1e-6 * np.abs(m64)).all()
        assert (np.abs(s - s64) <= 1e-6 * s64).all(), float((np.abs(s - s64) / s64).max())
        sm = sums.cpu().numpy()
        assert np.allclose(sm[:5], m.sum(axis=0), rtol=1e-12) and np.allclose(sm[5:10], s.sum(axis=0), rtol=1e-12) and sm[10] == 12


def test_analyze_radiance_stats_printout(K, golden, tmp_path, capsys):
    """The drop-in prints the same table numbers as the reference did on the same files (S:48-62)."""
    import re
    z = golden("golden_stats.npz")
    for i, p in enumerate(z["patches"]):
        np.save(tmp_path / f"patch_{i:03d}.npy", p)
    assert K.S.analyze_radiance_stats(str(tmp_path), int(z["num_samples"])) is None
    got = capsys.readouterr().out
    pat = r"Band (\d)\s*\|\s*([-\d.eE+nan]+)\s*\|\s*([-\d.eE+nan]+)"
    want_rows = re.findall(pat, str(z["stdout"]))
    got_rows = re.findall(pat, got)
    assert len(got_rows) == len(want_rows) == 5
    for a, b in zip(got_rows, want_rows):
        assert abs(float(a[1]) - float(b[1])) <= 1e-4 * abs(float(b[1])) + 1e-6
        assert abs(float(a[2]) - float(b[2])) <= 1e-4 * abs(float(b[2])) + 1e-6


def test_water_mask_and_tiling(K, golden, synth):
    z = golden("golden_cutter.npz")
    scene = synth.make_scene(int(z["scene_seed"]), *z["scene_shape"][1:], n_fill=3, n_cloud=3)
    data = scene.copy()
    masked = K.CUT.apply_water_mask(data, K.CUT.THRESHOLD_MIN, K.CUT.THRESHOLD_MAX)
    assert np.array_equal(np.packbits(np.isnan(masked)), z["masked_nan"])
    assert np.array_equal(np.packbits(np.isnan(data)), z["inplace_nan"])
    assert hashlib.sha256(np.nan_to_num(masked, nan=-1.0).tobytes()).hexdigest() == str(z["masked_sha256"])
    total, kept, ij, offsets, dev_scene = K.CUT.create_patches(masked, 256, 0.5, 0.0)
    assert total == int(z["total"]) and kept == int(z["kept"])
    assert np.array_equal(ij, z["kept_ij"][:, :2])
    # kept windows, read back through their offsets, are the reference's patches byte for byte
    w = masked.shape[2]
    flat = dev_scene.reshape(5, -1)
    for (i, j), off, h16 in zip(ij, offsets.tolist(), z["kept_patch_sha16"]):
        assert off == i * 128 * w + j * 128
        p = dev_scene[:, i * 128:i * 128 + 256, j * 128:j * 128 + 256].contiguous().cpu().numpy()
        assert hashlib.sha256(p.tobytes()).hexdigest()[:16] == str(h16)
    # general thresholds and a stride that does not divide the patch (direct kernel)
    keep, cnt = K.ops.keep_mask(torch.from_numpy(masked).cuda(), 96, 40, 0.02)
    ref = orc.keep_mask(masked, 96, 40 / 96, 0.02)
    assert np.array_equal(keep.cpu().numpy(), ref)
    keep, cnt = K.ops.keep_mask(torch.from_numpy(masked).cuda(), 128, 64, 0.3)
    assert np.array_equal(keep.cpu().numpy(), orc.keep_mask(masked, 128, 0.5, 0.3))


def test_scene_windows_degrade_like_cut_patches(K, golden, synth, bank):
    """Config 4 path: degrading kept windows in place (patch_offsets) == degrading the cut patches;
    the replicate halo clamps to the WINDOW, not to the scene."""
    kb, _ = bank
    z = golden("golden_cutter.npz")
    scene = synth.make_scene(int(z["scene_seed"]), *z["scene_shape"][1:], n_fill=3, n_cloud=3)
    masked = orc.apply_water_mask(scene.copy(), 1e-6, 7.0)
    total, kept, ij, offsets, dev_scene = K.CUT.create_patches(masked, 256, 0.5, 0.0)
    assert kept > 0
    h, w = masked.shape[1:]
    for algo, extents in (("tiled", None), ("auto", None), ("auto", (h, w)), ("tma", (h, w))):
        lr = K.ops.degrade_batch(dev_scene, torch.from_numpy(kb[2]).cuda(), factor=8, patch_offsets=offsets,
                                 patch_hw=(256, 256), strides=(h * w, w), algo=algo, scene_hw=extents,
                                 x_multiple=128).cpu().numpy()
        # with the scene extents known the windows stream through the TMA kernel
        assert K.lib.last_algo() == ("tma" if extents else "tiled")
        pick = list(range(min(6, kept))) + list(range(max(kept - 6, 0), kept))      # first and last windows (scene edges)
        for n in pick:
            i, j = ij[n]
            p = masked[:, i * 128:i * 128 + 256, j * 128:j * 128 + 256]
            ref = orc.apply_kernel_degradation(torch.from_numpy(np.ascontiguousarray(p)), torch.from_numpy(kb[2]), 8).numpy()
            check_pixels(lr[n], ref, p, exact_degrade(np.ascontiguousarray(p), kb[2], 8), name=f"window{n}")


@pytest.mark.parametrize("k,s,p", [(11, 8, 64), (13, 4, 128), (15, 2, 64), (21, 8, 256), (31, 4, 128), (31, 8, 512),
                                   (13, 8, 128), (21, 2, 128), (31, 2, 64), (11, 4, 512), (15, 8, 256), (13, 2, 256)])
def test_generic_streaming_kernel(K, synth, k, s, p):
    """degrade_stream<K, S> (BASELINE config 5 shapes): replicate and zero padding, noise epilogue, NaN footprint,
    against the reference call sites; bit-identical to itself under sharding."""
    n = 3
    kern = synth.softmax_kernels(k, 7 + k)
    hr = np.concatenate([synth.make_hr(2, 2000 + k + s, "textured", size=p), synth.make_hr(1, 2100 + k, "water", size=p)])
    hd = torch.from_numpy(hr).cuda()
    kd = torch.from_numpy(kern).cuda()
    lr = K.ops.degrade_batch(hd, kd, factor=s, algo="stream").cpu().numpy()
    assert K.lib.last_algo() == "stream"
    for i in range(n):
        ref = orc.apply_kernel_degradation(torch.from_numpy(hr[i]), torch.from_numpy(kern), s).numpy()
        check_pixels(lr[i], ref, hr[i], exact_degrade(hr[i], kern, s), name=f"stream k{k} s{s} p{p} #{i}")
    # zero padding (train_gemini.py:128) + sigma noise
    ho = p // s
    pool = (np.random.RandomState(3).standard_normal((5, 5, ho, ho)) * 0.5).astype(np.float32)
    sig = np.linspace(0.7, 1.0, 5, dtype=np.float32)[None]
    nidx = np.array([4, 0, 2], dtype=np.int32)
    lz = K.ops.degrade_batch(hd, kd, factor=s, pad_mode="zero", sigma=torch.from_numpy(sig), pool=torch.from_numpy(pool).cuda(),
                             nidx=nidx, noise_mode="sigma", algo="stream").cpu().numpy()
    lt = K.ops.degrade_batch(hd, kd, factor=s, pad_mode="zero", sigma=torch.from_numpy(sig), pool=torch.from_numpy(pool).cuda(),
                             nidx=nidx, noise_mode="sigma", algo="tiled").cpu().numpy()
    # zero padding + box mean is not a reference combination: held to the exact (fp64) value and to the tiled kernel
    for i in range(n):
        rngs = orc.band_range(hr[i])
        ex = exact_degrade(hr[i], kern, s, zero_pad=True) + sig[0][:, None, None].astype(np.float64) * pool[nidx[i]]
        slack = np.spacing(np.abs(ex).astype(np.float32)).astype(np.float64) / rngs
        # with zeros mixed into a level-80 / range-2 patch the border sums live at the radiance level, not at the
        # patch's dynamic range: the water patch gets the level-scaled bar there
        tol = 5e-6 if i < 2 else 5e-6 * float(np.abs(hr[i]).max()) / float(rngs.min())
        assert (np.abs(lz[i] - ex) / rngs <= tol + slack).all(), (k, s, p, i, float((np.abs(lz[i] - ex) / rngs).max()))
        assert (np.abs(lz[i].astype(np.float64) - lt[i]) / rngs <= 2e-4 + slack).all()      # sanity against the other kernel
    # NaN footprint
    hn = hr.copy()
    hn[0, 1, 0, 0] = np.nan
    hn[1, 3, p // 2, p // 3] = np.nan
    hn[2, 0, p - 1, p - 1] = np.nan
    ln = K.ops.degrade_batch(torch.from_numpy(hn).cuda(), kd, factor=s, algo="stream").cpu().numpy()
    for i in range(n):
        ref = orc.apply_kernel_degradation(torch.from_numpy(hn[i]), torch.from_numpy(kern), s).numpy()
        assert np.array_equal(np.isnan(ln[i]), np.isnan(ref)), (k, s, p, i)


@pytest.mark.parametrize("k,s,h,w", [(11, 2, 64, 64), (13, 2, 128, 128), (15, 4, 256, 256), (21, 2, 256, 256), (31, 4, 128, 128),
                                     (31, 2, 64, 64), (21, 4, 512, 512), (11, 4, 64, 64), (13, 2, 72, 200), (15, 2, 130, 66),
                                     (31, 2, 40, 300), (13, 4, 100, 36)])
def test_register_tile_kernel(K, synth, k, s, h, w):
    """degrade_reg<K, S> (the FP32-bound shapes of BASELINE config 5; any H / W, partial tiles, narrow groups):
    replicate and zero padding, noise epilogue, NaN footprint, strided views, against the reference call sites."""
    _check_stencil_kernel(K, synth, "reg", k, s, h, w)


@pytest.mark.parametrize("k,s,h,w", [(11, 2, 64, 64), (13, 2, 128, 128), (15, 4, 256, 256), (21, 2, 256, 256), (31, 4, 128, 128),
                                     (31, 2, 64, 64), (21, 4, 512, 512), (11, 4, 64, 64), (13, 2, 72, 208), (15, 2, 136, 80),
                                     (31, 2, 40, 304), (13, 4, 104, 48), (13, 8, 64, 64), (31, 8, 64, 64), (15, 8, 128, 128),
                                     (21, 8, 64, 64), (11, 8, 56, 48), (13, 4, 64, 64), (15, 4, 64, 64), (21, 4, 64, 64)])
def test_box_tile_kernel(K, synth, k, s, h, w):
    """degrade_box<K, S> (TMA box tiles: factor 2 / 4 sweep shapes, 64-wide patches at any factor; partial tiles,
    whole-band-per-warp mode, strided views): same checks as the register-tile kernel."""
    _check_stencil_kernel(K, synth, "box", k, s, h, w)


def test_box_tile_kernel_multi_kernel_noise(K, synth):
    """kidx / nidx / sigma through the box kernel (tile mode and band mode) against the tiled kernel and the oracle."""
    for (k, s, p, n) in ((13, 4, 128, 7), (11, 2, 64, 9), (21, 2, 256, 3)):
        bank = np.stack([synth.softmax_kernels(k, 50 + i) for i in range(4)])
        sig = np.linspace(0.5, 1.2, 4 * 5, dtype=np.float32).reshape(4, 5)
        hr = synth.make_hr(n, 4100 + k, "textured", size=p)
        ho = p // s
        pool = (np.random.RandomState(11).standard_normal((6, 5, ho, ho)) * 0.5).astype(np.float32)
        kidx = np.random.RandomState(12).randint(0, 4, n).astype(np.int32)
        nidx = np.random.RandomState(13).randint(0, 6, n).astype(np.int32)
        hd, kd = torch.from_numpy(hr).cuda(), torch.from_numpy(bank).cuda()
        lb = K.ops.degrade_batch(hd, kd, kidx=kidx, sigma=torch.from_numpy(sig), pool=torch.from_numpy(pool).cuda(),
                                 nidx=nidx, factor=s, noise_mode="sigma", algo="box").cpu().numpy()
        assert K.lib.last_algo() == "box"
        for i in range(n):
            ref = orc.apply_kernel_degradation(torch.from_numpy(hr[i]), torch.from_numpy(bank[kidx[i]]), s).numpy()
            nz = sig[kidx[i]][:, None, None].astype(np.float64) * pool[nidx[i]]
            ref = (ref + (sig[kidx[i]][:, None, None] * pool[nidx[i]]).astype(np.float32)).astype(np.float32)
            check_pixels(lb[i], ref, hr[i], exact_degrade(hr[i], bank[kidx[i]], s), noise=nz, name=f"box multi k{k} s{s} #{i}")


def _check_stencil_kernel(K, synth, algo, k, s, h, w):
    n = 3
    kern = synth.softmax_kernels(k, 7 + k)
    p = max(h, w)
    p = (p + 63) // 64 * 64
    full = np.concatenate([synth.make_hr(2, 3000 + k + s, "textured", size=p), synth.make_hr(1, 3100 + k, "water", size=p)])
    hr = np.ascontiguousarray(full[:, :, :h, :w])
    hd = torch.from_numpy(full).cuda()[:, :, :h, :w]           # a strided view: rows contiguous, row stride p
    kd = torch.from_numpy(kern).cuda()
    lr = K.ops.degrade_batch(hd, kd, factor=s, algo=algo).cpu().numpy()
    assert K.lib.last_algo() == algo and lr.shape == (n, 5, h // s, w // s)
    for i in range(n):
        ref = orc.apply_kernel_degradation(torch.from_numpy(hr[i]), torch.from_numpy(kern), s).numpy()
        check_pixels(lr[i], ref, hr[i], exact_degrade(hr[i], kern, s), name=f"{algo} k{k} s{s} {h}x{w} #{i}")
    # zero padding (train_gemini.py:128) + sigma noise: held to the exact (fp64) value
    ho, wo = h // s, w // s
    pool = (np.random.RandomState(3).standard_normal((5, 5, ho, wo)) * 0.5).astype(np.float32)
    sig = np.linspace(0.7, 1.0, 5, dtype=np.float32)[None]
    nidx = np.array([4, 0, 2], dtype=np.int32)
    lz = K.ops.degrade_batch(hd, kd, factor=s, pad_mode="zero", sigma=torch.from_numpy(sig), pool=torch.from_numpy(pool).cuda(),
                             nidx=nidx, noise_mode="sigma", algo=algo).cpu().numpy()
    for i in range(n):
        rngs = orc.band_range(hr[i])
        ex = exact_degrade(hr[i], kern, s, zero_pad=True) + sig[0][:, None, None].astype(np.float64) * pool[nidx[i]]
        slack = np.spacing(np.abs(ex).astype(np.float32)).astype(np.float64) / rngs
        tol = 5e-6 if i < 2 else 5e-6 * float(np.abs(hr[i]).max()) / float(rngs.min())
        assert (np.abs(lz[i] - ex) / rngs <= tol + slack).all(), (k, s, h, w, i, float((np.abs(lz[i] - ex) / rngs).max()))
    # NaN footprint
    hn = hr.copy()
    hn[0, 1, 0, 0] = np.nan
    hn[1, 3, h // 2, w // 3] = np.nan
    hn[2, 0, h - 1, w - 1] = np.nan
    ln = K.ops.degrade_batch(torch.from_numpy(hn).cuda(), kd, factor=s, algo=algo).cpu().numpy()
    for i in range(n):
        ref = orc.apply_kernel_degradation(torch.from_numpy(hn[i]), torch.from_numpy(kern), s).numpy()
        assert np.array_equal(np.isnan(ln[i]), np.isnan(ref)), (k, s, h, w, i)


def test_streaming_interior_path_equals_general_path(K, synth, monkeypatch):
    """The interior fast path of degrade_stream (one chunk in, one chunk out per step) must give the bits of the general
    path (KMSR_STREAM_NOFAST=1) on every <K, S, TX> variant, run to run -- the check that caught a chunk being handed
    back to the producer before the warp's queued shared-memory loads had read it -- and both must agree with the
    tiled kernel."""
    torch.manual_seed(5)
    for k in (11, 13, 15, 21, 31):
        kd = torch.from_numpy(synth.softmax_kernels(k, 7)).cuda()
        for s in (4, 8):
            for (h, w) in ((64, 64), (128, 128), (96, 512), (256, 1024)):
                n = 48 if h * w <= 128 * 128 else 16
                hr = torch.randn((n, 5, h, w), device="cuda") * 3 + 50
                for pad in ("replicate", "zero"):
                    monkeypatch.setenv("KMSR_STREAM_NOFAST", "1")
                    g = K.ops.degrade_batch(hr, kd, factor=s, pad_mode=pad, algo="stream")
                    monkeypatch.setenv("KMSR_STREAM_NOFAST", "0")
                    for _ in range(3):
                        f = K.ops.degrade_batch(hr, kd, factor=s, pad_mode=pad, algo="stream")
                        assert torch.equal(f, g), (k, s, h, w, pad)
                    t = K.ops.degrade_batch(hr, kd, factor=s, pad_mode=pad, algo="tiled")
                    assert float((f - t).abs().max()) <= 1e-4 * 25.0, (k, s, h, w, pad)


def test_nan_propagation_matches_reference(K, synth, bank):
    """A NaN pixel poisons exactly the LR pixels whose (clamped) window contains it."""
    kb, _ = bank
    hr = synth.make_hr(2, 1260, "textured")
    hr[0, 1, 0, 0] = np.nan            # corner: replicated into the halo
    hr[0, 3, 130, 77] = np.nan
    hr[1, 0, 255, 200] = np.nan
    for algo in ("tiled", "auto"):
        lr = K.ops.degrade_batch(torch.from_numpy(hr).cuda(), torch.from_numpy(kb[0]).cuda(), factor=8,
                                 algo=algo).cpu().numpy()
        for i in range(2):
            ref = orc.apply_kernel_degradation(torch.from_numpy(hr[i]), torch.from_numpy(kb[0]), 8).numpy()
            assert np.array_equal(np.isnan(lr[i]), np.isnan(ref)), algo
            ok = ~np.isnan(ref)
            rng_ = np.broadcast_to(orc.band_range(hr[i]), ref.shape)
            assert (np.abs(lr[i][ok] - ref[ok]) / rng_[ok]).max() <= PIX_TOL


def test_full_size_properties(K, synth, bank):
    """BASELINE config 2 size (4096 patches) through size-independent properties:
    shard invariance (bitwise), constant patches stay constant, additivity of the noise term."""
    kb, sb = bank
    n = 4096
    base = torch.from_numpy(synth.make_hr(64, 1270, "textured")).cuda()
    hr = base.repeat(n // 64, 1, 1, 1)
    hr += torch.arange(n, device="cuda", dtype=torch.float32).view(n, 1, 1, 1) * 0.01   # distinct patches
    pool = torch.from_numpy(synth.make_noise_pool(4096, 42)).cuda()
    kidx, nidx = K.rng.draw_multi_kernel_indices(n, 10, 4096, 42)
    kbd = torch.from_numpy(kb).cuda()
    full = K.ops.degrade_batch(hr, kbd, kidx=kidx, sigma=torch.from_numpy(sb), pool=pool, nidx=nidx, factor=8)
    assert full.shape == (n, 5, 32, 32) and bool(torch.isfinite(full).all())
    for world in (2, 8):
        parts = []
        for r in range(world):
            a, b = K.rng.shard_range(n, r, world)
            parts.append(K.ops.degrade_batch(hr[a:b], kbd, kidx=kidx[a:b], sigma=torch.from_numpy(sb), pool=pool,
                                             nidx=nidx[a:b], factor=8))
        assert torch.equal(torch.cat(parts), full), f"shards differ at world={world}"
    # the tiled and the TMA kernels agree to well inside the tolerance on every patch
    tiled = K.ops.degrade_batch(hr, kbd, kidx=kidx, sigma=torch.from_numpy(sb), pool=pool, nidx=nidx, factor=8,
                                algo="tiled")
    rngs = (hr.amax(dim=(2, 3)) - hr.amin(dim=(2, 3)))[:, :, None, None]
    assert float(((tiled - full).abs() / rngs).max()) <= PIX_TOL
    # constant patch -> constant output (normalised kernels sum to one)
    const = torch.full((16, 5, 256, 256), 37.25, device="cuda")
    out = K.ops.degrade_batch(const, kbd, kidx=np.arange(16, dtype=np.int32) % 10, factor=8)
    assert float((out - 37.25).abs().max()) <= 37.25 * 4e-7
    # checksum of checksums: per-band LR mean equals HR band mean up to boundary weighting
    plain = K.ops.degrade_batch(hr[:256], kbd, kidx=kidx[:256], factor=8)
    assert float((plain.mean(dim=(2, 3)) - hr[:256].mean(dim=(2, 3))).abs().max()) < 0.5


def test_c_abi_error_codes(K):
    """Bad arguments come back as negative codes with a message; nothing throws inside the library."""
    import ctypes as C
    lib = K.lib.lib()
    x = torch.zeros(1, 5, 64, 64, device="cuda")
    comp = torch.zeros(1, 5, 20, 20, device="cuda")
    ds = torch.zeros(1, 5, device="cuda")
    out = torch.zeros(1, 5, 8, 8, device="cuda")
    vp = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.kmsr_degrade_prepared(vp(x), 1, 5, 64, 64, 5 * 4096, 4096, 64, None, vp(comp), vp(ds), 1, 13, 13,
                                   None, None, None, 0, None, 8, 7, 0, 0, vp(out), 0, None)
    assert rc == K.lib.E_INVALID and "pad_mode" in K.lib.last_error()
    rc = lib.kmsr_degrade_prepared(vp(x), 1, 5, 64, 64, 5 * 4096, 4096, 64, None, vp(comp), vp(ds), 1, 13, 13,
                                   None, None, None, 0, None, 8, 0, 0, 1, vp(out), 0, None)
    assert rc == K.lib.E_INVALID and "noise" in K.lib.last_error()
    rc = lib.kmsr_degrade_prepared(vp(x), 1, 5, 64, 64, 5 * 4096, 4096, 64, None, vp(comp), vp(ds), 1, 13, 13,
                                   None, None, None, 0, None, 8, 0, 0, 0, vp(out), K.lib.ALGO_TMA, None)
    assert rc == K.lib.E_UNSUPPORTED
    rc = lib.kmsr_crop_sub(vp(x), vp(x), 5, 16, 64, None, None, 1, 32, vp(out), None)
    assert rc == K.lib.E_INVALID and "smaller than crop" in K.lib.last_error()
    with pytest.raises(K.lib.KmsrError):
        K.lib.check(rc)
    # empty batch is a no-op success
    assert lib.kmsr_degrade_prepared(None, 0, 5, 64, 64, 0, 0, 64, None, None, None, 1, 13, 13, None, None, None,
                                     0, None, 8, 0, 0, 0, None, 0, None) == 0


def test_folder_drivers_match_the_reference_chain(K, synth, bank, tmp_path, capsys):
    """C_30.process_landsat_folder -> E.process_files and D.build_noise_pool on .npz patch files: file naming, per-file
    skip semantics, RNG consumption and pixels equal to the reference functions applied file by file."""
    import os
    import random
    from kmsr_b200 import patch_io as pio
    kb, _ = bank
    src, blur_dir, train_dir, goci = [str(tmp_path / d) for d in ("patches", "blurred", "train", "goci")]
    for d in (src, goci):
        os.makedirs(d)
    hr = synth.make_hr(5, 3100, "textured")
    rs = np.random.RandomState(9)
    nav = {"latitude": rs.standard_normal((256, 256)).astype(np.float32), "longitude": rs.standard_normal((256, 256)).astype(np.float32)}
    for i in range(5):
        pio.write_groups(os.path.join(src, f"LC08_{i:02d}_denoised.npz"),
                         {"denoised": {b: hr[i, c] for c, b in enumerate(pio.BAND_NAMES)}, "navigation_data": nav})
    pio.write_groups(os.path.join(src, "LC08_97_denoised.npz"), {"geophysical_data": {"L_TOA_443": hr[0, 0]}})   # no 'denoised' group
    pio.write_groups(os.path.join(src, "LC08_98_denoised.npz"),                                                 # wrong size for E
                     {"denoised": {b: hr[0, c, :128, :128] for c, b in enumerate(pio.BAND_NAMES)}, "navigation_data": nav})
    kpath = str(tmp_path / "kernel.npy")
    np.save(kpath, kb[5])
    K.C30.process_landsat_folder(src, kpath, blur_dir)
    out = sorted(os.listdir(blur_dir))
    assert out == [f"LC08_{i:02d}_denoised_blurred.npz" for i in range(5)] + ["LC08_98_denoised_blurred.npz"]
    for i in range(5):
        got = pio.read_group_bands(os.path.join(blur_dir, out[i]), "blurred")
        ref = orc.apply_kernel_degradation(torch.from_numpy(hr[i]), torch.from_numpy(kb[5]), 8).numpy()
        check_pixels(got, ref, hr[i], name=out[i])
        assert np.array_equal(pio.read_group_bands(os.path.join(blur_dir, out[i]), "denoised"), hr[i])
    # E: listdir order binds the draws; the 128x128 file is rejected without a draw
    pool = synth.make_noise_pool(32, 42)
    ppath = str(tmp_path / "pool.npy")
    np.save(ppath, pool)
    ok, fail = K.E.process_files(blur_dir, ppath, train_dir, seed=42)
    assert (ok, fail) == (5, 1)
    order = [f for f in os.listdir(blur_dir) if f.endswith(".npz")]
    np.random.seed(42)
    for f in order:
        if "_98_" in f:
            continue
        idx = np.random.randint(0, len(pool))
        tr = os.path.join(train_dir, f.replace("_denoised_blurred.npz", "_train.npz"))
        assert os.path.isfile(tr)
        blurred = pio.read_group_bands(os.path.join(blur_dir, f), "blurred")
        assert np.array_equal(pio.read_group_bands(tr, "lr"), blurred + pool[idx])          # E:74, bit exact
        assert np.array_equal(pio.read_group_bands(tr, "hr"), pio.read_group_bands(os.path.join(blur_dir, f), "denoised"))
        assert set(pio.read_navigation(tr)) == {"latitude", "longitude"}
    # C_31: group 'hr' in, group 'lr' written into the same file
    c31 = str(tmp_path / "c31")
    os.makedirs(c31)
    for i in range(2):
        pio.write_groups(os.path.join(c31, f"p{i}.npz"), {"hr": {b: hr[i, c] for c, b in enumerate(pio.BAND_NAMES)}})
    K.C31.process_landsat_folder(c31, kpath, str(tmp_path / "c31_out"), downscale_factor=4)
    for i in range(2):
        got = pio.read_group_bands(os.path.join(c31, f"p{i}.npz"), "lr")
        ref = orc.apply_kernel_degradation(torch.from_numpy(hr[i]), torch.from_numpy(kb[5]), 4).numpy()
        assert got.shape == (5, 64, 64)
        check_pixels(got, ref, hr[i], name=f"c31 p{i}")
    # D: noise = geo - den, crops at CPython-random offsets in listdir order
    geo = synth.make_hr(3, 3200, "textured", size=96)
    den = geo - (rs.standard_normal(geo.shape) * 0.3).astype(np.float32)
    for i in range(3):
        pio.write_groups(os.path.join(goci, f"GK2_{i}.npz"), {"geophysical_data": {b: geo[i, c] for c, b in enumerate(pio.BAND_NAMES)},
                                                             "denoised": {b: den[i, c] for c, b in enumerate(pio.BAND_NAMES)}})
    pool2 = K.D.build_noise_pool(goci, str(tmp_path / "np" / "pool.npy"), str(tmp_path / "np" / "meta.npy"), samples_per_file=2,
                                 patch_size=32, seed=42)
    assert np.array_equal(np.load(str(tmp_path / "np" / "pool.npy")), pool2) and pool2.shape == (6, 5, 32, 32)
    random.seed(42)
    m = 0
    for f in [f for f in os.listdir(goci) if f.endswith(".npz")]:
        i = int(f.split("_")[1].split(".")[0])
        noise = geo[i] - den[i]
        for _ in range(2):
            top = random.randint(0, 96 - 32)
            left = random.randint(0, 96 - 32)
            assert np.array_equal(pool2[m], noise[:, top:top + 32, left:left + 32])
            m += 1
    meta = np.load(str(tmp_path / "np" / "meta.npy"), allow_pickle=True)
    assert len(meta) == 6 and meta[0]["patch_size"] == 32
    capsys.readouterr()


def test_create_patches_nc_writes_the_kept_windows(K, synth, tmp_path, capsys):
    """CUT.create_patches_nc: (total, kept), file names in raster order, patch contents and cropped navigation."""
    import os
    from kmsr_b200 import patch_io as pio
    scene = synth.make_scene(11, 640, 768, n_fill=2, n_cloud=2)
    masked = orc.apply_water_mask(scene.copy(), 1e-6, 7.0)
    nav = {"latitude": np.arange(640 * 768, dtype=np.float32).reshape(640, 768)}
    total, kept = K.CUT.create_patches_nc(masked, 256, 0.5, 0.0, str(tmp_path / "cut"), "LC09", {"navigation_data": nav, "source_file": "s.nc"},
                                          ext=".npz")
    ref_keep = orc.keep_mask(masked, 256, 0.5, 0.0)
    assert total == ref_keep.size and kept == int(ref_keep.sum())
    names = sorted(os.listdir(str(tmp_path / "cut")))
    want = [f"LC09_{i:03d}_{j:03d}.npz" for i in range(ref_keep.shape[0]) for j in range(ref_keep.shape[1]) if ref_keep[i, j]]
    assert names == want
    i, j = [int(t) for t in names[-1][5:12].split("_")]
    p = pio.read_group_bands(str(tmp_path / "cut" / names[-1]), "geophysical_data")
    assert np.array_equal(p, masked[:, i * 128:i * 128 + 256, j * 128:j * 128 + 256])
    assert np.array_equal(pio.read_navigation(str(tmp_path / "cut" / names[-1]))["latitude"], nav["latitude"][i * 128:i * 128 + 256, j * 128:j * 128 + 256])
    capsys.readouterr()


def test_per_patch_dynamic_kernels(K, synth):
    """a bank with one kernel per patch ([N,5,13,13], kidx = arange(N)) -- the output format of
    muti_kernel/train.py:118-187 -- goes through the same entry point."""
    n = 24
    hr = synth.make_hr(n, 3300, "textured")
    kb = synth.softmax_kernels(13, 99, n=n)                      # [N, 5, 13, 13]
    lr = K.ops.degrade_batch(torch.from_numpy(hr).cuda(), torch.from_numpy(kb).cuda(), kidx=np.arange(n, dtype=np.int32),
                             factor=8).cpu().numpy()
    assert K.lib.last_algo() == "tma"
    for i in range(0, n, 5):
        ref = orc.apply_kernel_degradation(torch.from_numpy(hr[i]), torch.from_numpy(kb[i]), 8).numpy()
        check_pixels(lr[i], ref, hr[i], name=f"dynamic kernel {i}")


def test_new_entry_points_reject_bad_arguments(K, synth, bank):
    """kmsr_degrade_windows / kmsr_degrade_stats_prepared / KMSR_ALGO_STREAM: argument errors come back as codes."""
    import ctypes as C
    lib = K.lib.lib()
    vp = lambda t: C.c_void_p(t.data_ptr())
    kb, _ = bank
    pb = K.ops.prepare_kernels(torch.from_numpy(kb[0]).cuda(), 8)
    scene = torch.zeros(5, 512, 512, device="cuda")
    offs = torch.zeros(1, dtype=torch.int64, device="cuda")
    out = torch.zeros(1, 5, 32, 32, device="cuda")
    # window larger than the scene
    rc = lib.kmsr_degrade_windows(vp(scene), 5, 128, 512, 128 * 512, 512, vp(offs), 1, 256, 256, 128, vp(pb.comp), vp(pb.dsum),
                                  1, 13, 13, None, None, None, 0, None, 8, 0, 0, 0, vp(out), 0, None)
    assert rc == K.lib.E_INVALID and "does not fit" in K.lib.last_error()
    # the TMA kernel needs the caller's alignment promise; without it AUTO falls back to the tiled kernel
    rc = lib.kmsr_degrade_windows(vp(scene), 5, 512, 512, 512 * 512, 512, vp(offs), 1, 256, 256, 1, vp(pb.comp), vp(pb.dsum),
                                  1, 13, 13, None, None, None, 0, None, 8, 0, 0, 0, vp(out), K.lib.ALGO_TMA, None)
    assert rc == K.lib.E_UNSUPPORTED and "multiples of 4" in K.lib.last_error()
    rc = lib.kmsr_degrade_windows(vp(scene), 5, 512, 512, 512 * 512, 512, vp(offs), 1, 256, 256, 1, vp(pb.comp), vp(pb.dsum),
                                  1, 13, 13, None, None, None, 0, None, 8, 0, 0, 0, vp(out), 0, None)
    assert rc == 0 and K.lib.last_algo() == "tiled"
    # statistics entry point: missing outputs / short workspace
    hr = torch.zeros(2, 5, 256, 256, device="cuda")
    lr = torch.zeros(2, 5, 32, 32, device="cuda")
    m = torch.zeros(2, 5, dtype=torch.float64, device="cuda")
    ws = torch.zeros(16, dtype=torch.uint8, device="cuda")
    rc = lib.kmsr_degrade_stats_prepared(vp(hr), 2, 5, 256, 256, 5 * 65536, vp(pb.comp), vp(pb.dsum), 1, 13, 13, None, None, None, 0,
                                         None, 8, 0, 0, 0, vp(lr), None, vp(m), None, vp(ws), 16, 0, None)
    assert rc == K.lib.E_INVALID
    rc = lib.kmsr_degrade_stats_prepared(vp(hr), 2, 5, 256, 256, 5 * 65536, vp(pb.comp), vp(pb.dsum), 1, 13, 13, None, None, None, 0,
                                         None, 8, 0, 0, 0, vp(lr), vp(m), vp(m), None, vp(ws), 16, 0, None)
    assert rc == K.lib.E_INVALID and "workspace" in K.lib.last_error()
    # the streaming kernel refuses what it does not cover (even kernel, decimation)
    with pytest.raises(K.lib.KmsrError):
        K.ops.degrade_batch(hr, torch.rand(5, 12, 12, device="cuda"), factor=8, algo="stream")
    with pytest.raises(K.lib.KmsrError):
        K.ops.degrade_batch(hr, torch.from_numpy(kb[0]).cuda(), factor=4, down_mode="decimate", pad_mode="zero", algo="stream")
    # ... and AUTO still serves them (tiled)
    K.ops.degrade_batch(hr, torch.rand(5, 12, 12, device="cuda"), factor=8)
    assert K.lib.last_algo() == "tiled"


def test_shapes_the_streaming_kernels_hand_to_each_other(K, synth, bank):
    """AUTO routing: headline shape -> tma, sweep shapes -> stream, odd sizes -> tiled; all three agree."""
    kb, _ = bank
    hr = torch.from_numpy(synth.make_hr(4, 3400, "textured")).cuda()
    kd = torch.from_numpy(kb[7]).cuda()
    a = K.ops.degrade_batch(hr, kd, factor=8)
    assert K.lib.last_algo() == "tma"
    b = K.ops.degrade_batch(hr, kd, factor=8, algo="stream")
    c = K.ops.degrade_batch(hr, kd, factor=8, algo="tiled")
    rngs = (hr.amax(dim=(2, 3)) - hr.amin(dim=(2, 3)))[:, :, None, None]
    assert float(((a - b).abs() / rngs).max()) <= 2e-6 and float(((a - c).abs() / rngs).max()) <= 5e-6
    K.ops.degrade_batch(hr[:, :, :250, :250].contiguous(), kd, factor=8)          # 250 is not a multiple of 8
    assert K.lib.last_algo() == "tiled"
    K.ops.degrade_batch(hr[:, :, :128, :128].contiguous(), kd, factor=4)
    assert K.lib.last_algo() == "box"
    K.ops.degrade_batch(hr[:, :, :128, :128].contiguous(), kd, factor=8)
    assert K.lib.last_algo() == "stream"
    # factor 2 / 4 and 64-wide patches go to the TMA box-tile kernel; what it refuses (H % 8, W % 16) falls to the
    # register-tile kernel (any H / W / strides)
    r2 = K.ops.degrade_batch(hr, kd, factor=2)
    assert K.lib.last_algo() == "box"
    s2 = K.ops.degrade_batch(hr, kd, factor=2, algo="stream")
    assert K.lib.last_algo() == "stream" and float(((r2 - s2).abs() / rngs).max()) <= 2e-6
    g2 = K.ops.degrade_batch(hr, kd, factor=2, algo="reg")
    assert K.lib.last_algo() == "reg" and float(((r2 - g2).abs() / rngs).max()) <= 2e-6
    K.ops.degrade_batch(hr[:, :, :64, :64].contiguous(), kd, factor=4)
    assert K.lib.last_algo() == "box"
    K.ops.degrade_batch(hr[:, :, :64, :64].contiguous(), kd, factor=8)
    assert K.lib.last_algo() == "box"
    K.ops.degrade_batch(hr[:, :, :100, :36].contiguous(), kd, factor=4)
    assert K.lib.last_algo() == "reg"
    # strided views: a channel slice of a wider tensor still streams (16-byte aligned strides)
    wide = torch.randn(4, 7, 256, 256, device="cuda") + 30.0
    v = wide[:, 1:6]
    d = K.ops.degrade_batch(v, kd, factor=8)
    assert K.lib.last_algo() == "tma"
    assert torch.equal(d, K.ops.degrade_batch(v.contiguous(), kd, factor=8))


def test_host_buffer_pipeline_equals_device_path(K, synth, bank):
    """pipeline.PairSynthesizer.run_host (chunked, double-buffered H2D / kernel / D2H over three streams) returns the
    bits of the one-launch device path, also when the batch is not a multiple of the chunk; synthesize_pairs draws
    the same indices as the oracle's composition."""
    from kmsr_b200 import pipeline
    kb, sb = bank
    n = 37
    hr = np.concatenate([synth.make_hr(20, 3500, "textured"), synth.make_hr(17, 3501, "water")])
    pool = synth.make_noise_pool(128, 42)
    kidx, nidx = K.rng.draw_multi_kernel_indices(n, 10, 128, 7)
    syn = pipeline.PairSynthesizer(kb, sb, pool, factor=8, chunk=8)
    dev = syn.run_device(torch.from_numpy(hr).cuda(), kidx, nidx).cpu()
    host_in = torch.from_numpy(hr).pin_memory()
    for _ in range(2):                                   # second pass reuses the staging buffers
        out = syn.run_host(host_in, kidx, nidx)
        torch.cuda.synchronize()
        assert torch.equal(out, dev)
    lr, k2, n2 = pipeline.synthesize_pairs(hr, kb, sb, pool, seed=7, factor=8, chunk=16)
    assert np.array_equal(k2, kidx) and np.array_equal(n2, nidx) and torch.equal(lr, dev)
    ref = orc.multi_kernel_pairs(hr[:4], kb, sb, pool, kidx[:4], nidx[:4], 8)
    for i in range(4):
        check_pixels(dev[i].numpy(), ref[i], hr[i], name=f"pipeline {i}")


def test_raw_scene_keep_grid_and_windows(K, golden, synth, bank):
    """kmsr_scene_keep_mask: the keep grid / NaN counts from the RAW scene equal water_mask + keep_mask (and the
    oracle); at nan_threshold 0 the kept windows degraded straight from the raw scene equal the masked ones bitwise."""
    kb, _ = bank
    scene = synth.make_scene(21, 896, 1152, n_fill=3, n_cloud=3)
    scene[2, 300:310, 400:420] = np.nan                    # a NaN that is not a fill value
    scene[0, 700, 100] = -9999.0                           # fill in one band only
    raw = torch.from_numpy(scene).cuda()
    masked_ref = orc.apply_water_mask(scene.copy(), 1e-6, 7.0)
    for thr in (0.0, 0.01):
        keep, cnt = K.ops.scene_keep_mask(raw, 1e-6, 7.0, 256, 128, thr)
        work = raw.clone()
        masked = K.ops.water_mask(work, 1e-6, 7.0)
        keep2, cnt2 = K.ops.keep_mask(masked, 256, 128, thr)
        assert torch.equal(keep, keep2) and torch.equal(cnt, cnt2)
        assert np.array_equal(keep.cpu().numpy(), orc.keep_mask(masked_ref, 256, 0.5, thr))
    assert torch.equal(raw, torch.from_numpy(scene).cuda()) or bool(torch.isnan(raw).any())     # the raw scene is not modified
    total, kept, ij, offs, dev_scene = K.CUT.create_patches_from_raw(scene)
    t2, k2, ij2, offs2, dev_masked = K.CUT.create_patches(masked_ref, 256, 0.5, 0.0)
    assert (total, kept) == (t2, k2) and np.array_equal(ij, ij2) and kept > 0
    h, w = scene.shape[1:]
    a = K.ops.degrade_batch(dev_scene, torch.from_numpy(kb[1]).cuda(), factor=8, patch_offsets=offs, patch_hw=(256, 256),
                            strides=(h * w, w), scene_hw=(h, w), x_multiple=128)
    b = K.ops.degrade_batch(dev_masked, torch.from_numpy(kb[1]).cuda(), factor=8, patch_offsets=offs2, patch_hw=(256, 256),
                            strides=(h * w, w), scene_hw=(h, w), x_multiple=128)
    assert K.lib.last_algo() == "tma" and torch.equal(a, b) and bool(torch.isfinite(a).all())
    with pytest.raises(K.lib.KmsrError):
        K.ops.scene_keep_mask(raw, 1e-6, 7.0, 96, 40, 0.0)            # P % stride != 0


def test_content_adaptive_pick_feeds_the_fused_kernel(K, golden, synth, bank):
    """f2 on the device: selector logits (libkmsr's 3xTF32 tensor-core convolutions) equal the reference's, the argmax
    pick drives the fused degrade + sigma-noise launch, and the result equals the oracle composition with the same picks."""
    from kmsr_b200.selector import Selector, degrade_content_adaptive
    kb, sb = bank
    z = golden("selector.npz")
    sel = Selector.from_npz(z, "cuda")
    hr = np.concatenate([synth.make_hr(4, 5100, "textured"), synth.make_hr(2, 5101, "water")])
    lg = sel.logits(torch.from_numpy(hr).cuda()).cpu().numpy()
    assert np.abs(lg - z["logits"][:6]).max() <= 1e-4 * np.abs(z["logits"]).max()
    pool = synth.make_noise_pool(64, 42)
    lr, kidx, nidx = degrade_content_adaptive(torch.from_numpy(hr), sel, kb, sb, pool, seed=42)
    assert np.array_equal(kidx, z["argmax"][:6]) and K.lib.last_algo() == "tma"
    _, n_ref = K.rng.draw_multi_kernel_indices(6, 10, 64, 42)
    assert np.array_equal(nidx, n_ref)
    ref = orc.multi_kernel_pairs(hr, kb, sb, pool, kidx, nidx, 8)
    for i in range(6):
        ex = exact_degrade(hr[i], kb[kidx[i]], 8)
        check_pixels(lr[i].numpy(), ref[i], hr[i], ex, noise=sb[kidx[i]][:, None, None].astype(np.float64) * pool[nidx[i]], name=f"adaptive {i}")


def test_wide_bands_through_the_headline_kernel(K, synth, bank):
    """W = 512 / 768: the TMA kernel walks a band in 256-column blocks (interior block edges read real neighbours, only the
    outer ones replicate), noise tile and output offsets follow the block; H up to 512."""
    kb, sb = bank
    rs = np.random.RandomState(77)
    for h, w in ((128, 512), (512, 512), (40, 768)):
        n = 3
        hr = (rs.standard_normal((n, 5, h, w)) * 3.0 + np.array([80, 70, 50, 25, 8])[None, :, None, None]).astype(np.float32)
        pool = (rs.standard_normal((7, 5, h // 8, w // 8)) * 0.5).astype(np.float32)
        kidx = np.array([3, 9, 0], dtype=np.int32)
        nidx = np.array([6, 0, 2], dtype=np.int32)
        lr = K.ops.degrade_batch(torch.from_numpy(hr).cuda(), torch.from_numpy(kb).cuda(), kidx=kidx, sigma=torch.from_numpy(sb),
                                 pool=torch.from_numpy(pool).cuda(), nidx=nidx, factor=8, noise_mode="sigma").cpu().numpy()
        assert K.lib.last_algo() == "tma"
        ref = orc.multi_kernel_pairs(hr, kb, sb, pool, kidx, nidx, 8)
        for i in range(n):
            ex = exact_degrade(hr[i], kb[kidx[i]], 8)
            check_pixels(lr[i], ref[i], hr[i], ex, noise=sb[kidx[i]][:, None, None].astype(np.float64) * pool[nidx[i]],
                         name=f"wide {h}x{w} #{i}")
        # NaN footprint across a block edge
        hn = hr.copy()
        hn[0, 2, h // 2, 255] = np.nan
        hn[1, 0, 0, 256] = np.nan
        ln = K.ops.degrade_batch(torch.from_numpy(hn).cuda(), torch.from_numpy(kb[1]).cuda(), factor=8).cpu().numpy()
        for i in range(2):
            r = orc.apply_kernel_degradation(torch.from_numpy(hn[i]), torch.from_numpy(kb[1]), 8).numpy()
            assert np.array_equal(np.isnan(ln[i]), np.isnan(r)), (h, w, i)


def test_selector_kernels_match_the_fp32_library_forward(K, golden, synth):
    """kmsr_selector_logits (BatchNorm folded on the host, 3xTF32-split mma convolutions, fused pooling, linear layer)
    against the torch fp32 forward of the same weights (TF32 disabled): logits to 2e-5 of their scale, identical argmax,
    deterministic run to run; patch sizes that leave partial tiles; the golden logits of the reference module."""
    from kmsr_b200.selector import Selector
    z = golden("selector.npz")
    sel = Selector.from_npz(z, "cuda")
    for n, h, w in ((6, 256, 256), (5, 100, 100), (3, 72, 200), (2, 17, 33)):
        rs = np.random.RandomState(h)
        x = torch.from_numpy((rs.standard_normal((n, 5, h, w)) * 3.0 + np.array([80, 70, 50, 25, 8])[None, :, None, None])
                             .astype(np.float32)).cuda()
        a = sel.logits(x)
        b = sel.logits_library(x)
        scale = float(b.abs().max())
        assert float((a - b).abs().max()) <= 2e-5 * scale, (n, h, w, float((a - b).abs().max()) / scale)
        assert torch.equal(a.argmax(1), b.argmax(1))
        assert torch.equal(a, sel.logits(x))
    hr = np.concatenate([synth.make_hr(4, 5100, "textured"), synth.make_hr(2, 5101, "water")])
    lg = sel.logits(torch.from_numpy(hr).cuda()).cpu().numpy()
    assert np.abs(lg - z["logits"][:6]).max() <= 1e-4 * np.abs(z["logits"]).max()
    assert np.array_equal(lg.argmax(1), z["argmax"][:6])
    assert sel.logits(torch.zeros((0, 5, 64, 64), device="cuda")).shape == (0, 10)


def test_selector_tcgen05_path(K, golden, synth):
    """kmsr_selector_logits_umma (csrc/selector_umma.cuh: tcgen05.mma kind::tf32 with TMEM accumulators, strided TMA boxes as
    the hi operand, the lo operand built in shared memory) against the torch fp32 forward (TF32 disabled), against the
    mma.sync kernels and against the golden logits of the reference module (train_gemini.py:35-39): logits to 5e-5 of
    their scale, identical argmax, bit-identical run to run; shapes the path refuses fall back (auto) or raise (umma)."""
    from kmsr_b200.selector import Selector
    z = golden("selector.npz")
    sel = Selector.from_npz(z, "cuda")
    worst = 0.0
    for n, h, w in ((6, 256, 256), (1, 256, 256), (301, 256, 256), (5, 128, 128), (3, 64, 256), (4, 256, 128), (2, 512, 256), (3, 256, 64)):
        assert K.lib.lib().kmsr_selector_umma_supported(h, w) == 1, (h, w)
        rs = np.random.RandomState(h + w)
        x = torch.from_numpy((rs.standard_normal((n, 5, h, w)) * 3.0 + np.array([80, 70, 50, 25, 8])[None, :, None, None])
                             .astype(np.float32)).cuda()
        a = sel.logits(x)
        assert sel.last_algo == "umma"
        b = sel.logits_library(x)
        c = sel.logits(x, algo="mma")
        assert sel.last_algo == "mma"
        scale = float(b.abs().max())
        worst = max(worst, float((a - b).abs().max()) / scale)
        assert float((a - b).abs().max()) <= 5e-5 * scale, (n, h, w, float((a - b).abs().max()) / scale)
        assert float((a - c).abs().max()) <= 5e-5 * scale
        assert torch.equal(a.argmax(1), b.argmax(1))
        assert torch.equal(a, sel.logits(x, algo="umma"))
    print(f"tcgen05 selector vs fp32 library forward: worst {worst:.2e} of the logit scale")
    hr = np.concatenate([synth.make_hr(4, 5100, "textured"), synth.make_hr(2, 5101, "water")])
    lg = sel.logits(torch.from_numpy(hr).cuda(), algo="umma").cpu().numpy()
    assert np.abs(lg - z["logits"][:6]).max() <= 1e-4 * np.abs(z["logits"]).max()
    assert np.array_equal(lg.argmax(1), z["argmax"][:6])
    # NaN patches give NaN logits, as the reference forward does, and do not disturb their neighbours
    xn = torch.from_numpy(hr).cuda()
    xn[1, 2, 100, 37] = float("nan")
    for algo in ("umma", "mma"):
        ln = sel.logits(xn, algo=algo).cpu().numpy()
        assert np.isnan(ln[1]).all() and np.isfinite(ln[[0, 2, 3, 4, 5]]).all(), algo
        if algo == "umma":
            assert np.array_equal(ln[[0, 2, 3, 4, 5]], lg[[0, 2, 3, 4, 5]])
    # the C ABI directly: argument errors come back as codes, nothing is launched
    import ctypes as C
    lib = K.lib.lib()
    sel._prepare_umma(torch.device("cuda", torch.cuda.current_device())) if sel._umma is None else None
    _, blobs, fc_w, fc_b = sel._umma
    vp = lambda t: C.c_void_p(t.data_ptr())
    xs = torch.zeros((2, 5, 256, 256), device="cuda")
    out = torch.zeros((2, 10), device="cuda")
    need = lib.kmsr_selector_umma_workspace_bytes(2, 256, 256)
    ws = torch.zeros(need, dtype=torch.uint8, device="cuda")
    args = lambda x_, ws_, n_ws: (vp(x_), 2, 256, 256, vp(blobs[0][0]), vp(blobs[0][1]), vp(blobs[1][0]), vp(blobs[1][1]),
                                  vp(blobs[2][0]), vp(blobs[2][1]), vp(fc_w), vp(fc_b), vp(out), ws_, n_ws, None)
    assert lib.kmsr_selector_logits_umma(*args(xs, vp(ws), need)) == 0
    assert lib.kmsr_selector_logits_umma(*args(xs, vp(ws), need - 1)) == K.lib.E_INVALID and "workspace" in K.lib.last_error()
    assert lib.kmsr_selector_logits_umma(*args(xs, None, need)) == K.lib.E_INVALID
    assert lib.kmsr_selector_logits_umma(*args(xs.view(-1)[1:], vp(ws), need)) == K.lib.E_ALIGN      # x not 16-byte aligned
    assert lib.kmsr_selector_logits_umma(vp(xs), 2, 100, 100, *args(xs, vp(ws), need)[4:]) == K.lib.E_UNSUPPORTED
    # shapes outside the path
    for h, w in ((100, 100), (17, 33), (128, 64), (256, 512)):
        assert K.lib.lib().kmsr_selector_umma_supported(h, w) == 0
        x = torch.randn((2, 5, h, w), device="cuda")
        sel.logits(x)
        assert sel.last_algo == "mma"
        with pytest.raises(K.lib.KmsrError):
            sel.logits(x, algo="umma")
    assert sel.logits(torch.zeros((0, 5, 256, 256), device="cuda"), algo="umma").shape == (0, 10)


@pytest.mark.parametrize("h,w,k,s,algo", [(64, 256, 13, 8, "tma"), (8, 256, 13, 8, "tma"), (16, 256, 13, 8, "tma"), (248, 256, 13, 8, "tma"),
                                          (64, 512, 13, 8, "tma"),
                                          (96, 128, 11, 4, "stream"), (8, 64, 15, 8, "stream"), (40, 512, 21, 2, "stream"),
                                          (24, 768, 13, 8, "stream")])
def test_non_square_and_short_patches(K, synth, bank, h, w, k, s, algo):
    """Rectangular / very short patches through the streaming kernels (few steps per band, rings barely filled)."""
    kb, _ = bank
    kern = kb[3] if k == 13 else synth.softmax_kernels(k, 5)
    n = 5
    rs = np.random.RandomState(h * 1000 + w)
    hr = (rs.standard_normal((n, 5, h, w)) * 3.0 + np.array([80, 70, 50, 25, 8])[None, :, None, None]).astype(np.float32)
    lr = K.ops.degrade_batch(torch.from_numpy(hr).cuda(), torch.from_numpy(kern).cuda(), factor=s, algo=algo).cpu().numpy()
    assert K.lib.last_algo() == algo and lr.shape == (n, 5, h // s, w // s)
    for i in range(n):
        ref = orc.apply_kernel_degradation(torch.from_numpy(hr[i]), torch.from_numpy(kern), s).numpy()
        check_pixels(lr[i], ref, hr[i], exact_degrade(hr[i], kern, s), name=f"{h}x{w} k{k} s{s} #{i}")


def test_indices_are_range_checked_at_the_boundary(K, synth, bank):
    """kidx >= nK / nidx >= nPool / crop offsets outside the image are out-of-bounds device reads inside the kernels:
    host arrays are rejected by the wrapper (IndexError, as noise_pool[idx] raises in the reference, E:72-74), device
    arrays through kmsr_validate_indices when the caller opts in, and the C entry returns KMSR_E_INVALID."""
    import ctypes as C
    kb, sb = bank
    hr = torch.from_numpy(synth.make_hr(4, 77, "textured")).cuda()
    pool = torch.from_numpy(synth.make_noise_pool(8, 1)).cuda()
    kd = torch.from_numpy(kb).cuda()
    good_k, good_n = np.array([0, 9, 3, 1], np.int32), np.array([7, 0, 2, 5], np.int32)
    K.ops.degrade_batch(hr, kd, kidx=good_k, sigma=torch.from_numpy(sb), pool=pool, nidx=good_n, noise_mode="sigma")
    for bad_k, bad_n in ((np.array([0, 10, 3, 1], np.int32), good_n), (good_k, np.array([7, 0, 8, 5], np.int32)),
                         (np.array([0, -1, 3, 1], np.int32), good_n)):
        with pytest.raises(IndexError):
            K.ops.degrade_batch(hr, kd, kidx=bad_k, sigma=torch.from_numpy(sb), pool=pool, nidx=bad_n, noise_mode="sigma")
    with pytest.raises(ValueError):                                   # one index per patch
        K.ops.degrade_batch(hr, kd, kidx=good_k[:3], sigma=torch.from_numpy(sb), pool=pool, nidx=good_n, noise_mode="sigma")
    # device-resident indices: trusted by default, checked on request
    bad_dev = torch.tensor([7, 0, 8, 5], dtype=torch.int32, device="cuda")
    with pytest.raises(IndexError):
        K.ops.degrade_batch(hr, kd, kidx=torch.from_numpy(good_k).cuda(), sigma=torch.from_numpy(sb), pool=pool, nidx=bad_dev,
                            noise_mode="sigma", validate=True)
    K.ops.degrade_batch(hr, kd, kidx=torch.from_numpy(good_k).cuda(), sigma=torch.from_numpy(sb), pool=pool,
                        nidx=torch.from_numpy(good_n).cuda(), noise_mode="sigma", validate=True)
    # the C entry itself
    scratch = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = K.lib.lib().kmsr_validate_indices(C.c_void_p(bad_dev.data_ptr()), 4, 8, C.c_void_p(scratch.data_ptr()), b"nidx", None)
    assert rc == K.lib.E_INVALID and "1 of 4" in K.lib.last_error()
    rc = K.lib.lib().kmsr_validate_indices(C.c_void_p(bad_dev.data_ptr()), 4, 9, C.c_void_p(scratch.data_ptr()), b"nidx", None)
    assert rc == 0
    # add_noise / crop_sub wrappers
    with pytest.raises(IndexError):
        K.ops.add_noise_batch(torch.zeros((4, 5, 32, 32), device="cuda"), pool, np.array([0, 1, 2, 8], np.int32))
    geo = torch.randn((5, 64, 80), device="cuda")
    den = torch.randn((5, 64, 80), device="cuda")
    K.ops.crop_sub(geo, den, [0, 32], [48, 0], 32)
    with pytest.raises(IndexError):
        K.ops.crop_sub(geo, den, [0, 33], [48, 0], 32)               # top + crop > H
    with pytest.raises(ValueError):
        K.ops.crop_sub(geo, den[:, :, :64], [0], [0], 32)            # geo - den would not broadcast (D:88)
    with pytest.raises(ValueError):
        K.ops.crop_sub(geo, den, [0], [0], 65)                       # D:44-45


def test_fused_statistics_on_shapes_the_headline_kernel_refuses(K, synth, bank):
    """degrade_batch_stats on a 512-wide band (the fused path needs W == 256): the two kernels run back to back and the
    statistics are real, not an unwritten workspace (an uninitialised probe once let this shape claim the fused path);
    asking for a kernel that cannot fuse is reported, not silently ignored."""
    kb, _ = bank
    hr = torch.randn((3, 5, 64, 512), device="cuda") * 2.0 + 40.0
    lr, m, s = K.ops.degrade_batch_stats(hr, torch.from_numpy(kb[1]).cuda(), factor=8)
    x = hr.double().flatten(2)
    assert float(((m - x.mean(dim=2)).abs() / x.mean(dim=2).abs()).max()) <= 1e-6
    assert float(((s - x.std(dim=2, unbiased=False)).abs() / x.std(dim=2, unbiased=False)).max()) <= 1e-6
    ref = K.ops.degrade_batch(hr, torch.from_numpy(kb[1]).cuda(), factor=8)
    assert torch.equal(lr, ref)
    # the headline shape through an explicitly requested non-fusing kernel: statistics still right
    hr2 = torch.randn((2, 5, 256, 256), device="cuda") + 10.0
    lr2, m2, s2 = K.ops.degrade_batch_stats(hr2, torch.from_numpy(kb[1]).cuda(), factor=8, algo="stream")
    assert K.lib.last_algo() == "stream"
    x2 = hr2.double().flatten(2)
    assert float(((m2 - x2.mean(dim=2)).abs() / x2.mean(dim=2).abs()).max()) <= 1e-6
    assert float(((s2 - x2.std(dim=2, unbiased=False)).abs() / x2.std(dim=2, unbiased=False)).max()) <= 1e-6


@pytest.mark.parametrize("k,s,down", [(15, 8, "boxmean"), (7, 8, "boxmean"), (13, 2, "boxmean"), (13, 3, "decimate")])
def test_tiled_kernel_pad_taps_never_meet_a_neighbouring_pixel(K, synth, k, s, down):
    """The tiled fallback pads a composite row to a multiple of 4 taps; a pad tap (weight 0) times a NaN / Inf pixel up to
    three columns right of an output's window would poison outputs the reference keeps finite.  Shapes with KWp > KW."""
    kern = synth.softmax_kernels(k, 3)
    h = w = 64
    hr = synth.make_hr(1, 5150 + k, "textured", size=64)[0]
    ho = len(range(0, h, s)) if down == "decimate" else h // s
    for (c, y, x, val) in ((0, 20, 37, np.nan), (2, 0, 5, np.inf), (4, 63, 63, np.nan), (1, 33, 0, -np.inf)):
        img = hr.copy()
        img[c, y, x] = val
        out = K.ops.degrade_batch(torch.from_numpy(img).cuda().unsqueeze(0), torch.from_numpy(kern).cuda(), factor=s,
                                  down_mode=down, algo="tiled")[0].cpu().numpy()
        # footprint by the definition: output (Y, X) of band c is non-finite iff its clamped window holds the pixel
        p = k // 2
        ys = np.arange(ho)[:, None, None] * s + np.arange(k + (s - 1 if down == "boxmean" else 0))[None, :, None] - p
        xs = np.arange(ho)[:, None, None] * s + np.arange(k + (s - 1 if down == "boxmean" else 0))[None, :, None] - p
        hit_y = (np.clip(ys, 0, h - 1) == y).any(axis=1)[:, 0]
        hit_x = (np.clip(xs, 0, w - 1) == x).any(axis=1)[:, 0]
        want = np.zeros((5, ho, ho), dtype=bool)
        want[c] = hit_y[:, None] & hit_x[None, :]
        assert np.array_equal(~np.isfinite(out), want), (k, s, down, c, y, x)


def test_two_rank_nccl_all_reduce_of_the_statistics(K, tmp_path):
    """Config 3 (pairs + fused statistics + the design's only collective) on two GPUs under torchrun: the all-reduced
    sums must equal one process summing the gathered per-patch values (asserted inside run_configs.config3).  Skipped on
    a single-GPU box; bench.py runs the same block at every N of the driver's scaling run."""
    import json
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "c3.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(root, "tests", "run_configs.py"), "--configs", "3", "--c3-patches", "512",
           "--c3-check", "64", "--reps", "1", "--out", str(out)]
    subprocess.run(cmd, check=True, timeout=600, cwd=root)
    r = json.load(open(out))[0]
    assert r["count"] == 1024 and r["stats_parity"]["allreduce_vs_single_process_rel"] <= 1e-12
    assert r["pixel_parity"]["checked_water"] > 0 and r["pixel_parity"]["textured"] <= 1e-5


def test_resident_noise_pool_follows_the_host_array(K, synth):
    """E.add_noise keeps the last host pool resident on the device.  The cache must notice a DIFFERENT pool of the same
    shape (a freed pool's address is commonly reused by the next np.load) and an in-place edit of the same array."""
    blurred = np.zeros((5, 32, 32), dtype=np.float32)
    pool_a = np.full((4, 5, 32, 32), 1.0, dtype=np.float32)
    np.random.seed(0)
    out = K.E.add_noise(blurred, pool_a)
    assert float(out.min()) == float(out.max()) == 1.0
    addr = pool_a.__array_interface__["data"][0]
    del pool_a
    pool_b = np.full((4, 5, 32, 32), 2.0, dtype=np.float32)          # usually lands where pool_a was
    out = K.E.add_noise(blurred, pool_b)
    assert float(out.min()) == float(out.max()) == 2.0, (addr, pool_b.__array_interface__["data"][0])
    pool_b[:] = 3.0                                                   # in-place edit of the cached array
    out = K.E.add_noise(blurred, pool_b)
    assert float(out.min()) == float(out.max()) == 3.0


def test_config2_every_one_of_the_4096_patches_against_the_oracle(K, synth, bank, golden):
    """BASELINE config 2 at its full size, patch by patch against the reference call sites ("parity on all
    4 096"): the bench workload's recipe (2048 textured + 2048 water patches, generated on the device), the job's own
    kidx / nidx (bit-exact against the stored reference draws), all 4096 x 5 x 32 x 32 LR pixels.  Textured patches are held
    to the plain north-star bar; water patches to the two-part rule, with the exact (fp64) evaluation of every 8th one and
    -- for all of them -- a bound from the sample's measured reference deviation; the fraction of water pixels over the
    plain bar is reported and bounded."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import run_configs as rc
    kb, sb = bank
    n = 4096
    g = golden("golden_rng.npz")
    kidx, nidx = K.rng.draw_multi_kernel_indices(n, 10, 4096, 42)
    assert np.array_equal(kidx, g["cfg2_kidx"]) and np.array_equal(nidx, g["cfg2_nidx"])
    pool = synth.make_noise_pool(4096, 42)
    dev = torch.device("cuda", torch.cuda.current_device())
    hr_d = rc.synth_hr_device(n, 1234, dev)                        # first half textured, second half water
    lr = K.ops.degrade_batch(hr_d, torch.from_numpy(kb).to(dev), kidx=kidx, sigma=torch.from_numpy(sb),
                             pool=torch.from_numpy(pool).to(dev), nidx=nidx, factor=8, noise_mode="sigma").cpu().numpy()
    assert K.lib.last_algo() == "tma"
    hr = hr_d.cpu().numpy()
    del hr_d
    torch.set_num_threads(os.cpu_count() or 1)
    rng = (hr.max(axis=(2, 3)) - hr.min(axis=(2, 3)))[:, :, None, None].astype(np.float64)
    worst_tex, worst_wat, over, ref_dev, ours_dev = 0.0, 0.0, 0, 0.0, 0.0
    for a0 in range(0, n, 256):                                    # the oracle, one patch per F.conv2d call (C_31:147)
        sl = slice(a0, a0 + 256)
        ref = orc.multi_kernel_pairs(hr[sl], kb, sb, pool, kidx[sl], nidx[sl], 8)
        e = np.abs(lr[sl].astype(np.float64) - ref) / rng[sl]
        assert np.isfinite(lr[sl]).all()
        if a0 < n // 2:
            worst_tex = max(worst_tex, float(e.max()))
        else:
            worst_wat = max(worst_wat, float(e.max()))
            over += int((e > PIX_TOL).sum())
            for i in range(a0, a0 + 256, 8):                       # exact value of every 8th water patch
                ex = exact_degrade(hr[i], kb[kidx[i]], 8) + sb[kidx[i]][:, None, None].astype(np.float64) * pool[nidx[i]]
                ulp = np.spacing(np.abs(ref[i - a0]).astype(np.float32)).astype(np.float64) / rng[i]
                d_ex = np.abs(lr[i].astype(np.float64) - ex) / rng[i]
                r_ex = np.abs(ref[i - a0].astype(np.float64) - ex) / rng[i]
                assert (d_ex <= 5e-6 + ulp).all(), (i, float(d_ex.max()))
                assert (e[i - a0] <= PIX_TOL + r_ex + ulp).all(), (i, float(e[i - a0].max()), float(r_ex.max()))
                ref_dev, ours_dev = max(ref_dev, float(r_ex.max())), max(ours_dev, float(d_ex.max()))
    frac_over = over / float((n // 2) * 5 * 32 * 32)
    print(f"config 2, 4096 patches: textured max {worst_tex:.2e}; water max {worst_wat:.2e}, {100 * frac_over:.4f} % of the water "
          f"pixels over 1e-5 x range; on every 8th water patch ours vs fp64 {ours_dev:.2e}, reference vs fp64 {ref_dev:.2e}")
    assert worst_tex <= PIX_TOL                                    # 2048 textured patches: the plain bar
    assert worst_wat <= PIX_TOL + ref_dev + 1e-6                   # water: the bar plus the reference's own measured deviation
    assert frac_over <= 1e-3 and ours_dev <= 5e-6 and ref_dev > PIX_TOL
