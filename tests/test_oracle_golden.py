"""Pin the oracle (oracle/kmsr_oracle.py + oracle/oracle.c) against outputs of the REAL reference.

The golden vectors in tests/golden/*.npz were produced by tests/golden/make_golden.py, which
imports the reference functions themselves (C_30/C_31 apply_kernel_degradation + load_kernel,
D.random_crop, E.add_noise, data_mean_std.analyze_radiance_stats, A_00 apply_water_mask +
create_patches_nc).  CPU only.
"""
import hashlib
import os
import random
import re

import numpy as np
import pytest
import torch

from oracle import kmsr_oracle as orc
from oracle import oracle_c

FLOAT_TOL = 2e-6      # x band range; ATen's CPU conv picks ISA-specific kernels, so not bit-pinned across hosts


def _range(img):
    return orc.band_range(img)


def test_shipped_bank_known_answers(golden):
    z = golden("moe_bank.npz")
    k, s = z["kernels"], z["sigmas"]
    assert k.shape == (10, 5, 13, 13) and k.dtype == np.float32
    assert s.shape == (10, 5) and s.dtype == np.float32
    assert np.abs(k.sum(axis=(2, 3), dtype=np.float64) - 1.0).max() < 2e-7      # 1 +- 1.2e-7
    assert 0.7 < s.min() and s.max() < 1.0
    assert z["regen_max_err"][0] < 1e-8 and z["regen_max_err"][1] < 2e-7        # moe_model.pth identity
    assert (z["kernel_per_band_iter2400"] == 0).sum() == 81                     # exact-zero taps fixture


def test_load_kernel_forms(golden):
    z = golden("golden_load_kernel.npz")
    for name in ("k3", "k2", "k4"):
        assert np.array_equal(orc.kernel_from_array_c30(z[f"{name}_in"]).numpy(), z[f"{name}_c30"])
        assert np.array_equal(orc.kernel_from_array_c31(z[f"{name}_in"]).numpy(), z[f"{name}_c31"])
    assert z["k4_c31"].shape == (5, 13, 13) and z["k4_c30"].shape == (6, 5, 13, 13)
    assert z["k2_c31"].shape == (1, 13, 13) and z["k2_c30"].shape == (13, 13)


def _case_inputs(z, name, synth):
    if f"{name}__img" in z.files:
        img = z[f"{name}__img"]
    else:
        img = synth.make_hr(1, int(z[f"{name}__seed"]), "textured")[0]
        assert hashlib.sha256(img.tobytes()).hexdigest() == str(z[f"{name}__sha256"]), \
            "synthetic generator is not reproducing the fixture bytes on this host"
    return img, z[f"{name}__kernel"], int(z[f"{name}__factor"])


def test_degrade_python_oracle_matches_reference(golden, synth):
    torch.set_num_threads(1)
    z = golden("golden_degrade.npz")
    for name in z["cases"]:
        img, kern, f = _case_inputs(z, name, synth)
        out = orc.apply_kernel_degradation(torch.from_numpy(img), torch.from_numpy(kern), f).numpy()
        ref = z[f"{name}__out"]
        assert out.shape == ref.shape, name
        err = orc.rel_err(out, ref, _range(img))
        assert err <= FLOAT_TOL, (name, err)


def test_degrade_error_types(golden):
    z = golden("golden_degrade.npz")
    got = {}
    for strict, tag in ((False, "c30"), (True, "c31")):
        for kern, kname in ((torch.zeros(4, 13, 13), "bands4"), (torch.zeros(1, 5, 13, 13), "ndim4")):
            try:
                orc.apply_kernel_degradation(torch.zeros(5, 64, 64), kern, 8, strict_ndim=strict)
                got[f"{tag}:{kname}"] = "none"
            except Exception as e:  # noqa: BLE001
                got[f"{tag}:{kname}"] = type(e).__name__
    want = {":".join(s.split(":")[:2]): s.split(":")[2] for s in z["error_types"]}
    # C_30 with a 4-D kernel does not raise ValueError: it falls through and fails later inside torch.
    assert want["c30:bands4"] == got["c30:bands4"] == "AssertionError"
    assert want["c31:bands4"] == got["c31:bands4"] == "AssertionError"
    assert want["c31:ndim4"] == got["c31:ndim4"] == "ValueError"
    assert want["c30:ndim4"] != "none" and got["c30:ndim4"] != "none"


def test_degrade_c_restatement_matches_reference(golden, synth):
    """Plain-C fp32 (reference operation order) and fp64 truth both sit within fp32 noise of the reference."""
    z = golden("golden_degrade.npz")
    for name in z["cases"]:
        img, kern, f = _case_inputs(z, name, synth)
        if kern.ndim == 2:
            kern = np.repeat(kern[None], img.shape[0], axis=0)
        kn = oracle_c.normalize_kernel(kern)
        ref = z[f"{name}__out"]
        rng = _range(img)
        out32 = oracle_c.degrade(img, kn, f)
        out64 = oracle_c.degrade(img, kn, f, f64=True)
        assert out32.shape == ref.shape, name
        # The bound is measured, not named: how far the REAL reference's stored output is from the exact (fp64) value of
        # its own formula on this case.  An fp32 evaluation in another order (plain C here) is as far from exact as the
        # reference is, so it can differ from the reference by up to the sum of the two.
        ref_dev = orc.rel_err(ref, out64, rng)
        c32_dev = orc.rel_err(out32, out64, rng)
        assert orc.rel_err(out32, ref, rng) <= 1e-6 + ref_dev + c32_dev, (name, orc.rel_err(out32, ref, rng), ref_dev, c32_dev)
        assert ref_dev <= (4e-6 if ref_dev < 5e-6 else 2e-5), (name, ref_dev)


def test_reference_order_sensitivity_exceeds_the_pixel_bar_on_water_patches(golden, synth):
    """Evidence for the amended pixel tolerance (Parity; README.md): on the low-dynamic-range "water" fixtures
    the REAL reference's stored outputs are themselves further than 1e-5 x range from the exact value of the reference's
    formula, and a second fp32 evaluation in the reference's own operation order (plain C, oracle/oracle.c) differs from
    the reference by more than 1e-5 x range -- i.e. the north-star bar `max |ours - ref| <= 1e-5 x range` is below the
    reference's own order sensitivity there, so no independent implementation can be held to it pixel by pixel.
    Machine independent: both sides are stored reference outputs and deterministic C arithmetic."""
    z = golden("golden_degrade.npz")
    seen = {}
    for name in z["cases"]:
        img, kern, f = _case_inputs(z, name, synth)
        if kern.ndim == 2:
            kern = np.repeat(kern[None], img.shape[0], axis=0)
        kn = oracle_c.normalize_kernel(kern)
        ref = z[f"{name}__out"]
        out32, out64 = oracle_c.degrade(img, kn, f), oracle_c.degrade(img, kn, f, f64=True)
        if out32.shape != ref.shape:
            continue
        rng = _range(img)
        d_c = np.abs(out32.astype(np.float64) - ref) / rng
        d_r = np.abs(ref.astype(np.float64) - out64) / rng
        seen[name] = (float(d_c.max()), float(d_r.max()), float((d_r > 1e-5).mean()), float(np.quantile(d_r, 0.999)))
    water = {k: v for k, v in seen.items() if v[1] > 5e-6}
    assert set(water) == {"p64_water_s8", "p256_water"}, water           # measured, and it is exactly the two water fixtures
    for name, (c_vs_ref, ref_vs_exact, frac, p999) in water.items():
        assert c_vs_ref > 1e-5, (name, c_vs_ref)              # same formula, same order, other fp32 implementation: over the bar
        assert ref_vs_exact > 1e-5, (name, ref_vs_exact)      # the reference itself is over the bar from exact
        assert 0.0005 < frac < 0.05 and p999 < 1.5e-5, (name, frac, p999)   # a tail of the pixels, bounded
    for name, v in seen.items():
        if name not in water:
            assert v[0] <= 2e-6 and v[1] <= 4e-6, (name, v)   # everywhere else both sit far inside the bar


def test_normalisation_semantics(golden):
    z = golden("golden_degrade.npz")
    kern = z["p64_nonpositive_sum__kernel"]
    kn = oracle_c.normalize_kernel(kern)
    assert np.array_equal(kn[2], kern[2]) and np.array_equal(kn[4], kern[4])     # sum <= 0: untouched (C_30:96)
    kt = orc.normalize_kernel(torch.from_numpy(kern), 5).numpy()
    assert np.abs(kt - kn).max() <= 2e-7 * np.abs(kn).max()


def test_numpy_randint_streams_bit_exact(golden):
    z = golden("golden_rng.npz")
    assert np.array_equal(orc.draw_noise_indices(256, 4096, 42), z["e_idx_seed42_pool4096_n256"])
    assert np.array_equal(oracle_c.numpy_randint_stream(42, 4096, 256), z["e_idx_seed42_pool4096_n256"])
    assert np.array_equal(oracle_c.numpy_randint_stream(7, 1000, 64), z["e_idx_seed7_pool1000_n64"])
    assert np.array_equal(oracle_c.numpy_randint_stream(5, 1, 8), z["e_idx_seed5_pool1_n8"])
    kidx, nidx = orc.draw_multi_kernel_indices(4096, 10, 4096, 42)
    assert np.array_equal(kidx, z["cfg2_kidx"]) and np.array_equal(nidx, z["cfg2_nidx"])
    ck, cn = oracle_c.numpy_two_randint_vectors(42, 10, 4096, 4096)
    assert np.array_equal(ck, z["cfg2_kidx"]) and np.array_equal(cn, z["cfg2_nidx"])
    # a vector draw equals the scalar draws one by one
    np.random.seed(42)
    assert np.array_equal(np.random.randint(0, 4096, size=256), z["e_idx_seed42_pool4096_n256"])


def test_add_noise_bit_exact(golden, synth):
    z = golden("golden_rng.npz")
    pool = synth.make_noise_pool(64, int(z["pool64_seed"]))
    np.random.seed(42)
    for b, want in zip(z["add_noise_blurred"], z["add_noise_out"]):
        assert np.array_equal(orc.add_noise(b, pool), want)
    idx = oracle_c.numpy_randint_stream(42, 64, 3)
    for b, want, i in zip(z["add_noise_blurred"], z["add_noise_out"], idx):
        assert np.array_equal(oracle_c.add_noise(b, pool[i]), want)


def test_crop_offsets_and_pool_bit_exact(golden):
    z = golden("golden_noise_pool.npz")
    shapes = [tuple(s) for s in z["shapes"]]
    assert np.array_equal(orc.draw_crop_offsets(shapes, 32, 2, 42), z["offsets"])
    assert np.array_equal(oracle_c.python_crop_offsets(42, shapes, 32, 2), z["offsets"])
    assert np.array_equal(oracle_c.python_crop_offsets(42, [(256, 300)] * 200, 32, 1),
                          z["offsets_seed42_256x300_n200"])
    geos = [z[f"geo{i}"] for i in range(len(shapes))]
    dens = [z[f"den{i}"] for i in range(len(shapes))]
    assert np.array_equal(orc.build_noise_pool(geos, dens, 2, 32, 42), z["pool"])
    k = 0
    for g, d in zip(geos, dens):
        for _ in range(2):
            top, left = z["offsets"][k]
            assert np.array_equal(oracle_c.crop_sub(g, d, top, left, 32), z["pool"][k])
            k += 1
    assert str(z["small_error"]) == "ValueError"
    with pytest.raises(ValueError):
        random.seed(0)
        orc.random_crop(np.zeros((5, 16, 64), np.float32), 32, 1)


def _parse_stats(text):
    rows = re.findall(r"Band (\d)\s*\|\s*([-\d.eE+nan]+)\s*\|\s*([-\d.eE+nan]+)", text)
    means = np.array([float(r[1]) for r in rows]); stds = np.array([float(r[2]) for r in rows])
    glob = float(re.findall(r":\s*([-\d.]+)\s*\n=+\s*$", text.strip() + "\n")[0]) if rows else None
    return means, stds, glob


def test_radiance_stats_matches_reference_printout(golden):
    z = golden("golden_stats.npz")
    patches = z["patches"]
    n = int(z["num_samples"])
    want_m, want_s, want_g = _parse_stats(str(z["stdout"]))
    assert len(want_m) == 5
    _, _, am, asd, g = orc.radiance_stats(list(patches), n)
    assert np.allclose(am, want_m, atol=6e-7, rtol=0) and np.allclose(asd, want_s, atol=6e-7, rtol=0)
    assert abs(g - want_g) <= 6e-7                                    # printed with 6 decimals
    # the C two-pass fp64 statistics agree with numpy's f32 pairwise result to f32 resolution
    for p in patches[:n]:
        m, s = oracle_c.band_stats_f64(p)
        assert np.allclose(m, np.nanmean(p, axis=(1, 2)), rtol=2e-6)
        assert np.allclose(s, np.nanstd(p, axis=(1, 2)), rtol=2e-5)


def test_water_mask_and_tiling(golden, synth):
    z = golden("golden_cutter.npz")
    scene = synth.make_scene(int(z["scene_seed"]), *z["scene_shape"][1:], n_fill=3, n_cloud=3)
    assert hashlib.sha256(scene.tobytes()).hexdigest() == str(z["scene_sha256"])
    data = scene.copy()
    masked = orc.apply_water_mask(data, 1e-6, 7.0)
    assert np.array_equal(np.packbits(np.isnan(masked)), z["masked_nan"])
    assert np.array_equal(np.packbits(np.isnan(data)), z["inplace_nan"])          # in-place -9999 -> NaN
    assert hashlib.sha256(np.nan_to_num(masked, nan=-1.0).tobytes()).hexdigest() == str(z["masked_sha256"])
    keep = orc.keep_mask(masked, 256, 0.5, 0.0)
    hp, wp, stride = orc.patch_grid(scene.shape[1], scene.shape[2])
    assert hp * wp == int(z["total"]) and keep.sum() == int(z["kept"])
    assert 0 < keep.sum() < keep.size                                              # fixture drops a known subset
    ij = np.argwhere(keep)
    assert np.array_equal(ij, z["kept_ij"][:, :2]) and np.array_equal(ij * stride, z["kept_ij"][:, 2:])
    cm = oracle_c.water_mask(scene, 1e-6, 7.0)
    assert np.array_equal(np.isnan(cm), np.isnan(masked))
    assert np.array_equal(np.nan_to_num(cm, nan=-1.0), np.nan_to_num(masked, nan=-1.0))
    assert np.array_equal(oracle_c.keep_mask(cm, 256, 128, 0.0), keep)
    for (i, j), h16 in zip(ij, z["kept_patch_sha16"]):
        p = masked[:, i * stride:i * stride + 256, j * stride:j * stride + 256]
        assert hashlib.sha256(np.ascontiguousarray(p).tobytes()).hexdigest()[:16] == str(h16)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_live_reference_cross_check(synth, bank):
    """When the reference tree is present, the oracle port equals it bit for bit on fresh inputs."""
    import contextlib, io, sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import _refload
    with contextlib.redirect_stdout(io.StringIO()):
        C30 = _refload.load("C30")
    torch.set_num_threads(1)
    kb, _ = bank
    hr = synth.make_hr(2, 99, "water", size=128)
    for i in range(2):
        a = C30.apply_kernel_degradation(torch.from_numpy(hr[i]), torch.from_numpy(kb[i + 3]), 4).numpy()
        b = orc.apply_kernel_degradation(torch.from_numpy(hr[i]), torch.from_numpy(kb[i + 3]), 4).numpy()
        assert np.array_equal(a, b)
