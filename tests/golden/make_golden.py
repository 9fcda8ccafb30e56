"""Generate tests/golden/*.npz by running the REAL reference functions (build container only).

    python tests/golden/make_golden.py

The reference is pure Python; it is imported read-only from /root/reference with
netCDF4/matplotlib stand-ins (tests/golden/_refload.py).  Outputs are small fixtures that
travel to the GPU box, where /root/reference does not exist.  Inputs of the small cases are
stored verbatim; the two full-size patches are regenerated from a seed and guarded by a
sha256 of their bytes.
"""
import contextlib
import hashlib
import importlib.util
import io
import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import _refload  # noqa: E402

spec = importlib.util.spec_from_file_location(
    "kmsr_synth", os.path.join(ROOT, "kernel-modeling-super-resolution_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

REF = _refload.REF_ROOT


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r = fn(*a, **k)
    return r, buf.getvalue()


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    torch.set_num_threads(1)
    with contextlib.redirect_stdout(io.StringIO()):
        C30, C31, D, E, S, CUT = (_refload.load(t) for t in ("C30", "C31", "D", "E", "S", "CUT"))

    # ---------------- shipped kernel / sigma bank (known-answer inputs) ----------------
    kbank = np.stack([np.load(f"{REF}/moe_kernels/kernel_{i}.npy") for i in range(10)]).astype(np.float32)
    sbank = np.stack([np.load(f"{REF}/moe_kernels/sigma_{i}.npy") for i in range(10)]).astype(np.float32)
    kpb = np.load(f"{REF}/output/single_kernel/kernelgan_out/kernel_per_band_iter2400.npy").astype(np.float32)
    k2d = np.load(f"{REF}/output/single_kernel/kernelgan_out/kernel_iter2400.npy").astype(np.float32)
    # regeneration identity from moe_model.pth (train_gemini.py:68-86, 241-249)
    sd = torch.load(f"{REF}/moe_kernels/moe_model.pth", map_location="cpu", weights_only=True)
    kb = sd["kernel_bank"]
    regen_k = torch.softmax(kb.view(10, 5, -1), dim=-1).view(10, 5, 13, 13).numpy()
    regen_s = torch.nn.functional.softplus(sd["sigma_bank"]).numpy()
    np.savez_compressed(os.path.join(HERE, "moe_bank.npz"), kernels=kbank, sigmas=sbank,
                        kernel_per_band_iter2400=kpb, kernel_iter2400=k2d,
                        regen_max_err=np.array([np.abs(regen_k - kbank).max(), np.abs(regen_s - sbank).max()]))

    # ---------------- load_kernel (a1) ----------------
    lk = {}
    with tempfile.TemporaryDirectory() as td:
        rs = np.random.RandomState(3)
        k4 = rs.rand(6, 5, 13, 13).astype(np.float64)
        for name, arr in (("k3", kbank[1]), ("k2", k2d), ("k4", k4)):
            p = os.path.join(td, name + ".npy")
            np.save(p, arr)
            lk[f"{name}_in"] = arr
            (t, _) = quiet(C30.load_kernel, p)
            lk[f"{name}_c30"] = t.numpy()
            (t, _) = quiet(C31.load_kernel, p)
            lk[f"{name}_c31"] = t.numpy()
    np.savez_compressed(os.path.join(HERE, "golden_load_kernel.npz"), **lk)

    # ---------------- apply_kernel_degradation (a2) ----------------
    g = {}
    cases = []

    def add_case(name, img, kern, f, store_input=True):
        out30 = C30.apply_kernel_degradation(torch.from_numpy(img), torch.from_numpy(kern), f).numpy()
        out31 = C31.apply_kernel_degradation(torch.from_numpy(img), torch.from_numpy(kern), f).numpy()
        assert np.array_equal(out30, out31, equal_nan=True), name       # C_30 == C_31 bit for bit
        if store_input:
            g[f"{name}__img"] = img
        g[f"{name}__kernel"] = kern
        g[f"{name}__factor"] = np.array(f)
        g[f"{name}__out"] = out30
        cases.append(name)

    small = synth.make_hr(3, 11, "textured", size=64)
    smallw = synth.make_hr(1, 12, "water", size=64)
    add_case("p64_k13_s8", small[0], kbank[0], 8)
    add_case("p64_k13_s4", small[1], kbank[3], 4)
    add_case("p64_k13_s2", small[2], kbank[7], 2)
    add_case("p64_k13_s6_is4", small[0], kbank[5], 6)                 # int(log2(6)) == 2
    add_case("p64_k13_s1", small[1], kbank[2], 1)                     # no pooling
    add_case("p64_water_s8", smallw[0], kbank[9], 8)
    add_case("p64_kernel2d", small[0], k2d, 8)                        # 2-D kernel broadcast C_30:83-85
    add_case("p64_zero_taps", small[2], kpb, 8)                       # shipped kernel with exact zeros
    unnorm = (kbank[4] * np.array([2.5, 1.0, 0.5, 3.0, 1.0], np.float32)[:, None, None]).astype(np.float32)
    add_case("p64_unnormalised", small[1], unnorm, 8)                 # C_30:96 divides by the sum
    neg = kbank[6].copy(); neg[2] = -neg[2]; neg[4] = 0.0
    add_case("p64_nonpositive_sum", small[0], neg, 8)                 # sum<=0 -> used as is
    add_case("p64_k11", small[1], synth.softmax_kernels(11), 4)
    add_case("p64_k21", small[2], synth.softmax_kernels(21), 2)
    odd = synth.make_hr(1, 13, "textured", size=72)[0][:, :70, :52].copy()
    add_case("odd_70x52_s8", odd, kbank[8], 8)                        # pooling floors odd dims
    add_case("c3_bands", synth.make_hr(1, 14, "textured", size=64, bands=3)[0], kbank[0][:3].copy(), 8)
    big_t = synth.make_hr(1, 1234, "textured")[0]
    big_w = synth.make_hr(1, 1235, "water")[0]
    add_case("p256_textured", big_t, kbank[0], 8, store_input=False)
    g["p256_textured__seed"] = np.array(1234); g["p256_textured__sha256"] = np.array(sha(big_t))
    add_case("p256_water", big_w, kbank[7], 8, store_input=True)      # stored verbatim: the stress case
    g["cases"] = np.array(cases)
    # error behaviour: band mismatch -> AssertionError (C_30:88); bad ndim -> ValueError only in C_31:69
    errs = []
    for mod, tag in ((C30, "c30"), (C31, "c31")):
        for kern, kname in ((torch.zeros(4, 13, 13), "bands4"), (torch.zeros(1, 5, 13, 13), "ndim4")):
            try:
                mod.apply_kernel_degradation(torch.zeros(5, 64, 64), kern, 8)
                errs.append(f"{tag}:{kname}:none")
            except Exception as e:  # noqa: BLE001
                errs.append(f"{tag}:{kname}:{type(e).__name__}")
    g["error_types"] = np.array(errs)
    np.savez_compressed(os.path.join(HERE, "golden_degrade.npz"), **g)

    # ---------------- RNG streams + add_noise (a5) ----------------
    r = {}
    pool = synth.make_noise_pool(64, 42)
    blurred = [C30.apply_kernel_degradation(torch.from_numpy(small[i]), torch.from_numpy(kbank[0]), 2).numpy()
               for i in range(3)]                                        # (5,32,32) each
    np.random.seed(42)
    outs = [E.add_noise(b, pool) for b in blurred]
    r["pool64_seed"] = np.array(42)
    r["add_noise_blurred"] = np.stack(blurred)
    r["add_noise_out"] = np.stack(outs)
    np.random.seed(42)
    r["e_idx_seed42_pool4096_n256"] = np.array([np.random.randint(0, 4096) for _ in range(256)])
    np.random.seed(7)
    r["e_idx_seed7_pool1000_n64"] = np.array([np.random.randint(0, 1000) for _ in range(64)])
    np.random.seed(5)
    r["e_idx_seed5_pool1_n8"] = np.array([np.random.randint(0, 1) for _ in range(8)])
    rs = np.random.RandomState(42)
    r["cfg2_kidx"] = rs.randint(0, 10, 4096)
    r["cfg2_nidx"] = rs.randint(0, 4096, 4096)
    np.savez_compressed(os.path.join(HERE, "golden_rng.npz"), **r)

    # ---------------- random_crop / noise pool arithmetic (a4) ----------------
    d = {}
    rs = np.random.RandomState(21)
    shapes = [(48, 40), (32, 32), (70, 33), (64, 64)]
    geos = [rs.standard_normal((5, h, w)).astype(np.float32) * 2 + 30 for (h, w) in shapes]
    dens = [gg - rs.standard_normal(gg.shape).astype(np.float32) * 0.3 for gg in geos]
    random.seed(42)
    crops, offs = [], []
    for gg, dd in zip(geos, dens):
        noise = gg - dd                                                  # D:88
        st = random.getstate()
        ps = D.random_crop(noise, 32, 2)                                 # D:91
        random.setstate(st)                                              # replay to record offsets
        for _ in range(2):
            offs.append((random.randint(0, noise.shape[1] - 32), random.randint(0, noise.shape[2] - 32)))
        crops.extend(ps)
    d["shapes"] = np.array(shapes)
    for i, (gg, dd) in enumerate(zip(geos, dens)):
        d[f"geo{i}"] = gg; d[f"den{i}"] = dd
    d["offsets"] = np.array(offs)
    d["pool"] = np.stack(crops, axis=0)                                  # D:110
    random.seed(42)
    big_offs = []
    for _ in range(200):
        big_offs.append((random.randint(0, 256 - 32), random.randint(0, 300 - 32)))
    d["offsets_seed42_256x300_n200"] = np.array(big_offs)
    try:
        D.random_crop(np.zeros((5, 16, 64), np.float32), 32, 1)
        d["small_error"] = np.array("none")
    except Exception as e:  # noqa: BLE001
        d["small_error"] = np.array(type(e).__name__)
    np.savez_compressed(os.path.join(HERE, "golden_noise_pool.npz"), **d)

    # ---------------- analyze_radiance_stats (a7) ----------------
    s = {}
    patches = np.concatenate([synth.make_hr(3, 31, "textured", size=64), synth.make_hr(3, 32, "water", size=64)])
    patches[1, 2, 5:9, 7] = np.nan                                       # nan-skipping path
    with tempfile.TemporaryDirectory() as td:
        for i, p in enumerate(patches):
            np.save(os.path.join(td, f"patch_{i:03d}.npy"), p)
        (ret, text) = quiet(S.analyze_radiance_stats, td, 5)              # first 5 of sorted glob
    assert ret is None
    s["patches"] = patches
    s["num_samples"] = np.array(5)
    s["stdout"] = np.array(text)
    np.savez_compressed(os.path.join(HERE, "golden_stats.npz"), **s)

    # ---------------- water mask + tiling (a8) ----------------
    c = {}
    scene = synth.make_scene(41, 640, 512, n_fill=3, n_cloud=3)
    kept = []
    CUT.save_patch_as_nc = lambda patch, path, meta, i, j, h0, w0: kept.append((i, j, h0, w0, sha(patch)[:16]))
    data = scene.copy()
    (masked, _) = quiet(CUT.apply_water_mask, data, CUT.THRESHOLD_MIN, CUT.THRESHOLD_MAX)
    with tempfile.TemporaryDirectory() as td:
        ((total, nkept), _) = quiet(CUT.create_patches_nc, masked, 256, 0.5, 0.0, td, "scene", {})
    c["scene_seed"] = np.array(41); c["scene_shape"] = np.array(scene.shape)
    c["scene_sha256"] = np.array(sha(scene))
    c["masked_nan"] = np.packbits(np.isnan(masked))
    c["masked_sha256"] = np.array(sha(np.nan_to_num(masked, nan=-1.0)))
    c["inplace_nan"] = np.packbits(np.isnan(data))                      # CUT:102 mutates its argument
    c["total"] = np.array(total); c["kept"] = np.array(nkept)
    c["kept_ij"] = np.array([(k[0], k[1], k[2], k[3]) for k in kept])
    c["kept_patch_sha16"] = np.array([k[4] for k in kept])
    np.savez_compressed(os.path.join(HERE, "golden_cutter.npz"), **c)

    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
