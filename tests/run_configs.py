#!/usr/bin/env python
"""BASELINE.json configs 1, 3, 4, 5 on the GPU (config 2 is bench.py): parity + throughput per config.

    python tests/run_configs.py --configs 1,3,4,5 [--out gpurun_out/configs.json]
    python -m torch.distributed.run --nproc-per-node N ... tests/run_configs.py --configs 3,4   # sharded

One JSON object per config on stdout (rank 0) and all of them in --out.  Timing: CUDA events on the current
stream, best of `--reps` after one warm-up, max over ranks.  The oracle (tests' checker) is used here only to
verify samples of what the GPU produced; this is a measurement + parity script (it lives under tests/ because it
uses the oracle), not part of the product.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))      # tests/ may use the oracle as a checker
sys.path.insert(0, ROOT)

import kmsr_b200.synth as synth  # noqa: E402
from kmsr_b200 import _lib, ops, rng, shard  # noqa: E402
from kmsr_b200 import A_00_patch_cutter_universal as CUT  # noqa: E402
from oracle import kmsr_oracle as orc  # noqa: E402
from oracle import oracle_c  # noqa: E402

C, P = 5, 256
HBM_PEAK = 6448.4
FP32_PEAK_TFLOPS = 2 * 36.6          # scratch/ffma2_probe.cu: 36.6 TFMA/s sustained with FFMA2 on this pool's B200;
                                     # config5 replaces it by the in-process probe (ops.fp32_peak_tflops) when it runs
try:
    HBM_PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def bank():
    z = np.load(os.path.join(ROOT, "tests", "golden", "moe_bank.npz"))
    return z["kernels"].astype(np.float32), z["sigmas"].astype(np.float32)


def synth_hr_device(n, seed, dev, size=P, water_from=None):
    """base + A*smooth + sn*white on the device (SURVEY 8d recipe); patches >= water_from use the water regime."""
    import torch.nn.functional as F
    g = torch.Generator(device=dev).manual_seed(seed)
    base = torch.tensor([80.0, 70.0, 50.0, 25.0, 8.0], device=dev).view(1, C, 1, 1)
    out = torch.empty((n, C, size, size), dtype=torch.float32, device=dev)
    wf = n // 2 if water_from is None else water_from
    step = max(1, min(256, (1 << 28) // (C * size * size)))
    for a in range(0, n, step):
        b = min(n, a + step)
        amp = torch.where(torch.arange(a, b, device=dev) < wf, 5.0, 0.3).view(-1, 1, 1, 1)
        coarse = torch.randn((b - a, C, max(size // 8, 2), max(size // 8, 2)), generator=g, device=dev)
        field = F.interpolate(coarse, size=(size, size), mode="bilinear", align_corners=False)
        white = torch.randn((b - a, C, size, size), generator=g, device=dev)
        out[a:b] = base + amp * field + (amp * 0.1) * white
    return out


def timed(fn, reps, world, before=None):
    import torch.distributed as dist
    if before is not None:
        before()
    fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        if before is not None:
            before()                                  # untimed: restores inputs that fn mutates
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        best = ms if best is None else min(best, ms)
    return best


def pix_parity(lr, ref, hr, exact=None):
    """max |ours-ref| / range, max |ours-exact| / range, max |ref-exact| / range."""
    rngs = orc.band_range(hr)
    out = {"ours_vs_ref": float((np.abs(lr.astype(np.float64) - ref) / rngs).max())}
    if exact is not None:
        out["ours_vs_exact"] = float((np.abs(lr.astype(np.float64) - exact) / rngs).max())
        out["ref_vs_exact"] = float((np.abs(ref.astype(np.float64) - exact) / rngs).max())
    return out


# ------------------------------------------------------------------------------------------------
def config1(args, rank, world, dev):
    """C_30 single-kernel apply, kernel_0.npy, 64 patches, no noise; the CPU reference path is timed."""
    kb, _ = bank()
    hr = np.concatenate([synth.make_hr(32, 1235, "textured"), synth.make_hr(32, 1236, "water")])
    t0 = time.perf_counter()
    ref = np.stack([orc.apply_kernel_degradation(torch.from_numpy(hr[i]), torch.from_numpy(kb[0]), 8).numpy() for i in range(64)])
    cpu_s = time.perf_counter() - t0
    hd = torch.from_numpy(hr).to(dev)
    pb = ops.prepare_kernels(torch.from_numpy(kb[0]).to(dev), 8)
    out = torch.empty((64, C, 32, 32), device=dev)
    ms = timed(lambda: ops.degrade_batch(hd, pb, factor=8, out=out), args.reps, 1)
    lr = out.cpu().numpy()
    exact = np.stack([oracle_c.degrade(hr[i], oracle_c.normalize_kernel(kb[0]), 8, f64=True) for i in range(64)])
    return {"config": 1, "workload": "C_30 single-kernel apply, kernel_0.npy, 64 patches [5,256,256], factor 8",
            "algo": _lib.last_algo(), "gpu_ms": ms, "gpu_pairs_per_s": 64 / (ms * 1e-3),
            "cpu_reference_pairs_per_s": 64 / cpu_s, "cpu_threads": torch.get_num_threads(),
            "parity_textured": pix_parity(lr[:32], ref[:32], hr[:32], exact[:32]),
            "parity_water": pix_parity(lr[32:], ref[32:], hr[32:], exact[32:]), "bar": 1e-5}


def config3(args, rank, world, dev):
    """E_make_train_data pair generation (C_30 blur + E noise, fused) + data_mean_std statistics + all-reduce."""
    kb, _ = bank()
    n_total = args.c3_patches * world
    a, b = rng.shard_range(n_total, rank, world)
    n = b - a
    pool_np = synth.make_noise_pool(4096, 42)
    nidx_all = rng.draw_noise_indices(n_total, 4096, 42)              # np.random.seed(42) + one draw per file (E:190, E:72)
    nidx = nidx_all[a:b]
    hr = synth_hr_device(n, 4000 + rank, dev)
    pool = torch.from_numpy(pool_np).to(dev)
    pb = ops.prepare_kernels(torch.from_numpy(kb[0]).to(dev), 8)
    nd = torch.from_numpy(nidx).to(dev)
    lr = torch.empty((n, C, 32, 32), device=dev)
    sums = torch.zeros(2 * C + 1, dtype=torch.float64, device=dev)
    res = {}

    def pairs_only():
        ops.degrade_batch(hr, pb, pool=pool, nidx=nd, factor=8, noise_mode="add", out=lr)

    def pairs_and_stats():
        sums.zero_()
        ops.degrade_batch(hr, pb, pool=pool, nidx=nd, factor=8, noise_mode="add", out=lr)
        m, s = ops.band_stats(hr, sums)
        shard.allreduce_stat_sums(sums)
        res["m"], res["s"] = m, s

    def fused():
        sums.zero_()
        _, m, s = ops.degrade_batch_stats(hr, pb, pool=pool, nidx=nd, factor=8, noise_mode="add", out=lr, sums=sums)
        shard.allreduce_stat_sums(sums)
        res["m"], res["s"] = m, s

    ms_pairs = timed(pairs_only, args.reps, world)
    ms_unfused = timed(pairs_and_stats, args.reps, world)
    lr_unfused = lr.clone()
    ms_all = timed(fused, args.reps, world)
    assert torch.equal(lr_unfused, lr), "fused-statistics kernel changed the LR pixels"
    avg_mean, avg_std, count = shard.finish_stats(sums)
    out = {"config": 3, "workload": f"E pair generation + per-band stats, {n_total} patches over {world} GPU(s) ({n} on rank 0)",
           "algo": _lib.last_algo(), "ms_pairs": ms_pairs, "pairs_per_s": n_total / (ms_pairs * 1e-3),
           "ms_pairs_plus_stats": ms_all, "pairs_per_s_with_stats": n_total / (ms_all * 1e-3),
           "ms_pairs_plus_stats_unfused": ms_unfused, "pairs_per_s_with_stats_unfused": n_total / (ms_unfused * 1e-3),
           "hbm_gbs_pairs": (4 * C * P * P + 2 * 4 * C * 1024) * n / (ms_pairs * 1e-3) / 1e9, "count": count,
           "collective": f"all_reduce(SUM) of {2 * C + 1} float64 over {world} rank(s), inside the timed region"}
    out["hbm_frac_pairs"] = out["hbm_gbs_pairs"] / HBM_PEAK

    # ---- statistics of the FULL set (every rank checks its own shard against fp64 on the device, worst over ranks) ----
    m64 = torch.empty((n, C), dtype=torch.float64, device=dev)
    s64 = torch.empty((n, C), dtype=torch.float64, device=dev)
    for a0 in range(0, n, 512):
        x = hr[a0:a0 + 512].double().flatten(2)
        m64[a0:a0 + 512] = x.mean(dim=2)
        s64[a0:a0 + 512] = x.std(dim=2, unbiased=False)
    worst = torch.stack([((res["m"] - m64).abs() / m64.abs()).max(), ((res["s"] - s64).abs() / s64).max()])
    # ---- the all-reduced sums against ONE process summing every rank's per-patch values (gathered) ----
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        gm = [torch.empty_like(res["m"]) for _ in range(world)]
        gs = [torch.empty_like(res["s"]) for _ in range(world)]
        dist.all_gather(gm, res["m"].contiguous())
        dist.all_gather(gs, res["s"].contiguous())
        gm, gs = torch.cat(gm), torch.cat(gs)
    else:
        gm, gs = res["m"], res["s"]
    one_mean = (gm.sum(dim=0) / gm.shape[0]).cpu().numpy()
    one_std = (gs.sum(dim=0) / gs.shape[0]).cpu().numpy()
    gather_rel = float(max(np.abs(one_mean - avg_mean).max() / np.abs(one_mean).max(),
                           np.abs(one_std - avg_std).max() / np.abs(one_std).max()))
    assert count == n_total and gather_rel <= 1e-12, (count, n_total, gather_rel)
    assert float(worst.max()) <= 1e-6, worst
    out["stats_parity"] = {"set": f"all {n_total} patches (worst over {world} rank(s))", "mean_rel": float(worst[0]),
                           "std_rel": float(worst[1]), "bar": 1e-6, "allreduce_vs_single_process_rel": gather_rel,
                           "allreduce_bar": 1e-12, "avg_mean": [float(v) for v in avg_mean],
                           "avg_std": [float(v) for v in avg_std]}
    if rank == 0:
        # pixels against the C_30 -> E chain of the oracle: half of the sample from the textured half of the shard,
        # half from the water half (first patches of each + random ones), so every world size checks both regimes
        half = max(1, args.c3_check // 2)
        rs = np.random.RandomState(5)
        tex = np.unique(np.concatenate([np.arange(min(half // 2, n // 2)), rs.choice(n // 2, min(half, n // 2), replace=False)]))[:half]
        wat = n // 2 + np.unique(np.concatenate([np.arange(min(half // 2, n - n // 2)),
                                                 rs.choice(n - n // 2, min(half, n - n // 2), replace=False)]))[:half]
        pick = np.concatenate([tex, wat])
        hs = hr[torch.from_numpy(pick).to(dev)].cpu().numpy()
        ls = lr[torch.from_numpy(pick).to(dev)].cpu().numpy()
        worst_px = {"textured": 0.0, "water": 0.0, "water_vs_exact": 0.0, "ref_water_vs_exact": 0.0}
        kn = oracle_c.normalize_kernel(kb[0])
        for j, i in enumerate(pick):
            ref = orc.apply_kernel_degradation(torch.from_numpy(hs[j]), torch.from_numpy(kb[0]), 8).numpy() + pool_np[nidx[i]]
            r = orc.band_range(hs[j])
            e = float((np.abs(ls[j].astype(np.float64) - ref) / r).max())
            if i < n // 2:
                worst_px["textured"] = max(worst_px["textured"], e)
            else:
                worst_px["water"] = max(worst_px["water"], e)
                if j % 8 == 0:
                    ex = oracle_c.degrade(hs[j], kn, 8, f64=True) + pool_np[nidx[i]].astype(np.float64)
                    worst_px["water_vs_exact"] = max(worst_px["water_vs_exact"], float((np.abs(ls[j] - ex) / r).max()))
                    worst_px["ref_water_vs_exact"] = max(worst_px["ref_water_vs_exact"], float((np.abs(ref - ex) / r).max()))
        out["pixel_parity"] = {"checked": int(len(pick)), "checked_textured": int(len(tex)), "checked_water": int(len(wat)),
                               **worst_px, "bar": 1e-5}
    return out


def device_scene(seed, H, W, dev):
    """[5,H,W] scene on the device: smooth radiance field, -9999 fill discs, NIR > 7 'cloud' discs (seeded)."""
    import torch.nn.functional as F
    g = torch.Generator(device=dev).manual_seed(seed)
    base = torch.tensor([80.0, 70.0, 50.0, 25.0, 3.0], device=dev).view(C, 1, 1)
    # per-band texture amplitude: a 256-pixel window sees a dynamic range of ~10-20 radiance units in the
    # visible bands; NIR stays inside the 1e-6 .. 7.0 water window (CUT:32-33) except under the cloud discs
    amp = torch.tensor([5.0, 5.0, 4.0, 2.0, 0.4], device=dev).view(C, 1, 1)
    coarse = torch.randn((1, C, H // 32, W // 32), generator=g, device=dev)
    scene = base + amp * F.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=False)[0]
    scene += 0.1 * amp * torch.randn((C, H, W), generator=g, device=dev)
    rs = np.random.RandomState(seed)
    yy = torch.arange(H, device=dev).view(H, 1)
    xx = torch.arange(W, device=dev).view(1, W)
    for kind in ("fill",) * 6 + ("cloud",) * 6:
        cy, cx = int(rs.randint(0, H)), int(rs.randint(0, W))
        r = int(rs.randint(max(2, H // 40), max(3, H // 12)))
        disc = (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
        if kind == "fill":
            scene[:, disc] = -9999.0
        else:
            scene[C - 1][disc] = 9.5
    return scene.contiguous()


def config4(args, rank, world, dev):
    """A_00 patch-cutter tiling of an 8k x 8k scene, then blur + downsample + noise on the kept windows in place."""
    kb, _ = bank()
    H = W = args.c4_size
    hp, wp, stride = CUT.patch_grid(H, W)
    per_rank = getattr(args, "c4_scene_per_rank", False)          # bench.py: one whole scene per GPU (weak scaling)
    i0, i1 = (0, hp) if per_rank else rng.shard_range(hp, rank, world)   # patch-grid rows of this rank
    full = device_scene(77 + (rank if per_rank else 0), H, W, dev)
    slab = full[:, i0 * stride:(i1 - 1) * stride + P, :].contiguous() if i1 > i0 else full[:, :0]
    del full
    raw = slab.clone()
    sh = slab.shape[1]
    pool = torch.from_numpy(synth.make_noise_pool(4096, 42)).to(dev)
    pb = ops.prepare_kernels(torch.from_numpy(kb[2]).to(dev), 8)
    masked = torch.empty_like(slab)
    state = {}

    def run():
        ops.water_mask(slab, 1e-6, 7.0, out=masked)                # mutates slab (CUT:102): restored untimed between reps
        keep, _ = ops.keep_mask(masked, P, stride, 0.0)
        ij = torch.nonzero(keep)
        offs = (ij[:, 0] * stride * W + ij[:, 1] * stride).to(torch.int64)
        k = int(offs.numel())
        nidx = torch.from_numpy(rng.draw_noise_indices(k, 4096, 42)).to(dev)
        lr = ops.degrade_batch(masked, pb, pool=pool, nidx=nidx, factor=8, noise_mode="add", patch_offsets=offs,
                               patch_hw=(P, P), strides=(sh * W, W), scene_hw=(sh, W), x_multiple=stride)
        state.update(keep=keep, ij=ij, lr=lr, nidx=nidx, k=k)

    ms = timed(run, args.reps, world, before=lambda: slab.copy_(raw))

    # the reference's nan_threshold = 0 lets mask + tiling collapse into one read of the raw scene
    # (kmsr_scene_keep_mask) with the kept windows degraded straight from it
    fstate = {}

    def run_fused():
        keep, _ = ops.scene_keep_mask(raw, 1e-6, 7.0, P, stride, 0.0)
        ij = torch.nonzero(keep)
        offs = (ij[:, 0] * stride * W + ij[:, 1] * stride).to(torch.int64)
        k = int(offs.numel())
        nidx = torch.from_numpy(rng.draw_noise_indices(k, 4096, 42)).to(dev)
        lr = ops.degrade_batch(raw, pb, pool=pool, nidx=nidx, factor=8, noise_mode="add", patch_offsets=offs,
                               patch_hw=(P, P), strides=(sh * W, W), scene_hw=(sh, W), x_multiple=stride)
        fstate.update(keep=keep, lr=lr, k=k)

    ms_fused = timed(run_fused, args.reps, world)
    assert torch.equal(fstate["keep"], state["keep"]) and torch.equal(fstate["lr"], state["lr"]), "fused scene path differs"

    def degrade_only():
        ops.degrade_batch(masked, pb, pool=pool, nidx=state["nidx"], factor=8, noise_mode="add",
                          patch_offsets=(state["ij"][:, 0] * stride * W + state["ij"][:, 1] * stride).to(torch.int64),
                          patch_hw=(P, P), strides=(sh * W, W), scene_hw=(sh, W), x_multiple=stride, out=state["lr"])
    algo_after = None
    ms_deg = timed(degrade_only, args.reps, world)
    algo_after = _lib.last_algo()
    kept = torch.tensor([state["k"]], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(kept)
    kept = int(kept.item())
    scene_bytes = 4 * C * H * W
    scenes = world if per_rank else 1
    out = {"config": 4, "workload": f"A_00 tiling of {scenes} [5,{H},{W}] scene(s) (stride 128) + degrade + noise on kept windows, {world} GPU(s)",
           "algo": algo_after, "scenes": scenes, "candidates": hp * wp * scenes, "kept": kept, "scenes_per_s_fused": scenes / (ms_fused * 1e-3), "ms_scene_total": ms, "ms_degrade_only": ms_deg,
           "ms_scene_total_fused": ms_fused, "kept_pairs_per_s_total_fused": kept / (ms_fused * 1e-3),
           "kept_pairs_per_s_total": kept / (ms * 1e-3), "kept_pairs_per_s_degrade": kept / (ms_deg * 1e-3),
           "unique_bytes_gbs_degrade": (scene_bytes * scenes + kept * 2 * 4 * C * 1024) / (ms_deg * 1e-3) / 1e9,
           "patchwise_bytes_gbs_degrade": kept * (4 * C * P * P + 2 * 4 * C * 1024) / (ms_deg * 1e-3) / 1e9}
    if rank == 0 and state["k"] > 0:
        ij = state["ij"].cpu().numpy()
        m = masked.cpu().numpy() if sh * W <= 4096 * 4096 else None
        worst, nchk = 0.0, 0
        pick = np.unique(np.linspace(0, state["k"] - 1, 12).astype(int))
        pool_np = synth.make_noise_pool(4096, 42)
        nidx = state["nidx"].cpu().numpy()
        lr = state["lr"][torch.from_numpy(pick).to(dev)].cpu().numpy()
        kn = oracle_c.normalize_kernel(kb[2])
        we = 0.0
        for j, nwin in enumerate(pick):
            i, jj = ij[nwin]
            p = (m[:, i * stride:i * stride + P, jj * stride:jj * stride + P] if m is not None
                 else masked[:, i * stride:i * stride + P, jj * stride:jj * stride + P].cpu().numpy())
            p = np.ascontiguousarray(p)
            assert not np.isnan(p).any()                               # kept windows hold no NaN (CUT:179-183)
            ref = orc.apply_kernel_degradation(torch.from_numpy(p), torch.from_numpy(kb[2]), 8).numpy() + pool_np[nidx[nwin]]
            ex = oracle_c.degrade(p, kn, 8, f64=True) + pool_np[nidx[nwin]].astype(np.float64)
            r = orc.band_range(p)
            worst = max(worst, float((np.abs(lr[j].astype(np.float64) - ref) / r).max()))
            we = max(we, float((np.abs(lr[j] - ex) / r).max()))
            nchk += 1
        # keep grid against the oracle on a corner of the slab (full-scene oracle is minutes of numpy)
        sub = raw[:, :1024, :1024].cpu().numpy()
        ref_keep = orc.keep_mask(orc.apply_water_mask(sub.copy(), 1e-6, 7.0), P, 0.5, 0.0)
        got = state["keep"][:ref_keep.shape[0], :ref_keep.shape[1]].cpu().numpy()
        out["parity"] = {"windows_checked": nchk, "ours_vs_ref": worst, "ours_vs_exact": we, "bar": 1e-5,
                         "keep_grid_corner_exact": bool(np.array_equal(got, ref_keep))}
    return out


def config5(args, rank, world, dev):
    """Roofline sweep: kernel size x patch size x factor at 1 GPU, HR batch >= --c5-gb GB."""
    global FP32_PEAK_TFLOPS
    fp32_src = "scratch/ffma2_probe.cu (FFMA2, measured earlier)"
    try:
        probe = ops.fp32_peak_tflops(dev)
        FP32_PEAK_TFLOPS = float(max(probe["burst"], probe["sustained"]))     # the larger one: never flatter a cell
        fp32_src = (f"kmsr_fp32_probe, packed FFMA2 chains with the degrade kernels' operand pattern, measured in this run: "
                    f"~2 ms launches {probe['burst']:.1f} TFLOP/s, one 150 ms launch {probe['sustained']:.1f} TFLOP/s (the larger is used)")
    except Exception:
        pass
    rows = []
    cells = getattr(args, "c5_cells", None)                       # bench.py: a few named cells instead of all 60
    for k in (11, 13, 15, 21, 31):
        kern = torch.from_numpy(synth.softmax_kernels(k, 7)).to(dev)
        for p in (64, 128, 256, 512):
            if cells is not None and not any(c[0] == k and c[1] == p for c in cells):
                continue
            n = max(8, int(args.c5_gb * 1e9 / (4 * C * p * p)))
            hr = synth_hr_device(n, 900 + p, dev, size=p)
            for s in (2, 4, 8):
                if cells is not None and (k, p, s) not in cells:
                    continue
                pb = ops.prepare_kernels(kern, s)
                ho = p // s
                out = torch.empty((n, C, ho, ho), device=dev)
                ms = timed(lambda: ops.degrade_batch(hr, pb, factor=s, out=out), args.reps, 1)
                auto_algo = _lib.last_algo()
                alts = {}
                for alg in [x for x in args.c5_algos.split(",") if x and x != auto_algo]:
                    try:
                        alts[alg] = timed(lambda: ops.degrade_batch(hr, pb, factor=s, out=out, algo=alg), args.reps, 1)
                    except Exception:
                        pass                                   # the kernel does not take this shape
                by = 4 * C * (p * p + ho * ho) * n
                fma = C * ho * ho * (k + s - 1) ** 2 * n
                rows.append({"k": k, "P": p, "s": s, "n": n, "algo": auto_algo, "ms": ms, "alt_ms": alts,
                             "pairs_per_s": n / (ms * 1e-3), "gbs": by / (ms * 1e-3) / 1e9,
                             "hbm_frac": by / (ms * 1e-3) / 1e9 / HBM_PEAK, "tflops": 2 * fma / (ms * 1e-3) / 1e12,
                             "fp32_frac": 2 * fma / (ms * 1e-3) / 1e12 / FP32_PEAK_TFLOPS,
                             "flop_per_byte": 2 * fma / by})
            del hr
    # spot parity on the extremes: 2 patches each against the oracle
    par = []
    for k, p, s in (((11, 64, 2), (31, 128, 4), (21, 512, 8), (15, 256, 2)) if cells is None else cells[:4]):
        kern = synth.softmax_kernels(k, 7)
        h = synth.make_hr(2, 31 + k, "textured", size=p)
        lr = ops.degrade_batch(torch.from_numpy(h).to(dev), torch.from_numpy(kern).to(dev), factor=s).cpu().numpy()
        e = 0.0
        for i in range(2):
            ref = orc.apply_kernel_degradation(torch.from_numpy(h[i]), torch.from_numpy(kern), s).numpy()
            e = max(e, float((np.abs(lr[i].astype(np.float64) - ref) / orc.band_range(h[i])).max()))
        par.append({"k": k, "P": p, "s": s, "ours_vs_ref": e})
    fr = sorted(max(r["hbm_frac"], r["fp32_frac"]) for r in rows)
    summary = {"cells": len(fr), "hr_gb_per_cell": args.c5_gb, "binding_roofline_frac_median": fr[len(fr) // 2] if len(fr) % 2 else
               0.5 * (fr[len(fr) // 2 - 1] + fr[len(fr) // 2]), "min": fr[0], "max": fr[-1],
               "cells_ge_0.70": sum(f >= 0.70 for f in fr), "cells_ge_0.50": sum(f >= 0.50 for f in fr), "cells_lt_0.35": sum(f < 0.35 for f in fr)}
    return {"config": 5, "workload": "roofline sweep k x P x s, 1 GPU", "hbm_peak_gbs": HBM_PEAK,
            "fp32_peak_tflops": FP32_PEAK_TFLOPS, "fp32_peak_source": fp32_src, "summary": summary,
            "rows": rows, "parity": par, "bar": 1e-5}


def selector_pick(args, rank, world, dev):
    """Row f2 (SelectorNet, train_gemini.py:14-39): the learned pick of `args.sel_patches` patches through the tcgen05
    kernels, the mma.sync kernels and the fp32 library forward (TF32 off), next to the fused degrade launch it feeds;
    picks compared with the library forward's and with the reference module's golden logits."""
    from kmsr_b200.selector import Selector
    z = np.load(os.path.join(ROOT, "tests", "golden", "selector.npz"))
    sel = Selector.from_npz(z, dev)
    kb, sb = bank()
    n = int(getattr(args, "sel_patches", 4096))
    hr = synth_hr_device(n, 4242 + rank, dev)
    pb = ops.prepare_kernels(torch.from_numpy(kb).to(dev), 8)
    ms_umma = timed(lambda: sel.logits(hr, algo="umma"), args.reps, world)
    ms_mma = timed(lambda: sel.logits(hr, algo="mma"), args.reps, world)
    a = sel.logits(hr, algo="umma")
    b = sel.logits_library(hr)
    ms_lib = timed(lambda: sel.logits_library(hr), 1, world)
    kidx = a.argmax(1).to(torch.int32)
    ms_deg = timed(lambda: ops.degrade_batch(hr, pb, kidx=kidx, factor=8), args.reps, world)
    scale = float(b.abs().max())
    hg = np.concatenate([synth.make_hr(4, 5100, "textured"), synth.make_hr(2, 5101, "water")])
    lg = sel.logits(torch.from_numpy(hg).to(dev), algo="umma").cpu().numpy()
    flop = 2.0 * n * (128 * 128 * 32 * 45 + 64 * 64 * 64 * 288 + 32 * 32 * 128 * 576)
    return {"row": "f2", "workload": f"SelectorNet pick of {n} patches [5,256,256] per GPU (3 convolutions + pooling + linear layer)",
            "algo": "tcgen05 (kind::tf32, accumulators and A operand in tensor memory, 3xTF32 split)",
            "ms_tcgen05": ms_umma, "ms_mma_sync": ms_mma, "ms_library_fp32": ms_lib, "ms_fused_degrade": ms_deg,
            "patches_per_s": n * world / (ms_umma * 1e-3), "useful_tflops": flop / (ms_umma * 1e-3) / 1e12,
            "tensor_tflops_3x": 3 * flop / (ms_umma * 1e-3) / 1e12,
            "logits_vs_library_over_scale": float((a - b).abs().max()) / scale,
            "picks_equal_library": bool(torch.equal(a.argmax(1), b.argmax(1))),
            "golden_logits_over_scale": float(np.abs(lg - z["logits"][:6]).max() / np.abs(z["logits"]).max()),
            "golden_picks_equal": bool(np.array_equal(lg.argmax(1), z["argmax"][:6])), "bar": 1e-4}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,3,4,5")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--c3-patches", type=int, default=12500, help="config 3 patches per GPU")
    ap.add_argument("--c3-check", type=int, default=2048)
    ap.add_argument("--c4-size", type=int, default=8192)
    ap.add_argument("--c5-gb", type=float, default=4.0)
    ap.add_argument("--c5-algos", default="", help="config 5: also time these kernels (comma list of box,reg,stream,tma)")
    args = ap.parse_args()
    assert torch.cuda.is_available(), "needs a CUDA device; there is no CPU fallback"
    rank, local, world = shard.init_distributed()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.set_num_threads(os.cpu_count() or 1)
    fns = {1: config1, 3: config3, 4: config4, 5: config5, 2: selector_pick}      # 2 = row f2's selector leg (config 2 itself is bench.py)
    results = []
    for c in [int(x) for x in args.configs.split(",")]:
        if world > 1 and c in (1, 5):
            continue
        t0 = time.perf_counter()
        r = fns[c](args, rank, world, dev)
        r["wall_s"] = time.perf_counter() - t0
        torch.cuda.empty_cache()
        if rank == 0:
            print(json.dumps(r))
            sys.stdout.flush()
            results.append(r)
    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(results, open(args.out, "w"), indent=1)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
