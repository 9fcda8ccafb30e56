"""Cost of the learned kernel pick (f2, library code) next to the fused degrade it feeds: 4096 patches on one B200."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kmsr_b200 import ops, synth
from kmsr_b200.selector import Selector
z = np.load(os.path.join(ROOT, "tests", "golden", "selector.npz"))
sel = Selector.from_npz(z, "cuda")
bank = np.load(os.path.join(ROOT, "tests", "golden", "moe_bank.npz"))
n = 4096
hr = torch.randn((n, 5, 256, 256), device="cuda") * 3 + 50
pb = ops.prepare_kernels(torch.from_numpy(bank["kernels"]).cuda(), 8)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out
t_lib, _ = timed(lambda: sel.logits_library(hr).argmax(1))
t_mma, _ = timed(lambda: sel.logits(hr, algo="mma").argmax(1))
t_sel, kidx = timed(lambda: sel.pick(hr))
assert sel.last_algo == "umma"
t_deg, _ = timed(lambda: ops.degrade_batch(hr, pb, kidx=kidx, factor=8))
import json
print(json.dumps({"workload": "SelectorNet pick of 4096 patches [5,256,256] on one B200", "ms_library_fp32": t_lib, "ms_libkmsr_mma_sync": t_mma, "ms_libkmsr": t_sel, "algo": sel.last_algo, "ms_fused_degrade": t_deg, "patches_per_s": n / t_sel * 1e3, "speedup_vs_library": t_lib / t_sel}))
print(f"library forward {t_lib:.2f} ms; mma.sync kernels {t_mma:.2f} ms; selector pick {t_sel:.2f} ms ({n / t_sel * 1e3:.0f} patches/s), fused degrade {t_deg:.3f} ms; selector / degrade = {t_sel / t_deg:.0f}x")
