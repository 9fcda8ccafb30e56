"""Repeat the tcgen05 selector on the same 4096 patches and on changing small batches: every run must reproduce the first
bit for bit (ring / barrier-phase races show up as run-to-run differences or hangs)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kmsr_b200.selector import Selector
z = np.load(os.path.join(ROOT, "tests", "golden", "selector.npz"))
sel = Selector.from_npz(z, "cuda")
g = torch.Generator(device="cuda").manual_seed(7)
x = torch.randn((4096, 5, 256, 256), device="cuda", generator=g) * 3 + 50
ref = sel.logits(x, algo="umma")
lib = sel.logits_library(x[:512])
assert float((ref[:512] - lib).abs().max()) <= 5e-5 * float(lib.abs().max())
bad = 0
for i in range(40):
    y = sel.logits(x, algo="umma")
    bad += int(not torch.equal(y, ref))
for n in (1, 2, 3, 7, 37, 148, 149, 1000):
    a = sel.logits(x[:n], algo="umma")
    bad += int(not torch.equal(a, ref[:n]))
xs = x[:600].reshape(600, 5, 256, 256)[:, :, :128, :128].contiguous()
r2 = sel.logits(xs, algo="umma")
for i in range(10):
    bad += int(not torch.equal(sel.logits(xs, algo="umma"), r2))
torch.cuda.synchronize()
print("selector stress: mismatching runs", bad)
sys.exit(1 if bad else 0)
