import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from kmsr_b200.selector import Selector
z = np.load("/root/repo/tests/golden/selector.npz")
sel = Selector.from_npz(z, "cuda")
x = torch.randn((4096, 5, 256, 256), device="cuda") * 3 + 50
sel.logits(x); torch.cuda.synchronize()
for reps in (1, 10, 120):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): sel.logits(x)
    e1.record(); torch.cuda.synchronize()
    print(f"{reps:4d} back-to-back picks of 4096 patches: {e0.elapsed_time(e1) / reps:.3f} ms each")
