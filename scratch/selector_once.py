import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from kmsr_b200.selector import Selector
z = np.load("/root/repo/tests/golden/selector.npz")
sel = Selector.from_npz(z, "cuda")
x = torch.randn((4096, 5, 256, 256), device="cuda") * 3 + 50
for _ in range(3): k = sel.pick(x)
torch.cuda.synchronize()
print("picks", torch.bincount(k.long(), minlength=10).tolist())
