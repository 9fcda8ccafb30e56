"""Library-side variants of the SelectorNet forward (f2): BN folded into the convolutions, cuDNN autotune, channels_last."""
import os, sys, numpy as np, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kmsr_b200.selector import Selector
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
z = np.load(os.path.join(ROOT, "tests", "golden", "selector.npz"))
sel = Selector.from_npz(z, "cuda")
n = 2048
x = torch.randn((n, 5, 256, 256), device="cuda") * 3 + 50
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out
t0, ref = timed(lambda: sel.logits(x))
print(f"as shipped: {t0:.1f} ms per {n}")
folded = []
for w, b, g, beta, mean, var in sel.layers:
    sc = (g.double() / torch.sqrt(var.double() + 1e-5))
    folded.append(((w.double() * sc[:, None, None, None]).float(), ((b.double() - mean.double()) * sc + beta.double()).float()))
@torch.no_grad()
def fwd(x, batch=256, cl=False):
    outs = []
    for a in range(0, x.shape[0], batch):
        h = x[a:a + batch]
        if cl: h = h.contiguous(memory_format=torch.channels_last)
        for w, b in folded:
            h = F.relu_(F.conv2d(h, w.contiguous(memory_format=torch.channels_last) if cl else w, b, stride=2, padding=1))
        outs.append(F.linear(h.mean(dim=(2, 3)), sel.fc_w, sel.fc_b))
    return torch.cat(outs)
for name, kw in (("folded", {}), ("folded b=512", {"batch": 512}), ("folded channels_last", {"cl": True})):
    t, out = timed(lambda: fwd(x, **kw))
    rel = float((out - ref).abs().max() / ref.abs().max())
    flips = int((out.argmax(1) != ref.argmax(1)).sum())
    print(f"{name}: {t:.1f} ms, max rel logit diff {rel:.1e}, argmax flips {flips}")
torch.backends.cudnn.benchmark = True
for name, kw in (("folded + cudnn.benchmark", {}), ("folded cl + benchmark", {"cl": True})):
    t, out = timed(lambda: fwd(x, **kw), reps=4)
    rel = float((out - ref).abs().max() / ref.abs().max())
    print(f"{name}: {t:.1f} ms, max rel logit diff {rel:.1e}, argmax flips {int((out.argmax(1) != ref.argmax(1)).sum())}")
