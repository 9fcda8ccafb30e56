"""Streaming kernel self-checks on many shapes (GPU only, no oracle): interior fast path vs general path must be
bit-identical (KMSR_STREAM_NOFAST toggles per call), and both agree with the tiled kernel."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmsr_b200.synth as synth
from kmsr_b200 import ops
torch.manual_seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    for k in (11, 13, 15, 21, 31):
        kd = torch.from_numpy(synth.softmax_kernels(k, 7)).cuda()
        for s in (2, 4, 8):
            for (h, w) in ((64, 64), (128, 128), (256, 256), (96, 512), (40, 768), (8 * s, 256), (256, 1024)):
                n = 64 if h * w <= 128 * 128 else 12
                hr = torch.randn((n, 5, h, w), device="cuda") * 3 + 50
                for pad in ("replicate", "zero"):
                    os.environ["KMSR_STREAM_NOFAST"] = "0"
                    a = ops.degrade_batch(hr, kd, factor=s, pad_mode=pad, algo="stream")
                    os.environ["KMSR_STREAM_NOFAST"] = "1"
                    b = ops.degrade_batch(hr, kd, factor=s, pad_mode=pad, algo="stream")
                    c = ops.degrade_batch(hr, kd, factor=s, pad_mode=pad, algo="tiled")
                    same = bool(torch.equal(a, b))
                    err = float((a - c).abs().max()) / 25.0
                    if not same or err > 1e-4:
                        bad += 1
                        print("BAD", k, s, h, w, pad, "fast==general", same, f"vs tiled {err:.2e}", flush=True)
print("checked, bad =", bad)
