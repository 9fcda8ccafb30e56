"""Generate tests/golden/selector.npz from the REAL reference (build container only).

    python tests/golden/make_selector_golden.py

Loads muti_kernel/train_gemini.py by path (its sibling imports `networks` / `loss` resolve from the same folder),
instantiates ContentAdaptiveDegradation, loads moe_kernels/moe_model.pth, and stores
  * the SelectorNet parameters / BatchNorm running statistics (train_gemini.py:14-39) as plain arrays,
  * the logits the reference selector (eval mode, CPU fp32) gives on 12 seeded synthetic patches, and
  * hard-selection outputs of ContentAdaptiveDegradation.forward internals for the same patches
    (argmax kernel index, effective sigma row) -- the quantities SURVEY.md 8 f2 composes.
The fixture travels to the GPU box; /root/reference does not.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import _refload  # noqa: E402

spec = importlib.util.spec_from_file_location("kmsr_synth", os.path.join(ROOT, "kernel-modeling-super-resolution_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)


def main():
    ref = _refload.REF_ROOT
    mk = os.path.join(ref, "kernel_from_lr_gan", "muti_kernel")
    sys.path.insert(0, mk)                                     # `from networks import ...`, `from loss import ...`
    _refload._stub_missing()
    sp = importlib.util.spec_from_file_location("_kmsr_ref_gem", os.path.join(mk, "train_gemini.py"))
    gem = importlib.util.module_from_spec(sp)
    sp.loader.exec_module(gem)
    model = gem.ContentAdaptiveDegradation(n_kernels=10, n_channels=5)
    sd = torch.load(os.path.join(ref, "moe_kernels", "moe_model.pth"), map_location="cpu", weights_only=True)
    model.load_state_dict(sd)
    model.eval()
    torch.set_num_threads(1)
    # radiance-like patches (far from what the shipped selector was trained on: train_gemini.py:169 feeds randn) and
    # zero-mean unit-scale fields of several smoothness / scale settings, which spread over the classes
    rs = np.random.RandomState(5103)
    z = []
    for i in range(12):
        w = rs.standard_normal((5, 64, 64))
        for _ in range(i % 4):                                  # progressively smoother
            w = 0.25 * (w + np.roll(w, 1, axis=1) + np.roll(w, 1, axis=2) + np.roll(np.roll(w, 1, axis=1), 1, axis=2))
        w = w / w.std() * [1.0, 0.5, 2.0, 0.2, 4.0][i % 5] + [0.0, 0.3, -0.5][i % 3]
        z.append(w.astype(np.float32))
    hr = np.concatenate([synth.make_hr(4, 5100, "textured"), synth.make_hr(2, 5101, "water")])
    small = np.stack(z)
    with torch.no_grad():
        logits = np.concatenate([model.selector(torch.from_numpy(hr)).numpy(), model.selector(torch.from_numpy(small)).numpy()])
        sig = model.get_effective_sigmas().numpy()
        kern = model.get_effective_kernels().numpy()
    out = {"small_inputs": small, "logits": logits, "argmax": logits.argmax(axis=1).astype(np.int32),
           "sigmas": sig, "kernels": kern}
    for k, v in sd.items():
        if k.startswith("selector."):
            out["w__" + k[len("selector."):].replace(".", "__")] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "selector.npz"), **out)
    print("argmax", out["argmax"], "margin", np.sort(logits, axis=1)[:, -1] - np.sort(logits, axis=1)[:, -2])
    print({k: v.shape for k, v in out.items() if k.startswith("w__")})


if __name__ == "__main__":
    main()
