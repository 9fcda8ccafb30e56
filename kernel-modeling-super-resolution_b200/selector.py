"""Content-adaptive kernel pick: SelectorNet inference + hard selection.

The reference's mixture-of-kernels model (muti_kernel/train_gemini.py) owns a small CNN -- three stride-2 3x3
convolutions with BatchNorm + ReLU, global average pool, linear layer (train_gemini.py:14-39) -- whose logits choose
among the 10 kernels / sigma rows of the bank (train_gemini.py:107-115).  Training uses a Gumbel-softmax over the
logits; the deterministic inference form is the hard pick `argmax(logits)` (the `hard=True`, temperature -> 0 limit).

SURVEY.md lists this as a *next* row that "could stay in PyTorch": the selector below is LIBRARY code (cuDNN / cuBLAS
through torch, fp32 with TF32 disabled so that the pick is reproducible), not a hand-written kernel; what it feeds --
the per-patch `kidx` -- goes into the fused sm_100a degrade kernel exactly like the random pick of config 2.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import ops, rng

_BN_EPS = 1e-5            # nn.BatchNorm2d default (train_gemini.py:20)


class Selector:
    """SelectorNet parameters as plain tensors.  `state` maps the reference's state-dict names without the
    `selector.` prefix (features.0.weight, features.1.running_mean, ..., classifier.bias) to arrays."""

    def __init__(self, state: dict, device=None):
        dev = torch.device(device) if device is not None else torch.device("cpu")
        t = {k: torch.as_tensor(np.asarray(v)).to(dev) for k, v in state.items() if "num_batches_tracked" not in k}
        self.layers = []
        for conv, bn in ((0, 1), (3, 4), (6, 7)):
            self.layers.append((t[f"features.{conv}.weight"].float(), t[f"features.{conv}.bias"].float(),
                                t[f"features.{bn}.weight"].float(), t[f"features.{bn}.bias"].float(),
                                t[f"features.{bn}.running_mean"].float(), t[f"features.{bn}.running_var"].float()))
        self.fc_w = t["classifier.weight"].float()
        self.fc_b = t["classifier.bias"].float()
        self.device = dev

    @classmethod
    def from_state_dict_file(cls, path: str, device=None) -> "Selector":
        """moe_model.pth as written by train_gemini.py:252 (keys `selector.*`)."""
        sd = torch.load(path, map_location="cpu", weights_only=True)
        return cls({k[len("selector."):]: v.numpy() for k, v in sd.items() if k.startswith("selector.")}, device)

    @classmethod
    def from_npz(cls, z, device=None) -> "Selector":
        """tests/golden/selector.npz layout: keys w__features__0__weight, ..."""
        return cls({k[3:].replace("__", "."): z[k] for k in z.files if k.startswith("w__")}, device)

    def to(self, device) -> "Selector":
        dev = torch.device(device)
        self.layers = [tuple(p.to(dev) for p in layer) for layer in self.layers]
        self.fc_w, self.fc_b, self.device = self.fc_w.to(dev), self.fc_b.to(dev), dev
        return self

    @torch.no_grad()
    def logits(self, x: torch.Tensor, batch: int = 256) -> torch.Tensor:
        """x [N,5,H,W] float32 on this selector's device -> logits [N,10] (eval-mode BatchNorm, fp32, no TF32)."""
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            outs = []
            for a in range(0, x.shape[0], batch):
                h = x[a:a + batch]
                for w, b, g, beta, mean, var in self.layers:
                    h = F.conv2d(h, w, b, stride=2, padding=1)
                    h = F.relu(F.batch_norm(h, mean, var, g, beta, training=False, eps=_BN_EPS))
                h = F.adaptive_avg_pool2d(h, 1).flatten(1)
                outs.append(F.linear(h, self.fc_w, self.fc_b))
            return torch.cat(outs) if outs else x.new_zeros((0, self.fc_w.shape[0]))
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old

    def pick(self, x: torch.Tensor) -> torch.Tensor:
        """Hard kernel pick per patch: argmax of the logits, int32 [N]."""
        return self.logits(x).argmax(dim=1).to(torch.int32)


def degrade_content_adaptive(patches: torch.Tensor, selector: Selector, kernel_bank, sigma_bank, noise_pool,
                             seed: int = 42, downscale_factor: int = 8, pad_mode: str = "replicate",
                             down_mode: str = "boxmean", nidx=None):
    """BASELINE config 2 with the learned pick instead of the random one: kidx = argmax(selector(hr)),
    nidx from RandomState(seed) (second draw of rng.draw_multi_kernel_indices, so the noise picks equal config 2's),
    lr[n,c] = degrade(hr[n], K[kidx[n]])[c] + sigma[kidx[n],c] * pool[nidx[n],c] in one fused launch.
    `pad_mode="zero", down_mode="decimate", downscale_factor=4` gives train_gemini.py:124-137's own degrade.
    Returns (lr, kidx, nidx)."""
    ops.require_cuda()
    dev = patches.device if patches.is_cuda else torch.device("cuda", torch.cuda.current_device())
    x = patches.to(device=dev, dtype=torch.float32)
    if selector.device != dev:
        selector.to(dev)
    kidx = selector.pick(x)
    kb = torch.as_tensor(kernel_bank)
    if nidx is None:
        _, nidx = rng.draw_multi_kernel_indices(x.shape[0], kb.shape[0], len(noise_pool), seed)
    lr = ops.degrade_batch(x, kb.to(dev), kidx=kidx, sigma=sigma_bank, pool=torch.as_tensor(noise_pool).to(dev), nidx=nidx,
                           factor=int(downscale_factor), pad_mode=pad_mode, down_mode=down_mode, noise_mode="sigma")
    return (lr if patches.is_cuda else lr.cpu()), kidx.cpu().numpy(), np.asarray(nidx)
