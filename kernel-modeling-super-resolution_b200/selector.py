"""Content-adaptive kernel pick: SelectorNet inference + hard selection.

The reference's mixture-of-kernels model (muti_kernel/train_gemini.py) owns a small CNN -- three stride-2 3x3
convolutions with BatchNorm + ReLU, global average pool, linear layer (train_gemini.py:14-39) -- whose logits choose
among the 10 kernels / sigma rows of the bank (train_gemini.py:107-115).  Training uses a Gumbel-softmax over the
logits; the deterministic inference form is the hard pick `argmax(logits)` (the `hard=True`, temperature -> 0 limit).

On a CUDA tensor the logits come from libkmsr's own kernels (`kmsr_selector_logits`, csrc/selector.cu): BatchNorm is
folded into the convolution weights here on the host, the weights are laid out and split into TF32 hi / lo parts for
the 3xTF32 tensor-core MMAs (fp32-level accuracy, so the pick does not depend on reduced precision), and the three
convolutions, the pooling and the linear layer run as four launches.  `logits_library` keeps the torch / cuDNN fp32
evaluation (TF32 disabled) -- the form SURVEY.md says "could stay in PyTorch" -- as the cross-check the tests use and
for CPU tensors.  What the pick feeds -- the per-patch `kidx` -- goes into the fused sm_100a degrade kernel exactly
like the random pick of config 2.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib as L
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
        self._cuda = None
        self._umma = None
        self.last_algo = None

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
        self._cuda = None
        self._umma = None
        return self

    # ---- libkmsr path ---------------------------------------------------------------------------------------------
    @staticmethod
    def _tf32(a: np.ndarray) -> np.ndarray:
        """cvt.rna.tf32.f32 on the host: round the magnitude to 10 mantissa bits, ties away from zero."""
        b = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
        return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)

    def _prepare(self, dev):
        """Fold BatchNorm (train_gemini.py:19-29, eval mode) into each convolution and build the split weight blobs
        [chunk][channel block][tap][quad][lane][4] (fragment-major, TF32 hi / lo parts) that conv_mma_kernel stages."""
        blobs = []
        for w, b, g, beta, mean, var in self.layers:
            w64, b64 = w.double().cpu().numpy(), b.double().cpu().numpy()
            sc = g.double().cpu().numpy() / np.sqrt(var.double().cpu().numpy() + _BN_EPS)
            wf = (w64 * sc[:, None, None, None]).astype(np.float32)             # [cout, cin, 3, 3]
            bf = ((b64 - mean.double().cpu().numpy()) * sc + beta.double().cpu().numpy()).astype(np.float32)
            cout, cin = wf.shape[0], wf.shape[1]
            nt = 8 if cout >= 64 else 4
            chunks, nblk = (cin + 7) // 8, cout // (8 * nt)
            hi = self._tf32(wf)
            lo = self._tf32(wf - hi)
            # fragment-major: blob[chunk, block, tap, quad, lane, 4]; lane (g = lane >> 2, t = lane & 3) owns, per tap, the
            # B values (part, j, h) = W_part[n = 8 (block * nt + j) + g][c = 8 chunk + t + 4 h][tap] at index part * 2 nt + 2 j + h
            wpad = np.zeros((2, cout, 8 * chunks, 9), dtype=np.float32)
            wpad[0, :, :cin] = hi.reshape(cout, cin, 9)
            wpad[1, :, :cin] = lo.reshape(cout, cin, 9)
            blob = np.zeros((chunks, nblk, 9, nt, 32, 4), dtype=np.float32)
            lane = np.arange(32)
            g, t = lane >> 2, lane & 3
            for part in range(2):
                for j in range(nt):
                    for h in range(2):
                        idx = part * 2 * nt + 2 * j + h
                        for ch in range(chunks):
                            for nb in range(nblk):
                                n_idx = 8 * (nb * nt + j) + g                       # [32]
                                c_idx = 8 * ch + t + 4 * h                           # [32]
                                blob[ch, nb, :, idx // 4, :, idx % 4] = wpad[part, n_idx, c_idx, :].T   # [9, 32]
            assert blob.size == L.check(int(L.lib().kmsr_selector_weight_floats(cin, cout)))
            blobs.append((torch.from_numpy(blob.reshape(-1)).to(dev), torch.from_numpy(bf).to(dev)))
        self._cuda = (dev, blobs, self.fc_w.to(dev).contiguous(), self.fc_b.to(dev).contiguous())

    def _folded(self):
        """(weight [cout, cin, 3, 3] float32, bias [cout] float32) per layer with eval-mode BatchNorm folded in (fp64)."""
        out = []
        for w, b, g, beta, mean, var in self.layers:
            w64, b64 = w.double().cpu().numpy(), b.double().cpu().numpy()
            sc = g.double().cpu().numpy() / np.sqrt(var.double().cpu().numpy() + _BN_EPS)
            out.append(((w64 * sc[:, None, None, None]).astype(np.float32),
                        ((b64 - mean.double().cpu().numpy()) * sc + beta.double().cpu().numpy()).astype(np.float32)))
        return out

    @classmethod
    def umma_stage_images(cls, wf: np.ndarray) -> np.ndarray:
        """Weight stages of conv_umma_kernel (csrc/selector_umma.cuh) for a folded weight [cout, cin, 3, 3]:
        float32 [stage][4 chunks][2 cout / 8][8][4] -- per stage of 16 values of K the shared-memory image of the operand
        [B_hi ; B_lo] in the K-major canonical layout without swizzle (core matrix = 8 rows x 16 bytes).  Rows < cout hold the
        TF32 hi part of output channel `row`, rows >= cout the lo part of channel `row - cout`.  K runs tap-major in groups
        of 16 input channels (stage = tap * cin / 16 + group, k = 16 group + 4 chunk + e); the 5-channel first layer puts the
        five (band, kernel row) triples T = 5 stage + j = 3 band + dy at k = 3 j + dx of a stage and zero at k = 15."""
        cout, cin = wf.shape[0], wf.shape[1]
        w9 = wf.reshape(cout, cin, 9)
        if cin == 5:
            stages = 3
            flat = np.zeros((cout, 3, 16), dtype=np.float32)
            flat[:, :, :15] = w9.reshape(cout, 3, 15)              # stage s: triples T = 5 s + j = 3 band + dy, k = 3 j + dx; k = 15: zero
            flat = flat.reshape(cout, 48)
        else:
            assert cin % 16 == 0, cin
            stages = 9 * (cin // 16)
            flat = np.ascontiguousarray(w9.transpose(0, 2, 1)).reshape(cout, 9 * cin)   # k = tap * cin + channel
        hi = cls._tf32(flat)
        lo = cls._tf32(flat - hi)
        cat = np.concatenate([hi, lo], axis=0)                                    # [2 cout, K]
        img = cat.reshape(2 * cout // 8, 8, stages, 4, 4).transpose(2, 3, 0, 1, 4)  # [stage][chunk][row group][row][e]
        return np.ascontiguousarray(img, dtype=np.float32)

    def _prepare_umma(self, dev):
        blobs = []
        for wf, bf in self._folded():
            img = self.umma_stage_images(wf)
            assert img.size == L.check(int(L.lib().kmsr_selector_umma_weight_floats(wf.shape[1], wf.shape[0])))
            blobs.append((torch.from_numpy(img.reshape(-1)).to(dev), torch.from_numpy(bf).to(dev)))
        self._umma = (dev, blobs, self.fc_w.to(dev).contiguous(), self.fc_b.to(dev).contiguous())

    @torch.no_grad()
    def logits(self, x: torch.Tensor, batch: int = 4096, algo: str = "auto") -> torch.Tensor:
        """x [N,5,H,W] float32 -> logits [N,10].  CUDA tensors go through libkmsr (3xTF32 tensor-core convolutions with
        folded BatchNorm); CPU tensors through the torch fp32 evaluation.  `algo`: "umma" = the tcgen05 / tensor-memory
        kernels (256 x 256, 128 x 128, ... patches: kmsr_selector_umma_supported), "mma" = the mma.sync kernels (any
        shape), "auto" = umma where the shape qualifies."""
        if not x.is_cuda:
            return self.logits_library(x)
        ops.require_cuda()
        dev = x.device
        x = x.to(torch.float32).contiguous()
        n, c, h, w = x.shape
        assert c == 5, f"SelectorNet takes 5 bands, got {c}"
        if algo not in ("auto", "umma", "mma"):
            raise ValueError(f"algo must be 'auto', 'umma' or 'mma', got {algo!r}")
        umma = algo == "umma" or (algo == "auto" and bool(L.lib().kmsr_selector_umma_supported(h, w)))
        if umma:
            if getattr(self, "_umma", None) is None or self._umma[0] != dev:
                self._prepare_umma(dev)
            _, blobs, fc_w, fc_b = self._umma
            ws_fn, fn = L.lib().kmsr_selector_umma_workspace_bytes, L.lib().kmsr_selector_logits_umma
        else:
            if getattr(self, "_cuda", None) is None or self._cuda[0] != dev:
                self._prepare(dev)
            _, blobs, fc_w, fc_b = self._cuda
            ws_fn, fn = L.lib().kmsr_selector_workspace_bytes, L.lib().kmsr_selector_logits
        self.last_algo = "umma" if umma else "mma"
        out = torch.empty((n, 10), dtype=torch.float32, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            for a in range(0, n, batch):
                m = min(batch, n - a)
                wsb = L.check(int(ws_fn(m, h, w)))
                ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
                L.check(fn(p(x[a:a + m]), m, h, w, p(blobs[0][0]), p(blobs[0][1]), p(blobs[1][0]), p(blobs[1][1]),
                           p(blobs[2][0]), p(blobs[2][1]), p(fc_w), p(fc_b), p(out[a:a + m]), p(ws), wsb, st))
        return out

    @torch.no_grad()
    def logits_library(self, x: torch.Tensor, batch: int = 256) -> torch.Tensor:
        """The same forward through torch (cuDNN / cuBLAS or CPU): eval-mode BatchNorm, fp32, TF32 disabled."""
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
