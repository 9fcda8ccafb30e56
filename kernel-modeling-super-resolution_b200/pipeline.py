"""Host-buffer pipeline: the end-to-end form of the pair-synthesis hot path.

`PairSynthesizer` is what a data-preparation job calls with HOST arrays (the reference reads every
patch from disk into host memory, C_30:152-162 / E:223-250): HR patches stream host -> device in
chunks over a copy stream while the fused blur + downsample + noise kernel runs on the previous
chunk, and LR patches stream back.  Kernel bank, sigmas and noise pool stay resident on the device
(replicated per GPU).  Indices are host-drawn (rng.py).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops, rng


class PairSynthesizer:
    def __init__(self, kernel_bank, sigma_bank=None, noise_pool=None, *, factor: int = 8,
                 pad_mode: str = "replicate", down_mode: str = "boxmean", noise_mode: str | None = None,
                 chunk: int = 512, device=None, algo: str = "auto"):
        ops.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        kb = torch.as_tensor(kernel_bank)
        if kb.ndim == 3:
            kb = kb.unsqueeze(0)
        self.bank = ops.prepare_kernels(kb.to(self.device), factor, down_mode)
        self.sigma = None if sigma_bank is None else torch.as_tensor(sigma_bank).to(self.device, torch.float32).contiguous()
        self.pool = None if noise_pool is None else torch.as_tensor(noise_pool).to(self.device, torch.float32).contiguous()
        if noise_mode is None:
            noise_mode = "none" if self.pool is None else ("sigma" if self.sigma is not None else "add")
        self.noise_mode = noise_mode
        self.factor, self.pad_mode, self.down_mode, self.chunk, self.algo = factor, pad_mode, down_mode, chunk, algo
        self._copy = torch.cuda.Stream(self.device)
        self._out = torch.cuda.Stream(self.device)
        self._bufs = None

    def _buffers(self, c, h, w, ho, wo):
        key = (c, h, w, ho, wo)
        if self._bufs is None or self._bufs[0] != key:
            mk = lambda *s: torch.empty(s, dtype=torch.float32, device=self.device)
            self._bufs = (key, [mk(self.chunk, c, h, w) for _ in range(2)], [mk(self.chunk, c, ho, wo) for _ in range(2)])
        return self._bufs[1], self._bufs[2]

    def run_device(self, hr: torch.Tensor, kidx=None, nidx=None, out=None) -> torch.Tensor:
        """Device-resident inputs: one fused launch over the whole batch."""
        return ops.degrade_batch(hr, self.bank, kidx=kidx, sigma=self.sigma if self.noise_mode == "sigma" else None,
                                 pool=self.pool, nidx=nidx if self.noise_mode != "none" else None,
                                 factor=self.factor, pad_mode=self.pad_mode, down_mode=self.down_mode,
                                 noise_mode=self.noise_mode, out=out, algo=self.algo)

    def run_host(self, hr_host: torch.Tensor, kidx=None, nidx=None, lr_host: torch.Tensor | None = None,
                 sync: bool = True) -> torch.Tensor:
        """hr_host: CPU float32 [N,C,H,W] (pinned for full-speed copies) -> lr_host CPU [N,C,Ho,Wo].

        Double-buffered: H2D of chunk i+1 overlaps the kernel on chunk i and the D2H of chunk i-1.
        With `sync` (default) the call returns once the last device-to-host copy has landed, so the returned CPU
        tensor can be read right away; `sync=False` returns while copies may still be in flight (the current stream
        has been made to wait for them: synchronise it, or the device, before touching `lr_host`).
        """
        from . import _lib as L
        n, c, h, w = hr_host.shape
        ho, wo = L.degrade_out_size(h, w, self.bank.kh, self.bank.kw, self.factor, self.bank.down_mode)
        if lr_host is None:
            lr_host = torch.empty((n, c, ho, wo), dtype=torch.float32).pin_memory()
        ins, outs = self._buffers(c, h, w, ho, wo)
        dev = self.device
        main = torch.cuda.current_stream(dev)
        kd = None if kidx is None else torch.as_tensor(np.asarray(kidx)).to(dev, torch.int32, non_blocking=True)
        nd = None if nidx is None else torch.as_tensor(np.asarray(nidx)).to(dev, torch.int32, non_blocking=True)
        in_ready = [torch.cuda.Event() for _ in range(2)]
        in_free = [torch.cuda.Event() for _ in range(2)]
        out_ready = [torch.cuda.Event() for _ in range(2)]
        out_free = [torch.cuda.Event() for _ in range(2)]
        self._copy.wait_stream(main)
        self._out.wait_stream(main)
        nchunks = (n + self.chunk - 1) // self.chunk
        for i in range(nchunks):
            a, b = i * self.chunk, min(n, (i + 1) * self.chunk)
            s = i & 1
            with torch.cuda.stream(self._copy):
                if i >= 2:
                    self._copy.wait_event(in_free[s])
                ins[s][:b - a].copy_(hr_host[a:b], non_blocking=True)
                in_ready[s].record(self._copy)
            main.wait_event(in_ready[s])
            if i >= 2:
                main.wait_event(out_free[s])
            self.run_device(ins[s][:b - a], None if kd is None else kd[a:b], None if nd is None else nd[a:b],
                            out=outs[s][:b - a])
            in_free[s].record(main)
            out_ready[s].record(main)
            with torch.cuda.stream(self._out):
                self._out.wait_event(out_ready[s])
                lr_host[a:b].copy_(outs[s][:b - a], non_blocking=True)
                out_free[s].record(self._out)
        main.wait_stream(self._out)
        if sync:
            self._out.synchronize()
        return lr_host


def synthesize_pairs(hr, kernel_bank, sigma_bank, noise_pool, seed: int = 42, factor: int = 8, chunk: int = 512):
    """One-call form of BASELINE config 2: host HR in, (lr, kidx, nidx) out; indices from RandomState(seed)."""
    t = hr if isinstance(hr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(hr, dtype=np.float32))
    kb = torch.as_tensor(kernel_bank)
    kidx, nidx = rng.draw_multi_kernel_indices(t.shape[0], kb.shape[0], len(noise_pool), seed)
    syn = PairSynthesizer(kb, sigma_bank, noise_pool, factor=factor, chunk=chunk)
    if t.is_cuda:
        return syn.run_device(t, kidx, nidx), kidx, nidx
    lr = syn.run_host(t, kidx, nidx)
    return lr, kidx, nidx
