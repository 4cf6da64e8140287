"""``FusedAdam``: torch.optim.Adam's update for all parameters in one sm_100a launch
(C ABI ``nerfb200_adam_step``).  Drop-in for the optimiser the reference builds in
``utils/__init__.py:16-18`` (``Adam(parameters, lr=hparams.lr, eps=eps, weight_decay=hparams.weight_decay)``):
same constructor arguments, same state keys (``step``, ``exp_avg``, ``exp_avg_sq``), same arithmetic
(fp32, bias-corrected, L2 weight decay added to the gradient; no amsgrad / maximize)."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .nerf import _stream_ptr


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._cache = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            cache = self._cache.get(gi)
            ptrs = tuple([p.data_ptr() for p in ps])
            if cache is None or cache["ptrs"] != ptrs:       # first step / parameters moved: validate, build the tables
                for p in ps:
                    if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or p.grad.is_sparse:
                        raise RuntimeError("FusedAdam needs contiguous float32 CUDA parameters with dense gradients")
                    st = self.state[p]
                    if len(st) == 0:
                        st["step"] = 0
                        st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                        st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                steps = {self.state[p]["step"] for p in ps}
                if len(steps) != 1:
                    raise RuntimeError("FusedAdam expects all parameters of a group to be stepped together")
                chunks = []
                for i0 in range(0, len(ps), 64):
                    ch = ps[i0:i0 + 64]
                    n = len(ch)
                    arr = lambda vals: (ctypes.c_void_p * n)(*vals)
                    chunks.append(dict(n=n, i0=i0, p=arr([p.data_ptr() for p in ch]),
                                       m=arr([self.state[p]["exp_avg"].data_ptr() for p in ch]),
                                       v=arr([self.state[p]["exp_avg_sq"].data_ptr() for p in ch]),
                                       numel=(ctypes.c_int64 * n)(*[p.numel() for p in ch])))
                cache = dict(ptrs=ptrs, chunks=chunks, step=steps.pop(), dev=ps[0].device)
                self._cache[gi] = cache
            cache["step"] += 1
            b1, b2 = group["betas"]
            with torch.cuda.device(cache["dev"]):
                for ch in cache["chunks"]:
                    gs = [p.grad for p in ps[ch["i0"]:ch["i0"] + ch["n"]]]
                    gs = [g if g.is_contiguous() else g.contiguous() for g in gs]
                    garr = (ctypes.c_void_p * ch["n"])(*[g.data_ptr() for g in gs])
                    _lib.check(lib.nerfb200_adam_step(ch["n"], ch["p"], garr, ch["m"], ch["v"], ch["numel"],
                                                      float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                                      float(group["weight_decay"]), cache["step"], _stream_ptr()),
                               "nerfb200_adam_step")
            for p in ps:                         # keep the public state (state_dict / checkpoints) in step
                self.state[p]["step"] = cache["step"]
        return loss
