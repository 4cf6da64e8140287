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

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            for i0 in range(0, len(ps), 64):
                chunk = ps[i0:i0 + 64]
                for p in chunk:
                    if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or p.grad.is_sparse:
                        raise RuntimeError("FusedAdam needs contiguous float32 CUDA parameters with dense gradients")
                    st = self.state[p]
                    if len(st) == 0:
                        st["step"] = 0
                        st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                        st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["step"] += 1
                steps = {self.state[p]["step"] for p in chunk}
                if len(steps) != 1:
                    raise RuntimeError("FusedAdam expects all parameters of a group to be stepped together")
                grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in chunk]
                n = len(chunk)
                arr = lambda ts: (ctypes.c_void_p * n)(*[ctypes.c_void_p(t.data_ptr()) for t in ts])
                numel = (ctypes.c_int64 * n)(*[p.numel() for p in chunk])
                b1, b2 = group["betas"]
                with torch.cuda.device(chunk[0].device):
                    _lib.check(lib.nerfb200_adam_step(n, arr(chunk), arr(grads), arr([self.state[p]["exp_avg"] for p in chunk]),
                                                      arr([self.state[p]["exp_avg_sq"] for p in chunk]), numel,
                                                      float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                                      float(group["weight_decay"]), steps.pop(), _stream_ptr()),
                               "nerfb200_adam_step")
        return loss
