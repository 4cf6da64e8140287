"""Training path of ``render_rays`` (reference: train.py:103-117 — loss.backward() through
models/rendering.py + models/nerf.py): a ``torch.autograd.Function`` around the fused kernel.

forward : ONE fused launch in "save" mode — it renders exactly like inference and additionally
          writes, per sample, the post-activation outputs of the 8 hidden layers and of the
          direction layer (fp16) plus raw sigma / rgb (fp32) to HBM (include/nerf_pl_b200.h,
          ``save_*`` fields).  No autograd graph is recorded for the 196,608 x 12 linear layers.
backward: (i) the compositing quadrature is re-evaluated on the saved (R, S) sigma / rgb tensors
          with torch ops to turn d(rgb, depth, opacity) into per-sample d(sigma), d(rgb);
          (ii) the MLP is back-propagated by hand, layer by layer, with fp16 tensor-core GEMMs
          (``torch.mm(..., out_dtype=float32)`` → cuBLAS) on the saved activations: wgrad =
          dY^T X, dgrad = dY W, ReLU masks from the saved outputs.  Per-sample gradients are
          scaled by a power of two chosen on the device (no host sync) so they sit in fp16's
          normal range; weight gradients are accumulated in fp32 and unscaled.
The sampling of the fine depths carries no gradient (models/rendering.py:225-227 ``.detach()``).

The hand-written sm_100a part is the forward + activation capture; the backward GEMMs are
library calls for now (DESIGN.md section 9 lists the fused tcgen05 dgrad/wgrad as the next step).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch

from . import _lib
from .nerf import _stream_ptr, nerf_parameters, packed_weights


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _embed(x: torch.Tensor, n_freqs: int) -> torch.Tensor:
    parts = [x]
    for k in range(n_freqs):
        parts += [torch.sin((2.0 ** k) * x), torch.cos((2.0 ** k) * x)]
    return torch.cat(parts, -1)


def _coarse_depths(rays, S, use_disp, perturb, pr):
    near, far = rays[:, 6:7], rays[:, 7:8]
    t = torch.linspace(0, 1, S, device=rays.device)
    z = 1 / (1 / near * (1 - t) + 1 / far * t) if use_disp else near * (1 - t) + far * t
    z = z.expand(rays.shape[0], S)
    if perturb > 0:
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        upper = torch.cat((mid, z[:, -1:]), -1)
        lower = torch.cat((z[:, :1], mid), -1)
        z = lower + (upper - lower) * (perturb * pr)
    return z


def _composite(sig, col, z, dnorm, noise, noise_std, white_back):
    """models/rendering.py:143-170 on (R,S) sigma, (R,S,3) rgb (differentiable torch ops)."""
    delta = torch.cat((z[:, 1:] - z[:, :-1], torch.full_like(z[:, :1], 1e10)), -1) * dnorm
    s = sig if noise is None else sig + noise * noise_std
    alpha = 1 - torch.exp(-delta * torch.relu(s))
    trans = torch.cumprod(torch.cat((torch.ones_like(alpha[:, :1]), 1 - alpha + 1e-10), -1), -1)[:, :-1]
    w = alpha * trans
    opac = w.sum(1)
    rgb = (w[..., None] * col).sum(-2)
    depth = (w * z).sum(-1)
    if white_back:
        rgb = rgb + 1 - opac[:, None]
    return rgb, depth, opac


def _mm32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """fp16 x fp16 -> fp32 GEMM on the tensor cores."""
    return torch.mm(a, b, out_dtype=torch.float32)


def _relu_backward(dh: torch.Tensor, act: torch.Tensor):
    """dpre = dh * (act > 0) plus its transpose, one fused kernel (C ABI nerfb200_relu_backward)."""
    S, C = dh.shape
    dpre = torch.empty_like(dh)
    dpre_t = torch.empty(C, S, dtype=dh.dtype, device=dh.device)
    lib = _lib.load()
    with torch.cuda.device(dh.device):
        _lib.check(lib.nerfb200_relu_backward(dh.data_ptr(), act.data_ptr(), S, C, dpre.data_ptr(), dpre_t.data_ptr(),
                                              _stream_ptr()), "nerfb200_relu_backward")
    return dpre, dpre_t


def _mlp_backward(params: List[torch.Tensor], acts: torch.Tensor, d_act: torch.Tensor, enc: torch.Tensor,
                  dir_enc: torch.Tensor, samples_per_ray: int, dsig: torch.Tensor, dpre_rgb: torch.Tensor
                  ) -> List[torch.Tensor]:
    """Hand-written backward of NeRF.forward (models/nerf.py:100-124) for S samples.
      params  : 24 fp32 tensors (state_dict order)
      acts    : (8, S, 256) fp16 outputs of xyz_encoding_1..8;  d_act: (S,128) fp16 output of dir_encoding
      enc     : (S, 64) fp16 encoded xyz (col 63 zero);  dir_enc: (R, 27) fp32 encoded directions
      dsig    : (S,) fp32 dL/dsigma;  dpre_rgb: (S,3) fp32 dL/d(rgb pre-sigmoid)
    Returns the 24 gradients in the same order."""
    W = [params[2 * i] for i in range(12)]
    grads: List[Optional[torch.Tensor]] = [None] * 24
    S = acts.shape[1]
    R = S // samples_per_ray
    ones = torch.ones(1, S, dtype=torch.float16, device=acts.device)

    amax = torch.maximum(dsig.abs().max(), dpre_rgb.abs().max()).clamp_min(1e-30)
    scale = torch.exp2(torch.floor(torch.log2(1024.0 / amax)))          # device scalar, power of two
    inv = 1.0 / scale
    g3 = (dpre_rgb * scale).half()                                       # (S,3)
    gs = (dsig * scale).half()                                           # (S,)
    h8 = acts[7]

    # rgb head: pre = W_rgb d + b_rgb
    grads[22] = _mm32(g3.t().contiguous(), d_act) * inv                   # (3,128)
    grads[23] = g3.float().sum(0) * inv
    dd, dd_t = _relu_backward(torch.mm(g3, W[11].half()), d_act)         # (S,128), (128,S)
    # dir_encoding: d = relu(W_d [final, dir] + b_d), final = W_f h8 + b_f (no activation)
    final = torch.addmm(params[17].half(), h8, W[8].half().t())          # (S,256) fp16
    gWd = torch.empty_like(W[9])
    gWd[:, :256] = _mm32(dd_t, final) * inv
    dd_ray = dd.view(R, samples_per_ray, 128).float().sum(1)             # direction is constant per ray
    gWd[:, 256:] = (dd_ray.t() @ dir_enc) * inv
    grads[18] = gWd
    grads[19] = _mm32(ones, dd).view(-1) * inv
    dfinal = torch.mm(dd, W[9][:, :256].half())                          # (S,256)
    grads[16] = _mm32(dfinal.t().contiguous(), h8) * inv
    grads[17] = _mm32(ones, dfinal).view(-1) * inv
    # sigma head + layer 8
    grads[20] = _mm32(gs.view(1, S), h8) * inv
    grads[21] = (gs.float().sum() * inv).view(1)
    dh = torch.addmm(gs.view(S, 1) * W[10].half().view(1, 256), dfinal, W[8].half())   # (S,256)
    for l in range(7, -1, -1):                                           # xyz_encoding_{l+1}
        dpre, dpre_t = _relu_backward(dh.contiguous(), acts[l])
        grads[2 * l + 1] = _mm32(ones, dpre).view(-1) * inv
        if l == 0:
            grads[0] = (_mm32(dpre_t, enc) * inv)[:, :63].contiguous()
        elif l == 4:                                                     # skip: input = [enc(63), h4]
            g = torch.empty_like(W[4])
            g[:, :63] = (_mm32(dpre_t, enc) * inv)[:, :63]
            g[:, 63:] = _mm32(dpre_t, acts[3]) * inv
            grads[8] = g
            dh = torch.mm(dpre, W[4][:, 63:].half())
        else:
            grads[2 * l] = _mm32(dpre_t, acts[l - 1]) * inv
            dh = torch.mm(dpre, W[l].half())
    return grads


class FusedRenderFunction(torch.autograd.Function):
    """rays + pre-drawn randoms + 48 parameter tensors -> the six result tensors."""

    @staticmethod
    def forward(ctx, cfg: Dict, rays, pr, nc, ur, nf, *params):
        models = cfg["models"]
        S_c, K = cfg["N_samples"], cfg["N_importance"]
        S_f = S_c + K
        n = rays.shape[0]
        dev = rays.device
        f32 = dict(dtype=torch.float32, device=dev)
        f16 = dict(dtype=torch.float16, device=dev)
        out = [torch.empty(n, 3, **f32), torch.empty(n, **f32), torch.empty(n, **f32)]
        if K > 0:
            out += [torch.empty(n, 3, **f32), torch.empty(n, **f32), torch.empty(n, **f32)]
        z_fine = torch.empty(n, S_f, **f32) if K > 0 else None
        act_c, d_c = torch.empty(8, n * S_c, 256, **f16), torch.empty(n * S_c, 128, **f16)
        sig_c, rgb_c = torch.empty(n * S_c, **f32), torch.empty(n * S_c, 3, **f32)
        act_f = d_f = sig_f = rgb_f = None
        if K > 0:
            act_f, d_f = torch.empty(8, n * S_f, 256, **f16), torch.empty(n * S_f, 128, **f16)
            sig_f, rgb_f = torch.empty(n * S_f, **f32), torch.empty(n * S_f, 3, **f32)
        lib = _lib.load()
        blob_c = packed_weights(models[0])
        blob_f = packed_weights(models[1]) if K > 0 else None
        args = _lib.RenderArgs(
            rays=rays.data_ptr(), n_rays=n, ray_stride=rays.stride(0),
            packed_coarse=blob_c.data_ptr(), packed_fine=_ptr(blob_f),
            n_samples=S_c, n_importance=K, use_disp=int(cfg["use_disp"]), perturb=cfg["perturb"],
            noise_std=cfg["noise_std"], white_back=int(cfg["white_back"]), test_time=0,
            perturb_rand=_ptr(pr), noise_coarse=_ptr(nc), u_rand=_ptr(ur), noise_fine=_ptr(nf),
            rgb_coarse=out[0].data_ptr(), depth_coarse=out[1].data_ptr(), opacity_coarse=out[2].data_ptr(),
            rgb_fine=_ptr(out[3]) if K > 0 else None, depth_fine=_ptr(out[4]) if K > 0 else None,
            opacity_fine=_ptr(out[5]) if K > 0 else None,
            z_fine=_ptr(z_fine), weights_coarse=None, weights_fine=None, status=None, max_ctas=0,
            save_act_coarse=act_c.data_ptr(), save_act_fine=_ptr(act_f), save_dir_coarse=d_c.data_ptr(),
            save_dir_fine=_ptr(d_f), save_sigma_coarse=sig_c.data_ptr(), save_sigma_fine=_ptr(sig_f),
            save_rgb_coarse=rgb_c.data_ptr(), save_rgb_fine=_ptr(rgb_f))
        with torch.cuda.device(dev):
            _lib.check(lib.nerfb200_render_rays(ctypes.byref(args), _stream_ptr()), "nerfb200_render_rays")
        ctx.cfg = cfg
        ctx.has_fine = K > 0
        ctx.opt = (pr, nc, nf, z_fine, act_c, d_c, sig_c, rgb_c, act_f, d_f, sig_f, rgb_f)
        ctx.save_for_backward(rays, *params)
        return tuple(out)

    @staticmethod
    def backward(ctx, *gouts):
        cfg = ctx.cfg
        rays, *params = ctx.saved_tensors
        pr, nc, nf, z_fine, act_c, d_c, sig_c, rgb_c, act_f, d_f, sig_f, rgb_f = ctx.opt
        S_c, K = cfg["N_samples"], cfg["N_importance"]
        n = rays.shape[0]
        o, d = rays[:, 0:3], rays[:, 3:6]
        dnorm = d.norm(dim=-1, keepdim=True)
        dir_enc = _embed(d, 4)
        noise_std = cfg["noise_std"]
        grads: List[Optional[torch.Tensor]] = []
        passes = [(params[:24], act_c, d_c, sig_c, rgb_c, _coarse_depths(rays, S_c, cfg["use_disp"], cfg["perturb"], pr),
                   nc, gouts[0:3])]
        if ctx.has_fine:
            passes.append((params[24:48], act_f, d_f, sig_f, rgb_f, z_fine, nf, gouts[3:6]))
        for prm, acts, d_act, sig, col, z, noise, (g_rgb, g_depth, g_opac) in passes:
            S = z.shape[1]
            if g_rgb is None and g_depth is None and g_opac is None:
                grads += [torch.zeros_like(p) for p in prm]
                continue
            with torch.enable_grad():
                sg = sig.view(n, S).detach().requires_grad_(True)
                cl = col.view(n, S, 3).detach().requires_grad_(True)
                rgb, depth, opac = _composite(sg, cl, z, dnorm, noise if noise_std > 0 else None, noise_std,
                                              cfg["white_back"])
                outs, gs_ = [], []
                for t, g in ((rgb, g_rgb), (depth, g_depth), (opac, g_opac)):
                    if g is not None:
                        outs.append(t)
                        gs_.append(g)
                dsg, dcl = torch.autograd.grad(outs, [sg, cl], gs_, allow_unused=True)
            dsg = torch.zeros_like(sg) if dsg is None else dsg
            dcl = torch.zeros_like(cl) if dcl is None else dcl
            dpre = (dcl * cl.detach() * (1 - cl.detach())).reshape(-1, 3)          # sigmoid'
            xyz = (o[:, None, :] + d[:, None, :] * z[:, :, None]).reshape(-1, 3)
            enc = torch.zeros(n * S, 64, dtype=torch.float16, device=rays.device)
            enc[:, :63] = _embed(xyz, 10).half()
            grads += _mlp_backward(list(prm), acts, d_act, enc, dir_enc, S, dsg.reshape(-1), dpre)
        if not ctx.has_fine:
            grads += [None] * 24
        return (None, None, None, None, None, None, *grads)


def render_rays_train(models, rays, N_samples, use_disp, perturb, noise_std, N_importance, white_back,
                      pr, nc, ur, nf) -> Dict[str, torch.Tensor]:
    """Differentiable render_rays (test_time=False) through FusedRenderFunction."""
    cfg = dict(models=list(models), N_samples=int(N_samples), N_importance=int(N_importance),
               use_disp=bool(use_disp), perturb=float(perturb), noise_std=float(noise_std),
               white_back=bool(white_back))
    params = nerf_parameters(models[0]) + (nerf_parameters(models[1]) if N_importance > 0 else
                                           [p for p in nerf_parameters(models[0])])
    outs = FusedRenderFunction.apply(cfg, rays, pr, nc, ur, nf, *params)
    res = {"rgb_coarse": outs[0], "depth_coarse": outs[1], "opacity_coarse": outs[2]}
    if N_importance > 0:
        res.update(rgb_fine=outs[3], depth_fine=outs[4], opacity_fine=outs[5])
    return res
