"""CPU oracle for the TRAINING step of the render_rays hot path: loss + hand-derived gradients.

TEST INFRASTRUCTURE ONLY (same rules as oracle/nerf_oracle.py: imported by tests/, smoke() and
bench.py's CPU legs, never by the product).

The reference obtains these gradients from torch autograd (train.py:103-117: ``results =
render_rays(...)``; ``loss = MSELoss(results, rgbs)`` (losses.py:9-14); ``loss.backward()``)
through models/nerf.py:100-124 and models/rendering.py:143-170.  This module restates the same
derivatives by hand in numpy, function by function, in the decomposition the CUDA backward uses
(compositing backward -> per-sample d sigma / d rgb -> MLP backward).  PINNED:
``tests/golden/grad_*.npz`` hold the 48 ``.grad`` tensors the unmodified reference produces
(``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py`` holds this restatement to them.
No gradient flows through the fine-depth sampling (models/rendering.py:225-227 ``.detach()``) nor
into the rays.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from . import nerf_oracle as orc

F32 = np.float32


# ------------------------------------------------------------------ models/nerf.py, forward with tape
def nerf_forward_tape(w: Dict[str, np.ndarray], x: np.ndarray) -> Dict[str, np.ndarray]:
    """NeRF.forward (models/nerf.py:100-124) keeping what the backward needs: the post-ReLU outputs
    h1..h8, `final`, the direction-layer output d, sigmoid(rgb) and raw sigma."""
    x = np.ascontiguousarray(x, dtype=F32)
    enc, dirs = x[:, :63], x[:, 63:90]
    tape = {"enc": enc, "dir": dirs}
    h = enc
    for i in range(8):
        if i == 4:
            h = np.concatenate([enc, h], -1)
        h = np.maximum(h @ w[f"xyz_encoding_{i + 1}.0.weight"].T + w[f"xyz_encoding_{i + 1}.0.bias"], F32(0))
        tape[f"h{i + 1}"] = h
    tape["sigma"] = (h @ w["sigma.weight"].T + w["sigma.bias"])[:, 0]
    final = h @ w["xyz_encoding_final.weight"].T + w["xyz_encoding_final.bias"]
    tape["final"] = final
    d = np.maximum(np.concatenate([final, dirs], -1) @ w["dir_encoding.0.weight"].T + w["dir_encoding.0.bias"], F32(0))
    tape["d"] = d
    pre = d @ w["rgb.0.weight"].T + w["rgb.0.bias"]
    tape["rgb"] = (F32(1) / (F32(1) + np.exp(-pre, dtype=F32))).astype(F32)
    return tape


def nerf_backward(w: Dict[str, np.ndarray], tape: Dict[str, np.ndarray], d_sigma: np.ndarray,
                  d_rgb: np.ndarray) -> Dict[str, np.ndarray]:
    """Derivative of NeRF.forward w.r.t. its 24 parameter tensors given dL/d(raw sigma) (S,) and
    dL/d(sigmoid rgb) (S,3).  Layer by layer in reverse (models/nerf.py:120 -> :100)."""
    g: Dict[str, np.ndarray] = {}
    rgb, d, h8 = tape["rgb"], tape["d"], tape["h8"]
    dpre = (d_rgb * rgb * (F32(1) - rgb)).astype(F32)                       # sigmoid' (models/nerf.py:81)
    g["rgb.0.weight"] = dpre.T @ d
    g["rgb.0.bias"] = dpre.sum(0)
    dd = (dpre @ w["rgb.0.weight"]) * (d > 0)                              # ReLU of dir_encoding (:77)
    cat = np.concatenate([tape["final"], tape["dir"]], -1)
    g["dir_encoding.0.weight"] = dd.T @ cat
    g["dir_encoding.0.bias"] = dd.sum(0)
    dfinal = dd @ w["dir_encoding.0.weight"][:, :256]
    g["xyz_encoding_final.weight"] = dfinal.T @ h8
    g["xyz_encoding_final.bias"] = dfinal.sum(0)
    g["sigma.weight"] = (d_sigma[:, None] * h8).sum(0, keepdims=True)
    g["sigma.bias"] = np.array([d_sigma.sum()], dtype=F32)
    dh = dfinal @ w["xyz_encoding_final.weight"] + d_sigma[:, None] * w["sigma.weight"]   # into h8
    for i in range(7, -1, -1):
        name = f"xyz_encoding_{i + 1}.0"
        dp = dh * (tape[f"h{i + 1}"] > 0)
        inp = tape["enc"] if i == 0 else (np.concatenate([tape["enc"], tape["h4"]], -1) if i == 4 else tape[f"h{i}"])
        g[name + ".weight"] = dp.T @ inp
        g[name + ".bias"] = dp.sum(0)
        if i > 0:
            W = w[name + ".weight"]
            dh = dp @ (W[:, 63:] if i == 4 else W)          # skip layer: only the hidden part carries on
    return {k: np.asarray(v, dtype=F32) for k, v in g.items()}


# ------------------------------------------------------------------ models/rendering.py:143-170, backward
def volume_render_backward(sigmas, rgbs, z_vals, dirs, noise, noise_std, white_back, g_rgb, g_depth=None,
                           g_opac=None) -> Tuple[np.ndarray, np.ndarray]:
    """Given dL/d(rgb (n,3)) [, dL/d(depth), dL/d(opacity)] of the quadrature return
    dL/d(sigmas) (n,S) and dL/d(rgbs) (n,S,3).

      w_i = alpha_i T_i,  T_i = prod_{j<i} (1 - alpha_j + 1e-10),  alpha_i = 1 - exp(-delta_i relu(s_i))
      dL/dw_i   = <g_rgb, c_i> + g_depth z_i + g_opac - [white_back] sum_ch g_rgb
      dL/dalpha_i = T_i dL/dw_i - (sum_{j>i} w_j dL/dw_j) / (1 - alpha_i + 1e-10)
      dalpha_i/ds_i = delta_i exp(-delta_i relu(s_i)) [s_i > 0]
    """
    f8 = np.float64
    n, S = sigmas.shape
    deltas = np.concatenate([z_vals[:, 1:] - z_vals[:, :-1], np.full((n, 1), 1e10, F32)], -1).astype(F32)
    deltas = (deltas * np.linalg.norm(dirs.astype(F32), axis=-1, keepdims=True)).astype(F32)
    s = sigmas.astype(F32) if noise is None else (sigmas + noise.astype(F32) * F32(noise_std)).astype(F32)
    e = np.exp(-deltas * np.maximum(s, F32(0)), dtype=F32)
    alpha = (F32(1) - e).astype(F32)
    om = (F32(1) - alpha + F32(1e-10)).astype(F32)
    T = np.cumprod(np.concatenate([np.ones((n, 1), F32), om], -1), -1, dtype=F32)[:, :-1]
    wts = (alpha * T).astype(F32)
    dw = (rgbs.astype(f8) * g_rgb[:, None, :].astype(f8)).sum(-1)
    if g_depth is not None:
        dw = dw + g_depth[:, None].astype(f8) * z_vals
    if g_opac is not None:
        dw = dw + g_opac[:, None].astype(f8)
    if white_back:
        dw = dw - g_rgb.astype(f8).sum(-1, keepdims=True)
    a = wts.astype(f8) * dw
    suffix = np.cumsum(a[:, ::-1], -1)[:, ::-1] - a                        # sum_{j>i} w_j dL/dw_j
    dalpha = T.astype(f8) * dw - suffix / om.astype(f8)
    dsig = dalpha * (deltas.astype(f8) * e.astype(f8)) * (s > 0)
    drgbs = wts[..., None].astype(f8) * g_rgb[:, None, :].astype(f8)
    return dsig.astype(F32), drgbs.astype(F32)


# ------------------------------------------------------------------ one training step
def render_rays_loss_grad(weights: List[Dict[str, np.ndarray]], rays: np.ndarray, target: np.ndarray,
                          N_samples=64, use_disp=False, perturb=1.0, noise_std=0.0, N_importance=64,
                          white_back=True, randoms: Optional[Dict[str, np.ndarray]] = None):
    """loss = mean((rgb_coarse - t)^2) + mean((rgb_fine - t)^2) (losses.py:9-14) through
    render_rays(test_time=False) and its gradients w.r.t. the parameters of both networks.
    Returns (loss, results dict, {'coarse.<key>': grad, 'fine.<key>': grad})."""
    randoms = randoms or {}
    res = orc.render_rays(weights, rays, N_samples, use_disp, perturb, noise_std, N_importance, white_back,
                          False, randoms, extras=True)
    rays = rays.astype(F32)
    o, d = rays[:, 0:3], rays[:, 3:6]
    n = rays.shape[0]
    dir_emb = orc.embed(d, 4)
    z_c = orc.coarse_depths(rays, N_samples, use_disp, perturb, randoms.get("perturb_rand"))
    passes = [("coarse", weights[0], z_c, randoms.get("noise_coarse"))]
    if N_importance > 0:
        passes.append(("fine", weights[1], res["z_vals_fine"], randoms.get("noise_fine")))
    loss = 0.0
    grads: Dict[str, np.ndarray] = {}
    for tag, w, z, noise in passes:
        S = z.shape[1]
        diff = (res[f"rgb_{tag}"] - target).astype(np.float64)
        loss += float((diff ** 2).mean())
        g_rgb = (2.0 * diff / diff.size).astype(F32)                      # d mean((rgb - t)^2) / d rgb
        xyz = (o[:, None, :] + d[:, None, :] * z[:, :, None]).astype(F32).reshape(-1, 3)
        x = np.concatenate([orc.embed(xyz, 10), np.repeat(dir_emb, S, axis=0)], -1)
        tape = nerf_forward_tape(w, x)
        dsig, drgbs = volume_render_backward(tape["sigma"].reshape(n, S), tape["rgb"].reshape(n, S, 3), z, d,
                                             noise if noise_std > 0 else None, noise_std, white_back, g_rgb)
        g = nerf_backward(w, tape, dsig.reshape(-1), drgbs.reshape(-1, 3))
        for k, v in g.items():
            grads[f"{tag}.{k}"] = v
    return loss, res, grads


def unpack_golden_grads(npz) -> Dict[str, np.ndarray]:
    """Inverse of tests/golden/make_golden.py:pack_grads."""
    out = {}
    for k in npz.files:
        if k.startswith("g16_"):
            key = k[4:]
            out[key] = npz[k].astype(F32) * F32(npz["gscale_" + key])
    return out


def grad_compare(a: Dict[str, np.ndarray], b: Dict[str, np.ndarray]):
    """Per-tensor (relative L2 error, cosine) and the global pair over all tensors."""
    rows = {}
    num = den = dot = na = nb_ = 0.0
    for k in sorted(b):
        x, y = a[k].astype(np.float64).ravel(), b[k].astype(np.float64).ravel()
        e, ny, nx = float(np.linalg.norm(x - y)), float(np.linalg.norm(y)), float(np.linalg.norm(x))
        rows[k] = (e / max(ny, 1e-30), float(x @ y) / max(nx * ny, 1e-30))
        num += e ** 2; den += ny ** 2; dot += float(x @ y); na += nx ** 2; nb_ += ny ** 2
    return rows, (float(np.sqrt(num / max(den, 1e-60))), dot / max(np.sqrt(na * nb_), 1e-30))
