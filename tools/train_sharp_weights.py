"""Produce TRAINED ("sharp") NeRF weights for parity tests by running the UNMODIFIED reference's own
training step (models/rendering.py render_rays + losses.py MSELoss + torch.optim.Adam, train.py:103-117)
on CPU against a procedural scene.  Random-init weights exercise the fp16 MLP and the final.dir folding
least; trained weights have larger norms and use the high positional frequencies.

    python tools/train_sharp_weights.py [steps]     # writes tests/golden/sharp_weights.npz (+ loss curve)

The scene is analytic (no dataset on the box): three soft spheres with a position-dependent
high-frequency colour pattern, white background; ground-truth pixel colours come from a 512-sample
quadrature of the analytic field along each ray.  Deterministic: seeds below.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import bench  # noqa: E402
from make_golden import import_reference  # noqa: E402

CENTERS = np.array([[0.0, 0.0, 0.0], [0.9, 0.3, -0.2], [-0.6, -0.7, 0.4]], np.float32)
RADII = np.array([0.8, 0.45, 0.55], np.float32)


def scene(x):
    """density (.., ) and colour (.., 3) of the analytic field at points x (.., 3)."""
    d = np.linalg.norm(x[..., None, :] - CENTERS, axis=-1)                      # (.., 3 spheres)
    sig = (40.0 / (1.0 + np.exp((d - RADII) * 30.0))).sum(-1)
    col = 0.5 + 0.5 * np.stack([np.sin(9.0 * x[..., 0] + 2.0 * x[..., 1]), np.sin(7.0 * x[..., 1] - 3.0 * x[..., 2]),
                                np.cos(8.0 * x[..., 2] + x[..., 0])], -1)
    return sig.astype(np.float32), col.astype(np.float32)


def ground_truth(rays, n=512):
    o, d = rays[:, :3], rays[:, 3:6]
    z = np.linspace(2.0, 6.0, n, dtype=np.float32)
    x = o[:, None, :] + d[:, None, :] * z[None, :, None]
    sig, col = scene(x)
    delta = np.full(n, z[1] - z[0], np.float32)
    alpha = 1 - np.exp(-sig * delta)
    T = np.cumprod(np.concatenate([np.ones_like(alpha[:, :1]), 1 - alpha + 1e-10], -1), -1)[:, :-1]
    w = alpha * T
    return ((w[..., None] * col).sum(1) + 1 - w.sum(1, keepdims=True)).astype(np.float32)


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    torch.set_num_threads(os.cpu_count() or 1)
    Embedding, NeRF, render_rays, _ = import_reference()
    torch.manual_seed(1234)
    models = [NeRF(), NeRF()]
    emb = [Embedding(3, 10), Embedding(3, 4)]
    opt = torch.optim.Adam([p for m in models for p in m.parameters()], lr=5e-4, eps=1e-8)     # opt.py:47-58 defaults
    curve = []
    t0 = time.time()
    for it in range(steps):
        rays = bench.blender_rays(1024, 5000 + it)
        tgt = torch.from_numpy(ground_truth(rays))
        out = render_rays(models, emb, torch.from_numpy(rays), 64, False, 1.0, 0.0, 64, 1024 * 32, True, test_time=False)
        loss = torch.nn.functional.mse_loss(out["rgb_coarse"], tgt) + torch.nn.functional.mse_loss(out["rgb_fine"], tgt)
        opt.zero_grad()
        loss.backward()
        opt.step()
        curve.append(float(loss))
        if it % 20 == 0:
            psnr = -10 * np.log10(float(torch.nn.functional.mse_loss(out["rgb_fine"], tgt)))
            print(f"step {it:4d} loss {float(loss):.5f} psnr_fine {psnr:.2f} dB  ({time.time() - t0:.0f} s)", flush=True)
    store = {"loss_curve": np.array(curve, np.float32), "steps": steps}
    for tag, m in zip(("coarse", "fine"), models):
        for k, v in m.state_dict().items():
            store[f"{tag}.{k}"] = v.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sharp_weights.npz"), **store)
    print("saved; final loss", curve[-1])


if __name__ == "__main__":
    main()
