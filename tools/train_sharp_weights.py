"""Train NeRF weights on a procedural scene with THIS repository's fused training step (forward + loss +
hand-written sm_100a backward and Adam update) and save them for parity tests (run on a B200 via gpurun).

    python tools/train_sharp_weights.py [steps] [out.npz]

Two purposes: (1) random-init weights exercise the fp16 MLP and the final.dir folding least — trained
weights have larger norms and use the high positional frequencies; tests/golden/make_golden.py runs
the UNMODIFIED reference on the saved weights and the GPU tests hold the kernels to those outputs;
(2) evidence that the loss goes down through the fused backward (the curve is stored with the weights).

The scene is analytic (no dataset on the box): three soft spheres with a position-dependent
high-frequency colour pattern, white background; ground-truth pixel colours come from a 512-sample
quadrature of the analytic field along each ray.  The recipe is the reference's (README.md:75-83,
104-111: 64+64 samples, perturb 1, batch 1024, Adam lr 5e-4; train.py:103-117) with the LLFF setting
noise_std 1: with noise_std 0 a default-initialised coarse network whose raw sigma starts negative
everywhere never receives a gradient (relu'(sigma) = 0, models/rendering.py:155 - the known dead-density
start of NeRF; a first run of this tool showed exactly that: mse_coarse stuck at 0.158 = white image).
Deterministic seeds below.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nerf_pl_b200 as nb  # noqa: E402

CENTERS = torch.tensor([[0.0, 0.0, 0.0], [0.9, 0.3, -0.2], [-0.6, -0.7, 0.4]])
RADII = torch.tensor([0.8, 0.45, 0.55])


def ground_truth(rays, n=512):
    """White-background colour of the analytic scene along each ray (n-sample quadrature on [2, 6])."""
    dev = rays.device
    o, d = rays[:, :3], rays[:, 3:6]
    z = torch.linspace(2.0, 6.0, n, device=dev)
    x = o[:, None, :] + d[:, None, :] * z[None, :, None]
    dist = (x[:, :, None, :] - CENTERS.to(dev)).norm(dim=-1)
    sig = (40.0 / (1.0 + torch.exp((dist - RADII.to(dev)) * 30.0))).sum(-1)
    col = 0.5 + 0.5 * torch.stack([torch.sin(9.0 * x[..., 0] + 2.0 * x[..., 1]), torch.sin(7.0 * x[..., 1] - 3.0 * x[..., 2]),
                                   torch.cos(8.0 * x[..., 2] + x[..., 0])], -1)
    alpha = 1 - torch.exp(-sig * (z[1] - z[0]))
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1 - alpha + 1e-10], -1), -1)[:, :-1]
    w = alpha * T
    return (w[..., None] * col).sum(1) + 1 - w.sum(1, keepdim=True)


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "sharp_weights.npz")
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    models = [nb.NeRF().to(dev), nb.NeRF().to(dev)]
    emb = [nb.Embedding(3, 10), nb.Embedding(3, 4)]
    opt = nb.FusedAdam([p for m in models for p in m.parameters()], lr=5e-4, eps=1e-8)
    n_views = 64
    views = [torch.from_numpy(bench.blender_rays(16384, 7000 + v)).to(dev) for v in range(n_views)]
    targets = [ground_truth(v) for v in views]
    gen = torch.Generator(device=dev).manual_seed(99)
    curve = []
    t0 = time.time()
    for it in range(steps):
        v = it % n_views
        idx = torch.randint(0, views[v].shape[0], (1024,), device=dev, generator=gen)
        out = nb.render_rays_loss(models, emb, views[v][idx], targets[v][idx], 64, False, 1.0, 1.0, 64, 32768, True,
                                  match_reference_rng=False)
        opt.zero_grad(set_to_none=True)
        out["loss"].backward()
        opt.step()
        if it % 50 == 0 or it == steps - 1:
            curve.append((it, float(out["loss"].detach()), float(out["psnr"]), float(out["mse_coarse"])))
            if it % 500 == 0 or it == steps - 1:
                print(f"step {it:5d} loss {curve[-1][1]:.5f} psnr_fine {curve[-1][2]:.2f} dB mse_coarse {curve[-1][3]:.5f} "
                      f"({time.time() - t0:.1f} s)", flush=True)
    torch.cuda.synchronize()
    print(f"{steps} steps in {time.time() - t0:.1f} s")
    # held-out view through the inference path
    with torch.no_grad():
        hv = torch.from_numpy(bench.blender_rays(16384, 9999)).to(dev)
        res = nb.render_rays(models, emb, hv, 64, False, 0, 0, 64, 32768, True, test_time=True)
        gt = ground_truth(hv)
        mse = float(((res["rgb_fine"] - gt) ** 2).mean())
        res_c = nb.render_rays(models, emb, hv, 64, False, 0, 0, 64, 32768, True, test_time=False)
        mse_c = float(((res_c["rgb_coarse"] - gt) ** 2).mean())
    print(f"held-out view psnr fine {-10 * np.log10(mse):.2f} dB, coarse {-10 * np.log10(mse_c):.2f} dB")
    store = {"loss_curve": np.array(curve, np.float32), "steps": steps, "heldout_psnr": np.float32(-10 * np.log10(mse))}
    for tag, m in zip(("coarse", "fine"), models):
        for k, p in m.state_dict().items():
            store[f"{tag}.{k}"] = p.detach().cpu().numpy()
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.savez_compressed(out_path, **store)
    print("saved", out_path)


if __name__ == "__main__":
    main()
