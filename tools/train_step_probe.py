"""Time one training step (train.py:103-117 shape: 1024 rays, 64+64, perturb=1, noise_std=0,
white_back; loss = MSE(rgb_coarse)+MSE(rgb_fine); Adam) through the drop-in's autograd path."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import nerf_pl_b200 as nb  # noqa: E402
from oracle import nerf_oracle as orc  # noqa: E402

dev = torch.device("cuda:0")
models = []
for s in (11, 12):
    m = nb.NeRF()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in orc.make_weights(s).items()})
    models.append(m.to(dev))
emb = [nb.Embedding(3, 10), nb.Embedding(3, 4)]
opt = torch.optim.Adam([p for m in models for p in m.parameters()], lr=5e-4)
rays = torch.from_numpy(bench.blender_rays(1024, 0)).to(dev)
tgt = torch.rand(1024, 3, device=dev)


IMPL = sys.argv[1] if len(sys.argv) > 1 else "fused"


def step():
    opt.zero_grad(set_to_none=True)
    out = nb.render_rays(models, emb, rays, 64, False, 1.0, 0.0, 64, 32768, True, autograd_impl=IMPL)
    loss = ((out["rgb_coarse"] - tgt) ** 2).mean() + ((out["rgb_fine"] - tgt) ** 2).mean()
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 10
for _ in range(n):
    l = step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print(f"train step [{IMPL}] fwd+bwd+Adam: {dt * 1e3:.2f} ms/step, "
      f"{1024 * 192 / dt:.3e} ray-samples/s, loss {float(l):.4f}")
