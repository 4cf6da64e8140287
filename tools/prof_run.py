"""Tiny driver for ncu captures: a few fused render_rays launches, nothing else.
    python tools/prof_run.py [n_rays] [reps] [test_time 0|1]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import nerf_pl_b200 as nb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tt = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
dev = torch.device("cuda:0")
models = []
for s in (11, 12):
    m = nb.NeRF()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in bench.synthetic_weights(s).items()})
    models.append(m.to(dev).eval())
emb = [nb.Embedding(3, 10), nb.Embedding(3, 4)]
rays = torch.from_numpy(bench.blender_rays(n, 0)).to(dev)
rnd = {"perturb_rand": torch.rand(n, 64, device=dev), "u_rand": torch.rand(n, 64, device=dev)}
with torch.no_grad():
    for _ in range(reps):
        nb.render_rays(models, emb, rays, 64, False, 1.0, 0.0, 64, 32768, True, test_time=tt, randoms=rnd)
torch.cuda.synchronize()
print("done", n, reps, tt)
