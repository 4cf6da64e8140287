"""ncu driver: the bench.py workload as single launches (1024 rays, training-mode forward), each
preceded by the 256 MiB L2 flush, so the --set full capture sees the DRAM traffic of a cold launch.
    ncu --set full -k regex:render_rays -s 2 -c 1 ... python tools/prof_bench1024.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import nerf_pl_b200 as nb  # noqa: E402

dev = torch.device("cuda:0")
models = []
for s in (11, 12):
    m = nb.NeRF()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in bench.synthetic_weights(s).items()})
    models.append(m.to(dev).eval())
emb = [nb.Embedding(3, 10), nb.Embedding(3, 4)]
n = 1024
rays = torch.from_numpy(bench.blender_rays(n, 0)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
with torch.no_grad():
    for i in range(4):
        rnd = {"perturb_rand": torch.rand(n, 64, device=dev), "u_rand": torch.rand(n, 64, device=dev)}
        flush.fill_(i)
        nb.render_rays(models, emb, rays, 64, False, 1.0, 0.0, 64, 32768, True, test_time=False, randoms=rnd)
torch.cuda.synchronize()
print("done")
