"""Where does the fused training step spend its time?  torch profiler table of one step."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import nerf_pl_b200 as nb  # noqa: E402

dev = torch.device("cuda:0")
models = []
for s in (11, 12):
    m = nb.NeRF()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in bench.synthetic_weights(s).items()})
    models.append(m.to(dev))
emb = [nb.Embedding(3, 10), nb.Embedding(3, 4)]
rays = torch.from_numpy(bench.blender_rays(1024, 0)).to(dev)
tgt = torch.rand(1024, 3, device=dev)


def step():
    for m in models:
        m.zero_grad(set_to_none=True)
    out = nb.render_rays(models, emb, rays, 64, False, 1.0, 0.0, 64, 32768, True)
    loss = ((out["rgb_coarse"] - tgt) ** 2).mean() + ((out["rgb_fine"] - tgt) ** 2).mean()
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
