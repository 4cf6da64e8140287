"""Host-side cost of the public render_rays call (bench.py's step), with cProfile.
    python tools/wrapper_profile.py [steps]"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import nerf_pl_b200 as nb  # noqa: E402
from oracle import nerf_oracle as orc  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = torch.device("cuda:0")
models = []
for s in (11, 12):
    m = nb.NeRF()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in orc.make_weights(s).items()})
    models.append(m.to(dev).eval())
emb = [nb.Embedding(3, 10), nb.Embedding(3, 4)]
rays = torch.from_numpy(bench.blender_rays(1024, 0)).to(dev)


def step():
    return nb.render_rays(models, emb, rays, 64, False, 1.0, 0.0, 64, 1024 * 32, True, test_time=False)


with torch.no_grad():
    for _ in range(20):
        step()
    torch.cuda.synchronize()
    # pure host cost: the GPU is given nothing else, so the queue never blocks the host
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"host time per call {t_host / steps * 1e6:.1f} us; incl. GPU drain {t_all / steps * 1e6:.1f} us")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(steps):
        step()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(18)
