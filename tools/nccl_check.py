"""torchrun --nproc-per-node 2 tools/nccl_check.py : ray-sharded render over NCCL == single-GPU render."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import nerf_pl_b200 as nb  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
models = []
for s in (11, 12):
    m = nb.NeRF()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in bench.synthetic_weights(s).items()})
    models.append(m.to(dev).eval())
emb = [nb.Embedding(3, 10), nb.Embedding(3, 4)]
for n in (4001, 80000):
    rays = torch.from_numpy(bench.blender_rays(n, 3)).to(dev)
    with torch.no_grad():
        single = nb.batched_inference(models, emb, rays, 64, 64, False, 32768, True)
        shard = nb.batched_inference(models, emb, rays, 64, 64, False, 32768, True, sharded=True)
    torch.cuda.synchronize()
    ok = all(torch.equal(single[k], shard[k]) for k in single)
    print(f"rank {dist.get_rank()} n={n} sharded==single: {ok}", flush=True)
    assert ok
dist.destroy_process_group()
print("NCCL_CHECK_OK")
