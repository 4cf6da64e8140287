import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_pl_b200 as nb
from oracle import nerf_oracle as orc
dev = torch.device("cuda:0")
ms = []
for s in (11, 12):
    net = nb.NeRF(); net.load_state_dict({k: torch.from_numpy(v) for k, v in orc.make_weights(s).items()}); ms.append(net.to(dev))
emb = [nb.Embedding(3, 10), nb.Embedding(3, 4)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rays = torch.from_numpy(orc.make_rays(n, 12)).to(dev)
g = torch.Generator(device=dev).manual_seed(3)
rnd = {"perturb_rand": torch.rand(n, 64, device=dev, generator=g), "u_rand": torch.rand(n, 64, device=dev, generator=g)}
keep = {k: v.clone() for k, v in rnd.items()}; r0 = rays.clone()
a = nb.render_rays(ms, emb, rays, 64, False, 1.0, 0.0, 64, 32768, True, randoms=rnd, autograd_impl="fused")
b = nb.render_rays(ms, emb, rays, 64, False, 1.0, 0.0, 64, 32768, True, randoms=rnd, autograd_impl="fused")
torch.cuda.synchronize()
print("inputs intact:", all(torch.equal(rnd[k], keep[k]) for k in rnd), torch.equal(rays, r0))
print("fused vs fused:", {k: float((a[k].detach() - b[k].detach()).abs().max()) for k in a})
