import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_pl_b200 as nb
from oracle import nerf_oracle as orc
dev = torch.device("cuda:0")
ws = [orc.make_weights(11), orc.make_weights(12)]
def build():
    out = []
    for w in ws:
        net = nb.NeRF(); net.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}); out.append(net.to(dev))
    return out
emb = [nb.Embedding(3, 10), nb.Embedding(3, 4)]
n = 256
rays = torch.from_numpy(orc.make_rays(n, 12)).to(dev)
g = torch.Generator(device=dev).manual_seed(3)
rnd = {"perturb_rand": torch.rand(n, 64, device=dev, generator=g), "u_rand": torch.rand(n, 64, device=dev, generator=g)}
m = build()
with torch.no_grad():
    inf = nb.render_rays(m, emb, rays, 64, False, 1.0, 0.0, 64, 32768, True, randoms=rnd, extras=True)
fused = nb.render_rays(m, emb, rays, 64, False, 1.0, 0.0, 64, 32768, True, randoms=rnd, autograd_impl="fused")
tor = nb.render_rays(m, emb, rays, 64, False, 1.0, 0.0, 64, 32768, True, randoms=rnd, autograd_impl="torch")
for k in tor:
    a, b, c = inf[k], fused[k].detach(), tor[k].detach()
    d1 = (a - b).abs().flatten(); d2 = (a - c).abs().flatten()
    print(f"{k:15s} inf-vs-fused max {float(d1.max()):.3e} | inf-vs-torch max {float(d2.max()):.3e} q98 {float(torch.quantile(d2, 0.98)):.3e} mean {float(d2.mean()):.3e}")
with torch.no_grad():
    inf2 = nb.render_rays(m, emb, rays, 64, False, 1.0, 0.0, 64, 32768, True, randoms=rnd, extras=True)
fused2 = nb.render_rays(m, emb, rays, 64, False, 1.0, 0.0, 64, 32768, True, randoms=rnd, autograd_impl="fused")
for k in tor:
    print(f"{k:15s} inf-vs-inf2 {float((inf[k]-inf2[k]).abs().max()):.3e} fused-vs-fused2 {float((fused[k].detach()-fused2[k].detach()).abs().max()):.3e}")
bad = ((inf['rgb_fine'] - fused['rgb_fine'].detach()).abs().max(1).values > 1e-4).nonzero().flatten()
print("rays with rgb_fine mismatch:", bad.tolist()[:40], "count", len(bad))
badc = ((inf['rgb_coarse'] - fused['rgb_coarse'].detach()).abs().max(1).values > 1e-4).nonzero().flatten()
print("rays with rgb_coarse mismatch:", badc.tolist()[:40], "count", len(badc))
i = int((inf['depth_coarse'] - tor['depth_coarse'].detach()).abs().argmax())
print("worst ray", i, "depth inf", float(inf['depth_coarse'][i]), "torch", float(tor['depth_coarse'][i]), "opac", float(inf['opacity_coarse'][i]), float(tor['opacity_coarse'][i]))
