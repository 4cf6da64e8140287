import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_pl_b200 as nb
from nerf_pl_b200 import _lib
from oracle import nerf_oracle as orc
dev = torch.device("cuda:0")
ms = []
for s in (11, 12):
    net = nb.NeRF(); net.load_state_dict({k: torch.from_numpy(v) for k, v in orc.make_weights(s).items()}); ms.append(net.to(dev))
n = 300
rays = torch.from_numpy(orc.make_rays(n, 12)).to(dev)
g = torch.Generator(device=dev).manual_seed(3)
pr = torch.rand(n, 64, device=dev, generator=g); ur = torch.rand(n, 64, device=dev, generator=g)
lib = _lib.load()
bc, bf = nb.packed_weights(ms[0]), nb.packed_weights(ms[1])
f32 = dict(dtype=torch.float32, device=dev); f16 = dict(dtype=torch.float16, device=dev)
def run(which):
    o = [torch.zeros(n, 3, **f32), torch.zeros(n, **f32), torch.zeros(n, **f32), torch.zeros(n, 3, **f32), torch.zeros(n, **f32), torch.zeros(n, **f32)]
    bufs = dict(act_c=torch.empty(8, n * 64, 256, **f16), act_f=torch.empty(8, n * 128, 256, **f16), d_c=torch.empty(n * 64, 128, **f16),
                d_f=torch.empty(n * 128, 128, **f16), sig_c=torch.empty(n * 64, **f32), sig_f=torch.empty(n * 128, **f32),
                rgb_c=torch.empty(n * 64, 3, **f32), rgb_f=torch.empty(n * 128, 3, **f32))
    p = lambda k: bufs[k].data_ptr() if k in which else None
    a = _lib.RenderArgs(rays=rays.data_ptr(), n_rays=n, ray_stride=8, packed_coarse=bc.data_ptr(), packed_fine=bf.data_ptr(),
                        n_samples=64, n_importance=64, use_disp=0, perturb=1.0, noise_std=0.0, white_back=1, test_time=0,
                        perturb_rand=pr.data_ptr(), u_rand=ur.data_ptr(), rgb_coarse=o[0].data_ptr(), depth_coarse=o[1].data_ptr(),
                        opacity_coarse=o[2].data_ptr(), rgb_fine=o[3].data_ptr(), depth_fine=o[4].data_ptr(), opacity_fine=o[5].data_ptr(),
                        save_act_coarse=p("act_c"), save_act_fine=p("act_f"), save_dir_coarse=p("d_c"), save_dir_fine=p("d_f"),
                        save_sigma_coarse=p("sig_c"), save_sigma_fine=p("sig_f"), save_rgb_coarse=p("rgb_c"), save_rgb_fine=p("rgb_f"))
    rc = lib.nerfb200_render_rays(ctypes.byref(a), None)
    torch.cuda.synchronize()
    assert rc == 0, lib.nerfb200_last_error()
    return o, bufs
ref, _ = run(())
for which in ((), ("sig_c",), ("sig_c", "sig_f", "rgb_c", "rgb_f"), ("d_c", "d_f"), ("act_c",), ("act_f",), ("act_c", "act_f"),
              ("act_c", "act_f", "d_c", "d_f", "sig_c", "sig_f", "rgb_c", "rgb_f")):
    o1, _ = run(which); o2, _ = run(which)
    d_ref = max(float((x - y).abs().max()) for x, y in zip(o1, ref))
    d_rep = max(float((x - y).abs().max()) for x, y in zip(o1, o2))
    print(f"{str(which):70s} vs inference {d_ref:.3e}  run-to-run {d_rep:.3e}")
