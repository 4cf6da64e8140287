"""ncu / timing driver for the training step (fused forward + loss, sm_100a backward, Adam):
    python tools/prof_train.py [n_rays] [steps] [time|plain]
'time' prints CUDA-event time per step and the host-side time of the pieces."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import nerf_pl_b200 as nb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else "plain"
dev = torch.device("cuda:0")
models = []
for s in (11, 12):
    m = nb.NeRF()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in bench.synthetic_weights(s).items()})
    models.append(m.to(dev))
emb = [nb.Embedding(3, 10), nb.Embedding(3, 4)]
rays = torch.from_numpy(bench.blender_rays(n, 0)).to(dev)
tgt = torch.rand(n, 3, device=dev)
opt = nb.FusedAdam([p for m in models for p in m.parameters()], lr=5e-4)


def step():
    opt.zero_grad(set_to_none=True)
    out = nb.render_rays_loss(models, emb, rays, tgt, 64, False, 1.0, 0.0, 64, 32768, True, match_reference_rng=False)
    out["loss"].backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
if mode == "time":
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"train step: device {e0.elapsed_time(e1) / steps:.3f} ms/step, host enqueue {t_host / steps * 1e3:.3f} ms/step")
    # forward-only pieces on the host
    with torch.no_grad():
        for m in models:
            m.requires_grad_(False)
        t0 = time.perf_counter()
        for _ in range(200):
            nb.render_rays(models, emb, rays, 64, False, 1.0, 0.0, 64, 32768, True, match_reference_rng=False)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"render_rays host-side: {(t1 - t0) / 200 * 1e6:.1f} us per call")
        pin = rays.cpu().pin_memory()
        hout = torch.empty(n, 10).pin_memory()
        for sync_each in (False, True):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(200):
                r = pin.to(dev, non_blocking=True)
                out = nb.render_rays(models, emb, r, 64, False, 1.0, 0.0, 64, 32768, True, match_reference_rng=False)
                flat = torch.cat((out["rgb_coarse"], out["depth_coarse"][:, None], out["opacity_coarse"][:, None],
                                  out["rgb_fine"], out["depth_fine"][:, None], out["opacity_fine"][:, None]), 1)
                hout.copy_(flat, non_blocking=True)
                if sync_each:
                    torch.cuda.current_stream().synchronize()
            torch.cuda.synchronize()
            print(f"e2e loop sync_each={sync_each}: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us per step")
        # the same with a ClockSampler thread polling NVML (what bench.py does while it times)
        cs = bench.ClockSampler(0)
        cs.start()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            r = pin.to(dev, non_blocking=True)
            out = nb.render_rays(models, emb, r, 64, False, 1.0, 0.0, 64, 32768, True, match_reference_rng=False)
            flat = torch.cat((out["rgb_coarse"], out["depth_coarse"][:, None], out["opacity_coarse"][:, None],
                              out["rgb_fine"], out["depth_fine"][:, None], out["opacity_fine"][:, None]), 1)
            hout.copy_(flat, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        print(f"e2e loop with NVML sampler thread: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us per step", cs.stop())
else:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
print("done")
