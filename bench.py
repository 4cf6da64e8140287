"""bench.py — ray-samples/s of the render_rays hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1]): Blender-lego-shaped 400x400 pinhole rays, N_samples=64,
N_importance=64, batch_size=1024 rays per GPU per step, training-mode forward
(perturb=1, noise_std=0, white_back, coarse rgb computed) = 1024 x (64 + 128) = 196,608 MLP
evaluations per GPU per step.  A "step" is one render_rays pass over one such batch.
Synthetic rays / random-init (pseudo-trained) weights: no dataset or checkpoint on the box.

One JSON line on stdout (rank 0):  value = device-timed whole-job ray-samples/s with the batch
resident in HBM; e2e = same through the public Python API with pinned HOST buffers (H2D of the
rays and D2H of rgb_fine inside the timed region); roofline = the fused kernel against the
measured dense tensor peak; cpu_baseline = the numpy oracle timed on the host cores.
`--impl reference` times the reference algorithm's CPU restatement (oracle/) on the host cores —
the reference itself (PyTorch + torchsearchsorted) cannot travel to the GPU box.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SAMPLES, N_IMPORTANCE = 64, 64
BATCH = 1024
SAMPLES_PER_RAY = N_SAMPLES + (N_SAMPLES + N_IMPORTANCE)          # 192
FLOP_PER_RAY_TRAIN = 192 * 1186816                               # test_time=False  
IMG_W = IMG_H = 400
CAMERA_ANGLE_X = 0.6911112070083618                              # lego transforms_*.json


def blender_rays(n, seed):
    """n random rays of a 400x400 Blender-style pinhole camera on a radius-4 sphere looking at the
    origin: unit directions, near=2, far=6 (reference datasets/ray_utils.py:5-43, blender.py:28-35)."""
    rs = np.random.RandomState(seed)
    focal = 0.5 * IMG_W / np.tan(0.5 * CAMERA_ANGLE_X)
    th, ph = rs.uniform(0, 2 * np.pi), rs.uniform(np.pi / 6, np.pi / 3)
    cam = 4.0 * np.array([np.cos(th) * np.sin(ph), np.sin(th) * np.sin(ph), np.cos(ph)])
    fwd = -cam / np.linalg.norm(cam)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0])); right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    px = rs.randint(0, IMG_W, n); py = rs.randint(0, IMG_H, n)
    dc = np.stack([(px - IMG_W / 2) / focal, -(py - IMG_H / 2) / focal, -np.ones(n)], -1)
    d = dc[:, :1] * right + dc[:, 1:2] * up - dc[:, 2:3] * fwd
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    o = np.broadcast_to(cam, (n, 3))
    return np.concatenate([o, d, np.full((n, 1), 2.0), np.full((n, 1), 6.0)], -1).astype(np.float32)


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region, in-process through NVML
    (nvidia-smi's own start-up is longer than a short timed region), every ~2 ms."""

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.stop_flag, self.thread, self.err = gpu_index, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES if it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:          # noqa: BLE001
            self.err = repr(e)
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, rs))
            except Exception as e:      # noqa: BLE001
                self.err = repr(e)
                return
            time.sleep(0.002)

    def stop(self):
        self.stop_flag = True
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + str(self.err)]}
        self.thread.join(timeout=1)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(k for k, bit in names.items() if any(r & bit for _, r in self.rows))
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.max_sm),
                "reasons": reasons, "samples": len(sm)}


_NERF_LAYERS = ([("xyz_encoding_1.0", 256, 63)] + [(f"xyz_encoding_{i}.0", 256, 256) for i in (2, 3, 4)]
                + [("xyz_encoding_5.0", 256, 319)] + [(f"xyz_encoding_{i}.0", 256, 256) for i in (6, 7, 8)]
                + [("xyz_encoding_final", 256, 256), ("dir_encoding.0", 128, 283), ("sigma", 1, 256), ("rgb.0", 3, 128)])


def synthetic_weights(seed):
    """Random-init NeRF state_dict (numpy): nn.Linear's default U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for the
    reference architecture (models/nerf.py:58-81), with the sigma / rgb heads scaled so that
    opacities and colours are not degenerate (there are no checkpoints to load).  bench.py's own
    generator: the b200 arm does not touch oracle/."""
    rs = np.random.RandomState(seed)
    w = {}
    for name, out_f, in_f in _NERF_LAYERS:
        bound = 1.0 / np.sqrt(in_f)
        w[name + ".weight"] = rs.uniform(-bound, bound, size=(out_f, in_f)).astype(np.float32)
        w[name + ".bias"] = rs.uniform(-bound, bound, size=(out_f,)).astype(np.float32)
    w["sigma.weight"] = (w["sigma.weight"] * np.float32(30.0)).astype(np.float32)
    w["sigma.bias"] = (w["sigma.bias"] + np.float32(0.5)).astype(np.float32)
    w["rgb.0.weight"] = (w["rgb.0.weight"] * np.float32(8.0)).astype(np.float32)
    return w


def ncu_traffic():
    """DRAM bytes (read + write) of one bench-shaped render_rays launch from the committed ncu
    --set full capture (profiles/*_ncu_traffic.json, written by tools/summarize_ncu.py); None if absent."""
    import glob
    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "*_ncu_traffic.json")))
    if not files:
        return None
    try:
        return float(json.load(open(files[-1]))["traffic_bytes_per_launch"])
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst cuBLAS 8192^3)"
    return 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"


_BEST_THREADS = None


def best_blas_threads():
    """The numpy/BLAS thread count that renders fastest on this host (all cores is not always
    best: on a 128-core box the 256-wide GEMMs oversubscribe).  Probed once, ~2 s."""
    global _BEST_THREADS
    if _BEST_THREADS is not None:
        return _BEST_THREADS
    cores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        _BEST_THREADS = cores
        return cores
    from oracle import nerf_oracle as orc
    ws = [synthetic_weights(11), synthetic_weights(12)]
    rays = blender_rays(128, 0)
    best, best_t = cores, float("inf")
    cands = sorted({c for c in (4, 8, 16, 32, 64, cores) if c <= cores})
    for c in cands:
        with threadpool_limits(limits=c):
            orc.render_rays(ws, rays[:16], N_SAMPLES, False, 0.0, 0.0, N_IMPORTANCE, True, False)
            t0 = time.perf_counter()
            orc.render_rays(ws, rays, N_SAMPLES, False, 0.0, 0.0, N_IMPORTANCE, True, False)
            t = time.perf_counter() - t0
        if t < best_t:
            best, best_t = c, t
    _BEST_THREADS = best
    return best


def cpu_oracle_throughput(n_rays, reps, seed=0):
    """ray-samples/s of the oracle on the host cores over `reps` batches of n_rays (same workload),
    with the BLAS thread count that is fastest on this host."""
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=best_blas_threads()):
            return _cpu_oracle_throughput(n_rays, reps, seed)
    except ImportError:
        return _cpu_oracle_throughput(n_rays, reps, seed)


def _cpu_oracle_throughput(n_rays, reps, seed=0):
    from oracle import nerf_oracle as orc      # the cpu_baseline / reference leg: the oracle is what is timed
    ws = [synthetic_weights(11), synthetic_weights(12)]
    rays = blender_rays(n_rays, seed)
    rs = np.random.RandomState(seed)
    rnd = {"perturb_rand": rs.rand(n_rays, N_SAMPLES).astype(np.float32),
           "u_rand": rs.rand(n_rays, N_IMPORTANCE).astype(np.float32)}
    orc.render_rays(ws, rays[:32], N_SAMPLES, False, 1.0, 0.0, N_IMPORTANCE, True, False,
                    {k: v[:32] for k, v in rnd.items()})       # warm BLAS threads
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        orc.render_rays(ws, rays, N_SAMPLES, False, 1.0, 0.0, N_IMPORTANCE, True, False, rnd)
        times.append(time.perf_counter() - t0)
    return n_rays * SAMPLES_PER_RAY / float(np.median(times)), float(np.median(times))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = best_blas_threads()
    n_rays = 256                      # bounded sample of the 1024-ray batch (same per-ray work)
    for _ in range(max(args.warmup, 1)):
        cpu_oracle_throughput(n_rays, 1)
    vals = [cpu_oracle_throughput(n_rays, 1, seed=i)[0] for i in range(args.steps)]
    v = float(np.median(vals))
    ms = n_rays * SAMPLES_PER_RAY / v * 1e3
    sample = (f"{n_rays} rays of the 1024-ray batch per step (64+128 samples/ray), numpy/BLAS on {cores} threads "
              f"(fastest of 4..{os.cpu_count()} on this host)")
    emit(({
        "impl": "reference", "metric": "ray-samples/sec (coarse+fine)", "value": v, "unit": "ray-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": v, "unit": "ray-samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "ray-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(n_gpus):
    return {"workload": "blender_lego_400x400 N_samples=64 N_importance=64 batch_size=1024/GPU, "
                        "render_rays training-mode forward (perturb=1, noise_std=0, white_back, coarse rgb)",
            "rays_per_step_per_gpu": BATCH, "global_batch": BATCH * n_gpus, "samples_per_ray": SAMPLES_PER_RAY,
            "parallelism": f"ray-sharded dp{n_gpus}, weights replicated, no data-path collective" if n_gpus > 1 else "single GPU",
            "l2": "flushed between timed steps (256 MiB write)"}


def run_b200(args):
    import torch
    import torch.distributed as dist

    import nerf_pl_b200 as nb
    from nerf_pl_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there at VERSION level
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    ws = [synthetic_weights(11), synthetic_weights(12)]   # random init: there are no checkpoints
    models = []
    for w in ws:
        m = nb.NeRF()
        m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
        models.append(m.to(dev).eval())
    emb = [nb.Embedding(3, 10), nb.Embedding(3, 4)]
    n_batches = 8
    host_rays = [torch.from_numpy(blender_rays(BATCH, 100 * rank + i)).pin_memory() for i in range(n_batches)]
    dev_rays = [r.to(dev) for r in host_rays]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    def step(rays):
        out = nb.render_rays(models, emb, rays, N_SAMPLES, False, 1.0, 0.0, N_IMPORTANCE, 1024 * 32, True,
                             test_time=False)
        return out       # rays are independent: no collective on the path

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(args.warmup):
            step(dev_rays[i % n_batches])
        barrier()
        # ---- device-timed steps (inputs resident in HBM); L2 flushed between steps, not timed
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        l0 = lib.nerfb200_launch_count()
        barrier()
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            ev[i][0].record()
            step(dev_rays[i % n_batches])
            ev[i][1].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
        launches = lib.nerfb200_launch_count() - l0
        step_ms = [a.elapsed_time(b) for a, b in ev]
        total_ms = float(sum(step_ms))

        # ---- kernel-only timing for the roofline (pre-drawn randoms: the only kernel between the
        # events is the fused render kernel, on torch's current stream, which is the launch stream)
        rnd = {"perturb_rand": torch.rand(BATCH, N_SAMPLES, device=dev),
               "u_rand": torch.rand(BATCH, N_IMPORTANCE, device=dev)}
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            kev[i][0].record()
            nb.render_rays(models, emb, dev_rays[i % n_batches], N_SAMPLES, False, 1.0, 0.0, N_IMPORTANCE,
                           1024 * 32, True, test_time=False, randoms=rnd)
            kev[i][1].record()
        torch.cuda.synchronize()
        kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))

        # ---- end to end through the public API with HOST buffers
        barrier()
        eev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        host_out = torch.empty(BATCH, 3).pin_memory()
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            eev[i][0].record()
            r = host_rays[i % n_batches].to(dev, non_blocking=True)
            out = step(r)
            host_out.copy_(out["rgb_fine"], non_blocking=True)
            eev[i][1].record()
            eev[i][1].synchronize()          # the result is read on the host every step
        barrier()
        e2e_ms = float(sum(a.elapsed_time(b) for a, b in eev))
        clocks = sampler.stop() if rank == 0 else None

    # max over ranks
    t = torch.tensor([total_ms, e2e_ms, kern_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kern_ms = [float(x) for x in t.tolist()]
    if rank == 0:
        samples = BATCH * SAMPLES_PER_RAY * world * args.steps
        value = samples / (total_ms * 1e-3)
        peak, peak_src = measured_peaks()
        ach = BATCH * FLOP_PER_RAY_TRAIN / (kern_ms * 1e-3) / 1e12
        cpu_v, cpu_t = cpu_oracle_throughput(256, 3)
        cores = best_blas_threads()
        line = {
            "metric": "ray-samples/sec (coarse+fine)", "value": value, "unit": "ray-samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 (fp32 accumulate, "
            "fp32 encoding/compositing)", "data": "synthetic", "config": workload_config(world),
            "clocks": clocks,
            "e2e": {"value": samples / (e2e_ms * 1e-3), "unit": "ray-samples/s",
                    "h2d_bytes_per_step": BATCH * 8 * 4, "d2h_bytes_per_step": BATCH * 3 * 4},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                         "frac": ach / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                         "kernel": "render_rays_kernel", "kernel_ms": kern_ms,
                         "flop_per_launch": BATCH * FLOP_PER_RAY_TRAIN},
            "cpu_baseline": {"value": cpu_v, "unit": "ray-samples/s", "cores": cores, "kind": "port",
                             "sample": f"256 rays of the same batch, 3 reps, median {cpu_t:.2f} s, numpy/BLAS on {cores} of "
                                       f"{os.cpu_count()} host threads (fastest setting)"},
            "wall_s_timed_region": t_wall,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def emit(obj):
    """The ONE JSON line goes to the process's original stdout; everything else that libraries
    print on fd 1 (e.g. NCCL's version banner) has been redirected to stderr by main()."""
    data = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
