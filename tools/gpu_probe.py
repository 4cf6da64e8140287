"""Bring-up probe for the tcgen05 engine (run on a B200 via gpurun; each step in its own process
so a device trap in one step does not hide the others).

    python tools/gpu_probe.py gemm | mlp | render | speed | speed1 | mmabench | contention | issue | cache
    python tools/gpu_probe.py timeline | tlsum        (needs a -DNERFB200_TIMELINE build: NERFB200_LIB=...)
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_pl_b200 as nb  # noqa: E402
from nerf_pl_b200 import _lib  # noqa: E402
from nerf_pl_b200.rendering import _render_with_graph  # noqa: E402


def make_models(seed=0, scale_heads=True):
    torch.manual_seed(seed)
    ms = [nb.NeRF().cuda(), nb.NeRF().cuda()]
    if scale_heads:
        with torch.no_grad():
            for m in ms:
                m.sigma.weight.mul_(30.0)
                m.sigma.bias.add_(0.5)
                m.rgb[0].weight.mul_(8.0)
    return ms


def make_rays(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    o = torch.tensor([0.0, 0.0, 4.0]) + 0.1 * torch.randn(n, 3, generator=g)
    d = torch.randn(n, 3, generator=g)
    d[:, 2] = -d[:, 2].abs() - 1.0
    d = d / d.norm(dim=-1, keepdim=True)
    near = torch.full((n, 1), 2.0)
    far = torch.full((n, 1), 6.0)
    return torch.cat([o, d, near, far], -1).cuda()


def step_gemm():
    lib = _lib.load()
    m = make_models()[0]
    blob = nb.packed_weights(m)
    torch.cuda.synchronize()
    params = nb.nerf_parameters(m)
    a = torch.randn(128, 64, device="cuda")
    a16 = a.half().float()
    # slice -> (weight, k offset, valid k, N)   (csrc/layout.h)
    cases = {
        0: (params[0], 0, 63, 256), 5: (params[4], 0, 64, 256), 8: (params[4], 192, 64, 256),
        13: (params[8], 0, 63, 256), 15: (params[8], 63 + 64, 64, 256),
        31: ((params[18][:, :256].double() @ params[16].double()).float(), 64, 64, 128),   # fused W' k-block 1
        34: (params[18], 256, 27, 128),
    }
    ok = True
    for mode in (0, 1):
        for sl, (W, koff, kval, N) in cases.items():
            d = torch.zeros(128, N, device="cuda")
            rc = lib.nerfb200_debug_gemm(a.data_ptr(), blob.data_ptr(), sl, mode, d.data_ptr(), None)
            torch.cuda.synchronize()
            Wk = torch.zeros(N, 64, device="cuda")
            Wk[:, :kval] = W[:, koff:koff + kval].detach().half().float()
            ref = a16 @ Wk.t()
            err = (d - ref).abs().max().item()
            print(f"gemm mode {mode} slice {sl:2d} N={N} rc={rc} max_abs_err={err:.3e} ref_absmax={ref.abs().max().item():.3f}")
            ok &= err < 1e-3
    print("GEMM_OK" if ok else "GEMM_FAIL")


def step_gemm_mn():
    """MN-major descriptor probe (the wgrad operand layout): sweep (LBO, SBO) candidates."""
    lib = _lib.load()
    torch.manual_seed(3)
    a = torch.randn(64, 128, device="cuda")
    b = torch.randn(64, 256, device="cuda")
    for fmt in (0, 1, 2, 3):
        ah = a.bfloat16().float() if fmt & 1 else a.half().float()
        bh = b.bfloat16().float() if fmt & 2 else b.half().float()
        ref = ah.t() @ bh
        for lbo, sbo in ((8192, 1024), (1024, 8192)) if fmt == 0 else ((8192, 1024),):
            d = torch.zeros(128, 256, device="cuda")
            rc = lib.nerfb200_debug_gemm_mn(a.data_ptr(), b.data_ptr(), lbo, sbo, fmt, d.data_ptr(), None)
            torch.cuda.synchronize()
            err = (d - ref).abs().max().item()
            print(f"gemm_mn fmt={fmt} lbo={lbo} sbo={sbo} rc={rc} max_abs_err={err:.3e} ref_absmax={ref.abs().max().item():.3f}"
                  + ("  <-- MATCH" if err < 2e-3 else ""))
    print("GEMM_MN_DONE")


def step_mlp():
    m = make_models()[0]
    torch.manual_seed(1)
    for n in (128, 1000, 40000):
        x = torch.randn(n, 90, device="cuda")
        x[:, :63].clamp_(-1, 1)
        with torch.no_grad():
            ref = nb.nerf_forward_torch(m, x, False)
            out = nb.nerf_forward_fused(m, x, False)
            refs = nb.nerf_forward_torch(m, x[:, :63].contiguous(), True)
            outs = nb.nerf_forward_fused(m, x[:, :63].contiguous(), True)
        torch.cuda.synchronize()
        e_rgb = (out[:, :3] - ref[:, :3]).abs().max().item()
        e_sig = ((out[:, 3] - ref[:, 3]).abs() / (1 + ref[:, 3].abs())).max().item()
        e_so = ((outs - refs).abs() / (1 + refs.abs())).max().item()
        print(f"mlp n={n}: rgb max_abs={e_rgb:.3e} sigma rel={e_sig:.3e} sigma_only rel={e_so:.3e}")
    print("MLP_DONE")


def step_render():
    ms = make_models()
    emb = [nb.Embedding(3, 10), nb.Embedding(3, 4)]
    for (n, K, tt, wb) in ((2, 64, False, True), (256, 64, False, True), (1001, 64, True, False), (512, 0, False, True)):
        rays = make_rays(n)
        with torch.no_grad():
            out = nb.render_rays(ms, emb, rays, 64, False, 0, 0, K, 32768, wb, test_time=tt, extras=True)
            zf = out.get("z_vals_fine")
            ref = _render_with_graph(ms, emb, rays, 64, K, False, 0.0, 0.0, wb, tt, None, None, None, zf)
        torch.cuda.synchronize()
        msg = []
        for k, v in ref.items():
            msg.append(f"{k}:{(out[k] - v).abs().max().item():.2e}")
        print(f"render n={n} K={K} tt={tt}: " + " ".join(msg))
        if zf is not None:
            print("   z_fine sorted:", bool((zf[:, 1:] >= zf[:, :-1]).all().item()),
                  "range", zf.min().item(), zf.max().item())
    print("RENDER_DONE")


def step_speed():
    ms = make_models()
    emb = [nb.Embedding(3, 10), nb.Embedding(3, 4)]
    for n in (1024, 32768, 160000):
        rays = make_rays(n)
        with torch.no_grad():
            for _ in range(3):
                nb.render_rays(ms, emb, rays, 64, False, 0, 0, 64, 32768, True, test_time=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                nb.render_rays(ms, emb, rays, 64, False, 0, 0, 64, 32768, True, test_time=True)
            e1.record()
            torch.cuda.synchronize()
        ms_per = e0.elapsed_time(e1) / reps
        sps = n * 192 / (ms_per * 1e-3)
        flops = n * 214794240 / (ms_per * 1e-3)
        print(f"speed n={n}: {ms_per:.3f} ms  {sps:.3e} ray-samples/s  {flops / 1e12:.1f} TFLOP/s")
    print("SPEED_DONE")


def step_mmabench():
    lib = _lib.load()
    names = ["SS N=256", "SS N=128", "TS N=128", "TS N=256", "TS N=128 alt D"]
    for n_ctas in (1, 148):
        out = torch.zeros(n_ctas, 8, dtype=torch.long, device="cuda")
        reps = 16
        for _ in range(2):
            rc = lib.nerfb200_debug_mma_bench(out.data_ptr(), n_ctas, reps, None)
        torch.cuda.synchronize()
        o = out.cpu().numpy()
        for v, nm in enumerate(names):
            cyc = o[:, v].astype(float)
            per = cyc / (reps * 16)
            print(f"mmabench ctas={n_ctas:3d} {nm:16s} rc={rc} cycles/MMA median {sorted(per)[len(per) // 2]:.1f} min {per.min():.1f} max {per.max():.1f}")
    print("MMABENCH_DONE")


def step_cache():
    import numpy as np
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import nerf_oracle as orc
    dev = torch.device("cuda:0")
    emb = [nb.Embedding(3, 10), nb.Embedding(3, 4)]
    m = [nb.NeRF().to(dev), nb.NeRF().to(dev)]
    rays = torch.from_numpy(orc.make_rays(64, 1)).to(dev)
    with torch.no_grad():
        o1 = nb.render_rays(m, emb, rays, 64, False, 0, 0, 64)
        a = o1["rgb_fine"].clone()
        v0 = m[1].rgb[0].bias._version
        m[1].rgb[0].bias.add_(1.0)
        print("version", v0, "->", m[1].rgb[0].bias._version)
        o2 = nb.render_rays(m, emb, rays, 64, False, 0, 0, 64)
        b = o2["rgb_fine"]
    torch.cuda.synchronize()
    print("a[:4]", a[:4].cpu().numpy())
    print("b[:4]", b[:4].cpu().numpy())
    print("opacity_fine", o2["opacity_fine"][:8].cpu().numpy())
    print("equal:", torch.equal(a, b), "nan:", bool(torch.isnan(a).any()), bool(torch.isnan(b).any()))


def step_timeline():
    import numpy as np
    os.environ["NERFB200_FLAGS"] = str(int(os.environ.get("NERFB200_FLAGS", "0")) | 2)
    lib = _lib.load()
    ms = make_models()
    emb = [nb.Embedding(3, 10), nb.Embedding(3, 4)]
    rays = make_rays(32768)
    tt = bool(int(os.environ.get("TT", "1")))
    with torch.no_grad():
        for _ in range(2):
            nb.render_rays(ms, emb, rays, 64, False, 0, 0, 64, 32768, True, test_time=tt)
    torch.cuda.synchronize()
    buf = np.zeros((3, 512, 2), dtype=np.int64)
    rc = lib.nerfb200_debug_timeline(buf.ctypes.data, buf.size)
    print("timeline rc", rc)
    for role, name in ((0, "EPI"), (1, "MMA")):
        ev = buf[role]
        n = int((ev[:, 1] != 0).sum())
        t0 = ev[0, 1]
        print(f"--- {name}: {n} events")
        prev = t0
        for i in range(min(n, 260)):
            tag, t = int(ev[i, 0]), int(ev[i, 1])
            print(f"{name} {i:4d} tag={tag:4d} t={t - t0:9d} dt={t - prev:7d}")
            prev = t
    print("TIMELINE_DONE")


def step_speed1():
    ms = make_models()
    emb = [nb.Embedding(3, 10), nb.Embedding(3, 4)]
    n = int(os.environ.get("SPEED_N", "160000"))
    rays = make_rays(n)
    with torch.no_grad():
        for _ in range(2):
            nb.render_rays(ms, emb, rays, 64, False, 0, 0, 64, 32768, True, test_time=True)
        torch.cuda.synchronize()
        best = 1e9
        filler = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if os.environ.get("SPEED_KERNEL"):
                filler.fill_(1)      # keeps the GPU busy while the wrapper's host work runs: e0..e1 is the kernel
            e0.record()
            nb.render_rays(ms, emb, rays, 64, False, 0, 0, 64, 32768, True, test_time=True)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
    print(f"speed1 n={n} ctas={os.environ.get('NERFB200_MAX_CTAS', 'all')}: best {best:.3f} ms  {n * 192 / (best * 1e-3):.3e} ray-samples/s")


def step_contention():
    lib = _lib.load()
    bgs = {0: "idle", 1: "ld", 2: "ld+cvt+st", 3: "st", 4: "per-kb ld/cvt/st", 5: "mbar poll", 6: "st+fence+arrive"}
    vs = {0: "SS", 1: "TS in-place A", 2: "TS packed A"}
    for n_ctas in (1,):
        for v in vs:
            for bg in bgs:
                out = torch.zeros(n_ctas, 4, dtype=torch.long, device="cuda")
                reps = 64
                for _ in range(2):
                    rc = lib.nerfb200_debug_mma_contention(out.data_ptr(), n_ctas, reps, bg, v, None)
                torch.cuda.synchronize()
                o = out.cpu().numpy().astype(float)
                cyc = np.median(o[:, 0])
                it = np.median(o[:, 1])
                print(f"contention ctas={n_ctas:3d} {vs[v]:14s} bg={bgs[bg]:18s} rc={rc} cyc/MMA {cyc / (reps * 16):7.1f}  "
                      f"bg iters {it:6.0f}  cyc/bg-iter {cyc / max(it, 1):8.1f}")
    print("CONTENTION_DONE")


def step_issue():
    lib = _lib.load()
    cases = [(0, 0, "back-to-back"), (1, 0, "commit/4"),
             (2, 50, "spin 50 after every 4th"), (2, 100, "spin 100 after every 4th"), (2, 200, "spin 200 after every 4th"),
             (2, 400, "spin 400 after every 4th"),
             (3, 30, "spin 30 after every MMA"), (3, 60, "spin 60 after every MMA"), (3, 120, "spin 120 after every MMA"),
             (4, 0, "2 waits+fence before every 4"), (8, 0, "1 wait+fence before every 4"),
             (5, 0, "2 waits+fence after j=0"), (6, 0, "2 waits+fence after j=1"), (7, 0, "2 waits+fence after j=2")]
    for mode, arg, nm in cases:
        out = torch.zeros(1, 4, dtype=torch.long, device="cuda")
        reps = 64
        for _ in range(2):
            rc = lib.nerfb200_debug_mma_contention(out.data_ptr(), 1, reps, -1 - arg, mode, None)
        torch.cuda.synchronize()
        o = out.cpu().numpy().astype(float)
        print(f"issue mode {mode} {nm:30s} rc={rc} cyc/MMA {o[0, 0] / (reps * 16):7.1f}  cyc/slice {o[0, 0] / (reps * 4):7.1f}")
    print("ISSUE_DONE")


def step_tlsum():
    """Timeline of CTA 0, summarised: per-tile phase durations (epilogue thread 0) and MMA-side waits."""
    os.environ["NERFB200_FLAGS"] = str(int(os.environ.get("NERFB200_FLAGS", "0")) | 2)
    lib = _lib.load()
    ms = make_models()
    emb = [nb.Embedding(3, 10), nb.Embedding(3, 4)]
    n = int(os.environ.get("TL_N", "32768"))
    rays = make_rays(n)
    tt = bool(int(os.environ.get("TT", "1")))
    perturb = float(os.environ.get("TL_PERTURB", "0"))
    with torch.no_grad():
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            nb.render_rays(ms, emb, rays, 64, False, perturb, 0, 64, 32768, True, test_time=tt)
            e1.record()
            torch.cuda.synchronize()
    print(f"kernel (events) {e0.elapsed_time(e1) * 1e3:.1f} us for n={n}")
    buf = np.zeros((3, 512, 2), dtype=np.int64)
    rc = lib.nerfb200_debug_timeline(buf.ctypes.data, buf.size)
    epi = [(int(a), int(b)) for a, b in buf[0] if b != 0]
    mma = [(int(a), int(b)) for a, b in buf[1] if b != 0]
    print("rc", rc, "events", len(epi), len(mma))
    t_entry = next((v for tg, v in epi if tg == 90), None)
    t91 = next((v for tg, v in epi if tg == 91), None)
    t99 = next((v for tg, v in epi if tg == 99), None)
    if t_entry and t91:
        print(f"setup (entry -> first group) {t91 - t_entry} cycles; entry -> end {t99 - t_entry if t99 else -1} cycles")
    # generic: print consecutive deltas grouped by (tag -> next tag), aggregated
    agg = {}
    seq = [(tg, v) for tg, v in epi if tg != 90]
    for (a, ta), (b, tb) in zip(seq, seq[1:]):
        agg.setdefault((a, b), []).append(tb - ta)
    print("EPI transitions (tag->tag: count, mean, min, max cycles)")
    for k in sorted(agg):
        v = agg[k]
        print(f"  {k[0]:3d}->{k[1]:3d}: n={len(v):4d} mean {np.mean(v):8.0f} min {min(v):7d} max {max(v):7d}  total {sum(v):9d}")
    # MMA side
    lay = {}
    cur = {}
    for tg, v in mma:
        if 100 <= tg < 200: cur[tg - 100] = v
        elif 200 <= tg < 300 and (tg - 200) in cur: lay.setdefault(tg - 200, {}).setdefault("issue", []).append(v - cur[tg - 200])
        elif 300 <= tg < 400: lay.setdefault(tg - 300, {}).setdefault("w_full", []).append(v)
        elif 400 <= tg < 500: lay.setdefault(tg - 400, {}).setdefault("w_akb", []).append(v)
    starts = [v for tg, v in mma if 100 <= tg < 200]
    tags = [tg for tg, v in mma if 100 <= tg < 200]
    per = {}
    for (t0, v0), v1 in zip(zip(tags, starts), starts[1:]):
        per.setdefault(t0 - 100, []).append(v1 - v0)
    wd = [v for tg, v in mma if tg == 600]
    we = [v for tg, v in mma if tg == 601]
    if wd:
        print(f"MMA tile start: wait d_free mean {np.mean(wd):.0f} max {max(wd)}; wait enc_full mean {np.mean(we):.0f} max {max(we)}  (n={len(wd)})")
        print("  per tile (d_free, enc):", list(zip(wd, we))[:24])
    print("MMA per layer: start->next start mean | issue span | wait full | wait a_kb (cycles)")
    for l in sorted(lay):
        d = lay[l]
        print(f"  L{l}: period {np.mean(per.get(l, [0])):7.0f} issue {np.mean(d.get('issue', [0])):7.0f} "
              f"w_full {np.mean(d.get('w_full', [0])):7.0f} (max {max(d.get('w_full', [0]))}) w_akb {np.mean(d.get('w_akb', [0])):7.0f}")
    if os.environ.get("TL_RAW"):
        ev = [(v, "EPI", tg) for tg, v in epi if tg != 90] + [(v, "MMA", tg) for tg, v in mma if tg < 300 or 500 <= tg < 600]
        ev.sort()
        lo = int(os.environ.get("TL_RAW_LO", "150"))
        t0 = ev[lo][0]
        prev = {"EPI": t0, "MMA": t0}
        for v, role, tg in ev[lo:lo + int(os.environ["TL_RAW"])]:
            pad = "" if role == "EPI" else "                      "
            print(f"{v - t0:8d} {pad}{role} {tg:4d} (+{v - prev[role]})")
            prev[role] = v
    print("TLSUM_DONE")


if __name__ == "__main__":
    t0 = time.time()
    {"gemm": step_gemm, "gemm_mn": step_gemm_mn, "mlp": step_mlp, "render": step_render, "speed": step_speed, "timeline": step_timeline, "mmabench": step_mmabench, "speed1": step_speed1, "contention": step_contention, "issue": step_issue, "tlsum": step_tlsum, "cache": step_cache}[sys.argv[1]]()
    print(f"[{sys.argv[1]}] {time.time() - t0:.1f}s")
