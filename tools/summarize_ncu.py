"""Summarise ncu captures brought back in gpurun_out/ into profiles/ (text, committed).

    python tools/summarize_ncu.py <round-tag>   e.g. r01
Reads gpurun_out/launches_<tag>.csv (per-launch gpu__time_duration) and
gpurun_out/prof_<tag>_render.ncu-rep (--set full capture of render_rays_kernel)."""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = []

lp = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 5 and r[0].strip('"').isdigit()]
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r[4].split("(")[0].split("<")[0].strip()
        if "distribution" in r[4]:
            name = "at::distribution_elementwise (torch.rand/randn)"
        if "vectorized_elementwise" in r[4] or "FillFunctor" in r[4]:
            name = "at::vectorized_elementwise (fill / L2 flush)"
        tot[name][0] += 1
        tot[name][1] += float(r[-1])
    total = sum(v[1] for v in tot.values())
    out.append(f"## Launch list ({os.path.basename(lp)}; ncu --metrics gpu__time_duration.sum, cold-cache, serialised)\n")
    out.append("| kernel | launches | total ns | share |\n|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1]:.0f} | {100 * v[1] / total:.1f}% |")
    rr = [float(r[-1]) for r in rows if "render_rays_kernel" in r[4]]
    if rr:
        out.append(f"\nrender_rays_kernel: {len(rr)} launches, mean {sum(rr) / len(rr) / 1e3:.1f} us "
                   f"(1024 rays x 192 samples each; ncu's clock, cold caches, serialised launches).\n")

# bench-shaped launch (1024 rays): DRAM traffic per launch for bench.py's roofline.traffic
bp = os.path.join(ROOT, "gpurun_out", f"prof_{tag}_bench1024.ncu-rep")
if os.path.exists(bp):
    import json
    raw = subprocess.run(["ncu", "-i", bp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}

    def to_bytes(key):
        u, v = d[key]
        f = float(v.replace(",", ""))
        return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    info = {"kernel": "render_rays_kernel", "launch": "1024 rays, 64+64, training-mode forward (bench.py workload), L2 cold",
            "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes_per_launch": rd + wr,
            "gpu_time_us": float(d["gpu__time_duration.sum"][1].replace(",", "")) * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(d["gpu__time_duration.sum"][0], 1),
            "tensor_pipe_active_pct": float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"][1]),
            "source": f"profiles/{tag}_ncu_summary.md (ncu --set full --clock-control none, {os.path.basename(bp)})"}
    json.dump(info, open(os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic.json"), "w"), indent=1)
    out.append(f"## Bench-shaped launch ({os.path.basename(bp)}: 1024 rays, L2 flushed before the launch)\n")
    out.append("| metric | value |\n|---|---|")
    for k, v in info.items():
        out.append(f"| {k} | {v} |")
    out.append("")

rp = os.path.join(ROOT, "gpurun_out", f"prof_{tag}_render.ncu-rep")
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic"]
    out.append(f"## Full capture ({os.path.basename(rp)}; ncu --set full --clock-control none, one launch: "
               f"32768 rays, 64+64 samples, training-mode forward)\n")
    out.append("| metric | unit | value |\n|---|---|---|")
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            out.append(f"| `{h}` | {u} | {v} |")
    src = subprocess.run(["ncu", "-i", rp, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    shdr, data = srows[1], srows[2:]
    i_s = shdr.index("# Samples")
    stall = [i for i, h in enumerate(shdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[i_s]) for r in data)
    out.append(f"\nWarp-stall samples (all warps incl. the spinning producer / MMA-issuer threads), total {total}:\n")
    out.append("| stall | share |\n|---|---|")
    for i in sorted(stall, key=lambda i: -sum(int(r[i]) for r in data)):
        sh = sum(int(r[i]) for r in data)
        if sh > 0.01 * total:
            out.append(f"| {shdr[i]} | {100 * sh / total:.1f}% |")
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "nerf_pl_b200", "libnerf_pl_b200.so")],
                          capture_output=True, text=True).stdout
    cnt = {k: sass.count(k) for k in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "SYNCS.PHASECHK", "FADD2", "F2FP.RELU", "HMMA.")}
    out.append("\nSASS mnemonics in libnerf_pl_b200.so (cuobjdump): " + ", ".join(f"{k} x{v}" for k, v in cnt.items()) + "\n")

# training-step capture: one --set full pass over the step's big kernels (tools/prof_train.py)
tp = os.path.join(ROOT, "gpurun_out", f"prof_{tag}_train.ncu-rep")
if os.path.exists(tp):
    raw = subprocess.run(["ncu", "-i", tp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (elapsed)"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
            ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
    out.append(f"## Training step, 1024 rays ({os.path.basename(tp)}; ncu --set full --clock-control none, one step: "
               f"fused forward in training mode + loss, then the backward kernels)\n")
    out.append("| kernel | " + " | ".join(c[1] for c in cols) + " |\n|---|" + "---|" * len(cols))
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        cells = []
        for key, _ in cols:
            i = hdr.index(key)
            cells.append(f"{r[i]} {units[i]}".strip())
        out.append(f"| `{r[kn].split('(')[0]}` | " + " | ".join(cells) + " |")
    out.append("")
lt = os.path.join(ROOT, "gpurun_out", f"launches_{tag}_train.csv")
if os.path.exists(lt):
    rows = [r for r in csv.reader(open(lt)) if len(r) > 5 and r[0].strip('"').isdigit()]
    per = defaultdict(lambda: [0, 0.0])
    n_steps = max(1, sum(1 for r in rows if "adam_kernel" in r[4]))
    for r in rows:
        name = r[4].split("(")[0].strip()
        if "at::" in name:
            name = "torch fill / rand (zero_grad buffers, random draws)"
        per[name][0] += 1
        per[name][1] += float(r[-1])
    total = sum(v[1] for v in per.values())
    out.append(f"## Launch list of the training step ({os.path.basename(lt)}, {n_steps} steps incl. warm-up; gpu__time_duration, serialised)\n")
    out.append("| kernel | launches / step | us / step | share |\n|---|---|---|---|")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0] / n_steps:.1f} | {v[1] / n_steps / 1e3:.1f} | {100 * v[1] / total:.1f}% |")
    out.append(f"\nsum {total / n_steps / 1e3:.0f} us per step.\n")

os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
dst = os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md")
open(dst, "w").write(f"# ncu summary {tag}\n\n" + "\n".join(out) + "\n")
print(open(dst).read())
