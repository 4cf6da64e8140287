"""One-call experiment batch (GPU box): contention microbenchmark, timelines, CTA sweep, variants.
Every leg runs in a subprocess with its own timeout so one failure cannot hang the call."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PY = sys.executable


def run(tag, args, env=None, timeout=120):
    e = dict(os.environ)
    e.update(env or {})
    print(f"===== {tag}", flush=True)
    try:
        r = subprocess.run([PY, "-u", *args], cwd=ROOT, env=e, timeout=timeout, capture_output=True, text=True)
        print(r.stdout[-6000:])
        if r.returncode:
            print("rc", r.returncode, r.stderr[-1500:])
    except subprocess.TimeoutExpired:
        print("TIMEOUT")
    sys.stdout.flush()


def main():
    legs = sys.argv[1:] or ["contention", "tl1024", "tl32k", "sweep", "variants"]
    V = os.path.join(ROOT, "nerf_pl_b200", "variants")
    if "contention" in legs:
        run("contention", ["tools/gpu_probe.py", "contention"])
    if "tl1024" in legs:
        run("timeline n=1024 train", ["tools/gpu_probe.py", "tlsum"], {"TL_N": "1024", "TT": "0", "TL_PERTURB": "1"})
    if "tl32k" in legs:
        run("timeline n=32768 test", ["tools/gpu_probe.py", "tlsum"], {"TL_N": "32768", "TT": "1"})
    if "sweep" in legs:
        for c in (148, 74, 37):
            run(f"speed ctas={c}", ["tools/gpu_probe.py", "speed1"], {"NERFB200_MAX_CTAS": str(c)})
    if "variants" in legs:
        for name in sorted(os.listdir(V)) if os.path.isdir(V) else []:
            if not name.endswith(".so"):
                continue
            env = {"NERFB200_LIB": os.path.join(V, name)}
            run(f"variant {name} speed", ["tools/gpu_probe.py", "speed1"], env)
            if "noload" in name:
                run(f"variant {name} timeline", ["tools/gpu_probe.py", "tlsum"], dict(env, TL_N="32768", TT="1"))


if __name__ == "__main__":
    main()
