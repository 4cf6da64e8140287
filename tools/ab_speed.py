"""A/B speed comparison of library builds on the same box (interleaved rounds).
Usage: python tools/ab_speed.py libA.so libB.so [...]      env AB_ROUNDS (3), AB_SIZES ("160000,1024")"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sys.argv[1:]
for rnd in range(int(os.environ.get("AB_ROUNDS", "3"))):
    for lib in libs:
        for n in os.environ.get("AB_SIZES", "160000,1024").split(","):
            env = dict(os.environ, NERFB200_LIB=lib, SPEED_N=n, SPEED_KERNEL="1")
            r = subprocess.run([sys.executable, "-u", "tools/gpu_probe.py", "speed1"], cwd=ROOT, env=env,
                               capture_output=True, text=True, timeout=120)
            line = [l for l in r.stdout.splitlines() if l.startswith("speed1")]
            print(rnd, os.path.basename(lib), line[0] if line else r.stderr[-300:], flush=True)
