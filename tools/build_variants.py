"""Build experiment variants of the library into nerf_pl_b200/variants/ (git-ignored, travels to
the GPU box).  Usage: python tools/build_variants.py name=-DFLAG[,-DFLAG2] ...
Select at run time with NERFB200_LIB=nerf_pl_b200/variants/lib_<name>.so."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerf_pl_b200 import _lib  # noqa: E402


def main():
    out_dir = os.path.join(ROOT, "nerf_pl_b200", "variants")
    os.makedirs(out_dir, exist_ok=True)
    procs = []
    for spec in sys.argv[1:]:
        name, _, flags = spec.partition("=")
        out = os.path.join(out_dir, f"lib_{name}.so")
        cmd = [_lib._nvcc(), *_lib.NVCC_FLAGS, *[f for f in flags.split(",") if f], "-o", out] + \
              [os.path.join(_lib.CSRC, s) for s in _lib.SOURCES]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        txt, _ = p.communicate()
        print(name, "rc", p.returncode, txt[-2000:] if p.returncode else "")
        if p.returncode:
            sys.exit(1)


if __name__ == "__main__":
    main()
