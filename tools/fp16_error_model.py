"""Error model of the fp16 MLP (CPU, numpy): the fp32 oracle with the operands of the ten big layers rounded to
fp16 (weights, activations, or both; fp32 accumulation) against the reference goldens.  This is what bounds the
parity tolerances of the fused kernel: it computes exactly these roundings.

    python tools/fp16_error_model.py [w|a|wa] [trained|random]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nerf_oracle as orc  # noqa: E402
from tests import cases  # noqa: E402

F32 = np.float32
mode = sys.argv[1] if len(sys.argv) > 1 else "wa"
which = sys.argv[2] if len(sys.argv) > 2 else "trained"
BIG = {f"xyz_encoding_{i}.0" for i in range(1, 9)} | {"xyz_encoding_final", "dir_encoding.0"}


def r16(a):
    return a.astype(np.float16).astype(F32)


def linear(w, name, x):
    W = w[name + ".weight"]
    if name in BIG:
        x = r16(x) if "a" in mode else x
        W = r16(W) if "w" in mode else W
    return (x @ W.T + w[name + ".bias"]).astype(F32)


orc._linear = linear
if which == "trained":
    ws = cases.trained_weights()
    todo = {k: (v[0], "blender", v[1], 64, v[2], False, v[3], v[4], True, v[5]) for k, v in cases.TRAINED_CASES.items()}
else:
    ws = cases.weights()
    todo = cases.CASES
for name, (n, kind, rseed, S, K, disp, perturb, noise, wb, tt) in todo.items():
    rays, randoms, ref = cases.load_case(name)
    out = orc.render_rays(ws, rays, S, disp, perturb, noise, K, wb, tt, randoms or None)
    print(mode, name, " ".join(f"{k}={np.abs(out[k] - ref[k]).max():.2e}/{np.abs(out[k] - ref[k]).mean():.1e}" for k in ref))
