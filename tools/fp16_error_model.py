"""Error model of the fp16 MLP (CPU, numpy): the fp32 oracle with the operands of the ten big layers rounded to
fp16 (weights, activations, or both; fp32 accumulation) against the reference goldens.  This is what bounds the
parity tolerances of the fused kernel: it computes exactly these roundings.  tests/test_oracle_golden.py pins the
numbers the tolerances in tests/test_gpu_parity.py and DESIGN.md section 5 rest on.

    python tools/fp16_error_model.py [w|a|wa] [trained|random]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nerf_oracle as orc  # noqa: E402
from tests import cases  # noqa: E402

F32 = np.float32
BIG = {f"xyz_encoding_{i}.0" for i in range(1, 9)} | {"xyz_encoding_final", "dir_encoding.0"}


def r16(a):
    return a.astype(np.float16).astype(F32)


class rounded_operands:
    """Context manager: oracle._linear with the big layers' operands rounded to fp16 ('w', 'a' or 'wa')."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.saved = orc._linear
        mode = self.mode

        def linear(w, name, x):
            W = w[name + ".weight"]
            if name in BIG:
                x = r16(x) if "a" in mode else x
                W = r16(W) if "w" in mode else W
            return (x @ W.T + w[name + ".bias"]).astype(F32)
        orc._linear = linear
        return self

    def __exit__(self, *exc):
        orc._linear = self.saved


def case_errors(mode, name, trained):
    """max |oracle(fp16-rounded operands) - reference| per result key for one golden case."""
    if trained:
        n, rseed, K, perturb, noise, tt = cases.TRAINED_CASES[name]
        ws, S, disp, wb = cases.trained_weights(), 64, False, True
    else:
        n, kind, rseed, S, K, disp, perturb, noise, wb, tt = cases.CASES[name]
        ws = cases.weights()
    rays, randoms, ref = cases.load_case(name)
    with rounded_operands(mode):
        out = orc.render_rays(ws, rays, S, disp, perturb, noise, K, wb, tt, randoms or None)
    return {k: (float(np.abs(out[k] - ref[k]).max()), float(np.abs(out[k] - ref[k]).mean())) for k in ref}


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "wa"
    trained = (sys.argv[2] if len(sys.argv) > 2 else "trained") == "trained"
    for name in (cases.TRAINED_CASES if trained else cases.CASES):
        e = case_errors(mode, name, trained)
        print(mode, name, " ".join(f"{k}={v[0]:.2e}/{v[1]:.1e}" for k, v in e.items()))
