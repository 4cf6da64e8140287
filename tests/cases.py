"""The golden render cases (mirrors tests/golden/make_golden.py CASES) and loaders."""
import os

import numpy as np

from oracle import nerf_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# name: (n_rays, kind, ray seed, N_samples, N_importance, use_disp, perturb, noise_std, white_back, test_time)
CASES = {
    "c1_coarse_only": (256, "blender", 1, 64, 0, False, 0.0, 0.0, True, False),
    "blender_64_64": (128, "blender", 2, 64, 64, False, 0.0, 0.0, True, False),
    "blender_64_64_test": (128, "blender", 3, 64, 64, False, 0.0, 0.0, True, True),
    "ndc_64_64_test": (96, "ndc", 4, 64, 64, False, 0.0, 0.0, False, True),
    "blender_train_rng": (64, "blender", 5, 64, 64, False, 1.0, 1.0, True, False),
    "blender_disp": (64, "blender", 6, 64, 64, True, 0.0, 0.0, True, False),
    "blender_64_128": (64, "blender", 7, 64, 128, False, 0.0, 0.0, True, True),
    "odd_rays": (33, "blender", 8, 64, 64, False, 0.0, 0.0, False, False),
    "ndc_perturb_128": (70, "ndc", 9, 64, 128, False, 1.0, 0.0, False, False),
    "blender_disp_perturb": (50, "blender", 10, 64, 64, True, 1.0, 0.0, True, True),
    # configs[3] (LLFF fern recipe, README.md:104-111): NDC rays, perturb 1, noise_std 1, white_back False
    "ndc_train_noise1": (80, "ndc", 13, 64, 64, False, 1.0, 1.0, False, False),
}
W_SEEDS = (11, 12)
RANDOM_KEYS = ("perturb_rand", "noise_coarse", "u_rand", "noise_fine")


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f"render_{name}.npz"))
    rays = z["rays"]
    randoms = {k: z[k] for k in RANDOM_KEYS if k in z.files}
    ref = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    return rays, randoms, ref


def weights():
    return [orc.make_weights(s) for s in W_SEEDS]


def error_stats(a, b):
    d = np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).ravel()
    return float(d.max()), float(np.percentile(d, 99.9)), float(d.mean())

# Gradient goldens (mirrors tests/golden/make_golden.py GRAD_CASES):
# name: (n_rays, kind, ray seed, N_importance, perturb, noise_std, white_back)
GRAD_CASES = {
    "grad_blender_noise0": (64, "blender", 31, 64, 1.0, 0.0, True),
    "grad_ndc_noise1": (48, "ndc", 32, 64, 1.0, 1.0, False),
}


def load_grad_case(name):
    """rays, target, randoms, reference loss, reference outputs, reference gradients ('coarse.<key>' / 'fine.<key>')."""
    from oracle import nerf_oracle_grad as og
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    randoms = {k: z[k] for k in RANDOM_KEYS if k in z.files}
    ref = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    return z["rays"], z["target"], randoms, float(z["loss"]), ref, og.unpack_golden_grads(z)


# Trained weights (tests/golden/trained_weights.npz: 8000 steps of this repository's own training step on a B200,
# tools/train_sharp_weights.py) through the unmodified reference (make_golden.py TRAINED_CASES / TRAINED_GRAD).
# name: (n_rays, ray seed, N_importance, perturb, noise_std, test_time)
TRAINED_CASES = {
    "trained_test": (192, 51, 64, 0.0, 0.0, True),
    "trained_train": (128, 52, 64, 1.0, 0.0, False),
    "trained_noise": (96, 53, 128, 1.0, 1.0, False),
}
TRAINED_GRAD = ("grad_trained", 64, 54, 64, 1.0, 0.0)


def have_trained():
    return os.path.exists(os.path.join(GOLDEN, "trained_weights.npz"))


def trained_weights():
    z = np.load(os.path.join(GOLDEN, "trained_weights.npz"))
    return [{k[len(t) + 1:]: z[k] for k in z.files if k.startswith(t + ".")} for t in ("coarse", "fine")]
