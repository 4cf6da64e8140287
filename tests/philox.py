"""Host replica (numpy) of the render kernel's in-kernel uniform generator (csrc/render_kernel.cuh philox_uniform,
include/nerf_pl_b200.h rng_in_kernel): Philox4x32-10, counter = {ray, i >> 2, stream, 0}, key = seed, word i & 3,
u = (x >> 8) * 2^-24.  The parity tests feed these numbers to the oracle."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, seed):
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3)]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def uniform(seed: int, n_rays: int, count: int, stream: int, ray0: int = 0) -> np.ndarray:
    """(n_rays, count) float32: element (r, i) as the kernel generates it for global ray index ray0 + r."""
    r = (np.arange(n_rays, dtype=np.uint64) + np.uint64(ray0))[:, None]
    i = np.arange(count, dtype=np.uint64)[None, :]
    words = philox4x32_10(np.broadcast_to(r, (n_rays, count)), np.broadcast_to(i >> np.uint64(2), (n_rays, count)),
                          np.full((n_rays, count), stream, np.uint64), np.zeros((n_rays, count), np.uint64), int(seed))
    sel = (i & np.uint64(3)).astype(np.int64)
    w = np.choose(np.broadcast_to(sel, (n_rays, count)), words)
    return ((w >> np.uint64(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def randoms(seed: int, n_rays: int, n_samples: int, n_importance: int):
    out = {"perturb_rand": uniform(seed, n_rays, n_samples, 0)}
    if n_importance > 0:
        out["u_rand"] = uniform(seed, n_rays, n_importance, 1)
    return out
