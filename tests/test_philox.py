"""The host replica of the in-kernel uniform generator against the published Philox4x32-10 known-answer vectors
(Random123 kat_vectors) and basic uniformity; CPU only.  The GPU test (test_gpu_parity.py) then shows that the kernel
generates exactly these numbers."""
import numpy as np

from tests import philox


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), 0, (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, 0xffffffffffffffff, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0x299f31d0 << 32) | 0xa4093822,
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*[[c] for c in ctr], key)
        assert tuple(int(g[0]) for g in got) == want


def test_uniform_mapping():
    u = philox.uniform(20260923, 2048, 128, 1)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 2e-3 and abs(u.std() - 12 ** -0.5) < 2e-3
    hist = np.histogram(u, bins=16, range=(0, 1))[0] / u.size
    assert np.abs(hist - 1 / 16).max() < 2e-3
    # rows are independent of the launch shape: ray r of a 2048-ray call == ray r of a call that starts at r
    np.testing.assert_array_equal(philox.uniform(20260923, 3, 128, 1, ray0=77), u[77:80])


def test_kernel_rng_option_resolves_without_tensors():
    """randoms='kernel' / {'seed': s} (nerf_pl_b200/rendering.py _resolve_randoms): no tensors are drawn for the uniform
    inputs, the seed is deterministic under torch.manual_seed and differs from call to call."""
    import torch

    from nerf_pl_b200 import rendering
    dev = torch.device("cpu")
    pr, nc, ur, nf, seed = rendering._resolve_randoms({"seed": 2 ** 64 + 5}, 8, 64, 64, 1.0, 0.0, dev, True)
    assert (pr, nc, ur, nf) == (None, None, None, None) and seed == 5
    torch.manual_seed(123)
    rendering._KERNEL_RNG_CALLS = 0
    a = [rendering._resolve_randoms("kernel", 8, 64, 64, 1.0, 0.0, dev, True)[4] for _ in range(3)]
    torch.manual_seed(123)
    rendering._KERNEL_RNG_CALLS = 0
    b = [rendering._resolve_randoms("kernel", 8, 64, 64, 1.0, 0.0, dev, True)[4] for _ in range(3)]
    assert a == b and len(set(a)) == 3 and all(0 <= s < 2 ** 64 for s in a)
    # the Gaussian inputs stay tensors
    pr, nc, ur, nf, seed = rendering._resolve_randoms({"seed": 1}, 4, 64, 64, 1.0, 1.0, dev, True)
    assert pr is None and ur is None and nc.shape == (4, 64) and nf.shape == (4, 128)
    import pytest
    with pytest.raises(ValueError):
        rendering._resolve_randoms("philox", 4, 64, 64, 1.0, 0.0, dev, True)
