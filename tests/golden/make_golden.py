"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(kwea123/nerf_pl, /root/reference, read-only) on CPU in the build container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz
    python tests/golden/make_golden.py NAME ...   # only the named render cases

The reference's renderer imports `torchsearchsorted`, whose native extension does not build
against torch 2.11; it is shimmed with torch.searchsorted, which the survey
verified to be bit-identical on the reference's own test grid.  Inputs come from the oracle's
deterministic generators (oracle/nerf_oracle.py make_weights / make_rays), so tests can rebuild
them from seeds and only rays + reference outputs are stored.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import nerf_oracle as orc  # noqa: E402

REF = os.environ.get("NERF_PL_REFERENCE", "/root/reference")


def import_reference():
    shim = types.ModuleType("torchsearchsorted")
    shim.searchsorted = lambda a, v, out=None, side="left": torch.searchsorted(
        a.contiguous(), v.contiguous(), right=(side == "right"))
    sys.modules["torchsearchsorted"] = shim
    # datasets/ray_utils.py:2 imports kornia.create_meshgrid (kornia is not installed here): the one
    # function it uses is restated (pixel grid, x = column index, y = row index, un-normalised)
    kor = types.ModuleType("kornia")

    def create_meshgrid(H, W, normalized_coordinates=False):
        assert not normalized_coordinates
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
        return torch.stack([xs, ys], -1)[None]
    kor.create_meshgrid = create_meshgrid
    sys.modules["kornia"] = kor
    sys.path.insert(0, REF)
    from models.nerf import Embedding, NeRF
    from models.rendering import render_rays, sample_pdf
    return Embedding, NeRF, render_rays, sample_pdf


def ref_model(NeRF, weights):
    m = NeRF()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    return m.eval()


# name: (n_rays, ray kind, ray seed, N_samples, N_importance, use_disp, perturb, noise_std,
#        white_back, test_time)
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
W_SEEDS = (11, 12)   # coarse, fine

# Gradient goldens: the reference's training step (train.py:103-117: results = render_rays(...);
# loss = MSELoss(results, rgbs) (losses.py:9-14); loss.backward()) with autograd enabled.
# name: (n_rays, ray kind, ray seed, N_importance, perturb, noise_std, white_back)
GRAD_CASES = {
    "grad_blender_noise0": (64, "blender", 31, 64, 1.0, 0.0, True),     # README Blender recipe (README.md:75-83)
    "grad_ndc_noise1": (48, "ndc", 32, 64, 1.0, 1.0, False),            # README LLFF recipe (README.md:104-111)
}


# Trained weights (tests/golden/trained_weights.npz, written on a B200 by tools/train_sharp_weights.py with THIS
# repository's training step; larger norms and high-frequency content than the random-init W_SEEDS) through the
# unmodified reference: two renders and one training-step gradient.
# name: (n_rays, ray seed, N_importance, perturb, noise_std, test_time)
TRAINED_CASES = {
    "trained_test": (192, 51, 64, 0.0, 0.0, True),
    "trained_train": (128, 52, 64, 1.0, 0.0, False),
    "trained_noise": (96, 53, 128, 1.0, 1.0, False),
}
TRAINED_GRAD = ("grad_trained", 64, 54, 64, 1.0, 0.0)     # name, n_rays, ray seed, N_importance, perturb, noise_std


def load_trained_weights():
    z = np.load(os.path.join(HERE, "trained_weights.npz"))
    return [{k[len(t) + 1:]: z[k] for k in z.files if k.startswith(t + ".")} for t in ("coarse", "fine")]


def trained_rays(n, seed):
    """Blender-style rays (origin on a radius-4 sphere, looking at the scene the weights were trained on)."""
    sys.path.insert(0, ROOT)
    import bench
    return bench.blender_rays(n, seed)


def make_trained_cases(NeRF, Embedding, render_rays, only):
    if not os.path.exists(os.path.join(HERE, "trained_weights.npz")):
        print("no trained_weights.npz: trained cases skipped")
        return
    ws = load_trained_weights()
    emb = [Embedding(3, 10), Embedding(3, 4)]
    S = 64
    for name, (n, rseed, K, perturb, noise, tt) in TRAINED_CASES.items():
        if only and name not in only:
            continue
        models = [ref_model(NeRF, w) for w in ws]
        rays = trained_rays(n, rseed)
        store = {"rays": rays}
        torch.manual_seed(3000 + rseed)
        if perturb > 0 or noise > 0:
            g = torch.get_rng_state()
            store["perturb_rand"] = torch.rand(n, S).numpy()
            store["noise_coarse"] = torch.randn(n, S).numpy()
            store["u_rand"] = torch.rand(n, K).numpy()
            store["noise_fine"] = torch.randn(n, S + K).numpy()
            torch.set_rng_state(g)
        with torch.no_grad():
            out = render_rays(models, emb, torch.from_numpy(rays), S, False, perturb, noise, K, 1024 * 32, True,
                              test_time=tt)
        for k, v in out.items():
            store["out_" + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, f"render_{name}.npz"), **store)
        print(name, {k: (tuple(v.shape), float(v.mean())) for k, v in out.items()})
    name, n, rseed, K, perturb, noise = TRAINED_GRAD
    if only and name not in only:
        return
    models = [ref_model(NeRF, w).train() for w in ws]
    rays = trained_rays(n, rseed)
    target = np.random.RandomState(500 + rseed).uniform(0, 1, (n, 3)).astype(np.float32)
    store = {"rays": rays, "target": target}
    torch.manual_seed(2000 + rseed)
    g = torch.get_rng_state()
    store["perturb_rand"] = torch.rand(n, S).numpy()
    store["noise_coarse"] = torch.randn(n, S).numpy()
    store["u_rand"] = torch.rand(n, K).numpy()
    store["noise_fine"] = torch.randn(n, S + K).numpy()
    torch.set_rng_state(g)
    out = render_rays(models, emb, torch.from_numpy(rays), S, False, perturb, noise, K, 1024 * 32, True, test_time=False)
    tgt = torch.from_numpy(target)
    loss = torch.nn.functional.mse_loss(out["rgb_coarse"], tgt) + torch.nn.functional.mse_loss(out["rgb_fine"], tgt)
    loss.backward()
    store["loss"] = np.float32(loss.item())
    for k, v in out.items():
        store["out_" + k] = v.detach().numpy()
    grads = {f"{tag}.{key}": prm.grad.numpy() for tag, m in zip(("coarse", "fine"), models) for key, prm in m.named_parameters()}
    store.update(pack_grads(grads))
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **store)
    print(name, "loss", float(loss.item()))


def pack_grads(grads):
    """48 fp32 gradient tensors -> per-tensor max-abs scale (fp32) + values / scale as fp16
    (relative precision 5e-4, far below the 5e-2 test tolerance; keeps the fixture ~2.4 MB)."""
    store = {}
    for key, g in grads.items():
        g = np.asarray(g, dtype=np.float32)
        sc = np.float32(max(float(np.abs(g).max()), 1e-30))
        store["gscale_" + key] = sc
        store["g16_" + key] = (g / sc).astype(np.float16)
    return store


def make_grad_cases(NeRF, Embedding, render_rays, only):
    for name, (n, kind, rseed, K, perturb, noise, wb) in GRAD_CASES.items():
        if only and name not in only:
            continue
        S = 64
        ws = [orc.make_weights(s) for s in W_SEEDS]
        models = [ref_model(NeRF, w).train() for w in ws]
        emb = [Embedding(3, 10), Embedding(3, 4)]
        rays = orc.make_rays(n, rseed, kind)
        target = np.random.RandomState(500 + rseed).uniform(0, 1, (n, 3)).astype(np.float32)
        store = {"rays": rays, "target": target}
        torch.manual_seed(2000 + rseed)
        g = torch.get_rng_state()
        store["perturb_rand"] = torch.rand(n, S).numpy()
        store["noise_coarse"] = torch.randn(n, S).numpy()
        store["u_rand"] = torch.rand(n, K).numpy()
        store["noise_fine"] = torch.randn(n, S + K).numpy()
        torch.set_rng_state(g)
        out = render_rays(models, emb, torch.from_numpy(rays), S, False, perturb, noise, K, 1024 * 32, wb,
                          test_time=False)
        tgt = torch.from_numpy(target)
        loss = torch.nn.functional.mse_loss(out["rgb_coarse"], tgt) + torch.nn.functional.mse_loss(out["rgb_fine"], tgt)
        loss.backward()
        store["loss"] = np.float32(loss.item())
        for k, v in out.items():
            store["out_" + k] = v.detach().numpy()
        grads = {}
        for tag, m in zip(("coarse", "fine"), models):
            for key, prm in m.named_parameters():
                grads[f"{tag}.{key}"] = prm.grad.numpy()
        store.update(pack_grads(grads))
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **store)
        gn = float(np.sqrt(sum(float((v.astype(np.float64) ** 2).sum()) for v in grads.values())))
        print(name, "loss", float(loss.item()), "grad norm", gn)


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    Embedding, NeRF, render_rays, sample_pdf = import_reference()
    ws = [orc.make_weights(s) for s in W_SEEDS]
    models = [ref_model(NeRF, w) for w in ws]
    emb = [Embedding(3, 10), Embedding(3, 4)]
    only = sys.argv[1:]
    make_grad_cases(NeRF, Embedding, render_rays, only)
    make_trained_cases(NeRF, Embedding, render_rays, only)
    for name, (n, kind, rseed, S, K, disp, perturb, noise, wb, tt) in CASES.items():
        if only and name not in only:
            continue
        rays = orc.make_rays(n, rseed, kind)
        store = {"rays": rays}
        torch.manual_seed(1000 + rseed)
        if perturb > 0 or noise > 0:
            # replay the reference's draws (models/rendering.py:203, :152, :39, :152) to record them
            g = torch.get_rng_state()
            if perturb > 0:
                store["perturb_rand"] = torch.rand(n, S).numpy()
            store["noise_coarse"] = torch.randn(n, S).numpy()
            if K > 0:
                if perturb > 0:
                    store["u_rand"] = torch.rand(n, K).numpy()
                store["noise_fine"] = torch.randn(n, S + K).numpy()
            torch.set_rng_state(g)
        with torch.no_grad():
            out = render_rays(models, emb, torch.from_numpy(rays), S, disp, perturb, noise, K, 1024 * 32,
                              wb, test_time=tt)
        for k, v in out.items():
            store["out_" + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, f"render_{name}.npz"), **store)
        print(name, {k: tuple(v.shape) for k, v in out.items()})

    if only:
        return
    # unit vectors: Embedding, NeRF.forward, sample_pdf
    rs = np.random.RandomState(21)
    x3 = rs.uniform(-6, 6, (64, 3)).astype(np.float32)
    with torch.no_grad():
        e10 = emb[0](torch.from_numpy(x3)).numpy()
        e4 = emb[1](torch.from_numpy(x3 / 6)).numpy()
        xin = np.concatenate([e10, e4], -1)
        full = models[0](torch.from_numpy(xin)).numpy()
        sig = models[1](torch.from_numpy(e10), sigma_only=True).numpy()
        bins = np.sort(rs.uniform(2, 6, (32, 63)).astype(np.float32), -1)
        wts = (rs.uniform(0, 1, (32, 62)) ** 4).astype(np.float32)
        wts[:4] = 0
        u = rs.uniform(0, 1, (32, 64)).astype(np.float32)
        sp_det = sample_pdf(torch.from_numpy(bins), torch.from_numpy(wts), 64, det=True).numpy()
        # non-det: replay the rand draw
        torch.manual_seed(77)
        u_ref = torch.rand(32, 48).numpy()
        torch.manual_seed(77)
        sp_rand = sample_pdf(torch.from_numpy(bins), torch.from_numpy(wts), 48, det=False).numpy()
    np.savez_compressed(os.path.join(HERE, "units.npz"), x3=x3, embed10=e10, embed4=e4, nerf_in=xin,
                        nerf_full=full, nerf_sigma=sig, pdf_bins=bins, pdf_weights=wts, pdf_det=sp_det,
                        pdf_u=u_ref, pdf_rand=sp_rand, linspace64=torch.linspace(0, 1, 64).numpy(),
                        linspace128=torch.linspace(0, 1, 128).numpy())
    print("units ok")

    # ray generation ("next" row): datasets/ray_utils.py run as datasets/blender.py / llff.py use it
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_ray_utils", os.path.join(REF, "datasets", "ray_utils.py"))
    ru = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ru)
    H, W, focal = 24, 36, 41.5
    th = 0.7
    c2w = np.array([[np.cos(th), 0, np.sin(th), 1.5], [0.2, 0.96, -0.1, -0.3], [-np.sin(th), 0.1, np.cos(th), 3.2]],
                   dtype=np.float32)
    dirs = ru.get_ray_directions(H, W, focal)
    ro, rd = ru.get_rays(dirs, torch.from_numpy(c2w))
    blender = torch.cat([ro, rd, 2.0 * torch.ones_like(ro[:, :1]), 6.0 * torch.ones_like(ro[:, :1])], 1).numpy()
    no, nd = ru.get_ndc_rays(H, W, focal, 1.0, ro, rd)
    ndc = torch.cat([no, nd, 0 * torch.ones_like(ro[:, :1]), 1 * torch.ones_like(ro[:, :1])], 1).numpy()
    np.savez_compressed(os.path.join(HERE, "raygen.npz"), H=H, W=W, focal=focal, c2w=c2w, blender=blender, ndc=ndc)
    print("raygen ok", blender.shape)


if __name__ == "__main__":
    main()
