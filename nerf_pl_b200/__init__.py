"""nerf_pl_b200 — B200-native (sm_100a) implementation of the volumetric-rendering hot path of
kwea123/nerf_pl: ``render_rays`` + ``Embedding`` / ``NeRF`` behind the reference's own Python
signatures, executed by hand-written tcgen05/TMEM CUDA kernels through a C-ABI library
(``include/nerf_pl_b200.h``).  See DESIGN.md and INTEGRATION.md."""
from .nerf import (Embedding, NeRF, invalidate_packed, nerf_forward_fused, nerf_forward_torch, nerf_parameters,
                   packed_weights)
from .inference import batched_inference, generate_rays, mse_psnr, query_sigma, render_image, to_uint8
from .optim import FusedAdam
from .rendering import render_rays, render_rays_host, render_rays_loss, sample_pdf, searchsorted, volume_render

__all__ = [
    "Embedding", "NeRF", "render_rays", "render_rays_loss", "render_rays_host", "FusedAdam", "invalidate_packed", "sample_pdf", "searchsorted", "volume_render",
    "nerf_forward_fused", "nerf_forward_torch", "nerf_parameters", "packed_weights",
    "batched_inference", "generate_rays", "render_image", "to_uint8", "query_sigma", "mse_psnr",
]
__version__ = "0.1.0"
