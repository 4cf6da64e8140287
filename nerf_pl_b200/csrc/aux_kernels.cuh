// Stand-alone kernels for the individual reference functions on the hot path
//  and the weight packer.  The render kernel
// fuses all of them; these entries exist so each row has its own parity test and so callers
// that use the pieces directly (extract_color_mesh.py:127-140) have a drop-in.
#pragma once
#include <cuda_bf16.h>

#include "render_kernel.cuh"

namespace nerfb200 {

// ------------------------------------------------------------ weight packer
struct PackParams {
  const float* p[kNumParams];   // device pointers, order in layout.h
  uint8_t* out;
  int bwd_bf16;                 // element type of the backward region: 1 = bf16, 0 = fp16
};

struct PackParams2 {
  PackParams net[2];            // blockIdx.y selects the network: a training step packs both images in one launch
};

// One thread per fp16 element of the slice region, then the fp32 tail.
__global__ void pack_weights_kernel(const __grid_constant__ PackParams2 pp2) {
  const PackParams& pp = pp2.net[blockIdx.y];
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n_half = kHalfRegionBytes / 2;
  if (idx < n_half) {
    int slice, n, k, N;
    const long long n256 = static_cast<long long>(kNumSlices256) * 256 * 64;
    if (idx < n256) {
      slice = static_cast<int>(idx / (256 * 64));
      const int rem = static_cast<int>(idx % (256 * 64));
      n = rem / 64; k = rem % 64; N = 256;
    } else {
      const long long j = idx - n256;
      slice = kNumSlices256 + static_cast<int>(j / (128 * 64));
      const int rem = static_cast<int>(j % (128 * 64));
      n = rem / 64; k = rem % 64; N = 128;
    }
    // slice -> (weight tensor, input-feature offset, in_features, valid k)
    const float* W; int ld, koff, kvalid;
    if (slice == 0) { W = pp.p[0]; ld = 63; koff = 0; kvalid = 63; }
    else if (slice <= 12) { const int l = 1 + (slice - 1) / 4; W = pp.p[2 * l]; ld = 256; koff = ((slice - 1) % 4) * 64; kvalid = 64; }
    else if (slice == 13) { W = pp.p[8]; ld = 319; koff = 0; kvalid = 63; }
    else if (slice <= 17) { W = pp.p[8]; ld = 319; koff = 63 + (slice - 14) * 64; kvalid = 64; }
    else if (slice <= 29) { const int l = 5 + (slice - 18) / 4; W = pp.p[2 * l]; ld = 256; koff = ((slice - 18) % 4) * 64; kvalid = 64; }
    else if (slice <= 33) { W = nullptr; ld = 0; koff = (slice - 30) * 64; kvalid = 64; }   // fused W' (below)
    else { W = pp.p[18]; ld = 283; koff = 256; kvalid = 27; }
    float v = 0.f;
    if (W == nullptr) {
      // W'[n][koff+k] = sum_m W_dir[n][m] * W_final[m][koff+k]   (fp32, layout.h)
      const float* wd = pp.p[18] + static_cast<long long>(n) * 283;
      const float* wf = pp.p[16] + koff + k;
      float acc = 0.f;
#pragma unroll 16
      for (int m = 0; m < 256; ++m) acc = fmaf(wd[m], wf[static_cast<long long>(m) * 256], acc);
      v = acc;
    } else if (k < kvalid) {
      v = W[static_cast<long long>(n) * ld + koff + k];
    }
    const uint32_t base = (slice < kNumSlices256) ? slice * kSliceBytes256
                                                   : kOffDir + (slice - kNumSlices256) * kSliceBytes128;
    (void)N;
    *reinterpret_cast<__half*>(pp.out + base + sw128_off(n, k)) = __float2half_rn(v);
    return;
  }
  const long long f = idx - n_half;
  if (f >= kF32Count) {
    // ---- backward region (layout.h): transposed slices B[n][k] = W[k0 + k][n0 + n], 16-bit
    const long long e = f - kF32Count;
    if (e >= static_cast<long long>(kNumSlicesBwd) * 256 * 64) return;
    const int slice = static_cast<int>(e / (256 * 64));
    const int rem = static_cast<int>(e % (256 * 64));
    const int k = rem / 256, n = rem % 256;       // n fastest: the reads below are coalesced over n
    float v;
    if (slice < 2) {
      // W'[m][n] = sum_j W_dir[m][j] W_final[j][n], m = slice * 64 + k
      const float* wd = pp.p[18] + static_cast<long long>(slice * 64 + k) * 283;
      const float* wf = pp.p[16] + n;
      float acc = 0.f;
#pragma unroll 16
      for (int j = 0; j < 256; ++j) acc = fmaf(wd[j], wf[static_cast<long long>(j) * 256], acc);
      v = acc;
    } else {
      const int step = (slice - 2) / 4, kb = (slice - 2) % 4;
      const int L = 8 - step;                                   // xyz_encoding_L, L = 8 .. 2
      const int ld = (L == 5) ? 319 : 256, n0 = (L == 5) ? 63 : 0;
      v = pp.p[2 * (L - 1)][static_cast<long long>(kb * 64 + k) * ld + n0 + n];
    }
    uint16_t bits;
    if (pp.bwd_bf16) bits = __bfloat16_as_ushort(__float2bfloat16_rn(v));
    else bits = __half_as_ushort(__float2half_rn(v));
    *reinterpret_cast<uint16_t*>(pp.out + kOffBwd + static_cast<uint32_t>(slice) * kSliceBytes256 + sw128_off(n, k)) = bits;
    return;
  }
  float* o = reinterpret_cast<float*>(pp.out + kHalfRegionBytes);
  float v = 0.f;
  const int i = static_cast<int>(f);
  if (i < kF32WSigma) {
    const int l = i / 256, n = i % 256;
    if (l < 8) v = pp.p[2 * l + 1][n];
    else if (n < 128) {
      // b'[n] = b_dir[n] + sum_m W_dir[n][m] * b_final[m]
      float acc = pp.p[19][n];
      const float* wd = pp.p[18] + static_cast<long long>(n) * 283;
      for (int m = 0; m < 256; ++m) acc = fmaf(wd[m], pp.p[17][m], acc);
      v = acc;
    }
  } else if (i < kF32BSigma) v = pp.p[20][i - kF32WSigma];
  else if (i < kF32WRgb) v = (i == kF32BSigma) ? pp.p[21][0] : 0.f;
  else if (i < kF32BRgb) v = pp.p[22][i - kF32WRgb];
  else if (i < kF32WDirPart) v = (i - kF32BRgb < 3) ? pp.p[23][i - kF32BRgb] : 0.f;
  else {
    const int j = i - kF32WDirPart;
    const int k = j / 128, n = j % 128;          // transposed: coalesced over n
    v = (k < 27) ? pp.p[18][static_cast<long long>(n) * 283 + 256 + k] : 0.f;
  }
  o[i] = v;
}

// ------------------------------------------------ NeRF.forward (models/nerf.py:83-124)
struct MlpParams {
  int raw_xyz;             // 1: x is (n, x_stride>=3) raw positions, encoded in-kernel; sigma only
  const float* x;          // (n, x_stride): embedded xyz (63) [+ embedded dir (27)]
  long long x_stride;
  long long n;
  const uint8_t* net;
  int sigma_only;
  float* out;              // (n, 4) rgb,sigma  or (n, 1) sigma
  int* status;
};

__global__ void __launch_bounds__(kThreads, 1) mlp_forward_kernel(const MlpParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Scratch* sc = reinterpret_cast<Scratch*>(smem + kSmemScratch);
  Barriers* bars = &sc->bars;
  if (!engine_setup(smem, bars)) {
    if (threadIdx.x == 0) report_fault(p.status, 101);
    return;
  }
  load_consts(smem, 0, p.net);
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long n_tiles = (p.n + 127) / 128;
  const bool so = p.sigma_only != 0;
  if (warp == kProducerWarp) {
    if (lane == 0) {
      RingState rs;
      for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) produce_tile(rs, smem, bars, p.net, so, !so);
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0) {
      RingState rs;
      MmaPhases ph;
      for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) mma_tile(rs, ph, smem, bars, so, !so);
    }
  } else {
    EpiCtx c;
    c.smem = smem; c.bars = bars; c.lane = lane;
    c.row = (warp & 3) * 32 + lane;
    c.part = warp >> 2;
    c.tmem_row = bars->tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    c.d_phase = 0;
    c.save_act = nullptr; c.save_d = nullptr; c.save_n = 0; c.save_row = -1;
    c.tl = nullptr;
    c.f32 = reinterpret_cast<const float*>(p.net + kHalfRegionBytes);
    c.cst = consts_ptr(smem, 0);
    uint8_t* enc = smem + kSmemEnc;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long gi = tile * 128 + c.row;
      const bool valid = gi < p.n;
      const float* xr = p.x + (valid ? gi : (p.n - 1)) * p.x_stride;
      if (p.raw_xyz) {
        // dense-grid sigma query (extract_color_mesh.py:127-140): encode the raw position here
        const float o[3] = {__ldg(xr), __ldg(xr + 1), __ldg(xr + 2)};
        const float zero[3] = {0.f, 0.f, 0.f};
        encode_row(enc, c.row, c.part, o, zero, 0.f);
      } else {
        const int k0 = c.part * (64 / kColSplit);
        for (int k = k0; k < k0 + 64 / kColSplit; ++k) {
          const float v = (k < kEncXyz) ? __ldg(xr + k) : 0.f;
          *reinterpret_cast<__half*>(enc + sw128_off(c.row, k)) = __float2half_rn(v);
        }
      }
      float sig_part, rgb_part[3];
      epi_run_tile(c, so, nullptr, so ? nullptr : xr + kEncXyz, sig_part, rgb_part);
      sc->sig_part[c.part][c.row] = sig_part;
      if (!so) {
        sc->rgb_part[c.part][0][c.row] = rgb_part[0];
        sc->rgb_part[c.part][1][c.row] = rgb_part[1];
        sc->rgb_part[c.part][2][c.row] = rgb_part[2];
      }
      epi_bar();
      if (c.part == 0 && valid) {
        float sg = c.cst[kF32BSigma];
        float pre[3] = {c.cst[kF32BRgb + 0], c.cst[kF32BRgb + 1], c.cst[kF32BRgb + 2]};
        for (int q = 0; q < kColSplit; ++q) {
          sg += sc->sig_part[q][c.row];
          if (!so) { pre[0] += sc->rgb_part[q][0][c.row]; pre[1] += sc->rgb_part[q][1][c.row]; pre[2] += sc->rgb_part[q][2][c.row]; }
        }
        if (so) {
          p.out[gi] = sg;
        } else {
          float4 o;
          o.x = sigmoid_ref(pre[0]);
          o.y = sigmoid_ref(pre[1]);
          o.z = sigmoid_ref(pre[2]);
          o.w = sg;
          *reinterpret_cast<float4*>(p.out + gi * 4) = o;
        }
      }
      epi_bar();
    }
  }
  engine_teardown(bars);
}

// ------------------------------------------------ Embedding.forward (models/nerf.py:21-38)
// x: (n, 3) -> out: (n, 3 + 6*n_freqs), channel order [x, sin f0 x, cos f0 x, ...].
__global__ void embed_kernel(const float* __restrict__ x, long long n, int n_freqs,
                             float* __restrict__ out) {
  const int C = 3 + 6 * n_freqs;
  const long long total = n * C;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / C;
    const int ch = static_cast<int>(idx - r * C);
    float v;
    if (ch < 3) {
      v = x[r * 3 + ch];
    } else {
      const int j = ch - 3;
      const int k = j / 6, fn = (j % 6) / 3, c = j % 3;
      const float a = __fmul_rn(exp2f(static_cast<float>(k)), x[r * 3 + c]);
      v = fn ? cosf(a) : sinf(a);
    }
    out[idx] = v;
  }
}

// ------------------------------------------------ searchsorted (torchsearchsorted)
// res[r, c] = #{k : a[r,k] <  v[r,c]} (side='left') or #{k : a[r,k] <= v[r,c]} ('right');
// a or v may have a single row that is broadcast (searchsorted.py:23-35, kernel.cu:83-107).
__global__ void searchsorted_kernel(const float* __restrict__ a, const float* __restrict__ v,
                                    long long* __restrict__ res, long long nrow_a, long long nrow_v,
                                    int ncol_a, int ncol_v, int side_right) {
  const long long nrow = nrow_a > nrow_v ? nrow_a : nrow_v;
  const long long total = nrow * ncol_v;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / ncol_v;
    const int c = static_cast<int>(idx - r * ncol_v);
    const float* ar = a + (nrow_a == 1 ? 0 : r) * ncol_a;
    const float val = v[(nrow_v == 1 ? 0 : r) * ncol_v + c];
    int lo = 0, hi = ncol_a;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const float x = __ldg(ar + mid);
      const bool go_right = side_right ? (x <= val) : (x < val);
      if (go_right) lo = mid + 1; else hi = mid;
    }
    res[idx] = lo;
  }
}

// ------------------------------------------------ sample_pdf (models/rendering.py:14-55)
// bins (R, nb = nw+1), weights (R, nw), u (R, K) sorted or not; out (R, K).  One warp per ray.
__global__ void sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights,
                                  const float* __restrict__ u, long long n_rays, int nw, int K,
                                  float* __restrict__ out) {
  extern __shared__ float sh[];
  const int wpb = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* cdf = sh + warp * (nw + 1);
  for (long long r = static_cast<long long>(blockIdx.x) * wpb + warp; r < n_rays;
       r += static_cast<long long>(gridDim.x) * wpb) {
    const float* w = weights + r * nw;
    const float* b = bins + r * (nw + 1);
    float part = 0.f;
    for (int i = lane; i < nw; i += 32) part += __fadd_rn(w[i], 1e-5f);
    const float total = warp_sum(part);
    if (lane == 0) {
      float run = 0.f;
      cdf[0] = 0.f;
      for (int i = 0; i < nw; ++i) {
        run = __fadd_rn(run, __fdiv_rn(__fadd_rn(w[i], 1e-5f), total));
        cdf[i + 1] = run;
      }
    }
    __syncwarp();
    for (int j = lane; j < K; j += 32) {
      const float uj = u[r * K + j];
      int lo = 0, hi = nw + 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] <= uj) lo = mid + 1; else hi = mid;
      }
      const int below = max(lo - 1, 0), above = min(lo, nw);
      const float c0 = cdf[below], c1 = cdf[above];
      float denom = __fsub_rn(c1, c0);
      if (denom < 1e-5f) denom = 1.f;
      const float t = __fdiv_rn(__fsub_rn(uj, c0), denom);
      out[r * K + j] = __fadd_rn(b[below], __fmul_rn(t, __fsub_rn(b[above], b[below])));
    }
    __syncwarp();
  }
}

// ------------------------------------------------ volume rendering (models/rendering.py:143-170)
// sigmas (R,S), rgbs (R,S,3) nullable, z (R,S), dirs (R,3), noise (R,S) nullable.
// One warp per ray.  S % 32 == 0, S <= 192.
__global__ void composite_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                 const float* __restrict__ z, const float* __restrict__ dirs,
                                 const float* __restrict__ noise, float noise_std, int white_back,
                                 long long n_rays, int S, float* __restrict__ weights,
                                 float* __restrict__ rgb_out, float* __restrict__ depth_out,
                                 float* __restrict__ opac_out) {
  extern __shared__ float sh[];
  const int wpb = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* base = sh + warp * (6 * S);
  float *sz = base, *ss = base + S, *sr = base + 2 * S, *sg = base + 3 * S, *sb = base + 4 * S,
        *sw = base + 5 * S;
  for (long long r = static_cast<long long>(blockIdx.x) * wpb + warp; r < n_rays;
       r += static_cast<long long>(gridDim.x) * wpb) {
    for (int i = lane; i < S; i += 32) {
      sz[i] = z[r * S + i];
      ss[i] = sigmas[r * S + i];
      if (rgbs != nullptr) {
        sr[i] = rgbs[(r * S + i) * 3 + 0];
        sg[i] = rgbs[(r * S + i) * 3 + 1];
        sb[i] = rgbs[(r * S + i) * 3 + 2];
      }
    }
    __syncwarp();
    const float dx = dirs[r * 3], dy = dirs[r * 3 + 1], dz = dirs[r * 3 + 2];
    const float dn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    const RayOut o = composite_ray(lane, S, sz, ss, sr, sg, sb, noise ? noise + r * S : nullptr,
                                   noise_std, dn, rgbs != nullptr, sw);
    __syncwarp();
    if (weights != nullptr)
      for (int i = lane; i < S; i += 32) weights[r * S + i] = sw[i];
    if (lane == 0) {
      opac_out[r] = o.opac;
      if (rgbs != nullptr) {
        const float add = white_back ? __fsub_rn(1.f, o.opac) : 0.f;
        rgb_out[r * 3 + 0] = o.r + add;
        rgb_out[r * 3 + 1] = o.g + add;
        rgb_out[r * 3 + 2] = o.b + add;
        depth_out[r] = o.depth;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------ ray generation (datasets/ray_utils.py:5-94)
// One thread per pixel: get_ray_directions (:16-22, no +0.5 pixel centre), get_rays (:41-46:
// rotate by c2w[:, :3], normalise, origin = c2w[:, 3]) and optionally get_ndc_rays (:75-92, as
// datasets/llff.py:236-241 applies it: near plane 1.0, then near/far columns 0/1).
// Writes the (H*W, 8) ray rows [o, d, near, far] the renderer consumes, so rays never cross PCIe.
struct RayGenParams {
  int H, W;
  float focal;
  float c2w[12];      // row-major (3, 4)
  float near, far;
  int ndc;
  float* rays;
};
__global__ void generate_rays_kernel(const RayGenParams p) {
  const long long total = static_cast<long long>(p.H) * p.W;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(idx / p.W), i = static_cast<int>(idx - static_cast<long long>(j) * p.W);
    const float dx = __fdiv_rn(static_cast<float>(i) - 0.5f * p.W, p.focal);
    const float dy = -__fdiv_rn(static_cast<float>(j) - 0.5f * p.H, p.focal);
    const float dz = -1.f;
    float d[3], o[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      d[r] = __fadd_rn(__fadd_rn(__fmul_rn(dx, p.c2w[4 * r + 0]), __fmul_rn(dy, p.c2w[4 * r + 1])),
                       __fmul_rn(dz, p.c2w[4 * r + 2]));
      o[r] = p.c2w[4 * r + 3];
    }
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
#pragma unroll
    for (int r = 0; r < 3; ++r) d[r] = __fdiv_rn(d[r], nrm);
    float near = p.near, far = p.far;
    if (p.ndc) {
      const float n1 = 1.0f;                                   // llff.py:238 near plane at 1.0
      const float tt = -__fdiv_rn(__fadd_rn(n1, o[2]), d[2]);
#pragma unroll
      for (int r = 0; r < 3; ++r) o[r] = __fadd_rn(o[r], __fmul_rn(tt, d[r]));
      const float ox_oz = __fdiv_rn(o[0], o[2]), oy_oz = __fdiv_rn(o[1], o[2]);
      const float sx = -1.f / (p.W / (2.f * p.focal)), sy = -1.f / (p.H / (2.f * p.focal));
      const float o0 = sx * ox_oz, o1 = sy * oy_oz, o2 = 1.f + 2.f * n1 / o[2];
      const float d0 = sx * (__fdiv_rn(d[0], d[2]) - ox_oz), d1 = sy * (__fdiv_rn(d[1], d[2]) - oy_oz);
      const float d2 = 1.f - o2;
      o[0] = o0; o[1] = o1; o[2] = o2; d[0] = d0; d[1] = d1; d[2] = d2;
      near = 0.f; far = 1.f;
    }
    float4* out = reinterpret_cast<float4*>(p.rays + idx * 8);
    out[0] = make_float4(o[0], o[1], o[2], d[0]);
    out[1] = make_float4(d[1], d[2], near, far);
  }
}

// ------------------------------------------------ float image -> uint8 (eval.py:126-128)
// img_pred_ = (clip(img_pred, 0, 1) * 255).astype(uint8)  (truncation, as numpy's astype does)
__global__ void to_uint8_kernel(const float* __restrict__ src, long long n, uint8_t* __restrict__ dst) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = fminf(fmaxf(src[i], 0.f), 1.f) * 255.f;
    dst[i] = static_cast<uint8_t>(v);
  }
}

// ------------------------------------------------ loss / metric epilogue (losses.py:9-14, metrics.py:4-13)
// out[0] = mean((rgb_coarse - t)^2), out[1] = mean((rgb_fine - t)^2) (0 if rgb_fine is null),
// out[2] = out[0] + out[1] (MSELoss.forward), out[3] = -10 log10(mse of the finest available)  (psnr).
// One block, fixed summation order (deterministic).
__global__ void __launch_bounds__(1024, 1) mse_psnr_kernel(const float* __restrict__ rgb_c,
                                                           const float* __restrict__ rgb_f,
                                                           const float* __restrict__ target, long long n_elem,
                                                           float* __restrict__ out) {
  __shared__ double red[2][32];
  double ac = 0.0, af = 0.0;
  for (long long i = threadIdx.x; i < n_elem; i += blockDim.x) {
    const float t = target[i];
    if (rgb_c != nullptr) { const float d = rgb_c[i] - t; ac += static_cast<double>(d) * d; }
    if (rgb_f != nullptr) { const float d = rgb_f[i] - t; af += static_cast<double>(d) * d; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    ac += __shfl_xor_sync(0xffffffffu, ac, o);
    af += __shfl_xor_sync(0xffffffffu, af, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ac; red[1][threadIdx.x >> 5] = af; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sc = 0.0, sf = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { sc += red[0][w]; sf += red[1][w]; }
    const float mc = static_cast<float>(sc / static_cast<double>(n_elem));
    const float mf = static_cast<float>(sf / static_cast<double>(n_elem));
    out[0] = mc; out[1] = mf; out[2] = mc + mf;
    out[3] = -10.f * log10f(rgb_f != nullptr ? mf : mc);
  }
}

}  // namespace nerfb200
