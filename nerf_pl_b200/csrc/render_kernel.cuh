// Fused render_rays (models/rendering.py:58-244): stratified depths -> positional encoding
// -> coarse MLP -> alpha compositing -> inverse-CDF resampling -> merge -> fine MLP ->
// compositing, one persistent CTA per SM, two rays ("a group") at a time.  sigma / rgb never
// leave the SM; HBM traffic is 32 B in and <= 40 B out per ray.
#pragma once
#include "mlp_engine.cuh"

namespace nerfb200 {

constexpr int kMaxSc = 64;      // max coarse samples / ray
constexpr int kMaxImp = 128;    // max importance samples / ray
constexpr int kMaxSf = 192;     // max fine samples / ray (N_samples + N_importance)
constexpr int kMaxRows = 2 * kMaxSf;

struct RenderParams {
  const float* rays;            // (n_rays, ray_stride) : o(3) d(3) near far
  long long ray_stride;         // in floats
  int n_rays;
  const uint8_t* net_coarse;    // packed images (layout.h)
  const uint8_t* net_fine;      // may be null when n_importance == 0
  int n_samples;                // S_c
  int n_importance;
  int use_disp;
  float perturb;
  float noise_std;
  int white_back;
  int test_time;
  const float* perturb_rand;    // (n_rays, S_c)   U[0,1)   required iff perturb > 0
  const float* noise_coarse;    // (n_rays, S_c)   N(0,1)   required iff noise_std > 0
  const float* noise_fine;      // (n_rays, S_f)   N(0,1)   required iff noise_std > 0 && fine
  const float* u_rand;          // (n_rays, N_imp) U[0,1)   required iff perturb > 0 && fine
  float* rgb_coarse;            // (n_rays,3) nullable
  float* depth_coarse;          // (n_rays)   nullable
  float* opacity_coarse;        // (n_rays)
  float* rgb_fine;              // (n_rays,3)
  float* depth_fine;            // (n_rays)
  float* opacity_fine;          // (n_rays)
  float* z_fine;                // (n_rays, S_f) optional: merged sorted depths
  float* weights_coarse;        // (n_rays, S_c) optional
  float* weights_fine;          // (n_rays, S_f) optional
  int* status;                  // device int: nonzero on device-detected error
  // training mode: per-sample intermediates for the backward pass (all null = inference)
  __half* save_act_c;           // [8][n_rays*S_c][256] fp16: coarse h1..h8
  __half* save_act_f;           // [8][n_rays*S_f][256]
  __half* save_d_c;             // [n_rays*S_c][128] fp16: direction-layer activation
  __half* save_d_f;             // [n_rays*S_f][128]
  float* save_sig_c;            // [n_rays*S_c] raw sigma
  float* save_sig_f;            // [n_rays*S_f]
  float* save_rgb_c;            // [n_rays*S_c][3] sigmoid(rgb)
  float* save_rgb_f;            // [n_rays*S_f][3]
  unsigned flags;               // experiment switches (NERFB200_FLAGS), 0 in production
  long long* timeline;          // experiment: device timeline buffer (flags & 2), else null
};

struct alignas(16) Scratch {
  Barriers bars;                       //   96
  alignas(16) float dirbias[2][kDirW]; // 1024   per-ray b_dir + W_dir[:,256:283] . dir_enc
  float sig_part[kColSplit][128];      // [column group][row] partial sigma-head sums
  float rgb_part[kColSplit][3][128];   // partial rgb-head sums
  float ray[2][8];                     //   64
  float dnorm[2];
  float pad0[2];
  float direnc[2][28];                 //  224
  float z[kMaxRows];                   // 1536   depths of the current pass, [ray][S]
  float sigma[kMaxRows];               // 1536   sigma of the current pass; overwritten in place by the
                                       //        compositing weights (same index, same lane)
  float rgb[3][kMaxRows];              // 4608
  float cdf[2][kMaxSc];                //  512
  float znew[2][kMaxImp];              // 1024   u (sorted) then the new depths
  float zc[2][kMaxSc];                 //  512   coarse depths kept for the merge
};
static_assert(sizeof(Scratch) <= kScratchBytes, "scratch does not fit");

// torch.linspace(0, 1, n)[i] in fp32 (symmetric two-sided evaluation).
__device__ __forceinline__ float linspace01(int i, int n) {
  if (n <= 1) return 0.f;
  const float step = __fdiv_rn(1.f, static_cast<float>(n - 1));
  return (i < n / 2) ? __fmul_rn(step, static_cast<float>(i))
                     : __fsub_rn(1.f, __fmul_rn(step, static_cast<float>(n - 1 - i)));
}

// models/rendering.py:189-193
__device__ __forceinline__ float z_base(float near, float far, int i, int n, bool use_disp) {
  const float t = linspace01(i, n);
  const float omt = __fsub_rn(1.f, t);
  if (!use_disp) return __fadd_rn(__fmul_rn(near, omt), __fmul_rn(far, t));
  const float a = __fmul_rn(__fdiv_rn(1.f, near), omt);
  const float b = __fmul_rn(__fdiv_rn(1.f, far), t);
  return __fdiv_rn(1.f, __fadd_rn(a, b));
}

// Embedding.forward (models/nerf.py:33-38) of the 3 values x into out[3 + 6*n_freqs].
__device__ __forceinline__ void embed3(const float x[3], int n_freqs, float* out) {
  out[0] = x[0]; out[1] = x[1]; out[2] = x[2];
  float f = 1.f;
  for (int k = 0; k < n_freqs; ++k, f *= 2.f) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s, co;
      sincosf(f * x[c], &s, &co);
      out[3 + 6 * k + c] = s;
      out[3 + 6 * k + 3 + c] = co;
    }
  }
}

// sin / cos of arguments up to a few thousand radians (2^9 * |x|): two-constant Cody-Waite
// reduction by 2 pi (6.28125 is exact in 8 bits, so n * 6.28125 is exact for n < 2^16), then the
// MUFU approximations on [-pi, pi] (abs error 2^-21.4).  Total abs error ~1e-6, three orders of
// magnitude below the fp16 rounding (4.9e-4) the encoded features receive anyway.
__device__ __forceinline__ void fast_sincos(float a, float& s, float& c) {
  const float n = rintf(a * 0.15915494309189535f);
  float r = fmaf(n, -6.28125f, a);
  r = fmaf(n, -1.9353071795864769e-3f, r);
  s = __sinf(r);
  c = __cosf(r);
}

// Encoded xyz of one sample row into the ENC tile; the kColSplit threads of a row split the ten
// frequencies.  Feature order models/nerf.py:33-38: [x, sin f0 x, cos f0 x, sin f1 x, ..].
__device__ __forceinline__ void encode_row(uint8_t* enc, int row, int part, const float o[3],
                                           const float d[3], float z) {
  float x[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(o[c], __fmul_rn(d[c], z));   // rendering.py:206
  if (part == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      *reinterpret_cast<__half*>(enc + sw128_off(row, c)) = __float2half_rn(x[c]);
  } else if (part == kColSplit - 1) {
    *reinterpret_cast<__half*>(enc + sw128_off(row, 63)) = __float2half_rn(0.f);
  }
  // frequencies [k0, k1): 2 groups -> 5+5, 4 groups -> 3+3+2+2
  const int k0 = (kColSplit == 2) ? part * 5 : (part < 2 ? part * 3 : 6 + (part - 2) * 2);
  const int k1 = (kColSplit == 2) ? k0 + 5 : (part < 2 ? k0 + 3 : k0 + 2);
  float f = static_cast<float>(1 << k0);
  for (int k = k0; k < k1; ++k, f *= 2.f) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s, co;
      fast_sincos(f * x[c], s, co);
      *reinterpret_cast<__half*>(enc + sw128_off(row, 3 + 6 * k + c)) = __float2half_rn(s);
      *reinterpret_cast<__half*>(enc + sw128_off(row, 3 + 6 * k + 3 + c)) = __float2half_rn(co);
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Volume-rendering quadrature for one ray by one warp (models/rendering.py:143-170).
// Each lane owns P = S/32 consecutive samples.  Writes weights to w[0..S).
struct RayOut { float r, g, b, depth, opac; };
__device__ __forceinline__ RayOut composite_ray(int lane, int S, const float* z, const float* sigma,
                                                const float* r, const float* g, const float* b,
                                                const float* __restrict__ noise, float noise_std,
                                                float dnorm, bool want_rgb, float* w) {
  const int P = S >> 5;          // 2..6
  float alpha[6], tloc[6];
  float prod = 1.f;
  for (int p = 0; p < P; ++p) {
    const int i = lane * P + p;
    float delta = (i < S - 1) ? __fsub_rn(z[i + 1], z[i]) : 1e10f;
    delta = __fmul_rn(delta, dnorm);
    float s = sigma[i];
    if (noise != nullptr) s = __fadd_rn(s, __fmul_rn(noise[i], noise_std));
    const float a = __fsub_rn(1.f, expf(-__fmul_rn(delta, fmaxf(s, 0.f))));
    alpha[p] = a;
    tloc[p] = prod;
    prod = __fmul_rn(prod, __fadd_rn(__fsub_rn(1.f, a), 1e-10f));
  }
  // exclusive multiplicative scan of the per-lane products
  float incl = prod;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= v;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 1.f;
  RayOut out{0.f, 0.f, 0.f, 0.f, 0.f};
  for (int p = 0; p < P; ++p) {
    const int i = lane * P + p;
    const float wi = alpha[p] * (excl * tloc[p]);
    w[i] = wi;
    out.opac += wi;
    if (want_rgb) {
      out.r = fmaf(wi, r[i], out.r);
      out.g = fmaf(wi, g[i], out.g);
      out.b = fmaf(wi, b[i], out.b);
      out.depth = fmaf(wi, z[i], out.depth);
    }
  }
  out.opac = warp_sum(out.opac);
  if (want_rgb) {
    out.r = warp_sum(out.r); out.g = warp_sum(out.g); out.b = warp_sum(out.b);
    out.depth = warp_sum(out.depth);
  }
  return out;
}

// sample_pdf, first half (models/rendering.py:28-33) for one ray by one warp:
//   weights = w[1..S-2] + 1e-5 -> pdf -> cdf[0..S-2] (cdf[0] = 0), S-1 entries.
__device__ __forceinline__ void pdf_to_cdf_ray(int lane, int S, const float* w, float* cdf) {
  const int nw = S - 2;            // N_samples_
  float part = 0.f;
  for (int i = lane; i < nw; i += 32) part += __fadd_rn(w[1 + i], 1e-5f);
  const float total = warp_sum(part);
  // per-lane contiguous segments + warp scan
  const int per = (nw + 31) >> 5;
  float loc[4];
  float run = 0.f;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (p < per) {
      const int i = lane * per + p;
      const float pdf = (i < nw) ? __fdiv_rn(__fadd_rn(w[1 + i], 1e-5f), total) : 0.f;
      run += pdf;
      loc[p] = run;
    }
  }
  float incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) { excl = 0.f; cdf[0] = 0.f; }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (p < per) {
      const int i = lane * per + p;
      if (i < nw) cdf[i + 1] = excl + loc[p];
    }
  }
}

// sample_pdf, second half (models/rendering.py:42-54) for one sample position u:
//   inds = searchsorted(cdf, u, 'right'); below/above clamps; linear interpolation between the
//   bin mid-points (bins[k] = 0.5 (z[k] + z[k+1])).
__device__ __forceinline__ float inverse_cdf(int S, const float* zc, const float* cdf, float u) {
  const int nw = S - 2;
  int lo = 0, hi = nw + 1;         // S-1 cdf entries, last valid index nw
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  const int below = max(lo - 1, 0);
  const int above = min(lo, nw);
  const float c0 = cdf[below], c1 = cdf[above];
  const float b0 = 0.5f * __fadd_rn(zc[below], zc[below + 1]);
  const float b1 = 0.5f * __fadd_rn(zc[above], zc[above + 1]);
  float denom = __fsub_rn(c1, c0);
  if (denom < 1e-5f) denom = 1.f;
  const float tt = __fdiv_rn(__fsub_rn(u, c0), denom);
  return __fadd_rn(b0, __fmul_rn(tt, __fsub_rn(b1, b0)));
}

template <bool kSave>
__global__ void __launch_bounds__(kThreads, 1) render_rays_kernel(const RenderParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Scratch* sc = reinterpret_cast<Scratch*>(smem + kSmemScratch);
  Barriers* bars = &sc->bars;
#ifdef NERFB200_TIMELINE
  const long long t_entry = clock64();
#endif
  if (!engine_setup(smem, bars)) {
    if (threadIdx.x == 0) atomicExch(p.status, 101);
    return;
  }
  load_consts(smem, 0, p.net_coarse);
  load_consts(smem, 1, p.net_fine);
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int Sc = p.n_samples;
  const int K = p.n_importance;
  const int Sf = Sc + K;
  const bool fine = K > 0;
  const bool coarse_sigma_only = p.test_time != 0;
  // Balanced contiguous ray ranges: CTA b renders rays [my_lo, my_lo + my_n), two at a time; an odd
  // count ends with a single-ray group that only runs the tiles it needs.
  const int per_cta = p.n_rays / static_cast<int>(gridDim.x), rem_cta = p.n_rays % static_cast<int>(gridDim.x);
  const int my_lo = static_cast<int>(blockIdx.x) * per_cta + min(static_cast<int>(blockIdx.x), rem_cta);
  const int my_n = per_cta + (static_cast<int>(blockIdx.x) < rem_cta ? 1 : 0);
  const int n_groups = (my_n + 1) >> 1;
  auto rays_in_group = [&](int g) { return min(2, my_n - 2 * g); };
  auto tiles_of = [](int n_r, int S) { return (n_r * S + 127) >> 7; };

  if (warp == kProducerWarp) {
    if (lane == 0) {
      RingState rs;
      for (int g = 0; g < n_groups; ++g) {
        const int nr = rays_in_group(g);
        for (int t = 0; t < tiles_of(nr, Sc); ++t)
          produce_tile(rs, smem, bars, p.net_coarse, coarse_sigma_only, false);
        if (fine)
          for (int t = 0; t < tiles_of(nr, Sf); ++t) produce_tile(rs, smem, bars, p.net_fine, false, false);
      }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0) {
      RingState rs;
      MmaPhases ph;
      Timeline tlm{(blockIdx.x == 0) ? p.timeline : nullptr, {0, 0, 0}};
      for (int g = 0; g < n_groups; ++g) {
        const int nr = rays_in_group(g);
        for (int t = 0; t < tiles_of(nr, Sc); ++t) mma_tile(rs, ph, smem, bars, coarse_sigma_only, false, &tlm);
        if (fine)
          for (int t = 0; t < tiles_of(nr, Sf); ++t) mma_tile(rs, ph, smem, bars, false, false, &tlm);
      }
    }
  } else {
    EpiCtx c;
    c.smem = smem;
    c.bars = bars;
    c.lane = lane;
    c.row = (warp & 3) * 32 + lane;
    c.part = warp >> 2;
    c.tmem_row = bars->tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    c.d_phase = 0;
    c.flags = p.flags;
    c.save_act = nullptr; c.save_d = nullptr; c.save_n = 0; c.save_row = -1;
    Timeline tle{(blockIdx.x == 0 && threadIdx.x == 0) ? p.timeline : nullptr, {0, 0, 0}};
    c.tl = &tle;
#ifdef NERFB200_TIMELINE
    tl_val(c.tl, 0, 90, t_entry);
#endif
    NERFB200_TL_MARK(c.tl, 0, 91);
    const int t = threadIdx.x;   // 0..255
    uint8_t* enc = smem + kSmemEnc;

    for (int g = 0; g < n_groups; ++g) {
      const int ray0 = my_lo + 2 * g;
      const bool valid1 = rays_in_group(g) == 2;
      const int rid[2] = {ray0, valid1 ? ray0 + 1 : ray0};
      // ---- rays, direction embedding (models/rendering.py:179-186)
      NERFB200_TL_MARK(c.tl, 0, 30);
      if (t < 16) sc->ray[t >> 3][t & 7] = __ldg(p.rays + static_cast<long long>(rid[t >> 3]) * p.ray_stride + (t & 7));
      epi_bar();
      if (t < 2) {
        const float* d = &sc->ray[t][3];
        sc->dnorm[t] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
      } else if (t >= 32 && t < 32 + 30) {
        // Embedding(3,4)(rays_d) (models/rendering.py:186): one (ray, coord, freq) per thread, accurate sincosf
        const int q = t - 32, r = q / 15, cc = (q % 15) / 5, k = q % 5;     // k == 4: the raw value
        const float dv = sc->ray[r][3 + cc];
        if (k == 4) {
          sc->direnc[r][cc] = dv;
        } else {
          float sn, cs;
          sincosf(__fmul_rn(static_cast<float>(1 << k), dv), &sn, &cs);
          sc->direnc[r][3 + 6 * k + cc] = sn;
          sc->direnc[r][3 + 6 * k + 3 + cc] = cs;
        }
      }
      // ---- coarse depths (models/rendering.py:189-204)
      for (int e = t; e < 2 * Sc; e += kEpiThreads) {
        const int r = e / Sc, i = e - r * Sc;
        const float nr = sc->ray[r][6], fr = sc->ray[r][7];
        float z = z_base(nr, fr, i, Sc, p.use_disp != 0);
        if (p.perturb > 0.f) {
          const float zl = (i > 0) ? z_base(nr, fr, i - 1, Sc, p.use_disp != 0) : z;
          const float zu = (i < Sc - 1) ? z_base(nr, fr, i + 1, Sc, p.use_disp != 0) : z;
          const float lower = (i > 0) ? __fmul_rn(0.5f, __fadd_rn(zl, z)) : z;
          const float upper = (i < Sc - 1) ? __fmul_rn(0.5f, __fadd_rn(z, zu)) : z;
          const float pr = __fmul_rn(p.perturb, __ldg(p.perturb_rand + static_cast<long long>(rid[r]) * Sc + i));
          z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), pr));
        }
        sc->z[e] = z;
        sc->zc[r][i] = z;
      }
      NERFB200_TL_MARK(c.tl, 0, 31);
      epi_bar();

      // ================= two passes: coarse, fine =================
      for (int pass = 0; pass < (fine ? 2 : 1); ++pass) {
        const int S = pass ? Sf : Sc;
        const int tiles = tiles_of(valid1 ? 2 : 1, S);
        const bool sigma_only = (pass == 0) && coarse_sigma_only;
        const uint8_t* blob = pass ? p.net_fine : p.net_coarse;
        c.f32 = reinterpret_cast<const float*>(blob + kHalfRegionBytes);
        c.cst = consts_ptr(smem, pass);
        c.save_act = kSave ? (pass ? p.save_act_f : p.save_act_c) : nullptr;
        c.save_d = kSave ? (pass ? p.save_d_f : p.save_d_c) : nullptr;
        c.save_n = static_cast<long long>(p.n_rays) * S;
        if (!sigma_only) {
          // per-ray direction bias: b_dir + W_dir[:, 256:283] . dir_embedded   (fp32)
          if (t < 256) {
            const int r = t >> 7, n = t & 127;
            const float* wd = c.f32 + kF32WDirPart + n;           // [j][n]: coalesced over n
            float wv[27];
#pragma unroll
            for (int j = 0; j < 27; ++j) wv[j] = __ldg(wd + j * 128);
            float acc = c.cst[kF32Bias + 8 * 256 + n];
#pragma unroll
            for (int j = 0; j < 27; ++j) acc = fmaf(wv[j], sc->direnc[r][j], acc);
            sc->dirbias[r][n] = acc;
          }
          NERFB200_TL_MARK(c.tl, 0, 27);
          epi_bar();
        }
        for (int tile = 0; tile < tiles; ++tile) {
          const int gr = tile * 128 + c.row;
          const int r = gr / S;
          NERFB200_TL_MARK(c.tl, 0, 10);
          encode_row(enc, c.row, c.part, &sc->ray[r][0], &sc->ray[r][3], sc->z[gr]);
          const long long grow = (r == 0 || valid1) ? static_cast<long long>(rid[r]) * S + (gr - r * S) : -1;
          c.save_row = grow;
          float sig_part, rgb_part[3];
          epi_run_tile<kSave>(c, sigma_only, sc->dirbias[r], nullptr, sig_part, rgb_part);
          sc->sig_part[c.part][c.row] = sig_part;
          if (!sigma_only) {
            sc->rgb_part[c.part][0][c.row] = rgb_part[0];
            sc->rgb_part[c.part][1][c.row] = rgb_part[1];
            sc->rgb_part[c.part][2][c.row] = rgb_part[2];
          }
          epi_bar();
          // combine the column groups' partial head sums: group 0 -> sigma, groups 1..3 -> r, g, b
          if (c.part == 0) {
            float sg = c.cst[kF32BSigma];
#pragma unroll
            for (int q = 0; q < kColSplit; ++q) sg += sc->sig_part[q][c.row];
            sc->sigma[gr] = sg;
            if (kSave && grow >= 0) {
              float* ss = pass ? p.save_sig_f : p.save_sig_c;
              if (ss != nullptr) ss[grow] = sg;
            }
          }
          if (!sigma_only) {
            for (int ch = c.part - (kColSplit == 4 ? 1 : 0); ch < 3 && ch >= 0; ch += (kColSplit == 4 ? 3 : 1)) {
              if (kColSplit == 2 && c.part != 0) break;
              float pre = c.cst[kF32BRgb + ch];
#pragma unroll
              for (int q = 0; q < kColSplit; ++q) pre += sc->rgb_part[q][ch][c.row];
              const float col = sigmoid_ref(pre);
              sc->rgb[ch][gr] = col;
              if (kSave && grow >= 0) {
                float* sr = pass ? p.save_rgb_f : p.save_rgb_c;
                if (sr != nullptr) sr[grow * 3 + ch] = col;
              }
            }
          }
          epi_bar();
        }
        // ---- compositing: warp r renders ray r
        NERFB200_TL_MARK(c.tl, 0, 20);
        if (warp < 2) {
          const int r = warp;
          const float* nz = nullptr;
          if (p.noise_std > 0.f)
            nz = (pass ? p.noise_fine : p.noise_coarse) + static_cast<long long>(rid[r]) * S;
          const RayOut o = composite_ray(lane, S, sc->z + r * S, sc->sigma + r * S, sc->rgb[0] + r * S,
                                         sc->rgb[1] + r * S, sc->rgb[2] + r * S, nz, p.noise_std,
                                         sc->dnorm[r], !sigma_only, sc->sigma + r * S);
          __syncwarp();
          NERFB200_TL_MARK(c.tl, 0, 22);
          const bool wr = (r == 0) || valid1;
          if (wr) {
            const long long ri = rid[r];
            float* wout = pass ? p.weights_fine : p.weights_coarse;
            if (wout != nullptr)
              for (int i = lane; i < S; i += 32) wout[ri * S + i] = sc->sigma[r * S + i];
            if (lane == 0) {
              float add = (p.white_back != 0) ? __fsub_rn(1.f, o.opac) : 0.f;
              if (pass == 0) {
                p.opacity_coarse[ri] = o.opac;
                if (!sigma_only) {
                  p.rgb_coarse[3 * ri + 0] = o.r + add;
                  p.rgb_coarse[3 * ri + 1] = o.g + add;
                  p.rgb_coarse[3 * ri + 2] = o.b + add;
                  p.depth_coarse[ri] = o.depth;
                }
              } else {
                p.opacity_fine[ri] = o.opac;
                p.rgb_fine[3 * ri + 0] = o.r + add;
                p.rgb_fine[3 * ri + 1] = o.g + add;
                p.rgb_fine[3 * ri + 2] = o.b + add;
                p.depth_fine[ri] = o.depth;
              }
            }
          }
          NERFB200_TL_MARK(c.tl, 0, 23);
          // ---- hierarchical resampling, part 1 (models/rendering.py:28-33): pdf -> cdf
          if (pass == 0 && fine) pdf_to_cdf_ray(lane, Sc, sc->sigma + r * Sc, sc->cdf[r]);
          NERFB200_TL_MARK(c.tl, 0, 24);
        }
        epi_bar();
        NERFB200_TL_MARK(c.tl, 0, 21);
        if (pass == 0 && fine) {
          // ---- part 2 (:36-54): one u per thread -> inverse-CDF depth; u sorted first when random
          if (t < 2 * K) {
            const int r = t / K, j = t - r * K;
            float uj;
            int slot = j;
            if (p.perturb > 0.f) {
              const float* ur = p.u_rand + static_cast<long long>(rid[r]) * K;
              uj = __ldg(ur + j);
              slot = 0;
#pragma unroll 8
              for (int q = 0; q < K; ++q) {
                const float uq = __ldg(ur + q);
                slot += (uq < uj) || (uq == uj && q < j);
              }
            } else {
              uj = linspace01(j, K);
            }
            sc->znew[r][slot] = inverse_cdf(Sc, sc->zc[r], sc->cdf[r], uj);
          }
          epi_bar();
          // ---- merge: z_fine = sort(cat(z_coarse, z_new)) (:229) as a rank computation:
          // position = number of elements that are smaller, ties broken by index in the
          // concatenation (any tie order gives the same sorted VALUES, which is all torch.sort's
          // output carries).  Both lists are sorted except for possible 1-ulp inversions from
          // rounding, so every element first checks its predecessor; if no inversion exists
          // anywhere (the common case) ranks come from two binary searches, otherwise from
          // exhaustive counting.  Both paths give identical results on sorted input.
          {
            bool inv = false;
            for (int e = t; e < 2 * Sf; e += kEpiThreads) {
              const int r = e / Sf, i = e - r * Sf;
              if (i != 0 && i != Sc)
                inv |= (i < Sc) ? (sc->zc[r][i] < sc->zc[r][i - 1]) : (sc->znew[r][i - Sc] < sc->znew[r][i - Sc - 1]);
            }
            int any_inv;
            asm volatile(
                "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                "barrier.red.or.pred p, 1, %2, q;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                : "=r"(any_inv) : "r"(static_cast<int>(inv)), "n"(kEpiThreads) : "memory");
            for (int e = t; e < 2 * Sf; e += kEpiThreads) {
              const int r = e / Sf, i = e - r * Sf;
              const float* zc = sc->zc[r];
              const float* zn = sc->znew[r];
              const float v = (i < Sc) ? zc[i] : zn[i - Sc];
              int rank;
              if (!any_inv) {
                // lower_bound in the other list for coarse elements (coarse first on ties),
                // upper_bound for new elements
                const float* other = (i < Sc) ? zn : zc;
                int lo = 0, hi = (i < Sc) ? K : Sc;
                while (lo < hi) {
                  const int mid = (lo + hi) >> 1;
                  const float x = other[mid];
                  const bool right = (i < Sc) ? (x < v) : (x <= v);
                  if (right) lo = mid + 1; else hi = mid;
                }
                rank = lo + ((i < Sc) ? i : i - Sc);
              } else {
                rank = 0;
#pragma unroll 8
                for (int q = 0; q < Sc; ++q) rank += (zc[q] < v) || (zc[q] == v && q < i);
#pragma unroll 8
                for (int q = 0; q < K; ++q) rank += (zn[q] < v) || (zn[q] == v && (q + Sc) < i);
              }
              sc->z[r * Sf + rank] = v;
            }
          }
          NERFB200_TL_MARK(c.tl, 0, 25);
          epi_bar();
          if (p.z_fine != nullptr) {
            for (int e = t; e < 2 * Sf; e += kEpiThreads) {
              const int r = e / Sf, i = e - r * Sf;
              if (r == 0 || valid1) p.z_fine[static_cast<long long>(rid[r]) * Sf + i] = sc->z[e];
            }
          }
        }
      }
    }
    NERFB200_TL_MARK(c.tl, 0, 99);
  }
  engine_teardown(bars);
}

}  // namespace nerfb200
