// Fused render_rays (models/rendering.py:58-244): stratified depths -> positional encoding
// -> coarse MLP -> alpha compositing -> inverse-CDF resampling -> merge -> fine MLP ->
// compositing, one persistent CTA per SM.  A CTA walks its rays two at a time ("a group"); the
// MLP tiles of neighbouring groups are interleaved so that everything that is not the MLP runs on
// two helper warps concurrently with it (see the kernel's comment below).  sigma / rgb never
// leave the SM; HBM traffic is 32 B in and <= 40 B out per ray.
#pragma once
#include "mlp_engine.cuh"

namespace nerfb200 {

constexpr int kMaxSc = 128;     // max coarse samples / ray
constexpr int kMaxImp = 128;    // max importance samples / ray
constexpr int kMaxSf = 192;     // max fine samples / ray (N_samples + N_importance)
constexpr int kMaxRows = 2 * kMaxSf;

// Everything one render pass (coarse or fine) leaves behind in the training workspace / what the
// backward launches produce there (csrc/bwd_kernels.cuh).  "tiled": the layout of layout.h.
struct PassBufs {
  long long n;            // samples in the pass = n_rays * S
  long long n_pad;        // rounded up to 128
  int S;
  // written by the forward launch
  uint8_t* enc;           // tiled (n_pad, 64) fp16: encoded xyz, column 63 zero
  uint8_t* act;           // 8 x tiled (n_pad, 256) fp16: outputs of xyz_encoding_1..8 (layer l at l * n_pad * 512)
  uint2* mask;            // [8][n_pad][4]: ReLU sign bits of the 64 columns thread (row, column group) owns
  uint8_t* d;             // tiled (n_pad, 128) fp16: output of dir_encoding
  float* sigma;           // (n_pad) raw sigma
  float* rgb;             // (n_pad, 3) sigmoid(rgb)
  float* z;               // (n_rays, S) depths of the pass
  // written by the backward launches
  float* dsigma;          // (n_pad)    dL / d sigma
  float* dprergb;         // (n_pad, 3) dL / d (rgb before the sigmoid)
  uint8_t* dd;            // tiled (n_pad, 128) 16-bit: dL / d (dir_encoding pre-activation) (scaled)
  uint8_t* dpre;          // 8 x tiled (n_pad, 256) 16-bit: dL / d (xyz_encoding_l pre-activation) (scaled)
};

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
  // training mode (train != 0, requires test_time == 0): the per-sample intermediates the backward
  // needs are written into the training workspace (PassBufs of the coarse / fine pass)
  int train;
  PassBufs tr[2];
  float* z_coarse;              // (n_rays, S_c) optional: the (stratified) coarse depths
  // fused loss epilogue (losses.py:9-14, metrics.py:4-13), all null = off
  const float* target;          // (n_rays, 3)
  float* loss_part;             // [gridDim.x][2] per-CTA sums of squared errors (coarse, fine)
  float* loss_out;              // [4]: mse_coarse, mse_fine, mse_coarse + mse_fine, psnr of the finest pass
  unsigned* loss_counter;       // zero-initialised ticket counter (reset by the kernel)
  unsigned flags;               // experiment switches (NERFB200_FLAGS), 0 in production
  long long* timeline;          // experiment: device timeline buffer (flags & 2), else null
};

// Shared-memory scratch of the stand-alone kernels (aux_kernels.cuh: NeRF.forward, probes); the
// render kernel has its own RenderScratch below.
struct alignas(16) Scratch {
  Barriers bars;
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

// ---------------------------------------------------------------------------------------------
// The render kernel.  Warp roles (kRenderThreads = 640, 96 registers per thread):
//   0..15  MLP epilogue (csrc/mlp_engine.cuh) + sigma / rgb heads of every tile
//   16     weight producer     17  tcgen05 issuer
//   18, 19 helper warps: everything that is not the MLP - ray set-up, stratified depths, positional
//          encoding of the NEXT tile into the other ENC buffer, alpha compositing, inverse-CDF
//          resampling, merge, result stores - concurrently with the MLP of the current tile.
// Tiles are issued in the order C(0), C(1), F(0), C(2), F(1), ...: the coarse tile of group g+1 and
// the fine tiles of group g-1 run between the coarse tile of group g and its fine tiles, so the
// dependent chain "coarse sigma -> weights -> cdf -> new depths -> encoding" is off the tensor
// core's critical path.  Hand-over: enc_full[b] (helpers -> issuer + epilogue: ENC buffer b and
// the group's direction bias are written), out_full[b] (epilogue -> helpers: sigma / rgb of the
// 128 rows of the tile are in out_*[b]).  out_full of tile q-2 also tells the helpers that ENC
// buffer (q & 1) is no longer read.
constexpr int kHelperWarp0 = kMmaWarp + 1;       // 18
constexpr int kHelperWarps = 2;                  // 20 warps x 96 registers fill the register file
constexpr int kHelperThreads = kHelperWarps * 32;
constexpr int kRenderThreads = (kHelperWarp0 + kHelperWarps) * 32;   // 640
static_assert(kHelperWarps >= 2, "compositing uses one helper warp per ray of the group");
// Hand-over arrivals: one elected lane per warp after __syncwarp() (production), or every writing
// thread itself (-DNERFB200_SANITIZE: compute-sanitizer's racecheck follows per-thread
// arrive -> wait chains but not "other lanes' stores -> __syncwarp -> elected arrive").
#ifdef NERFB200_SANITIZE
constexpr bool kHandoffPerThread = true;
#else
constexpr bool kHandoffPerThread = false;
#endif
constexpr int kGroupSlots = 3;                   // groups in flight: g-1 (fine), g (resampling), g+1 (coarse)

struct alignas(16) GroupState {
  float ray[2][8];
  float dnorm[2];
  float pad0[2];
  float direnc[2][28];
  alignas(16) float dirbias[2][2][kDirW];   // [pass][ray][n]: b_dir + W_dir[:,256:283] . dir_enc
  float zc[2][kMaxSc];                      // coarse depths (kept for the merge)
  float z[kMaxRows];                        // fine depths [ray][S_f] (merged, sorted)
  float sigma[kMaxRows];                    // sigma of the current pass, overwritten in place by the weights
  float rgb[3][kMaxRows];
  float cdf[2][kMaxSc];
  float znew[2][kMaxImp];                   // u (sorted) then the new depths
};

struct alignas(16) RenderScratch {
  Barriers bars;
  uint64_t enc_full[2];
  uint64_t out_full[2];
  float sig_part[2][kColSplit][128];        // [tile buffer][column group][row] partial sigma-head sums
  float rgb_part[2][kColSplit][3][128];     //   (double-buffered: the epi_bar of tile q+1 separates the reads of tile q
                                            //    from the writes of tile q+2)
  float out_sigma[2][128];                  // [tile buffer][row]
  float out_rgb[2][3][128];
  GroupState gs[kGroupSlots];
};
static_assert(sizeof(RenderScratch) <= kScratchBytes, "render scratch does not fit");
static_assert(sizeof(GroupState) % 16 == 0, "GroupState keeps float4 alignment");

__device__ __forceinline__ void helper_bar() {
  asm volatile("bar.sync 2, %0;" ::"n"(kHelperThreads) : "memory");
}

struct Tile { int q, g, pass, tile; bool last; };

// Tile order of one CTA (see above); every role walks the same sequence.
struct TileSeq {
  int G, my_n, Sc, Sf;
  bool fine;
  int g = 0, phase = 0, t = 0, q = 0;
  __device__ __forceinline__ TileSeq(int G_, int my_n_, int Sc_, int Sf_, bool fine_)
      : G(G_), my_n(my_n_), Sc(Sc_), Sf(Sf_), fine(fine_) {}
  __device__ __forceinline__ int nr(int gg) const { return min(2, my_n - 2 * gg); }
  __device__ __forceinline__ static int tiles_of(int n_r, int S) { return (n_r * S + 127) >> 7; }
  __device__ __forceinline__ bool next(Tile& o) {
    while (g <= G) {
      if (phase == 0) {
        const int nt = (g < G) ? tiles_of(nr(g), Sc) : 0;
        if (t < nt) { o = Tile{q, g, 0, t, t + 1 == nt}; ++t; ++q; return true; }
        phase = 1; t = 0;
      } else {
        const int nt = (fine && g >= 1) ? tiles_of(nr(g - 1), Sf) : 0;
        if (t < nt) { o = Tile{q, g - 1, 1, t, t + 1 == nt}; ++t; ++q; return true; }
        phase = 0; t = 0; ++g;
      }
    }
    return false;
  }
};

template <bool kSave>
__global__ void __launch_bounds__(kRenderThreads, 1) render_rays_kernel(const RenderParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  RenderScratch* sc = reinterpret_cast<RenderScratch*>(smem + kSmemScratch);
  Barriers* bars = &sc->bars;
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&sc->enc_full[b]), kHandoffPerThread ? kHelperThreads : kHelperWarps);
      mbar_init(smem_u32(&sc->out_full[b]), kHandoffPerThread ? kEpiThreads : kEpiWarps);
    }
  }
  if (!engine_setup(smem, bars)) {
    if (threadIdx.x == 0) report_fault(p.status, 101);
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
  TileSeq seq(n_groups, my_n, Sc, Sf, fine);
  Tile tl;

  if (warp >= kEpiWarps && warp < kHelperWarp0) {
    if (warp == kProducerWarp && lane == 0) {
      RingState rs;
      if (kSave) rs.n = 2;
      while (seq.next(tl))
        produce_tile(rs, smem, bars, tl.pass ? p.net_fine : p.net_coarse, tl.pass == 0 && coarse_sigma_only, false);
    } else if (warp == kMmaWarp && lane == 0) {
      {
      RingState rs;
      if (kSave) rs.n = 2;
      MmaPhases ph;
      Timeline tlm{(blockIdx.x == 0) ? p.timeline : nullptr, {0, 0, 0}};
      while (seq.next(tl)) {
        const int b = tl.q & 1;
        mma_tile(rs, ph, smem, bars, tl.pass == 0 && coarse_sigma_only, false, &tlm, b ? kSmemEnc1 : kSmemEnc,
                 smem_u32(&sc->enc_full[b]), static_cast<uint32_t>(tl.q >> 1) & 1u);
      }
      }
    }
  } else if (warp < kEpiWarps) {
    // ================================ MLP epilogue warps ================================
    EpiCtx c;
    c.smem = smem;
    c.bars = bars;
    c.lane = lane;
    c.row = (warp & 3) * 32 + lane;
    c.part = warp >> 2;
    c.tmem_row = bars->tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    c.d_phase = 0;
    c.save_act = nullptr; c.save_mask = nullptr; c.save_d = nullptr; c.save_n = 0; c.save_row = -1;
    c.early = true;     // the accumulator is handed back as soon as a tile's last layer is read
    // training mode: the weight ring runs with 2 stages, the third stage's 32 KiB stages the activation stores
    c.stage.base = smem + kSmemRing + 2 * kSliceBytes256 + (warp & 3) * 4096;
    c.stage.buf = 0; c.stage.rg = warp & 3; c.stage.lane = lane; c.stage.part = warp >> 2;
    Timeline tle{(blockIdx.x == 0 && threadIdx.x == 0) ? p.timeline : nullptr, {0, 0, 0}};
    c.tl = &tle;
    while (seq.next(tl)) {
      const int b = tl.q & 1;
      const int pass = tl.pass;
      const int S = pass ? Sf : Sc;
      const bool valid1 = seq.nr(tl.g) == 2;
      const int ray0 = my_lo + 2 * tl.g;
      const bool sigma_only = (pass == 0) && coarse_sigma_only;
      const uint8_t* blob = pass ? p.net_fine : p.net_coarse;
      c.f32 = reinterpret_cast<const float*>(blob + kHalfRegionBytes);
      c.cst = consts_ptr(smem, pass);
      c.save_act = kSave ? p.tr[pass].act : nullptr;
      c.save_mask = kSave ? p.tr[pass].mask : nullptr;
      c.save_d = kSave ? p.tr[pass].d : nullptr;
      c.save_n = kSave ? p.tr[pass].n_pad : 0;

      const GroupState& gs = sc->gs[tl.g % kGroupSlots];
      const int gr = tl.tile * 128 + c.row;
      const int r = (gr >= S) ? 1 : 0;       // rows past the group's 2 S samples (S < 64 k) are padding
      const long long grow = (gr < 2 * S && (r == 0 || valid1)) ? static_cast<long long>(ray0 + r) * S + (gr - r * S) : -1;
      c.save_row = grow;
      c.save_g0 = __shfl_sync(0xffffffffu, grow, 0);      // rows come in whole rays of S = 32 k samples: a 32-row group is all valid or all padding
      // ENC buffer b and the group's direction bias are ready
      mbar_wait(smem_u32(&sc->enc_full[b]), static_cast<uint32_t>(tl.q >> 1) & 1u, 7);
      if (kSave && c.part == 0 && lane == 0 && c.save_g0 >= 0) {
        // the encoded-input tile is the B operand of the wgrad of layers 1 and 5: the shared-memory image IS the
        // tiled layout (row & 7 == global row & 7), so this row group's 32 rows go out as one 4 KiB bulk store
        uint8_t* dst = p.tr[pass].enc + tiled_block_off(static_cast<unsigned long long>(c.save_g0 >> 6), 0, 1) +
                       (c.save_g0 & 63) * 128;
        bulk_s2g(dst, smem_u32(smem + (b ? kSmemEnc1 : kSmemEnc) + (c.row >> 5) * 4096), 4096);
        bulk_commit();       // (the staging rounds' wait_group.read also covers this read; ENC is rewritten two tiles later)
      }
      float sig_part, rgb_part[3];
      epi_run_tile<kSave>(c, sigma_only, gs.dirbias[pass][r], nullptr, sig_part, rgb_part);
      sc->sig_part[b][c.part][c.row] = sig_part;
      if (!sigma_only) {
        sc->rgb_part[b][c.part][0][c.row] = rgb_part[0];
        sc->rgb_part[b][c.part][1][c.row] = rgb_part[1];
        sc->rgb_part[b][c.part][2][c.row] = rgb_part[2];
      }
      epi_bar();
      // combine the column groups' partial head sums: group 0 -> sigma, groups 1..3 -> r, g, b
      if (c.part == 0) {
        float sg = c.cst[kF32BSigma];
#pragma unroll
        for (int q = 0; q < kColSplit; ++q) sg += sc->sig_part[b][q][c.row];
        sc->out_sigma[b][c.row] = sg;
        if (kSave && grow >= 0) {
          p.tr[pass].sigma[grow] = sg;
        }
      } else if (!sigma_only) {
        const int ch = c.part - 1;
        float pre = c.cst[kF32BRgb + ch];
#pragma unroll
        for (int q = 0; q < kColSplit; ++q) pre += sc->rgb_part[b][q][ch][c.row];
        const float col = sigmoid_ref(pre);
        sc->out_rgb[b][ch][c.row] = col;
        if (kSave && grow >= 0) {
          p.tr[pass].rgb[grow * 3 + ch] = col;
        }
      }
      if (kHandoffPerThread) {
        mbar_arrive(smem_u32(&sc->out_full[b]));
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sc->out_full[b]));
      }
    }
    if (kSave) bulk_wait_all();      // this thread's bulk stores (if any) have left shared memory and landed
  } else if (warp >= kHelperWarp0) {
    // ================================== helper warps ===================================
    const int ht = threadIdx.x - kHelperWarp0 * 32;   // 0..127
    const int hw = ht >> 5;
    TileSeq cons_seq(n_groups, my_n, Sc, Sf, fine);
    Tile tc;
    float loss_c = 0.f, loss_f = 0.f;   // lane 0 of each helper warp: squared error of its rays (coarse, fine)
    int consumed = 0;           // tiles consumed so far (== q of the next tile to consume)
    int cpos0 = 0, cpos1 = 0, cpos2 = 0;   // position of the (last) coarse tile of the group in slot 0 / 1 / 2

    // ---- consume one finished tile: sigma / rgb rows -> group arrays; after the last tile of a
    // pass: compositing (models/rendering.py:143-170), result stores, and after the coarse pass
    // the hierarchical resampling (:28-54) and the merge (:229)
    auto consume = [&](const Tile& tt) {
      const int b = tt.q & 1;
      const int pass = tt.pass;
      const int S = pass ? Sf : Sc;
      const bool valid1 = seq.nr(tt.g) == 2;
      const int ray0 = my_lo + 2 * tt.g;
      const int rid[2] = {ray0, valid1 ? ray0 + 1 : ray0};
      const bool sigma_only = (pass == 0) && coarse_sigma_only;
      GroupState& gs = sc->gs[tt.g % kGroupSlots];
      mbar_wait(smem_u32(&sc->out_full[b]), static_cast<uint32_t>(tt.q >> 1) & 1u, 8);
      for (int row = ht; row < 128; row += kHelperThreads) {
        const int gr = tt.tile * 128 + row;
        if (gr < 2 * S) {
          gs.sigma[gr] = sc->out_sigma[b][row];
          if (!sigma_only) {
            gs.rgb[0][gr] = sc->out_rgb[b][0][row];
            gs.rgb[1][gr] = sc->out_rgb[b][1][row];
            gs.rgb[2][gr] = sc->out_rgb[b][2][row];
          }
        }
      }
      helper_bar();
      if (!tt.last) return;
      // ---- compositing: helper warp r renders ray r
      if (hw < 2) {
        const int r = hw;
        const float* zr = pass ? (gs.z + r * S) : gs.zc[r];
        const float* nz = nullptr;
        if (p.noise_std > 0.f)
          nz = (pass ? p.noise_fine : p.noise_coarse) + static_cast<long long>(rid[r]) * S;
        const RayOut o = composite_ray(lane, S, zr, gs.sigma + r * S, gs.rgb[0] + r * S, gs.rgb[1] + r * S,
                                       gs.rgb[2] + r * S, nz, p.noise_std, gs.dnorm[r], !sigma_only,
                                       gs.sigma + r * S);
        __syncwarp();
        const bool wr = (r == 0) || valid1;
        if (wr) {
          const long long ri = rid[r];
          float* wout = pass ? p.weights_fine : p.weights_coarse;
          if (wout != nullptr)
            for (int i = lane; i < S; i += 32) wout[ri * S + i] = gs.sigma[r * S + i];
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
            if (p.target != nullptr && !sigma_only) {      // losses.py:9-14: squared error of this ray
              const float e0 = (o.r + add) - __ldg(p.target + 3 * ri), e1 = (o.g + add) - __ldg(p.target + 3 * ri + 1),
                          e2 = (o.b + add) - __ldg(p.target + 3 * ri + 2);
              const float se = e0 * e0 + e1 * e1 + e2 * e2;
              if (pass == 0) loss_c += se; else loss_f += se;
            }
          }
        }
        // ---- hierarchical resampling, part 1 (models/rendering.py:28-33): pdf -> cdf
        if (pass == 0 && fine) pdf_to_cdf_ray(lane, Sc, gs.sigma + r * Sc, gs.cdf[r]);
      }
      helper_bar();
      if (pass != 0 || !fine) return;
      // ---- part 2 (:36-54): one u per thread -> inverse-CDF depth; u sorted first when random
      for (int t = ht; t < 2 * K; t += kHelperThreads) {
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
        gs.znew[r][slot] = inverse_cdf(Sc, gs.zc[r], gs.cdf[r], uj);
      }
      helper_bar();
      // ---- merge: z_fine = sort(cat(z_coarse, z_new)) (:229) as a rank computation:
      // position = number of elements that are smaller, ties broken by index in the
      // concatenation (any tie order gives the same sorted VALUES, which is all torch.sort's
      // output carries).  Both lists are sorted except for possible 1-ulp inversions from
      // rounding, so every element first checks its predecessor; if no inversion exists
      // anywhere (the common case) ranks come from two binary searches, otherwise from
      // exhaustive counting.  Both paths give identical results on sorted input.
      {
        bool inv = false;
        for (int e = ht; e < 2 * Sf; e += kHelperThreads) {
          const int r = e / Sf, i = e - r * Sf;
          if (i != 0 && i != Sc)
            inv |= (i < Sc) ? (gs.zc[r][i] < gs.zc[r][i - 1]) : (gs.znew[r][i - Sc] < gs.znew[r][i - Sc - 1]);
        }
        int any_inv;
        asm volatile(
            "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
            "barrier.red.or.pred p, 2, %2, q;\n\tselp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(any_inv) : "r"(static_cast<int>(inv)), "n"(kHelperThreads) : "memory");
        for (int e = ht; e < 2 * Sf; e += kHelperThreads) {
          const int r = e / Sf, i = e - r * Sf;
          const float* zc = gs.zc[r];
          const float* zn = gs.znew[r];
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
          gs.z[r * Sf + rank] = v;
        }
      }
      helper_bar();
      if (p.z_fine != nullptr) {
        for (int e = ht; e < 2 * Sf; e += kHelperThreads) {
          const int r = e / Sf, i = e - r * Sf;
          if (r == 0 || valid1) p.z_fine[static_cast<long long>(rid[r]) * Sf + i] = gs.z[e];
        }
      }
    };

    // ---- rays, direction embedding, stratified depths, direction bias of one group
    auto setup_group = [&](int g) {
      GroupState& gs = sc->gs[g % kGroupSlots];
      const bool valid1 = seq.nr(g) == 2;
      const int ray0 = my_lo + 2 * g;
      const int rid[2] = {ray0, valid1 ? ray0 + 1 : ray0};
      // models/rendering.py:179-186
      if (ht < 16) gs.ray[ht >> 3][ht & 7] = __ldg(p.rays + static_cast<long long>(rid[ht >> 3]) * p.ray_stride + (ht & 7));
      helper_bar();
      if (ht < 2) {
        const float* d = &gs.ray[ht][3];
        gs.dnorm[ht] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
      } else if (ht >= 32 && ht < 32 + 30) {
        static_assert(kHelperThreads >= 62, "direction embedding uses helper threads 32..61");
        // Embedding(3,4)(rays_d) (models/rendering.py:186): one (ray, coord, freq) per thread, accurate sincosf
        const int q = ht - 32, r = q / 15, cc = (q % 15) / 5, k = q % 5;     // k == 4: the raw value
        const float dv = gs.ray[r][3 + cc];
        if (k == 4) {
          gs.direnc[r][cc] = dv;
        } else {
          float sn, cs;
          sincosf(__fmul_rn(static_cast<float>(1 << k), dv), &sn, &cs);
          gs.direnc[r][3 + 6 * k + cc] = sn;
          gs.direnc[r][3 + 6 * k + 3 + cc] = cs;
        }
      }
      // ---- coarse depths (models/rendering.py:189-204)
      for (int e = ht; e < 2 * Sc; e += kHelperThreads) {
        const int r = e / Sc, i = e - r * Sc;
        const float nr = gs.ray[r][6], fr = gs.ray[r][7];
        float z = z_base(nr, fr, i, Sc, p.use_disp != 0);
        if (p.perturb > 0.f) {
          const float zl = (i > 0) ? z_base(nr, fr, i - 1, Sc, p.use_disp != 0) : z;
          const float zu = (i < Sc - 1) ? z_base(nr, fr, i + 1, Sc, p.use_disp != 0) : z;
          const float lower = (i > 0) ? __fmul_rn(0.5f, __fadd_rn(zl, z)) : z;
          const float upper = (i < Sc - 1) ? __fmul_rn(0.5f, __fadd_rn(z, zu)) : z;
          const float pr = __fmul_rn(p.perturb, __ldg(p.perturb_rand + static_cast<long long>(rid[r]) * Sc + i));
          z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), pr));
        }
        gs.zc[r][i] = z;
        if (p.z_coarse != nullptr && (r == 0 || valid1)) p.z_coarse[static_cast<long long>(rid[r]) * Sc + i] = z;
      }
      helper_bar();
      // ---- per-ray direction bias of both networks: b_dir + W_dir[:, 256:283] . dir_embedded (fp32)
      for (int pass = 0; pass < (fine ? 2 : 1); ++pass) {
        if (pass == 0 && coarse_sigma_only) continue;
        const uint8_t* blob = pass ? p.net_fine : p.net_coarse;
        const float* f32 = reinterpret_cast<const float*>(blob + kHalfRegionBytes);
        const float* cst = consts_ptr(smem, pass);
        for (int e = ht; e < 2 * kDirW; e += kHelperThreads) {
          const int r = e >> 7, n = e & 127;
          const float* wd = f32 + kF32WDirPart + n;           // [j][n]: coalesced over n
          float acc = cst[kF32Bias + 8 * 256 + n];
#pragma unroll
          for (int j = 0; j < 27; ++j) acc = fmaf(__ldg(wd + j * 128), gs.direnc[r][j], acc);
          gs.dirbias[pass][r][n] = acc;
        }
      }
    };

    while (seq.next(tl)) {
      int need = tl.q - 2;
      if (tl.pass == 1 && tl.tile == 0) {
        const int slot = tl.g % kGroupSlots;
        need = max(need, slot == 0 ? cpos0 : (slot == 1 ? cpos1 : cpos2));
      }
      while (consumed <= need) {
        cons_seq.next(tc);
        consume(tc);
        ++consumed;
      }
      // ---- prepare tile q: (new group: set-up) + positional encoding into ENC buffer q & 1
      const int b = tl.q & 1;
      GroupState& gs = sc->gs[tl.g % kGroupSlots];
      if (tl.pass == 0) {
        if (tl.tile == 0) setup_group(tl.g);
        if (tl.last) {
          const int slot = tl.g % kGroupSlots;
          if (slot == 0) cpos0 = tl.q; else if (slot == 1) cpos1 = tl.q; else cpos2 = tl.q;
        }
      }
      {
        const int S = tl.pass ? Sf : Sc;
        uint8_t* enc = smem + (b ? kSmemEnc1 : kSmemEnc);
#pragma unroll 1
        for (int row = ht; row < 128; row += kHelperThreads) {
          const int gr = tl.tile * 128 + row;
          const int r = (gr >= S) ? 1 : 0;   // padding rows encode whatever is there; their results are dropped
          const float zval = tl.pass ? gs.z[gr] : gs.zc[r][min(gr - r * S, kMaxSc - 1)];
#pragma unroll 1
          for (int part = 0; part < kColSplit; ++part) encode_row(enc, row, part, &gs.ray[r][0], &gs.ray[r][3], zval);
        }
      }
      fence_proxy_async();
      if (kHandoffPerThread) {
        mbar_arrive(smem_u32(&sc->enc_full[b]));
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sc->enc_full[b]));
      }
    }
    while (cons_seq.next(tc)) consume(tc);
    // ---- fused loss epilogue: per-CTA partial sums, the last CTA to finish reduces them in a fixed
    // order (deterministic) and writes MSELoss / psnr (losses.py:9-14, metrics.py:4-13)
    if (p.target != nullptr) {
      float* lsum = &sc->sig_part[0][0][0];       // the epilogue warps are done with it
      helper_bar();
      if (lane == 0 && hw < 2) { lsum[2 * hw] = loss_c; lsum[2 * hw + 1] = loss_f; }
      helper_bar();
      if (ht == 0) {
        p.loss_part[2 * blockIdx.x] = lsum[0] + lsum[2];
        p.loss_part[2 * blockIdx.x + 1] = lsum[1] + lsum[3];
        __threadfence();
        const unsigned ticket = atomicAdd(p.loss_counter, 1u);
        if (ticket == gridDim.x - 1) {
          __threadfence();
          double sc_ = 0.0, sf_ = 0.0;
          for (unsigned i = 0; i < gridDim.x; ++i) {
            sc_ += static_cast<double>(*reinterpret_cast<volatile float*>(p.loss_part + 2 * i));
            sf_ += static_cast<double>(*reinterpret_cast<volatile float*>(p.loss_part + 2 * i + 1));
          }
          const double ne = 3.0 * static_cast<double>(p.n_rays);
          const float mc = static_cast<float>(sc_ / ne), mf = static_cast<float>(sf_ / ne);
          p.loss_out[0] = mc;
          p.loss_out[1] = fine ? mf : 0.f;
          p.loss_out[2] = fine ? mc + mf : mc;
          p.loss_out[3] = -10.f * log10f(fine ? mf : mc);
          *p.loss_counter = 0u;
        }
      }
    }
  }
  engine_teardown(bars);
}

}  // namespace nerfb200
