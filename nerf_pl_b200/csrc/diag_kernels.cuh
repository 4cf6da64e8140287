// Diagnostics and microbenchmarks (bring-up probes, tcgen05 issue-rate / contention benchmarks).
// NOT part of the product library: compiled only with -DNERFB200_DIAG (tools/build_variants.py
// diag=-DNERFB200_DIAG); the entry points are declared in include/nerf_pl_b200_diag.h.
#pragma once
#include <cuda_bf16.h>

#include "aux_kernels.cuh"

namespace nerfb200 {

// ------------------------------------------------ diagnostics: one K=64 slice through the engine
// d[128 x N] = fp16(a[128 x 64]) . slice^T, N = 256 (slices 0..33) or 128 (34..38), read back from
// TMEM unmodified.  mode 0: A staged in the ENC shared-memory tile (SS MMA, as layer 1 / the skip
// part of layer 5); mode 1: A written to TMEM with tcgen05.st (TS MMA, as every hidden layer).
// Isolates descriptor / swizzle / TMEM layouts from the layer protocol.
__global__ void __launch_bounds__(kThreads, 1) gemm_probe_kernel(const float* __restrict__ a,
                                                                 const uint8_t* __restrict__ blob,
                                                                 int slice, int mode, float* __restrict__ d,
                                                                 int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Scratch* sc = reinterpret_cast<Scratch*>(smem + kSmemScratch);
  Barriers* bars = &sc->bars;
  if (!engine_setup(smem, bars)) {
    if (threadIdx.x == 0) report_fault(status, 101);
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool wide = slice < kNumSlices256;
  const int N = wide ? 256 : 128;
  if (warp == kProducerWarp) {
    if (lane == 0) {
      const uint32_t full = smem_u32(&bars->full[0]);
      const uint32_t bytes = wide ? kSliceBytes256 : kSliceBytes128;
      const uint8_t* src = wide ? blob + static_cast<size_t>(slice) * kSliceBytes256
                                : blob + kOffDir + static_cast<size_t>(slice - kNumSlices256) * kSliceBytes128;
      mbar_wait(smem_u32(&bars->empty[0]), 1, 11);
      mbar_arrive_expect_tx(full, bytes);
      for (uint32_t c = 0; c < bytes; c += 8192) bulk_g2s(smem_u32(smem + kSmemRing) + c, src + c, 8192, full);
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0) {
      mbar_wait(smem_u32(&bars->d_free), 0, 12);
      tc_fence_after();
      mbar_wait(smem_u32(&bars->full[0]), 0, 13);
      tc_fence_after();
      const uint64_t bdesc = make_desc_sw128(smem_u32(smem + kSmemRing));
      const uint32_t idesc = make_idesc_f16(N);
      if (mode == 0) {
        const uint64_t adesc = make_desc_sw128(smem_u32(smem + kSmemEnc));
        for (int j = 0; j < 4; ++j) umma_f16(bars->tmem_base + kTmemD, adesc + 2 * j, bdesc + 2 * j, idesc, j != 0);
      } else {
        for (int j = 0; j < 4; ++j)
          umma_f16_ts(bars->tmem_base + kTmemD, bars->tmem_base + kTmemA + 8 * j, bdesc + 2 * j, idesc, j != 0);
      }
      umma_commit(smem_u32(&bars->d_ready));
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
    if (mode == 0) {
      uint8_t* enc = smem + kSmemEnc;
      for (int k = c.part * (64 / kColSplit); k < (c.part + 1) * (64 / kColSplit); ++k)
        *reinterpret_cast<__half*>(enc + sw128_off(c.row, k)) = __float2half_rn(a[c.row * 64 + k]);
    } else if (c.part < 2) {
      uint32_t h[16];   // parts 0,1: 32 K-values each -> 16 packed columns at A + 16*part
      for (int i = 0; i < 16; ++i)
        h[i] = cvt_f16x2(a[c.row * 64 + c.part * 32 + 2 * i], a[c.row * 64 + c.part * 32 + 2 * i + 1]);
      tmem_st16(c.tmem_row + kTmemA + 16 * c.part, h);
    }
    tmem_st_wait();
    epi_signal_tile_start(c);
    epi_wait_d(c);
    const int ncol = N / kColSplit;
    for (int cc = 0; cc < ncol; cc += 32) {
      uint32_t r[32];
      tmem_ld32(c.tmem_row + kTmemD + c.part * ncol + cc, r);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) d[c.row * N + c.part * ncol + cc + i] = __uint_as_float(r[i]);
    }
  }
  engine_teardown(bars);
}

// ------------------------------------------------ diagnostics: MN-major operands (wgrad layout)
// d[128 x 256] = a^T b, a (64 samples, 128 features), b (64 samples, 256 features): both staged as
// [64-feature block][64 rows][128 B] SWIZZLE_128B images (block stride 8 KiB), read MN-major.
__global__ void __launch_bounds__(kThreads, 1) gemm_mn_probe_kernel(const float* __restrict__ a,
                                                                    const float* __restrict__ b, uint32_t lbo,
                                                                    uint32_t sbo, uint32_t fmt, float* __restrict__ d, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Scratch* sc = reinterpret_cast<Scratch*>(smem + kSmemScratch);
  Barriers* bars = &sc->bars;
  if (!engine_setup(smem, bars)) {
    if (threadIdx.x == 0) report_fault(status, 101);
    return;
  }
  const int warp = threadIdx.x >> 5;
  uint8_t* sa = smem + kSmemRing;
  uint8_t* sb = smem + kSmemRing + 32768;
  if (warp < kEpiWarps) {
    for (int i = threadIdx.x; i < 64 * 128; i += kEpiThreads) {
      const int s = i / 128, m = i % 128;
      uint8_t* dst = sa + (m >> 6) * 8192 + sw128_off(s, m & 63);
      if (fmt & 1u) *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(a[i]);     // A as bf16
      else *reinterpret_cast<__half*>(dst) = __float2half_rn(a[i]);
    }
    for (int i = threadIdx.x; i < 64 * 256; i += kEpiThreads) {
      const int s = i / 256, n = i % 256;
      uint8_t* dst = sb + (n >> 6) * 8192 + sw128_off(s, n & 63);
      if (fmt & 2u) *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(b[i]);     // B as bf16
      else *reinterpret_cast<__half*>(dst) = __float2half_rn(b[i]);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(smem_u32(&bars->d_free));
    mbar_wait(smem_u32(&bars->d_ready), 0, 41);
    tc_fence_after();
    const uint32_t row = bars->tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const int part = warp >> 2, r_ = (warp & 3) * 32 + (threadIdx.x & 31);
    for (int cc = 0; cc < 64; cc += 32) {
      uint32_t r[32];
      tmem_ld32(row + part * 64 + cc, r);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) d[r_ * 256 + part * 64 + cc + i] = __uint_as_float(r[i]);
    }
  } else if (warp == kMmaWarp && (threadIdx.x & 31) == 0) {
    mbar_wait(smem_u32(&bars->d_free), 0, 42);
    tc_fence_after();
    // operand formats: idesc bits [7,10) = A, [10,13) = B; 0 = f16, 1 = bf16 (mixed = the wgrad's bf16 x f16)
    const uint32_t idesc = make_idesc_f16_mn(256) | ((fmt & 1u) << 7) | (((fmt >> 1) & 1u) << 10);
    for (int j = 0; j < 4; ++j) {      // 16 samples = two 8-row groups = 2048 B per K step
      const uint64_t ad = make_desc_mn_sw128(smem_u32(sa) + j * 2048, lbo, sbo);
      const uint64_t bd = make_desc_mn_sw128(smem_u32(sb) + j * 2048, lbo, sbo);
      umma_f16(bars->tmem_base, ad, bd, idesc, j != 0);
    }
    umma_commit(smem_u32(&bars->d_ready));
  }
  engine_teardown(bars);
}

// ------------------------------------------------ diagnostics: raw tcgen05.mma issue rate
// out[block*8 + v] = SM cycles for `reps` x 16 back-to-back MMAs (K=16 each) of variant v:
//   0: SS N=256   1: SS N=128   2: TS N=128   3: TS N=256   4: TS N=128 alternating D0/D1
// Operands are whatever is in smem / TMEM (timing only).
__global__ void __launch_bounds__(kThreads, 1) mma_bench_kernel(long long* __restrict__ out, int reps,
                                                                int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Scratch* sc = reinterpret_cast<Scratch*>(smem + kSmemScratch);
  Barriers* bars = &sc->bars;
  if (!engine_setup(smem, bars)) {
    if (threadIdx.x == 0) report_fault(status, 101);
    return;
  }
  // zero the operand memory so the timing is not skewed by NaN/denormal paths
  for (uint32_t i = threadIdx.x; i < kSmemScratch / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == kMmaWarp * 32) {
    const uint32_t tmem = bars->tmem_base;
    const uint64_t adesc = make_desc_sw128(smem_u32(smem + kSmemEnc));
    const uint64_t bdesc = make_desc_sw128(smem_u32(smem + kSmemRing));
    uint32_t phase = 0;
    for (int v = 0; v < 5; ++v) {
      const uint32_t idesc = (v == 0 || v == 3) ? make_idesc_f16(256) : make_idesc_f16(128);
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t j = i & 3;
          const uint64_t b = bdesc + 2 * j + ((i >> 2) * 2048);      // next 32 KiB stage per K block
          if (v <= 1) umma_f16(tmem + kTmemD, adesc + 2 * j, b, idesc, 1u);
          else if (v == 4) umma_f16_ts(tmem + kTmemD + ((i & 4) ? 128 : 0), tmem + kTmemA + (i >> 2) * 32 + j * 8, b, idesc, 1u);
          else umma_f16_ts(tmem + kTmemD, tmem + kTmemA + (i >> 2) * 32 + j * 8, b, idesc, 1u);
        }
      }
      umma_commit(smem_u32(&bars->d_ready));
      mbar_wait(smem_u32(&bars->d_ready), phase, 21);
      phase ^= 1;
      out[blockIdx.x * 8 + v] = clock64() - t0;
    }
  }
  engine_teardown(bars);
}

// ------------------------------------------------ tcgen05 / TMEM contention microbenchmark
// The MMA thread issues `reps` x 16 MMAs (K=16, M=128, N=256) into accumulator region 0 while
// the 16 epilogue warps run an epilogue-shaped background load on region 1 (columns 256..511):
//   bg 0: idle   1: tcgen05.ld only (4 x16 per iteration = this warp's share of a 128x256 fp32
//   accumulator)   2: ld + bias/ReLU/convert + tcgen05.st (in place, 8 columns per K block)
//   3: tcgen05.st only   4: as 2 but one K block at a time (ld 16, convert, st 8)
// variant 0: SS (A from smem)   1: TS, A read from region 1 in the in-place layout
//   (K block kb, step j -> columns 256 + kb*64 + j*16)   2: TS, A from columns 256 + kb*32 + j*8
// out[block*4 + 0] = cycles of the MMA loop, [1] = background iterations of warp 0 during it,
// [2] = checksum (keeps the loads alive).
__global__ void __launch_bounds__(kThreads, 1) mma_contention_kernel(long long* __restrict__ out, int reps,
                                                                     int bg, int variant, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Scratch* sc = reinterpret_cast<Scratch*>(smem + kSmemScratch);
  Barriers* bars = &sc->bars;
  if (!engine_setup(smem, bars)) {
    if (threadIdx.x == 0) report_fault(status, 101);
    return;
  }
  for (uint32_t i = threadIdx.x; i < kSmemScratch / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  volatile int* stop = reinterpret_cast<volatile int*>(&sc->z[0]);
  volatile int* go = reinterpret_cast<volatile int*>(&sc->z[1]);
  if (threadIdx.x == 0) { *stop = 0; *go = 0; }
  fence_proxy_async();
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem = bars->tmem_base;
  if (warp < kEpiWarps) {
    // zero both regions once (defined operands)
    const uint32_t row = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const int part = warp >> 2;
    uint32_t zero[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) zero[i] = 0;
    for (int cg = 0; cg < 8; ++cg) tmem_st16(row + part * 128 + cg * 16, zero);
    tmem_st_wait();
    tc_fence_before();
    epi_bar();
    if (threadIdx.x == 0) *go = 1;
    long long iters = 0;
    uint32_t chk = 0;
    uint32_t r[4][16];
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = 0;
    while (*stop == 0) {
      if (bg == 1 || bg == 2) {
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) tmem_ld16(row + 256 + kb * 64 + part * 16, r[kb]);
        tmem_ld_wait();
        if (bg == 1) {
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) chk ^= r[kb][0] ^ r[kb][7] ^ r[kb][15];
        }
      }
      if (bg == 2) {
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float a, b;
            add_f32x2(a, b, __uint_as_float(r[kb][2 * i]), __uint_as_float(r[kb][2 * i + 1]), 0.25f, 0.5f);
            h[i] = cvt_f16x2_relu(a, b);
          }
          tmem_st8(row + 256 + kb * 64 + part * 16, h);
          tmem_st_wait();
        }
      }
      if (bg == 3) {
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          tmem_st8(row + 256 + kb * 64 + part * 16, h);
          tmem_st_wait();
        }
      }
      if (bg == 4) {
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          tmem_ld16(row + 256 + kb * 64 + part * 16, r[0]);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float a, b;
            add_f32x2(a, b, __uint_as_float(r[0][2 * i]), __uint_as_float(r[0][2 * i + 1]), 0.25f, 0.5f);
            h[i] = cvt_f16x2_relu(a, b);
          }
          tmem_st8(row + 256 + kb * 64 + part * 16, h);
          tmem_st_wait();
        }
      }
      if (bg == 5) {   // what the real epilogue does while the MMA runs: poll an mbarrier that does not flip
        uint32_t done;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(&bars->a_ready)), "r"(0u) : "memory");
        chk ^= done;
      }
      if (bg == 6) {   // st8 + wait + fence + elected arrive, the real hand-over sequence
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          tmem_st8(row + 256 + kb * 64 + part * 16, h);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->a_kb[kb]));
        }
      }
      if (bg == 0) __nanosleep(200);
      ++iters;
    }
    tc_fence_before();
    if (threadIdx.x == 0) {
      out[blockIdx.x * 4 + 1] = iters;
      out[blockIdx.x * 4 + 2] = chk;
    }
  } else if (threadIdx.x == kMmaWarp * 32) {
    while (*go == 0) {}
    tc_fence_after();
    const uint64_t adesc = make_desc_sw128(smem_u32(smem + kSmemEnc));
    const uint64_t bdesc = make_desc_sw128(smem_u32(smem + kSmemRing));
    const uint32_t idesc = make_idesc_f16(256);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t j = i & 3, kb = i >> 2;
        const uint64_t b = bdesc + 2 * j + (kb * 2048);
        if (variant == 0) umma_f16(tmem + kTmemD, adesc + 2 * j, b, idesc, 1u);
        else if (variant == 1) umma_f16_ts(tmem + kTmemD, tmem + 256 + kb * 64 + j * 16, b, idesc, 1u);
        else umma_f16_ts(tmem + kTmemD, tmem + 256 + kb * 32 + j * 8, b, idesc, 1u);
      }
    }
    umma_commit(smem_u32(&bars->d_ready));
    mbar_wait(smem_u32(&bars->d_ready), 0, 21);
    out[blockIdx.x * 4 + 0] = clock64() - t0;
    *stop = 1;
  }
  engine_teardown(bars);
}

// ------------------------------------------------ tcgen05 issue-pattern microbenchmark
// One thread issues reps x 16 TS MMAs (M=128, N=256, K=16) with bookkeeping around them, to find
// how far the issuing thread may run ahead of the tensor pipe (queue depth) and what the engine's
// per-slice waits cost.  MODE (compile time, so the loop stays lean):
//   0 back-to-back                         1 commit after every 4
//   2 busy-wait `arg` cycles after every 4th MMA        3 busy-wait `arg` cycles after every MMA
//   4 two mbarrier waits (already complete) + fence BEFORE every 4
//   5 the same waits + fence between MMA 0 and MMA 1 of every 4 (software-pipelined)
//   6 as 5, between MMA 1 and MMA 2       7 as 5, between MMA 2 and 3
//   8 one mbarrier wait + fence before every 4
// out[block*4+0] = cycles.
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) mma_issue_kernel(long long* __restrict__ out, int reps, int arg,
                                                                int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Scratch* sc = reinterpret_cast<Scratch*>(smem + kSmemScratch);
  Barriers* bars = &sc->bars;
  if (!engine_setup(smem, bars)) {
    if (threadIdx.x == 0) report_fault(status, 101);
    return;
  }
  for (uint32_t i = threadIdx.x; i < kSmemScratch / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  const uint32_t tmem = bars->tmem_base;
  if (warp < kEpiWarps) {
    const uint32_t row = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t zero[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) zero[i] = 0;
    for (int cg = 0; cg < 8; ++cg) tmem_st16(row + (warp >> 2) * 128 + cg * 16, zero);
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  if (threadIdx.x == kMmaWarp * 32) {
    tc_fence_after();
    const uint64_t bdesc = make_desc_sw128(smem_u32(smem + kSmemRing));
    const uint32_t idesc = make_idesc_f16(256);
    const uint32_t bar_a = smem_u32(&bars->a_ready), bar_b = smem_u32(&bars->d_free);
    auto waits = [&]() {
      mbar_wait(bar_a, 1, 31);        // fresh barrier: parity 1 is "already complete"
      if (MODE != 8) mbar_wait(bar_b, 1, 32);
      tc_fence_after();
    };
    auto spin = [&](int n) {
      const long long t = clock64();
      while (clock64() - t < n) {}
    };
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        if (MODE == 4 || MODE == 8) waits();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          umma_f16_ts(tmem + kTmemD, tmem + kTmemA + kb * 32 + j * 8, bdesc + kb * 2048 + 2 * j, idesc, 1u);
          if (MODE == 3) spin(arg);
          if ((MODE == 5 && j == 0) || (MODE == 6 && j == 1) || (MODE == 7 && j == 2)) waits();
        }
        if (MODE == 2) spin(arg);
        if (MODE >= 1) umma_commit(smem_u32(&bars->empty[kb]));
      }
    }
    umma_commit(smem_u32(&bars->d_ready));
    mbar_wait(smem_u32(&bars->d_ready), 0, 21);
    out[blockIdx.x * 4 + 0] = clock64() - t0;
  }
  engine_teardown(bars);
}

}  // namespace nerfb200
