// The NeRF MLP (models/nerf.py:83-124) as a warp-specialised tcgen05 tile engine.
//
// One CTA owns one 128-row tile of samples at a time.  The hidden activations never touch
// shared memory: each layer's fp32 accumulator lives in TMEM, the epilogue warps read it
// (tcgen05.ld), apply bias / ReLU, convert to fp16 and write it back to TMEM (tcgen05.st) as
// the A operand of the next layer's tcgen05.mma (A-from-TMEM form).  Shared memory only holds
// the encoded-input tile and the weight ring, so its bandwidth is spent on weights alone.
//
//   TMEM (512 columns allocated, lane = tile row):
//     [  0,256) D   fp32 accumulator, 256 output features
//     [256,384) A   fp16 activations (2 per 32-bit column), the next layer's A operand;
//                   rewritten by the epilogue once the layer's MMAs have retired
//   warps 0..15 epilogue: warp w owns TMEM lanes 32(w&3).. and column group w>>2
//   warp 16     weight producer: streams the packed 32 KiB K-slices (layout.h) through a
//               3-stage ring with cp.async.bulk + mbarriers
//   warp 17     MMA issuer (one thread): tcgen05.mma M=128, N=256|128, K=16; A from TMEM for
//               hidden K blocks (128 cycles per K step when issued back to back), from the
//               ENC shared-memory tile for the encoded-input slices
//   smem also keeps the fp32 biases and head weights of both networks resident (24 KiB), so the
//   epilogue reads them with broadcast LDS instead of L1/L2 loads
//
// Hand-over between layers: a tile has ONE accumulator, so its MMA and
// epilogue cannot fully overlap, but the hand-over is pipelined at K-block granularity: every
// epilogue warp first drains its 4 x 16 accumulator columns into registers (tcgen05.ld), then
// converts the 16 columns belonging to K block 0, stores them (tcgen05.st) and arrives on
// a_kb[0]; then K block 1, ...  The MMA issuer starts the next layer's K block 0 as soon as
// a_kb[0] completes (which also proves that every warp has drained the accumulator, so the first
// MMA may overwrite it) while the warps are still converting blocks 1..3.
// Measured alternatives are recorded in DESIGN.md section 4 (A in shared memory; N=128 split
// accumulators; sequential hand-over; two tiles alternating on one accumulator).
#pragma once
#include <cuda_fp16.h>

#include "layout.h"
#include "ptx.cuh"

namespace nerfb200 {

constexpr int kEpiWarps = 16;                 // 4 per scheduler: latency hiding for the serial epilogue
constexpr int kEpiThreads = kEpiWarps * 32;   // 512
constexpr int kColSplit = kEpiWarps / 4;      // column groups per accumulator row (4 lane quadrants each)
constexpr int kColsPer = 256 / kColSplit;     // accumulator columns per epilogue thread (64)
constexpr int kProducerWarp = kEpiWarps;
constexpr int kMmaWarp = kEpiWarps + 1;
constexpr int kThreads = (kEpiWarps + 2) * 32;   // 576
// Ring depth: 3, 4 and 5 stages measure the same (the weights come from L2 and one layer of
// look-ahead is enough), so the ring takes 96 KiB and the rest of shared memory holds the second
// ENC tile and the per-group state of the render kernel's helper warps.
#ifndef NERFB200_STAGES
#define NERFB200_STAGES 3
#endif
constexpr int kStages = NERFB200_STAGES;
constexpr int kTmemCols = 512;
constexpr uint32_t kTmemD = 0, kTmemA = 256;

constexpr uint32_t kSmemEnc = 0;                     // [128 x 64] fp16     16 KiB
constexpr uint32_t kSmemRing = 16384;                // kStages x 32 KiB
constexpr uint32_t kSmemEnc1 = kSmemRing + kStages * kSliceBytes256;      // second ENC tile (render kernel: double buffer)
constexpr uint32_t kSmemConsts = kSmemEnc1 + 16384;                       // fp32 biases + heads of two networks
constexpr uint32_t kConstFloats = kF32WDirPart;      // biases, sigma head, rgb head of one network
constexpr uint32_t kConstRegion = 24576;
constexpr uint32_t kSmemScratch = kSmemConsts + kConstRegion;
static_assert(2 * kConstFloats * 4 <= kConstRegion, "constants of two networks must fit");
static_assert(kSmemEnc1 % 1024 == 0 && kSmemConsts % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");
constexpr uint32_t kSmemTotal = 232448;              // 227 KiB (max opt-in)
constexpr uint32_t kScratchBytes = kSmemTotal - kSmemScratch;

constexpr int kLayersFull = 9;        // L1..L8, fused final.dir
constexpr int kLayersSigma = 8;       // L1..L8

struct Barriers {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t a_ready;        // not used by the engine (the microbenchmarks poll it as a barrier that never completes)
  uint64_t d_free;         // all epilogue warps -> MMA : "the previous tile's accumulator is read out (and, when the
                           //   epilogue writes the ENC tile itself, ENC is written)"
  uint64_t a_kb[4];        // all epilogue warps -> MMA : "A columns of K block kb written"
                           //   (every warp has drained its accumulator columns before its first arrive)
  uint64_t d_ready;        // MMA -> epilogue : "accumulator complete"
  uint32_t tmem_base;
  uint32_t pad[1];
};
static_assert(sizeof(Barriers) % 16 == 0, "Barriers must keep 16-byte alignment of what follows");

// Optional device timeline (experiments only): CTA 0 appends (tag, clock) pairs per role.
struct Timeline {
  long long* buf;      // [3 roles][kTlMax][2]
  int n[3];
};
constexpr int kTlMax = 512;
__device__ __forceinline__ void tl_mark(Timeline* tl, int role, int tag) {
  if (tl == nullptr || tl->buf == nullptr) return;
  int i = tl->n[role];
  if (i >= kTlMax) return;
  tl->buf[(role * kTlMax + i) * 2 + 0] = tag;
  tl->buf[(role * kTlMax + i) * 2 + 1] = clock64();
  tl->n[role] = i + 1;
}
__device__ __forceinline__ void tl_val(Timeline* tl, int role, int tag, long long value) {
  if (tl == nullptr || tl->buf == nullptr) return;
  int i = tl->n[role];
  if (i >= kTlMax) return;
  tl->buf[(role * kTlMax + i) * 2 + 0] = tag;
  tl->buf[(role * kTlMax + i) * 2 + 1] = value;
  tl->n[role] = i + 1;
}

struct RingState {
  uint32_t stage = 0, phase = 0;
  uint32_t n = kStages;      // stages in use (the training-mode forward runs a 2-stage ring and stages its
                             // activation stores through the third stage's 32 KiB)
  __device__ __forceinline__ void advance() {
    if (++stage == n) { stage = 0; phase ^= 1; }
  }
};

// ----------------------------------------------------------------- set-up
// Called by all threads at kernel start.  Returns false (uniformly) on misaligned smem.
__device__ __forceinline__ bool engine_setup(uint8_t* smem, Barriers* bars) {
  const int warp = threadIdx.x >> 5;
  if ((smem_u32(smem) & 1023u) != 0) return false;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(smem_u32(&bars->full[i]), 1);
      mbar_init(smem_u32(&bars->empty[i]), 1);
    }
    mbar_init(smem_u32(&bars->a_ready), kEpiWarps);
    mbar_init(smem_u32(&bars->d_free), kEpiWarps);
    for (int k = 0; k < 4; ++k) mbar_init(smem_u32(&bars->a_kb[k]), kEpiWarps);   // all warps work on one K block at a time
    mbar_init(smem_u32(&bars->d_ready), 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(smem_u32(&bars->tmem_base), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return true;
}
__device__ __forceinline__ void engine_teardown(Barriers* bars) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(bars->tmem_base, kTmemCols);
  }
}

// Copy the fp32 constants (biases + heads) of a network image into shared-memory slot `slot`.
// Called by all threads before the first tile; followed by a __syncthreads().
__device__ __forceinline__ void load_consts(uint8_t* smem, int slot, const uint8_t* __restrict__ blob) {
  if (blob == nullptr) return;
  const float4* src = reinterpret_cast<const float4*>(blob + kHalfRegionBytes);
  float4* dst = reinterpret_cast<float4*>(smem + kSmemConsts) + slot * (kConstFloats / 4);
  for (uint32_t i = threadIdx.x; i < kConstFloats / 4; i += blockDim.x) dst[i] = __ldg(src + i);
}
__device__ __forceinline__ const float* consts_ptr(uint8_t* smem, int slot) {
  return reinterpret_cast<const float*>(smem + kSmemConsts) + slot * kConstFloats;
}

// --------------------------------------------------------------- producer
// One thread.  Streams the slices of one network for one tile, in consumption order.
__device__ __forceinline__ void produce_tile(RingState& rs, uint8_t* smem, Barriers* bars,
                                             const uint8_t* __restrict__ blob, bool sigma_only,
                                             bool dir_slice) {
  const int n256 = sigma_only ? kNumSlicesSigmaOnly : kNumSlices256;
  for (int i = 0; i < n256; ++i) {
    mbar_wait(smem_u32(&bars->empty[rs.stage]), rs.phase ^ 1, 1);
    const uint32_t full = smem_u32(&bars->full[rs.stage]);
#ifdef NERFB200_SANITIZE
    // the previous copy into this stage has landed (implied by `empty`, which the tensor core
    // signals after reading it; stated explicitly for racecheck, which cannot see tcgen05.commit)
    mbar_wait(full, rs.phase ^ 1, 1);
#endif
    const uint32_t dst = smem_u32(smem + kSmemRing + rs.stage * kSliceBytes256);
    mbar_arrive_expect_tx(full, kSliceBytes256);
    const uint8_t* src = blob + static_cast<size_t>(i) * kSliceBytes256;
#pragma unroll
    for (int c = 0; c < 4; ++c) bulk_g2s(dst + c * 8192, src + c * 8192, 8192, full);
    rs.advance();
  }
  if (!sigma_only) {
    const int n128 = dir_slice ? 5 : 4;
    for (int i = 0; i < n128; ++i) {
      mbar_wait(smem_u32(&bars->empty[rs.stage]), rs.phase ^ 1, 2);
      const uint32_t full = smem_u32(&bars->full[rs.stage]);
#ifdef NERFB200_SANITIZE
      mbar_wait(full, rs.phase ^ 1, 2);
#endif
      const uint32_t dst = smem_u32(smem + kSmemRing + rs.stage * kSliceBytes256);
      mbar_arrive_expect_tx(full, kSliceBytes128);
      const uint8_t* src = blob + kOffDir + static_cast<size_t>(i) * kSliceBytes128;
#pragma unroll
      for (int c = 0; c < 2; ++c) bulk_g2s(dst + c * 8192, src + c * 8192, 8192, full);
      rs.advance();
    }
  }
}

// -------------------------------------------------------------------- MMA
// One thread.  Issues all MMAs of one tile.
//
// The tensor pipe queues only about two MMAs ahead of the one it executes (measured with
// nerfb200_debug_mma_contention: ~200 cycles of issuer-side work per four MMAs are hidden, 400 are
// not), so every instruction between two tcgen05.mma counts: the layer / slice structure is fully
// unrolled at compile time (template flags instead of run-time ones), the descriptors are formed
// from one per-stage add, and the only run-time state is the ring stage and the barrier phases.
// With the run-time loop this replaced the issue cadence was 673 cycles per four MMAs; the pipe
// itself needs 4 x 128.
struct MmaPhases { uint32_t d_free = 0, a_kb = 0; };   // a_kb: the 4 K-block barriers flip together

#ifdef NERFB200_TIMELINE
#define NERFB200_TL_MARK(tl, role, tag) tl_mark(tl, role, tag)
#else
#define NERFB200_TL_MARK(tl, role, tag) ((void)0)
#endif

// enc_off: byte offset of this tile's ENC buffer; enc_bar != 0: additionally wait for that
// mbarrier (parity enc_parity) before the first MMA ("ENC written by the helper warps").
template <bool kSigmaOnly, bool kDirSlice>
__device__ __forceinline__ void mma_tile_t(RingState& rs, MmaPhases& ph, uint8_t* smem, Barriers* bars,
                                           Timeline* tl, uint32_t enc_off, uint32_t enc_bar, uint32_t enc_parity) {
  const uint32_t tmem = bars->tmem_base;
  const uint32_t d_tmem = tmem + kTmemD;
  const uint32_t a_tmem = tmem + kTmemA;
  const uint64_t enc_desc = make_desc_sw128(smem_u32(smem + enc_off));
  const uint64_t ring_desc = make_desc_sw128(smem_u32(smem + kSmemRing));
  const uint32_t full0 = smem_u32(&bars->full[0]);
  const uint32_t empty0 = smem_u32(&bars->empty[0]);
  const uint32_t akb0 = smem_u32(&bars->a_kb[0]);
  const uint32_t d_ready = smem_u32(&bars->d_ready);
  const uint32_t idesc256 = make_idesc_f16(256), idesc128 = make_idesc_f16(128);
  constexpr int n_layers = kSigmaOnly ? kLayersSigma : kLayersFull;
#pragma unroll
  for (int l = 0; l < n_layers; ++l) {
    // The first MMA of a layer overwrites the accumulator, so the previous epilogue must have
    // drained it.  Layer 0: "tile start" (d_free, once per tile).  Later layers: a_kb[0] - every
    // warp arrives on it only after its tcgen05.ld of the whole accumulator has completed.  This
    // also covers layer 4, whose first slice (encoded input) needs no A columns at all.
    if (l == 0) {
#ifdef NERFB200_TIMELINE
      const long long w0 = clock64();
#endif
      mbar_wait(smem_u32(&bars->d_free), ph.d_free, 3);
      ph.d_free ^= 1;
#ifdef NERFB200_TIMELINE
      const long long w1 = clock64();
#endif
      if (enc_bar != 0) mbar_wait(enc_bar, enc_parity, 9);
#ifdef NERFB200_TIMELINE
      tl_val(tl, 1, 600, w1 - w0);
      tl_val(tl, 1, 601, clock64() - w1);
#endif
    } else {
      mbar_wait(akb0, ph.a_kb, 6);
    }
    tc_fence_after();
    NERFB200_TL_MARK(tl, 1, 100 + l);
    const int n_slices = (l == 0) ? 1 : (l == 4) ? 5 : (l == 8 && kDirSlice) ? 5 : 4;
    const uint32_t idesc = (l == 8) ? idesc128 : idesc256;
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      if (s < n_slices) {
        const bool from_enc = (l == 0) || (l == 4 && s == 0) || (l == 8 && s == 4);
        const int kb = (l == 4) ? s - 1 : s;
        const uint32_t stage = rs.stage;
        mbar_wait(full0 + 8u * stage, rs.phase, 4);
        if (!from_enc && kb > 0) mbar_wait(akb0 + 8u * kb, ph.a_kb, 6);
        tc_fence_after();
        const uint64_t bdesc = ring_desc + static_cast<uint64_t>(stage * (kSliceBytes256 >> 4));
        if (from_enc) {
#pragma unroll
          for (int j = 0; j < 4; ++j)   // +32 B per K=16 step inside the 128-byte swizzle row
            umma_f16(d_tmem, enc_desc + 2 * j, bdesc + 2 * j, idesc, (s | j) != 0 ? 1u : 0u);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)   // A: 64 K-values per block = 32 columns, 8 per K=16 step
            umma_f16_ts(d_tmem, a_tmem + kb * 32 + j * 8, bdesc + 2 * j, idesc, (s | j) != 0 ? 1u : 0u);
        }
        umma_commit(empty0 + 8u * stage);
        rs.advance();
        NERFB200_TL_MARK(tl, 1, 500 + s);
      }
    }
    umma_commit(d_ready);
    if (l != 0) ph.a_kb ^= 1;
    NERFB200_TL_MARK(tl, 1, 200 + l);
  }
}

__device__ __forceinline__ void mma_tile(RingState& rs, MmaPhases& ph, uint8_t* smem, Barriers* bars,
                                         bool sigma_only, bool dir_slice, Timeline* tl = nullptr,
                                         uint32_t enc_off = kSmemEnc, uint32_t enc_bar = 0, uint32_t enc_parity = 0) {
  if (sigma_only) mma_tile_t<true, false>(rs, ph, smem, bars, tl, enc_off, enc_bar, enc_parity);
  else if (dir_slice) mma_tile_t<false, true>(rs, ph, smem, bars, tl, enc_off, enc_bar, enc_parity);
  else mma_tile_t<false, false>(rs, ph, smem, bars, tl, enc_off, enc_bar, enc_parity);
}

// ---- line-coalesced stores of an epilogue result (training mode / backward chain) -------------
// The four warps that own the same 32 tile rows (column groups 0..3) assemble one [32 rows x 128 B]
// block of the tiled layout (layout.h) in shared memory - each thread 16-byte chunks of its row,
// already at their swizzled positions - and one of them hands the 4 KiB block to the TMA store
// engine.  kBufs staging blocks per row group are used round-robin: a block is rewritten only after
// the bulk store issued kBufs rounds earlier has finished reading it (cp.async.bulk.wait_group.read
// kBufs - 1), so the copy engine's ~400-cycle read latency is off the warps' path for kBufs > 1.
// Two named barriers per round: "the block is free" and "the block is complete".
constexpr int kStageBar0 = 3;          // named barriers 3..6: row groups 0..3 (128 threads each)
constexpr uint32_t kStageBufBytes = 4 * 4096;      // one buffer: 4 row groups x 4 KiB
struct StageCtx {
  uint8_t* base;        // this row group's block in buffer 0; buffer b is kStageBufBytes * b further
  uint32_t buf;         // next buffer
  int rg, lane, part;
};
template <int kBufs>
__device__ __forceinline__ uint8_t* stage_begin(StageCtx& st) {
  if (st.part == 0 && st.lane == 0) bulk_wait_read_n<kBufs - 1>();
  named_bar_sync(kStageBar0 + st.rg, 128);
  return st.base + st.buf * kStageBufBytes;
}
// gdst: the block's place in HBM (null = the rows are padding, nothing is stored)
template <int kBufs>
__device__ __forceinline__ void stage_end(StageCtx& st, uint8_t* block, uint8_t* gdst) {
  fence_proxy_async();
  named_bar_sync(kStageBar0 + st.rg, 128);
#ifdef NERFB200_EXP_NOSTORE       // experiment: staging and barriers, but nothing handed to the copy engine
  gdst = nullptr;
#endif
  if (st.part == 0 && st.lane == 0 && gdst != nullptr) {
    bulk_s2g(gdst, smem_u32(block), 4096);
    bulk_commit();
  }
  st.buf = (st.buf + 1 == kBufs) ? 0u : st.buf + 1;
}
// one round for two adjacent 16-byte chunks (chunk0, chunk0 + 1) per thread
template <int kBufs>
__device__ __forceinline__ void stage_store(StageCtx& st, uint4 c0, uint4 c1, uint32_t chunk0, uint8_t* gdst) {
#ifdef NERFB200_EXP_NOSTAGE       // experiment: no staging at all (results are NOT stored)
  return;
#endif
  uint8_t* block = stage_begin<kBufs>(st);
  const uint32_t sw = static_cast<uint32_t>(st.lane & 7);      // == global row & 7 (row groups are 32-aligned)
  uint8_t* row = block + st.lane * 128;
  *reinterpret_cast<uint4*>(row + ((chunk0 ^ sw) << 4)) = c0;
  *reinterpret_cast<uint4*>(row + (((chunk0 + 1u) ^ sw) << 4)) = c1;
  stage_end<kBufs>(st, block, gdst);
}
constexpr int kFwdStageBufs = 2;       // training-mode forward: the third ring stage holds 2 x 16 KiB

// --------------------------------------------------------------- epilogue
struct EpiCtx {
  uint8_t* smem;
  Barriers* bars;
  const float* __restrict__ f32;   // fp32 region of the current network image (global)
  const float* cst;                // its first kConstFloats floats, resident in shared memory
  uint32_t tmem_row;               // tmem base + (lane quadrant << 16)
  uint32_t d_phase;
  int row;                         // 0..127 : tile row == TMEM lane
  int part;                        // 0..kColSplit-1 : which accumulator column group this warp drains
  int lane;
  Timeline* tl;                    // non-null only for the one traced thread
  // training mode ("save"): post-activation outputs of layers 1..8 and of the direction layer
  // are also written to HBM for the backward pass
  uint8_t* save_act;               // 8 x tiled (save_n, 256) fp16 (layout.h), null = off
  uint2* save_mask;                // [8][save_n][4]: sign bits of this thread's 64 pre-activations per layer
  uint8_t* save_d;                 // tiled (save_n, 128) fp16: output of dir_encoding
  long long save_n;                // padded rows per layer
  long long save_row;              // this thread's global sample row, -1 = padding row
  long long save_g0;               // global sample row of this thread's 32-row group (lane 0), -1 = padding group
  StageCtx stage;                  // shared-memory staging of the stores (training mode)
  // Accumulator release.  false: d_free is signalled at tile start (the epilogue also wrote the
  // ENC tile).  true (render kernel: ENC comes from the helper warps): d_free is signalled as soon
  // as the LAST layer of a tile has been read out of tensor memory, so the next tile's first
  // layer runs while this tile's heads are still being evaluated; `prime` = signal once at the
  // start of the very first tile.
  bool early = false;
  bool prime = true;
};

__device__ __forceinline__ void epi_bar() {   // all 256 epilogue threads
  asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
}
// Tile start: the ENC tile has been written with generic stores (async-proxy fence needed) and
// the previous tile's accumulator reads have retired.
__device__ __forceinline__ void epi_signal_tile_start(EpiCtx& c) {
  fence_proxy_async();
  tc_fence_before();
  __syncwarp();
  if (c.lane == 0) mbar_arrive(smem_u32(&c.bars->d_free));
  NERFB200_TL_MARK(c.tl, 0, 6);
}
// This warp has read its share of the tile's last accumulator: the next tile may overwrite it.
__device__ __forceinline__ void epi_release_accumulator(EpiCtx& c) {
  tc_fence_before();
  __syncwarp();
  if (c.lane == 0) mbar_arrive(smem_u32(&c.bars->d_free));
}
// This warp's A columns of K block kb are stored in TMEM (and, optionally, ENC smem rewritten).
__device__ __forceinline__ void epi_signal_kb(EpiCtx& c, int kb, bool smem_written) {
  if (smem_written) fence_proxy_async();
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (c.lane == 0) mbar_arrive(smem_u32(&c.bars->a_kb[kb]));
}
__device__ __forceinline__ void epi_wait_d(EpiCtx& c) {
  mbar_wait(smem_u32(&c.bars->d_ready), c.d_phase, 5);
  c.d_phase ^= 1;
  tc_fence_after();
}

__device__ __forceinline__ void add_f32x2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{ .reg .b64 a, b, d; mov.b64 a, {%2,%3}; mov.b64 b, {%4,%5}; add.rn.f32x2 d, a, b; mov.b64 {%0,%1}, d; }"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ uint32_t cvt_f16x2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t cvt_f16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// NeRF.forward mode: the ENC tile is dead after layer 5; reuse it for this row's embedded direction
// (27 values, columns 27..63 zero), each column group writes its share of the 64 columns.
__device__ __forceinline__ void write_dir_row(EpiCtx& c, const float* __restrict__ dir_row) {
  uint8_t* enc = c.smem + kSmemEnc;
  constexpr int kPer = 64 / kColSplit;
  const int k0 = c.part * kPer;
#pragma unroll 4
  for (int k = k0; k < k0 + kPer; ++k) {
    const float v = (k < kEncDir) ? __ldg(dir_row + k) : 0.f;
    *reinterpret_cast<__half*>(enc + sw128_off(c.row, k)) = __float2half_rn(v);
  }
}

// Hidden-layer epilogue of this thread's kColsPer accumulator columns.
//   kStore = false: last layer of a sigma-only tile (nothing to hand to the tensor core).
//   dir_row != nullptr (layer 8 in NeRF.forward mode): also rewrite the ENC tile with this row's
//   embedded direction before the last signal.
template <bool kRelu, bool kSigma, bool kStore, bool kSave = false>
__device__ __forceinline__ void epi_hidden(EpiCtx& c, int l, const float* bias, const float* wsig,
                                           float& sig_acc, const float* __restrict__ dir_row = nullptr) {
  NERFB200_TL_MARK(c.tl, 0, 1);
  // the biases of K block 0 are on the critical path "accumulator complete -> first K block handed
  // over": fetch them while waiting for the accumulator (measured: +4 % on the full image)
  float4 pb[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) pb[j] = *reinterpret_cast<const float4*>(bias + c.part * 16 + 4 * j);
  epi_wait_d(c);
  NERFB200_TL_MARK(c.tl, 0, 2);
  {
    // K-block interleaved: every warp owns kColsPer/4 columns of EACH 64-wide K block, so the
    // blocks complete one after the other and the MMA of the next layer can start on block 0
    // while blocks 1..3 are still being converted.
    static_assert(kColsPer / 4 == 16, "the hand-over assumes 16 epilogue warps (16 columns per K block per thread)");
    uint32_t r[4][16];
    uint32_t sgn_lo = 0, sgn_hi = 0;   // kSave: sign bits of the even / odd pre-activations, first in = top bit
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) tmem_ld16(c.tmem_row + kTmemD + kb * 64 + c.part * 16, r[kb]);
    tmem_ld_wait();
    if (!kStore && c.early) epi_release_accumulator(c);
    NERFB200_TL_MARK(c.tl, 0, 3);
    uint32_t hs[kSave ? 4 : 1][8];      // training mode: the layer's fp16 outputs, staged to HBM after the hand-over
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
      const int n0 = kb * 64 + c.part * 16;
      uint32_t (&h)[8] = hs[kSave ? kb : 0];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float4 b0 = (kb == 0) ? pb[2 * j] : *reinterpret_cast<const float4*>(bias + n0 + 8 * j);
        const float4 b1 = (kb == 0) ? pb[2 * j + 1] : *reinterpret_cast<const float4*>(bias + n0 + 8 * j + 4);
        float v[8];
        add_f32x2(v[0], v[1], __uint_as_float(r[kb][8 * j + 0]), __uint_as_float(r[kb][8 * j + 1]), b0.x, b0.y);
        add_f32x2(v[2], v[3], __uint_as_float(r[kb][8 * j + 2]), __uint_as_float(r[kb][8 * j + 3]), b0.z, b0.w);
        add_f32x2(v[4], v[5], __uint_as_float(r[kb][8 * j + 4]), __uint_as_float(r[kb][8 * j + 5]), b1.x, b1.y);
        add_f32x2(v[6], v[7], __uint_as_float(r[kb][8 * j + 6]), __uint_as_float(r[kb][8 * j + 7]), b1.z, b1.w);
        if (kSigma) {
          const float4 w0 = *reinterpret_cast<const float4*>(wsig + n0 + 8 * j);
          const float4 w1 = *reinterpret_cast<const float4*>(wsig + n0 + 8 * j + 4);
          sig_acc = fmaf(fmaxf(v[0], 0.f), w0.x, sig_acc); sig_acc = fmaf(fmaxf(v[1], 0.f), w0.y, sig_acc);
          sig_acc = fmaf(fmaxf(v[2], 0.f), w0.z, sig_acc); sig_acc = fmaf(fmaxf(v[3], 0.f), w0.w, sig_acc);
          sig_acc = fmaf(fmaxf(v[4], 0.f), w1.x, sig_acc); sig_acc = fmaf(fmaxf(v[5], 0.f), w1.y, sig_acc);
          sig_acc = fmaf(fmaxf(v[6], 0.f), w1.z, sig_acc); sig_acc = fmaf(fmaxf(v[7], 0.f), w1.w, sig_acc);
        }
        if (kSave) {
          // backward ReLU masks (bwd_kernels.cuh epi_chain_step): pair i of K block kb ends up at bit
          // 31 - (8 kb + i) of sgn_lo (even element) / sgn_hi (odd element); one SHF per value
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            sgn_lo = __funnelshift_l(__float_as_uint(v[2 * i]), sgn_lo, 1);
            sgn_hi = __funnelshift_l(__float_as_uint(v[2 * i + 1]), sgn_hi, 1);
          }
        }
        if (kRelu) {
          h[4 * j + 0] = cvt_f16x2_relu(v[0], v[1]); h[4 * j + 1] = cvt_f16x2_relu(v[2], v[3]);
          h[4 * j + 2] = cvt_f16x2_relu(v[4], v[5]); h[4 * j + 3] = cvt_f16x2_relu(v[6], v[7]);
        } else {
          h[4 * j + 0] = cvt_f16x2(v[0], v[1]); h[4 * j + 1] = cvt_f16x2(v[2], v[3]);
          h[4 * j + 2] = cvt_f16x2(v[4], v[5]); h[4 * j + 3] = cvt_f16x2(v[6], v[7]);
        }
      }
      if (kStore) {
        tmem_st8(c.tmem_row + kTmemA + n0 / 2, h);
        if (kb == 3 && dir_row != nullptr) write_dir_row(c, dir_row);
        epi_signal_kb(c, kb, kb == 3 && dir_row != nullptr);
        NERFB200_TL_MARK(c.tl, 0, 40 + kb);
      }
    }
    // training mode: the stores happen AFTER all four K blocks have been handed to the tensor core, i.e.
    // behind the next layer's MMAs (staging per K block inside the loop paced the MMA stream: measured
    // 6800 cycles per layer instead of 3350)
    if (kSave && c.save_act != nullptr) {
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        // tiled layout: this row group's 32 rows of column block kb are 4 KiB contiguous in HBM
        uint8_t* gdst = nullptr;
        if (c.save_g0 >= 0)
          gdst = c.save_act + static_cast<long long>(l) * c.save_n * 512 +
                 tiled_block_off(static_cast<unsigned long long>(c.save_g0 >> 6), kb, 4) + (c.save_g0 & 63) * 128;
        stage_store<kFwdStageBufs>(c.stage, make_uint4(hs[kb][0], hs[kb][1], hs[kb][2], hs[kb][3]),
                                   make_uint4(hs[kb][4], hs[kb][5], hs[kb][6], hs[kb][7]), 2u * c.part, gdst);
      }
    }
    if (kSave && c.save_mask != nullptr && c.save_row >= 0)
      c.save_mask[(static_cast<long long>(l) * c.save_n + c.save_row) * 4 + c.part] = make_uint2(sgn_lo, sgn_hi);
  }
  NERFB200_TL_MARK(c.tl, 0, 5);
}

// dir_encoding epilogue (N=128; this thread's 128/kColSplit columns) fused with the rgb head
// (models/nerf.py:119-120): d = relu(acc + dbias[n]); rgb_acc[c] += d * w_rgb[c][n].
// dbias is either the per-ray vector (bias + direction part, shared memory) or b' (shared memory).
template <bool kSave = false>
__device__ __forceinline__ void epi_dir(EpiCtx& c, const float* dbias, const float* wrgb,
                                        float (&rgb_acc)[3]) {
  constexpr int kCols = 128 / kColSplit;     // 32 or 64
  constexpr int kChunks = kCols / 32;
  const int n0 = c.part * kCols;
  uint32_t r[kChunks][32];
  uint32_t dsave[kSave ? 16 * kChunks : 1];
#pragma unroll
  for (int u = 0; u < kChunks; ++u) tmem_ld32(c.tmem_row + kTmemD + n0 + 32 * u, r[u]);
  tmem_ld_wait();
  if (c.early) epi_release_accumulator(c);
#pragma unroll
  for (int u = 0; u < kChunks; ++u) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + u * 32 + 4 * j;
      const float4 b = *reinterpret_cast<const float4*>(dbias + n);
      const float4 wr = *reinterpret_cast<const float4*>(wrgb + n);
      const float4 wg = *reinterpret_cast<const float4*>(wrgb + 128 + n);
      const float4 wb = *reinterpret_cast<const float4*>(wrgb + 256 + n);
      const float v0 = fmaxf(__uint_as_float(r[u][4 * j + 0]) + b.x, 0.f);
      const float v1 = fmaxf(__uint_as_float(r[u][4 * j + 1]) + b.y, 0.f);
      const float v2 = fmaxf(__uint_as_float(r[u][4 * j + 2]) + b.z, 0.f);
      const float v3 = fmaxf(__uint_as_float(r[u][4 * j + 3]) + b.w, 0.f);
      rgb_acc[0] = fmaf(v0, wr.x, rgb_acc[0]); rgb_acc[0] = fmaf(v1, wr.y, rgb_acc[0]);
      rgb_acc[0] = fmaf(v2, wr.z, rgb_acc[0]); rgb_acc[0] = fmaf(v3, wr.w, rgb_acc[0]);
      rgb_acc[1] = fmaf(v0, wg.x, rgb_acc[1]); rgb_acc[1] = fmaf(v1, wg.y, rgb_acc[1]);
      rgb_acc[1] = fmaf(v2, wg.z, rgb_acc[1]); rgb_acc[1] = fmaf(v3, wg.w, rgb_acc[1]);
      rgb_acc[2] = fmaf(v0, wb.x, rgb_acc[2]); rgb_acc[2] = fmaf(v1, wb.y, rgb_acc[2]);
      rgb_acc[2] = fmaf(v2, wb.z, rgb_acc[2]); rgb_acc[2] = fmaf(v3, wb.w, rgb_acc[2]);
      if (kSave) {
        dsave[u * 16 + 2 * j] = cvt_f16x2(v0, v1);
        dsave[u * 16 + 2 * j + 1] = cvt_f16x2(v2, v3);
      }
    }
  }
  if (kSave && c.save_d != nullptr) {
    // output of dir_encoding, tiled (save_n, 128): this thread's 32 columns are 4 chunks of column
    // block part / 2; one staging round per column block
    static_assert(kChunks == 1, "the dir-layer save assumes 16 epilogue warps (32 columns per thread)");
#pragma unroll
    for (int fb = 0; fb < 2; ++fb) {
      uint8_t* gdst = nullptr;
      if (c.save_g0 >= 0)
        gdst = c.save_d + tiled_block_off(static_cast<unsigned long long>(c.save_g0 >> 6), fb, 2) +
               (c.save_g0 & 63) * 128;
      uint8_t* block = stage_begin<kFwdStageBufs>(c.stage);
      if ((c.part >> 1) == fb) {
        const uint32_t sw = static_cast<uint32_t>(c.lane & 7);
        uint8_t* row = block + c.lane * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(row + ((((c.part & 1) * 4u + q) ^ sw) << 4)) =
              make_uint4(dsave[4 * q], dsave[4 * q + 1], dsave[4 * q + 2], dsave[4 * q + 3]);
      }
      stage_end<kFwdStageBufs>(c.stage, block, gdst);
    }
  }
}

// Run the epilogue side of one tile.  Pre-condition: the caller has written the encoded
// input into the ENC tile (all 256 threads) but has NOT signalled yet.
//   dbias      : per-row direction bias vector for this thread's ray (smem, 128 floats), or
//                nullptr to use b_dir from the image (dir_slice mode adds the direction
//                part through the tensor core instead).
//   dir_row    : dir_slice mode only - this row's 27 embedded direction values (global).
// Outputs partial sums (this thread's column group): sigma and rgb pre-activation.
template <bool kSave = false>
__device__ __forceinline__ void epi_run_tile(EpiCtx& c, bool sigma_only, const float* dbias,
                                             const float* __restrict__ dir_row, float& sig_part,
                                             float (&rgb_part)[3]) {
  const float* bias = c.cst + kF32Bias;
  const float* wsig = c.cst + kF32WSigma;
  sig_part = 0.f;
  rgb_part[0] = rgb_part[1] = rgb_part[2] = 0.f;
  float dummy = 0.f;
  if (!c.early || c.prime) epi_signal_tile_start(c);
  c.prime = false;
  for (int l = 0; l < 7; ++l) epi_hidden<true, false, true, kSave>(c, l, bias + l * 256, nullptr, dummy);
  if (sigma_only) {
    epi_hidden<true, true, false, kSave>(c, 7, bias + 7 * 256, wsig, sig_part);
    return;   // the accumulator is released with the next tile's ENC write
  }
  // layer 8's activations feed the fused final.dir layer (layout.h); in NeRF.forward mode the ENC
  // tile is rewritten with this row's embedded direction for the extra K slice
  epi_hidden<true, true, true, kSave>(c, 7, bias + 7 * 256, wsig, sig_part, dir_row);
  epi_wait_d(c);
  epi_dir<kSave>(c, dbias != nullptr ? dbias : (bias + 8 * 256), c.cst + kF32WRgb, rgb_part);
}

__device__ __forceinline__ float sigmoid_ref(float x) { return 1.f / (1.f + expf(-x)); }

}  // namespace nerfb200
