// Thin inline-PTX wrappers for the sm_100a features the renderer uses:
// mbarrier, 1-D bulk async copy (TMA engine, UBLKCP), tcgen05 MMA / TMEM.
// Hand-written; no CUTLASS/CuTe dependency.  Compile with
//   -gencode arch=compute_100a,code=sm_100a
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nerfb200 {

// Set by a device-side watchdog when an mbarrier wait exceeds its budget.
// Checked by the host wrapper after launch errors (diagnostic only).
__device__ int g_watchdog_code = 0;

// Device-side fault report into the caller's status word (may live in mapped host memory).
__device__ __forceinline__ void report_fault(int* status, int code) {
  if (status == nullptr) return;
  *reinterpret_cast<volatile int*>(status) = code;
  __threadfence_system();
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ----------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hang
// (a hung GPU box costs a strike).  ~2 s at 2 GHz.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int code) {
  if (mbar_try_wait(bar, parity)) return;
#ifdef NERFB200_SANITIZE
  while (!mbar_try_wait(bar, parity)) {}     // tool runs are 100x slower: no time limit
  return;
#endif
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      g_watchdog_code = code;
      __threadfence_system();
      __trap();
    }
  }
}

// Make generic-proxy smem writes visible to the async proxy (UMMA / TMA readers).
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------ bulk copy (global->smem)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// ------------------------------------------------------ bulk copy (smem->global, TMA store engine)
// Writes full 128-byte lines however the data was produced: the epilogue threads own one 16-byte
// chunk of 32 different rows each, which as plain st.global costs 32 L1 wavefronts per warp
// instruction; staged through shared memory one thread hands a contiguous block to the copy engine.
__device__ __forceinline__ void bulk_s2g(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores this thread has committed have finished READING shared memory (the source may be reused)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... all but the N most recently committed ones
template <int N>
__device__ __forceinline__ void bulk_wait_read_n() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// --------------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, fp16 operands, fp32 accumulate, one CTA.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the A operand in tensor memory (lane = row, 32-bit column c holds K elements
// 2c (low half) and 2c+1 (high half)); B from shared memory.
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued MMAs of this thread retire.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row
// (lane-quadrant base + i), columns [col, col+32).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 16-column load / 8-column store variants (K-block interleaved epilogue).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread i writes row (lane-quadrant base + i).
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, K-major, SWIZZLE_128B, 16-bit elements:
// rows are 128 B apart, 8-row groups 1024 B apart (SBO), 16-byte chunk c of row r
// lives at chunk position c ^ (r & 7).  Tile base must be 1024-byte aligned.
// Field layout (PTX ISA "matrix descriptor", sm_100): [0,14) addr>>4, [16,30) LBO>>4,
// [32,46) SBO>>4, [46,48) version=1, [61,64) layout (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;            // LBO (ignored for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO
  d |= static_cast<uint64_t>(1) << 46;            // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: fp16 A/B (K-major both), fp32 D, M=128.
__host__ __device__ constexpr uint32_t make_idesc_f16(int n) {
  return (1u << 4)                                  // D format = F32
         | (0u << 7) | (0u << 10)                   // A, B format = F16
         | (0u << 15) | (0u << 16)                  // A, B K-major
         | (static_cast<uint32_t>(n >> 3) << 17)    // N / 8
         | (static_cast<uint32_t>(128 >> 4) << 24); // M / 16
}

// MN-major shared-memory descriptor, SWIZZLE_128B, 16-bit elements (the canonical layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): one row per K index (= sample), 128 B =
// 64 consecutive M/N elements per row, 16-byte chunk c of row r at chunk position c ^ (r & 7);
// 8-row groups SBO bytes apart along K, 64-element blocks LBO bytes apart along M/N.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor with both operands MN-major (bits 15 / 16), fp16 in, fp32 out, M=128.
__host__ __device__ constexpr uint32_t make_idesc_f16_mn(int n) {
  return make_idesc_f16(n) | (1u << 15) | (1u << 16);
}

// Byte offset of element (row, k) inside a [rows x 64] fp16 SWIZZLE_128B block.
__host__ __device__ __forceinline__ constexpr uint32_t sw128_off(uint32_t row, uint32_t k) {
  return row * 128u + ((((k >> 3) ^ (row & 7u)) & 7u) << 4) + ((k & 7u) << 1);
}

}  // namespace nerfb200
