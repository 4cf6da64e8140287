/* nerf_pl_b200 diagnostics — bring-up probes and tcgen05 microbenchmarks.
 *
 * NOT part of the drop-in boundary and NOT in the product library: these entry points exist only
 * in builds made with -DNERFB200_DIAG (python tools/build_variants.py diag=-DNERFB200_DIAG, then
 * NERFB200_LIB=nerf_pl_b200/variants/lib_diag.so).  Same conventions as nerf_pl_b200.h.
 */
#ifndef NERF_PL_B200_DIAG_H_
#define NERF_PL_B200_DIAG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One K=64 weight slice of a packed image (csrc/layout.h) against a (128,64) fp32 A tile
 * through the tcgen05 engine: d (128, N), N = 256 for slices 0..33 and 128 for 34..38.
 * mode 0 stages A in shared memory (SS MMA), mode 1 in tensor memory (TS MMA).  Unit-test hook
 * for the operand layouts; not part of the reference API. */
int nerfb200_debug_gemm(const float* a, const void* packed, int32_t slice, int32_t mode, float* d,
                        void* stream);
/* Raw tcgen05.mma issue-rate microbenchmark (timing only): out_dev (n_ctas, 8) int64 device
 * buffer; column v = SM cycles for reps x 16 MMAs of variant v (csrc/aux_kernels.cuh). */
int nerfb200_debug_mma_bench(int64_t* out_dev, int32_t n_ctas, int32_t reps, void* stream);
/* tcgen05.mma vs. concurrent tcgen05.ld/st microbenchmark (timing only): out_dev (n_ctas, 4)
 * int64; [0] = SM cycles of reps x 16 MMAs, [1] = background iterations meanwhile
 * (bg / variant codes: csrc/aux_kernels.cuh mma_contention_kernel). */
int nerfb200_debug_mma_contention(int64_t* out_dev, int32_t n_ctas, int32_t reps, int32_t bg, int32_t variant,
                                  void* stream);
/* Experiment hook: with NERFB200_FLAGS bit 1 set, CTA 0 of the last render launch records
 * (tag, SM clock) pairs for its epilogue / MMA roles; this copies 3*512*2 int64 to host. */
int nerfb200_debug_timeline(int64_t* host_out, int64_t n_values);
/* MN-major operand probe (the layout the wgrad kernel uses): d (128, 256) = a^T b for
 * a (64, 128), b (64, 256) fp32 inputs rounded to fp16, both staged in shared memory as
 * [64-feature block][64 rows = samples][128 B] SWIZZLE_128B images and read by tcgen05.mma with
 * MN-major descriptors (lbo / sbo in bytes).  fmt bit 0: A staged as bf16, bit 1: B as bf16
 * (mixed formats = what the wgrad kernel issues: bf16 gradients x fp16 activations). */
int nerfb200_debug_gemm_mn(const float* a, const float* b, int32_t lbo, int32_t sbo, int32_t fmt, float* d,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERF_PL_B200_DIAG_H_ */
