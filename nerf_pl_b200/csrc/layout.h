// Packed weight image of one NeRF MLP (reference: models/nerf.py:61-81) as the fused kernels
// consume it.  One image per network (coarse / fine).
//
//  [fp16 region]  "Half-slices" in consumption order.  A half-slice is a [128 x 64] fp16 block
//                 (128 output features x 64 input features) in the UMMA K-major SWIZZLE_128B
//                 layout (row n at n*128 B, 16-byte chunk c at position c ^ (n & 7)), i.e. exactly
//                 the shared-memory image, so a plain 1-D bulk copy (cp.async.bulk) stages it.
//                 Every 256-wide layer is computed as two N=128 halves (n0 = outputs 0..127,
//                 n1 = 128..255); for each half the K blocks follow in order:
//       layer 0  xyz_encoding_1      n0:[enc]            n1:[enc]             hs  0.. 1
//       layer 1-3 xyz_encoding_2..4  n0:[k0 k1 k2 k3]    n1:[k0 k1 k2 k3]     hs  2..25
//       layer 4  xyz_encoding_5      n0:[enc k0..k3]     n1:[enc k0..k3]      hs 26..35
//       layer 5-7 xyz_encoding_6..8                                            hs 36..59
//       layer 8  xyz_encoding_final                                            hs 60..67
//       layer 9  dir_encoding        n0:[k0 k1 k2 k3 dir]                      hs 68..72
//                 "enc" = the 63 encoded-xyz input columns (col 63 zero; for layer 4 these are
//                 W[:, 0:63], the hidden part is W[:, 63:319]); "dir" = W_dir[:, 256:283]
//                 (cols 27..63 zero), used only by the stand-alone NeRF.forward entry.
//  [fp32 region]  biases (10 x 256), sigma head, rgb head, dir_encoding W[:, 256:283].
#pragma once
#include <cstdint>

namespace nerfb200 {

constexpr int kW = 256;            // hidden width            (models/nerf.py:43 W=256)
constexpr int kEncXyz = 63;        // in_channels_xyz         (models/nerf.py:44)
constexpr int kEncDir = 27;        // in_channels_dir
constexpr int kDirW = 128;         // W//2                    (models/nerf.py:74)

constexpr uint32_t kHsBytes = 128 * 128;          // 16 KiB per half-slice
constexpr int kNumLayers = 10;
constexpr int kNumHs = 73;
constexpr int kNumHsSigmaOnly = 60;               // layers 0..7
// first half-slice of each layer and number of K blocks per N-half
__host__ __device__ constexpr int hs_layer_start(int l) {
  return l == 0 ? 0 : l <= 4 ? 2 + 8 * (l - 1) : l <= 8 ? 36 + 8 * (l - 5) : 68;
}
__host__ __device__ constexpr int hs_items_per_half(int l) { return l == 0 ? 1 : (l == 4 || l == 9) ? 5 : 4; }

constexpr uint32_t kHalfRegionBytes = kNumHs * kHsBytes;     // 1,196,032

// fp32 region (offsets in floats from the start of the region)
constexpr int kNumBiasRows = 10;                 // b1..b8, b_final, b_dir(128 used)
constexpr int kF32Bias = 0;                      // [10][256]
constexpr int kF32WSigma = kF32Bias + kNumBiasRows * 256;   // [256]
constexpr int kF32BSigma = kF32WSigma + 256;     // [4]  (1 used)
constexpr int kF32WRgb = kF32BSigma + 4;         // [3][128]
constexpr int kF32BRgb = kF32WRgb + 3 * 128;     // [4]  (3 used)
constexpr int kF32WDirPart = kF32BRgb + 4;       // [128][28] (27 used)
constexpr int kF32Count = kF32WDirPart + 128 * 28;
constexpr uint32_t kPackedBytes = kHalfRegionBytes + kF32Count * 4;

// Parameter order of the 24 tensors handed to the pack routine
// (state_dict order of models/nerf.py NeRF):
//  0..15  xyz_encoding_{1..8}.0.{weight,bias}
//  16,17  xyz_encoding_final.{weight,bias}
//  18,19  dir_encoding.0.{weight,bias}
//  20,21  sigma.{weight,bias}
//  22,23  rgb.0.{weight,bias}
constexpr int kNumParams = 24;

}  // namespace nerfb200
