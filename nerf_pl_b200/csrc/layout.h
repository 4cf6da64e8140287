// Packed weight image of one NeRF MLP (reference: models/nerf.py:61-81) as the
// fused kernels consume it.  One image per network (coarse / fine).
//
//  [fp16 region]  K-slices in consumption order.  A slice is a [N x 64] fp16 block in
//                 the UMMA K-major SWIZZLE_128B layout (row n at n*128 B, 16-byte chunk c
//                 at position c ^ (n & 7)), i.e. exactly the shared-memory image, so a
//                 plain 1-D bulk copy (cp.async.bulk) stages it.
//       slice 0        xyz_encoding_1  W[:, 0:63]           (col 63 zero)      N=256
//       slice 1..12    xyz_encoding_2..4, 4 K-blocks each                       N=256
//       slice 13       xyz_encoding_5  W[:, 0:63]  (skip: encoded-input part)   N=256
//       slice 14..17   xyz_encoding_5  W[:, 63:319] (hidden part)               N=256
//       slice 18..29   xyz_encoding_6..8                                        N=256
//       slice 30..33   xyz_encoding_final                                       N=256
//       slice 34..37   dir_encoding    W[:, 0:256]                              N=128
//       slice 38       dir_encoding    W[:, 256:283] (cols 27..63 zero)         N=128
//  [fp32 region]  biases (10 x 256), sigma head, rgb head, dir_encoding W[:, 256:283].
#pragma once
#include <cstdint>

namespace nerfb200 {

constexpr int kW = 256;            // hidden width            (models/nerf.py:43 W=256)
constexpr int kEncXyz = 63;        // in_channels_xyz         (models/nerf.py:44)
constexpr int kEncDir = 27;        // in_channels_dir
constexpr int kDirW = 128;         // W//2                    (models/nerf.py:74)

constexpr uint32_t kSliceBytes256 = 256 * 128;   // 32 KiB
constexpr uint32_t kSliceBytes128 = 128 * 128;   // 16 KiB
constexpr int kNumSlices256 = 34;
constexpr int kNumSlicesSigmaOnly = 30;          // layers 1..8 only
constexpr int kNumSlices128 = 5;                 // 4 hidden + 1 direction-part slice

constexpr uint32_t kOffDir = kNumSlices256 * kSliceBytes256;                 // 1,114,112
constexpr uint32_t kHalfRegionBytes = kOffDir + kNumSlices128 * kSliceBytes128;  // 1,196,032

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
