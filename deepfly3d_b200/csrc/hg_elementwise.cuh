// Launchers of the elementwise hourglass stages (see hg_elementwise.cu).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace df3d {

constexpr int kStemKPadCols = 192;  // 7*7*3 = 147 patch columns, zero padded to 3 x 64
constexpr int kStemKGray = 64;      // gray fast path: 7*7 = 49 patch columns, zero padded to 64
constexpr int kStemGrayMaxW = 1024; // widest input row the gray fast path stages in shared memory

int launch_stem_im2col(const void* img, int dtype, const uint8_t* flip, int B, int H, int W, const float mean[3],
                       __nv_bfloat16* out, cudaStream_t s);
// gray fast path (uint8 input, one common mean): patch column = ky*7 + kx
int launch_stem_im2col_gray(const uint8_t* img, const uint8_t* flip, int B, int H, int W, float mean, __nv_bfloat16* out,
                            cudaStream_t s);
int launch_maxpool_bn_relu(const __nv_bfloat16* in, int B, int H, int W, int C, const float* scale, const float* shift,
                           __nv_bfloat16* out_raw, __nv_bfloat16* out_act, cudaStream_t s);

}  // namespace df3d
