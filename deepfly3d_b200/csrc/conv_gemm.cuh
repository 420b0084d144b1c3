// Implicit-GEMM convolution (1x1 and 3x3, stride 1, "same" padding) on NHWC bf16 activations for
// sm_100a: TMA -> 128B-swizzled shared memory -> tcgen05.mma (fp32 accumulators in TMEM) ->
// tcgen05.ld epilogue with the next layer's BatchNorm/ReLU, the residual add and the bf16
// rounding fused.  This is the hot loop that replaces the torch conv2d/batch_norm/relu calls of
// the hourglass inside df2d (reference call site df3d/core.py:177-185).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace df3d {

struct ConvParams {
  CUtensorMap tmA;    // activations, 4-D (C, W, H, N), box (64, tw, th, nb), SWIZZLE_128B, OOB = 0
  CUtensorMap tmB;    // weights, 2-D (K, CoutPad) K-major, box (64, BN), SWIZZLE_128B
  CUtensorMap tmRes;  // residual / out_raw / out_act: same 4-D box as tmA, one 64-channel slab per
  CUtensorMap tmRaw;  // TMA load (residual) or TMA store (outputs)
  CUtensorMap tmAct;
  CUtensorMap tmRes2; // optional second residual at HALF resolution (nearest x2 up-sampled in the
                      // epilogue): box (64, tw/2, th/2, nb) of the (C, W/2, H/2, N) tensor
  int has_res2;
  // K-concatenation (1x1 convs only): K blocks [kb_split, K/64) are read from a SECOND activation tensor through tmA2
  // (same spatial size), e.g. conv3(t2) + downsample(x) of a bottleneck with a projection shortcut as ONE GEMM with
  // the weights [W3 | Wds] side by side.  0: unused.
  CUtensorMap tmA2;
  int kb_split;
  // pool2 != 0: the tile is 2x2 max-pooled in the epilogue (from the staged raw tile in shared memory) and only
  // the pooled tensors go to HBM: tmPoolRaw = max, tmPoolAct = bf16(relu(max * scale2 + shift2)), boxes
  // (64, tw/2, th/2, nb) of the (C, W/2, H/2, N) tensors.  Needs out_raw (its tensor map is not used), nb == 1.
  CUtensorMap tmPoolRaw, tmPoolAct;
  int pool2;
  // halo9 != 0 (3x3 conv, 64 input channels, BN = 64, 8 x 16 tiles): the nine taps are row-shifted UMMA views of ONE
  // TMA-loaded (tw + 2) x (th + 2) halo tile (tmHalo) and the nine 64 x 64 weight tiles stay resident in shared
  // memory -- an N = 64 shared-memory MMA reads 6 KB per 32 clocks, so with nine A boxes + nine weight tiles
  // streamed per tile the 128 B/clk shared-memory port was the bound (432 KB per tile; 239 KB this way)
  CUtensorMap tmHalo;
  int halo9;
  int n_stages;       // A/B ring depth (filled by launch_conv_gemm from the shared-memory budget)
  int n_res_slots;    // residual ring depth (0 without residual)
  int taps;         // 1 (1x1) or 9 (3x3)
  int kc_per_tap;   // CinPad / 64
  int H, W, B;      // spatial size and number of images of this launch
  int tw, th, nb;   // M tile = nb images x th rows x tw cols = 128 pixels
  int tiles_x, tiles_y, tiles_b;
  int n_tiles_n;    // CoutPad / BN
  // epilogue:  v = acc*scale1[c] + shift1[c] (+ residual) ; relu1 ; out_raw = bf16(v) ;
  //            out_f32 = v ; out_act = bf16(relu(bf16(v)*scale2[c] + shift2[c]))
  const float* scale1;
  const float* shift1;
  const float* scale2;
  const float* shift2;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out_raw;
  __nv_bfloat16* out_act;
  float* out_f32;
  // fp32 path only (score head): arg-max fused into the epilogue -- per (image, channel) one 64-bit key
  // (order-preserving bits of the value << 32 | ~flat pixel index), combined with atomicMax; first occurrence
  // wins ties like the stand-alone kernel.  [B][BN] keys, zero before the launch; needs nb == 1.  May be set
  // with or without out_f32 (the heat-map itself is only stored when the caller asks for it).
  unsigned long long* amax_keys;
  int amax_k;  // real channels (<= BN)
  int res_ld, raw_ld, act_ld, f32_ld;  // channel strides (elements per pixel)
  int relu1;
};

// host side ------------------------------------------------------------------------------------
int tma_init();  // resolves cuTensorMapEncodeTiled through the runtime (no libcuda link dependency)
int make_tmap_act(CUtensorMap* out, const void* base, int C, int W, int H, int N, int tw, int th, int nb);
// same, but the box need not hold 128 pixels (half-resolution residual slabs)
int make_tmap_box(CUtensorMap* out, const void* base, int C, int W, int H, int N, int bw, int bh, int bn);
int make_tmap_wgt(CUtensorMap* out, const void* base, int K, int CoutPad, int BN);
int launch_conv_gemm(const ConvParams& p, int BN, int num_sms, cudaStream_t stream);
int conv_gemm_configure();  // cudaFuncSetAttribute for the dynamic shared memory of every instantiation

}  // namespace df3d
