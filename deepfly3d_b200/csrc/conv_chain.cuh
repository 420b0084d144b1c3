// Convolution *chains* for sm_100a: one implicit-GEMM conv (1x1 or 3x3, the "head") followed by up to
// four point-wise (1x1) convs evaluated on the same 128-pixel tile without leaving the SM.
//
// Every hourglass bottleneck is  relu(bn) -> 1x1 -> relu(bn) -> 3x3 -> relu(bn) -> 1x1 (+ residual);
// unfused, the 128/256-channel tensors between those convs make the 1x1 layers HBM-bound.  In a
// chain, the epilogue of stage i turns its fp32 accumulator (TMEM) into the bf16 activation the next
// conv consumes and stores it back into TMEM as the A operand of the next tcgen05.mma (A-from-TMEM
// form), so only the tensors that other layers need (the residual stream, the next block's 3x3
// input) are written to HBM.  Replaces the torch conv2d/batch_norm/relu/add calls of the hourglass
// inside df2d (reference call site df3d/core.py:177-185).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace df3d {

constexpr int kMaxChain = 5;

struct ChainStage {
  CUtensorMap tmB;     // weights, 2-D (K, N) K-major, box (64, 64) = one CTA's half of a 128-row block, SWIZZLE_128B
  CUtensorMap tmRes;   // residual added in this stage's epilogue: 4-D box (64, tw, th, nb)
  CUtensorMap tmRes2;  // second residual at half resolution (nearest x2 up-sample + add)
  CUtensorMap tmOutQ;  // stages with out_raw: out_raw as a TMA store target, box = one warp's quarter of
                       // the tile (64 channels x 32 consecutive pixels, see make_tmap_quarter)
  // epilogue:  v = acc*scale1[c] + shift1[c] (+ residual) (+ up(residual2)) ; relu1 ;
  //            raw = bf16(v) ; act = bf16(relu(raw*scale2[c] + shift2[c]))
  const float* scale1;
  const float* shift1;
  const float* scale2;
  const float* shift2;
  __nv_bfloat16* out_raw;  // bf16 output of this stage, NHWC with `n` channels per pixel, or null (stored through tmOutQ
                           // from the residual slab when the stage has a residual, else by 256-bit stores)
  // 2x2 max-pool of out_raw taken in the epilogue (stages with a residual and out_raw on 8-wide tiles): each epilogue
  // warp pools its own 4 x 8 pixel quarter of the slab it has just written in place and stores
  //   pool_raw = max,  pool_act = bf16(relu(max * pool_scale[c] + pool_shift[c]))      (NHWC at half resolution, n channels)
  // -- what maxpool_bn_relu_kernel computes from the stored tensor, without the extra pass over HBM
  __nv_bfloat16* pool_raw;
  __nv_bfloat16* pool_act;
  const float* pool_scale;
  const float* pool_shift;
  int n;         // output channels: 128 or 256
  int kblocks;   // K / 64 (head: taps * Cin/64; later stages: n of the previous stage / 64)
  // Later stages may take `ss_kblocks` MORE K blocks whose A operand is a second activation tensor read from shared
  // memory (tmA2, same tile as the head: box (64, tw, th, nb)) instead of the previous stage's result in tensor memory:
  //   acc = W[:, :64 kblocks] * x_prev + W[:, 64 kblocks:] * a2
  // -- e.g. fc(conv3(t) + h) = (W_fc W_3) t + W_fc h with t from the previous stage and h from HBM.  These MMAs do not
  // depend on the previous epilogue and are issued first.  Needs n = 256.
  CUtensorMap tmA2;
  int ss_kblocks;
  int relu1;
  int unit_scale;  // scale1 is identically 1 (conv without a folded BatchNorm): the epilogue only adds shift1
  int has_res, has_res2;
  int x_src;     // operand handed to the next stage: 0 none (last stage), 1 raw, 2 act
  // tensor-memory plan (filled by launch_conv_chain, see plan_tmem)
  int col[2][2]; // [tile parity][half]: accumulator columns of channels [0,128) and [128,256) (the latter unused for
                 // n = 128); plans whose column assignment alternates between consecutive tiles differ by parity
  int hz_stage;  // epilogue that must have drained those columns before this stage is issued (-1: implied)
  int hz_delta;  // ... of this tile (0) or of the previous one (1)
  // filled by launch_conv_chain: a stored stage WITHOUT a residual writes its output into a blank slab of the
  // residual ring and leaves through tmOutQ like the stages with one (needs tmOutQ).  256-bit stores from registers
  // cost the L1 data pipe ~6x the wavefronts of the same bytes staged through shared memory, and that pipe (LSU +
  // tensor-core operand reads) is what bounds the chains
  int stage_out;
  int aff_off;   // float offset of this stage's constants in shared memory (filled by the launcher)
  int pool_off;  // ... of the pooled output's scale / shift (filled by the launcher)
  int epi_kind;  // specialised epilogue variant (filled by the launcher)
};

struct ChainParams {
  CUtensorMap tmA;  // head activations, 4-D (C, W, H, N), box (64, tw, th, nb), SWIZZLE_128B, OOB = 0
  CUtensorMap tmHalo;  // halo mode: box (64, tw + 2, th + 2, 1) of the same tensor (see conv_chain.cu)
  ChainStage st[kMaxChain];
  int n_chain;
  int halo;        // != 0: the 3x3 head reads nine row-shifted views of one halo tile instead of nine TMA boxes
  int taps;        // head: 1 or 9
  int kc_per_tap;  // head: Cin / 64
  int H, W, B;
  int tw, th, nb;  // M tile = nb images x th rows x tw cols = 128 pixels
  int tiles_x, tiles_y, tiles_b;
  // shared-memory carve-up (filled by launch_conv_chain)
  int head_after;  // the head GEMM of the next tile is issued behind this stage (filled by launch_conv_chain)
  int n_m;         // operand ring slots
  int slot_bytes;  // per CTA: 16 KB of weights (two 64-row sub-tiles), or a head K block (A tile + its sub-tiles)
  int tx_shift, ty_shift;  // log2 of tiles_x / tiles_y when they are powers of two, else -1
  int n_slabs;     // residual slabs (TMA prefetch ring)
  int slab_bytes;  // 16384 (+4096 with a half-resolution residual)
  int aff_bytes;
  // profiling only (tools/chain_probe.cu): when non-null, CTA 0 records clock64() time stamps of its MMA
  // issuer ([0, 4096)) and of epilogue warps 4 and 8 ([4096, 8192), [8192, 12288))
  unsigned long long* dbg;
  int dbg_exec;  // probe only: the MMA warp also waits for (and stamps) the completion of every stage >= 1
};

int conv_chain_configure();
int launch_conv_chain(const ChainParams& p, int num_sms, cudaStream_t stream);
// TMA box of 32 consecutive pixels of a (tw, th, nb) tile: (64, tw, 32/tw, 1) or (64, tw, th, 32/(tw*th))
int make_tmap_quarter(CUtensorMap* out, const void* base, int C, int W, int H, int N, int tw, int th, int nb);

}  // namespace df3d
