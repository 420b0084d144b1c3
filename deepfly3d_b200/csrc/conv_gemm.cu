// tcgen05 implicit-GEMM convolution kernel (see conv_gemm.cuh).
//
// GEMM view:  D[M = pixels, N = Cout] = sum over taps, cin of A[pixel shifted by tap, cin] * W[cout, tap, cin]
//   * A tile: 128 pixels (nb x th x tw box of the NHWC tensor) x 64 channels, loaded by ONE 4-D TMA
//     per (tap, channel block); the 3x3 halo / zero padding is the TMA's out-of-bounds zero fill
//     (coordinates x0+dx-1, y0+dy-1 may be -1 or W/H).
//   * B tile: BN output channels x 64 K, 2-D TMA from the packed [CoutPad][taps*CinPad] weights.
//   * one CTA per SM, persistent over tiles, 12 warps:
//       warp 0  TMA producer of the A/B ring          warp 1  MMA issuer (one thread) + TMEM owner
//       warp 2  TMA producer of the residual ring     warp 3  idle
//       warps 4..11  epilogue: two groups of four warps (one TMEM lane quarter each); group g takes the
//                    tiles of accumulator stage g (every other tile).  With one group the small-K convs
//                    (1x1, K = 64..256: the 128x128-resolution layers, the pooled conv1s) were bound by the
//                    epilogue -- 4 200 cycles per 128-pixel tile for an HBM cost of 1 900 (measured).
//   * TMEM holds two accumulator stages (2 x BN fp32 columns): the epilogue of tile i overlaps the
//     MMAs of tile i+1 and the epilogue of tile i-1.
//   * epilogue, per 64-channel slab: tcgen05.ld -> scale/shift (+ residual slab from shared memory,
//     prefetched by TMA) -> ReLU / bf16 rounding -> 128B-swizzled shared-memory slab -> TMA store.
//     Every global access of the kernel is a TMA bulk transfer (fully coalesced, asynchronous); the
//     only direct stores left are the fp32 score maps of the last stack.
#include <cstdlib>

#include "conv_gemm.cuh"
#include "epilogue.cuh"
#include "sm100.cuh"

namespace df3d {

using namespace sm100;

constexpr int kConvThreads = 384;
constexpr int kEpiWarp0 = 4;            // first epilogue warp
constexpr int kEpiThreads = 128;
constexpr int kTileM = 128;
constexpr int kABytes = kTileM * 128;   // 128 rows x 64 bf16
constexpr int kSlabBytes = kTileM * 128;  // one 64-channel bf16 slab of an output / residual tile
constexpr int kHalfSlabBytes = kSlabBytes / 4;  // same slab at half resolution (32 pixels)
constexpr int kMaxStages = 8;
constexpr int kMaxResSlots = 4;
constexpr int kAffBytes = 4 * 256 * 4;  // scale1, shift1, scale2, shift2 for up to 256 channels
constexpr int kBarBytes = 512;
constexpr int kSmemLimit = 232448;      // 227 KB opt-in maximum per CTA
constexpr int kHalo9Pitch = 10;                          // halo row of an 8-wide tile
constexpr int kHalo9Bytes = 18 * kHalo9Pitch * 128;      // (16 + 2) x (8 + 2) pixels x 64 channels
constexpr int kHalo9Slot = 23552;                        // ... padded to a multiple of 1024 (swizzle atom alignment)

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_round(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }

template <int BN>
__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvParams p) {
  constexpr int kBBytes = BN * 128;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
  constexpr int kSlabs = BN >= 64 ? BN / 64 : 1;   // 64-channel slabs per tile (bf16 outputs)

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 128B swizzle: 1024-byte aligned tiles
  const int n_stages = p.n_stages, n_res = p.n_res_slots;
  const bool pool2 = p.pool2 != 0;
  const bool has_res = n_res > 0, has_raw = p.out_raw != nullptr, has_act = p.out_act != nullptr || pool2;
  const bool has_res2 = has_res && p.has_res2 != 0;
  const uint32_t res_slot_bytes = kSlabBytes + (has_res2 ? kHalfSlabBytes : 0);
  // carve-up: [A/B ring | nine resident weight tiles + halo ring][residual ring][raw out x2][act out x2][affine][barriers]
  const bool halo9 = p.halo9 != 0;
  const uint32_t halo_ring = smem_base + 9u * kBBytes;
  const uint32_t res_base = halo9 ? halo_ring + n_stages * kHalo9Slot : smem_base + n_stages * kStageBytes;
  const uint32_t raw_base = res_base + n_res * res_slot_bytes;
  const uint32_t act_base = raw_base + (has_raw ? 2 * kSlabBytes : 0);
  const uint32_t aff_base = act_base + (has_act ? 2 * kSlabBytes : 0);
  const uint32_t bar_base = aff_base + kAffBytes;
  auto a_addr = [&](int s) { return smem_base + s * kStageBytes; };
  auto b_addr = [&](int s) { return smem_base + s * kStageBytes + kABytes; };
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 2 + s); };
  auto rfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 4 + s); };
  auto rempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 4 + kMaxResSlots + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 4 + 2 * kMaxResSlots);
  const uint32_t wfull_bar = tmem_slot + 8u;  // halo9: the resident weights have landed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  const int total_tiles = m_tiles * p.n_tiles_n;
  const int num_kb = p.taps * p.kc_per_tap;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.tmA);
    prefetch_tensormap(&p.tmB);
    if (has_res) prefetch_tensormap(&p.tmRes);
    if (has_res2) prefetch_tensormap(&p.tmRes2);
    if (has_raw && !pool2) prefetch_tensormap(&p.tmRaw);
    if (has_act && !pool2) prefetch_tensormap(&p.tmAct);
    if (pool2) {
      prefetch_tensormap(&p.tmPoolRaw);
      prefetch_tensormap(&p.tmPoolAct);
    }
    if (p.kb_split > 0) prefetch_tensormap(&p.tmA2);
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);  // one arrive per epilogue warp
    }
    for (int s = 0; s < n_res; ++s) {
      mbar_init(rfull_bar(s), 1);
      mbar_init(rempty_bar(s), 4);
    }
    mbar_init(wfull_bar, 1);
    if (halo9) prefetch_tensormap(&p.tmHalo);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  // per-channel epilogue constants -> shared memory (CoutPad <= 256)
  {
    const int ncol = p.n_tiles_n * BN;
    for (int i = threadIdx.x; i < ncol; i += kConvThreads) {
      float s1 = p.scale1[i], h1 = p.shift1[i], s2 = 0.f, h2 = 0.f;
      if (p.scale2) {
        s2 = p.scale2[i];
        h2 = p.shift2[i];
      }
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(aff_base + 4u * i), "f"(s1));
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(aff_base + 1024u + 4u * i), "f"(h1));
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(aff_base + 2048u + 4u * i), "f"(s2));
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(aff_base + 3072u + 4u * i), "f"(h2));
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  auto decode_tile = [&](int tile, int& nt, int& x0, int& y0, int& n0) {
    nt = tile % p.n_tiles_n;
    int mt = tile / p.n_tiles_n;
    const int tx = mt % p.tiles_x;
    mt /= p.tiles_x;
    const int ty = mt % p.tiles_y;
    const int tb = mt / p.tiles_y;
    x0 = tx * p.tw;
    y0 = ty * p.th;
    n0 = tb * p.nb;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ A/B producer
    if (lane == 0 && halo9) {
      mbar_arrive_expect_tx(wfull_bar, 9u * kBBytes);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(smem_base + tap * kBBytes, &p.tmB, wfull_bar, tap * 64, 0);
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int nt, x0, y0, n0;
        decode_tile(tile, nt, x0, y0, n0);
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_arrive_expect_tx(full_bar(stage), kHalo9Bytes);
        tma_load_4d(halo_ring + stage * kHalo9Slot, &p.tmHalo, full_bar(stage), 0, x0 - 1, y0 - 1, n0);
        if (++stage == (uint32_t)n_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    } else if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int nt, x0, y0, n0;
        decode_tile(tile, nt, x0, y0, n0);
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / p.kc_per_tap, kc = kb - tap * p.kc_per_tap;
          int dx = 0, dy = 0;
          if (p.taps == 9) {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
          }
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), kStageBytes);
          const bool second = p.kb_split > 0 && kb >= p.kb_split;  // K-concatenation: the second activation tensor
          tma_load_4d(a_addr(stage), second ? &p.tmA2 : &p.tmA, full_bar(stage), (second ? kc - p.kb_split : kc) * 64, x0 + dx,
                      y0 + dy, n0);
          tma_load_2d(b_addr(stage), &p.tmB, full_bar(stage), kb * 64, nt * BN);
          if (++stage == (uint32_t)n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs the loop (warp-uniform control flow and operands: descriptors live in
    // uniform registers, no per-lane waterfall around each tcgen05 instruction); one elected lane
    // issues.  The issue rate of this loop bounds the tensor pipe.
    {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
      const uint64_t desc_hi = umma_smem_desc_sw128(0);
      uint32_t stage = 0, phase = 0, it = 0;
      if (halo9) {
        // A operand of tap (dy, dx) = the halo rows shifted by dy * 10 + dx: 8-row groups 1280 B apart starting at a
        // row offset (the 128B swizzle is a function of the absolute shared-memory address, which is what TMA wrote)
        const uint64_t desc_halo = (desc_hi & ~((uint64_t)0x3FFF << 32)) | ((uint64_t)((kHalo9Pitch * 128) >> 4) << 32);
        mbar_wait_warp(wfull_bar, 0u);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
          const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
          mbar_wait_warp(tempty_bar(as), aphase ^ 1u);
          mbar_wait_warp(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * BN;
          const uint32_t hbase = halo_ring + stage * kHalo9Slot;
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t a_ad = hbase + (uint32_t)((tap / 3) * kHalo9Pitch + tap % 3) * 128u;
              const uint64_t adesc = desc_halo | (uint64_t)((a_ad >> 4) & 0x3FFFu);
              const uint64_t bdesc = desc_hi | (uint64_t)(((smem_base + tap * kBBytes) >> 4) & 0x3FFFu);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (tap | k) != 0 ? 1u : 0u);
            }
            umma_commit(empty_bar(stage));
            umma_commit(tfull_bar(as));
          }
          __syncwarp();
          if (++stage == (uint32_t)n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      } else
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
        mbar_wait_warp(tempty_bar(as), aphase ^ 1u);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_warp(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t adesc = desc_hi | (uint64_t)((a_addr(stage) >> 4) & 0x3FFFu);
          const uint64_t bdesc = desc_hi | (uint64_t)((b_addr(stage) >> 4) & 0x3FFFu);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)  // 4 x (K = 16): +32 bytes inside the 128B swizzle atom
              umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(empty_bar(stage));  // frees the smem stage once these MMAs retire
          }
          __syncwarp();
          if (++stage == (uint32_t)n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (elect_one()) umma_commit(tfull_bar(as));  // accumulator ready for the epilogue
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ residual producer
    if (lane == 0 && has_res) {
      uint32_t slot = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int nt, x0, y0, n0;
        decode_tile(tile, nt, x0, y0, n0);
        for (int sl = 0; sl < kSlabs; ++sl) {
          mbar_wait(rempty_bar(slot), phase ^ 1u);
          mbar_arrive_expect_tx(rfull_bar(slot), res_slot_bytes);
          tma_load_4d(res_base + slot * res_slot_bytes, &p.tmRes, rfull_bar(slot), nt * BN + sl * 64, x0, y0, n0);
          if (has_res2)
            tma_load_4d(res_base + slot * res_slot_bytes + kSlabBytes, &p.tmRes2, rfull_bar(slot), nt * BN + sl * 64,
                        x0 >> 1, y0 >> 1, n0);
          if (++slot == (uint32_t)n_res) {
            slot = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue (warps 4..11)
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int grp = (warp - kEpiWarp0) >> 2;  // accumulator stage / tile parity of this group
    const int m = q * 32 + lane;       // row of the tile = pixel
    const bool leader = (q == 0) && (lane == 0);
    const uint32_t bar_a = 1u + 2u * (uint32_t)grp, bar_b = 2u + 2u * (uint32_t)grp;
    const uint32_t row_off = (uint32_t)m * 128u;
    const uint32_t sw = (uint32_t)(m & 7);
    // row of this pixel's parent in the half-resolution residual slab (box tw/2 x th/2 x nb)
    uint32_t row2_off = 0, sw2 = 0;
    if (has_res2) {
      const int w = m % p.tw, h = (m / p.tw) % p.th, nl = m / (p.tw * p.th);
      const int r2 = (nl * (p.th >> 1) + (h >> 1)) * (p.tw >> 1) + (w >> 1);
      row2_off = (uint32_t)r2 * 128u;
      sw2 = (uint32_t)(r2 & 7);
    }
    // fast path (epilogue.cuh): two pixel rows per lane -- lanes l and l + 16 share rows r16, r16 + 16 of the
    // quarter and split a slab's 64 channels (every per-channel constant is fetched once for two rows)
    const uint32_t hh = (uint32_t)lane >> 4;
    uint32_t row_off2[2], sw_2[2];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const int m2 = q * 32 + 16 * w + (lane & 15);
      row_off2[w] = (uint32_t)m2 * 128u;
      sw_2[w] = (uint32_t)(m2 & 7);
    }
    const uint8_t* const smg = smem_raw + (smem_base - smem_u32(smem_raw));  // generic view of the carve-up
    // the combinations the hourglass plans use most; anything else takes the generic loop below
    const bool fast = has_raw && !has_res2 && !(has_res && p.relu1) && !p.out_f32;
    uint32_t it = 0, rslot = 0, rphase = 0;
    const uint32_t obuf = (uint32_t)grp;  // one staging buffer per group and output kind
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
      if (as != (uint32_t)grp) {  // the other group's tile: only keep the residual ring position in step
        if (has_res && !p.out_f32) {
          rslot += kSlabs;
          while (rslot >= (uint32_t)n_res) {
            rslot -= (uint32_t)n_res;
            rphase ^= 1u;
          }
        }
        continue;
      }
      int nt, x0, y0, n0;
      decode_tile(tile, nt, x0, y0, n0);
      mbar_wait_warp(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;

      if (p.out_f32 || p.amax_keys) {
        // fp32 score maps of the last stack (BN = 32): direct, predicated stores and / or the fused arg-max
        const int x = x0 + m % p.tw, y = y0 + (m / p.tw) % p.th, n = n0 + m / (p.tw * p.th);
        uint32_t r[32];
        tmem_ld_32x32(t_row, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          v[c] = fmaf(__uint_as_float(r[c]), lds_f32(aff_base + 4u * (nt * BN + c)), lds_f32(aff_base + 1024u + 4u * (nt * BN + c)));
          if (p.relu1) v[c] = fmaxf(v[c], 0.0f);
        }
        if (p.out_f32 && n < p.B) {
          float* o = p.out_f32 + (((size_t)n * p.H + y) * p.W + x) * p.f32_ld + nt * BN;
#pragma unroll
          for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (p.amax_keys) {
          // heat-map arg-max without the heat-map (reference read-out README.md:404, df2d behind core.py:177-185):
          // per channel the warp's maximum (redux) and its lowest lane = lowest flat index inside the tile; lane c
          // keeps channel c's key, the four warps meet in shared memory, one atomicMax per (tile, channel)
          const uint32_t myflat = (uint32_t)(y * p.W + x);
          unsigned long long key = 0ull;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            if (c < p.amax_k) {  // warp-uniform
              const uint32_t b = __float_as_uint(v[c] + 0.0f);  // -0 -> +0: equal values, first index wins
              const uint32_t o = b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);  // order-preserving
              const uint32_t mx = __reduce_max_sync(0xffffffffu, o);
              const uint32_t first = (uint32_t)__ffs(__ballot_sync(0xffffffffu, o == mx)) - 1u;
              const uint32_t fl = __shfl_sync(0xffffffffu, myflat, (int)first);
              if (lane == c) key = ((unsigned long long)mx << 32) | (unsigned long long)(0xffffffffu - fl);
            }
          }
          // scratch: the scale2 / shift2 quarter of the constant area (unused on this path): [group][warp][32] keys
          const uint32_t scratch = aff_base + 2048u + (uint32_t)(grp * 4 + q) * 256u + (uint32_t)lane * 8u;
          asm volatile("st.shared.u64 [%0], %1;" ::"r"(scratch), "l"(key) : "memory");
          named_bar_sync(bar_a, kEpiThreads);
          if (q == 0 && lane < p.amax_k) {
            unsigned long long best = key;
#pragma unroll
            for (int w = 1; w < 4; ++w) {
              unsigned long long k2;
              asm volatile("ld.shared.u64 %0, [%1];" : "=l"(k2) : "r"(scratch + (uint32_t)w * 256u));
              best = k2 > best ? k2 : best;
            }
            atomicMax(p.amax_keys + (size_t)n0 * BN + lane, best);
          }
          named_bar_sync(bar_b, kEpiThreads);  // the scratch may be rewritten by this group's next tile
        }
      } else {
#pragma unroll 1
        for (int sl = 0; sl < kSlabs; ++sl) {
          if (has_res) mbar_wait_warp(rfull_bar(rslot), rphase);
          const uint32_t rbuf = res_base + rslot * res_slot_bytes + row_off;
          const uint32_t rbuf2 = res_base + rslot * res_slot_bytes + kSlabBytes + row2_off;
          const uint32_t raw_buf = raw_base + obuf * kSlabBytes + row_off;
          const uint32_t act_buf = act_base + obuf * kSlabBytes + row_off;
          // the group's previous bulk store must have read the staging buffer out (a tile ago for one-slab tiles)
          if (leader) tma_store_wait_read<0>();
          named_bar_sync(bar_a, kEpiThreads);
          if (fast) {
            EpiRow row[2];
#pragma unroll
            for (int w = 0; w < 2; ++w) {
              row[w].rrow_s = res_base + rslot * res_slot_bytes + row_off2[w];
              row[w].rrow2_s = 0;
              row[w].sw = sw_2[w];
              row[w].sw2 = 0;
              row[w].out = nullptr;
              row[w].store = false;
              row[w].raw_s = raw_base + obuf * kSlabBytes + row_off2[w];
              row[w].act_s = act_base + obuf * kSlabBytes + row_off2[w];
            }
            const int c0 = nt * BN + sl * 64 + (int)hh * 32;  // this lane's first channel
            const float4* const c1 = reinterpret_cast<const float4*>(smg + (aff_base - smem_base)) + (c0 >> 2);
            const float4 *h1 = c1 + 64, *c2 = c1 + 128, *h2 = c1 + 192;  // scale1 | shift1 | scale2 | shift2, 256 floats (64 float4) each
            const uint32_t t_slab = t_row + sl * 64;
            //                      UNIT   RES    RES2   RELU   XSRC OUT   STAGED
            if (has_res) {
              if (has_act) epi_slab<false, true, false, false, 2, true, true>(t_slab, c1, h1, c2, h2, row, 0u, hh);
              else         epi_slab<false, true, false, false, 0, true, true>(t_slab, c1, h1, c2, h2, row, 0u, hh);
            } else if (p.relu1) {
              if (has_act) epi_slab<false, false, false, true, 2, true, true>(t_slab, c1, h1, c2, h2, row, 0u, hh);
              else         epi_slab<false, false, false, true, 0, true, true>(t_slab, c1, h1, c2, h2, row, 0u, hh);
            } else {
              if (has_act && !pool2) epi_slab<false, false, false, false, 2, true, true>(t_slab, c1, h1, c2, h2, row, 0u, hh);
              else                   epi_slab<false, false, false, false, 0, true, true>(t_slab, c1, h1, c2, h2, row, 0u, hh);
            }
            if (pool2) {
              // 2x2 max-pool of this warp's quarter of the staged raw tile (32 consecutive pixels = whole pooling windows:
              // two rows of a 16-wide tile or four rows of an 8-wide one -> 8 pooled pixels x 8 sixteen-byte chunks, two
              // items per lane), then the next BatchNorm + ReLU on the pooled value, like maxpool_bn_relu_kernel
              __syncwarp();
              const uint32_t raw_tile = raw_base + obuf * kSlabBytes, pool_tile = act_base + obuf * kSlabBytes;
#pragma unroll
              for (int it2 = 0; it2 < 2; ++it2) {
                const uint32_t item = (uint32_t)lane + 32u * it2;
                const uint32_t ppx = item >> 3, chunk = item & 7u;
                uint32_t m0;
                if (p.tw == 16) m0 = 32u * q + 2u * ppx;
                else            m0 = 32u * q + (ppx >> 2) * 16u + 2u * (ppx & 3u);
                const uint32_t dn = (uint32_t)p.tw;  // pixel below
                auto ldc = [&](uint32_t mm) { return lds128(raw_tile + mm * 128u + ((chunk ^ (mm & 7u)) << 4)); };
                const uint4 a = ldc(m0), b = ldc(m0 + 1u), c = ldc(m0 + dn), d = ldc(m0 + dn + 1u);
                auto mx2 = [](uint32_t x, uint32_t y) {
                  uint32_t r;
                  asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
                  return r;
                };
                uint4 mx;
                mx.x = mx2(mx2(a.x, b.x), mx2(c.x, d.x));
                mx.y = mx2(mx2(a.y, b.y), mx2(c.y, d.y));
                mx.z = mx2(mx2(a.z, b.z), mx2(c.z, d.z));
                mx.w = mx2(mx2(a.w, b.w), mx2(c.w, d.w));
                const uint32_t r8 = 8u * q + ppx;  // row of the pooled tile (box tw/2 x th/2)
                const uint32_t off = r8 * 128u + ((chunk ^ (r8 & 7u)) << 4);
                sts128(pool_tile + off, mx);
                const int cc = nt * BN + sl * 64 + (int)chunk * 8;
                const uint4 sa = lds128(aff_base + 2048u + 4u * cc), sb = lds128(aff_base + 2048u + 4u * cc + 16u);
                const uint4 ha = lds128(aff_base + 3072u + 4u * cc), hb = lds128(aff_base + 3072u + 4u * cc + 16u);
                const uint32_t s2[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                const uint32_t h2v[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
                const uint32_t mw[4] = {mx.x, mx.y, mx.z, mx.w};
                uint32_t ow[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float lo = fmaf(bf_lo(mw[e]), __uint_as_float(s2[2 * e]), __uint_as_float(h2v[2 * e]));
                  const float hi = fmaf(bf_hi(mw[e]), __uint_as_float(s2[2 * e + 1]), __uint_as_float(h2v[2 * e + 1]));
                  ow[e] = pack2_relu(lo, hi);
                }
                sts128(pool_tile + 4096u + off, make_uint4(ow[0], ow[1], ow[2], ow[3]));
              }
            }
          } else
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            tmem_ld_32x32(t_row + sl * 64 + half * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // 8 channels = one 16-byte chunk of the swizzled row
              const int c = nt * BN + sl * 64 + half * 32 + j * 8;
              const uint32_t chunk = ((uint32_t)(half * 4 + j) ^ sw) << 4;
              float v[8];
              {
                const uint4 sa = lds128(aff_base + 4u * c), sb = lds128(aff_base + 4u * c + 16u);
                const uint4 ha = lds128(aff_base + 1024u + 4u * c), hb = lds128(aff_base + 1024u + 4u * c + 16u);
                const uint32_t s1[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                const uint32_t h1[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  v[e] = fmaf(__uint_as_float(r[j * 8 + e]), __uint_as_float(s1[e]), __uint_as_float(h1[e]));
              }
              if (has_res) {
                const uint4 rr = lds128(rbuf + chunk);
                const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(rh[e]);
                  v[2 * e] += f.x;
                  v[2 * e + 1] += f.y;
                }
              }
              if (has_res2) {  // up1 + nearest_x2(low3): the sum is rounded to bf16 first, like a stored up1
                const uint4 rr = lds128(rbuf2 + (((uint32_t)(half * 4 + j) ^ sw2) << 4));
                const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(rh[e]);
                  v[2 * e] = bf16_round(v[2 * e]) + f.x;
                  v[2 * e + 1] = bf16_round(v[2 * e + 1]) + f.y;
                }
              }
              if (p.relu1) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.0f);
              }
              if (has_raw) {
                uint4 o;
                o.x = pack_bf16x2(v[0], v[1]);
                o.y = pack_bf16x2(v[2], v[3]);
                o.z = pack_bf16x2(v[4], v[5]);
                o.w = pack_bf16x2(v[6], v[7]);
                sts128(raw_buf + chunk, o);
              }
              if (has_act) {
                float w[8];
                const uint4 sa = lds128(aff_base + 2048u + 4u * c), sb = lds128(aff_base + 2048u + 4u * c + 16u);
                const uint4 ha = lds128(aff_base + 3072u + 4u * c), hb = lds128(aff_base + 3072u + 4u * c + 16u);
                const uint32_t s2[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                const uint32_t h2[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  w[e] = fmaxf(fmaf(bf16_round(v[e]), __uint_as_float(s2[e]), __uint_as_float(h2[e])), 0.0f);
                uint4 o;
                o.x = pack_bf16x2(w[0], w[1]);
                o.y = pack_bf16x2(w[2], w[3]);
                o.z = pack_bf16x2(w[4], w[5]);
                o.w = pack_bf16x2(w[6], w[7]);
                sts128(act_buf + chunk, o);
              }
            }
          }
          if (has_res) {  // residual slab consumed: hand the slot back to its producer
            __syncwarp();
            if (lane == 0) mbar_arrive(rempty_bar(rslot));
            if (++rslot == (uint32_t)n_res) {
              rslot = 0;
              rphase ^= 1u;
            }
          }
          fence_proxy_async();                 // generic-proxy smem writes -> visible to the TMA store
          named_bar_sync(bar_b, kEpiThreads);
          if (leader) {
            const int c0 = nt * BN + sl * 64;
            if (pool2) {
              tma_store_4d(&p.tmPoolRaw, act_base + obuf * kSlabBytes, c0, x0 >> 1, y0 >> 1, n0);
              tma_store_4d(&p.tmPoolAct, act_base + obuf * kSlabBytes + 4096u, c0, x0 >> 1, y0 >> 1, n0);
            } else {
              if (has_raw) tma_store_4d(&p.tmRaw, raw_base + obuf * kSlabBytes, c0, x0, y0, n0);
              if (has_act) tma_store_4d(&p.tmAct, act_base + obuf * kSlabBytes, c0, x0, y0, n0);
            }
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
    }
    if (leader) tma_store_wait_read<0>();      // shared memory must outlive the last bulk store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

int tma_init() {
  if (g_encode) return DF3D_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  DF3D_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  DF3D_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, DF3D_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return DF3D_OK;
}

int make_tmap_act(CUtensorMap* out, const void* base, int C, int W, int H, int N, int tw, int th, int nb) {
  if (int e = tma_init()) return e;
  DF3D_REQUIRE(C % 64 == 0, DF3D_EINVAL, "make_tmap_act: channels must be a multiple of 64 (got %d)", C);
  DF3D_REQUIRE(tw * th * nb == kTileM && tw <= 256 && th <= 256 && nb <= 256, DF3D_EINVAL, "make_tmap_act: bad tile %dx%dx%d", tw, th, nb);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)nb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DF3D_REQUIRE(r == CUDA_SUCCESS, DF3D_ECUDA, "cuTensorMapEncodeTiled(act C=%d W=%d H=%d N=%d) failed: %d", C, W, H, N, (int)r);
  return DF3D_OK;
}

int make_tmap_box(CUtensorMap* out, const void* base, int C, int W, int H, int N, int bw, int bh, int bn) {
  if (int e = tma_init()) return e;
  DF3D_REQUIRE(C % 64 == 0 && bw >= 1 && bh >= 1 && bn >= 1 && bw <= 256 && bh <= 256 && bn <= 256, DF3D_EINVAL,
               "make_tmap_box: bad box %dx%dx%d", bw, bh, bn);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DF3D_REQUIRE(r == CUDA_SUCCESS, DF3D_ECUDA, "cuTensorMapEncodeTiled(box C=%d W=%d H=%d N=%d) failed: %d", C, W, H, N, (int)r);
  return DF3D_OK;
}

int make_tmap_wgt(CUtensorMap* out, const void* base, int K, int CoutPad, int BN) {
  if (int e = tma_init()) return e;
  DF3D_REQUIRE(K % 64 == 0 && CoutPad % BN == 0 && BN <= 256, DF3D_EINVAL, "make_tmap_wgt: bad shape K=%d Cout=%d BN=%d", K, CoutPad, BN);
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)CoutPad};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DF3D_REQUIRE(r == CUDA_SUCCESS, DF3D_ECUDA, "cuTensorMapEncodeTiled(wgt K=%d Cout=%d) failed: %d", K, CoutPad, (int)r);
  return DF3D_OK;
}

int conv_gemm_configure() {
  DF3D_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  DF3D_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  DF3D_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  DF3D_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  return DF3D_OK;
}

int launch_conv_gemm(const ConvParams& p_in, int BN, int num_sms, cudaStream_t stream) {
  ConvParams p = p_in;
  const int total = p.tiles_x * p.tiles_y * p.tiles_b * p.n_tiles_n;
  if (total <= 0) return DF3D_OK;
  DF3D_REQUIRE(p.n_tiles_n * BN <= 256, DF3D_EUNSUPPORTED, "launch_conv_gemm: more than 256 output channels");
  DF3D_REQUIRE(!(p.out_f32 && (p.out_raw || p.out_act || p.residual)), DF3D_EUNSUPPORTED,
               "launch_conv_gemm: the fp32 output path takes no residual / bf16 outputs");
  DF3D_REQUIRE(p.out_f32 || p.amax_keys || BN >= 64, DF3D_EUNSUPPORTED, "launch_conv_gemm: bf16 outputs need BN >= 64");
  DF3D_REQUIRE(!p.amax_keys || (BN == 32 && p.n_tiles_n == 1 && p.nb == 1 && p.amax_k >= 1 && p.amax_k <= 32 &&
                                !(p.out_raw || p.out_act || p.residual)),
               DF3D_EUNSUPPORTED, "launch_conv_gemm: the fused arg-max needs the fp32 path (BN = 32), one image per tile");
  // shared-memory budget -> ring depths
  const int stage_bytes = kABytes + BN * 128;
  DF3D_REQUIRE(!p.pool2 || (p.out_raw && !p.out_act && !p.residual && !p.relu1 && !p.out_f32 && p.scale2 && p.nb == 1 &&
                            (p.tw == 16 || p.tw == 8) && p.th % 2 == 0 && BN >= 64),
               DF3D_EUNSUPPORTED, "launch_conv_gemm: the pooled epilogue needs a plain bf16 output, one image per tile, 16- or 8-wide tiles");
  DF3D_REQUIRE(p.kb_split == 0 || (p.taps == 1 && p.kb_split < p.kc_per_tap), DF3D_EUNSUPPORTED,
               "launch_conv_gemm: K-concatenation needs a 1x1 conv and a split inside its K blocks");
  const int fixed = 1024 + kAffBytes + kBarBytes + (p.out_raw ? 2 * kSlabBytes : 0) + ((p.out_act || p.pool2) ? 2 * kSlabBytes : 0);
  DF3D_REQUIRE(!p.has_res2 || (p.residual && p.tw % 2 == 0 && p.th % 2 == 0), DF3D_EUNSUPPORTED,
               "launch_conv_gemm: the half-resolution residual needs a full-resolution residual and an even tile");
  const int res_slot = kSlabBytes + (p.has_res2 ? kHalfSlabBytes : 0);
  int n_res = p.residual ? kMaxResSlots : 0;
  int n_stages = 0;
  if (p.halo9) {
    DF3D_REQUIRE(p.taps == 9 && p.kc_per_tap == 1 && BN == 64 && p.n_tiles_n == 1 && p.tw == 8 && p.th == 16 && p.nb == 1 &&
                     p.kb_split == 0,
                 DF3D_EUNSUPPORTED, "launch_conv_gemm: halo mode needs a 3x3 conv with 64 input and output channels on 8 x 16 tiles");
    for (;; --n_res) {
      n_stages = (kSmemLimit - fixed - n_res * res_slot - 9 * BN * 128) / kHalo9Slot;
      if (n_stages >= 2 || n_res <= (p.residual ? 1 : 0)) break;
    }
    if (n_stages > 4) n_stages = 4;
  } else
  for (;; --n_res) {
    n_stages = (kSmemLimit - fixed - n_res * res_slot) / stage_bytes;
    if (n_stages >= 2 || n_res <= (p.residual ? 1 : 0)) break;
  }
  if (n_stages > kMaxStages) n_stages = kMaxStages;
  if (const char* env = getenv("DF3D_CONV_STAGES")) {  // profiling knob: cap the ring depth
    const int v = atoi(env);
    if (v >= 2 && v < n_stages) n_stages = v;
  }
  DF3D_REQUIRE(n_stages >= 2, DF3D_EUNSUPPORTED, "launch_conv_gemm: shared-memory budget too small for BN=%d", BN);
  p.n_stages = n_stages;
  p.n_res_slots = n_res;
  const int smem = (p.halo9 ? 9 * BN * 128 + n_stages * kHalo9Slot : n_stages * stage_bytes) + n_res * res_slot + fixed;
  const int grid = total < num_sms ? total : num_sms;
  switch (BN) {
    case 32: conv_gemm_kernel<32><<<grid, kConvThreads, smem, stream>>>(p); break;
    case 64: conv_gemm_kernel<64><<<grid, kConvThreads, smem, stream>>>(p); break;
    case 128: conv_gemm_kernel<128><<<grid, kConvThreads, smem, stream>>>(p); break;
    case 256: conv_gemm_kernel<256><<<grid, kConvThreads, smem, stream>>>(p); break;
    default: DF3D_REQUIRE(false, DF3D_EUNSUPPORTED, "launch_conv_gemm: BN=%d not instantiated", BN);
  }
  DF3D_LAUNCH_CHECK("conv_gemm_kernel");
  return DF3D_OK;
}

}  // namespace df3d
