// tcgen05 convolution-chain kernel (see conv_chain.cuh).
//
// One persistent CTA per SM, 12 warps, 128-pixel tiles:
//   warp 0  TMA producer of the head ring: A tiles (one 4-D box per tap and 64-channel block; the
//           3x3 halo is the TMA's out-of-bounds zero fill) and the head's weight half-tiles
//   warp 1  MMA issuer (one thread) + TMEM owner
//   warp 2  TMA producer of the residual slabs (64 channels x 128 pixels, + the half-resolution slab
//           of the up-sample branch), a ring that prefetches across stages and tiles
//   warp 3  TMA producer of the weight ring of stages >= 1 (prefetched while the head GEMM runs)
//   warps 4..11  epilogue: two groups of four warps (one TMEM lane quarter each), group g takes the
//           64-channel slabs sl = g, g+2 of every stage
//
// Tensor memory (512 columns):  P = [0,128)  Q = [128,384)  accumulators,  X = [384,512) the bf16
// operand of the next stage (128 lanes x up to 256 channels, two per column).  Stage i:
//   MMA   : D(acc_i) = A_i * W_i^T, A_0 from shared memory (TMA), A_i (i >= 1) from X
//   epilogue: tcgen05.ld acc_i -> scale/shift (+ residuals) -> ReLU / bf16 rounding
//             -> optional bf16 store straight from registers (each thread owns one pixel row: 64
//                contiguous bytes per 32 channels = two full-sector 256-bit stores)
//             -> optional next-BatchNorm + ReLU -> tcgen05.st into X
// The chain of one tile is sequential (stage i+1 needs the whole operand of stage i), but the head
// GEMM of the next tile is issued right behind the last stage and overlaps its epilogue, and both
// rings keep prefetching across tiles.
#include <cstdlib>

#include "conv_chain.cuh"
#include "sm100.cuh"

namespace df3d {

using namespace sm100;

constexpr int kChainThreads = 384;
constexpr int kEpiWarp0c = 4;
constexpr int kUnitBytes = 16384;       // 128 rows x 64 bf16: one A tile or one 128-row weight half-tile
constexpr int kMaxM = 8, kMaxW = 4, kMaxSlabs = 8;
constexpr int kColP = 0, kColQ = 128, kColX = 384;
constexpr int kChainSmemLimit = 232448;  // 227 KB opt-in maximum per CTA
constexpr int kChainBarBytes = 512;
// specialised epilogues (see epi_slab)
enum { kEpiReluX = 0, kEpiReluOut, kEpiResOutAct, kEpiResUpOutAct, kEpiResOut, kEpiResUpOut, kEpiResX };

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 32 bytes (one full sector) to global memory
__device__ __forceinline__ void stg256(void* ptr, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// relu + round-to-nearest-even + pack in one instruction: {hi, lo} -> bf16x2 (lo in the low half)
__device__ __forceinline__ uint32_t pack2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ float bf_lo(uint32_t x) { return __uint_as_float(x << 16); }
__device__ __forceinline__ float bf_hi(uint32_t x) { return __uint_as_float(x & 0xffff0000u); }

// Epilogue of one 64-channel slab of one stage for one pixel row (one thread), fully specialised:
//   UNIT  scale1 == 1 (conv without a folded BatchNorm): v = acc + shift1
//   RES   + residual (bf16, swizzled slab row)     RES2  + nearest-x2 up-sampled half-resolution residual
//   RELU  relu after the adds                      XSRC  0: no operand, 1: bf16(v), 2: relu(bn2(bf16(v)))
//   OUT   bf16(v) to global memory
// r: the 64 fp32 accumulator columns; c1/c2: this slab's first channel in the constant arrays.
template <bool UNIT, bool RES, bool RES2, bool RELU, int XSRC, bool OUT>
__device__ __forceinline__ void epi_slab(const uint32_t (&r)[2][32], const float4* __restrict__ sc1,
                                         const float4* __restrict__ sh1, const float4* __restrict__ sc2,
                                         const float4* __restrict__ sh2, const uint8_t* __restrict__ rrow,
                                         const uint8_t* __restrict__ rrow2, uint32_t sw, uint32_t sw2, uint32_t x_addr,
                                         uint8_t* out, bool store) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t xp[16], op[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // 8 channels = one 16-byte chunk of the swizzled slab row
      const int c4 = (half * 32 + j * 8) >> 2;
      float v[8];
      {
        const float4 ha = sh1[c4], hb = sh1[c4 + 1];
        const float t1[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
        if (UNIT) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[half][j * 8 + e]) + t1[e];
        } else {
          const float4 sa = sc1[c4], sb = sc1[c4 + 1];
          const float s1[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = fmaf(__uint_as_float(r[half][j * 8 + e]), s1[e], t1[e]);
        }
      }
      if (RES) {
        const uint4 rr = *reinterpret_cast<const uint4*>(rrow + (((uint32_t)(half * 4 + j) ^ sw) << 4));
        const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[2 * e] += bf_lo(rw[e]);
          v[2 * e + 1] += bf_hi(rw[e]);
        }
      }
      if (RES2) {  // up1 + nearest_x2(low3): the sum is rounded to bf16 first, like a stored up1
        const uint4 rr = *reinterpret_cast<const uint4*>(rrow2 + (((uint32_t)(half * 4 + j) ^ sw2) << 4));
        const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t pk = pack2(v[2 * e], v[2 * e + 1]);
          v[2 * e] = bf_lo(pk) + bf_lo(rw[e]);
          v[2 * e + 1] = bf_hi(pk) + bf_hi(rw[e]);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) op[j * 4 + e] = RELU ? pack2_relu(v[2 * e], v[2 * e + 1]) : pack2(v[2 * e], v[2 * e + 1]);
      if (XSRC == 2) {  // act = relu(bn(bf16(v))): the rounded value is the packed one
        const float4 sa = sc2[c4], sb = sc2[c4 + 1], ha = sh2[c4], hb = sh2[c4 + 1];
        const float s2[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
        const float t2[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t ow = op[j * 4 + e];
          xp[j * 4 + e] = pack2_relu(fmaf(bf_lo(ow), s2[2 * e], t2[2 * e]), fmaf(bf_hi(ow), s2[2 * e + 1], t2[2 * e + 1]));
        }
      }
    }
    if (XSRC == 1) tmem_st_32x16(x_addr + half * 16, op);
    if (XSRC == 2) tmem_st_32x16(x_addr + half * 16, xp);
    if (OUT) {
      if (store) {  // 32 channels = 64 contiguous bytes of this thread's pixel: two full 32-byte sectors
        stg256(out + half * 64, op);
        stg256(out + half * 64 + 32, op + 8);
      }
    }
  }
}

__global__ void __launch_bounds__(kChainThreads, 1) conv_chain_kernel(const __grid_constant__ ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const sm = smem_raw + (smem_base - smem_u32(smem_raw));  // same place as a __shared__ pointer: plain
                                                                    // loads/stores the compiler may schedule
  const uint32_t m_base = smem_base;
  const uint32_t w_base = m_base + (uint32_t)p.n_m * kUnitBytes;
  const uint32_t s_base = w_base + (uint32_t)p.n_w * kUnitBytes;
  const uint32_t aff_base = s_base + (uint32_t)p.n_slabs * (uint32_t)p.slab_bytes;
  const uint32_t bar_base = aff_base + (uint32_t)p.aff_bytes;
  auto mfull = [&](uint32_t s) { return bar_base + 8u * s; };
  auto mempty = [&](uint32_t s) { return bar_base + 8u * (kMaxM + s); };
  auto wfull = [&](uint32_t s) { return bar_base + 8u * (2 * kMaxM + s); };
  auto wempty = [&](uint32_t s) { return bar_base + 8u * (2 * kMaxM + kMaxW + s); };
  auto sfull = [&](uint32_t s) { return bar_base + 8u * (2 * kMaxM + 2 * kMaxW + s); };
  auto sempty = [&](uint32_t s) { return bar_base + 8u * (2 * kMaxM + 2 * kMaxW + kMaxSlabs + s); };
  auto rfull = [&](uint32_t r) { return bar_base + 8u * (2 * kMaxM + 2 * kMaxW + 2 * kMaxSlabs + r); };
  auto rempty = [&](uint32_t r) { return bar_base + 8u * (2 * kMaxM + 2 * kMaxW + 2 * kMaxSlabs + 2 + r); };
  const uint32_t xfull = bar_base + 8u * (2 * kMaxM + 2 * kMaxW + 2 * kMaxSlabs + 4);
  const uint32_t tmem_slot = xfull + 8u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  const int n_chain = p.n_chain;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.tmA);
    for (int i = 0; i < n_chain; ++i) {
      prefetch_tensormap(&p.st[i].tmB);
      if (p.st[i].has_res) prefetch_tensormap(&p.st[i].tmRes);
      if (p.st[i].has_res2) prefetch_tensormap(&p.st[i].tmRes2);
    }
    for (int s = 0; s < p.n_m; ++s) {
      mbar_init(mfull(s), 1);
      mbar_init(mempty(s), 1);
    }
    for (int s = 0; s < p.n_w; ++s) {
      mbar_init(wfull(s), 1);
      mbar_init(wempty(s), 1);
    }
    for (int s = 0; s < p.n_slabs; ++s) {
      mbar_init(sfull(s), 1);
      mbar_init(sempty(s), 4);  // one arrive per warp of the epilogue group that read it
    }
    for (int r = 0; r < 2; ++r) {
      mbar_init(rfull(r), 1);
      mbar_init(rempty(r), 8);  // one arrive per epilogue warp
    }
    mbar_init(xfull, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  // per-channel epilogue constants of every stage -> shared memory: [scale1 n][shift1 n][scale2 n][shift2 n]
  for (int i = 0; i < n_chain; ++i) {
    const ChainStage& st = p.st[i];
    const uint32_t a0 = aff_base + 4u * (uint32_t)st.aff_off;
    for (int c = threadIdx.x; c < st.n; c += kChainThreads) {
      const float s2 = st.x_src == 2 ? st.scale2[c] : 0.f, h2 = st.x_src == 2 ? st.shift2[c] : 0.f;
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(a0 + 4u * c), "f"(st.scale1[c]));
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(a0 + 4u * (st.n + c)), "f"(st.shift1[c]));
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(a0 + 4u * (2 * st.n + c)), "f"(s2));
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(a0 + 4u * (3 * st.n + c)), "f"(h2));
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  auto decode_tile = [&](int tile, int& x0, int& y0, int& n0) {
    const int tx = tile % p.tiles_x;
    tile /= p.tiles_x;
    const int ty = tile % p.tiles_y;
    const int tb = tile / p.tiles_y;
    x0 = tx * p.tw;
    y0 = ty * p.th;
    n0 = tb * p.nb;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ head ring producer
    if (lane == 0) {
      const int kb0 = p.st[0].kblocks, nh0 = p.st[0].n >> 7;
      uint32_t u = 0, ph = 0;
      auto advance = [&]() {
        if (++u == (uint32_t)p.n_m) {
          u = 0;
          ph ^= 1u;
        }
      };
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int x0, y0, n0;
        decode_tile(tile, x0, y0, n0);
        for (int kb = 0; kb < kb0; ++kb) {
          const int tap = kb / p.kc_per_tap, kc = kb - tap * p.kc_per_tap;
          int dx = 0, dy = 0;
          if (p.taps == 9) {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
          }
          mbar_wait(mempty(u), ph ^ 1u);
          mbar_arrive_expect_tx(mfull(u), kUnitBytes);
          tma_load_4d(m_base + u * kUnitBytes, &p.tmA, mfull(u), kc * 64, x0 + dx, y0 + dy, n0);
          advance();
          for (int h = 0; h < nh0; ++h) {
            mbar_wait(mempty(u), ph ^ 1u);
            mbar_arrive_expect_tx(mfull(u), kUnitBytes);
            tma_load_2d(m_base + u * kUnitBytes, &p.st[0].tmB, mfull(u), kb * 64, h * 128);
            advance();
          }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ weight ring producer (stages >= 1)
    if (lane == 0 && n_chain > 1) {
      uint32_t u = 0, ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        for (int i = 1; i < n_chain; ++i) {
          const int kbn = p.st[i].kblocks, nh = p.st[i].n >> 7;
          for (int kb = 0; kb < kbn; ++kb)
            for (int h = 0; h < nh; ++h) {
              mbar_wait(wempty(u), ph ^ 1u);
              mbar_arrive_expect_tx(wfull(u), kUnitBytes);
              tma_load_2d(w_base + u * kUnitBytes, &p.st[i].tmB, wfull(u), kb * 64, h * 128);
              if (++u == (uint32_t)p.n_w) {
                u = 0;
                ph ^= 1u;
              }
            }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ residual / staging slab producer
    if (lane == 0 && p.n_slabs > 0) {
      uint32_t u = 0, ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int x0, y0, n0;
        decode_tile(tile, x0, y0, n0);
        for (int i = 0; i < n_chain; ++i) {
          const ChainStage& st = p.st[i];
          if (!st.has_res) continue;
          const int nsl = st.n >> 6;
          for (int sl = 0; sl < nsl; ++sl) {
            mbar_wait(sempty(u), ph ^ 1u);
            const uint32_t slab = s_base + u * (uint32_t)p.slab_bytes;
            mbar_arrive_expect_tx(sfull(u), kUnitBytes + (st.has_res2 ? kUnitBytes / 4 : 0));
            tma_load_4d(slab, &st.tmRes, sfull(u), sl * 64, x0, y0, n0);
            if (st.has_res2) tma_load_4d(slab + kUnitBytes, &st.tmRes2, sfull(u), sl * 64, x0 >> 1, y0 >> 1, n0);
            if (++u == (uint32_t)p.n_slabs) {
              u = 0;
              ph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // One thread; the loop bodies are kept minimal (no local arrays, no unrolling across blocks):
    // the issue rate of this thread bounds the tensor pipe.
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint64_t desc_hi = umma_smem_desc_sw128(0);  // everything but the 14-bit start address
      const uint32_t n_m = (uint32_t)p.n_m, n_w = (uint32_t)p.n_w;
      const int kb0 = p.st[0].kblocks, nh0 = p.st[0].n >> 7;
      const uint32_t d0 = tmem_base + (uint32_t)p.st[0].acc_col;
      const uint32_t reg0 = p.st[0].acc_col == kColP ? 0u : 1u;
      uint32_t mu = 0, mph = 0, wu = 0, wph = 0, xuse = 0, use0 = 0, use1 = 0;
      unsigned long long* const dbg = blockIdx.x == 0 ? p.dbg : nullptr;
      int di = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        // ---- head: A and B from shared memory
        {
          uint32_t& use = reg0 ? use1 : use0;
          if (dbg && di < 4000) dbg[di++] = clock64();  // [tile start]
          mbar_wait(rempty(reg0), (use & 1u) ^ 1u);  // the epilogue has drained this accumulator
          ++use;
          tc_fence_after();
          if (dbg && di < 4000) dbg[di++] = clock64();  // [head accumulator free]
#pragma unroll 1
          for (int kb = 0; kb < kb0; ++kb) {
            mbar_wait(mfull(mu), mph);
            const uint32_t ua = mu;
            const uint64_t adesc = desc_hi | (uint64_t)(((m_base + mu * kUnitBytes) >> 4) & 0x3FFFu);
            if (++mu == n_m) {
              mu = 0;
              mph ^= 1u;
            }
#pragma unroll 1
            for (int h = 0; h < nh0; ++h) {
              mbar_wait(mfull(mu), mph);
              tc_fence_after();
              const uint64_t bdesc = desc_hi | (uint64_t)(((m_base + mu * kUnitBytes) >> 4) & 0x3FFFu);
              const uint32_t d = d0 + h * 128;
#pragma unroll
              for (int k = 0; k < 4; ++k)  // 4 x (K = 16): +32 bytes inside the 128B swizzle atom
                umma_bf16(d, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
              umma_commit(mempty(mu));  // frees the weight half-tile once these MMAs retire
              if (++mu == n_m) {
                mu = 0;
                mph ^= 1u;
              }
            }
            umma_commit(mempty(ua));
          }
          umma_commit(rfull(reg0));
          if (dbg && di < 4000) dbg[di++] = clock64();  // [head issued]
        }
        // ---- later stages: A from tensor memory (X), B from the weight ring
#pragma unroll 1
        for (int i = 1; i < n_chain; ++i) {
          const int kbn = p.st[i].kblocks, nh = p.st[i].n >> 7, col = p.st[i].acc_col;
          const uint32_t reg = col == kColP ? 0u : 1u;
          uint32_t& use = reg ? use1 : use0;
          mbar_wait(rempty(reg), (use & 1u) ^ 1u);
          ++use;
          mbar_wait(xfull, xuse & 1u);  // operand of this stage is complete in tensor memory
          ++xuse;
          tc_fence_after();
          if (dbg && di < 4000) dbg[di++] = clock64();  // [stage i operand ready]
#pragma unroll 1
          for (int kb = 0; kb < kbn; ++kb) {
            const uint32_t xa = tmem_base + kColX + kb * 32;
#pragma unroll 1
            for (int h = 0; h < nh; ++h) {
              mbar_wait(wfull(wu), wph);
              tc_fence_after();
              const uint64_t bdesc = desc_hi | (uint64_t)(((w_base + wu * kUnitBytes) >> 4) & 0x3FFFu);
              const uint32_t d = tmem_base + (uint32_t)col + h * 128;
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16_ts(d, xa + k * 8, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
              umma_commit(wempty(wu));
              if (++wu == n_w) {
                wu = 0;
                wph ^= 1u;
              }
            }
          }
          umma_commit(rfull(reg));  // accumulator ready; the operand in X may be overwritten
          if (dbg && di < 4000) dbg[di++] = clock64();  // [stage i issued]
        }
      }
    }
  } else if (warp >= kEpiWarp0c) {
    // ------------------------------------------------------------------ epilogue (warps 4..11)
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int grp = (warp - kEpiWarp0c) >> 2; // slab parity this group handles
    const int m = q * 32 + lane;              // row of the tile = pixel
    const uint32_t row_off = (uint32_t)m * 128u;
    const uint32_t sw = (uint32_t)(m & 7);
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    // position of this thread's pixel inside the tile, and the row of its parent in the
    // half-resolution residual slab (box tw/2 x th/2 x nb)
    const int pw = m % p.tw, phh = (m / p.tw) % p.th, pn = m / (p.tw * p.th);
    uint32_t row2_off, sw2;
    {
      const int r2 = (pn * (p.th >> 1) + (phh >> 1)) * (p.tw >> 1) + (pw >> 1);
      row2_off = (uint32_t)r2 * 128u;
      sw2 = (uint32_t)(r2 & 7);
    }
    uint32_t use0 = 0, use1 = 0;
    uint32_t scount = 0;
    unsigned long long* const dbg = (blockIdx.x == 0 && lane == 0 && q == 0) ? p.dbg : nullptr;
    int di = 4096 * (1 + grp);
    const int dend = di + 4000;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int x0, y0, n0;
      decode_tile(tile, x0, y0, n0);
      const bool in_batch = (n0 + pn) < p.B;  // partially filled multi-image tiles: skip the stores
      const size_t pixel = ((size_t)(n0 + pn) * p.H + (y0 + phh)) * p.W + (x0 + pw);
      for (int i = 0; i < n_chain; ++i) {
        const ChainStage& st = p.st[i];
        const uint32_t reg = st.acc_col == kColP ? 0u : 1u;
        const bool has_res = st.has_res != 0;
        const int x_src = st.x_src, kind = st.epi_kind;
        const int nsl = st.n >> 6;
        uint8_t* const out_row = st.out_raw ? reinterpret_cast<uint8_t*>(st.out_raw) + pixel * (size_t)st.n * 2 : nullptr;
        const float4* const sc1 = reinterpret_cast<const float4*>(sm + (aff_base - smem_base)) + (st.aff_off >> 2);
        const float4* const sh1 = sc1 + (st.n >> 2);
        const float4* const sc2 = sh1 + (st.n >> 2);
        const float4* const sh2 = sc2 + (st.n >> 2);
        {
          uint32_t& use = reg ? use1 : use0;
          mbar_wait(rfull(reg), use & 1u);
          ++use;
        }
        tc_fence_after();
        if (dbg && di < dend) dbg[di++] = clock64();  // [stage i accumulator ready]
        const uint32_t t_row = tmem_base + lane_base + (uint32_t)st.acc_col;
        const uint32_t x_row = tmem_base + lane_base + kColX;
#pragma unroll 1
        for (int sl = grp; sl < nsl; sl += 2) {
          uint32_t su = 0, slab = s_base;
          uint32_t r[2][32];
          tmem_ld_32x32(t_row + sl * 64, r[0]);  // both halves in flight while the residual slab is awaited
          tmem_ld_32x32(t_row + sl * 64 + 32, r[1]);
          if (has_res) {
            const uint32_t idx = scount + (uint32_t)sl;
            su = idx % (uint32_t)p.n_slabs;
            mbar_wait(sfull(su), (idx / (uint32_t)p.n_slabs) & 1u);
            slab = s_base + su * (uint32_t)p.slab_bytes;
          }
          const uint8_t* const rrow = sm + (slab - smem_base) + row_off;
          const uint8_t* const rrow2 = sm + (slab - smem_base) + kUnitBytes + row2_off;
          const float4 *c1 = sc1 + sl * 16, *h1 = sh1 + sl * 16, *c2 = sc2 + sl * 16, *h2 = sh2 + sl * 16;
          const uint32_t xa = x_row + sl * 32;
          uint8_t* const o = out_row + sl * 128;
          tmem_ld_wait();
          switch (kind) {  //         UNIT   RES    RES2   RELU  XSRC OUT
            case kEpiReluX:    epi_slab<false, false, false, true, 1, false>(r, c1, h1, c2, h2, rrow, rrow2, sw, sw2, xa, o, in_batch); break;
            case kEpiReluOut:  epi_slab<false, false, false, true, 0, true>(r, c1, h1, c2, h2, rrow, rrow2, sw, sw2, xa, o, in_batch); break;
            case kEpiResOutAct: epi_slab<true, true, false, false, 2, true>(r, c1, h1, c2, h2, rrow, rrow2, sw, sw2, xa, o, in_batch); break;
            case kEpiResUpOutAct: epi_slab<true, true, true, false, 2, true>(r, c1, h1, c2, h2, rrow, rrow2, sw, sw2, xa, o, in_batch); break;
            case kEpiResOut:   epi_slab<true, true, false, false, 0, true>(r, c1, h1, c2, h2, rrow, rrow2, sw, sw2, xa, o, in_batch); break;
            case kEpiResUpOut: epi_slab<true, true, true, false, 0, true>(r, c1, h1, c2, h2, rrow, rrow2, sw, sw2, xa, o, in_batch); break;
            case kEpiResX:     epi_slab<true, true, false, false, 1, false>(r, c1, h1, c2, h2, rrow, rrow2, sw, sw2, xa, o, in_batch); break;
            default: break;  // launch_conv_chain rejects anything else
          }
          if (has_res) {  // slab consumed by this warp
            __syncwarp();
            if (lane == 0) mbar_arrive(sempty(su));
          }
        }
        if (has_res) scount += (uint32_t)nsl;
        if (x_src) tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (x_src) mbar_arrive(xfull);
          mbar_arrive(rempty(reg));
        }
        if (dbg && di < dend) dbg[di++] = clock64();  // [stage i epilogue done]
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------ host
int conv_chain_configure() {
  DF3D_CUDA(cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmemLimit));
  return DF3D_OK;
}

int launch_conv_chain(const ChainParams& p_in, int num_sms, cudaStream_t stream) {
  ChainParams p = p_in;
  const int total = p.tiles_x * p.tiles_y * p.tiles_b;
  if (total <= 0) return DF3D_OK;
  DF3D_REQUIRE(p.n_chain >= 1 && p.n_chain <= kMaxChain, DF3D_EINVAL, "launch_conv_chain: bad chain length %d", p.n_chain);
  DF3D_REQUIRE(p.tw * p.th * p.nb == 128, DF3D_EINVAL, "launch_conv_chain: tile must hold 128 pixels");
  int aff_floats = 0;
  bool any_slab = false, any_res2 = false;
  for (int i = 0; i < p.n_chain; ++i) {
    ChainStage& st = p.st[i];
    DF3D_REQUIRE(st.n == 128 || st.n == 256, DF3D_EUNSUPPORTED, "launch_conv_chain: stage %d has %d output channels (128 or 256)", i, st.n);
    DF3D_REQUIRE(i == 0 || st.kblocks * 64 == p.st[i - 1].n, DF3D_EINVAL, "launch_conv_chain: stage %d K does not match stage %d N", i, i - 1);
    DF3D_REQUIRE((i + 1 < p.n_chain) == (st.x_src != 0), DF3D_EINVAL, "launch_conv_chain: x_src must be set on every stage but the last");
    DF3D_REQUIRE(!st.has_res2 || (st.has_res && p.tw % 2 == 0 && p.th % 2 == 0), DF3D_EUNSUPPORTED,
                 "launch_conv_chain: the half-resolution residual needs a full-resolution residual and an even tile");
    DF3D_REQUIRE(st.x_src || st.out_raw, DF3D_EINVAL, "launch_conv_chain: the last stage must store its output");
    {
      const bool relu = st.relu1 != 0, res = st.has_res != 0, up = st.has_res2 != 0, out = st.out_raw != nullptr;
      const bool unit = st.unit_scale != 0;
      int kind = -1;
      if (!unit && !res && !up && relu && st.x_src == 1 && !out) kind = kEpiReluX;
      if (!unit && !res && !up && relu && st.x_src == 0 && out) kind = kEpiReluOut;
      if (unit && res && !up && !relu && st.x_src == 2 && out) kind = kEpiResOutAct;
      if (unit && res && up && !relu && st.x_src == 2 && out) kind = kEpiResUpOutAct;
      if (unit && res && !up && !relu && st.x_src == 0 && out) kind = kEpiResOut;
      if (unit && res && up && !relu && st.x_src == 0 && out) kind = kEpiResUpOut;
      if (unit && res && !up && !relu && st.x_src == 1 && !out) kind = kEpiResX;
      DF3D_REQUIRE(kind >= 0, DF3D_EUNSUPPORTED,
                   "launch_conv_chain: stage %d has no specialised epilogue (unit %d res %d up %d relu %d x %d out %d)", i,
                   (int)unit, (int)res, (int)up, (int)relu, st.x_src, (int)out);
      st.epi_kind = kind;
    }
    st.aff_off = aff_floats;
    aff_floats += 4 * st.n;
    any_slab |= st.has_res != 0;
    any_res2 |= st.has_res2 != 0;
  }
  // tensor-memory regions: head in P when it is 128 wide, else Q; 256-wide stages in Q; a 128-wide
  // later stage takes the region the head does not use so that the next tile's head GEMM can overlap
  // its epilogue
  const int head_col = p.st[0].n == 128 ? kColP : kColQ;
  p.st[0].acc_col = head_col;
  for (int i = 1; i < p.n_chain; ++i) p.st[i].acc_col = p.st[i].n == 256 ? kColQ : (head_col == kColP ? kColQ : kColP);
  // shared-memory budget
  p.aff_bytes = (aff_floats * 4 + 255) & ~255;
  p.slab_bytes = kUnitBytes + (any_res2 ? kUnitBytes / 4 : 0);
  p.n_slabs = any_slab ? (p.taps == 9 ? 4 : 6) : 0;
  const int fixed = 1024 + p.aff_bytes + kChainBarBytes + p.n_slabs * p.slab_bytes;
  const int units = (kChainSmemLimit - fixed) / kUnitBytes;
  p.n_w = p.n_chain > 1 ? (units >= 9 ? 3 : 2) : 0;
  if (const char* env = getenv("DF3D_CHAIN_NS")) {  // profiling knobs: slab / weight-ring / head-ring depth
    const int v = atoi(env);
    if (any_slab && v >= 2 && v <= kMaxSlabs) p.n_slabs = v;
  }
  if (const char* env = getenv("DF3D_CHAIN_NW")) {
    const int v = atoi(env);
    if (p.n_chain > 1 && v >= 2 && v <= kMaxW) p.n_w = v;
  }
  const int fixed2 = 1024 + p.aff_bytes + kChainBarBytes + p.n_slabs * p.slab_bytes;
  p.n_m = (kChainSmemLimit - fixed2) / kUnitBytes - p.n_w;
  if (p.n_m > kMaxM) p.n_m = kMaxM;
  if (const char* env = getenv("DF3D_CHAIN_NM")) {
    const int v = atoi(env);
    if (v >= 3 && v < p.n_m) p.n_m = v;
  }
  DF3D_REQUIRE(p.n_m >= 1 + (p.st[0].n >> 7) && p.n_w <= kMaxW, DF3D_EUNSUPPORTED, "launch_conv_chain: shared-memory budget too small");
  const int smem = fixed2 + (p.n_m + p.n_w) * kUnitBytes;
  const int grid = total < num_sms ? total : num_sms;
  conv_chain_kernel<<<grid, kChainThreads, smem, stream>>>(p);
  DF3D_LAUNCH_CHECK("conv_chain_kernel");
  return DF3D_OK;
}

}  // namespace df3d
