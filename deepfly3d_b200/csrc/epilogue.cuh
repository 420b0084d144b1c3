// Epilogue of one 64-channel slab of a tcgen05 accumulator, shared by the conv-chain kernel (conv_chain.cu) and
// the implicit-GEMM conv kernel (conv_gemm.cu): tcgen05.ld -> per-channel affine (+ residuals) -> ReLU / bf16
// rounding -> output (global / in place in the residual slab / shared-memory staging tile) and / or the next
// BatchNorm + ReLU (operand of the next GEMM in tensor memory, or a second staging tile).
// Replaces the torch batch_norm / relu / add calls around every conv of the hourglass inside df2d (reference
// call site df3d/core.py:177-185).
#pragma once
#include <cuda_bf16.h>

#include "sm100.cuh"

namespace df3d {

using namespace sm100;

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

// Two fp32 lanes per instruction (FADD2 / FFMA2): same IEEE results as the scalar forms, half the issue slots.
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 a, b, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tadd.rn.f32x2 d, a, b;\n\tmov.b64 {%0, %1}, d;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 a, b, c, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmov.b64 c, {%6, %7};\n\t"
      "fma.rn.f32x2 d, a, b, c;\n\tmov.b64 {%0, %1}, d;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 bf2_unpack(uint32_t x) { return make_float2(bf_lo(x), bf_hi(x)); }
// bf16x2 + bf16x2 with ONE rounding (HADD2.BF16): bit-identical to unpacking both, adding in fp32 and packing again
// -- the fp32 sum of two bf16 values is exact unless their exponents are more than 16 apart, and then neither
// rounding moves the larger operand (checked on 8 M random pairs, tools/check_bf16_add.py) -- in 1 instruction for 6
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

// Operands of one group of 8 channels of a slab, fetched from shared memory two groups ahead of their use
// (unused members are dead code in the specialisations that do not need them).
// Per-channel constants of one group of 8 channels (unused members are dead code in the specialisations
// that do not need them).
struct EpiGroup {
  float4 sc1[2], sh1[2], sc2[2], sh2[2];
  uint4 res[2], up[2];  // residual / half-resolution residual chunk of the two rows
};
// What the epilogue needs to know about one of the two pixel rows a lane works on
struct EpiRow {
  uint32_t rrow_s;       // residual slab row (shared-memory address), 128 bytes, 16-byte chunks swizzled with sw;
                         // also where the in-place output goes
  uint32_t rrow2_s;      // row of the parent pixel in the half-resolution slab, swizzled with sw2
  uint32_t sw, sw2;
  uint8_t* out;          // global address of this pixel's 64 channels of the slab (stages stored from registers)
  bool store;
  uint32_t raw_s, act_s; // STAGED mode (conv_gemm): this pixel's rows of the raw / activated output staging tiles
};

// Epilogue of one 64-channel slab of one stage, fully specialised:
//   UNIT  scale1 == 1 (conv without a folded BatchNorm): v = acc + shift1
//   RES   + residual (bf16, swizzled slab row)     RES2  + nearest-x2 up-sampled half-resolution residual
//   RELU  relu after the adds                      XSRC  0: no operand, 1: bf16(v), 2: relu(bn2(bf16(v)))
//   OUT   bf16(v) to global memory
// Lane mapping (tcgen05.ld/st shape .16x32bx2): lanes l and l + 16 of a warp share a pixel row and split
// its channels -- lane l owns channels [32 h, 32 h + 32) of the slab, h = l / 16, for the TWO rows
// r16 = l % 16 and r16 + 16 of the warp's 32-row quarter.  With the plain .32x32b shape (one row, all 64
// channels per lane) every lane needs every per-channel constant of the slab, and broadcast loads do not
// come cheaper: an LDS.128 of one address still costs four shared-memory wavefronts -- the constants were
// 4 500 of the 9 500 wavefronts per tile on the shared-memory port that bounds this kernel.  Two rows per
// lane halve that (each constant is fetched once and used for both rows); the residual / output chunks
// stay 16 bytes per access.
// t_slab: TMEM address (lane field = first lane of the quarter) of the slab's 64 fp32 accumulator columns;
// the operand of the next stage is written back in place over the first 32 of them (x_addr), which is
// safe because both rows' accumulators are in registers before the first tcgen05.st.
// c1/h1/c2/h2: the constant arrays at this lane's first channel.
//   STAGED  conv_gemm.cu: bf16(v) and / or the activated value go to 128B-swizzled shared-memory staging tiles
//           (TMA stores follow) instead of global memory / tensor memory
//   INPLACE the output goes into the slab row (rrow_s) although there is no residual in it (blank slab of the ring)
template <bool UNIT, bool RES, bool RES2, bool RELU, int XSRC, bool OUT, bool STAGED = false, bool INPLACE = false>
__device__ __forceinline__ void epi_slab(uint32_t t_slab, const float4* __restrict__ sc1,
                                         const float4* __restrict__ sh1, const float4* __restrict__ sc2,
                                         const float4* __restrict__ sh2, const EpiRow (&row)[2], uint32_t x_addr,
                                         uint32_t h) {
  // (sc1 .. sh2 are pointers derived from a __shared__ array: the compiler emits LDS for them directly; the
  // slab rows are 32-bit shared addresses -- a generic pointer rebuilt per slab cost an S2UR of the cluster
  // CTA id and a dependent address chain each time)
  auto fetch = [&](EpiGroup& g, int j) {  // j: group of 8 channels of this lane's half, 0..3
    g.sh1[0] = sh1[2 * j];
    g.sh1[1] = sh1[2 * j + 1];
    if (!UNIT) {
      g.sc1[0] = sc1[2 * j];
      g.sc1[1] = sc1[2 * j + 1];
    }
    if (XSRC == 2) {
      g.sc2[0] = sc2[2 * j];
      g.sc2[1] = sc2[2 * j + 1];
      g.sh2[0] = sh2[2 * j];
      g.sh2[1] = sh2[2 * j + 1];
    }
    const uint32_t ck = 4u * h + (uint32_t)j;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      if (RES) g.res[w] = lds128(row[w].rrow_s + ((ck ^ row[w].sw) << 4));
      if (RES2) g.up[w] = lds128(row[w].rrow2_s + ((ck ^ row[w].sw2) << 4));
    }
  };
  uint32_t acc[2][32];
  tmem_ld_16x32bx2_x32(t_slab, acc[0]);                 // rows r16:      this lane's 32 channels
  tmem_ld_16x32bx2_x32(t_slab + (16u << 16), acc[1]);   // rows r16 + 16
  EpiGroup g[2];
  uint32_t held[2][4];  // stages stored from registers: the even group's 16 bytes wait for the odd group's
  (void)held;
  fetch(g[0], 0);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j + 1 < 4) fetch(g[(j + 1) & 1], j + 1);
    const EpiGroup& G = g[j & 1];
    const float2 t1[4] = {make_float2(G.sh1[0].x, G.sh1[0].y), make_float2(G.sh1[0].z, G.sh1[0].w),
                          make_float2(G.sh1[1].x, G.sh1[1].y), make_float2(G.sh1[1].z, G.sh1[1].w)};
    const float2 s1[4] = {make_float2(G.sc1[0].x, G.sc1[0].y), make_float2(G.sc1[0].z, G.sc1[0].w),
                          make_float2(G.sc1[1].x, G.sc1[1].y), make_float2(G.sc1[1].z, G.sc1[1].w)};
    const float2 s2[4] = {make_float2(G.sc2[0].x, G.sc2[0].y), make_float2(G.sc2[0].z, G.sc2[0].w),
                          make_float2(G.sc2[1].x, G.sc2[1].y), make_float2(G.sc2[1].z, G.sc2[1].w)};
    const float2 t2[4] = {make_float2(G.sh2[0].x, G.sh2[0].y), make_float2(G.sh2[0].z, G.sh2[0].w),
                          make_float2(G.sh2[1].x, G.sh2[1].y), make_float2(G.sh2[1].z, G.sh2[1].w)};
    const uint32_t chunk = 4u * h + (uint32_t)j;  // 16-byte chunk of the 128-byte slab row
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const EpiRow& R = row[w];
      const uint32_t rw[4] = {G.res[w].x, G.res[w].y, G.res[w].z, G.res[w].w};
      const uint32_t uw[4] = {G.up[w].x, G.up[w].y, G.up[w].z, G.up[w].w};
      const uint32_t* a = acc[w] + 8 * j;
      uint32_t op[4], xp[4];
      (void)xp;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 ac = make_float2(__uint_as_float(a[2 * e]), __uint_as_float(a[2 * e + 1]));
        float2 v = UNIT ? fadd2(ac, t1[e]) : ffma2(ac, s1[e], t1[e]);
        if (RES) v = fadd2(v, bf2_unpack(rw[e]));
        if (RES2 && !RELU) {  // up1 + nearest_x2(low3): the sum is rounded to bf16 first, like a stored up1
          op[e] = add_bf16x2(pack2(v.x, v.y), uw[e]);
        } else {
          if (RES2) v = fadd2(bf2_unpack(pack2(v.x, v.y)), bf2_unpack(uw[e]));
          op[e] = RELU ? pack2_relu(v.x, v.y) : pack2(v.x, v.y);
        }
        if (XSRC == 2) {  // act = relu(bn(bf16(v))): the rounded value is the packed one
          const float2 x = ffma2(bf2_unpack(op[e]), s2[e], t2[e]);
          xp[e] = pack2_relu(x.x, x.y);
        }
      }
      // 8 channels = 4 packed operand columns: lanes < 16 write columns [4 j, 4 j + 4), the others 16 further
      if (STAGED) {
        if (OUT) sts128(R.raw_s + ((chunk ^ R.sw) << 4), make_uint4(op[0], op[1], op[2], op[3]));
        if (XSRC == 2) sts128(R.act_s + ((chunk ^ R.sw) << 4), make_uint4(xp[0], xp[1], xp[2], xp[3]));
        continue;
      }
      if (XSRC == 1) tmem_st_16x32bx2_x4(x_addr + ((uint32_t)(16 * w) << 16) + 4 * j, op);
      if (XSRC == 2) tmem_st_16x32bx2_x4(x_addr + ((uint32_t)(16 * w) << 16) + 4 * j, xp);
      if (OUT && (RES || INPLACE)) {
        // Stored stage with a residual: bf16(v) replaces the residual chunk it was computed from, in place in
        // the slab; the warp's 32 rows then leave with one TMA store (run_stage).  256-bit stores from
        // registers cost the LSU data pipe ~50 wavefronts per instruction (measured: 4 900 of 9 500 per tile).
        sts128(R.rrow_s + ((chunk ^ R.sw) << 4), make_uint4(op[0], op[1], op[2], op[3]));
      } else if (OUT) {
        // two groups = 16 channels = 32 contiguous bytes of this pixel: one full sector per store.  (Staging in
        // shared memory for a TMA store was measured too: the 32 KB of staging cost two slots of the weight
        // ring or a residual slab.)
        if ((j & 1) == 0) {
#pragma unroll
          for (int e = 0; e < 4; ++e) held[w][e] = op[e];
        } else if (R.store) {
          const uint32_t both[8] = {held[w][0], held[w][1], held[w][2], held[w][3], op[0], op[1], op[2], op[3]};
          stg256(R.out + ((chunk - 1u) << 4), both);
        }
      }
    }
  }
}

}  // namespace df3d
