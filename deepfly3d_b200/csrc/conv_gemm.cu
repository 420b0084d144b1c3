// tcgen05 implicit-GEMM convolution kernel (see conv_gemm.cuh).
//
// GEMM view:  D[M = pixels, N = Cout] = sum over taps, cin of A[pixel shifted by tap, cin] * W[cout, tap, cin]
//   * A tile: 128 pixels (nb x th x tw box of the NHWC tensor) x 64 channels, loaded by ONE 4-D TMA
//     per (tap, channel block); the 3x3 halo / zero padding is the TMA's out-of-bounds zero fill
//     (coordinates x0+dx-1, y0+dy-1 may be -1 or W/H).
//   * B tile: BN output channels x 64 K, 2-D TMA from the packed [CoutPad][taps*CinPad] weights.
//   * one CTA per SM, persistent over tiles; warp 0 = TMA producer, warp 1 = MMA issuer (one
//     elected thread) + TMEM owner, warps 2..5 = epilogue (one TMEM lane quarter each).
//   * TMEM holds two accumulator stages (2 x BN fp32 columns) so the epilogue of tile i overlaps
//     the MMAs of tile i+1.
#include "conv_gemm.cuh"
#include "sm100.cuh"

namespace df3d {

using namespace sm100;

constexpr int kConvThreads = 192;
constexpr int kTileM = 128;
constexpr int kABytes = kTileM * 128;  // 128 rows x 64 bf16

template <int BN>
struct ConvCfg {
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;  // power of two for BN in {32,64,128,256}
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_round(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }

template <int BN>
__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned tiles
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  auto a_addr = [&](int s) { return smem_base + s * Cfg::kStageBytes; };
  auto b_addr = [&](int s) { return smem_base + s * Cfg::kStageBytes + kABytes; };
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  const int total_tiles = m_tiles * p.n_tiles_n;
  const int num_kb = p.taps * p.kc_per_tap;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.tmA);
    prefetch_tensormap(&p.tmB);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles_n;
        int mt = tile / p.n_tiles_n;
        const int tx = mt % p.tiles_x;
        mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int tb = mt / p.tiles_y;
        const int x0 = tx * p.tw, y0 = ty * p.th, n0 = tb * p.nb;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / p.kc_per_tap, kc = kb - tap * p.kc_per_tap;
          int dx = 0, dy = 0;
          if (p.taps == 9) {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
          }
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
          tma_load_4d(a_addr(stage), &p.tmA, full_bar(stage), kc * 64, x0 + dx, y0 + dy, n0);
          tma_load_2d(b_addr(stage), &p.tmB, full_bar(stage), kb * 64, nt * BN);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t adesc = umma_smem_desc_sw128(a_addr(stage));
          const uint64_t bdesc = umma_smem_desc_sw128(b_addr(stage));
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 4 x (K = 16): +32 bytes inside the 128B swizzle atom
            umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(empty_bar(stage));  // frees the smem stage once these MMAs retire
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(tfull_bar(as));  // accumulator ready for the epilogue
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
      const int nt = tile % p.n_tiles_n;
      int mt = tile / p.n_tiles_n;
      const int tx = mt % p.tiles_x;
      mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int tb = mt / p.tiles_y;
      const int x = tx * p.tw + m % p.tw;
      const int y = ty * p.th + (m / p.tw) % p.th;
      const int n = tb * p.nb + m / (p.tw * p.th);
      const bool valid = n < p.B;
      const size_t pix = ((size_t)n * p.H + y) * p.W + x;

      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + c0, r);
        tmem_ld_wait();
        if (valid) {
          const int cbase = nt * BN + c0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {  // 8 channels at a time
            const int c = cbase + j * 8;
            float v[8];
            const float4 s1a = __ldg(reinterpret_cast<const float4*>(p.scale1 + c));
            const float4 s1b = __ldg(reinterpret_cast<const float4*>(p.scale1 + c + 4));
            const float4 h1a = __ldg(reinterpret_cast<const float4*>(p.shift1 + c));
            const float4 h1b = __ldg(reinterpret_cast<const float4*>(p.shift1 + c + 4));
            const float s1[8] = {s1a.x, s1a.y, s1a.z, s1a.w, s1b.x, s1b.y, s1b.z, s1b.w};
            const float h1[8] = {h1a.x, h1a.y, h1a.z, h1a.w, h1b.x, h1b.y, h1b.z, h1b.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = fmaf(__uint_as_float(r[j * 8 + e]), s1[e], h1[e]);
            if (p.residual) {
              const uint4 rr = __ldg(reinterpret_cast<const uint4*>(p.residual + pix * p.res_ld + c));
              const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(rh[e]);
                v[2 * e] += f.x;
                v[2 * e + 1] += f.y;
              }
            }
            if (p.relu1) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.0f);
            }
            if (p.out_f32) {
              float4* o = reinterpret_cast<float4*>(p.out_f32 + pix * p.f32_ld + c);
              o[0] = make_float4(v[0], v[1], v[2], v[3]);
              o[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
            if (p.out_raw) {
              uint4 o;
              o.x = pack_bf16x2(v[0], v[1]);
              o.y = pack_bf16x2(v[2], v[3]);
              o.z = pack_bf16x2(v[4], v[5]);
              o.w = pack_bf16x2(v[6], v[7]);
              *reinterpret_cast<uint4*>(p.out_raw + pix * p.raw_ld + c) = o;
            }
            if (p.out_act) {
              const float4 s2a = __ldg(reinterpret_cast<const float4*>(p.scale2 + c));
              const float4 s2b = __ldg(reinterpret_cast<const float4*>(p.scale2 + c + 4));
              const float4 h2a = __ldg(reinterpret_cast<const float4*>(p.shift2 + c));
              const float4 h2b = __ldg(reinterpret_cast<const float4*>(p.shift2 + c + 4));
              const float s2[8] = {s2a.x, s2a.y, s2a.z, s2a.w, s2b.x, s2b.y, s2b.z, s2b.w};
              const float h2[8] = {h2a.x, h2a.y, h2a.z, h2a.w, h2b.x, h2b.y, h2b.z, h2b.w};
              float w[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) w[e] = fmaxf(fmaf(bf16_round(v[e]), s2[e], h2[e]), 0.0f);
              uint4 o;
              o.x = pack_bf16x2(w[0], w[1]);
              o.y = pack_bf16x2(w[2], w[3]);
              o.z = pack_bf16x2(w[4], w[5]);
              o.w = pack_bf16x2(w[6], w[7]);
              *reinterpret_cast<uint4*>(p.out_act + pix * p.act_ld + c) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
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
  DF3D_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<32>::kSmemBytes));
  DF3D_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<64>::kSmemBytes));
  DF3D_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<128>::kSmemBytes));
  DF3D_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<256>::kSmemBytes));
  return DF3D_OK;
}

int launch_conv_gemm(const ConvParams& p, int BN, int num_sms, cudaStream_t stream) {
  const int total = p.tiles_x * p.tiles_y * p.tiles_b * p.n_tiles_n;
  if (total <= 0) return DF3D_OK;
  const int grid = total < num_sms ? total : num_sms;
  switch (BN) {
    case 32: conv_gemm_kernel<32><<<grid, kConvThreads, ConvCfg<32>::kSmemBytes, stream>>>(p); break;
    case 64: conv_gemm_kernel<64><<<grid, kConvThreads, ConvCfg<64>::kSmemBytes, stream>>>(p); break;
    case 128: conv_gemm_kernel<128><<<grid, kConvThreads, ConvCfg<128>::kSmemBytes, stream>>>(p); break;
    case 256: conv_gemm_kernel<256><<<grid, kConvThreads, ConvCfg<256>::kSmemBytes, stream>>>(p); break;
    default: DF3D_REQUIRE(false, DF3D_EUNSUPPORTED, "launch_conv_gemm: BN=%d not instantiated", BN);
  }
  DF3D_LAUNCH_CHECK("conv_gemm_kernel");
  return DF3D_OK;
}

}  // namespace df3d
