// Heat-map decode: per-(image, joint) first-occurrence argmax + peak value.
// Replaces the read-out inside df2d.inference.inference_folder (reference call site
// df3d/core.py:177-185, rule at README.md:404).  HBM-bound scan: every element is read once
// with 128-bit loads, reduced with warp shuffles.
#include <cuda_bf16.h>

#include "common.cuh"

namespace df3d {

struct Best {
  float v;
  int i;
};

// first occurrence wins: larger value, or equal value with the smaller flat index
__device__ __forceinline__ void take(Best& a, float v, int i) {
  if (v > a.v || (v == a.v && i < a.i)) {
    a.v = v;
    a.i = i;
  }
}

__device__ __forceinline__ Best warp_best(Best b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, b.v, o);
    int oi = __shfl_xor_sync(0xffffffffu, b.i, o);
    take(b, ov, oi);
  }
  return b;
}

constexpr int kArgmaxThreads = 256;

// One CTA per (b,k) plane of an NCHW tensor.  n = H*W.
template <typename T>
__global__ void __launch_bounds__(kArgmaxThreads)
argmax_nchw_kernel(const T* __restrict__ hm, int n, int32_t* __restrict__ idx, float* __restrict__ conf) {
  const T* plane = hm + (size_t)blockIdx.x * n;
  Best best{-INFINITY, 0x7fffffff};
  constexpr int V = 16 / sizeof(T);  // elements per 128-bit load
  const bool vec_ok = (n % V == 0) && ((reinterpret_cast<uintptr_t>(plane) & 15) == 0);
  if (vec_ok) {
    const uint4* p4 = reinterpret_cast<const uint4*>(plane);
    for (int q = threadIdx.x; q < n / V; q += kArgmaxThreads) {
      uint4 raw = __ldg(p4 + q);
      if constexpr (sizeof(T) == 4) {
        const float* f = reinterpret_cast<const float*>(&raw);
#pragma unroll
        for (int e = 0; e < 4; ++e) take(best, f[e], q * 4 + e);
      } else {
        const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
        for (int e = 0; e < 8; ++e) take(best, __bfloat162float(h[e]), q * 8 + e);
      }
    }
  } else {
    for (int q = threadIdx.x; q < n; q += kArgmaxThreads) {
      float v;
      if constexpr (sizeof(T) == 4) v = plane[q]; else v = __bfloat162float(plane[q]);
      take(best, v, q);
    }
  }
  best = warp_best(best);
  __shared__ float sv[kArgmaxThreads / 32];
  __shared__ int si[kArgmaxThreads / 32];
  if ((threadIdx.x & 31) == 0) {
    sv[threadIdx.x >> 5] = best.v;
    si[threadIdx.x >> 5] = best.i;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    Best b{-INFINITY, 0x7fffffff};
    if (threadIdx.x < kArgmaxThreads / 32) b = Best{sv[threadIdx.x], si[threadIdx.x]};
    b = warp_best(b);
    if (threadIdx.x == 0) {
      idx[blockIdx.x] = b.i;
      conf[blockIdx.x] = b.v;
    }
  }
}

// One CTA per image of an NHWC float32 tensor with Cpad (multiple of 4, <= 32) channels.
// A thread owns one channel quad; Cpad/4 consecutive threads cover a pixel (coalesced).
__global__ void __launch_bounds__(kArgmaxThreads)
argmax_nhwc_kernel(const float* __restrict__ hm, int npix, int Cpad, int K,
                   int32_t* __restrict__ idx, float* __restrict__ conf) {
  const int tpp = Cpad >> 2;                   // threads per pixel
  const int quad = threadIdx.x % tpp;
  const int prow = threadIdx.x / tpp;
  const int pstep = kArgmaxThreads / tpp;
  const float4* base = reinterpret_cast<const float4*>(hm + (size_t)blockIdx.x * npix * Cpad);
  Best b[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) b[e] = Best{-INFINITY, 0x7fffffff};
  for (int p = prow; p < npix; p += pstep) {
    float4 v = __ldg(base + (size_t)p * tpp + quad);
    take(b[0], v.x, p);
    take(b[1], v.y, p);
    take(b[2], v.z, p);
    take(b[3], v.w, p);
  }
  __shared__ float sv[kArgmaxThreads][4];
  __shared__ int si[kArgmaxThreads][4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    sv[threadIdx.x][e] = b[e].v;
    si[threadIdx.x][e] = b[e].i;
  }
  __syncthreads();
  // thread c (< K) scans the pstep partial results of its channel in pixel order
  if (threadIdx.x < K) {
    const int c = threadIdx.x;
    Best r{-INFINITY, 0x7fffffff};
    for (int g = 0; g < pstep; ++g) {
      int t = g * tpp + (c >> 2);
      take(r, sv[t][c & 3], si[t][c & 3]);
    }
    idx[(size_t)blockIdx.x * K + c] = r.i;
    conf[(size_t)blockIdx.x * K + c] = r.v;
  }
}

}  // namespace df3d

extern "C" int df3d_heatmap_argmax(const void* hm_dev, int dtype, int B, int K, int H, int W,
                                   int32_t* idx_dev, float* conf_dev, void* stream) {
  using namespace df3d;
  DF3D_REQUIRE(B >= 0 && K > 0 && H > 0 && W > 0, DF3D_EINVAL, "df3d_heatmap_argmax: bad shape");
  if (B == 0) return DF3D_OK;  // empty batch: nothing to decode (pointers may be null)
  DF3D_REQUIRE(hm_dev && idx_dev && conf_dev, DF3D_EINVAL, "df3d_heatmap_argmax: null pointer");
  DF3D_REQUIRE(dtype == 0 || dtype == 1, DF3D_EINVAL, "df3d_heatmap_argmax: dtype must be 0 (f32) or 1 (bf16)");
  DF3D_REQUIRE((long long)H * W < (1ll << 30), DF3D_EINVAL, "df3d_heatmap_argmax: plane too large");
  if (B == 0) return DF3D_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == 0)
    argmax_nchw_kernel<float><<<B * K, kArgmaxThreads, 0, s>>>(static_cast<const float*>(hm_dev), H * W, idx_dev, conf_dev);
  else
    argmax_nchw_kernel<__nv_bfloat16><<<B * K, kArgmaxThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(hm_dev), H * W, idx_dev, conf_dev);
  DF3D_LAUNCH_CHECK("argmax_nchw_kernel");
  return DF3D_OK;
}

// Decode of the keys the score-head epilogue accumulates (conv_gemm.cu, ConvParams::amax_keys): key =
// order-preserving bits of the peak << 32 | ~flat index.  Also clears the keys for the next forward.
__global__ void argmax_keys_decode_kernel(unsigned long long* __restrict__ keys, int n, int Cpad, int K,
                                          int32_t* __restrict__ idx, float* __restrict__ conf) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n * Cpad) return;
  const int b = g / Cpad, c = g % Cpad;
  const unsigned long long key = keys[g];
  keys[g] = 0ull;
  if (c >= K) return;
  const uint32_t o = (uint32_t)(key >> 32);
  const uint32_t bits = o ^ ((o >> 31) ? 0x80000000u : 0xffffffffu);
  idx[(size_t)b * K + c] = (int32_t)(0xffffffffu - (uint32_t)key);
  conf[(size_t)b * K + c] = __uint_as_float(bits);
}

namespace df3d {
int launch_argmax_keys_decode(unsigned long long* keys, int B, int Cpad, int K, int32_t* idx, float* conf, cudaStream_t s) {
  if (B == 0) return DF3D_OK;
  const int n = B * Cpad;
  argmax_keys_decode_kernel<<<(n + 255) / 256, 256, 0, s>>>(keys, B, Cpad, K, idx, conf);
  DF3D_LAUNCH_CHECK("argmax_keys_decode_kernel");
  return DF3D_OK;
}
}  // namespace df3d

extern "C" int df3d_heatmap_argmax_nhwc(const float* hm_dev, int B, int H, int W, int Cpad, int K,
                                        int32_t* idx_dev, float* conf_dev, void* stream) {
  using namespace df3d;
  DF3D_REQUIRE(B >= 0 && H > 0 && W > 0, DF3D_EINVAL, "df3d_heatmap_argmax_nhwc: bad shape");
  if (B == 0) return DF3D_OK;
  DF3D_REQUIRE(hm_dev && idx_dev && conf_dev, DF3D_EINVAL, "df3d_heatmap_argmax_nhwc: null pointer");
  DF3D_REQUIRE(Cpad % 4 == 0 && Cpad >= 4 && Cpad <= 32 && K >= 1 && K <= Cpad, DF3D_EINVAL,
               "df3d_heatmap_argmax_nhwc: need Cpad%%4==0, Cpad<=32, K<=Cpad");
  DF3D_REQUIRE(kArgmaxThreads % (Cpad / 4) == 0, DF3D_EINVAL, "df3d_heatmap_argmax_nhwc: Cpad/4 must divide 256");
  if (B == 0) return DF3D_OK;
  argmax_nhwc_kernel<<<B, kArgmaxThreads, 0, static_cast<cudaStream_t>(stream)>>>(hm_dev, H * W, Cpad, K, idx_dev, conf_dev);
  DF3D_LAUNCH_CHECK("argmax_nhwc_kernel");
  return DF3D_OK;
}
