// HBM-bound elementwise stages of the hourglass forward (NHWC bf16, 128-bit accesses):
//   stem_im2col  : image -> [B][H/2][W/2][192] patch matrix of the 7x7/2 stem conv (K = 147, zero padded)
//                  with the uint8 -> float normalisation and the left-right mirror of the
//                  "flipped" cameras (reference df3d/core.py:179) fused into the gather
//   maxpool_bn_relu      : 2x2/2 max-pool -> raw + relu(bn(raw)) for the next bottleneck
//   upsample_add_bn_relu : up1 + nearest_x2(low3) -> raw + relu(bn(raw))
#include <cuda_bf16.h>

#include "common.cuh"
#include "hg_elementwise.cuh"

namespace df3d {

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __bfloat1622float2(h[e]);
    f[2 * e] = t.x;
    f[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack2(f[0], f[1]);
  o.y = pack2(f[2], f[3]);
  o.z = pack2(f[4], f[5]);
  o.w = pack2(f[6], f[7]);
  return o;
}
// act = bf16(relu(bf16(raw) * s + t)) on 8 channels starting at c
__device__ __forceinline__ uint4 bn_relu8(const uint4& raw_bits, const float* __restrict__ scale,
                                          const float* __restrict__ shift, int c) {
  float r[8];
  unpack8(raw_bits, r);
  const float4 sa = __ldg(reinterpret_cast<const float4*>(scale + c));
  const float4 sb = __ldg(reinterpret_cast<const float4*>(scale + c + 4));
  const float4 ta = __ldg(reinterpret_cast<const float4*>(shift + c));
  const float4 tb = __ldg(reinterpret_cast<const float4*>(shift + c + 4));
  const float s[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
  const float t[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
  float o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = fmaxf(fmaf(r[e], s[e], t[e]), 0.0f);
  return pack8(o);
}

constexpr int kStemK = 147, kStemKPad = 192;

__global__ void __launch_bounds__(256)
stem_im2col_kernel(const void* __restrict__ img, int dtype, const uint8_t* __restrict__ flip, int B, int H, int W,
                   float m0, float m1, float m2, __nv_bfloat16* __restrict__ out) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * Ho * Wo * (kStemKPad / 8);
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const int chunk = (int)(g % (kStemKPad / 8));
  long long pixel = g / (kStemKPad / 8);
  const int ox = (int)(pixel % Wo);
  const int oy = (int)((pixel / Wo) % Ho);
  const int b = (int)(pixel / ((long long)Wo * Ho));
  const bool fl = flip ? (flip[b] != 0) : false;
  const float mean[3] = {m0, m1, m2};
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = chunk * 8 + e;
    float val = 0.0f;
    if (k < kStemK) {
      const int tap = k / 3, c = k - tap * 3;
      const int ky = tap / 7, kx = tap - ky * 7;
      const int iy = 2 * oy + ky - 3;
      int ix = 2 * ox + kx - 3;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        if (fl) ix = W - 1 - ix;
        if (dtype == 0) {
          val = (float)static_cast<const uint8_t*>(img)[((size_t)b * H + iy) * W + ix] / 255.0f - mean[c];
        } else {
          val = static_cast<const float*>(img)[(((size_t)b * 3 + c) * H + iy) * W + ix];
        }
      }
    }
    v[e] = val;
  }
  *reinterpret_cast<uint4*>(out + (size_t)pixel * kStemKPad + chunk * 8) = pack8(v);
}

__global__ void __launch_bounds__(256)
maxpool_bn_relu_kernel(const __nv_bfloat16* __restrict__ in, int B, int H, int W, int C,
                       const float* __restrict__ scale, const float* __restrict__ shift,
                       __nv_bfloat16* __restrict__ out_raw, __nv_bfloat16* __restrict__ out_act) {
  const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
  const long long total = (long long)B * Ho * Wo * C8;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const int c = (int)(g % C8) * 8;
  long long pixel = g / C8;
  const int ox = (int)(pixel % Wo);
  const int oy = (int)((pixel / Wo) % Ho);
  const int b = (int)(pixel / ((long long)Wo * Ho));
  const __nv_bfloat16* base = in + (((size_t)b * H + 2 * oy) * W + 2 * ox) * C + c;
  uint4 q[4];
  q[0] = __ldg(reinterpret_cast<const uint4*>(base));
  q[1] = __ldg(reinterpret_cast<const uint4*>(base + C));
  q[2] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)W * C));
  q[3] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)W * C + C));
  uint4 mx;
  {
    const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(&q[0]);
    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&q[1]);
    const __nv_bfloat162* c2 = reinterpret_cast<const __nv_bfloat162*>(&q[2]);
    const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&q[3]);
    __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(&mx);
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = __hmax2(__hmax2(a[e], b2[e]), __hmax2(c2[e], d2[e]));
  }
  const size_t off = (size_t)pixel * C + c;
  *reinterpret_cast<uint4*>(out_raw + off) = mx;
  if (out_act) *reinterpret_cast<uint4*>(out_act + off) = bn_relu8(mx, scale, shift, c);
}

__global__ void __launch_bounds__(256)
upsample_add_bn_relu_kernel(const __nv_bfloat16* __restrict__ up1, const __nv_bfloat16* __restrict__ low, int B, int H,
                            int W, int C, const float* __restrict__ scale, const float* __restrict__ shift,
                            __nv_bfloat16* __restrict__ out_raw, __nv_bfloat16* __restrict__ out_act) {
  const int C8 = C / 8;
  const long long total = (long long)B * H * W * C8;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const int c = (int)(g % C8) * 8;
  long long pixel = g / C8;
  const int x = (int)(pixel % W);
  const int y = (int)((pixel / W) % H);
  const int b = (int)(pixel / ((long long)W * H));
  const size_t off = (size_t)pixel * C + c;
  const size_t loff = (((size_t)b * (H / 2) + (y >> 1)) * (W / 2) + (x >> 1)) * C + c;
  float a[8], l[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(up1 + off)), a);
  unpack8(__ldg(reinterpret_cast<const uint4*>(low + loff)), l);
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] += l[e];
  const uint4 raw = pack8(a);
  *reinterpret_cast<uint4*>(out_raw + off) = raw;
  if (out_act) *reinterpret_cast<uint4*>(out_act + off) = bn_relu8(raw, scale, shift, c);
}

static unsigned grid_for(long long total) { return (unsigned)((total + 255) / 256); }

int launch_stem_im2col(const void* img, int dtype, const uint8_t* flip, int B, int H, int W, const float mean[3],
                       __nv_bfloat16* out, cudaStream_t s) {
  const long long total = (long long)B * (H / 2) * (W / 2) * (kStemKPad / 8);
  if (total == 0) return DF3D_OK;
  stem_im2col_kernel<<<grid_for(total), 256, 0, s>>>(img, dtype, flip, B, H, W, mean[0], mean[1], mean[2], out);
  DF3D_LAUNCH_CHECK("stem_im2col_kernel");
  return DF3D_OK;
}

int launch_maxpool_bn_relu(const __nv_bfloat16* in, int B, int H, int W, int C, const float* scale, const float* shift,
                           __nv_bfloat16* out_raw, __nv_bfloat16* out_act, cudaStream_t s) {
  const long long total = (long long)B * (H / 2) * (W / 2) * (C / 8);
  if (total == 0) return DF3D_OK;
  maxpool_bn_relu_kernel<<<grid_for(total), 256, 0, s>>>(in, B, H, W, C, scale, shift, out_raw, out_act);
  DF3D_LAUNCH_CHECK("maxpool_bn_relu_kernel");
  return DF3D_OK;
}

int launch_upsample_add_bn_relu(const __nv_bfloat16* up1, const __nv_bfloat16* low, int B, int H, int W, int C,
                                const float* scale, const float* shift, __nv_bfloat16* out_raw,
                                __nv_bfloat16* out_act, cudaStream_t s) {
  const long long total = (long long)B * H * W * (C / 8);
  if (total == 0) return DF3D_OK;
  upsample_add_bn_relu_kernel<<<grid_for(total), 256, 0, s>>>(up1, low, B, H, W, C, scale, shift, out_raw, out_act);
  DF3D_LAUNCH_CHECK("upsample_add_bn_relu_kernel");
  return DF3D_OK;
}

}  // namespace df3d
