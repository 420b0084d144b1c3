// HBM-bound elementwise stages of the hourglass forward (NHWC bf16, 128-bit accesses):
//   stem_im2col  : image -> [B][H/2][W/2][192] patch matrix of the 7x7/2 stem conv (K = 147, zero padded)
//                  with the uint8 -> float normalisation and the left-right mirror of the
//                  "flipped" cameras (reference df3d/core.py:179) fused into the gather
//   maxpool_bn_relu      : 2x2/2 max-pool -> raw + relu(bn(raw)) for the next bottleneck
// (the hourglass' nearest x2 up-sample + add lives in the conv epilogue, see conv_gemm.cu)
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
constexpr int kStemPix = 32;                      // output pixels of one row segment per CTA
constexpr int kStemCols = 2 * kStemPix + 5;       // input columns feeding them (7-wide window, stride 2)

// One CTA = 32 consecutive output pixels of one output row.  The 7 x 69 x 3 input window is staged
// in shared memory once (normalised, mirrored, zero padded), then the 32 x 24 sixteen-byte chunks of
// the patch matrix are written fully coalesced.
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const void* __restrict__ img, int dtype, const uint8_t* __restrict__ flip, int B, int H, int W,
                   float m0, float m1, float m2, __nv_bfloat16* __restrict__ out) {
  __shared__ float win[3][7][kStemCols + 1];
  const int Ho = H / 2, Wo = W / 2;
  const int segs = Wo / kStemPix;
  int blk = blockIdx.x;
  const int seg = blk % segs;
  blk /= segs;
  const int oy = blk % Ho;
  const int b = blk / Ho;
  const int ox0 = seg * kStemPix;
  const bool fl = flip ? (flip[b] != 0) : false;
  const float mean[3] = {m0, m1, m2};
  for (int i = threadIdx.x; i < 7 * kStemCols; i += 256) {
    const int ky = i / kStemCols, cx = i - ky * kStemCols;
    const int iy = 2 * oy + ky - 3;
    int ix = 2 * ox0 + cx - 3;
    float v[3] = {0.0f, 0.0f, 0.0f};
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      if (fl) ix = W - 1 - ix;
      if (dtype == 0) {
        const float g = (float)static_cast<const uint8_t*>(img)[((size_t)b * H + iy) * W + ix] / 255.0f;
        v[0] = g - mean[0];
        v[1] = g - mean[1];
        v[2] = g - mean[2];
      } else {
        const float* f = static_cast<const float*>(img) + (size_t)b * 3 * H * W + (size_t)iy * W + ix;
        v[0] = f[0];
        v[1] = f[(size_t)H * W];
        v[2] = f[(size_t)2 * H * W];
      }
    }
    win[0][ky][cx] = v[0];
    win[1][ky][cx] = v[1];
    win[2][ky][cx] = v[2];
  }
  __syncthreads();
  __nv_bfloat16* row = out + (((size_t)b * Ho + oy) * Wo + ox0) * kStemKPad;
  for (int i = threadIdx.x; i < kStemPix * (kStemKPad / 8); i += 256) {
    const int px = i / (kStemKPad / 8), chunk = i - px * (kStemKPad / 8);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = chunk * 8 + e;
      float val = 0.0f;
      if (k < kStemK) {
        const int tap = k / 3, c = k - tap * 3;
        const int ky = tap / 7, kx = tap - ky * 7;
        val = win[c][ky][2 * px + kx];
      }
      v[e] = val;
    }
    *reinterpret_cast<uint4*>(row + (size_t)i * 8) = pack8(v);
  }
}

// Gray fast path of the stem: the three input planes are the same image (x/255 - mean with one
// common mean), so the 7x7x3 conv is a 7x7x1 conv with the weights summed over the input channel and
// the patch matrix has 49 (padded to 64) columns instead of 147 (192).
constexpr int kGrayMaxW = kStemGrayMaxW;  // widest input row staged in shared memory

// One CTA = one output row of one image.  The seven input rows feeding it are staged in shared memory as
// bf16 bits through a 256-entry table of bf16(u8 / 255 - mean) (the same expression and rounding as the
// per-element path, evaluated once per grey level), zero padded, mirrored for the flipped cameras; then
// every thread gathers 16-byte chunks (8 patch columns) of the patch matrix, written fully coalesced.
// (The first version used one CTA per 32 output pixels, a constant-memory tap table indexed per lane and
// scalar byte loads: 5.7 ms per 1792 images, 680 GB/s.)
__global__ void __launch_bounds__(256)
stem_im2col_gray_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ flip, int B, int H, int W, float mean,
                        __nv_bfloat16* __restrict__ out) {
  __shared__ uint16_t lut[256];
  __shared__ __align__(16) uint16_t win[7][kGrayMaxW + 8];  // column c holds input x = c - 3
  const int Ho = H / 2, Wo = W / 2;
  const int oy = blockIdx.x % Ho, b = blockIdx.x / Ho;
  const bool fl = flip ? (flip[b] != 0) : false;
  {
    const __nv_bfloat16 h = __float2bfloat16_rn((float)threadIdx.x / 255.0f - mean);
    lut[threadIdx.x] = *reinterpret_cast<const uint16_t*>(&h);
  }
  for (int i = threadIdx.x; i < 7 * 8; i += 256) {  // the zero padding left and right of the rows
    const int ky = i >> 3, c = i & 7;
    win[ky][c < 3 ? c : W + c] = 0;
  }
  __syncthreads();
  const int vec_per_row = W >> 4;  // 16 pixels per 128-bit load (W is a multiple of 64)
  for (int i = threadIdx.x; i < 7 * vec_per_row; i += 256) {
    const int ky = i / vec_per_row, vx = i - ky * vec_per_row;
    const int iy = 2 * oy + ky - 3;
    uint4 q = make_uint4(0, 0, 0, 0);
    const bool inside = iy >= 0 && iy < H;
    if (inside) q = __ldg(reinterpret_cast<const uint4*>(img + ((size_t)b * H + iy) * W) + vx);
    const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const uint32_t px = (w4[e >> 2] >> (8 * (e & 3))) & 0xffu;
      const int ix = 16 * vx + e;
      win[ky][3 + (fl ? W - 1 - ix : ix)] = inside ? lut[px] : (uint16_t)0;
    }
  }
  __syncthreads();
  __nv_bfloat16* row = out + ((size_t)b * Ho + oy) * Wo * kStemKGray;
  for (int item = threadIdx.x; item < Wo * 8; item += 256) {
    const int px = item >> 3, chunk = item & 7;
    uint32_t v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = chunk * 8 + e;       // patch column = ky * 7 + kx, 49 real columns
      const int ky = (k * 37) >> 8;      // k / 7 for k < 64
      const int kx = k - 7 * ky;
      v[e] = k < 49 ? (uint32_t)win[ky][2 * px + kx] : 0u;
    }
    uint4 o;
    o.x = v[0] | (v[1] << 16);
    o.y = v[2] | (v[3] << 16);
    o.z = v[4] | (v[5] << 16);
    o.w = v[6] | (v[7] << 16);
    *reinterpret_cast<uint4*>(row + (size_t)item * 8) = o;
  }
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

static unsigned grid_for(long long total) { return (unsigned)((total + 255) / 256); }

int launch_stem_im2col(const void* img, int dtype, const uint8_t* flip, int B, int H, int W, const float mean[3],
                       __nv_bfloat16* out, cudaStream_t s) {
  if (B == 0) return DF3D_OK;
  DF3D_REQUIRE((W / 2) % kStemPix == 0, DF3D_EUNSUPPORTED, "stem_im2col: input width must be a multiple of %d", 2 * kStemPix);
  const long long blocks = (long long)B * (H / 2) * ((W / 2) / kStemPix);
  DF3D_REQUIRE(blocks < (1ll << 31), DF3D_EUNSUPPORTED, "stem_im2col: too many blocks");
  stem_im2col_kernel<<<(unsigned)blocks, 256, 0, s>>>(img, dtype, flip, B, H, W, mean[0], mean[1], mean[2], out);
  DF3D_LAUNCH_CHECK("stem_im2col_kernel");
  return DF3D_OK;
}

int launch_stem_im2col_gray(const uint8_t* img, const uint8_t* flip, int B, int H, int W, float mean, __nv_bfloat16* out,
                            cudaStream_t s) {
  if (B == 0) return DF3D_OK;
  DF3D_REQUIRE((W / 2) % kStemPix == 0, DF3D_EUNSUPPORTED, "stem_im2col: input width must be a multiple of %d", 2 * kStemPix);
  DF3D_REQUIRE(W % 64 == 0 && W <= kGrayMaxW && (reinterpret_cast<uintptr_t>(img) & 15) == 0, DF3D_EUNSUPPORTED,
               "stem_im2col (gray): input width %d must be a multiple of 64, at most %d, rows 16-byte aligned", W, kGrayMaxW);
  const long long blocks = (long long)B * (H / 2);
  DF3D_REQUIRE(blocks < (1ll << 31), DF3D_EUNSUPPORTED, "stem_im2col: too many blocks");
  stem_im2col_gray_kernel<<<(unsigned)blocks, 256, 0, s>>>(img, flip, B, H, W, mean, out);
  DF3D_LAUNCH_CHECK("stem_im2col_gray_kernel");
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

}  // namespace df3d
