// Image ingest on the device: bilinear resize of uint8 gray images to the network input.
// Replaces the host-side cv2.resize(..., INTER_LINEAR) of the image loader in front of
// df2d.inference.inference_folder (reference call site df3d/core.py:177-185).
// Bit-identical to OpenCV's fixed-point INTER_LINEAR for 8-bit images (11-bit coefficients, int32 rows,
// 22-bit vertical cast; oracle/ingest.py is the CPU restatement, pinned against cv2.resize): the same
// frames reach the hourglass whether they were resized on the host or here.
// HBM-bound: every source byte is read once through L2, 1 byte written per output pixel.
#include "common.cuh"

namespace df3d {

// OpenCV: fx = (float)((d + 0.5) * scale - 0.5) in double, s = floor(fx), fx -= s; columns snap to the border
// pixel with weight 2048, rows keep the fraction and clamp the two indices separately (oracle/ingest.py).
// The double expression is evaluated without FMA contraction, like the host compiler does.
__device__ __forceinline__ void lin_coef(int d, double scale, int src, bool clamp_fraction, int& i0, int& i1, int& a0,
                                         int& a1) {
  const float f0 = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
  int s = (int)floorf(f0);
  float f = __fsub_rn(f0, (float)s);
  if (clamp_fraction) {
    if (s < 0) {
      s = 0;
      f = 0.0f;
    }
    if (s >= src - 1) {
      s = src - 1;
      f = 0.0f;
    }
  }
  i0 = min(max(s, 0), src - 1);
  i1 = min(max(s + 1, 0), src - 1);
  a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));  // saturate_cast<short>: round half to even
  a1 = __float2int_rn(__fmul_rn(f, 2048.0f));
}

// One thread = four consecutive output pixels of one row (one 32-bit store).
__global__ void __launch_bounds__(256)
resize_gray_u8_kernel(const uint8_t* __restrict__ src, int B, int Hs, int Ws, uint8_t* __restrict__ dst, int Hd, int Wd,
                      double scale_x, double scale_y) {
  const int wq = Wd >> 2;
  const long long n = (long long)B * Hd * wq;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int xq = (int)(g % wq);
  const int y = (int)((g / wq) % Hd);
  const int b = (int)(g / ((long long)wq * Hd));
  int y0, y1, b0, b1;
  lin_coef(y, scale_y, Hs, false, y0, y1, b0, b1);
  const uint8_t* r0 = src + ((size_t)b * Hs + y0) * Ws;
  const uint8_t* r1 = src + ((size_t)b * Hs + y1) * Ws;
  uint32_t packed = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int x0, x1, a0, a1;
    lin_coef(4 * xq + e, scale_x, Ws, true, x0, x1, a0, a1);
    const int h0 = (int)__ldg(r0 + x0) * a0 + (int)__ldg(r0 + x1) * a1;  // horizontal pass, scale 2^11
    const int h1 = (int)__ldg(r1 + x0) * a0 + (int)__ldg(r1 + x1) * a1;
    int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;  // vertical pass, 22-bit cast
    v = min(max(v, 0), 255);
    packed |= (uint32_t)v << (8 * e);
  }
  *reinterpret_cast<uint32_t*>(dst + ((size_t)b * Hd + y) * Wd + 4 * xq) = packed;
}

}  // namespace df3d

extern "C" int df3d_resize_gray_u8(const uint8_t* src_dev, int B, int Hs, int Ws, uint8_t* dst_dev, int Hd, int Wd,
                                   void* stream) {
  using namespace df3d;
  DF3D_REQUIRE(B >= 0 && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0, DF3D_EINVAL, "df3d_resize_gray_u8: bad shape");
  if (B == 0) return DF3D_OK;
  DF3D_REQUIRE(src_dev && dst_dev, DF3D_EINVAL, "df3d_resize_gray_u8: null pointer");
  DF3D_REQUIRE(Wd % 4 == 0 && (reinterpret_cast<uintptr_t>(dst_dev) & 3) == 0, DF3D_EUNSUPPORTED,
               "df3d_resize_gray_u8: output width %d must be a multiple of 4 and the output 4-byte aligned", Wd);
  const long long n = (long long)B * Hd * (Wd / 4);
  DF3D_REQUIRE((n + 255) / 256 < (1ll << 31), DF3D_EUNSUPPORTED, "df3d_resize_gray_u8: too many pixels");
  resize_gray_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src_dev, B, Hs, Ws, dst_dev, Hd, Wd, (double)Ws / (double)Wd, (double)Hs / (double)Hd);
  DF3D_LAUNCH_CHECK("resize_gray_u8_kernel");
  return DF3D_OK;
}
