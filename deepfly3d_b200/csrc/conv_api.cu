// Operator-level entry point: one convolution layer through the tcgen05 implicit-GEMM kernel.
// Not on the hot path (packs + uploads the weights on every call); exists so that every layer
// shape of the hourglass can be checked against a float32 reference in isolation.
#include <vector>

#include "conv_gemm.cuh"

namespace df3d {
static inline uint16_t f2bf_host(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
}  // namespace df3d

using namespace df3d;

extern "C" int df3d_conv2d_nhwc_bf16(const void* in_dev, int B, int H, int W, int Cin, const float* w_host, int Cout,
                                     int ksize, const float* scale1_host, const float* shift1_host, int relu1,
                                     const void* residual_dev, void* out_dev, const float* scale2_host,
                                     const float* shift2_host, void* out_act_dev, void* stream) {
  DF3D_REQUIRE(in_dev && w_host && scale1_host && shift1_host && out_dev, DF3D_EINVAL, "df3d_conv2d_nhwc_bf16: null pointer");
  DF3D_REQUIRE(ksize == 1 || ksize == 3, DF3D_EUNSUPPORTED, "df3d_conv2d_nhwc_bf16: ksize must be 1 or 3");
  DF3D_REQUIRE(Cin % 64 == 0 && Cin >= 64, DF3D_EUNSUPPORTED, "df3d_conv2d_nhwc_bf16: Cin must be a multiple of 64");
  DF3D_REQUIRE(Cout == 64 || Cout == 128 || Cout == 256, DF3D_EUNSUPPORTED,
               "df3d_conv2d_nhwc_bf16: Cout must be 64, 128 or 256");
  DF3D_REQUIRE(B >= 1 && H >= 1 && W >= 1, DF3D_EINVAL, "df3d_conv2d_nhwc_bf16: bad shape");
  DF3D_REQUIRE(!out_act_dev || (scale2_host && shift2_host), DF3D_EINVAL, "df3d_conv2d_nhwc_bf16: out_act needs scale2/shift2");
  int tw = W < 16 ? W : 16;
  int th = 128 / tw;
  if (th > H) th = H;
  DF3D_REQUIRE(128 % (tw * th) == 0 && W % tw == 0 && H % th == 0, DF3D_EUNSUPPORTED,
               "df3d_conv2d_nhwc_bf16: H, W must be powers of two (or multiples of the 16x8 tile)");
  const int nb = 128 / (tw * th);
  if (int e = tma_init()) return e;
  if (int e = conv_gemm_configure()) return e;
  int dev = 0, num_sms = 148;
  DF3D_CUDA(cudaGetDevice(&dev));
  DF3D_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));

  const int taps = ksize * ksize;
  const size_t K = (size_t)taps * Cin;
  std::vector<uint16_t> wp((size_t)Cout * K, 0);
  for (int co = 0; co < Cout; ++co)
    for (int ci = 0; ci < Cin; ++ci)
      for (int t = 0; t < taps; ++t) wp[(size_t)co * K + (size_t)t * Cin + ci] = f2bf_host(w_host[((size_t)co * Cin + ci) * taps + t]);
  std::vector<float> aff((size_t)4 * Cout, 0.0f);
  for (int i = 0; i < Cout; ++i) {
    aff[i] = scale1_host[i];
    aff[Cout + i] = shift1_host[i];
    if (scale2_host) aff[2 * Cout + i] = scale2_host[i];
    if (shift2_host) aff[3 * Cout + i] = shift2_host[i];
  }
  uint16_t* d_w = nullptr;
  float* d_a = nullptr;
  DF3D_CUDA(cudaMalloc(&d_w, wp.size() * 2));
  cudaError_t ce = cudaMalloc(&d_a, aff.size() * 4);
  if (ce == cudaSuccess) ce = cudaMemcpy(d_w, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(d_a, aff.data(), aff.size() * 4, cudaMemcpyHostToDevice);
  int rc = DF3D_OK;
  if (ce != cudaSuccess) {
    set_error("df3d_conv2d_nhwc_bf16: upload failed: %s", cudaGetErrorString(ce));
    rc = DF3D_ECUDA;
  }
  ConvParams p;
  memset(&p, 0, sizeof(p));
  if (!rc) rc = make_tmap_act(&p.tmA, in_dev, Cin, W, H, B, tw, th, nb);
  if (!rc) rc = make_tmap_wgt(&p.tmB, d_w, (int)K, Cout, Cout);
  if (!rc && residual_dev) rc = make_tmap_act(&p.tmRes, residual_dev, Cout, W, H, B, tw, th, nb);
  if (!rc) rc = make_tmap_act(&p.tmRaw, out_dev, Cout, W, H, B, tw, th, nb);
  if (!rc && out_act_dev) rc = make_tmap_act(&p.tmAct, out_act_dev, Cout, W, H, B, tw, th, nb);
  if (!rc) {
    p.taps = taps;
    p.kc_per_tap = Cin / 64;
    p.H = H;
    p.W = W;
    p.B = B;
    p.tw = tw;
    p.th = th;
    p.nb = nb;
    p.tiles_x = W / tw;
    p.tiles_y = H / th;
    p.tiles_b = (B + nb - 1) / nb;
    p.n_tiles_n = 1;
    p.scale1 = d_a;
    p.shift1 = d_a + Cout;
    p.scale2 = d_a + 2 * Cout;
    p.shift2 = d_a + 3 * Cout;
    p.relu1 = relu1;
    p.residual = static_cast<const __nv_bfloat16*>(residual_dev);
    p.res_ld = Cout;
    p.out_raw = static_cast<__nv_bfloat16*>(out_dev);
    p.raw_ld = Cout;
    p.out_act = static_cast<__nv_bfloat16*>(out_act_dev);
    p.act_ld = Cout;
    rc = launch_conv_gemm(p, Cout, num_sms, static_cast<cudaStream_t>(stream));
  }
  // the temporaries must outlive the kernel: this utility entry point synchronises
  cudaError_t se = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  if (!rc && se != cudaSuccess) {
    set_error("df3d_conv2d_nhwc_bf16: kernel failed: %s", cudaGetErrorString(se));
    rc = DF3D_ECUDA;
  }
  if (d_w) cudaFree(d_w);
  if (d_a) cudaFree(d_a);
  return rc;
}
