// JPEG decode on the device (luminance plane) through nvJPEG -- the first half of the ingest row
// (the reference expands videos to camera_C_img_I.jpg, df3d/core.py:446-459, and
// df2d reads them back on the host behind inference_folder, core.py:177-185).  nvJPEG is library code: this
// file only binds it.  It is loaded lazily with dlopen so that libdf3d_b200.so itself never depends on it --
// the default loader path (host decode, bit-identical to the reference's libjpeg read) needs nothing here.
// Parity: nvJPEG's inverse DCT is not libjpeg-turbo's; decoded frames differ by a few grey levels
// (tests/test_gpu_ingest.py states the measured bound), so this path is opt-in.
#include <dlfcn.h>
#include <nvjpeg.h>

#include "common.cuh"

struct df3d_jpeg {
  void* dl = nullptr;
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  nvjpegStatus_t (*create_simple)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*destroy)(nvjpegHandle_t) = nullptr;
  nvjpegStatus_t (*state_create)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*state_destroy)(nvjpegJpegState_t) = nullptr;
  nvjpegStatus_t (*get_info)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
  nvjpegStatus_t (*decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*,
                           cudaStream_t) = nullptr;
};

extern "C" void df3d_jpeg_destroy(df3d_jpeg* j) {
  if (!j) return;
  if (j->state && j->state_destroy) j->state_destroy(j->state);
  if (j->handle && j->destroy) j->destroy(j->handle);
  if (j->dl) dlclose(j->dl);
  delete j;
}

extern "C" int df3d_jpeg_create(df3d_jpeg** out) {
  using namespace df3d;
  DF3D_REQUIRE(out, DF3D_EINVAL, "df3d_jpeg_create: null pointer");
  *out = nullptr;
  df3d_jpeg* j = new df3d_jpeg();
  const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so"};
  for (const char* n : names) {
    j->dl = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (j->dl) break;
  }
  if (!j->dl) {
    delete j;
    set_error("df3d_jpeg_create: nvJPEG not found (%s)", dlerror());
    return DF3D_EUNSUPPORTED;
  }
#define DF3D_SYM(field, name)                                                   \
  *reinterpret_cast<void**>(&j->field) = dlsym(j->dl, name);                    \
  if (!j->field) {                                                              \
    set_error("df3d_jpeg_create: symbol %s missing in nvJPEG", name);           \
    df3d_jpeg_destroy(j);                                                       \
    return DF3D_EUNSUPPORTED;                                                   \
  }
  DF3D_SYM(create_simple, "nvjpegCreateSimple")
  DF3D_SYM(destroy, "nvjpegDestroy")
  DF3D_SYM(state_create, "nvjpegJpegStateCreate")
  DF3D_SYM(state_destroy, "nvjpegJpegStateDestroy")
  DF3D_SYM(get_info, "nvjpegGetImageInfo")
  DF3D_SYM(decode, "nvjpegDecode")
#undef DF3D_SYM
  nvjpegStatus_t st = j->create_simple(&j->handle);
  if (st == NVJPEG_STATUS_SUCCESS) st = j->state_create(j->handle, &j->state);
  if (st != NVJPEG_STATUS_SUCCESS) {
    set_error("df3d_jpeg_create: nvJPEG initialisation failed (status %d)", (int)st);
    df3d_jpeg_destroy(j);
    return DF3D_ECUDA;
  }
  *out = j;
  return DF3D_OK;
}

extern "C" int df3d_jpeg_info(df3d_jpeg* j, const uint8_t* data, size_t len, int* width, int* height) {
  using namespace df3d;
  DF3D_REQUIRE(j && data && width && height, DF3D_EINVAL, "df3d_jpeg_info: null pointer");
  int ncomp = 0, w[NVJPEG_MAX_COMPONENT] = {}, h[NVJPEG_MAX_COMPONENT] = {};
  nvjpegChromaSubsampling_t ss;
  const nvjpegStatus_t st = j->get_info(j->handle, data, len, &ncomp, &ss, w, h);
  DF3D_REQUIRE(st == NVJPEG_STATUS_SUCCESS, DF3D_EINVAL, "df3d_jpeg_info: not a decodable JPEG stream (nvJPEG status %d)", (int)st);
  *width = w[0];
  *height = h[0];
  return DF3D_OK;
}

// data / lens: HOST arrays of n compressed streams; dst_dev: (n, H, W) uint8 on the device.  Every image must
// be H x W.  Work is submitted on `stream`; nvJPEG's Huffman stage runs on the calling host thread.
extern "C" int df3d_jpeg_decode_gray(df3d_jpeg* j, const uint8_t* const* data, const size_t* lens, int n, uint8_t* dst_dev,
                                     int H, int W, void* stream) {
  using namespace df3d;
  DF3D_REQUIRE(j && (n == 0 || (data && lens && dst_dev)), DF3D_EINVAL, "df3d_jpeg_decode_gray: null pointer");
  DF3D_REQUIRE(n >= 0 && H > 0 && W > 0, DF3D_EINVAL, "df3d_jpeg_decode_gray: bad shape");
  for (int i = 0; i < n; ++i) {
    int w = 0, h = 0;
    if (int e = df3d_jpeg_info(j, data[i], lens[i], &w, &h)) return e;
    DF3D_REQUIRE(w == W && h == H, DF3D_EINVAL, "df3d_jpeg_decode_gray: image %d is %d x %d, expected %d x %d", i, w, h, W, H);
    nvjpegImage_t img = {};
    img.channel[0] = dst_dev + (size_t)i * H * W;
    img.pitch[0] = (size_t)W;
    const nvjpegStatus_t st = j->decode(j->handle, j->state, data[i], lens[i], NVJPEG_OUTPUT_Y, &img, static_cast<cudaStream_t>(stream));
    DF3D_REQUIRE(st == NVJPEG_STATUS_SUCCESS, DF3D_ECUDA, "df3d_jpeg_decode_gray: nvjpegDecode failed on image %d (status %d)", i, (int)st);
  }
  return DF3D_OK;
}
