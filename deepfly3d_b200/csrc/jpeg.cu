// JPEG decode on the device (luminance plane) through nvJPEG -- the first half of the ingest row
// (the reference expands videos to camera_C_img_I.jpg, df3d/core.py:446-459, and
// df2d reads them back on the host behind inference_folder, core.py:177-185).  nvJPEG is library code: this
// file only binds it.  It is loaded lazily with dlopen so that libdf3d_b200.so itself never depends on it --
// the default loader path (host decode, bit-identical to the reference's libjpeg read) needs nothing here.
// Backends, probed in this order at df3d_jpeg_create (df3d_jpeg_backend reports the one in use):
//   3  NVJPEG_BACKEND_HARDWARE   the GPU's NVJPG engines, nvjpegDecodeBatched over the whole frame block: no
//                                host work per image beyond parsing the headers
//   2  NVJPEG_BACKEND_GPU_HYBRID only on request: batched decode with the Huffman stage on the SMs
//   0  NVJPEG_BACKEND_DEFAULT    nvjpegDecode per image (Huffman stage on the calling host thread)
// Parity: nvJPEG's inverse DCT is not libjpeg-turbo's; decoded frames differ by a few grey levels
// (tests/test_gpu_ingest.py states the measured bound), so this path is opt-in.
#include <dlfcn.h>
#include <nvjpeg.h>

#include "common.cuh"

struct df3d_jpeg {
  void* dl = nullptr;
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  int backend = 0;       // nvjpegBackend_t in use
  int batch_ready = 0;   // batch size the state was initialised for (batched backends)
  nvjpegStatus_t (*create_simple)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*create_ex)(nvjpegBackend_t, nvjpegDevAllocator_t*, nvjpegPinnedAllocator_t*, unsigned int, nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*batched_init)(nvjpegHandle_t, nvjpegJpegState_t, int, int, nvjpegOutputFormat_t) = nullptr;
  nvjpegStatus_t (*batched)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char* const*, const size_t*, nvjpegImage_t*,
                            cudaStream_t) = nullptr;
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

// backend: -1 = probe (hardware engines first), else the nvjpegBackend_t to use (0 default, 3 hardware)
extern "C" int df3d_jpeg_create_backend(df3d_jpeg** out, int backend) {
  using namespace df3d;
  DF3D_REQUIRE(out, DF3D_EINVAL, "df3d_jpeg_create: null pointer");
  DF3D_REQUIRE(backend == -1 || backend == 0 || backend == 2 || backend == 3, DF3D_EINVAL, "df3d_jpeg_create: backend %d (want -1, 0, 2 or 3)",
               backend);
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
  DF3D_SYM(create_ex, "nvjpegCreateEx")
  DF3D_SYM(batched_init, "nvjpegDecodeBatchedInitialize")
  DF3D_SYM(batched, "nvjpegDecodeBatched")
#undef DF3D_SYM
  nvjpegStatus_t st = NVJPEG_STATUS_NOT_INITIALIZED;
  if (backend == -1 || backend == 3) {  // NVJPG engines: refused with ARCH_MISMATCH where the GPU / driver has none
    st = j->create_ex(NVJPEG_BACKEND_HARDWARE, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &j->handle);
    if (st == NVJPEG_STATUS_SUCCESS) {
      j->backend = 3;
    } else {
      j->handle = nullptr;
      if (backend == 3) {
        set_error("df3d_jpeg_create: no hardware JPEG backend here (nvJPEG status %d)", (int)st);
        df3d_jpeg_destroy(j);
        return DF3D_EUNSUPPORTED;
      }
    }
  }
  if (!j->handle && backend == 2) {  // Huffman stage on the SMs (nvJPEG takes that route for batches above 50 images)
    st = j->create_ex(NVJPEG_BACKEND_GPU_HYBRID, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &j->handle);
    if (st != NVJPEG_STATUS_SUCCESS) {
      set_error("df3d_jpeg_create: no GPU-hybrid JPEG backend here (nvJPEG status %d)", (int)st);
      j->handle = nullptr;
      df3d_jpeg_destroy(j);
      return DF3D_EUNSUPPORTED;
    }
    j->backend = 2;
  }
  if (!j->handle) st = j->create_simple(&j->handle);
  if (st == NVJPEG_STATUS_SUCCESS) st = j->state_create(j->handle, &j->state);
  if (st != NVJPEG_STATUS_SUCCESS) {
    set_error("df3d_jpeg_create: nvJPEG initialisation failed (status %d)", (int)st);
    df3d_jpeg_destroy(j);
    return DF3D_ECUDA;
  }
  *out = j;
  return DF3D_OK;
}

extern "C" int df3d_jpeg_create(df3d_jpeg** out) { return df3d_jpeg_create_backend(out, -1); }

extern "C" int df3d_jpeg_backend(const df3d_jpeg* j) { return j ? j->backend : -1; }

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
  }
  if (j->backend == 3 || j->backend == 2) {
    // the engines take the block in batches; the state is (re)initialised when the batch size changes
    constexpr int kBatch = 256;
    nvjpegImage_t imgs[kBatch];
    for (int i0 = 0; i0 < n; i0 += kBatch) {
      const int nb = n - i0 < kBatch ? n - i0 : kBatch;
      if (j->batch_ready != nb) {
        const nvjpegStatus_t st = j->batched_init(j->handle, j->state, nb, 1, NVJPEG_OUTPUT_Y);
        DF3D_REQUIRE(st == NVJPEG_STATUS_SUCCESS, DF3D_ECUDA, "df3d_jpeg_decode_gray: nvjpegDecodeBatchedInitialize(%d) failed (status %d)", nb, (int)st);
        j->batch_ready = nb;
      }
      for (int i = 0; i < nb; ++i) {
        imgs[i] = nvjpegImage_t{};
        imgs[i].channel[0] = dst_dev + (size_t)(i0 + i) * H * W;
        imgs[i].pitch[0] = (size_t)W;
      }
      const nvjpegStatus_t st = j->batched(j->handle, j->state, data + i0, lens + i0, imgs, static_cast<cudaStream_t>(stream));
      DF3D_REQUIRE(st == NVJPEG_STATUS_SUCCESS, DF3D_ECUDA, "df3d_jpeg_decode_gray: nvjpegDecodeBatched failed on images %d..%d (status %d)", i0,
                   i0 + nb - 1, (int)st);
    }
    return DF3D_OK;
  }
  for (int i = 0; i < n; ++i) {
    nvjpegImage_t img = {};
    img.channel[0] = dst_dev + (size_t)i * H * W;
    img.pitch[0] = (size_t)W;
    const nvjpegStatus_t st = j->decode(j->handle, j->state, data[i], lens[i], NVJPEG_OUTPUT_Y, &img, static_cast<cudaStream_t>(stream));
    DF3D_REQUIRE(st == NVJPEG_STATUS_SUCCESS, DF3D_ECUDA, "df3d_jpeg_decode_gray: nvjpegDecode failed on image %d (status %d)", i, (int)st);
  }
  return DF3D_OK;
}
