// 19 -> 38 joint packing of the decoded arg-max indices.
// Replaces df3d/core.py:187-203 (Core.pose2d_estimation) and the pixel scaling of core.py:247.
// Pure elementwise; one thread per output (camera, frame, joint).
#include "common.cuh"

namespace df3d {

struct PackParams {
  int C, T, K, Hh, Wh, img_w, img_h;
  int slot_of_cam[DF3D_MAX_CAMS];  // position of camera id in camera_ordering (inverse permutation)
};

__global__ void pack_points2d_kernel(const int32_t* __restrict__ idx, PackParams p,
                                     double* __restrict__ points2d, double* __restrict__ pts_xy) {
  const int J = 2 * p.K;
  const long long n = (long long)p.C * p.T * J;
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int j = (int)(g % J);
  const int t = (int)((g / J) % p.T);
  const int c = (int)(g / ((long long)J * p.T));
  const int slot = p.slot_of_cam[c];  // camera c sits at camera_ordering[slot]

  double row = 0.0, col = 0.0;
  // core.py:190: ordering[:3] fill joints [0,K) ; core.py:191: ordering[4:] fill joints [K,2K)
  bool filled = (slot >= 0 && slot < 3 && j < p.K) || (slot >= 4 && j >= p.K);
  // core.py:194-195: ordering[2] keeps joints < 15 only, ordering[4] keeps joints < K+15 only
  if (slot == 2 && j >= 15) filled = false;
  if (slot == 4 && j >= p.K + 15) filled = false;
  if (filled) {
    const int k = j < p.K ? j : j - p.K;
    const int flat = idx[((size_t)c * p.T + t) * p.K + k];
    row = (double)(flat / p.Wh) / (double)p.Hh;
    col = (double)(flat % p.Wh) / (double)p.Wh;
  }
  // core.py:198-199: un-flip the column of the three left-hand-side cameras (all 38 joints,
  // so blanked joints become (0, 1))
  if (slot >= 4) col = 1.0 - col;
  points2d[g * 2 + 0] = row;
  points2d[g * 2 + 1] = col;
  if (pts_xy) {
    // core.py:247: points2d * image_shape[::-1] = (row*H, col*W); pyba uses (x, y) = (col, row)
    pts_xy[g * 2 + 0] = col * (double)p.img_w;
    pts_xy[g * 2 + 1] = row * (double)p.img_h;
  }
}

}  // namespace df3d

extern "C" int df3d_pack_points2d(const int32_t* idx_dev, int C, int T, int K, int Hh, int Wh,
                                  const int* camera_ordering, int img_w, int img_h,
                                  double* points2d_dev, double* pts_xy_dev, void* stream) {
  using namespace df3d;
  DF3D_REQUIRE(idx_dev && camera_ordering && points2d_dev, DF3D_EINVAL, "df3d_pack_points2d: null pointer");
  DF3D_REQUIRE(C == 7, DF3D_EUNSUPPORTED, "df3d_pack_points2d: the reference packing is defined for 7 cameras, got %d", C);
  DF3D_REQUIRE(T >= 0 && K >= 15 && Hh > 0 && Wh > 0 && img_w > 0 && img_h > 0, DF3D_EINVAL, "df3d_pack_points2d: bad shape");
  PackParams p{C, T, K, Hh, Wh, img_w, img_h, {}};
  bool seen[DF3D_MAX_CAMS] = {};
  for (int s = 0; s < C; ++s) {
    int cam = camera_ordering[s];
    DF3D_REQUIRE(cam >= 0 && cam < C && !seen[cam], DF3D_EINVAL, "df3d_pack_points2d: camera_ordering is not a permutation");
    seen[cam] = true;
    p.slot_of_cam[cam] = s;
  }
  if (T == 0) return DF3D_OK;
  const long long n = (long long)C * T * 2 * K;
  const int threads = 256;
  pack_points2d_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      idx_dev, p, points2d_dev, pts_xy_dev);
  DF3D_LAUNCH_CHECK("pack_points2d_kernel");
  return DF3D_OK;
}
