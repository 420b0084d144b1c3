// Post-processing of the pose tracks on the device (fp64): procrustes registration of the triangulated skeleton
// and the reference's temporal filters.  Replaces
//   df3d/procrustes.py:51-263 (procrustes_seperate -> procrustes -> __procrustes) + df3d/plot_util.py:85-91
//       (normalize_pose_3d), run by Core.save (df3d/core.py:358) and Core.get_points3d (core.py:339-341);
//   df3d/signal_util.py:31-100, 103-132 (OneEuroFilter, filter_batch, filter_batch_2d), run by
//       Core.get_points3d (core.py:342);
//   df3d/signal_util.py:135-160 (smooth_pose2d), run by Core.smooth_points2d (core.py:286-296).
//
// Procrustes needs medians over ALL frames (bone lengths, all points, the alignment joints): 66 selection
// problems of T or 19 T values.  They run as one radix select per column (order-preserving 64-bit keys, eight
// 8-bit digits from the most significant, a 256-bin shared-memory histogram per pass); numpy's rule for an even
// count (mean of the two middle values) is reproduced.  The filters are recurrences in time: one thread per
// (joint, coordinate) track walks the frames; consecutive threads read consecutive doubles of a frame.
#include <cmath>
#include <vector>

#include "common.cuh"

namespace df3d {

constexpr int kHalf = 19;          // joints per body half (df3d/skeleton_fly.py: 38 joints)
constexpr int kLegs = 3, kBones = 12, kAlign = 6;
constexpr int kColsPerHalf = kBones + 3 + 3 * kAlign;  // 33 medians per half
constexpr int kProcParams = 16;    // per half: scale, median[3], Q[9] (row-major), c[3]

__device__ __forceinline__ int align_joint(int a) {  // BODY_COXA and COXA_FEMUR of the three legs (procrustes.py:55)
  return (a >> 1) * 5 + (a & 1);
}

// column layout of one half inside the scratch buffer (doubles): 12 bone columns of T, 3 coordinate columns of
// 19 T, 18 alignment columns of T
__host__ __device__ inline size_t half_doubles(int T) { return (size_t)(kBones + 3 * kAlign) * T + (size_t)3 * kHalf * T; }
__host__ __device__ inline size_t col_offset(int col, int T) {
  if (col < kBones) return (size_t)col * T;
  if (col < kBones + 3) return (size_t)kBones * T + (size_t)(col - kBones) * kHalf * T;
  return (size_t)kBones * T + (size_t)3 * kHalf * T + (size_t)(col - kBones - 3) * T;
}
__host__ __device__ inline int col_length(int col, int T) { return (col >= kBones && col < kBones + 3) ? kHalf * T : T; }

__global__ void proc_fill_kernel(const double* __restrict__ pts, int T, double* __restrict__ scratch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * T) return;
  const int h = i / T, t = i - h * T;
  const double* p = pts + ((size_t)t * 2 * kHalf + (size_t)h * kHalf) * 3;
  double* s = scratch + (size_t)h * half_doubles(T);
  for (int leg = 0; leg < kLegs; ++leg)
    for (int b = 0; b < 4; ++b) {
      const double* a0 = p + (leg * 5 + b) * 3;
      const double dx = a0[3] - a0[0], dy = a0[4] - a0[1], dz = a0[5] - a0[2];
      s[col_offset(leg * 4 + b, T) + t] = sqrt(dx * dx + dy * dy + dz * dz);
    }
  for (int j = 0; j < kHalf; ++j)
    for (int d = 0; d < 3; ++d) s[col_offset(kBones + d, T) + (size_t)t * kHalf + j] = p[j * 3 + d];
  for (int a = 0; a < kAlign; ++a)
    for (int d = 0; d < 3; ++d) s[col_offset(kBones + 3 + a * 3 + d, T) + t] = p[align_joint(a) * 3 + d];
}

__device__ __forceinline__ unsigned long long dkey(double v) {  // order-preserving map double -> u64
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// k-th smallest (0-based) key of a column: the whole CTA cooperates
__device__ unsigned long long radix_select(const double* __restrict__ col, int n, int k, unsigned int* hist, unsigned long long* s_prefix,
                                           int* s_k) {
  unsigned long long prefix = 0, mask = 0;
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned long long key = dkey(col[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffull], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int kk = k, b = 0;
      for (; b < 256; ++b) {
        if (kk < (int)hist[b]) break;
        kk -= (int)hist[b];
      }
      *s_prefix = prefix | ((unsigned long long)b << shift);
      *s_k = kk;
    }
    __syncthreads();
    prefix = *s_prefix;
    k = *s_k;
    mask |= 0xffull << shift;
    __syncthreads();
  }
  return prefix;
}

// one CTA per column: numpy.median (mean of the two middle values for an even count)
__global__ void __launch_bounds__(1024) proc_median_kernel(const double* __restrict__ scratch, int T, double* __restrict__ medians) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_k;
  const int h = blockIdx.x / kColsPerHalf, col = blockIdx.x % kColsPerHalf;
  const double* c = scratch + (size_t)h * half_doubles(T) + col_offset(col, T);
  const int n = col_length(col, T);
  const double lo = dkey_inv(radix_select(c, n, (n - 1) / 2, hist, &s_prefix, &s_k));
  double med = lo;
  if ((n & 1) == 0) {
    const double hi = dkey_inv(radix_select(c, n, n / 2, hist, &s_prefix, &s_k));
    med = (lo + hi) / 2.0;
  }
  if (threadIdx.x == 0) medians[blockIdx.x] = med;
}

// right singular vectors V and A V of a 3x3 matrix (one-sided Jacobi), then Q = V U^T
__device__ void kabsch_q(const double (&A_in)[3][3], double (&Q)[3][3]) {
  double G[3][3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      G[i][j] = A_in[i][j];
      V[i][j] = i == j ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < 3; ++i) {
          alpha += G[i][p] * G[i][p];
          beta += G[i][q] * G[i][q];
          gamma += G[i][p] * G[i][q];
        }
        if (gamma != 0.0 && fabs(gamma) > 1e-17 * sqrt(alpha * beta)) {
          rotated = true;
          const double zeta = (beta - alpha) / (2.0 * gamma);
          const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
          for (int i = 0; i < 3; ++i) {
            const double gp = G[i][p], gq = G[i][q];
            G[i][p] = c * gp - s * gq;
            G[i][q] = s * gp + c * gq;
            const double vp = V[i][p], vq = V[i][q];
            V[i][p] = c * vp - s * vq;
            V[i][q] = s * vp + c * vq;
          }
        }
      }
    if (!rotated) break;
  }
  // G = U S (columns): Q = V U^T = sum_k v_k u_k^T
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Q[i][j] = 0.0;
  for (int k = 0; k < 3; ++k) {
    double nrm = 0;
    for (int i = 0; i < 3; ++i) nrm += G[i][k] * G[i][k];
    nrm = sqrt(nrm);
    if (!(nrm > 0.0)) continue;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Q[i][j] += V[i][k] * (G[j][k] / nrm);
  }
}

// one thread per half: scale, centre, rigid fit (procrustes.py:92-151 / 154-263 with scaling=False, reflection='best')
__global__ void proc_fit_kernel(const double* __restrict__ medians, const double* __restrict__ tmpl, double* __restrict__ params) {
  const int h = threadIdx.x;
  if (h >= 2) return;
  const double* m = medians + h * kColsPerHalf;
  const double* tm = tmpl + h * (kBones + 3 * kAlign);  // template: 12 median bone lengths, 18 median alignment coordinates
  double ratio[kBones];
  for (int k = 0; k < kBones; ++k) ratio[k] = tm[k] / m[k];
  for (int i = 1; i < kBones; ++i) {  // insertion sort, then numpy's even-count median
    const double v = ratio[i];
    int j = i - 1;
    while (j >= 0 && ratio[j] > v) {
      ratio[j + 1] = ratio[j];
      --j;
    }
    ratio[j + 1] = v;
  }
  const double scale = (ratio[kBones / 2 - 1] + ratio[kBones / 2]) / 2.0;
  const double med[3] = {m[kBones + 0], m[kBones + 1], m[kBones + 2]};
  double P[kAlign][3], Tt[kAlign][3], mp[3] = {0, 0, 0}, mt[3] = {0, 0, 0};
  for (int a = 0; a < kAlign; ++a)
    for (int d = 0; d < 3; ++d) {
      P[a][d] = (m[kBones + 3 + a * 3 + d] - med[d]) * scale;
      Tt[a][d] = tm[kBones + a * 3 + d];
      mp[d] += P[a][d];
      mt[d] += Tt[a][d];
    }
  for (int d = 0; d < 3; ++d) {
    mp[d] /= kAlign;
    mt[d] /= kAlign;
  }
  double np_ = 0, nt = 0;
  for (int a = 0; a < kAlign; ++a)
    for (int d = 0; d < 3; ++d) {
      P[a][d] -= mp[d];
      Tt[a][d] -= mt[d];
      np_ += P[a][d] * P[a][d];
      nt += Tt[a][d] * Tt[a][d];
    }
  np_ = sqrt(np_);
  nt = sqrt(nt);
  double A[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0;
      for (int a = 0; a < kAlign; ++a) acc += (Tt[a][i] / nt) * (P[a][j] / np_);
      A[i][j] = acc;  // t0^T s0
    }
  double Q[3][3];
  kabsch_q(A, Q);
  double* o = params + h * kProcParams;
  o[0] = scale;
  for (int d = 0; d < 3; ++d) o[1 + d] = med[d];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o[4 + i * 3 + j] = Q[i][j];
  for (int j = 0; j < 3; ++j) o[13 + j] = mt[j] - (mp[0] * Q[0][j] + mp[1] * Q[1][j] + mp[2] * Q[2][j]);
}

__global__ void proc_apply_kernel(const double* __restrict__ pts, int n_joints, const double* __restrict__ params, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_joints) return;
  const int h = (i % (2 * kHalf)) / kHalf;
  const double* o = params + h * kProcParams;
  const double c0 = (pts[(size_t)i * 3 + 0] - o[1]) * o[0];
  const double c1 = (pts[(size_t)i * 3 + 1] - o[2]) * o[0];
  const double c2 = (pts[(size_t)i * 3 + 2] - o[3]) * o[0];
  for (int j = 0; j < 3; ++j) out[(size_t)i * 3 + j] = c0 * o[4 + j] + c1 * o[7 + j] + c2 * o[10 + j] + o[13 + j];
}

// ------------------------------------------------------------------------------------------------ filters
// OneEuroFilter of the reference, one thread per track, bit-exact: every product / sum is a separate IEEE
// operation like in the Python source (no fused multiply-add)
__global__ void one_euro_kernel(const double* __restrict__ pts, int T, int N, double freq0, double mincutoff, double beta,
                                double dcutoff, int t_first, double* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const double two_pi = 2.0 * 3.141592653589793;
  double freq = freq0;
  double last = 0.0, x_y = 0.0, x_s = 0.0, dx_s = 0.0;
  bool have_last = false, have_x = false, have_dx = false;
  for (int i = 0; i < T; ++i) {
    const double ts = __dmul_rn((double)(i + t_first), 0.1);
    if (have_last && last != 0.0 && ts != 0.0) freq = __ddiv_rn(1.0, __dsub_rn(ts, last));  // `if self.__lasttime and timestamp`
    last = ts;
    have_last = true;
    const double x = pts[(size_t)i * N + n];
    const double dx = have_x ? __dmul_rn(__dsub_rn(x, x_y), freq) : 0.0;
    const double te = __ddiv_rn(1.0, freq);
    const double a_d = __ddiv_rn(1.0, __dadd_rn(1.0, __ddiv_rn(__ddiv_rn(1.0, __dmul_rn(two_pi, dcutoff)), te)));
    const double edx = have_dx ? __dadd_rn(__dmul_rn(a_d, dx), __dmul_rn(__dsub_rn(1.0, a_d), dx_s)) : dx;
    dx_s = edx;
    have_dx = true;
    const double cutoff = __dadd_rn(mincutoff, __dmul_rn(beta, fabs(edx)));
    const double a = __ddiv_rn(1.0, __dadd_rn(1.0, __ddiv_rn(__ddiv_rn(1.0, __dmul_rn(two_pi, cutoff)), te)));
    const double s = have_x ? __dadd_rn(__dmul_rn(a, x), __dmul_rn(__dsub_rn(1.0, a), x_s)) : x;
    x_y = x;
    x_s = s;
    have_x = true;
    out[(size_t)i * N + n] = s;
  }
}

constexpr int kMaxGaussTaps = 129;
struct GaussTaps {
  double w[kMaxGaussTaps];
  int radius;
};

// smooth_pose2d: one thread per (frame, track)
__global__ void smooth_pose2d_kernel(const double* __restrict__ pts, int T, int N, int window, double std_thr, GaussTaps g,
                                     double* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)T * N) return;
  const int t = (int)(i / N), n = (int)(i % N);
  const int half = window / 2;
  auto seg = [&](int k) {  // padded[t + pad - half + k]: the edge frames repeated
    int f = t - half + k;
    f = f < 0 ? 0 : (f > T - 1 ? T - 1 : f);
    return pts[(size_t)f * N + n];
  };
  double mean = 0.0;
  for (int k = 0; k < window; ++k) mean += seg(k);
  mean /= window;
  double var = 0.0;
  for (int k = 0; k < window; ++k) {
    const double d = seg(k) - mean;
    var += d * d;
  }
  const double sd = sqrt(var / window);
  double v = seg(half);  // sigma 0.1: radius 0, the centre tap itself
  if (sd < std_thr) {
    v = 0.0;
    for (int k = -g.radius; k <= g.radius; ++k) {
      int idx = half + k;
      idx = idx < 0 ? 0 : (idx > window - 1 ? window - 1 : idx);  // mode='nearest' on the window
      v += g.w[k + g.radius] * seg(idx);
    }
  }
  out[i] = v;
}

}  // namespace df3d

using namespace df3d;

extern "C" size_t df3d_procrustes_workspace_bytes(int T) {
  if (T < 1) return 0;
  return (2 * half_doubles(T) + 2 * kColsPerHalf + 2 * kProcParams) * sizeof(double) + 256;
}

extern "C" int df3d_procrustes(const double* pts3d_dev, int T, int J, const double* template_medians_dev, double* out_dev,
                               void* workspace_dev, size_t workspace_bytes, void* stream) {
  DF3D_REQUIRE(pts3d_dev && template_medians_dev && out_dev && workspace_dev, DF3D_EINVAL, "df3d_procrustes: null pointer");
  DF3D_REQUIRE(J == 2 * kHalf, DF3D_EUNSUPPORTED, "df3d_procrustes: the skeleton has %d joints, got %d", 2 * kHalf, J);
  DF3D_REQUIRE(T >= 1 && (long long)T * kHalf < (1ll << 31), DF3D_EINVAL, "df3d_procrustes: bad T");
  DF3D_REQUIRE((reinterpret_cast<uintptr_t>(workspace_dev) & 255) == 0, DF3D_EINVAL, "df3d_procrustes: workspace must be 256-byte aligned");
  DF3D_REQUIRE(workspace_bytes >= df3d_procrustes_workspace_bytes(T) - 256, DF3D_ENOMEM, "df3d_procrustes: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* scratch = static_cast<double*>(workspace_dev);
  double* medians = scratch + 2 * half_doubles(T);
  double* params = medians + 2 * kColsPerHalf;
  proc_fill_kernel<<<ceil_div(2 * T, 128), 128, 0, s>>>(pts3d_dev, T, scratch);
  DF3D_LAUNCH_CHECK("proc_fill_kernel");
  proc_median_kernel<<<2 * kColsPerHalf, 1024, 0, s>>>(scratch, T, medians);
  DF3D_LAUNCH_CHECK("proc_median_kernel");
  proc_fit_kernel<<<1, 32, 0, s>>>(medians, template_medians_dev, params);
  DF3D_LAUNCH_CHECK("proc_fit_kernel");
  proc_apply_kernel<<<ceil_div(T * J, 256), 256, 0, s>>>(pts3d_dev, T * J, params, out_dev);
  DF3D_LAUNCH_CHECK("proc_apply_kernel");
  return DF3D_OK;
}

extern "C" int df3d_one_euro_filter(const double* pts_dev, int T, int n_tracks, double freq, double mincutoff, double beta,
                                    double dcutoff, int t_first, double* out_dev, void* stream) {
  DF3D_REQUIRE(pts_dev && out_dev, DF3D_EINVAL, "df3d_one_euro_filter: null pointer");
  DF3D_REQUIRE(T >= 0 && n_tracks >= 0, DF3D_EINVAL, "df3d_one_euro_filter: bad size");
  DF3D_REQUIRE(freq > 0 && mincutoff > 0 && dcutoff > 0, DF3D_EINVAL, "df3d_one_euro_filter: freq, mincutoff and dcutoff should be > 0");
  DF3D_REQUIRE(t_first == 0 || t_first == 1, DF3D_EINVAL, "df3d_one_euro_filter: t_first must be 0 or 1");
  if (T == 0 || n_tracks == 0) return DF3D_OK;
  one_euro_kernel<<<ceil_div(n_tracks, 64), 64, 0, static_cast<cudaStream_t>(stream)>>>(pts_dev, T, n_tracks, freq, mincutoff, beta,
                                                                                     dcutoff, t_first, out_dev);
  DF3D_LAUNCH_CHECK("one_euro_kernel");
  return DF3D_OK;
}

extern "C" int df3d_smooth_pose2d(const double* pts_dev, int T, int n_tracks, int window_size, double std_thr, double* out_dev,
                                  void* stream) {
  DF3D_REQUIRE(pts_dev && out_dev, DF3D_EINVAL, "df3d_smooth_pose2d: null pointer");
  DF3D_REQUIRE(T >= 0 && n_tracks >= 0 && window_size >= 2 && window_size <= 64 && window_size % 2 == 0, DF3D_EINVAL,
               "df3d_smooth_pose2d: bad size");
  if (T == 0 || n_tracks == 0) return DF3D_OK;
  // scipy.ndimage.gaussian_filter1d(sigma=7, truncate=4): radius 28, weights exp(-0.5 x^2 / sigma^2) normalised
  GaussTaps g;
  const double sigma = 7.0;
  g.radius = (int)(4.0 * sigma + 0.5);
  double sum = 0.0;
  for (int k = -g.radius; k <= g.radius; ++k) {
    g.w[k + g.radius] = std::exp(-0.5 / (sigma * sigma) * (double)(k * k));
    sum += g.w[k + g.radius];
  }
  for (int k = 0; k <= 2 * g.radius; ++k) g.w[k] /= sum;
  const size_t total = (size_t)T * n_tracks;
  smooth_pose2d_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(pts_dev, T, n_tracks, window_size,
                                                                                                    std_thr, g, out_dev);
  DF3D_LAUNCH_CHECK("smooth_pose2d_kernel");
  return DF3D_OK;
}
