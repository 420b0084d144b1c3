// Batched multi-view DLT triangulation (fp64).
// Replaces pyba CameraNetwork.triangulate() (reference call site df3d/core.py:355).
//
// One thread per (frame, joint).  The (2V x 4) DLT system is never materialised: each row
// x*P[2]-P[0] / y*P[2]-P[1] is folded into a 4x4 upper-triangular factor with Givens rotations
// (A = QR has the same right singular vectors as R and does not square the condition number the
// way A^T A would -- entries span 1e4..1e6 because the lens is long, SURVEY hard part 5), then a
// one-sided Jacobi SVD of R gives the right singular vector of the smallest singular value.
// Loads are 128-bit and coalesced: consecutive threads read consecutive (x,y) pairs.
#include "common.cuh"
#include "geom.cuh"

namespace df3d {

__device__ __forceinline__ void givens_fold(double (&Rm)[4][4], double (&row)[4]) {
  // eliminate `row` into the upper-triangular Rm
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double a = Rm[k][k], b = row[k];
    if (b != 0.0) {
      const double r = hypot(a, b);
      const double c = a / r, s = b / r;
#pragma unroll
      for (int m = k; m < 4; ++m) {
        const double u = Rm[k][m], v = row[m];
        Rm[k][m] = c * u + s * v;
        row[m] = -s * u + c * v;
      }
    }
  }
}

// smallest right singular vector of the 4x4 matrix G (destroyed); one-sided Jacobi (Hestenes)
__device__ __forceinline__ void smallest_right_singular_vector(double (&G)[4][4], double (&x)[4]) {
  double V[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;

  for (int sweep = 0; sweep < 40; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          alpha += G[i][p] * G[i][p];
          beta += G[i][q] * G[i][q];
          gamma += G[i][p] * G[i][q];
        }
        if (gamma != 0.0 && fabs(gamma) > 1e-16 * sqrt(alpha * beta)) {
          rotated = true;
          const double zeta = (beta - alpha) / (2.0 * gamma);
          const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const double gp = G[i][p], gq = G[i][q];
            G[i][p] = c * gp - s * gq;
            G[i][q] = s * gp + c * gq;
            const double vp = V[i][p], vq = V[i][q];
            V[i][p] = c * vp - s * vq;
            V[i][q] = s * vp + c * vq;
          }
        }
      }
    }
    if (!rotated) break;
  }
  int kmin = 0;
  double nmin = INFINITY;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double n = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) n += G[i][j] * G[i][j];
    if (n < nmin) {
      nmin = n;
      kmin = j;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    // select without dynamic register indexing
    double v = V[i][0];
    if (kmin == 1) v = V[i][1];
    if (kmin == 2) v = V[i][2];
    if (kmin == 3) v = V[i][3];
    x[i] = v;
  }
}

// P is read from global memory (it lives on the device, e.g. right after bundle adjustment, so
// no host round trip is needed) and staged in shared memory once per CTA.
__global__ void __launch_bounds__(128)
triangulate_dlt_kernel(const double* __restrict__ Pg, const double2* __restrict__ pts_xy, int C, int TJ,
                          double* __restrict__ pts3d) {
  __shared__ double sP[DF3D_MAX_CAMS * 12];
  for (int i = threadIdx.x; i < C * 12; i += blockDim.x) sP[i] = Pg[i];
  __syncthreads();
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= TJ) return;
  double Rm[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Rm[i][j] = 0.0;
  int views = 0;
  for (int c = 0; c < C; ++c) {
    const double2 xy = __ldg(pts_xy + (size_t)c * TJ + g);
    if (xy.x != 0.0 && xy.y != 0.0) {
      ++views;
      const double* P = sP + c * 12;
      double r0[4], r1[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        r0[m] = xy.x * P[8 + m] - P[m];
        r1[m] = xy.y * P[8 + m] - P[4 + m];
      }
      givens_fold(Rm, r0);
      givens_fold(Rm, r1);
    }
  }
  double X[3] = {0.0, 0.0, 0.0};
  if (views >= 2) {
    double xh[4];
    smallest_right_singular_vector(Rm, xh);
    X[0] = xh[0] / xh[3];
    X[1] = xh[1] / xh[3];
    X[2] = xh[2] / xh[3];
  }
  pts3d[(size_t)g * 3 + 0] = X[0];
  pts3d[(size_t)g * 3 + 1] = X[1];
  pts3d[(size_t)g * 3 + 2] = X[2];
}

__global__ void projection_matrices_kernel(const double* __restrict__ cam_rt, const double* __restrict__ intr, int C,
                                           double* __restrict__ P, double* __restrict__ Rout) {
  const int c = threadIdx.x;
  if (c >= C) return;
  double R[3][3];
  rodrigues_dev(cam_rt + c * 6, R);
  const double* t = cam_rt + c * 6 + 3;
  const double fx = intr[c * 4 + 0], fy = intr[c * 4 + 1], cx = intr[c * 4 + 2], cy = intr[c * 4 + 3];
  double* Pc = P + c * 12;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const double a = m < 3 ? R[0][m] : t[0];
    const double b = m < 3 ? R[1][m] : t[1];
    const double d = m < 3 ? R[2][m] : t[2];
    Pc[m] = fx * a + cx * d;
    Pc[4 + m] = fy * b + cy * d;
    Pc[8 + m] = d;
  }
  if (Rout) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Rout[c * 9 + i * 3 + j] = R[i][j];
  }
}

}  // namespace df3d

extern "C" int df3d_triangulate_dlt(const double* P_dev, const double* pts_xy_dev, int C, int T, int J,
                                    double* pts3d_dev, void* stream) {
  using namespace df3d;
  DF3D_REQUIRE(P_dev && pts_xy_dev && pts3d_dev, DF3D_EINVAL, "df3d_triangulate_dlt: null pointer");
  DF3D_REQUIRE(C >= 1 && C <= DF3D_MAX_CAMS, DF3D_EINVAL, "df3d_triangulate_dlt: C must be in [1,%d]", DF3D_MAX_CAMS);
  DF3D_REQUIRE(T >= 0 && J >= 1 && (long long)T * J < (1ll << 31), DF3D_EINVAL, "df3d_triangulate_dlt: bad T/J");
  DF3D_REQUIRE((reinterpret_cast<uintptr_t>(pts_xy_dev) & 15) == 0, DF3D_EINVAL, "df3d_triangulate_dlt: pts_xy must be 16-byte aligned");
  if (T == 0) return DF3D_OK;
  const int TJ = T * J;
  const int threads = 128;
  triangulate_dlt_kernel<<<ceil_div(TJ, threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      P_dev, reinterpret_cast<const double2*>(pts_xy_dev), C, TJ, pts3d_dev);
  DF3D_LAUNCH_CHECK("triangulate_dlt_kernel");
  return DF3D_OK;
}

extern "C" int df3d_projection_matrices(const double* cam_rt_dev, const double* intr_dev, int C,
                                        double* P_dev, double* R_dev, void* stream) {
  using namespace df3d;
  DF3D_REQUIRE(cam_rt_dev && intr_dev && P_dev, DF3D_EINVAL, "df3d_projection_matrices: null pointer");
  DF3D_REQUIRE(C >= 1 && C <= DF3D_MAX_CAMS, DF3D_EINVAL, "df3d_projection_matrices: C must be in [1,%d]", DF3D_MAX_CAMS);
  projection_matrices_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(cam_rt_dev, intr_dev, C, P_dev, R_dev);
  DF3D_LAUNCH_CHECK("projection_matrices_kernel");
  return DF3D_OK;
}
