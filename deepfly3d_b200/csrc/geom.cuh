// Camera model shared by triangulation and bundle adjustment (fp64, device side).
// Pin-hole without distortion, extrinsics as Rodrigues vector + translation -- the
// parameterisation pyba hands to SciPy.
#pragma once
#include "common.cuh"

namespace df3d {

__device__ __forceinline__ void rodrigues_dev(const double* r, double (&R)[3][3]) {
  const double th2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  const double th = sqrt(th2);
  if (th < 1e-300) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) R[i][j] = (i == j) ? 1.0 : 0.0;
    return;
  }
  const double k[3] = {r[0] / th, r[1] / th, r[2] / th};
  double s, c;
  sincos(th, &s, &c);
  const double v = 1.0 - c;
  R[0][0] = c + v * k[0] * k[0];
  R[0][1] = v * k[0] * k[1] - s * k[2];
  R[0][2] = v * k[0] * k[2] + s * k[1];
  R[1][0] = v * k[1] * k[0] + s * k[2];
  R[1][1] = c + v * k[1] * k[1];
  R[1][2] = v * k[1] * k[2] - s * k[0];
  R[2][0] = v * k[2] * k[0] - s * k[1];
  R[2][1] = v * k[2] * k[1] + s * k[0];
  R[2][2] = c + v * k[2] * k[2];
}

// Per-camera constants staged in shared memory: R (9), M (9), t (3), fx fy cx cy (4).
// M = (r r^T + (R^T - I)[r]x) / |r|^2 gives d(R X)/dr = -R [X]x M  (Gallego & Yezzi 2015).
constexpr int kCamStride = 25;

__device__ __forceinline__ void stage_camera(const double* rt, const double* intr4, double* out) {
  double R[3][3];
  rodrigues_dev(rt, R);
  const double r0 = rt[0], r1 = rt[1], r2 = rt[2];
  const double th2 = r0 * r0 + r1 * r1 + r2 * r2;
  double M[3][3];
  if (th2 < 1e-24) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) M[i][j] = (i == j) ? 1.0 : 0.0;
  } else {
    const double rv[3] = {r0, r1, r2};
    const double K[3][3] = {{0.0, -r2, r1}, {r2, 0.0, -r0}, {-r1, r0, 0.0}};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double acc = rv[i] * rv[j];
#pragma unroll
        for (int k = 0; k < 3; ++k) acc += (R[k][i] - (k == i ? 1.0 : 0.0)) * K[k][j];
        M[i][j] = acc / th2;
      }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      out[i * 3 + j] = R[i][j];
      out[9 + i * 3 + j] = M[i][j];
    }
  out[18] = rt[3];
  out[19] = rt[4];
  out[20] = rt[5];
  out[21] = intr4[0];
  out[22] = intr4[1];
  out[23] = intr4[2];
  out[24] = intr4[3];
}

// residual only
__device__ __forceinline__ void project_residual(const double* cam, const double (&X)[3], double ox, double oy,
                                                 double (&r)[2]) {
  const double xc = cam[0] * X[0] + cam[1] * X[1] + cam[2] * X[2] + cam[18];
  const double yc = cam[3] * X[0] + cam[4] * X[1] + cam[5] * X[2] + cam[19];
  const double zc = cam[6] * X[0] + cam[7] * X[1] + cam[8] * X[2] + cam[20];
  const double iz = 1.0 / zc;
  r[0] = cam[21] * (xc * iz) + cam[23] - ox;
  r[1] = cam[22] * (yc * iz) + cam[24] - oy;
}

// residual + Jacobians: Jc (2x6: d/d rvec, d/d tvec), Jp (2x3: d/dX)
__device__ __forceinline__ void project_jacobian(const double* cam, const double (&X)[3], double ox, double oy,
                                                 double (&r)[2], double (&Jc)[2][6], double (&Jp)[2][3]) {
  const double* R = cam;
  const double* M = cam + 9;
  const double xc = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + cam[18];
  const double yc = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + cam[19];
  const double zc = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + cam[20];
  const double iz = 1.0 / zc;
  const double x = xc * iz, y = yc * iz;
  const double fx = cam[21], fy = cam[22];
  r[0] = fx * x + cam[23] - ox;
  r[1] = fy * y + cam[24] - oy;
  // d(u,v)/dXc
  const double dp[2][3] = {{fx * iz, 0.0, -fx * x * iz}, {0.0, fy * iz, -fy * y * iz}};
#pragma unroll
  for (int a = 0; a < 2; ++a) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      Jp[a][j] = dp[a][0] * R[0 * 3 + j] + dp[a][1] * R[1 * 3 + j] + dp[a][2] * R[2 * 3 + j];
      Jc[a][3 + j] = dp[a][j];
    }
  }
  // G = R [X]x ; dXc/dr = -G M ; Jr = dp * dXc/dr = -(dp R) [X]x M = -Jp [X]x M
  double JX[2][3];  // Jp [X]x
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    JX[a][0] = Jp[a][1] * X[2] - Jp[a][2] * X[1];
    JX[a][1] = Jp[a][2] * X[0] - Jp[a][0] * X[2];
    JX[a][2] = Jp[a][0] * X[1] - Jp[a][1] * X[0];
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Jc[a][j] = -(JX[a][0] * M[0 * 3 + j] + JX[a][1] * M[1 * 3 + j] + JX[a][2] * M[2 * 3 + j]);
}

}  // namespace df3d
