// Bundle adjustment of C camera extrinsics + all 3-D points (fp64, everything on the device).
// Replaces pyba CameraNetwork.bundle_adjust(update_intrinsic=False, update_distort=False)
// (reference call site df3d/core.py:249; recipe reconstructed in SURVEY.md Appendix B).
//
// Levenberg-Marquardt in the column-norm (Jacobi) scaled space SciPy's x_scale='jac' uses, with a
// small constant damping, solved through the Schur complement on the 6C camera unknowns:
//
//   ba_linearize : one thread per 3-D point.  Analytic Jacobian (Rodrigues + pin-hole), per-point
//                  V (3x3), g_p, per-observation W (6x3); contributions to U_c, g_c and to the
//                  reduced system  S~ = sum_j W_j M_j W_j^T,  b~ = sum_j W_j M_j g_pj  with
//                  M_j = D_p (D_p V_j D_p + lambda I)^-1 D_p  are reduced across the warp with
//                  shuffles, accumulated in per-warp shared-memory tiles (single writer, no
//                  atomics -> bitwise reproducible), summed per block, and the last block to finish
//                  adds the per-block partials in a fixed order.
//   ba_solve     : one CTA.  Camera scaling D_c from diag(U), forms
//                  D_c (U - S~) D_c + lambda I, Cholesky, candidate cameras.
//   ba_evaluate  : one thread per point.  Back-substitution for the point step and the candidate
//                  cost; same last-block reduction.
//   ba_decide    : accept / reject, lambda update, ftol test (dF < ftol * F like SciPy's TRF).
//
// The state lives in the caller's workspace, so a whole solve is a fixed sequence of launches
// with no host synchronisation; the `sys` and `cost` buffers are the only data a frame-sharded
// multi-GPU run has to all-reduce (see include/df3d_b200.h).
#include "common.cuh"
#include "geom.cuh"

namespace df3d {

constexpr int kBAThreads = 128;
constexpr int kBAWarps = kBAThreads / 32;
constexpr int kBAMaxBlocks = 148;

struct BAState {
  double lambda, F, F0, ftol;
  int iter, accepted, max_iters, done, status, n_obs, accept_flag, pad;
};

__host__ __device__ inline int sys_doubles(int C) { return 36 * C + 6 * C + 36 * C * C + 6 * C + 2; }
// offsets inside `sys`
__host__ __device__ inline int off_U(int) { return 0; }
__host__ __device__ inline int off_gc(int C) { return 36 * C; }
__host__ __device__ inline int off_S(int C) { return 42 * C; }
__host__ __device__ inline int off_b(int C) { return 42 * C + 36 * C * C; }
__host__ __device__ inline int off_cost(int C) { return 48 * C + 36 * C * C; }

struct BAWorkspace {  // carved out of the caller's buffer
  BAState* state;
  unsigned int* counters;  // [0] linearize ticket, [1] evaluate ticket
  double* cam;             // C*6 current
  double* cam_new;         // C*6 candidate
  double* dcam;            // C*6 unscaled camera step
  double* sinv_c;          // C*6 running max of camera column norms
  double* sinv_p;          // TJ*3 running max of point column norms
  double* X_new;           // TJ*3 candidate points
  double* partials;        // kBAMaxBlocks * sys_doubles(C)
  double* cost_partials;   // kBAMaxBlocks * 2
  double* sys_local;       // sys_doubles(C): used by the single-GPU driver
  double* cost_local;      // 2
};

static size_t ba_workspace_layout(int C, int T, int J, char* base, BAWorkspace* ws) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~size_t(255);
    return base ? base + o : nullptr;
  };
  const size_t TJ = (size_t)T * J;
  BAWorkspace w;
  w.state = reinterpret_cast<BAState*>(take(sizeof(BAState)));
  w.counters = reinterpret_cast<unsigned int*>(take(4 * sizeof(unsigned int)));
  w.cam = reinterpret_cast<double*>(take(C * 6 * 8));
  w.cam_new = reinterpret_cast<double*>(take(C * 6 * 8));
  w.dcam = reinterpret_cast<double*>(take(C * 6 * 8));
  w.sinv_c = reinterpret_cast<double*>(take(C * 6 * 8));
  w.sinv_p = reinterpret_cast<double*>(take(TJ * 3 * 8));
  w.X_new = reinterpret_cast<double*>(take(TJ * 3 * 8));
  w.partials = reinterpret_cast<double*>(take((size_t)kBAMaxBlocks * sys_doubles(C) * 8));
  w.cost_partials = reinterpret_cast<double*>(take((size_t)kBAMaxBlocks * 2 * 8));
  w.sys_local = reinterpret_cast<double*>(take((size_t)sys_doubles(C) * 8));
  w.cost_local = reinterpret_cast<double*>(take(2 * 8));
  if (ws) *ws = w;
  return off;
}

static int ba_grid(int TJ) {
  int g = ceil_div(TJ, kBAThreads);
  return g < 1 ? 1 : (g > kBAMaxBlocks ? kBAMaxBlocks : g);
}

// ---------------------------------------------------------------------------------------------
__global__ void ba_begin_kernel(const double* __restrict__ cam_rt, int C, int TJ, df3d_ba_opts opts, BAWorkspace ws) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g == 0) {
    BAState s;
    s.lambda = opts.lambda0;
    s.F = -1.0;
    s.F0 = -1.0;
    s.ftol = opts.ftol;
    s.iter = 0;
    s.accepted = 0;
    s.max_iters = opts.max_iters;
    s.done = 0;
    s.status = 0;
    s.n_obs = 0;
    s.accept_flag = 0;
    s.pad = 0;
    *ws.state = s;
    ws.counters[0] = ws.counters[1] = ws.counters[2] = ws.counters[3] = 0u;
  }
  if (g < C * 6) {
    ws.cam[g] = cam_rt[g];
    ws.cam_new[g] = cam_rt[g];
    ws.dcam[g] = 0.0;
    ws.sinv_c[g] = 0.0;
  }
  for (int i = g; i < TJ * 3; i += gridDim.x * blockDim.x) ws.sinv_p[i] = 0.0;
}

// 3x3 symmetric positive definite inverse (adjugate); returns false if not invertible
__device__ __forceinline__ bool inv3_sym(const double (&A)[3][3], double (&Ai)[3][3]) {
  const double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
  const double c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
  const double c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
  const double det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
  if (!(fabs(det) > 0.0)) return false;
  const double id = 1.0 / det;
  Ai[0][0] = c00 * id;
  Ai[0][1] = Ai[1][0] = c01 * id;
  Ai[0][2] = Ai[2][0] = c02 * id;
  Ai[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id;
  Ai[1][2] = Ai[2][1] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
  Ai[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
  return true;
}

// Last-block-done reduction of per-block partial vectors of length n (fixed summation order).
__device__ __forceinline__ void reduce_partials_last_block(const double* partials, int n, double* out,
                                                           unsigned int* ticket) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      double acc = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) acc += partials[(size_t)b * n + e];
      out[e] = acc;
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

// ---------------------------------------------------------------------------------------------
// dynamic shared memory: cams [C*kCamStride] | per-warp system tiles [kBAWarps * nsys]
__global__ void __launch_bounds__(kBAThreads)
ba_linearize_kernel(const double* __restrict__ intr, const double2* __restrict__ pts_xy,
                    const double* __restrict__ pts3d, int C, int TJ, BAWorkspace ws, double* __restrict__ sys_out) {
  extern __shared__ double smem[];
  if (ws.state->done) return;
  const double lambda = ws.state->lambda;
  const int nsys = sys_doubles(C);
  const int n6 = 6 * C;
  double* s_cam = smem;
  double* s_sys = smem + C * kCamStride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* my = s_sys + warp * nsys;  // this warp's accumulator tile

  if (threadIdx.x < C) stage_camera(ws.cam + threadIdx.x * 6, intr + threadIdx.x * 4, s_cam + threadIdx.x * kCamStride);
  for (int e = threadIdx.x; e < kBAWarps * nsys; e += kBAThreads) s_sys[e] = 0.0;
  __syncthreads();

  double* aU = my + off_U(C);
  double* agc = my + off_gc(C);
  double* aS = my + off_S(C);
  double* ab = my + off_b(C);
  double* acost = my + off_cost(C);

  for (int base = blockIdx.x * kBAThreads; base < TJ; base += gridDim.x * kBAThreads) {
    const int g = base + threadIdx.x;
    const bool valid = g < TJ;
    double X[3] = {0.0, 0.0, 0.0};
    if (valid) {
      X[0] = pts3d[(size_t)g * 3 + 0];
      X[1] = pts3d[(size_t)g * 3 + 1];
      X[2] = pts3d[(size_t)g * 3 + 2];
    }
    double V[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double gp[3] = {0, 0, 0};
    double W[DF3D_MAX_CAMS][6][3];
    unsigned mask = 0;
    double cost = 0.0;
    int nobs = 0;

    // pass 1: per-camera blocks, V, gp, W
    for (int c = 0; c < C; ++c) {
      bool vis = false;
      double2 xy = make_double2(0.0, 0.0);
      if (valid) {
        xy = __ldg(pts_xy + (size_t)c * TJ + g);
        vis = (xy.x != 0.0) && (xy.y != 0.0);
      }
      if (!__any_sync(0xffffffffu, vis)) continue;
      double r[2] = {0, 0}, Jc[2][6], Jp[2][3];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
#pragma unroll
        for (int i = 0; i < 6; ++i) Jc[a][i] = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) Jp[a][i] = 0.0;
      }
      if (vis) {
        project_jacobian(s_cam + c * kCamStride, X, xy.x, xy.y, r, Jc, Jp);
        mask |= 1u << c;
        ++nobs;
        cost += r[0] * r[0] + r[1] * r[1];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          gp[i] += Jp[0][i] * r[0] + Jp[1][i] * r[1];
#pragma unroll
          for (int j = 0; j < 3; ++j) V[i][j] += Jp[0][i] * Jp[0][j] + Jp[1][i] * Jp[1][j];
        }
      }
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) W[c][i][j] = Jc[0][i] * Jp[0][j] + Jc[1][i] * Jp[1][j];
      // U_c (upper triangle) and g_c: warp-shuffle reduce, lane 0 is the only writer of `my`
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double gv = warp_sum(Jc[0][i] * r[0] + Jc[1][i] * r[1]);
        if (lane == 0) agc[c * 6 + i] += gv;
#pragma unroll
        for (int j = i; j < 6; ++j) {
          const double uv = warp_sum(Jc[0][i] * Jc[0][j] + Jc[1][i] * Jc[1][j]);
          if (lane == 0) aU[c * 36 + i * 6 + j] += uv;
        }
      }
    }

    // point scaling (running max of column norms, SciPy compute_jac_scale) and M = D (DVD + lam I)^-1 D
    double Mm[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    if (valid) {
      double d[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double s = fmax(ws.sinv_p[(size_t)g * 3 + i], sqrt(V[i][i]));
        ws.sinv_p[(size_t)g * 3 + i] = s;
        d[i] = (s == 0.0) ? 1.0 : 1.0 / s;
      }
      if (mask) {
        double Vh[3][3], Vi[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) Vh[i][j] = d[i] * V[i][j] * d[j] + (i == j ? lambda : 0.0);
        if (inv3_sym(Vh, Vi)) {
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) Mm[i][j] = d[i] * Vi[i][j] * d[j];
        }
      }
    }

    // pass 2: reduced system, upper block triangle (a <= b)
    for (int a = 0; a < C; ++a) {
      const bool va = (mask >> a) & 1u;
      if (!__any_sync(0xffffffffu, va)) continue;
      double WM[6][3];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k)
          WM[i][k] = va ? (W[a][i][0] * Mm[0][k] + W[a][i][1] * Mm[1][k] + W[a][i][2] * Mm[2][k]) : 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double bv = warp_sum(WM[i][0] * gp[0] + WM[i][1] * gp[1] + WM[i][2] * gp[2]);
        if (lane == 0) ab[a * 6 + i] += bv;
      }
      for (int b = a; b < C; ++b) {
        const bool vb = va && ((mask >> b) & 1u);
        if (!__any_sync(0xffffffffu, vb)) continue;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            double v = vb ? (WM[i][0] * W[b][j][0] + WM[i][1] * W[b][j][1] + WM[i][2] * W[b][j][2]) : 0.0;
            v = warp_sum(v);
            if (lane == 0) aS[(a * 6 + i) * n6 + b * 6 + j] += v;
          }
      }
    }
    const double cs = warp_sum(cost);
    const double ns = warp_sum((double)nobs);
    if (lane == 0) {
      acost[0] += 0.5 * cs;
      acost[1] += ns;
    }
  }
  __syncthreads();
  // block partial = fixed-order sum of the warp tiles
  double* part = ws.partials + (size_t)blockIdx.x * nsys;
  for (int e = threadIdx.x; e < nsys; e += kBAThreads) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < kBAWarps; ++w) acc += s_sys[w * nsys + e];
    part[e] = acc;
  }
  reduce_partials_last_block(ws.partials, nsys, sys_out, ws.counters + 0);
}

// ---------------------------------------------------------------------------------------------
// One CTA.  Builds and solves the scaled reduced camera system.
constexpr int kSolveThreads = 256;
constexpr int kMaxN = 6 * DF3D_MAX_CAMS;

__global__ void __launch_bounds__(kSolveThreads)
ba_solve_kernel(int C, BAWorkspace ws, const double* __restrict__ sys) {
  __shared__ double A[kMaxN][kMaxN + 1];
  __shared__ double rhs[kMaxN];
  __shared__ double dsc[kMaxN];
  BAState* st = ws.state;
  if (st->done) return;
  const int n = 6 * C;
  const double lambda = st->lambda;
  const double* U = sys + off_U(C);
  const double* gc = sys + off_gc(C);
  const double* S = sys + off_S(C);
  const double* bt = sys + off_b(C);

  if (threadIdx.x < n) {
    const int c = threadIdx.x / 6, i = threadIdx.x % 6;
    double s = fmax(ws.sinv_c[threadIdx.x], sqrt(U[c * 36 + i * 6 + i]));
    ws.sinv_c[threadIdx.x] = s;
    dsc[threadIdx.x] = (s == 0.0) ? 1.0 : 1.0 / s;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n * n; e += kSolveThreads) {
    const int row = e / n, col = e % n;
    const int lo = row < col ? row : col, hi = row < col ? col : row;  // stored upper triangle
    double v = -S[lo * n + hi];
    if (lo / 6 == hi / 6) {
      const int c = lo / 6;
      v += U[c * 36 + (lo % 6) * 6 + (hi % 6)];
    }
    v *= dsc[row] * dsc[col];
    if (row == col) v += lambda;
    A[row][col] = v;
  }
  if (threadIdx.x < n) rhs[threadIdx.x] = dsc[threadIdx.x] * (-gc[threadIdx.x] + bt[threadIdx.x]);
  __syncthreads();

  // in-place Cholesky A = L L^T (lower), right-looking, one column per step
  for (int k = 0; k < n; ++k) {
    if (threadIdx.x == 0) A[k][k] = sqrt(fmax(A[k][k], 1e-300));
    __syncthreads();
    const double dk = A[k][k];
    for (int i = k + 1 + threadIdx.x; i < n; i += kSolveThreads) A[i][k] /= dk;
    __syncthreads();
    const int m = n - k - 1;
    for (int e = threadIdx.x; e < m * m; e += kSolveThreads) {
      const int i = k + 1 + e / m, j = k + 1 + e % m;
      if (j <= i) A[i][j] -= A[i][k] * A[j][k];
    }
    __syncthreads();
  }
  // forward / backward substitution by one thread (n <= 48)
  if (threadIdx.x == 0) {
    for (int i = 0; i < n; ++i) {
      double acc = rhs[i];
      for (int j = 0; j < i; ++j) acc -= A[i][j] * rhs[j];
      rhs[i] = acc / A[i][i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double acc = rhs[i];
      for (int j = i + 1; j < n; ++j) acc -= A[j][i] * rhs[j];
      rhs[i] = acc / A[i][i];
    }
    if (st->iter == 0 && st->F < 0.0) {
      st->F = st->F0 = sys[off_cost(C)];
      st->n_obs = (int)(sys[off_cost(C) + 1] + 0.5);
    }
  }
  __syncthreads();
  if (threadIdx.x < n) {
    const double d = dsc[threadIdx.x] * rhs[threadIdx.x];
    ws.dcam[threadIdx.x] = d;
    ws.cam_new[threadIdx.x] = ws.cam[threadIdx.x] + d;
  }
}

// ---------------------------------------------------------------------------------------------
// Back-substitution + candidate cost.  dynamic smem: cams (old) | cams (new) | dcam
__global__ void __launch_bounds__(kBAThreads)
ba_evaluate_kernel(const double* __restrict__ intr, const double2* __restrict__ pts_xy,
                   const double* __restrict__ pts3d, int C, int TJ, BAWorkspace ws, double* __restrict__ cost_out) {
  extern __shared__ double smem[];
  if (ws.state->done) return;
  const double lambda = ws.state->lambda;
  double* s_old = smem;
  double* s_new = smem + C * kCamStride;
  double* s_dc = s_new + C * kCamStride;
  __shared__ double s_part[kBAWarps];
  if (threadIdx.x < C) {
    stage_camera(ws.cam + threadIdx.x * 6, intr + threadIdx.x * 4, s_old + threadIdx.x * kCamStride);
    stage_camera(ws.cam_new + threadIdx.x * 6, intr + threadIdx.x * 4, s_new + threadIdx.x * kCamStride);
  }
  if (threadIdx.x < 6 * C) s_dc[threadIdx.x] = ws.dcam[threadIdx.x];
  __syncthreads();

  double cost = 0.0;
  for (int g = blockIdx.x * kBAThreads + threadIdx.x; g < TJ; g += gridDim.x * kBAThreads) {
    double X[3] = {pts3d[(size_t)g * 3 + 0], pts3d[(size_t)g * 3 + 1], pts3d[(size_t)g * 3 + 2]};
    double V[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double rhs[3] = {0, 0, 0};  // -(g_p + sum_c W_c^T dc)
    unsigned mask = 0;
    for (int c = 0; c < C; ++c) {
      const double2 xy = __ldg(pts_xy + (size_t)c * TJ + g);
      if (xy.x == 0.0 || xy.y == 0.0) continue;
      mask |= 1u << c;
      double r[2], Jc[2][6], Jp[2][3];
      project_jacobian(s_old + c * kCamStride, X, xy.x, xy.y, r, Jc, Jp);
      // r + Jc dc  (first-order residual after the camera step)
      double q[2];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        double acc = r[a];
#pragma unroll
        for (int i = 0; i < 6; ++i) acc += Jc[a][i] * s_dc[c * 6 + i];
        q[a] = acc;
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        rhs[i] -= Jp[0][i] * q[0] + Jp[1][i] * q[1];
#pragma unroll
        for (int j = 0; j < 3; ++j) V[i][j] += Jp[0][i] * Jp[0][j] + Jp[1][i] * Jp[1][j];
      }
    }
    double Xn[3] = {X[0], X[1], X[2]};
    if (mask) {
      double d[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double s = ws.sinv_p[(size_t)g * 3 + i];  // already updated by ba_linearize
        d[i] = (s == 0.0) ? 1.0 : 1.0 / s;
      }
      double Vh[3][3], Vi[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Vh[i][j] = d[i] * V[i][j] * d[j] + (i == j ? lambda : 0.0);
      if (inv3_sym(Vh, Vi)) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < 3; ++j) acc += d[i] * Vi[i][j] * d[j] * rhs[j];
          Xn[i] += acc;
        }
      }
    }
    ws.X_new[(size_t)g * 3 + 0] = Xn[0];
    ws.X_new[(size_t)g * 3 + 1] = Xn[1];
    ws.X_new[(size_t)g * 3 + 2] = Xn[2];
    for (int c = 0; c < C; ++c) {
      if (!((mask >> c) & 1u)) continue;
      const double2 xy = __ldg(pts_xy + (size_t)c * TJ + g);
      double r[2];
      project_residual(s_new + c * kCamStride, Xn, xy.x, xy.y, r);
      cost += r[0] * r[0] + r[1] * r[1];
    }
  }
  cost = warp_sum(cost);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = cost;
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < kBAWarps; ++w) acc += s_part[w];
    ws.cost_partials[blockIdx.x * 2 + 0] = 0.5 * acc;
    ws.cost_partials[blockIdx.x * 2 + 1] = 0.0;
  }
  reduce_partials_last_block(ws.cost_partials, 2, cost_out, ws.counters + 1);
}

__global__ void ba_decide_kernel(int C, BAWorkspace ws, const double* __restrict__ cost) {
  BAState* st = ws.state;
  if (threadIdx.x != 0) return;
  st->accept_flag = 0;
  if (st->done) return;
  const double Fn = cost[0];
  st->iter += 1;
  if (st->n_obs == 0) {
    st->done = 1;
    st->status = 1;
    return;
  }
  if (Fn < st->F) {
    const double dF = st->F - Fn;
    const double Fold = st->F;
    st->F = Fn;
    st->accepted += 1;
    st->accept_flag = 1;
    for (int i = 0; i < 6 * C; ++i) ws.cam[i] = ws.cam_new[i];
    if (dF < st->ftol * Fold) {
      st->done = 1;
      st->status = 1;
    }
  } else {
    st->lambda = fmin(st->lambda * 10.0, 1e8);
  }
  if (st->iter >= st->max_iters) st->done = 1;
}

__global__ void ba_apply_points_kernel(int n, BAWorkspace ws, double* __restrict__ pts3d) {
  if (!ws.state->accept_flag) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) pts3d[i] = ws.X_new[i];
}

__global__ void ba_end_kernel(double* __restrict__ cam_rt, int C, BAWorkspace ws, df3d_ba_report* report) {
  if (threadIdx.x < 6 * C) cam_rt[threadIdx.x] = ws.cam[threadIdx.x];
  if (threadIdx.x == 0 && report) {
    const BAState s = *ws.state;
    report->cost0 = s.F0;
    report->cost = s.F;
    report->lambda = s.lambda;
    report->iters = s.iter;
    report->accepted = s.accepted;
    report->n_obs = s.n_obs;
    report->status = s.status;
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBAThreads)
reprojection_error_kernel(const double* __restrict__ cam_rt, const double* __restrict__ intr,
                          const double2* __restrict__ pts_xy, const double* __restrict__ pts3d, int C, int TJ,
                          double* __restrict__ out) {
  __shared__ double s_cam[DF3D_MAX_CAMS * kCamStride];
  if (threadIdx.x < C) stage_camera(cam_rt + threadIdx.x * 6, intr + threadIdx.x * 4, s_cam + threadIdx.x * kCamStride);
  __syncthreads();
  double sum = 0.0, cnt = 0.0;
  for (int g = blockIdx.x * kBAThreads + threadIdx.x; g < TJ; g += gridDim.x * kBAThreads) {
    const double X[3] = {pts3d[(size_t)g * 3 + 0], pts3d[(size_t)g * 3 + 1], pts3d[(size_t)g * 3 + 2]};
    for (int c = 0; c < C; ++c) {
      const double2 xy = __ldg(pts_xy + (size_t)c * TJ + g);
      if (xy.x == 0.0 || xy.y == 0.0) continue;
      double r[2];
      project_residual(s_cam + c * kCamStride, X, xy.x, xy.y, r);
      sum += sqrt(r[0] * r[0] + r[1] * r[1]);
      cnt += 1.0;
    }
  }
  sum = warp_sum(sum);
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) {  // a printed metric: one atomic per warp is fine here
    atomicAdd(out + 0, sum);
    atomicAdd(out + 1, cnt);
  }
}

static int check_common(const char* fn, int C, int T, int J) {
  DF3D_REQUIRE(C >= 1 && C <= DF3D_MAX_CAMS, DF3D_EINVAL, "%s: C must be in [1,%d]", fn, DF3D_MAX_CAMS);
  DF3D_REQUIRE(T >= 1 && J >= 1 && (long long)T * J * 3 < (1ll << 31), DF3D_EINVAL, "%s: bad T/J", fn);
  return DF3D_OK;
}

}  // namespace df3d

using namespace df3d;

extern "C" size_t df3d_bundle_adjust_workspace_bytes(int C, int T, int J) {
  if (C < 1 || C > DF3D_MAX_CAMS || T < 1 || J < 1) return 0;
  return ba_workspace_layout(C, T, J, nullptr, nullptr) + 256;
}

extern "C" size_t df3d_ba_system_doubles(int C) { return (C < 1 || C > DF3D_MAX_CAMS) ? 0 : (size_t)sys_doubles(C); }

static int get_ws(const char* fn, int C, int T, int J, void* workspace_dev, size_t workspace_bytes, BAWorkspace* ws) {
  DF3D_REQUIRE(workspace_dev, DF3D_EINVAL, "%s: null workspace", fn);
  DF3D_REQUIRE((reinterpret_cast<uintptr_t>(workspace_dev) & 255) == 0, DF3D_EINVAL, "%s: workspace must be 256-byte aligned", fn);
  const size_t need = ba_workspace_layout(C, T, J, static_cast<char*>(workspace_dev), ws);
  if (workspace_bytes != (size_t)-1)
    DF3D_REQUIRE(workspace_bytes >= need, DF3D_ENOMEM, "%s: workspace too small (%zu < %zu bytes)", fn, workspace_bytes, need);
  return DF3D_OK;
}

extern "C" int df3d_ba_begin(const double* cam_rt_dev, const df3d_ba_opts* opts, int C, int T, int J,
                             void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (int e = check_common("df3d_ba_begin", C, T, J)) return e;
  DF3D_REQUIRE(cam_rt_dev, DF3D_EINVAL, "df3d_ba_begin: null pointer");
  df3d_ba_opts o{20, 1e-4, 1e-6};
  if (opts) o = *opts;
  DF3D_REQUIRE(o.max_iters >= 1 && o.max_iters <= 1000 && o.ftol >= 0.0 && o.lambda0 > 0.0, DF3D_EINVAL,
               "df3d_ba_begin: bad options (max_iters in [1,1000], ftol >= 0, lambda0 > 0)");
  BAWorkspace ws;
  if (int e = get_ws("df3d_ba_begin", C, T, J, workspace_dev, workspace_bytes, &ws)) return e;
  ba_begin_kernel<<<ba_grid(T * J), kBAThreads, 0, static_cast<cudaStream_t>(stream)>>>(cam_rt_dev, C, T * J, o, ws);
  DF3D_LAUNCH_CHECK("ba_begin_kernel");
  return DF3D_OK;
}

extern "C" int df3d_ba_linearize(const double* intr_dev, const double* pts_xy_dev, const double* pts3d_dev,
                                 int C, int T, int J, void* workspace_dev, double* sys_dev, void* stream) {
  if (int e = check_common("df3d_ba_linearize", C, T, J)) return e;
  DF3D_REQUIRE(intr_dev && pts_xy_dev && pts3d_dev && sys_dev, DF3D_EINVAL, "df3d_ba_linearize: null pointer");
  BAWorkspace ws;
  if (int e = get_ws("df3d_ba_linearize", C, T, J, workspace_dev, (size_t)-1, &ws)) return e;
  const size_t smem = ((size_t)C * kCamStride + (size_t)kBAWarps * sys_doubles(C)) * sizeof(double);
  DF3D_CUDA(cudaFuncSetAttribute(ba_linearize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ba_linearize_kernel<<<ba_grid(T * J), kBAThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      intr_dev, reinterpret_cast<const double2*>(pts_xy_dev), pts3d_dev, C, T * J, ws, sys_dev);
  DF3D_LAUNCH_CHECK("ba_linearize_kernel");
  return DF3D_OK;
}

extern "C" int df3d_ba_solve(int C, void* workspace_dev, const double* sys_dev, void* stream) {
  DF3D_REQUIRE(C >= 1 && C <= DF3D_MAX_CAMS, DF3D_EINVAL, "df3d_ba_solve: C must be in [1,%d]", DF3D_MAX_CAMS);
  DF3D_REQUIRE(workspace_dev && sys_dev, DF3D_EINVAL, "df3d_ba_solve: null pointer");
  BAWorkspace ws;
  ba_workspace_layout(C, 1, 1, static_cast<char*>(workspace_dev), &ws);  // camera-side fields do not depend on T,J
  ba_solve_kernel<<<1, kSolveThreads, 0, static_cast<cudaStream_t>(stream)>>>(C, ws, sys_dev);
  DF3D_LAUNCH_CHECK("ba_solve_kernel");
  return DF3D_OK;
}

extern "C" int df3d_ba_evaluate(const double* intr_dev, const double* pts_xy_dev, const double* pts3d_dev,
                                int C, int T, int J, void* workspace_dev, double* cost_dev, void* stream) {
  if (int e = check_common("df3d_ba_evaluate", C, T, J)) return e;
  DF3D_REQUIRE(intr_dev && pts_xy_dev && pts3d_dev && cost_dev, DF3D_EINVAL, "df3d_ba_evaluate: null pointer");
  BAWorkspace ws;
  if (int e = get_ws("df3d_ba_evaluate", C, T, J, workspace_dev, (size_t)-1, &ws)) return e;
  const size_t smem = ((size_t)2 * C * kCamStride + 6 * C) * sizeof(double);
  ba_evaluate_kernel<<<ba_grid(T * J), kBAThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      intr_dev, reinterpret_cast<const double2*>(pts_xy_dev), pts3d_dev, C, T * J, ws, cost_dev);
  DF3D_LAUNCH_CHECK("ba_evaluate_kernel");
  return DF3D_OK;
}

extern "C" int df3d_ba_decide(int C, int T, int J, void* workspace_dev, const double* cost_dev,
                              double* pts3d_dev, void* stream) {
  if (int e = check_common("df3d_ba_decide", C, T, J)) return e;
  DF3D_REQUIRE(cost_dev && pts3d_dev, DF3D_EINVAL, "df3d_ba_decide: null pointer");
  BAWorkspace ws;
  if (int e = get_ws("df3d_ba_decide", C, T, J, workspace_dev, (size_t)-1, &ws)) return e;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ba_decide_kernel<<<1, 32, 0, s>>>(C, ws, cost_dev);
  DF3D_LAUNCH_CHECK("ba_decide_kernel");
  const int n = T * J * 3;
  int grid = ceil_div(n, 256);
  if (grid > 4 * kBAMaxBlocks) grid = 4 * kBAMaxBlocks;
  ba_apply_points_kernel<<<grid, 256, 0, s>>>(n, ws, pts3d_dev);
  DF3D_LAUNCH_CHECK("ba_apply_points_kernel");
  return DF3D_OK;
}

extern "C" int df3d_ba_end(double* cam_rt_dev, int C, void* workspace_dev, df3d_ba_report* report_dev, void* stream) {
  DF3D_REQUIRE(C >= 1 && C <= DF3D_MAX_CAMS, DF3D_EINVAL, "df3d_ba_end: C must be in [1,%d]", DF3D_MAX_CAMS);
  DF3D_REQUIRE(cam_rt_dev && workspace_dev, DF3D_EINVAL, "df3d_ba_end: null pointer");
  BAWorkspace ws;
  ba_workspace_layout(C, 1, 1, static_cast<char*>(workspace_dev), &ws);
  ba_end_kernel<<<1, 64, 0, static_cast<cudaStream_t>(stream)>>>(cam_rt_dev, C, ws, report_dev);
  DF3D_LAUNCH_CHECK("ba_end_kernel");
  return DF3D_OK;
}

extern "C" int df3d_bundle_adjust(double* cam_rt_dev, const double* intr_dev, const double* pts_xy_dev,
                                  int C, int T, int J, const df3d_ba_opts* opts, double* pts3d_dev,
                                  df3d_ba_report* report_dev, void* workspace_dev, size_t workspace_bytes,
                                  void* stream) {
  if (int e = check_common("df3d_bundle_adjust", C, T, J)) return e;
  DF3D_REQUIRE(cam_rt_dev && intr_dev && pts_xy_dev && pts3d_dev, DF3D_EINVAL, "df3d_bundle_adjust: null pointer");
  DF3D_REQUIRE((reinterpret_cast<uintptr_t>(pts_xy_dev) & 15) == 0, DF3D_EINVAL, "df3d_bundle_adjust: pts_xy must be 16-byte aligned");
  BAWorkspace ws;
  if (int e = get_ws("df3d_bundle_adjust", C, T, J, workspace_dev, workspace_bytes, &ws)) return e;
  df3d_ba_opts o{20, 1e-4, 1e-6};
  if (opts) o = *opts;
  if (int e = df3d_ba_begin(cam_rt_dev, &o, C, T, J, workspace_dev, workspace_bytes, stream)) return e;
  // fixed launch sequence; kernels become no-ops once the device-side state says `done`
  for (int it = 0; it < o.max_iters; ++it) {
    if (int e = df3d_ba_linearize(intr_dev, pts_xy_dev, pts3d_dev, C, T, J, workspace_dev, ws.sys_local, stream)) return e;
    if (int e = df3d_ba_solve(C, workspace_dev, ws.sys_local, stream)) return e;
    if (int e = df3d_ba_evaluate(intr_dev, pts_xy_dev, pts3d_dev, C, T, J, workspace_dev, ws.cost_local, stream)) return e;
    if (int e = df3d_ba_decide(C, T, J, workspace_dev, ws.cost_local, pts3d_dev, stream)) return e;
  }
  return df3d_ba_end(cam_rt_dev, C, workspace_dev, report_dev, stream);
}

extern "C" int df3d_reprojection_error(const double* cam_rt_dev, const double* intr_dev,
                                       const double* pts_xy_dev, const double* pts3d_dev, int C, int T, int J,
                                       double* out_dev, void* stream) {
  if (int e = check_common("df3d_reprojection_error", C, T, J)) return e;
  DF3D_REQUIRE(cam_rt_dev && intr_dev && pts_xy_dev && pts3d_dev && out_dev, DF3D_EINVAL, "df3d_reprojection_error: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DF3D_CUDA(cudaMemsetAsync(out_dev, 0, 2 * sizeof(double), s));
  reprojection_error_kernel<<<ba_grid(T * J), kBAThreads, 0, s>>>(
      cam_rt_dev, intr_dev, reinterpret_cast<const double2*>(pts_xy_dev), pts3d_dev, C, T * J, out_dev);
  DF3D_LAUNCH_CHECK("reprojection_error_kernel");
  return DF3D_OK;
}
