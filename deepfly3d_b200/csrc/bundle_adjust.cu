// Bundle adjustment of C camera extrinsics + all 3-D points (fp64, everything on the device).
// Replaces pyba CameraNetwork.bundle_adjust(update_intrinsic=False, update_distort=False)
// (reference call site df3d/core.py:249; recipe reconstructed in SURVEY.md Appendix B): pyba hands the
// problem to scipy.optimize.least_squares(method='trf', x_scale='jac', ftol=1e-4, tr_solver='lsmr').
//
// The bundle-adjustment gauge is free (no camera is held fixed), so WHERE the solver ends up depends on
// its step rule; this file therefore follows SciPy's trust-region-reflective iteration for unbounded
// problems (scipy/optimize/_lsq/trf.py::trf_no_bounds, the tr_solver == 'lsmr' branch) statement by
// statement, with one substitution: the regularised Gauss-Newton step that SciPy gets from LSMR,
//     gn_h = argmin |J_h p - f|^2 + reg |p|^2 ,   J_h = J diag(1 / column norm),
// is either that LSMR iteration itself, restated on the device (solver 1, ba_lsmr_kernel: what reproduces the
// reference's golden file to 1.3e-6 mm), or solved exactly through the Schur complement on the 6C camera unknowns
// (solver 0: 1-2e-5 mm from SciPy at T = 15 ... 1000 frames -- the distance between SciPy with its default LSMR
// tolerance and SciPy with a converged LSMR; tools/ba_proto.py is the CPU prototype).
//
// One outer iteration = a fixed sequence of launches, no host synchronisation (device-side flags turn the
// kernels into no-ops once the solver has terminated, and skip the linearisation after a rejected step).  Every
// per-point pass leaves one partial vector per block of 128 points (warp shuffles -> per-warp shared-memory tiles
// with a single writer -> block partial); ba_finish sums the blocks in an order fixed by the block count alone and
// its last block runs the scalar logic of the pass: bit-reproducible, and independent of how the blocks were spread
// over GPUs (df3d_ba_sharded_*: each rank runs a slice of the blocks, the caller all-gathers the partials).
//   ba_gradient : one thread per 3-D point.  Analytic Jacobian (Rodrigues + pin-hole); gradient, column
//                 norms (Jacobi scaling, running maximum like compute_jac_scale), cost, and the pieces of
//                 |J_h g_h|^2; finish: the Cauchy-step regularisation `reg`
//   ba_schur    : one thread per point.  Per-point M = D_p (D_p V D_p + reg I)^-1 D_p; reduced system
//                 S~ = sum W M W^T,  b~ = sum W M g_p
//   ba_solve    : one CTA.  D_c (U - S~) D_c + reg I, Cholesky + one refinement step -> camera part of gn_h
//                 (solver 1: ba_lsmr instead of ba_schur + ba_solve)
//   ba_backsub  : one thread per point.  Point part of gn_h, the Gram quantities of span{g_h, gn_h}; finish:
//                 the 2-D sub-problem, solved inside the trust region
//   ba_step     : one thread per point.  Candidate x + D step_h, its cost; finish: SciPy's ratio test, radius
//                 update and termination rule, and the 2-D problem again after a rejected step
//   ba_apply    : copies the candidate points after an accepted step
#include <cooperative_groups.h>

#include "common.cuh"
#include "geom.cuh"

namespace df3d {

constexpr int kBAThreads = 128;
constexpr int kBAWarps = kBAThreads / 32;
constexpr int kBAMaxBlocks = 4 * 148;  // one partial vector per block, summed by ba_finish_kernel: bounded, but enough
                                      // blocks to keep several per SM in flight (the passes are fp64 latency-bound)
constexpr int kMaxN = 6 * DF3D_MAX_CAMS;
constexpr int kLsmrRed = 64;  // values of one grid-wide reduction of the LSMR kernel (2 scalars + 6 C camera sums, padded)

struct BAState {
  double Delta, F, F0, ftol, xtol, gtol, reg;
  double gg, JgJg;        // |g_h|^2, |J_h g_h|^2
  double T00, T10, T11;   // orthonormal basis of span{g_h, gn_h}: q1 = T00 g_h, q2 = T10 g_h + T11 gn_h
  double B00, B01, B11;   // J_h restricted to that basis, squared
  double gS0, gS1;        // gradient in that basis
  double coef_g, coef_gn; // step_h = coef_g g_h + coef_gn gn_h
  double pred, sh_norm;   // predicted reduction and |step_h| of the current candidate
  int iter, accepted, max_iters, done, status, n_obs, accept_flag, need_lin, first, solver;
  int lsmr_itn, lsmr_istop;  // of the last LSMR solve (solver 1)
};

// ba_gradient's reduction vector
__host__ __device__ inline int g_doubles(int C) { return 36 * C + 6 * C + 6 * C + 6; }
__host__ __device__ inline int g_off_U(int) { return 0; }
__host__ __device__ inline int g_off_gc(int C) { return 36 * C; }
__host__ __device__ inline int g_off_w(int C) { return 42 * C; }
__host__ __device__ inline int g_off_sc(int C) { return 48 * C; }  // cost, n_obs, b^T V b, |g_h points|^2, sum (x sinv)^2, [max] |g|_inf
// ba_schur's reduction vector
__host__ __device__ inline int sys_doubles(int C) { return 36 * C + 6 * C + 36 * C * C + 6 * C + 2; }
__host__ __device__ inline int off_U(int) { return 0; }
__host__ __device__ inline int off_gc(int C) { return 36 * C; }
__host__ __device__ inline int off_S(int C) { return 42 * C; }
__host__ __device__ inline int off_b(int C) { return 42 * C + 36 * C * C; }
__host__ __device__ inline int off_cost(int C) { return 48 * C + 36 * C * C; }

struct BAWorkspace {  // carved out of the caller's buffer
  BAState* state;
  unsigned int* counters;  // last-block tickets of ba_finish_kernel, one per pass
  double* cam;             // C*6 current cameras
  double* sinv_c;          // C*6 camera column norms (running maximum)
  double* gc;              // C*6 camera gradient
  double* ac;              // C*6 gc / sinv^2  (camera part of D g_h)
  double* ghc;             // C*6 gc / sinv    (camera part of g_h)
  double* gnc;             // C*6 camera part of gn_h
  double* dcn;             // C*6 gnc / sinv   (camera part of D gn_h)
  double* sinv_p;          // TJ*3 point column norms (running maximum)
  double* gp;              // TJ*3 point gradient
  double* gnp;             // TJ*3 point part of gn_h
  double* X_new;           // TJ*3 candidate points
  double* partials;        // kBAMaxBlocks * max(sys_doubles, g_doubles)
  double* red;             // reduced vector of the last reducing kernel
  // LSMR (solver 1): u lives in residual space, v / h / hbar in the scaled unknown space (point parts; x = gnp)
  double* lu;              // C*TJ*2
  double* lv;              // TJ*3
  double* lh;              // TJ*3
  double* lhb;             // TJ*3
  double* lpart;           // 2 * kBAMaxBlocks * kLsmrRed
  // Block space of the per-point passes: a launch covers the virtual blocks [vb0, vb0 + gridDim.x) of `vgrid`.  One GPU:
  // vb0 = 0 and gridDim.x = vgrid.  Frame-sharded (`sharded`): each
  // rank launches its own slice of the blocks and the caller all-gathers the partials before ba_finish_kernel sums them
  // -- same blocks, same order, same bits as on one GPU.
  int vb0, vgrid, sharded, pad;
};

__device__ void gradient_tail(int C, BAWorkspace ws);
__device__ void backsub_tail(int C, BAWorkspace ws);
__device__ void step_tail(int C, BAWorkspace ws, const double* s_cn);

static int ba_grid(int TJ) {
  int g = ceil_div(TJ, kBAThreads);
  return g < 1 ? 1 : (g > kBAMaxBlocks ? kBAMaxBlocks : g);
}

static size_t ba_workspace_layout(int C, int T, int J, char* base, BAWorkspace* ws) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~size_t(255);
    return base ? base + o : nullptr;
  };
  const size_t TJ = (size_t)T * J;
  const size_t nred = (size_t)(sys_doubles(C) > g_doubles(C) ? sys_doubles(C) : g_doubles(C));
  BAWorkspace w;
  w.state = reinterpret_cast<BAState*>(take(sizeof(BAState)));
  w.counters = reinterpret_cast<unsigned int*>(take(8 * sizeof(unsigned int)));
  w.cam = reinterpret_cast<double*>(take(C * 6 * 8));
  w.sinv_c = reinterpret_cast<double*>(take(C * 6 * 8));
  w.gc = reinterpret_cast<double*>(take(C * 6 * 8));
  w.ac = reinterpret_cast<double*>(take(C * 6 * 8));
  w.ghc = reinterpret_cast<double*>(take(C * 6 * 8));
  w.gnc = reinterpret_cast<double*>(take(C * 6 * 8));
  w.dcn = reinterpret_cast<double*>(take(C * 6 * 8));
  w.sinv_p = reinterpret_cast<double*>(take(TJ * 3 * 8));
  w.gp = reinterpret_cast<double*>(take(TJ * 3 * 8));
  w.gnp = reinterpret_cast<double*>(take(TJ * 3 * 8));
  w.X_new = reinterpret_cast<double*>(take(TJ * 3 * 8));
  w.partials = reinterpret_cast<double*>(take((size_t)kBAMaxBlocks * nred * 8));
  w.red = reinterpret_cast<double*>(take(nred * 8));
  w.lu = reinterpret_cast<double*>(take((size_t)C * TJ * 2 * 8));
  w.lv = reinterpret_cast<double*>(take(TJ * 3 * 8));
  w.lh = reinterpret_cast<double*>(take(TJ * 3 * 8));
  w.lhb = reinterpret_cast<double*>(take(TJ * 3 * 8));
  w.lpart = reinterpret_cast<double*>(take((size_t)2 * kBAMaxBlocks * kLsmrRed * 8));
  w.vb0 = 0;
  w.vgrid = ba_grid((int)TJ);
  w.sharded = 0;
  w.pad = 0;
  if (ws) *ws = w;
  return off;
}


// ---------------------------------------------------------------------------------------------
__global__ void ba_begin_kernel(const double* __restrict__ cam_rt, int C, df3d_ba_opts opts, BAWorkspace ws) {
  const int g = threadIdx.x;
  if (g == 0) {
    BAState s;
    memset(&s, 0, sizeof(s));
    s.F = s.F0 = -1.0;
    s.ftol = opts.ftol;
    s.xtol = opts.xtol;
    s.gtol = opts.gtol;
    s.max_iters = opts.max_iters;
    s.need_lin = 1;
    s.first = 1;
    s.solver = opts.solver;
    *ws.state = s;
    for (int i = 0; i < 8; ++i) ws.counters[i] = 0u;
  }
  if (g < C * 6) ws.cam[g] = cam_rt[g];
}

// 3x3 symmetric positive definite inverse through its Cholesky factor (backward stable for the
// ill-conditioned per-point blocks: two views of a 16 000 px lens barely constrain the depth)
__device__ __forceinline__ bool inv3_spd(const double (&A)[3][3], double (&Ai)[3][3]) {
  if (!(A[0][0] > 0.0)) return false;
  const double l00 = sqrt(A[0][0]);
  const double l10 = A[1][0] / l00, l20 = A[2][0] / l00;
  const double d1 = A[1][1] - l10 * l10;
  if (!(d1 > 0.0)) return false;
  const double l11 = sqrt(d1);
  const double l21 = (A[2][1] - l20 * l10) / l11;
  const double d2 = A[2][2] - l20 * l20 - l21 * l21;
  if (!(d2 > 0.0)) return false;
  const double l22 = sqrt(d2);
  // Li = L^-1 (lower), A^-1 = Li^T Li
  const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
  const double i10 = -l10 * i00 * i11;
  const double i21 = -l21 * i11 * i22;
  const double i20 = -(l20 * i00 + l21 * i10) * i22;
  Ai[0][0] = i00 * i00 + i10 * i10 + i20 * i20;
  Ai[0][1] = Ai[1][0] = i10 * i11 + i20 * i21;
  Ai[0][2] = Ai[2][0] = i20 * i22;
  Ai[1][1] = i11 * i11 + i21 * i21;
  Ai[1][2] = Ai[2][1] = i21 * i22;
  Ai[2][2] = i22 * i22;
  return true;
}

// M = D (D V D + reg I)^-1 D for a point seen by ONE camera.  V = Jp^T Jp has rank 2 and the explicit 3x3 inverse
// carries a 1 / reg eigenvalue along the viewing ray; every use of M applies it to a vector of the row space of
// Jp (W^T x = Jp^T Jc x, g_p = Jp^T r), where that eigenvalue multiplies an exact zero -- in floating point a
// rounding residue, amplified by 1 / reg ~ 1e12 into the reduced camera system.  The restriction of M to the row
// space has a closed form without the blow-up (push-through identity):
//   M = D Jh^T (G + reg I)^-1 G^-1 Jh D,   Jh = Jp D,  G = Jh Jh^T (2 x 2, well conditioned).
// SciPy's LSMR step is the minimum-norm one and never moves along the ray either.
__device__ __forceinline__ bool single_view_M(const double (&Jp)[2][3], const double (&d)[3], double reg, double (&Mm)[3][3]) {
  double Jh[2][3];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int k = 0; k < 3; ++k) Jh[a][k] = Jp[a][k] * d[k];
  const double g00 = Jh[0][0] * Jh[0][0] + Jh[0][1] * Jh[0][1] + Jh[0][2] * Jh[0][2];
  const double g01 = Jh[0][0] * Jh[1][0] + Jh[0][1] * Jh[1][1] + Jh[0][2] * Jh[1][2];
  const double g11 = Jh[1][0] * Jh[1][0] + Jh[1][1] * Jh[1][1] + Jh[1][2] * Jh[1][2];
  const double det = g00 * g11 - g01 * g01;
  const double r00 = g00 + reg, r11 = g11 + reg;
  const double detr = r00 * r11 - g01 * g01;
  if (!(det > 0.0) || !(detr > 0.0)) return false;
  // H = (G + reg I)^-1 G^-1 (the two inverses commute: H is symmetric)
  const double a00 = r11 / detr, a01 = -g01 / detr, a11 = r00 / detr;
  const double b00 = g11 / det, b01 = -g01 / det, b11 = g00 / det;
  const double h00 = a00 * b00 + a01 * b01;
  const double h01 = 0.5 * ((a00 * b01 + a01 * b11) + (a01 * b00 + a11 * b01));
  const double h11 = a01 * b01 + a11 * b11;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Mm[i][j] = d[i] * d[j] * (Jh[0][i] * (h00 * Jh[0][j] + h01 * Jh[1][j]) + Jh[1][i] * (h01 * Jh[0][j] + h11 * Jh[1][j]));
  return true;
}

// M = D (D V D + reg I)^-1 D of one point (zero when the point has no observation or the block is singular)
__device__ __forceinline__ void point_M(unsigned mask, const double (&V)[3][3], const double (&JpLast)[2][3], const double (&d)[3],
                                        double reg, double (&Mm)[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Mm[i][j] = 0.0;
  if (!mask) return;
  if (__popc(mask) == 1) {
    single_view_M(JpLast, d, reg, Mm);
    return;
  }
  double Vh[3][3], Vi[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Vh[i][j] = d[i] * V[i][j] * d[j] + (i == j ? reg : 0.0);
  if (inv3_spd(Vh, Vi)) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Mm[i][j] = d[i] * Vi[i][j] * d[j];
  }
}

// Every per-point pass ends with one partial vector per block (block_partial); ba_finish_kernel sums the blocks in a
// fixed order -- entries [0, n_sum) are summed, [n_sum, n) take the maximum --, and its last block runs the scalar logic
// that follows the pass.  (Round 2 first summed inside the last block of the pass itself: 592 blocks x
// 2 102 doubles through ONE block cost ~0.6 ms per evaluation at 8 000 frames, two thirds of the whole solve.)
// block partial = fixed-order sum (or max) of the per-warp tiles
__device__ __forceinline__ void block_partial(const double* s_tiles, int n_sum, int n, double* part) {
  __syncthreads();
  for (int e = threadIdx.x; e < n; e += kBAThreads) {
    double acc = s_tiles[e];
#pragma unroll
    for (int w = 1; w < kBAWarps; ++w) acc = e < n_sum ? acc + s_tiles[w * n + e] : fmax(acc, s_tiles[w * n + e]);
    part[e] = acc;
  }
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// SciPy solve_trust_region_2d: min 0.5 p^T B p + g^T p subject to |p| <= Delta.  Same minimiser; found through the
// eigen-decomposition of B and bisection on the secular equation instead of the roots of a quartic.
__device__ void solve_tr_2d(double b00, double b01, double b11, double g0, double g1, double Delta, double (&p)[2]) {
  // eigenvalues w0 <= w1, eigenvectors as a rotation (c, s)
  const double tr = 0.5 * (b00 + b11), df = 0.5 * (b00 - b11);
  const double rad = sqrt(df * df + b01 * b01);
  const double w0 = tr - rad, w1 = tr + rad;
  double vx, vy;  // eigenvector of w1
  if (rad == 0.0) {
    vx = 1.0;
    vy = 0.0;
  } else if (df >= 0.0) {
    vx = df + rad;
    vy = b01;
  } else {
    vx = b01;
    vy = rad - df;
  }
  const double vn = sqrt(vx * vx + vy * vy);
  vx /= vn;
  vy /= vn;
  // q1 = (vx, vy) for w1, q0 = (-vy, vx) for w0
  const double gq0 = -vy * g0 + vx * g1, gq1 = vx * g0 + vy * g1;
  double y0, y1;
  bool inside = false;
  if (w0 > 0.0) {
    y0 = -gq0 / w0;
    y1 = -gq1 / w1;
    inside = (y0 * y0 + y1 * y1) <= Delta * Delta;
  }
  if (!inside) {
    const double lo0 = fmax(0.0, -w0);
    auto f = [&](double mu) {
      const double a = gq0 / (w0 + mu), b = gq1 / (w1 + mu);
      return sqrt(a * a + b * b) - Delta;
    };
    double lo = lo0, hi = lo0 + fmax(1.0, fabs(w1));
    int guard = 0;
    while (f(hi) > 0.0 && guard++ < 2000) hi = lo0 + 2.0 * (hi - lo0);
    for (int it = 0; it < 200; ++it) {
      const double mid = 0.5 * (lo + hi);
      if (f(mid) > 0.0)
        lo = mid;
      else
        hi = mid;
    }
    const double mu = 0.5 * (lo + hi);
    y0 = -gq0 / (w0 + mu);
    y1 = -gq1 / (w1 + mu);
  }
  p[0] = -vy * y0 + vx * y1;
  p[1] = vx * y0 + vy * y1;
}

// candidate step inside the current trust region (thread 0 of a last block)
__device__ void solve_subproblem(BAState* st) {
  double p[2] = {0.0, 0.0};
  if (st->T11 == 0.0) {  // gn_h parallel to g_h: one dimension
    if (st->B00 > 0.0) p[0] = -st->gS0 / st->B00;
    if (!(st->B00 > 0.0) || fabs(p[0]) > st->Delta) p[0] = st->gS0 > 0.0 ? -st->Delta : st->Delta;
  } else {
    solve_tr_2d(st->B00, st->B01, st->B11, st->gS0, st->gS1, st->Delta, p);
  }
  st->coef_g = p[0] * st->T00 + p[1] * st->T10;
  st->coef_gn = p[1] * st->T11;
  st->pred = -(0.5 * (p[0] * (st->B00 * p[0] + st->B01 * p[1]) + p[1] * (st->B01 * p[0] + st->B11 * p[1])) +
               st->gS0 * p[0] + st->gS1 * p[1]);
  st->sh_norm = sqrt(p[0] * p[0] + p[1] * p[1]);
}

// ---------------------------------------------------------------------------------------------
// dynamic shared memory: cams [C*kCamStride] | per-warp tiles [kBAWarps * g_doubles(C)]
__global__ void __launch_bounds__(kBAThreads)
ba_gradient_kernel(const double* __restrict__ intr, const double2* __restrict__ pts_xy,
                   const double* __restrict__ pts3d, int C, int TJ, BAWorkspace ws) {
  extern __shared__ double smem[];
  BAState* st = ws.state;
  if (st->done || !st->need_lin) return;
  const bool first = st->first != 0;
  const int n = g_doubles(C), n_sum = n - 1;
  double* s_cam = smem;
  double* s_t = smem + C * kCamStride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* my = s_t + warp * n;

  if (threadIdx.x < C) stage_camera(ws.cam + threadIdx.x * 6, intr + threadIdx.x * 4, s_cam + threadIdx.x * kCamStride);
  for (int e = threadIdx.x; e < kBAWarps * n; e += kBAThreads) s_t[e] = 0.0;
  __syncthreads();
  double* aU = my + g_off_U(C);
  double* agc = my + g_off_gc(C);
  double* aw = my + g_off_w(C);
  double* asc = my + g_off_sc(C);

  for (int base = (blockIdx.x + ws.vb0) * kBAThreads; base < TJ; base += ws.vgrid * kBAThreads) {
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
    unsigned mask = 0;
    double cost = 0.0;
    int nobs = 0;
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
    // Jacobi scaling of the point columns: column norm, running maximum (SciPy compute_jac_scale)
    double b[3] = {0, 0, 0}, ggp = 0.0, bVb = 0.0, dsq = 0.0, gmax = 0.0;
    if (valid) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double cn = sqrt(V[i][i]);
        double s = first ? (cn == 0.0 ? 1.0 : cn) : fmax(ws.sinv_p[(size_t)g * 3 + i], cn);
        ws.sinv_p[(size_t)g * 3 + i] = s;
        ws.gp[(size_t)g * 3 + i] = gp[i];
        b[i] = gp[i] / (s * s);
        ggp += (gp[i] / s) * (gp[i] / s);
        dsq += (X[i] * s) * (X[i] * s);
        gmax = fmax(gmax, fabs(gp[i]));
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) bVb += b[i] * (V[i][0] * b[0] + V[i][1] * b[1] + V[i][2] * b[2]);
    }
    // w_c = sum_obs Jc^T (Jp b): the cross term of |J_h g_h|^2
    for (int c = 0; c < C; ++c) {
      const bool vis = (mask >> c) & 1u;
      if (!__any_sync(0xffffffffu, vis)) continue;
      double wv[6] = {0, 0, 0, 0, 0, 0};
      if (vis) {
        const double2 xy = __ldg(pts_xy + (size_t)c * TJ + g);
        double r[2], Jc[2][6], Jp[2][3];
        project_jacobian(s_cam + c * kCamStride, X, xy.x, xy.y, r, Jc, Jp);
        const double q0 = Jp[0][0] * b[0] + Jp[0][1] * b[1] + Jp[0][2] * b[2];
        const double q1 = Jp[1][0] * b[0] + Jp[1][1] * b[1] + Jp[1][2] * b[2];
#pragma unroll
        for (int i = 0; i < 6; ++i) wv[i] = Jc[0][i] * q0 + Jc[1][i] * q1;
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double v = warp_sum(wv[i]);
        if (lane == 0) aw[c * 6 + i] += v;
      }
    }
    const double cs = warp_sum(cost), ns = warp_sum((double)nobs), bs = warp_sum(bVb), gs = warp_sum(ggp), ds = warp_sum(dsq);
    const double gm = warp_max(gmax);
    if (lane == 0) {
      asc[0] += 0.5 * cs;
      asc[1] += ns;
      asc[2] += bs;
      asc[3] += gs;
      asc[4] += ds;
      asc[5] = fmax(asc[5], gm);
    }
  }
  block_partial(s_t, n_sum, n, ws.partials + (size_t)(blockIdx.x + ws.vb0) * n);  // ba_finish_kernel(0) goes on from here
}

// ---- after the reduction, one thread: camera scaling, gradient norms, Cauchy-step regularisation
__device__ void gradient_tail(int C, BAWorkspace ws) {
  BAState* st = ws.state;
  const bool first = st->first != 0;
  const double* red = ws.red;
  const double* U = red + g_off_U(C);
  const double* gc = red + g_off_gc(C);
  const double* w = red + g_off_w(C);
  const double* sc = red + g_off_sc(C);
  const int n6 = 6 * C;
  double gg = sc[3], gmax = sc[5], dsq = sc[4];
  double aUa = 0.0, aw2 = 0.0;
  for (int i = 0; i < n6; ++i) {
    const int c = i / 6, k = i % 6;
    const double cn = sqrt(U[c * 36 + k * 6 + k]);
    const double s = first ? (cn == 0.0 ? 1.0 : cn) : fmax(ws.sinv_c[i], cn);
    ws.sinv_c[i] = s;
    ws.gc[i] = gc[i];
    ws.ac[i] = gc[i] / (s * s);
    ws.ghc[i] = gc[i] / s;
    gg += ws.ghc[i] * ws.ghc[i];
    gmax = fmax(gmax, fabs(gc[i]));
    dsq += (ws.cam[i] * s) * (ws.cam[i] * s);
  }
  for (int c = 0; c < C; ++c)
    for (int i = 0; i < 6; ++i) {
      double row = 0.0;
      for (int j = 0; j < 6; ++j) {
        const double u = i <= j ? U[c * 36 + i * 6 + j] : U[c * 36 + j * 6 + i];
        row += u * ws.ac[c * 6 + j];
      }
      aUa += ws.ac[c * 6 + i] * row;
      aw2 += ws.ac[c * 6 + i] * w[c * 6 + i];
    }
  const double JgJg = aUa + 2.0 * aw2 + sc[2];
  if (first) {
    st->F = st->F0 = sc[0];
    st->n_obs = (int)(sc[1] + 0.5);
    st->Delta = sqrt(dsq);
    if (st->Delta == 0.0) st->Delta = 1.0;
    st->first = 0;
  }
  st->gg = gg;
  st->JgJg = JgJg;
  if (st->n_obs == 0 || gmax < st->gtol) {  // trf: g_norm < gtol
    st->done = 1;
    st->status = 1;
    return;
  }
  // reg_term = -min_{0 <= t <= Delta / |g_h|} (a t^2 + b t) / Delta^2,  a = |J_h g_h|^2 / 2, b = -|g_h|^2
  const double a = 0.5 * JgJg, bq = -gg;
  const double to_tr = st->Delta / sqrt(gg);
  double best = fmin(0.0, to_tr * (a * to_tr + bq));
  if (a != 0.0) {
    const double ext = -0.5 * bq / a;
    if (ext > 0.0 && ext < to_tr) best = fmin(best, ext * (a * ext + bq));
  }
  st->reg = -best / (st->Delta * st->Delta);
}

// ---------------------------------------------------------------------------------------------
// dynamic shared memory: cams [C*kCamStride] | per-warp system tiles [kBAWarps * sys_doubles(C)]
__global__ void __launch_bounds__(kBAThreads)
ba_schur_kernel(const double* __restrict__ intr, const double2* __restrict__ pts_xy,
                const double* __restrict__ pts3d, int C, int TJ, BAWorkspace ws) {
  extern __shared__ double smem[];
  if (ws.state->done || !ws.state->need_lin || ws.state->solver == 1) return;
  const double reg = ws.state->reg;
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

  for (int base = (blockIdx.x + ws.vb0) * kBAThreads; base < TJ; base += ws.vgrid * kBAThreads) {
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
    double JpLast[2][3] = {{0, 0, 0}, {0, 0, 0}};  // Jacobian of the last visible camera (the only one of a 1-view point)
    unsigned mask = 0;

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
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int i = 0; i < 3; ++i) JpLast[a][i] = Jp[a][i];
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

    // M = D (D V D + reg I)^-1 D with the point scaling fixed by ba_gradient
    double Mm[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    if (valid && mask) {
      double d[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) d[i] = 1.0 / ws.sinv_p[(size_t)g * 3 + i];
      point_M(mask, V, JpLast, d, reg, Mm);
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
  }
  block_partial(s_sys, nsys, nsys, ws.partials + (size_t)(blockIdx.x + ws.vb0) * nsys);  // ba_finish_kernel(1) sums them
}

// ---------------------------------------------------------------------------------------------
// One CTA.  Builds and solves the scaled reduced camera system for the camera part of gn_h.
constexpr int kSolveThreads = 256;

__global__ void __launch_bounds__(kSolveThreads) ba_solve_kernel(int C, BAWorkspace ws) {
  __shared__ double A[kMaxN][kMaxN + 1];   // Cholesky factor (lower) after the factorisation
  __shared__ double A0[kMaxN][kMaxN + 1];  // the matrix itself, for the refinement step
  __shared__ double rhs[kMaxN], sol[kMaxN], res[kMaxN];
  __shared__ double dsc[kMaxN];
  BAState* st = ws.state;
  if (st->done || !st->need_lin || st->solver == 1) return;
  const int n = 6 * C;
  const double reg = st->reg;
  const double* sys = ws.red;
  const double* U = sys + off_U(C);
  const double* gc = sys + off_gc(C);
  const double* S = sys + off_S(C);
  const double* bt = sys + off_b(C);

  if (threadIdx.x < n) dsc[threadIdx.x] = 1.0 / ws.sinv_c[threadIdx.x];
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
    if (row == col) v += reg;
    A[row][col] = v;
    A0[row][col] = v;
  }
  if (threadIdx.x < n) rhs[threadIdx.x] = dsc[threadIdx.x] * (gc[threadIdx.x] - bt[threadIdx.x]);
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
  // solve, then one step of iterative refinement (the system carries eigenvalues down to `reg`: the gauge
  // directions and cameras without observations)
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) {
      if (threadIdx.x < n) {
        double acc = rhs[threadIdx.x];
        for (int j = 0; j < n; ++j) acc -= A0[threadIdx.x][j] * sol[j];
        res[threadIdx.x] = acc;
      }
    } else if (threadIdx.x < n) {
      res[threadIdx.x] = rhs[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 0; i < n; ++i) {
        double acc = res[i];
        for (int j = 0; j < i; ++j) acc -= A[i][j] * res[j];
        res[i] = acc / A[i][i];
      }
      for (int i = n - 1; i >= 0; --i) {
        double acc = res[i];
        for (int j = i + 1; j < n; ++j) acc -= A[j][i] * res[j];
        res[i] = acc / A[i][i];
      }
    }
    __syncthreads();
    if (threadIdx.x < n) sol[threadIdx.x] = pass == 0 ? res[threadIdx.x] : sol[threadIdx.x] + res[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x < n) {
    ws.gnc[threadIdx.x] = sol[threadIdx.x];
    ws.dcn[threadIdx.x] = dsc[threadIdx.x] * sol[threadIdx.x];
  }
}

// ---------------------------------------------------------------------------------------------
// Point part of gn_h + Gram quantities.  dynamic smem: cams | ac [6C] | dcn [6C] | tiles [kBAWarps * 4]
__global__ void __launch_bounds__(kBAThreads)
ba_backsub_kernel(const double* __restrict__ intr, const double2* __restrict__ pts_xy,
                  const double* __restrict__ pts3d, int C, int TJ, BAWorkspace ws) {
  extern __shared__ double smem[];
  BAState* st = ws.state;
  if (st->done || !st->need_lin) return;
  const double reg = st->reg;
  double* s_cam = smem;
  double* s_ac = smem + C * kCamStride;
  double* s_dc = s_ac + 6 * C;
  double* s_t = s_dc + 6 * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < C) stage_camera(ws.cam + threadIdx.x * 6, intr + threadIdx.x * 4, s_cam + threadIdx.x * kCamStride);
  if (threadIdx.x < 6 * C) {
    s_ac[threadIdx.x] = ws.ac[threadIdx.x];
    s_dc[threadIdx.x] = ws.dcn[threadIdx.x];
  }
  if (threadIdx.x < kBAWarps * 4) s_t[threadIdx.x] = 0.0;
  __syncthreads();

  double a_ggn = 0.0, a_gngn = 0.0, a_JgJgn = 0.0, a_JgnJgn = 0.0;
  for (int g = (blockIdx.x + ws.vb0) * kBAThreads + threadIdx.x; g < TJ; g += ws.vgrid * kBAThreads) {
    const double X[3] = {pts3d[(size_t)g * 3 + 0], pts3d[(size_t)g * 3 + 1], pts3d[(size_t)g * 3 + 2]};
    double V[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double rp[3] = {0, 0, 0};  // sum Jp^T (r - Jc dcn)
    double JpLast[2][3] = {{0, 0, 0}, {0, 0, 0}};
    unsigned mask = 0;
    for (int c = 0; c < C; ++c) {
      const double2 xy = __ldg(pts_xy + (size_t)c * TJ + g);
      if (xy.x == 0.0 || xy.y == 0.0) continue;
      mask |= 1u << c;
      double r[2], Jc[2][6], Jp[2][3];
      project_jacobian(s_cam + c * kCamStride, X, xy.x, xy.y, r, Jc, Jp);
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i) JpLast[a][i] = Jp[a][i];
      double q[2];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        double acc = r[a];
#pragma unroll
        for (int i = 0; i < 6; ++i) acc -= Jc[a][i] * s_dc[c * 6 + i];
        q[a] = acc;
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        rp[i] += Jp[0][i] * q[0] + Jp[1][i] * q[1];
#pragma unroll
        for (int j = 0; j < 3; ++j) V[i][j] += Jp[0][i] * Jp[0][j] + Jp[1][i] * Jp[1][j];
      }
    }
    double pp[3] = {0, 0, 0};  // scaled point part of gn_h
    double d[3], gh[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      d[i] = 1.0 / ws.sinv_p[(size_t)g * 3 + i];
      gh[i] = ws.gp[(size_t)g * 3 + i] * d[i];
    }
    if (st->solver == 1) {  // gn_h comes from the LSMR kernel
#pragma unroll
      for (int i = 0; i < 3; ++i) pp[i] = ws.gnp[(size_t)g * 3 + i];
    } else if (mask) {  // pp = (D V D + reg I)^-1 D rp = D^-1 M rp
      double Mm[3][3];
      point_M(mask, V, JpLast, d, reg, Mm);
#pragma unroll
      for (int i = 0; i < 3; ++i) pp[i] = (Mm[i][0] * rp[0] + Mm[i][1] * rp[1] + Mm[i][2] * rp[2]) / d[i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      ws.gnp[(size_t)g * 3 + i] = pp[i];
      a_ggn += gh[i] * pp[i];
      a_gngn += pp[i] * pp[i];
    }
    // J_h g_h and J_h gn_h of this point's observations
    const double bp[3] = {gh[0] * d[0], gh[1] * d[1], gh[2] * d[2]};
    const double dp[3] = {pp[0] * d[0], pp[1] * d[1], pp[2] * d[2]};
    for (int c = 0; c < C; ++c) {
      if (!((mask >> c) & 1u)) continue;
      const double2 xy = __ldg(pts_xy + (size_t)c * TJ + g);
      double r[2], Jc[2][6], Jp[2][3];
      project_jacobian(s_cam + c * kCamStride, X, xy.x, xy.y, r, Jc, Jp);
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        double jg = Jp[a][0] * bp[0] + Jp[a][1] * bp[1] + Jp[a][2] * bp[2];
        double jn = Jp[a][0] * dp[0] + Jp[a][1] * dp[1] + Jp[a][2] * dp[2];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          jg += Jc[a][i] * s_ac[c * 6 + i];
          jn += Jc[a][i] * s_dc[c * 6 + i];
        }
        a_JgJgn += jg * jn;
        a_JgnJgn += jn * jn;
      }
    }
  }
  a_ggn = warp_sum(a_ggn);
  a_gngn = warp_sum(a_gngn);
  a_JgJgn = warp_sum(a_JgJgn);
  a_JgnJgn = warp_sum(a_JgnJgn);
  if (lane == 0) {
    s_t[warp * 4 + 0] = a_ggn;
    s_t[warp * 4 + 1] = a_gngn;
    s_t[warp * 4 + 2] = a_JgJgn;
    s_t[warp * 4 + 3] = a_JgnJgn;
  }
  block_partial(s_t, 4, 4, ws.partials + (size_t)(blockIdx.x + ws.vb0) * 4);  // ba_finish_kernel(2) goes on from here
}

// ---- after the reduction, one thread: the 2-D sub-problem in an orthonormal basis of span{g_h, gn_h}
__device__ void backsub_tail(int C, BAWorkspace ws) {
  BAState* st = ws.state;
  double g_gn = ws.red[0], gn_gn = ws.red[1];
  const double Jg_Jgn = ws.red[2], Jgn_Jgn = ws.red[3];
  for (int i = 0; i < 6 * C; ++i) {
    g_gn += ws.ghc[i] * ws.gnc[i];
    gn_gn += ws.gnc[i] * ws.gnc[i];
  }
  const double gg = st->gg, JgJg = st->JgJg;
  const double n1 = sqrt(gg);
  const double c12 = g_gn / n1;
  const double n2sq = gn_gn - c12 * c12;
  const double n2 = n2sq > 0.0 ? sqrt(n2sq) : 0.0;
  st->T00 = 1.0 / n1;
  if (n2 > 1e-14 * sqrt(gn_gn)) {
    st->T10 = -c12 / (n1 * n2);
    st->T11 = 1.0 / n2;
  } else {
    st->T10 = st->T11 = 0.0;
  }
  // B_S = T Bg T^T with Bg = [[JgJg, Jg_Jgn], [Jg_Jgn, Jgn_Jgn]];  g_S = T [gg, g_gn]
  const double t00 = st->T00, t10 = st->T10, t11 = st->T11;
  st->B00 = t00 * t00 * JgJg;
  st->B01 = t00 * (t10 * JgJg + t11 * Jg_Jgn);
  st->B11 = t10 * t10 * JgJg + 2.0 * t10 * t11 * Jg_Jgn + t11 * t11 * Jgn_Jgn;
  st->gS0 = t00 * gg;
  st->gS1 = t10 * gg + t11 * g_gn;
  solve_subproblem(st);
}

// ---------------------------------------------------------------------------------------------
// solver 1: the regularised Gauss-Newton step the way SciPy gets it -- scipy.sparse.linalg.lsmr(J_h, f, damp =
// sqrt(reg), atol = btol = 1e-6, conlim = 1e8) with x0 = 0 (scipy/sparse/linalg/_isolve/lsmr.py), statement by
// statement: Golub-Kahan bidiagonalisation with the two plane rotations per step, the same estimates of |r|, |A^T r|,
// |A|, cond(A), |x| and the same stopping rules, so that the TRUNCATED iterate SciPy stops at is reproduced (its
// distance to the exact step is what separates the exact solver from the golden file: 5e-5 mm).  J_h is never formed:
// per observation the 2 x 9 analytic Jacobian is recomputed.  One persistent cooperative kernel per solve; each
// iteration is three passes over the points with a grid-wide barrier after each (|u|, A^T u + |v|, |x|); every block
// reduces the per-block partials in the same fixed order, so all blocks carry identical scalars and camera vectors
// (in shared memory) and take the same branches.  Point vectors live in the workspace (x = gnp).
namespace cg = cooperative_groups;

__device__ __forceinline__ void sym_ortho(double a, double b, double& c, double& s, double& r) {
  auto sgn = [](double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : 0.0); };
  if (b == 0.0) {
    c = sgn(a);
    s = 0.0;
    r = fabs(a);
  } else if (a == 0.0) {
    c = 0.0;
    s = sgn(b);
    r = fabs(b);
  } else if (fabs(b) > fabs(a)) {
    const double tau = a / b;
    s = sgn(b) / sqrt(1.0 + tau * tau);
    c = s * tau;
    r = b / s;
  } else {
    const double tau = b / a;
    c = sgn(a) / sqrt(1.0 + tau * tau);
    s = c * tau;
    r = a / c;
  }
}

struct LsmrScalars {
  double alpha, beta, inv_beta;
  double zetabar, alphabar, rho, rhobar, cbar, sbar;
  double betadd, betad, rhodold, tautildeold, thetatilde, zeta, d;
  double normA2, maxrbar, minrbar, normA, condA, normx, normr, normar, normb;
  double c_hbar, c_x, c_h;  // coefficients of this iteration's vector updates
  int itn, istop;
};

__global__ void __launch_bounds__(kBAThreads)
ba_lsmr_kernel(const double* __restrict__ intr, const double2* __restrict__ pts_xy, const double* __restrict__ pts3d, int C,
               int TJ, BAWorkspace ws, double atol, double btol, double conlim, int maxiter) {
  extern __shared__ double smem[];
  cg::grid_group grid = cg::this_grid();
  BAState* st = ws.state;
  if (st->done || !st->need_lin || st->solver != 1) return;  // uniform over the grid
  const int n6 = 6 * C;
  double* s_cam = smem;                       // staged cameras
  double* s_dc = s_cam + C * kCamStride;      // 1 / camera column norms
  double* s_vc = s_dc + n6;                   // camera parts of v, h, hbar, x (identical in every block)
  double* s_hc = s_vc + n6;
  double* s_hbc = s_hc + n6;
  double* s_xc = s_hbc + n6;
  double* s_red = s_xc + n6;                  // reduced values [kLsmrRed]
  double* s_tile = s_red + kLsmrRed;          // per-warp tiles [kBAWarps][kLsmrRed]
  __shared__ LsmrScalars S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double damp = sqrt(st->reg);
  if (threadIdx.x < C) stage_camera(ws.cam + threadIdx.x * 6, intr + threadIdx.x * 4, s_cam + threadIdx.x * kCamStride);
  if (threadIdx.x < n6) {
    s_dc[threadIdx.x] = 1.0 / ws.sinv_c[threadIdx.x];
    s_vc[threadIdx.x] = s_hc[threadIdx.x] = s_hbc[threadIdx.x] = s_xc[threadIdx.x] = 0.0;
  }
  __syncthreads();
  int flip = 0;
  // Grid-wide sum of entries [0, n) of the warp tiles -> s_red, in an order that is the same in every block: block
  // partial = tiles in warp order; after the grid barrier warp w of every block adds the blocks b = w, w + 4, ... (lane l
  // takes entries l and l + 32: the 43 values of a block are consecutive doubles), the four warp sums are added in
  // warp order.  Partials of other SMs are read through L2 (__ldcg); two buffers alternate so that a block that is
  // still reading the previous reduction is never overwritten.
  auto reduce = [&](int n) {
    __syncthreads();
    double* part = ws.lpart + ((size_t)flip * kBAMaxBlocks + blockIdx.x) * kLsmrRed;
    for (int e = threadIdx.x; e < n; e += kBAThreads) {
      double acc = s_tile[e];
#pragma unroll
      for (int w = 1; w < kBAWarps; ++w) acc += s_tile[w * kLsmrRed + e];
      part[e] = acc;
    }
    __threadfence();
    grid.sync();
    const double* all = ws.lpart + (size_t)flip * kBAMaxBlocks * kLsmrRed;
    if (n <= 2) {
      // two scalars: thread t adds the blocks t, t + 128, ... (independent L2 loads), then the 32 lanes of a warp are
      // added in lane order by a shuffle tree (fixed shape) and the four warps in warp order below
      double a0 = 0.0, a1 = 0.0;
      for (unsigned b = threadIdx.x; b < gridDim.x; b += kBAThreads) {
        a0 += __ldcg(all + (size_t)b * kLsmrRed);
        a1 += __ldcg(all + (size_t)b * kLsmrRed + 1);
      }
      a0 = warp_sum(a0);
      a1 = warp_sum(a1);
      if (lane == 0) {
        s_tile[warp * kLsmrRed] = a0;
        s_tile[warp * kLsmrRed + 1] = a1;
      }
    } else {
      double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
      for (unsigned b = warp; b < gridDim.x; b += kBAWarps) {
        if (lane < n) a0 += __ldcg(all + (size_t)b * kLsmrRed + lane);
        if (lane + 32 < n) a1 += __ldcg(all + (size_t)b * kLsmrRed + lane + 32);
      }
      s_tile[warp * kLsmrRed + lane] = a0;        // the tiles are free: their content went into `part`
      if (lane + 32 < kLsmrRed) s_tile[warp * kLsmrRed + lane + 32] = a1;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < n; e += kBAThreads) {
      double acc = s_tile[e];
#pragma unroll
      for (int w = 1; w < kBAWarps; ++w) acc += s_tile[w * kLsmrRed + e];
      s_red[e] = acc;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kBAWarps * kLsmrRed; e += kBAThreads) s_tile[e] = 0.0;
    flip ^= 1;
    __syncthreads();
  };
  for (int e = threadIdx.x; e < kBAWarps * kLsmrRed; e += kBAThreads) s_tile[e] = 0.0;
  __syncthreads();
  double* my = s_tile + warp * kLsmrRed;
  const int stride = gridDim.x * kBAThreads;
  // reduction slots: [0] |u|^2 or |v_p|^2, [1] |x_p|^2, [2, 2 + 6C) camera sums of A^T u
  constexpr int kCamSlot = 2;

  // ---- u = b = f (residuals), beta = |b|
  {
    double acc = 0.0;
    for (int g = blockIdx.x * kBAThreads + threadIdx.x; g < TJ; g += stride) {
      const double X[3] = {pts3d[(size_t)g * 3 + 0], pts3d[(size_t)g * 3 + 1], pts3d[(size_t)g * 3 + 2]};
      for (int c = 0; c < C; ++c) {
        const double2 xy = __ldg(pts_xy + (size_t)c * TJ + g);
        if (xy.x == 0.0 || xy.y == 0.0) continue;
        double r[2];
        project_residual(s_cam + c * kCamStride, X, xy.x, xy.y, r);
        ws.lu[((size_t)c * TJ + g) * 2 + 0] = r[0];
        ws.lu[((size_t)c * TJ + g) * 2 + 1] = r[1];
        acc += r[0] * r[0] + r[1] * r[1];
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) my[0] += acc;
  }
  reduce(1);
  if (threadIdx.x == 0) {
    S.normb = S.beta = sqrt(s_red[0]);
    S.inv_beta = S.beta > 0.0 ? 1.0 / S.beta : 0.0;
    S.alpha = 0.0;
    S.itn = 0;
    S.istop = 0;
  }
  __syncthreads();

  // one pass "v = A^T (u / beta) - beta v": point parts to ws.lv (not normalised yet), camera sums + |v_p|^2 reduced
  auto pass_atu = [&](double beta_old) {
    const double ib = S.inv_beta;
    for (int base = blockIdx.x * kBAThreads; base < TJ; base += stride) {
      const int g = base + threadIdx.x;
      const bool valid = g < TJ;
      double X[3] = {0, 0, 0}, tp[3] = {0, 0, 0};
      if (valid) {
        X[0] = pts3d[(size_t)g * 3 + 0];
        X[1] = pts3d[(size_t)g * 3 + 1];
        X[2] = pts3d[(size_t)g * 3 + 2];
      }
      for (int c = 0; c < C; ++c) {
        bool vis = false;
        double2 xy = make_double2(0.0, 0.0);
        if (valid) {
          xy = __ldg(pts_xy + (size_t)c * TJ + g);
          vis = (xy.x != 0.0) && (xy.y != 0.0);
        }
        if (!__any_sync(0xffffffffu, vis)) continue;
        double tc[6] = {0, 0, 0, 0, 0, 0};
        if (vis) {
          double r[2], Jc[2][6], Jp[2][3];
          project_jacobian(s_cam + c * kCamStride, X, xy.x, xy.y, r, Jc, Jp);
          const double u0 = ws.lu[((size_t)c * TJ + g) * 2 + 0] * ib, u1 = ws.lu[((size_t)c * TJ + g) * 2 + 1] * ib;
#pragma unroll
          for (int j = 0; j < 3; ++j) tp[j] += Jp[0][j] * u0 + Jp[1][j] * u1;
#pragma unroll
          for (int i = 0; i < 6; ++i) tc[i] = Jc[0][i] * u0 + Jc[1][i] * u1;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const double v = warp_sum(tc[i]);
          if (lane == 0) my[kCamSlot + c * 6 + i] += v;
        }
      }
      double nv = 0.0;
      if (valid) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const double old = beta_old != 0.0 ? ws.lv[(size_t)g * 3 + j] : 0.0;
          const double vn = tp[j] / ws.sinv_p[(size_t)g * 3 + j] - beta_old * old;
          ws.lv[(size_t)g * 3 + j] = vn;
          nv += vn * vn;
        }
      }
      nv = warp_sum(nv);
      if (lane == 0) my[0] += nv;
    }
    reduce(kCamSlot + n6);
    if (threadIdx.x < n6) s_vc[threadIdx.x] = s_dc[threadIdx.x] * s_red[kCamSlot + threadIdx.x] - beta_old * s_vc[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
      double a2 = s_red[0];
      for (int i = 0; i < n6; ++i) a2 += s_vc[i] * s_vc[i];
      S.alpha = sqrt(a2);
    }
    __syncthreads();
    if (threadIdx.x < n6 && S.alpha > 0.0) s_vc[threadIdx.x] *= 1.0 / S.alpha;
    __syncthreads();
  };

  // ---- v = A^T u / alpha ; h = v ; hbar = 0 ; x = 0
  // (every point vector is only ever touched by the thread that owns the point -- same g -> thread mapping in every
  // pass -- so the only data that crosses threads are the reduction partials)
  pass_atu(0.0);
  {
    const double ia = S.alpha > 0.0 ? 1.0 / S.alpha : 0.0;
    for (int g = blockIdx.x * kBAThreads + threadIdx.x; g < TJ; g += stride)
      for (int j = 0; j < 3; ++j) {
        const size_t i = (size_t)g * 3 + j;
        const double v = ws.lv[i] * ia;
        ws.lv[i] = v;
        ws.lh[i] = v;
        ws.lhb[i] = 0.0;
        ws.gnp[i] = 0.0;
      }
    if (threadIdx.x < n6) {
      s_hc[threadIdx.x] = s_vc[threadIdx.x];
      s_hbc[threadIdx.x] = 0.0;
      s_xc[threadIdx.x] = 0.0;
    }
  }
  if (threadIdx.x == 0) {
    S.zetabar = S.alpha * S.beta;
    S.alphabar = S.alpha;
    S.rho = S.rhobar = S.cbar = 1.0;
    S.sbar = 0.0;
    S.betadd = S.beta;
    S.betad = 0.0;
    S.rhodold = 1.0;
    S.tautildeold = S.thetatilde = S.zeta = S.d = 0.0;
    S.normA2 = S.alpha * S.alpha;
    S.maxrbar = 0.0;
    S.minrbar = 1e100;
    S.normA = sqrt(S.normA2);
    S.condA = 1.0;
    S.normx = 0.0;
    S.normr = S.beta;
    S.normar = S.alpha * S.beta;
    if (S.normar == 0.0 || S.normb == 0.0) S.istop = -1;  // x = 0 is the answer
  }
  __syncthreads();
  const double ctol = conlim > 0.0 ? 1.0 / conlim : 0.0;

  // Each iteration is TWO passes over the points: the vector updates of iteration k (h, hbar, x need this
  // iteration's rotations) ride in front of the "u = A v - alpha u" pass of iteration k + 1, whose reduction then
  // carries |x_k| as well, and iteration k's stopping rules are evaluated there.  When they fire, x is final (the u
  // that was just overwritten is not used again).
  bool pending = false;    // iteration S.itn's vector updates and stopping test are outstanding
  bool v_scaled = true;    // ws.lv holds v / alpha already
  while (S.istop == 0) {
    {
      const double alpha = S.alpha, ib = S.inv_beta;
      const double chb = S.c_hbar, cx = S.c_x, ch = S.c_h;
      const double ia = (!v_scaled && S.alpha > 0.0) ? 1.0 / S.alpha : 1.0;
      double acc = 0.0, accx = 0.0;
      for (int g = blockIdx.x * kBAThreads + threadIdx.x; g < TJ; g += stride) {
        const double X[3] = {pts3d[(size_t)g * 3 + 0], pts3d[(size_t)g * 3 + 1], pts3d[(size_t)g * 3 + 2]};
        double vp[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const size_t i = (size_t)g * 3 + j;
          double v = ws.lv[i];
          if (!v_scaled) {
            v *= ia;
            ws.lv[i] = v;
          }
          if (pending) {  // hbar = h + c_hbar hbar ; x += c_x hbar ; h = v + c_h h
            const double h = ws.lh[i];
            const double hb = ws.lhb[i] * chb + h;
            const double x = ws.gnp[i] + cx * hb;
            ws.lhb[i] = hb;
            ws.gnp[i] = x;
            ws.lh[i] = h * ch + v;
            accx += x * x;
          }
          vp[j] = v / ws.sinv_p[i];
        }
        for (int c = 0; c < C; ++c) {
          const double2 xy = __ldg(pts_xy + (size_t)c * TJ + g);
          if (xy.x == 0.0 || xy.y == 0.0) continue;
          double r[2], Jc[2][6], Jp[2][3];
          project_jacobian(s_cam + c * kCamStride, X, xy.x, xy.y, r, Jc, Jp);
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            double av = Jp[a][0] * vp[0] + Jp[a][1] * vp[1] + Jp[a][2] * vp[2];
#pragma unroll
            for (int i = 0; i < 6; ++i) av += Jc[a][i] * (s_dc[c * 6 + i] * s_vc[c * 6 + i]);
            const size_t o = ((size_t)c * TJ + g) * 2 + a;
            const double un = av - alpha * (ws.lu[o] * ib);
            ws.lu[o] = un;
            acc += un * un;
          }
        }
      }
      acc = warp_sum(acc);
      accx = warp_sum(accx);
      if (lane == 0) {
        my[0] += acc;
        my[1] += accx;
      }
      __syncthreads();  // every thread has read s_vc / the coefficients before the camera parts move on
      if (pending && threadIdx.x < n6) {
        const int i = threadIdx.x;
        const double hb = s_hbc[i] * chb + s_hc[i];
        s_hbc[i] = hb;
        s_xc[i] += cx * hb;
        s_hc[i] = s_hc[i] * ch + s_vc[i];
      }
    }
    reduce(2);
    if (threadIdx.x == 0) {
      if (pending) {  // iteration S.itn's stopping rules
        double x2 = s_red[1];
        for (int i = 0; i < n6; ++i) x2 += s_xc[i] * s_xc[i];
        S.normx = sqrt(x2);
        const double test1 = S.normr / S.normb;
        const double test2 = (S.normA * S.normr) != 0.0 ? S.normar / (S.normA * S.normr) : INFINITY;
        const double test3 = 1.0 / S.condA;
        const double t1 = test1 / (1.0 + S.normA * S.normx / S.normb);
        const double rtol = btol + atol * S.normA * S.normx / S.normb;
        int istop = 0;
        if (S.itn >= maxiter) istop = 7;
        if (1.0 + test3 <= 1.0) istop = 6;
        if (1.0 + test2 <= 1.0) istop = 5;
        if (1.0 + t1 <= 1.0) istop = 4;
        if (test3 <= ctol) istop = 3;
        if (test2 <= atol) istop = 2;
        if (test1 <= rtol) istop = 1;
        S.istop = istop;
      }
      if (S.istop == 0) {
        S.beta = sqrt(s_red[0]);
        S.inv_beta = S.beta > 0.0 ? 1.0 / S.beta : 0.0;
      }
    }
    __syncthreads();
    if (S.istop != 0) break;
    v_scaled = true;
    // ---- v = A^T u - beta v ; alpha = |v|
    if (S.beta > 0.0) {
      pass_atu(S.beta);
      v_scaled = false;
    }
    // ---- rotations and estimates of this iteration (every block, identical inputs)
    if (threadIdx.x == 0) {
      S.itn += 1;
      const double alpha = S.alpha, beta = S.beta;
      double chat, shat, alphahat, c, s, rho;
      sym_ortho(S.alphabar, damp, chat, shat, alphahat);
      const double rhoold = S.rho;
      sym_ortho(alphahat, beta, c, s, rho);
      const double thetanew = s * alpha;
      S.alphabar = c * alpha;
      const double rhobarold = S.rhobar, zetaold = S.zeta;
      const double thetabar = S.sbar * rho, rhotemp = S.cbar * rho;
      double cbar, sbar, rhobar;
      sym_ortho(S.cbar * rho, thetanew, cbar, sbar, rhobar);
      S.cbar = cbar;
      S.sbar = sbar;
      S.rhobar = rhobar;
      S.rho = rho;
      S.zeta = cbar * S.zetabar;
      S.zetabar = -sbar * S.zetabar;
      S.c_hbar = -(thetabar * rho / (rhoold * rhobarold));
      S.c_x = S.zeta / (rho * rhobar);
      S.c_h = -(thetanew / rho);
      // estimate of |r|
      const double betaacute = chat * S.betadd, betacheck = -shat * S.betadd;
      const double betahat = c * betaacute;
      S.betadd = -s * betaacute;
      const double thetatildeold = S.thetatilde;
      double ctildeold, stildeold, rhotildeold;
      sym_ortho(S.rhodold, thetabar, ctildeold, stildeold, rhotildeold);
      S.thetatilde = stildeold * rhobar;
      S.rhodold = ctildeold * rhobar;
      S.betad = -stildeold * S.betad + ctildeold * betahat;
      S.tautildeold = (zetaold - thetatildeold * S.tautildeold) / rhotildeold;
      const double taud = (S.zeta - S.thetatilde * S.tautildeold) / S.rhodold;
      S.d = S.d + betacheck * betacheck;
      S.normr = sqrt(S.d + (S.betad - taud) * (S.betad - taud) + S.betadd * S.betadd);
      // |A|, cond(A)
      S.normA2 = S.normA2 + beta * beta;
      S.normA = sqrt(S.normA2);
      S.normA2 = S.normA2 + alpha * alpha;
      S.maxrbar = fmax(S.maxrbar, rhobarold);
      if (S.itn > 1) S.minrbar = fmin(S.minrbar, rhobarold);
      S.condA = fmax(S.maxrbar, rhotemp) / fmin(S.minrbar, rhotemp);
      S.normar = fabs(S.zetabar);
    }
    __syncthreads();
    pending = true;
  }
  // ---- result: camera part of gn_h (point part is ws.gnp already)
  if (blockIdx.x == 0) {
    if (threadIdx.x < n6) {
      ws.gnc[threadIdx.x] = s_xc[threadIdx.x];
      ws.dcn[threadIdx.x] = s_dc[threadIdx.x] * s_xc[threadIdx.x];
    }
    if (threadIdx.x == 0) {
      st->lsmr_itn = S.itn;
      st->lsmr_istop = S.istop;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Candidate x + D step_h and its cost; the last block decides.  dynamic smem: cams (new) | cam_new [6C]
__global__ void __launch_bounds__(kBAThreads)
ba_step_kernel(const double* __restrict__ intr, const double2* __restrict__ pts_xy,
               const double* __restrict__ pts3d, int C, int TJ, BAWorkspace ws) {
  extern __shared__ double smem[];
  BAState* st = ws.state;
  if (st->done) return;
  double* s_new = smem;
  double* s_cn = smem + C * kCamStride;
  double* s_t = s_cn + 6 * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double cg = st->coef_g, cn = st->coef_gn;
  if (threadIdx.x < 6 * C)
    s_cn[threadIdx.x] = ws.cam[threadIdx.x] + (cg * ws.ghc[threadIdx.x] + cn * ws.gnc[threadIdx.x]) / ws.sinv_c[threadIdx.x];
  if (threadIdx.x < kBAWarps * 3) s_t[threadIdx.x] = 0.0;
  __syncthreads();
  if (threadIdx.x < C) stage_camera(s_cn + threadIdx.x * 6, intr + threadIdx.x * 4, s_new + threadIdx.x * kCamStride);
  __syncthreads();

  double cost = 0.0, stepsq = 0.0, xsq = 0.0;
  for (int g = (blockIdx.x + ws.vb0) * kBAThreads + threadIdx.x; g < TJ; g += ws.vgrid * kBAThreads) {
    double Xn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double x = pts3d[(size_t)g * 3 + i];
      const double s = ws.sinv_p[(size_t)g * 3 + i];
      const double step = (cg * (ws.gp[(size_t)g * 3 + i] / s) + cn * ws.gnp[(size_t)g * 3 + i]) / s;
      Xn[i] = x + step;
      stepsq += step * step;
      xsq += x * x;
      ws.X_new[(size_t)g * 3 + i] = Xn[i];
    }
    for (int c = 0; c < C; ++c) {
      const double2 xy = __ldg(pts_xy + (size_t)c * TJ + g);
      if (xy.x == 0.0 || xy.y == 0.0) continue;
      double r[2];
      project_residual(s_new + c * kCamStride, Xn, xy.x, xy.y, r);
      cost += r[0] * r[0] + r[1] * r[1];
    }
  }
  cost = warp_sum(cost);
  stepsq = warp_sum(stepsq);
  xsq = warp_sum(xsq);
  if (lane == 0) {
    s_t[warp * 3 + 0] = 0.5 * cost;
    s_t[warp * 3 + 1] = stepsq;
    s_t[warp * 3 + 2] = xsq;
  }
  block_partial(s_t, 3, 3, ws.partials + (size_t)(blockIdx.x + ws.vb0) * 3);  // ba_finish_kernel(3) goes on from here
}

// ---- after the reduction, one thread: trf_no_bounds' inner loop body after f_new = fun(x_new); s_cn = the candidate cameras
__device__ void step_tail(int C, BAWorkspace ws, const double* s_cn) {
  BAState* st = ws.state;
  const double Fn = ws.red[0];
  double stepsq_t = ws.red[1], xsq_t = ws.red[2];
  for (int i = 0; i < 6 * C; ++i) {
    const double d = s_cn[i] - ws.cam[i];
    stepsq_t += d * d;
    xsq_t += ws.cam[i] * ws.cam[i];
  }
  st->iter += 1;
  st->accept_flag = 0;
  st->need_lin = 0;
  if (!isfinite(Fn)) {
    st->Delta = 0.25 * st->sh_norm;
  } else {
    const double actual = st->F - Fn, pred = st->pred;
    double ratio;
    if (pred > 0.0)
      ratio = actual / pred;
    else if (pred == 0.0 && actual == 0.0)
      ratio = 1.0;
    else
      ratio = 0.0;
    double Delta_new = st->Delta;
    if (ratio < 0.25)
      Delta_new = 0.25 * st->sh_norm;
    else if (ratio > 0.75 && st->sh_norm > 0.95 * st->Delta)
      Delta_new = 2.0 * st->Delta;
    const bool f_ok = actual < st->ftol * st->F && ratio > 0.25;
    const bool x_ok = sqrt(stepsq_t) < st->xtol * (st->xtol + sqrt(xsq_t));
    const int term = (f_ok && x_ok) ? 4 : f_ok ? 2 : x_ok ? 3 : 0;
    if (term) {
      st->done = 1;
      st->status = term;
    } else {
      st->Delta = Delta_new;
    }
    if (actual > 0.0) {
      st->F = Fn;
      st->accepted += 1;
      st->accept_flag = st->iter;
      st->need_lin = 1;
      for (int i = 0; i < 6 * C; ++i) ws.cam[i] = s_cn[i];
    }
  }
  if (st->iter >= st->max_iters) st->done = 1;
  if (!st->done && !st->need_lin) solve_subproblem(st);  // rejected: same model, smaller region
}

// accept_flag holds the number of the iteration whose candidate was accepted (0: none)
// (frame-sharded: only the points of this rank's blocks have a candidate; `nb` = blocks of this rank)
__global__ void ba_apply_points_kernel(int n, int iter, BAWorkspace ws, double* __restrict__ pts3d, int nb) {
  if (ws.state->accept_flag != iter) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (ws.sharded) {
      const int vb = ((i / 3) / kBAThreads) % ws.vgrid;  // the block whose grid-stride loop owns this point
      if (vb < ws.vb0 || vb >= ws.vb0 + nb) continue;
    }
    pts3d[i] = ws.X_new[i];
  }
}

// Behind every per-point pass (and, frame-sharded, behind the all-gather of the partials of every rank): the sum over ALL
// blocks in a fixed order that depends on the number of blocks only -- kFinChunks runs of consecutive blocks per entry,
// each summed in block order by one thread (eight loads in flight), the runs then added in run order --, 16 entries per
// block of 128 threads; the last block to finish runs the scalar logic of the pass.  pass 0 gradient, 1 Schur system,
// 2 back-substitution, 3 step.  grid = ceil(entries / kFinEntries).
constexpr int kFinChunks = 8, kFinEntries = kBAThreads / kFinChunks;

__global__ void __launch_bounds__(kBAThreads) ba_finish_kernel(int pass, int C, BAWorkspace ws) {
  __shared__ double s_cn[kMaxN];
  __shared__ double s_run[kFinChunks][kFinEntries];
  __shared__ bool s_last;
  BAState* st = ws.state;
  if (st->done) return;
  if (pass != 3 && !st->need_lin) return;
  if (pass == 1 && st->solver == 1) return;
  const int n = pass == 0 ? g_doubles(C) : pass == 1 ? sys_doubles(C) : pass == 2 ? 4 : 3;
  const int n_sum = pass == 0 ? n - 1 : n;
  const int el = threadIdx.x % kFinEntries, run = threadIdx.x / kFinEntries;
  const int e = blockIdx.x * kFinEntries + el;
  const int len = (ws.vgrid + kFinChunks - 1) / kFinChunks;
  const int b0 = run * len, b1 = min(ws.vgrid, b0 + len);
  const bool is_sum = e < n_sum;
  double acc = is_sum ? 0.0 : -1.0;
  if (e < n) {
    const double* part = ws.partials + e;
    int b = b0;
    for (; b + 8 <= b1; b += 8) {
      double v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldcg(part + (size_t)(b + k) * n);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = is_sum ? acc + v[k] : fmax(acc, v[k]);
    }
    for (; b < b1; ++b) {
      const double v = __ldcg(part + (size_t)b * n);
      acc = is_sum ? acc + v : fmax(acc, v);
    }
  }
  s_run[run][el] = acc;
  __syncthreads();
  if (run == 0 && e < n) {
    double tot = s_run[0][el];
#pragma unroll
    for (int r = 1; r < kFinChunks; ++r) tot = is_sum ? tot + s_run[r][el] : fmax(tot, s_run[r][el]);
    ws.red[e] = tot;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(ws.counters + pass, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;   // (state is only written below, after every block has read it above and taken its ticket)
  __threadfence();
  if (pass == 3) {  // the candidate cameras, as ba_step_kernel builds them
    const double cg = st->coef_g, cn = st->coef_gn;
    if (threadIdx.x < 6 * C)
      s_cn[threadIdx.x] = ws.cam[threadIdx.x] + (cg * ws.ghc[threadIdx.x] + cn * ws.gnc[threadIdx.x]) / ws.sinv_c[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  ws.counters[pass] = 0u;
  if (pass == 0) gradient_tail(C, ws);
  if (pass == 2) backsub_tail(C, ws);
  if (pass == 3) step_tail(C, ws, s_cn);
}

__global__ void ba_end_kernel(double* __restrict__ cam_rt, int C, BAWorkspace ws, df3d_ba_report* report) {
  if (threadIdx.x < 6 * C) cam_rt[threadIdx.x] = ws.cam[threadIdx.x];
  if (threadIdx.x == 0 && report) {
    const BAState s = *ws.state;
    report->cost0 = s.F0;
    report->cost = s.F;
    report->reg = s.reg;
    report->iters = s.iter;
    report->accepted = s.accepted;
    report->n_obs = s.n_obs;
    report->status = s.status;
    report->lsmr_itn = s.lsmr_itn;
    report->lsmr_istop = s.lsmr_istop;
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

static int ba_finish_grid(int pass, int C) {
  const int n = pass == 0 ? g_doubles(C) : pass == 1 ? sys_doubles(C) : pass == 2 ? 4 : 3;
  return ceil_div(n, kFinEntries);
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

extern "C" int df3d_bundle_adjust_launches(const df3d_ba_opts* opts) {
  const int iters = opts ? opts->max_iters : 20;
  const int per_iter = (opts && opts->solver == 0) ? 10 : 8;  // gradient, finish, (schur, finish, solve | lsmr), backsub, finish, step, finish, apply
  return 2 + per_iter * iters;
}

extern "C" int df3d_bundle_adjust(double* cam_rt_dev, const double* intr_dev, const double* pts_xy_dev,
                                  int C, int T, int J, const df3d_ba_opts* opts, double* pts3d_dev,
                                  df3d_ba_report* report_dev, void* workspace_dev, size_t workspace_bytes,
                                  void* stream) {
  if (int e = check_common("df3d_bundle_adjust", C, T, J)) return e;
  DF3D_REQUIRE(cam_rt_dev && intr_dev && pts_xy_dev && pts3d_dev, DF3D_EINVAL, "df3d_bundle_adjust: null pointer");
  DF3D_REQUIRE((reinterpret_cast<uintptr_t>(pts_xy_dev) & 15) == 0, DF3D_EINVAL, "df3d_bundle_adjust: pts_xy must be 16-byte aligned");
  DF3D_REQUIRE(workspace_dev, DF3D_EINVAL, "df3d_bundle_adjust: null workspace");
  DF3D_REQUIRE((reinterpret_cast<uintptr_t>(workspace_dev) & 255) == 0, DF3D_EINVAL, "df3d_bundle_adjust: workspace must be 256-byte aligned");
  BAWorkspace ws;
  const size_t need = ba_workspace_layout(C, T, J, static_cast<char*>(workspace_dev), &ws);
  DF3D_REQUIRE(workspace_bytes >= need, DF3D_ENOMEM, "df3d_bundle_adjust: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
  df3d_ba_opts o{20, 1e-4, 1e-8, 1e-8, 1};
  if (opts) o = *opts;
  DF3D_REQUIRE(o.max_iters >= 1 && o.max_iters <= 1000 && o.ftol >= 0.0 && o.xtol >= 0.0 && o.gtol >= 0.0 &&
                   (o.solver == 0 || o.solver == 1),
               DF3D_EINVAL, "df3d_bundle_adjust: bad options (max_iters in [1,1000], tolerances >= 0, solver 0 or 1)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int TJ = T * J;
  const int grid = ba_grid(TJ);
  const size_t smem_g = ((size_t)C * kCamStride + (size_t)kBAWarps * g_doubles(C)) * sizeof(double);
  const size_t smem_s = ((size_t)C * kCamStride + (size_t)kBAWarps * sys_doubles(C)) * sizeof(double);
  const size_t smem_b = ((size_t)C * kCamStride + 12 * C + kBAWarps * 4) * sizeof(double);
  const size_t smem_e = ((size_t)C * kCamStride + 6 * C + kBAWarps * 3) * sizeof(double);
  DF3D_CUDA(cudaFuncSetAttribute(ba_schur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
  const double2* xy = reinterpret_cast<const double2*>(pts_xy_dev);
  // solver 1 (LSMR like SciPy): a persistent cooperative kernel, every block resident at once
  const size_t smem_l = ((size_t)C * kCamStride + 5 * 6 * C + kLsmrRed + kBAWarps * kLsmrRed) * sizeof(double);
  int lsmr_grid = grid;
  double lsmr_tol = 1e-6, lsmr_conlim = 1e8;
  long long lsmr_rows = 2ll * C * TJ, lsmr_cols = 6ll * C + 3ll * TJ;
  int lsmr_maxiter = (int)(lsmr_rows < lsmr_cols ? lsmr_rows : lsmr_cols);
  if (o.solver == 1) {
    int dev = 0, sms = 0, per_sm = 0;
    DF3D_CUDA(cudaGetDevice(&dev));
    DF3D_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DF3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ba_lsmr_kernel, kBAThreads, smem_l));
    DF3D_REQUIRE(per_sm >= 1, DF3D_EUNSUPPORTED, "df3d_bundle_adjust: the LSMR kernel does not fit on this device");
    if (lsmr_grid > per_sm * sms) lsmr_grid = per_sm * sms;
  }
  ba_begin_kernel<<<1, 64, 0, s>>>(cam_rt_dev, C, o, ws);
  DF3D_LAUNCH_CHECK("ba_begin_kernel");
  const int n3 = TJ * 3;
  int agrid = ceil_div(n3, 256);
  if (agrid > 4 * kBAMaxBlocks) agrid = 4 * kBAMaxBlocks;
  // fixed launch sequence; kernels become no-ops once the device-side state says `done`
  for (int it = 0; it < o.max_iters; ++it) {
    ba_gradient_kernel<<<grid, kBAThreads, smem_g, s>>>(intr_dev, xy, pts3d_dev, C, TJ, ws);
    ba_finish_kernel<<<ba_finish_grid(0, C), kBAThreads, 0, s>>>(0, C, ws);
    if (o.solver == 1) {
      void* args[] = {(void*)&intr_dev, (void*)&xy, (void*)&pts3d_dev, (void*)&C, (void*)&TJ, (void*)&ws, (void*)&lsmr_tol,
                      (void*)&lsmr_tol, (void*)&lsmr_conlim, (void*)&lsmr_maxiter};
      DF3D_CUDA(cudaLaunchCooperativeKernel((const void*)ba_lsmr_kernel, dim3(lsmr_grid), dim3(kBAThreads), args, smem_l, s));
    } else {
      ba_schur_kernel<<<grid, kBAThreads, smem_s, s>>>(intr_dev, xy, pts3d_dev, C, TJ, ws);
      ba_finish_kernel<<<ba_finish_grid(1, C), kBAThreads, 0, s>>>(1, C, ws);
      ba_solve_kernel<<<1, kSolveThreads, 0, s>>>(C, ws);
    }
    ba_backsub_kernel<<<grid, kBAThreads, smem_b, s>>>(intr_dev, xy, pts3d_dev, C, TJ, ws);
    ba_finish_kernel<<<ba_finish_grid(2, C), kBAThreads, 0, s>>>(2, C, ws);
    ba_step_kernel<<<grid, kBAThreads, smem_e, s>>>(intr_dev, xy, pts3d_dev, C, TJ, ws);
    ba_finish_kernel<<<ba_finish_grid(3, C), kBAThreads, 0, s>>>(3, C, ws);
    ba_apply_points_kernel<<<agrid, 256, 0, s>>>(n3, it + 1, ws, pts3d_dev, 0);
    DF3D_LAUNCH_CHECK("bundle adjustment iteration");
  }
  ba_end_kernel<<<1, 64, 0, s>>>(cam_rt_dev, C, ws, report_dev);
  DF3D_LAUNCH_CHECK("ba_end_kernel");
  return DF3D_OK;
}

// ---------------------------------------------------------------------------------------------
// Frame-sharded form of df3d_bundle_adjust (solver 0), see include/df3d_b200.h.
static int ba_pass_doubles(int pass, int C) { return pass == 0 ? g_doubles(C) : pass == 1 ? sys_doubles(C) : pass == 2 ? 4 : 3; }

extern "C" int df3d_ba_sharded_plan(int C, int T, int J, int world, int* n_blocks, size_t* partials_offset, int* pass_doubles) {
  if (int e = check_common("df3d_ba_sharded_plan", C, T, J)) return e;
  DF3D_REQUIRE(n_blocks && partials_offset && pass_doubles && world >= 1, DF3D_EINVAL, "df3d_ba_sharded_plan: bad arguments");
  const int grid = ba_grid(T * J);
  DF3D_REQUIRE(grid % world == 0, DF3D_EUNSUPPORTED,
               "df3d_ba_sharded_plan: %d blocks of points do not split evenly over %d ranks (solve replicated instead)", grid, world);
  BAWorkspace ws;
  ba_workspace_layout(C, T, J, reinterpret_cast<char*>(uintptr_t(256)), &ws);  // offsets relative to a fictitious base
  *n_blocks = grid;
  *partials_offset = reinterpret_cast<uintptr_t>(ws.partials) - 256;
  for (int p = 0; p < 4; ++p) pass_doubles[p] = ba_pass_doubles(p, C);
  return DF3D_OK;
}

static int ba_sharded_ws(const char* fn, int C, int T, int J, void* workspace_dev, size_t workspace_bytes, int rank, int world,
                         BAWorkspace* ws, int* nb) {
  if (int e = check_common(fn, C, T, J)) return e;
  DF3D_REQUIRE(workspace_dev && (reinterpret_cast<uintptr_t>(workspace_dev) & 255) == 0, DF3D_EINVAL, "%s: workspace must be 256-byte aligned", fn);
  const size_t need = ba_workspace_layout(C, T, J, static_cast<char*>(workspace_dev), ws);
  DF3D_REQUIRE(workspace_bytes >= need, DF3D_ENOMEM, "%s: workspace too small (%zu < %zu bytes)", fn, workspace_bytes, need);
  DF3D_REQUIRE(world >= 1 && rank >= 0 && rank < world && ws->vgrid % world == 0, DF3D_EINVAL, "%s: rank %d of %d over %d blocks", fn, rank,
               world, ws->vgrid);
  *nb = ws->vgrid / world;
  ws->vb0 = rank * *nb;
  ws->sharded = 1;
  return DF3D_OK;
}

extern "C" int df3d_ba_sharded_begin(const double* cam_rt_dev, int C, int T, int J, const df3d_ba_opts* opts, void* workspace_dev,
                                     size_t workspace_bytes, void* stream) {
  BAWorkspace ws;
  int nb;
  if (int e = ba_sharded_ws("df3d_ba_sharded_begin", C, T, J, workspace_dev, workspace_bytes, 0, 1, &ws, &nb)) return e;
  DF3D_REQUIRE(cam_rt_dev, DF3D_EINVAL, "df3d_ba_sharded_begin: null pointer");
  df3d_ba_opts o{20, 1e-4, 1e-8, 1e-8, 0};
  if (opts) o = *opts;
  DF3D_REQUIRE(o.max_iters >= 1 && o.max_iters <= 1000 && o.ftol >= 0.0 && o.xtol >= 0.0 && o.gtol >= 0.0 && o.solver == 0, DF3D_EINVAL,
               "df3d_ba_sharded_begin: bad options (max_iters in [1,1000], tolerances >= 0, solver 0: the LSMR solver runs replicated)");
  ba_begin_kernel<<<1, 64, 0, static_cast<cudaStream_t>(stream)>>>(cam_rt_dev, C, o, ws);
  DF3D_LAUNCH_CHECK("ba_begin_kernel");
  return DF3D_OK;
}

extern "C" int df3d_ba_sharded_pass(int pass, int rank, int world, const double* intr_dev, const double* pts_xy_dev, const double* pts3d_dev,
                                    int C, int T, int J, void* workspace_dev, size_t workspace_bytes, void* stream) {
  BAWorkspace ws;
  int nb;
  if (int e = ba_sharded_ws("df3d_ba_sharded_pass", C, T, J, workspace_dev, workspace_bytes, rank, world, &ws, &nb)) return e;
  DF3D_REQUIRE(pass >= 0 && pass <= 3 && intr_dev && pts_xy_dev && pts3d_dev, DF3D_EINVAL, "df3d_ba_sharded_pass: bad arguments");
  DF3D_REQUIRE((reinterpret_cast<uintptr_t>(pts_xy_dev) & 15) == 0, DF3D_EINVAL, "df3d_ba_sharded_pass: pts_xy must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int TJ = T * J;
  const double2* xy = reinterpret_cast<const double2*>(pts_xy_dev);
  const size_t smem_g = ((size_t)C * kCamStride + (size_t)kBAWarps * g_doubles(C)) * sizeof(double);
  const size_t smem_s = ((size_t)C * kCamStride + (size_t)kBAWarps * sys_doubles(C)) * sizeof(double);
  const size_t smem_b = ((size_t)C * kCamStride + 12 * C + kBAWarps * 4) * sizeof(double);
  const size_t smem_e = ((size_t)C * kCamStride + 6 * C + kBAWarps * 3) * sizeof(double);
  if (pass == 0) ba_gradient_kernel<<<nb, kBAThreads, smem_g, s>>>(intr_dev, xy, pts3d_dev, C, TJ, ws);
  if (pass == 1) {
    DF3D_CUDA(cudaFuncSetAttribute(ba_schur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
    ba_schur_kernel<<<nb, kBAThreads, smem_s, s>>>(intr_dev, xy, pts3d_dev, C, TJ, ws);
  }
  if (pass == 2) ba_backsub_kernel<<<nb, kBAThreads, smem_b, s>>>(intr_dev, xy, pts3d_dev, C, TJ, ws);
  if (pass == 3) ba_step_kernel<<<nb, kBAThreads, smem_e, s>>>(intr_dev, xy, pts3d_dev, C, TJ, ws);
  DF3D_LAUNCH_CHECK("bundle adjustment pass");
  return DF3D_OK;
}

// after the all-gather of the pass's partials: the fixed-order sum over ALL blocks and the scalar logic behind it; behind
// pass 1 the reduced camera system is solved, behind pass 3 (of iteration `iter`, counted from 1) an accepted candidate
// replaces this rank's points
extern "C" int df3d_ba_sharded_finish(int pass, int iter, int rank, int world, double* pts3d_dev, int C, int T, int J, void* workspace_dev,
                                      size_t workspace_bytes, void* stream) {
  BAWorkspace ws;
  int nb;
  if (int e = ba_sharded_ws("df3d_ba_sharded_finish", C, T, J, workspace_dev, workspace_bytes, rank, world, &ws, &nb)) return e;
  DF3D_REQUIRE(pass >= 0 && pass <= 3 && pts3d_dev, DF3D_EINVAL, "df3d_ba_sharded_finish: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ba_finish_kernel<<<ba_finish_grid(pass, C), kBAThreads, 0, s>>>(pass, C, ws);
  if (pass == 1) ba_solve_kernel<<<1, kSolveThreads, 0, s>>>(C, ws);
  if (pass == 3) {
    const int n3 = T * J * 3;
    int agrid = ceil_div(n3, 256);
    if (agrid > 4 * kBAMaxBlocks) agrid = 4 * kBAMaxBlocks;
    ba_apply_points_kernel<<<agrid, 256, 0, s>>>(n3, iter, ws, pts3d_dev, nb);
  }
  DF3D_LAUNCH_CHECK("bundle adjustment finish");
  return DF3D_OK;
}

extern "C" int df3d_ba_sharded_end(double* cam_rt_dev, int C, int T, int J, df3d_ba_report* report_dev, void* workspace_dev,
                                   size_t workspace_bytes, void* stream) {
  BAWorkspace ws;
  int nb;
  if (int e = ba_sharded_ws("df3d_ba_sharded_end", C, T, J, workspace_dev, workspace_bytes, 0, 1, &ws, &nb)) return e;
  DF3D_REQUIRE(cam_rt_dev, DF3D_EINVAL, "df3d_ba_sharded_end: null pointer");
  ba_end_kernel<<<1, 64, 0, static_cast<cudaStream_t>(stream)>>>(cam_rt_dev, C, ws, report_dev);
  DF3D_LAUNCH_CHECK("ba_end_kernel");
  return DF3D_OK;
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
