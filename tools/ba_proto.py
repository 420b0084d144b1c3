"""CPU prototype of the device bundle-adjustment step rule (numpy, vectorised) next to the SciPy-TRF
oracle, on the SURVEY 8(d) config-3 synthetic geometry.  Development tool: it is how the solver in
csrc/bundle_adjust.cu was chosen and how its distance to the oracle was measured before going to
the GPU.  Not imported by the package or the tests.

    python tools/ba_proto.py --frames 256 [--mode lm|trf]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import geometry as g  # noqa: E402
from oracle import synth  # noqa: E402


def rodrigues_batch(r):
    out = np.empty((r.shape[0], 3, 3))
    for i in range(r.shape[0]):
        out[i] = g.rodrigues(r[i])
    return out


def jacobians(cam, intr, X, cam_idx, pt_idx, obs):
    """Analytic residuals + Jacobians per observation: r (N,2), Jc (N,2,6), Jp (N,2,3)."""
    C = cam.shape[0]
    R = rodrigues_batch(cam[:, :3])
    # M = (r r^T + (R^T - I)[r]x) / |r|^2   (d(RX)/dr = -R [X]x M)
    M = np.empty((C, 3, 3))
    for c in range(C):
        r = cam[c, :3]
        th2 = r @ r
        if th2 < 1e-24:
            M[c] = np.eye(3)
        else:
            K = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
            M[c] = (np.outer(r, r) + (R[c].T - np.eye(3)) @ K) / th2
    Rk, Mk, tk = R[cam_idx], M[cam_idx], cam[cam_idx, 3:]
    Xk = X[pt_idx]
    Xc = np.einsum("nij,nj->ni", Rk, Xk) + tk
    iz = 1.0 / Xc[:, 2]
    x, y = Xc[:, 0] * iz, Xc[:, 1] * iz
    fx, fy, cx, cy = intr[cam_idx, 0, 0], intr[cam_idx, 1, 1], intr[cam_idx, 0, 2], intr[cam_idx, 1, 2]
    res = np.stack([fx * x + cx - obs[:, 0], fy * y + cy - obs[:, 1]], axis=1)
    dp = np.zeros((len(x), 2, 3))
    dp[:, 0, 0] = fx * iz
    dp[:, 0, 2] = -fx * x * iz
    dp[:, 1, 1] = fy * iz
    dp[:, 1, 2] = -fy * y * iz
    Jp = np.einsum("nai,nij->naj", dp, Rk)
    JX = np.cross(Jp, Xk[:, None, :])  # Jp [X]x rows: (Jp x X)
    Jr = -np.einsum("nai,nij->naj", JX, Mk)
    Jc = np.concatenate([Jr, dp], axis=2)
    return res, Jc, Jp


def residuals(cam, intr, X, cam_idx, pt_idx, obs):
    R = rodrigues_batch(cam[:, :3])
    Xc = np.einsum("nij,nj->ni", R[cam_idx], X[pt_idx]) + cam[cam_idx, 3:]
    x, y = Xc[:, 0] / Xc[:, 2], Xc[:, 1] / Xc[:, 2]
    fx, fy, cx, cy = intr[cam_idx, 0, 0], intr[cam_idx, 1, 1], intr[cam_idx, 0, 2], intr[cam_idx, 1, 2]
    return np.stack([fx * x + cx - obs[:, 0], fy * y + cy - obs[:, 1]], axis=1)


def _pairs(pt_idx):
    """Index pairs (a, b) of observations that share a point (including a == b)."""
    order = np.argsort(pt_idx, kind="stable")
    po = pt_idx[order]
    starts = np.flatnonzero(np.r_[True, po[1:] != po[:-1]])
    ends = np.r_[starts[1:], len(po)]
    maxv = (ends - starts).max()
    oa_all, ob_all = [], []
    for a in range(maxv):
        ia = starts + a
        for b in range(maxv):
            ib = starts + b
            ok = (ia < ends) & (ib < ends)
            oa_all.append(order[ia[ok]])
            ob_all.append(order[ib[ok]])
    return np.concatenate(oa_all), np.concatenate(ob_all)


def solve_tr_2d(B, gS, Delta):
    """min 0.5 p^T B p + g^T p  s.t. |p| <= Delta in two dimensions (what scipy's solve_trust_region_2d
    returns), through the eigen-decomposition of B and the secular equation."""
    w, Q = np.linalg.eigh(B)
    gq = Q.T @ gS
    if w[0] > 0:
        p = -gq / w
        if p @ p <= Delta ** 2:
            return Q @ p
    # boundary: |(B + mu I)^-1 g| = Delta, mu > max(0, -w0)
    lo = max(0.0, -w[0])
    f = lambda mu: np.sqrt(np.sum((gq / (w + mu)) ** 2)) - Delta
    hi = lo + 1.0
    while f(hi) > 0:
        hi = lo + 2 * (hi - lo)
    a, b = lo, hi
    for _ in range(200):
        mid = 0.5 * (a + b)
        if f(mid) > 0:
            a = mid
        else:
            b = mid
    mu = 0.5 * (a + b)
    return Q @ (-gq / (w + mu))


def trf_schur(cam0, intr, X0, cam_idx, pt_idx, obs, ftol=1e-4, xtol=1e-8, gtol=1e-8, max_iters=50, verbose=True):
    """scipy.optimize._lsq.trf.trf_no_bounds (tr_solver='lsmr', x_scale='jac') with the LSMR solve of the
    regularised Gauss-Newton step replaced by the exact Schur-complement solve -- the device algorithm."""
    C, NP = cam0.shape[0], X0.shape[0]
    n6 = 6 * C
    cam, X = cam0.copy(), X0.copy()
    sinv_c = sinv_p = None
    oa, ob = _pairs(pt_idx)
    r, Jc, Jp = jacobians(cam, intr, X, cam_idx, pt_idx, obs)
    F = 0.5 * np.sum(r * r)
    Delta = None
    nfev = 1
    status = 0
    for it in range(max_iters):
        # ---- K1: gradient, column norms
        U = np.zeros((C, 6, 6))
        np.add.at(U, cam_idx, np.einsum("nai,naj->nij", Jc, Jc))
        gc = np.zeros((C, 6))
        np.add.at(gc, cam_idx, np.einsum("nai,na->ni", Jc, r))
        V = np.zeros((NP, 3, 3))
        np.add.at(V, pt_idx, np.einsum("nai,naj->nij", Jp, Jp))
        gp = np.zeros((NP, 3))
        np.add.at(gp, pt_idx, np.einsum("nai,na->ni", Jp, r))
        cn_c = np.sqrt(np.einsum("cii->ci", U)).ravel()
        cn_p = np.sqrt(np.einsum("nii->ni", V))
        if sinv_c is None:
            sinv_c = np.where(cn_c == 0, 1.0, cn_c)
            sinv_p = np.where(cn_p == 0, 1.0, cn_p)
            Delta = np.sqrt(np.sum((cam.ravel() * sinv_c) ** 2) + np.sum((X * sinv_p) ** 2))
            if Delta == 0:
                Delta = 1.0
        else:
            sinv_c = np.maximum(sinv_c, cn_c)
            sinv_p = np.maximum(sinv_p, cn_p)
        if max(np.abs(gc).max(), np.abs(gp).max()) < gtol:
            status = 1
            break
        dc, dp = 1.0 / sinv_c, 1.0 / sinv_p
        ghc, ghp = dc * gc.ravel(), dp * gp
        # ---- K2: regularisation from the Cauchy step
        Jg = np.einsum("nai,ni->na", Jc, (dc * ghc).reshape(C, 6)[cam_idx]) + np.einsum("nai,ni->na", Jp, (dp * ghp)[pt_idx])
        gg = ghc @ ghc + np.sum(ghp * ghp)
        JgJg = np.sum(Jg * Jg)
        a_, b_ = 0.5 * JgJg, -gg
        to_tr = Delta / np.sqrt(gg)
        ts = [0.0, to_tr]
        if a_ != 0:
            ext = -0.5 * b_ / a_
            if 0 < ext < to_tr:
                ts.append(ext)
        ag_value = min(t * (a_ * t + b_) for t in ts)
        reg = -ag_value / Delta ** 2
        # ---- K3/K4: Schur complement of (J_h^T J_h + reg I) gn_h = J_h^T f
        Vh = dp[:, :, None] * V * dp[:, None, :] + reg * np.eye(3)
        Mm = dp[:, :, None] * np.linalg.inv(Vh) * dp[:, None, :]
        W = np.einsum("nai,naj->nij", Jc, Jp)
        WM = np.einsum("nij,njk->nik", W, Mm[pt_idx])
        bt = np.zeros((C, 6))
        np.add.at(bt, cam_idx, np.einsum("nik,nk->ni", WM, gp[pt_idx]))
        S = np.zeros((n6, n6))
        blk = np.einsum("nik,njk->nij", WM[oa], W[ob])
        idx_r = cam_idx[oa][:, None] * 6 + np.arange(6)[None, :]
        idx_c = cam_idx[ob][:, None] * 6 + np.arange(6)[None, :]
        np.add.at(S, (idx_r[:, :, None], idx_c[:, None, :]), blk)
        A = -S
        for c in range(C):
            A[6 * c:6 * c + 6, 6 * c:6 * c + 6] += U[c]
        A = dc[:, None] * A * dc[None, :] + reg * np.eye(n6)
        rhs = dc * (gc.ravel() - bt.ravel())
        gnc = np.linalg.solve(A, rhs)                     # scaled camera part of gn_h
        # ---- K5: back-substitution (scaled point part), Gram quantities
        q = r - np.einsum("nai,ni->na", Jc, (dc * gnc).reshape(C, 6)[cam_idx])
        rp = np.zeros((NP, 3))
        np.add.at(rp, pt_idx, np.einsum("nai,na->ni", Jp, q))
        gnp = np.einsum("nij,nj->ni", Mm, rp) / dp         # scaled point part: M = D (..)^-1 D -> divide one D out
        Jgn = np.einsum("nai,ni->na", Jc, (dc * gnc).reshape(C, 6)[cam_idx]) + np.einsum("nai,ni->na", Jp, (dp * gnp)[pt_idx])
        g_gn = ghc @ gnc + np.sum(ghp * gnp)
        gn_gn = gnc @ gnc + np.sum(gnp * gnp)
        Jg_Jgn = np.sum(Jg * Jgn)
        Jgn_Jgn = np.sum(Jgn * Jgn)
        # orthonormal basis of span{g_h, gn_h} by Gram-Schmidt on the Gram matrix: q1 = g/|g|, q2 = (gn - (q1.gn) q1)/|.|
        n1 = np.sqrt(gg)
        c12 = g_gn / n1
        n2 = np.sqrt(max(gn_gn - c12 * c12, 0.0))
        # q1 = Tm[0,0] g ; q2 = Tm[1,0] g + Tm[1,1] gn
        Tm = np.array([[1.0 / n1, 0.0], [-c12 / (n1 * n2), 1.0 / n2]]) if n2 > 0 else np.array([[1.0 / n1, 0.0], [0.0, 0.0]])
        Bg = np.array([[JgJg, Jg_Jgn], [Jg_Jgn, Jgn_Jgn]])
        B_S = Tm @ Bg @ Tm.T
        g_S = Tm @ np.array([gg, g_gn])
        actual = -1.0
        while actual <= 0:
            p_S = solve_tr_2d(B_S, g_S, Delta)
            coef = Tm.T @ p_S                              # step_h = coef[0] g_h + coef[1] gn_h
            pred = -(0.5 * p_S @ B_S @ p_S + g_S @ p_S)
            shc, shp = coef[0] * ghc + coef[1] * gnc, coef[0] * ghp + coef[1] * gnp
            cam_new = cam + (dc * shc).reshape(C, 6)
            X_new = X + dp * shp
            r_new = residuals(cam_new, intr, X_new, cam_idx, pt_idx, obs)
            nfev += 1
            F_new = 0.5 * np.sum(r_new * r_new)
            sh_norm = np.sqrt(p_S @ p_S)
            actual = F - F_new
            ratio = actual / pred if pred > 0 else (1.0 if pred == actual == 0 else 0.0)
            Delta_new = Delta
            if ratio < 0.25:
                Delta_new = 0.25 * sh_norm
            elif ratio > 0.75 and sh_norm > 0.95 * Delta:
                Delta_new = 2.0 * Delta
            step_norm = np.sqrt(np.sum((dc * shc) ** 2) + np.sum((dp * shp) ** 2))
            x_norm = np.sqrt(np.sum(cam ** 2) + np.sum(X ** 2))
            f_ok = actual < ftol * F and ratio > 0.25
            x_ok = step_norm < xtol * (xtol + x_norm)
            status = 4 if (f_ok and x_ok) else 2 if f_ok else 3 if x_ok else 0
            if verbose:
                print(f"  trf it {it}: F {F:.6f} -> {F_new:.6f} ratio {ratio:.4f} Delta {Delta:.4e} reg {reg:.3e} |step| {step_norm:.3e}")
            if status:
                break
            Delta = Delta_new
        if actual > 0:
            cam, X, F = cam_new, X_new, F_new
            r, Jc, Jp = jacobians(cam, intr, X, cam_idx, pt_idx, obs)
        if status:
            break
    return cam, X, dict(nfev=nfev, cost=F, status=status)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--lam", type=float, default=1e-6)
    ap.add_argument("--no-scipy", action="store_true")
    a = ap.parse_args()
    calib, pts_xy, Xgt = synth.config3_geometry(a.frames, seed=a.seed)
    C = 7
    P0 = g.projection_matrices(calib["R"], calib["tvec"], calib["intr"])
    t0 = time.time()
    X0 = g.triangulate_dlt(P0, pts_xy)
    print(f"DLT {time.time() - t0:.2f}s")
    cam_idx, pt_idx, obs = g.ba_observations(pts_xy)
    cam0 = np.stack([np.concatenate([g.rodrigues_inv(calib["R"][c]), calib["tvec"][c]]) for c in range(C)])
    t0 = time.time()
    cam_lm, X_lm, info = trf_schur(cam0, calib["intr"], X0.reshape(-1, 3), cam_idx, pt_idx, obs)
    print(f"TRF/Schur {time.time() - t0:.2f}s {info}")
    R_lm = rodrigues_batch(cam_lm[:, :3])
    Xt_lm = g.triangulate_dlt(g.projection_matrices(R_lm, cam_lm[:, 3:], calib["intr"]), pts_xy)
    if a.no_scipy:
        return
    t0 = time.time()
    Ro, to, sol = g.bundle_adjust(calib["R"], calib["tvec"], calib["intr"], pts_xy, return_info=True)
    print(f"SciPy TRF {time.time() - t0:.2f}s nfev {sol.nfev} cost {sol.cost:.6f} status {sol.status}")
    Xo = g.triangulate_dlt(g.projection_matrices(Ro, to, calib["intr"]), pts_xy)
    print(f"T={a.frames}: max|X_lm - X_scipy| = {np.abs(Xt_lm - Xo).max():.3e}   "
          f"max|R| {np.abs(R_lm - Ro).max():.3e} max|t| {np.abs(cam_lm[:, 3:] - to).max():.3e}")


if __name__ == "__main__":
    main()
