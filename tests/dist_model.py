"""Test-only numpy model of the row-block ADMM iteration, with the
cross-rank sums done by torch.distributed (gloo on CPU).  It mirrors, step by step, what the
device kernels do in multi-GPU mode -- local prox / products, one summed n-vector, summed
y-side scalars, replicated controller -- so that the decomposition itself (what is local,
what is exchanged, in which order) can be checked against the single-process oracle without a
GPU.  Float64 only; uses the oracle for the scalar prox / objective."""
import numpy as np
import torch
import torch.distributed as dist


def allsum(v):
    t = torch.from_numpy(np.atleast_1d(np.asarray(v, dtype=np.float64)).copy())
    dist.all_reduce(t)
    return t.numpy()


def solve_rowblock(O, A_loc, m_global, f_loc, g, rho=1.0, abs_tol=1e-4, rel_tol=1e-4, max_iter=2500,
                   adaptive_rho=True, gap_stop=True):
    """f_loc, g: (h, a, b, c, d, e) tuples with array entries (local rows / all columns)."""
    mloc, n = A_loc.shape
    m = m_global
    A = A_loc.copy()
    # ---- equilibration (Sinkhorn on A.^2) with the global constants -------------------------
    B = A * A
    d = np.ones(mloc); e = np.ones(n)
    ce = 1e-4 * (m + n) / m; cd = 1e-4 * (m + n) / n
    for _ in range(50):
        e = m / (allsum(B.T @ d) + ce)          # exchange: n-vector
        d = n / (B @ e + cd)                    # local
    d = np.sqrt(d); e = np.sqrt(e)
    fro = allsum(np.sum(d * d * (B @ (e * e))))[0]     # exchange: scalar
    normA = np.sqrt(fro) / np.sqrt(min(m, n))
    A = (d[:, None] * A * e[None, :]) / normA
    d = d / np.sqrt(normA); e = e / np.sqrt(normA)
    # ---- norm estimate --------------------------------------------------------------------------
    x = O.rand(n, np.float64)
    est = 0.0
    for _ in range(50):
        last = est
        Sx = A @ x
        x = allsum(A.T @ Sx)                    # exchange: n-vector
        nSx = np.sqrt(allsum(Sx @ Sx)[0])       # exchange: scalar
        nx = np.linalg.norm(x)
        x = x / nx
        est = nx / nSx
        if abs(last - est) < 1e-4 * est:
            break
    nrmA = est
    # ---- factor (replicated) ---------------------------------------------------------------------
    G = allsum((A.T @ A).ravel()).reshape(n, n)     # exchange: n x n, one-time
    M = np.linalg.inv(np.eye(n) + G)
    # ---- descriptors ---------------------------------------------------------------------------------
    fh, fa, fb, fc, fd, fe = [np.broadcast_to(np.asarray(v, dtype=np.float64 if i else np.int32), (mloc,)).copy() for i, v in enumerate(f_loc)]
    gh, ga, gb, gc, gd, ge = [np.broadcast_to(np.asarray(v, dtype=np.float64 if i else np.int32), (n,)).copy() for i, v in enumerate(g)]
    fc = np.maximum(fc, 0); fe = np.maximum(fe, 0); gc = np.maximum(gc, 0); ge = np.maximum(ge, 0)
    F = (fh, fa / d, fb, fc, fd / d, fe / (d * d))
    Gd = (gh, ga * e, gb, gc, gd * e, ge * (e * e))
    x = np.zeros(n); y = np.zeros(mloc); xt = np.zeros(n); yt = np.zeros(mloc)
    sq = np.sqrt
    delta, xi, k, kd, ku = 1.05, 1.0, 0, 0, 0
    alpha = 1.7
    while True:
        xp, yp = x, y
        x12 = O.prox_vec(Gd, rho, x - xt); y12 = O.prox_vec(F, rho, y - yt)
        wx = (x - xt) - x12; wy = (y - yt) - y12
        ysum = allsum([wy @ y12, wy @ wy, y12 @ y12])     # exchange: 3 scalars (packed with the next 2 on device)
        gap = abs(wx @ x12 + ysum[0])
        eps_gap = sq(m + n) * abs_tol + rel_tol * sq(wx @ wx + ysum[1]) * sq(x12 @ x12 + ysum[2])
        eps_pri = sq(m) * abs_tol + rel_tol * sq(ysum[2])
        eps_dua = rho * (sq(n) * abs_tol + rel_tol * sq(wx @ wx))
        tx = xt + alpha * x12 + (1 - alpha) * xp; ty = yt + alpha * y12 + (1 - alpha) * yp
        u = tx + allsum(A.T @ ty)                           # exchange: n-vector
        x = M @ u
        y = A @ x
        ys2 = allsum([(yp - y) @ (yp - y), (y12 - y) @ (y12 - y)])   # exchange: 2 scalars
        nrm_s = rho * (nrmA * sq(ys2[0]) + np.linalg.norm(xp - x))
        nrm_r = nrmA * np.linalg.norm(x12 - x) + sq(ys2[1])
        exact = nrm_r < 10 * eps_pri and nrm_s < 10 * eps_dua
        if exact:
            r = A @ x12 - y12
            nrm_r = sq(allsum(r @ r)[0])                    # exchange: scalar
            s = (x12 + xt - xp) + allsum(A.T @ (y12 + yt - yp))   # exchange: n-vector
            nrm_s = rho * np.linalg.norm(s)
        conv = exact and nrm_r < eps_pri and nrm_s < eps_dua and (not gap_stop or gap < eps_gap)
        if conv or k == max_iter - 1:
            break
        xt = tx - x; yt = ty - y
        if adaptive_rho:
            if k > 0 and k % 50 == 0 and eps_pri > 0 and eps_dua > 0:
                pn, dn = nrm_r / eps_pri, nrm_s / eps_dua
                if pn > 0 and dn > 0:
                    imb = pn / dn
                    if imb > 10 or imb < 0.1:
                        rho_new = min(1e4, max(1e-4, rho * min(1.5, max(0.67, sq(imb)))))
                        if abs(rho_new - rho) / rho > 0.05:
                            xt = xt * (rho / rho_new); yt = yt * (rho / rho_new); rho = rho_new
            elif nrm_s < xi * eps_dua and nrm_r > xi * eps_pri and 0.8 * k > kd:
                if rho < 1e4:
                    rho *= delta; xt = xt / delta; yt = yt / delta; delta *= 1.01; ku = k
            elif nrm_s > xi * eps_dua and nrm_r < xi * eps_pri and 0.8 * k > ku:
                if rho > 1e-4:
                    rho /= delta; xt = xt * delta; yt = yt * delta; delta *= 1.01; kd = k
            elif nrm_s < xi * eps_dua and nrm_r < xi * eps_pri:
                xi *= 0.9
            else:
                delta = 1.05
        k += 1
    optval = O.func_vec(Gd, x12) + allsum(O.func_vec(F, y12))[0]
    lam = -rho * (y12 + yt - yp) * d
    return dict(x=x12 * e, y=y12 / d, l=lam, optval=optval, iterations=k, status=0 if conv else 3)
