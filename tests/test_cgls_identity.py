"""CPU suite: the identity behind the indirect projector's residual recurrence.

The device path forms y = A x at the end of a projection as t_y - r, r being the residual CGLS keeps up to
date (reference src/cpu/include/cgls.h:275: r -= alpha q), and chains the projections of an ADMM run through
y_prev (graph_solver.cuh: cgls_epilogue); a product refreshes y every 16th iteration.  Restated here in numpy,
in fp32 and fp64: over a chain of warm-started projections the recurrence stays within rounding of the
product, with and without the refresh, and the iterates of the two variants agree."""
import numpy as np
import pytest


def cgls(A, b, x, r, shift, tol, maxit, dt):
    """cgls::Solve (cgls.h:201-323) on the shifted problem, started from x with the residual r = b - A x given."""
    s = (A.T @ r - dt(shift) * x).astype(dt)
    p = s.copy()
    norms0 = float(np.linalg.norm(s.astype(np.float64)))
    gamma = norms0 * norms0
    if norms0 < np.finfo(dt).eps:
        return x, r, 0
    k = 0
    for k in range(maxit):
        q = (A @ p).astype(dt)
        delta = float(np.dot(q.astype(np.float64), q.astype(np.float64))) + shift * float(np.dot(p.astype(np.float64), p.astype(np.float64)))
        alpha = dt(gamma / delta)
        x = (x + alpha * p).astype(dt)
        r = (r - alpha * q).astype(dt)
        s = (A.T @ r - dt(shift) * x).astype(dt)
        norms = float(np.linalg.norm(s.astype(np.float64)))
        gamma1, gamma = gamma, norms * norms
        p = (s + dt(gamma / gamma1) * p).astype(dt)
        if norms <= norms0 * tol:
            break
    return x, r, k + 1


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("refresh", [16, 10 ** 9])
def test_y_from_the_residual_recurrence_tracks_the_product(dt, refresh):
    rng = np.random.default_rng(5)
    m, n, K = 600, 150, 80
    A = (rng.standard_normal((m, n)) / np.sqrt(m)).astype(dt)
    eps = np.finfo(dt).eps
    x_rec = np.zeros(n, dt); y_rec = np.zeros(m, dt)      # chain that never multiplies for y (but for the refresh)
    x_mul = np.zeros(n, dt); y_mul = np.zeros(m, dt)      # reference chain: y = A x by a product every time
    worst = 0.0
    for it in range(K):
        tx = rng.standard_normal(n).astype(dt); ty = rng.standard_normal(m).astype(dt)
        # (projector_cgls.cpp:60-78 in the variables of the device path: unknown dx = x - t_x, start residual
        #  t_y - A x_prev, which is t_y - y_prev)
        out = []
        for x_prev, y_prev, rec in ((x_rec, y_rec, True), (x_mul, y_mul, False)):
            dx = (x_prev - tx).astype(dt)
            r0 = (ty - y_prev).astype(dt)
            dx, r, _ = cgls(A, None, dx, r0, 1.0, 1e-6, 200, dt)
            x = (tx + dx).astype(dt)
            if rec and (it + 1) % refresh != 0 and it > 0:
                y = (ty - r).astype(dt)
            else:
                y = (A @ x).astype(dt)
            out.append((x, y))
        (x_rec, y_rec), (x_mul, y_mul) = out
        worst = max(worst, float(np.linalg.norm((y_rec - A @ x_rec).astype(np.float64)) / np.linalg.norm(y_rec.astype(np.float64))))
    # the recurrence stays within a few roundings of the product over 80 chained projections (observed: 10 eps
    # with the refresh, 25 eps without) ...
    assert worst < 100 * eps
    # ... and the two chains end at the same point
    assert np.linalg.norm((x_rec - x_mul).astype(np.float64)) / np.linalg.norm(x_mul.astype(np.float64)) < 100 * eps
