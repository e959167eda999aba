"""Shared helpers for the tests: a small Anderson-accelerated fixed-point solver standing in for the
reference's `solve!` (NLsolve :anderson, m = 8, beta = 0.85, src/solve.jl:160-196) and state builders."""
import numpy as np


def anderson(fixed_point, x0, m=8, beta=0.85, tol=1e-9, maxiter=200, verbose=False):
    """Solve R(x) = g(x) - x = 0.  fixed_point(x) returns R(x)."""
    x = x0.copy()
    Xs, Rs = [], []
    for it in range(maxiter):
        R = fixed_point(x)
        err = np.max(np.abs(R))
        if verbose:
            print(f"  anderson it {it:3d} |R|_inf = {err:.3e}")
        if err < tol:
            return x, it, err
        Xs.append(x.copy())
        Rs.append(R.copy())
        if len(Xs) > m + 1:
            Xs.pop(0)
            Rs.pop(0)
        if len(Xs) == 1:
            x = x + beta * R
        else:
            dR = np.stack([Rs[i + 1] - Rs[i] for i in range(len(Rs) - 1)], axis=1)
            dX = np.stack([Xs[i + 1] - Xs[i] for i in range(len(Xs) - 1)], axis=1)
            gamma, *_ = np.linalg.lstsq(dR, R, rcond=None)
            x = x + beta * R - (dX + beta * dR) @ gamma
    return x, maxiter, err


def oracle_fixed_point(o, S, strategy):
    """fixed_point!(R, x, S) on the flattened [F; Σ] for an OracleSolver (src/solve.jl:119-157)."""
    nF = len(S.F)

    def fp(x):
        S.F.unflatten(x[:nF])
        S.Σ[...] = x[nF:].reshape(S.Σ.shape, order="F")
        o.iterate_solver(S, strategy, True)
        return np.concatenate([S.F.flatten(), S.Σ.ravel(order="F")]) - x
    return fp


def flatten_solver(S):
    return np.concatenate([S.F.flatten(), S.Σ.ravel(order="F")])
