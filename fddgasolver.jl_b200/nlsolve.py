"""Anderson-accelerated fixed-point iteration standing in for `nlsolve(...; method = :anderson, m, beta, ftol, iterations)`
(NLsolve.jl is a dependency of the reference, not in its tree; call sites src/solve.jl:160-196, src/mfRG.jl:287-294).
Host-side driver code: it only combines residual vectors; all heavy work happens inside `fixed_point(x)`."""
import numpy as np


class Result:
    def __init__(self, zero, f_converged, iterations, residual_norm):
        self.zero, self.f_converged, self.iterations, self.residual_norm = zero, f_converged, iterations, residual_norm


def anderson(fixed_point, x0, *, m=50, beta=0.85, ftol=1e-4, iterations=40, show_trace=False, droptol=1e10):
    """Solve R(x) = 0 where fixed_point(x) returns the residual R(x) = g(x) - x.  Converged when |R|_inf <= ftol.
    droptol (NLsolve default 1e10): while the least-squares matrix of residual differences has a condition number above it, the
    OLDEST history column is dropped, as NLsolve's anderson does on its QR factor."""
    x = np.array(x0, dtype=np.complex128, copy=True)
    Xs, Rs = [], []
    err = np.inf
    for it in range(1, iterations + 1):
        R = np.asarray(fixed_point(x))
        err = float(np.max(np.abs(R)))
        if show_trace:
            print(f"  anderson {it:3d}  |R|_inf = {err:.6e}", flush=True)
        if err <= ftol:
            return Result(x, True, it, err)
        Xs.append(x.copy())
        Rs.append(R.copy())
        if len(Xs) > m + 1:
            Xs.pop(0)
            Rs.pop(0)
        if len(Xs) == 1:
            x = x + beta * R
        else:
            if droptol is not None:
                while len(Xs) > 2:
                    Rfac = np.linalg.qr(np.stack([Rs[i + 1] - Rs[i] for i in range(len(Rs) - 1)], axis=1), mode="r")
                    if np.linalg.cond(Rfac) <= droptol:
                        break
                    Xs.pop(0)
                    Rs.pop(0)
            dR = np.stack([Rs[i + 1] - Rs[i] for i in range(len(Rs) - 1)], axis=1)
            dX = np.stack([Xs[i + 1] - Xs[i] for i in range(len(Xs) - 1)], axis=1)
            gamma, *_ = np.linalg.lstsq(dR, R, rcond=None)
            x = x + beta * R - (dX + beta * dR) @ gamma
    return Result(x, False, iterations, err)
