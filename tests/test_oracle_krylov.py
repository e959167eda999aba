"""Pins oracle.dqgmres (restated from Saad & Wu 1996; the reference calls Krylov.dqgmres, src/mfRG.jl:147-151, whose source is a
dependency outside the tree) against scipy's GMRES and the direct solution on random complex systems."""
import numpy as np
import pytest


class _Dense:
    def __init__(self, M):
        self.M = M

    def matvec(self, x):
        return self.M @ x


@pytest.mark.parametrize("memory", [100, 12, 3])
def test_dqgmres_solves_complex_system(orc, memory):
    rng = np.random.default_rng(0)
    n = 80
    M = np.eye(n) + 0.3 * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(n)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x, st = orc.dqgmres(_Dense(M), b, memory=memory, atol=1e-11, rtol=1e-11, itmax=300)
    assert st["solved"]
    assert np.linalg.norm(M @ x - b) < 1e-9
    assert np.linalg.norm(x - np.linalg.solve(M, b)) < 1e-9
    assert len(st["residuals"]) == st["niter"] + 1 and st["residuals"][0] == pytest.approx(np.linalg.norm(b))


def test_full_memory_dqgmres_is_gmres(orc):
    """with memory >= number of iterations the orthogonalisation is complete: iterates and residual norms are GMRES's"""
    import scipy.sparse.linalg as sl
    rng = np.random.default_rng(1)
    n = 50
    M = np.eye(n) + 0.4 * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(n)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    res = []
    sl.gmres(M, b, rtol=1e-30, atol=0.0, restart=n, maxiter=1, callback=lambda r: res.append(r), callback_type="pr_norm")
    x, st = orc.dqgmres(_Dense(M), b, memory=n, atol=0.0, rtol=1e-13, itmax=n)
    m = min(len(res), st["niter"], 12)
    mine = np.array(st["residuals"][1:m + 1]) / np.linalg.norm(b)
    assert np.max(np.abs(mine - np.array(res[:m])) / np.array(res[:m])) < 1e-8
    # truncated orthogonalisation: the quasi-residual estimate stays within sqrt(m + 1) of the true residual (Saad & Wu, Prop. 4.1)
    x3, st3 = orc.dqgmres(_Dense(M), b, memory=3, atol=0.0, rtol=1e-6, itmax=200)
    true = np.linalg.norm(M @ x3 - b)
    assert true <= np.sqrt(st3["niter"] + 1) * st3["residuals"][-1] * (1 + 1e-8)


def test_zero_rhs(orc):
    x, st = orc.dqgmres(_Dense(np.eye(4)), np.zeros(4), memory=2)
    assert st["solved"] and st["niter"] == 0 and not np.any(x)
