"""Host driver logic that needs no GPU: the Anderson fixed-point iteration (nlsolve stand-in) and argument validation of the mirror."""
import numpy as np
import pytest


def test_anderson_solves_linear_and_nonlinear_fixed_points():
    from fddgasolver_jl_b200.nlsolve import anderson
    rng = np.random.default_rng(0)
    n = 30
    M = 0.5 * rng.standard_normal((n, n)) / np.sqrt(n) + 0.2j * rng.standard_normal((n, n)) / np.sqrt(n)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    res = anderson(lambda x: M @ x + b - x, np.zeros(n, dtype=complex), m=8, beta=0.85, ftol=1e-12, iterations=200)
    assert res.f_converged and np.linalg.norm(res.zero - np.linalg.solve(np.eye(n) - M, b)) < 1e-10
    # plain damped iteration would need ~10x more steps than the accelerated one
    assert res.iterations < 60
    res2 = anderson(lambda x: np.cos(x) - x, np.array([0.3 + 0j]), m=5, beta=1.0, ftol=1e-13, iterations=50)
    assert res2.f_converged and abs(res2.zero[0] - 0.7390851332151607) < 1e-12
    res3 = anderson(lambda x: 2.0 * x + 1.0 - x, np.array([1.0 + 0j]), m=0, beta=1.0, ftol=1e-12, iterations=5)
    assert not res3.f_converged and res3.iterations == 5


def test_mirror_rejects_unknown_strategies_without_a_device():
    import fddgasolver_jl_b200 as fd

    class Dummy:
        def length_F(self):
            return 4
    with pytest.raises(ValueError):
        fd.mfRGLinearMap(Dummy(), "scPA")
    with pytest.raises(TypeError):
        fd.dqgmres(object(), np.zeros(4))
    with pytest.raises(AssertionError):
        fd.iterate_solver(Dummy(), "no_such_strategy")
    with pytest.raises(ValueError):
        fd.SDE(Dummy(), "no_such_strategy")
