"""Pins the CPU oracle against the GOLDEN NUMBERS stored in the reference's tests (local / SIAM solver).

  test/test_siam_scPA.jl:28-31, 60-63   converged scPA: Σ at ν = πT·[-3,-1,1,3] and K1 of all three channels at
                                        2πT·(-2):1:2 (MeshFunction call = linear interpolation), half filling and doped
  test/test_siam_fdPA.jl:52-84          fdPA(reference -> target) vs scPA(target) with the reference's tolerances and the
                                        golden Σ(πT)
The reference stops NLsolve-Anderson at ftol = 1e-4, so its numbers carry ~1e-5 of convergence error: tolerance 1e-4 here.

Two observations about the reference at HEAD (documented in DESIGN.md section 2):
  * the scPA goldens are reproduced with mΠν_factor = 1 (bubble fermionic box = K1 box), not with the current default 6
    (src/ParquetSolver.jl:92): with 6 the K1 values differ by 3-7 %;
  * the fdPA golden / tolerances are reproduced when the reference Hartree term is subtracted ONCE; exactly as coded
    (src/SDE.jl:13-24) Σ is off by ≈ (n0 - 1/2) U.
"""
import numpy as np
import pytest

from helpers import anderson, flatten_solver


def _fp(o, S, strategy):
    nF = len(S.F)

    def fp(x):
        S.F.unflatten(x[:nF])
        S.Σ[...] = x[nF:].reshape(S.Σ.shape, order="F")
        o.iterate_solver_local(S, strategy, True)
        return np.concatenate([S.F.flatten(), S.Σ.ravel(order="F")]) - x
    return fp


def _solve(o, S, strategy):
    x, it, err = anderson(_fp(o, S, strategy), flatten_solver(S), tol=1e-10)
    assert err < 1e-10
    nF = len(S.F)
    S.F.unflatten(x[:nF])
    S.Σ[...] = x[nF:].reshape(S.Σ.shape, order="F")


def _siam(o, nG, nK1, nK2, nK3, *, T, U, e, Δ, D, factor=1):
    from fddgasolver_jl_b200.types import RefVertex
    Gb = o.siam_bare_Green(T, nG, e=e, Δ=Δ, D=D)
    S = o.OracleLocalSolver(nK1, nK2, nK3, Gb, np.zeros_like(Gb), np.zeros_like(Gb), RefVertex(T, U), T=T, mΠν_factor=factor)
    S.init_sym_grp()
    return S


GOLD_HALF = dict(
    Σ=[-0.052138235296134906, -0.03838544776344314, 0.03838544776344314, 0.052138235296134906],
    γa=[0.13203850929270397, 0.5403615530152339, 0.2333246221064017, 0.09056300899983459],
    γp=[-0.10420799999591804, -0.2403951910434166, -0.15592452265704748, -0.07622568434721624],
    γt=[0.013898648018808482, 0.1499562726081748, 0.03867632419082161, 0.007160400841240708])
GOLD_DOPED = dict(
    Σ=[-0.0389123277075552 - 0.16855090184215607j, -0.025252640312580586 - 0.17429637478745583j,
       0.025252640312580586 - 0.17429637478745583j, 0.0389123277075552 - 0.16855090184215607j],
    γa=[0.11925962005661812 + 8.57514999021054e-5j, 0.416232811242488 + 3.319936929625957e-5j,
        0.20353141073439696 - 8.209974062547027e-5j, 0.08259294412067451 - 7.660306021755952e-5j],
    γp=[-0.12570450372739117 + 0.06583917638195431j, -0.24548654160724023 + 0.021014183409764874j,
        -0.17578586892780296 - 0.05408344507941078j, -0.09544981624337806 - 0.06686768343644132j],
    γt=[0.011016969129875598 + 7.455601769977568e-5j, 0.09533547032272821 + 2.4477973720973318e-5j,
        0.028843799251846686 - 6.260706942592786e-5j, 0.005828299701446311 - 7.32415771550901e-5j])


@pytest.mark.parametrize("e,nK2,gold", [(0.0, (6, 6), GOLD_HALF), (0.5, (7, 6), GOLD_DOPED)])
def test_siam_scPA_golden_numbers(orc, e, nK2, gold):
    T, nmax = 0.1, 6
    nG, nK1 = 6 * nmax, 4 * nmax
    S = _siam(orc, nG, nK1, nK2, nK2, T=T, U=1.0, e=e, Δ=np.pi / 5, D=10.0)
    _solve(orc, S, "scPA")
    got = [S.Σ[n + nG, 0] for n in (-2, -1, 0, 1)]
    assert np.max(np.abs(np.array(got) - np.array(gold["Σ"]))) < 1e-4
    xs = [-4 * np.pi * T + i for i in range(4)]          # `2π*T .* -2:2` parses as the range (-4πT):1:2
    for name, g in (("γa", S.F.γa), ("γp", S.F.γp), ("γt", S.F.γt)):
        vals = [orc.interp_boson(g.K1[:, 0], T, nK1, x) for x in xs]
        assert np.max(np.abs(np.array(vals) - np.array(gold[name]))) < 1e-4, name


def test_siam_fdPA_reference_test_and_golden_sigma(orc):
    from fddgasolver_jl_b200.types import RefVertex
    T, U, nmax = 0.1, 1.0, 12
    nG = nK1 = 8 * nmax
    S0 = _siam(orc, nG, nK1, (nmax, nmax), (nmax, nmax), T=T, U=U, e=-0.3, Δ=np.pi / 3, D=10.0)
    _solve(orc, S0, "scPA")
    orc.Dyson(S0)
    # trivial case: zero reference -> fdPA == scPA (test_siam_fdPA.jl:27-33)
    S0fd = _siam(orc, nG, nK1, (nmax, nmax), (nmax, nmax), T=T, U=U, e=-0.3, Δ=np.pi / 3, D=10.0)
    _solve(orc, S0fd, "fdPA")
    assert np.max(np.abs(S0fd.Σ - S0.Σ)) < 1e-10 and np.max(np.abs(S0fd.F.flatten() - S0.F.flatten())) < 1e-10
    S = _siam(orc, nG, nK1, (nmax, nmax), (nmax, nmax), T=T, U=U, e=0.5, Δ=np.pi / 5, D=20.0)
    _solve(orc, S, "scPA")
    Gb2 = orc.siam_bare_Green(T, nG, e=0.5, Δ=np.pi / 5, D=20.0)
    orc.QUIRK_E1 = False
    try:
        Sfd = orc.OracleLocalSolver(nK1, (nmax, nmax), (nmax, nmax), Gb2, S0.G, S0.Σ, S0.F, T=T, mΠν_factor=1)
        Sfd.init_sym_grp()
        _solve(orc, Sfd, "fdPA")
        Sfd2 = orc.OracleLocalSolver(64, (8, 8), (8, 8), Gb2, S0.G, S0.Σ, S0.F, T=T, mΠν_factor=1)
        Sfd2.init_sym_grp()
        _solve(orc, Sfd2, "fdPA")
    finally:
        orc.QUIRK_E1 = True
    assert np.max(np.abs(Sfd.Σ - S.Σ)) < 3e-5                                  # :52
    for ch in range(3):
        for cls, tol in (("K1", 5e-4), ("K2", 1e-3), ("K3", 1e-3)):          # :54-62
            d = getattr(Sfd.F.channel(ch), cls) + getattr(Sfd.F0.channel(ch), cls) - getattr(S.F.channel(ch), cls)
            assert np.max(np.abs(d)) < tol, (ch, cls)
    assert np.max(np.abs(Sfd2.Σ - S.Σ)) < 3e-3                                 # :83
    assert abs(Sfd2.Σ[nG, 0] - (0.024643001835742997 - 0.17494219707558506j)) < 5e-5   # :84 golden Σ(πT)
