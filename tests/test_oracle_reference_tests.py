"""Pins the CPU oracle against the reference's OWN tests (restated; Julia cannot run here).

Each test cites the reference test it restates.  These run on CPU in the `-m "not gpu"` suite.
"""
import numpy as np
import pytest

from helpers import anderson, flatten_solver, oracle_fixed_point

INF = (2 ** 31 - 1) // 4


# ------------------------------------------------------------------ test/test_hubbard.jl
def test_golden_occupations(orc):
    """test/test_hubbard.jl:84-88: compute_occupation(hubbard_bare_Green(...)) golden numbers"""
    class S:
        nG, LG, T = 20, 8, 0.5
    for mu, ref in [(-4.0, 0.0502663698543071), (-2.0, 0.2057188296739284), (0.0, 0.5),
                    (4.0, 1 - 0.0502663698543071), (2.0, 1 - 0.2057188296739284)]:
        G = orc.hubbard_bare_Green(0.5, 20, 8, μ=mu, t1=1.0)
        assert abs(orc.compute_occupation(S, G) - ref) < 1e-13


def test_bare_green_closed_forms_and_dyson(orc):
    """test/test_hubbard.jl:26-33, 51-62: -i*Gbare = 1/(iν + μ - ε_k) at special k; Dyson"""
    import fddgasolver_jl_b200 as fd
    T, t1, mu, nG, L = 0.5, 1.3, 0.2, 5, 8
    G = orc.hubbard_bare_Green(T, nG, L, μ=mu, t1=t1)
    assert np.array_equal(G, fd.hubbard_bare_Green(T, nG, L, μ=mu, t1=t1)) or np.max(np.abs(G - fd.hubbard_bare_Green(T, nG, L, μ=mu, t1=t1))) < 1e-15
    for n in range(-2, 3):
        nu = (2 * n + 1) * np.pi * T
        for (ix, iy), eps in {(0, 0): -4 * t1, (0, 2): -2 * t1, (0, 4): 0.0, (0, 6): -2 * t1, (4, 4): 4 * t1}.items():
            assert abs(-1j * G[n + nG, ix + L * iy] - 1 / (1j * nu + mu - eps)) < 1e-14
    Σ = np.full_like(G, -0.5 + 0.2j)

    class S:
        pass
    S.G, S.Σ, S.Gbare = np.zeros_like(G), Σ, G
    orc.Dyson(S)
    assert np.max(np.abs(S.G - 1 / (1 / G + Σ))) < 1e-15


def test_bubbles_momentum_space_identity_and_real_space_agreement(orc):
    """test/test_hubbard.jl:66-78: Πpp = G(ν,k) G(Ω-ν,P-k), Πph = G(Ω+ν,P+k) G(ν,k);
    test/test_nonlocal_vertex.jl:33-38 (spirit): real-space == momentum-space bubbles when L = LG and all frequencies are in the box"""
    from fddgasolver_jl_b200.types import RefVertex
    T, nG, L = 0.5, 12, 4
    G = orc.hubbard_bare_Green(T, nG, L, μ=0.2, t1=1.3)
    S = orc.OracleSolver(4, (2, 2), (2, 2), L, G, G, np.zeros_like(G), RefVertex(T, 1.0), T=T, compute_bubbles=False)
    pp, ph = np.zeros_like(S.Πpp), np.zeros_like(S.Πph)
    orc.bubbles_momentum_space(S, pp, ph, G)
    W, v, iP, ik = 3, -2, 2 + L * 1, 3 + L * 2          # boson 3, fermion -2
    Px, Py, kx, ky = 2, 1, 3, 2
    a, b = W - v - 1, W + v
    Gk = G[v + nG, ik]
    assert abs(pp[W + 3, v + 4, iP, ik] - Gk * G[a + nG, (Px - kx) % L + L * ((Py - ky) % L)]) < 1e-15
    assert abs(ph[W + 3, v + 4, iP, ik] - Gk * G[b + nG, (Px + kx) % L + L * ((Py + ky) % L)]) < 1e-15
    # L odd avoids the half-weight construction: real space == momentum space exactly (all frequencies in the G box)
    L3 = 3
    G3 = orc.hubbard_bare_Green(T, nG, L3, μ=0.2, t1=1.3)
    S3 = orc.OracleSolver(4, (2, 2), (2, 2), L3, G3, G3, np.zeros_like(G3), RefVertex(T, 1.0), T=T, compute_bubbles=False)
    a1, a2, b1, b2 = (np.zeros_like(S3.Πpp) for _ in range(4))
    orc.bubbles_momentum_space(S3, a1, a2, G3)
    orc.bubbles_real_space(S3, b1, b2, G3)
    assert np.max(np.abs(a1 - b1)) < 1e-13 and np.max(np.abs(a2 - b2)) < 1e-13


# ------------------------------------------------------------------ test/test_channel.jl, test/test_nonlocal_2_vertex.jl
def _rand_nl2(T, nK1, nK2, nK3, L, U=3.0, seed=0):
    from fddgasolver_jl_b200.types import NL2_Vertex, RefVertex
    rng = np.random.default_rng(seed)
    F = NL2_Vertex(RefVertex(T, U), T, nK1, nK2, nK3, L)
    for g in F.channels():
        for a in g.arrays():
            a[...] = rng.random(a.shape) + 1j * rng.random(a.shape)
    return F


def test_mesh_lengths():
    """test/test_channel.jl:16: boson mesh 2N-1 points, fermion mesh 2N points"""
    from fddgasolver_jl_b200.types import NL2_Channel
    g = NL2_Channel(0.5, 5, (4, 3), (2, 3), 3)
    assert g.K1.shape == (9, 9) and g.K2.shape == (7, 6, 9, 9) and g.K3.shape == (3, 6, 6, 9)


def test_nl2_channel_evaluator_with_fold_back(orc):
    """test/test_nonlocal_2_vertex.jl:20-36"""
    T, L = 0.5, 3
    F = _rand_nl2(T, 5, (4, 3), (2, 3), L)
    g = F.γp
    W, v, w = 1, 2, -1
    P_, k_, q_ = (-1, 1), (0, 5), (4, -2)
    iP, ik, iq = (P_[0] % L) + L * (P_[1] % L), (k_[0] % L) + L * (k_[1] % L), (q_[0] % L) + L * (q_[1] % L)
    K1, K2a, K2b, K3 = g.K1[W + 4, iP], g.K2[W + 3, v + 3, iP, ik], g.K2[W + 3, w + 3, iP, iq], g.K3[W + 1, v + 3, w + 3, iP]
    ev = lambda vv, ww: orc.eval_channel(F, L, 0, W, vv, ww, P_, k_, q_)
    assert abs(ev(v, w) - (K1 + K2a + K2b + K3)) < 1e-14
    assert abs(ev(INF, w) - (K1 + K2b)) < 1e-14
    assert abs(ev(v, INF) - (K1 + K2a)) < 1e-14
    assert abs(ev(INF, INF) - K1) < 1e-14


def test_nl2_vertex_channel_maps(orc):
    """test/test_nonlocal_2_vertex.jl:84-97: only K1 non-zero -> F in channel Ch = U + sum of K1's at converted arguments"""
    from fddgasolver_jl_b200.types import pCh, tCh, aCh, pSp
    T, U, L = 0.5, 3.0, 3
    F = _rand_nl2(T, 10, (4, 3), (2, 1), L, U=U)
    for g in F.channels():
        g.K2[...] = 0
        g.K3[...] = 0
    W, v, w = 1, 2, -1
    P, k, q = (-1, 1), (1, 1), (0, 1)

    def K1(g, m, mom):           # call semantics: 0 outside the mesh, momentum folded
        if abs(m) > 9:
            return 0.0
        return g.K1[m + 9, (mom[0] % L) + L * (mom[1] % L)]
    add = lambda *xs: tuple(sum(c) for c in zip(*xs))
    neg = lambda x: (-x[0], -x[1])
    ev = lambda vv, ww, ch: orc.eval_vertex(F, L, W, vv, ww, P, k, q, ch, pSp)
    for ch, g in ((pCh, F.γp), (tCh, F.γt), (aCh, F.γa)):
        assert abs(ev(INF, INF, ch) - (U + K1(g, W, P))) < 1e-14
    # frequencies: B(m) - F(n) - F(n') = B(m-n-n'-1); F(n) - F(n') = B(n-n'); B(m) + F(n) + F(n') = B(m+n+n'+1)
    exp_p = U + K1(F.γp, W, P) + K1(F.γt, W - v - w - 1, add(P, neg(k), neg(q))) + K1(F.γa, v - w, add(k, neg(q)))
    exp_t = U + K1(F.γt, W, P) + K1(F.γp, W + v + w + 1, add(P, k, q)) + K1(F.γa, w - v, add(q, neg(k)))
    exp_a = U + K1(F.γa, W, P) + K1(F.γp, W + v + w + 1, add(P, k, q)) + K1(F.γt, v - w, add(k, neg(q)))
    assert abs(ev(v, w, pCh) - exp_p) < 1e-14
    assert abs(ev(v, w, tCh) - exp_t) < 1e-14
    assert abs(ev(v, w, aCh) - exp_a) < 1e-14


@pytest.mark.parametrize("v,w", [(1, -2), (INF, -2), (1, INF)])
def test_swave_evaluation_equals_explicit_average(orc, v, w):
    """test/test_nonlocal_2_vertex.jl:114-222: F(..., kSW, q) / (k, kSW) / (kSW, kSW) == explicit BZ averages,
    for every channel, spin and gamma-switch combination"""
    from fddgasolver_jl_b200.types import pCh, tCh, aCh, pSp, xSp, dSp
    T, L = 0.5, 3
    F = _rand_nl2(T, 6, (3, 3), (2, 2), L, seed=4)
    W, P, k0, q0 = 0, (1, 2), (2, 0), (1, 1)
    pts = [(i, j) for j in range(L) for i in range(L)]
    for ch in (pCh, tCh, aCh):
        for sp in (pSp, xSp, dSp):
            for fl in ((True, True, True, True), (False, True, False, True), (True, False, True, False)):
                kw = dict(F0=fl[0], γp=fl[1], γt=fl[2], γa=fl[3])
                full = lambda k, q: orc.eval_vertex(F, L, W, v, w, P, k, q, ch, sp, **kw)
                avg_k = np.mean([full(k, q0) for k in pts])
                avg_q = np.mean([full(k0, q) for q in pts])
                avg_kq = np.mean([full(k, q) for k in pts for q in pts])
                assert abs(orc.eval_vertex(F, L, W, v, w, P, "sw", q0, ch, sp, **kw) - avg_k) < 1e-13
                assert abs(orc.eval_vertex(F, L, W, v, w, P, k0, "sw", ch, sp, **kw) - avg_q) < 1e-13
                assert abs(orc.eval_vertex(F, L, W, v, w, P, "sw", "sw", ch, sp, **kw) - avg_kq) < 1e-13


# ------------------------------------------------------------------ test/test_nonlocal_2_fdPA.jl
def _pa_solver(orc, mu, t2=0.0):
    from fddgasolver_jl_b200.types import RefVertex
    T, U, t1, nmax, nq, LG = 0.5, 2.0, 1.0, 3, 3, 24
    Gb = orc.hubbard_bare_Green(T, 6 * nmax, LG, μ=mu, t1=t1, t2=t2)
    S = orc.OracleSolver(6 * nmax, (nmax, nmax), (nmax, nmax), nq, Gb, np.zeros_like(Gb), np.zeros_like(Gb), RefVertex(T, U), T=T)
    S.init_sym_grp()
    return S


def _solve(orc, S, strategy):
    x, it, err = anderson(oracle_fixed_point(orc, S, strategy), flatten_solver(S), tol=1e-10)
    assert err < 1e-10, (it, err)
    nF = len(S.F)
    S.F.unflatten(x[:nF])
    S.Σ[...] = x[nF:].reshape(S.Σ.shape, order="F")
    return x


@pytest.fixture(scope="module")
def converged_reference(orc):
    S0 = _pa_solver(orc, 0.0)
    x0 = _solve(orc, S0, "scPA")
    return S0, x0


def test_fdPA_equals_scPA_for_zero_reference(orc, converged_reference):
    """test/test_nonlocal_2_fdPA.jl:27-40: G0 = Σ0 = Π0 = 0 -> fdPA and scPA give identical Σ and F (1e-10)"""
    S0, x0 = converged_reference
    S0_fd = _pa_solver(orc, 0.0)
    x1 = _solve(orc, S0_fd, "fdPA")
    assert np.max(np.abs(x1 - x0)) < 1e-10


def test_converged_fdPA_matches_scPA_of_target(orc, converged_reference):
    """test/test_nonlocal_2_fdPA.jl:43-72 with the reference's tolerances.  Uses the SDE L kernels "as commented"
    (own-channel γ only): restated *as coded* this reference test fails."""
    from fddgasolver_jl_b200.types import pCh, tCh, aCh
    S0, _ = converged_reference
    orc.Dyson(S0)
    S = _pa_solver(orc, 0.5, -0.3)
    _solve(orc, S, "scPA")
    Gb = orc.hubbard_bare_Green(0.5, 18, 24, μ=0.5, t1=1.0, t2=-0.3)
    orc.lib().orc_set_quirk_E2(0)
    try:
        Sfd = orc.OracleSolver(18, (3, 3), (3, 3), 3, Gb, S0.G, S0.Σ, S0.F, T=0.5)
        Sfd.init_sym_grp()
        _solve(orc, Sfd, "fdPA")
    finally:
        orc.lib().orc_set_quirk_E2(1)
    assert np.max(np.abs(Sfd.Σ - S.Σ)) < 3e-4
    for ch in (pCh, tCh, aCh):
        for cls, tol in (("K1", 2e-3), ("K2", 4e-3), ("K3", 2e-3)):
            d = getattr(Sfd.F.channel(ch), cls) + getattr(Sfd.F0.channel(ch), cls) - getattr(S.F.channel(ch), cls)
            assert np.max(np.abs(d)) < tol, (ch, cls, np.max(np.abs(d)))
    # as coded (E2) the same comparison is off by two orders of magnitude in Σ -- documented, not asserted tightly
    Sfd2 = orc.OracleSolver(18, (3, 3), (3, 3), 3, Gb, S0.G, S0.Σ, S0.F, T=0.5)
    Sfd2.init_sym_grp()
    _solve(orc, Sfd2, "fdPA")
    assert np.max(np.abs(Sfd2.Σ - S.Σ)) > 1e-2


def test_chemical_potential_inverts_golden_occupations(orc):
    """compute_hubbard_chemical_potential (src/dyson.jl:45-57) must invert the golden occupations of test/test_hubbard.jl:84-88"""
    class S:
        nG, LG, T = 20, 8, 0.5
    S.Σ = np.zeros((40, 64), dtype=np.complex128, order="F")
    for mu, occ in [(-2.0, 0.2057188296739284), (0.0, 0.5), (2.0, 1 - 0.2057188296739284)]:
        got = orc.compute_hubbard_chemical_potential(occ, S, {"t1": 1.0})
        assert abs(got - mu) < 1e-11, (mu, got)
    with pytest.raises(ValueError):
        orc.compute_hubbard_chemical_potential(1.5, S, {"t1": 1.0})
