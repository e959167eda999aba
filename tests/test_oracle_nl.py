"""The s-wave solver (NL_ParquetSolver, nl_method = 1 of script/run_Wu_point.jl) in the CPU oracle, pinned against the
reference's own tests for it: test/test_nonlocal_vertex.jl, test/test_nonlocal_fdPA.jl, test/test_nonlocal_symmetry.jl.
Vertices are held in the oracle's own containers (oracle/otypes.py)."""
import numpy as np
import pytest

from helpers import anderson, flatten_solver, oracle_fixed_point

INF = (2 ** 31 - 1) // 4


def _rand_nl(T, nK1, nK2, nK3, L, U=3.0, seed=0):
    from otypes import NL_Vertex, RefVertex
    rng = np.random.default_rng(seed)
    F = NL_Vertex(RefVertex(T, U), T, nK1, nK2, nK3, L)
    for g in F.channels():
        for a in g.arrays():
            a[...] = rng.random(a.shape) + 1j * rng.random(a.shape)
    return F


def test_nl_bubbles_equal_momentum_average_of_nl2_bubbles(orc):
    """test/test_nonlocal_vertex.jl:40-50: Π[Ω, ν, P] of bubbles_real_space!(::NL_MF_Π) == (1 / N_k) Σ_k Π[Ω, ν, P, k] of the NL2
    bubbles.  As coded the two defaults differ (use_G_tail = true in src/nonlocal/bubble.jl:93, false in
    src/nonlocal_2/bubble.jl:48), so the identity holds with the tail switched off; with the tail on only the entries whose
    Green functions leave the mesh change, by exactly the 1/ν·1/ν' (or G·1/ν) term at R = 0."""
    from otypes import RefVertex
    T, nG, LG, L = 0.5, 5, 6, 3
    G = orc.hubbard_bare_Green(T, nG, LG, μ=0.2, t1=1.0)
    S2 = orc.OracleSolver(5, (2, 2), (2, 2), L, G, G, np.zeros_like(G), RefVertex(T, 1.0), T=T, mΠν_factor=1, compute_bubbles=False)
    S1 = orc.OracleNLSolver(5, (2, 2), (2, 2), L, G, G, np.zeros_like(G), RefVertex(T, 1.0), T=T, mΠν_factor=1, compute_bubbles=False)
    assert S1.Πpp.shape == (9, 10, 9) and S2.Πpp.shape == (9, 10, 9, 9)
    orc.bubbles_real_space(S2, S2.Πpp, S2.Πph, G)
    orc.bubbles_real_space(S1, S1.Πpp, S1.Πph, G, use_G_tail=False)
    assert np.max(np.abs(S1.Πpp - S2.Πpp.mean(axis=3))) < 1e-14
    assert np.max(np.abs(S1.Πph - S2.Πph.mean(axis=3))) < 1e-14
    # the tail: Πpp[Ω, ν, P] += G_loc-or-tail(Ω - ν) * G_loc-or-tail(ν) at R = 0 wherever one of the two leaves the G mesh
    a, b = np.zeros_like(S1.Πpp), np.zeros_like(S1.Πph)
    orc.bubbles_real_space(S1, a, b, G)
    Gloc = G.mean(axis=1)
    gt = lambda n: Gloc[n + nG] if -nG <= n < nG else 1.0 / ((2 * n + 1) * np.pi * T)
    g0 = lambda n: Gloc[n + nG] if -nG <= n < nG else 0.0
    for iW, W in enumerate(range(-4, 5)):
        for iw, w in enumerate(range(-5, 5)):
            dpp = gt(W - w - 1) * gt(w) - g0(W - w - 1) * g0(w)
            dph = gt(W + w) * gt(w) - g0(W + w) * g0(w)
            assert np.max(np.abs(a[iW, iw, :] - S1.Πpp[iW, iw, :] - dpp)) < 1e-14
            assert np.max(np.abs(b[iW, iw, :] - S1.Πph[iW, iw, :] - dph)) < 1e-14


def test_nl_channel_evaluator(orc):
    """test/test_nonlocal_vertex.jl:55-95 (NL_Channel: shapes, call with fold-back, infinite frequencies, K switches)"""
    T, L = 0.5, 3
    F = _rand_nl(T, 5, (4, 3), (2, 3), L)
    g = F.γp
    assert g.K1.shape == (9, 9) and g.K2.shape == (7, 6, 9) and g.K3.shape == (3, 6, 6, 9)
    W, v, w = 1, 2, -1
    P_ = (-1, 1)
    iP = (P_[0] % L) + L * (P_[1] % L)
    K1, K2a, K2b, K3 = g.K1[W + 4, iP], g.K2[W + 3, v + 3, iP], g.K2[W + 3, w + 3, iP], g.K3[W + 1, v + 3, w + 3, iP]
    ev = lambda vv, ww, **kw: orc.eval_channel(F, L, 0, W, vv, ww, P_, (0, 0), (0, 0), **kw)
    assert abs(ev(v, w) - (K1 + K2a + K2b + K3)) < 1e-14
    assert abs(ev(INF, w) - (K1 + K2b)) < 1e-14
    assert abs(ev(v, INF) - (K1 + K2a)) < 1e-14
    assert abs(ev(INF, INF) - K1) < 1e-14
    # K switches (the pieces `reduce!` relies on, :99-108)
    assert abs(ev(v, w, K1=False, K2=False) - K3) < 1e-14
    assert abs(ev(INF, w, K1=False, K3=False) - K2b) < 1e-14
    assert abs(ev(v, INF, K1=False, K3=False) - K2a) < 1e-14
    assert abs(ev(INF, INF, K2=False, K3=False) - K1) < 1e-14
    # the fermionic momenta are ignored (:84-95)
    for vv in (v, INF):
        for ww in (w, INF):
            for K in ((True, True, True), (False, True, True), (True, False, True), (True, True, False), (False, False, True)):
                kw = dict(K1=K[0], K2=K[1], K3=K[2])
                assert orc.eval_channel(F, L, 0, W, vv, ww, P_, (1, 2), (2, 0), **kw) == ev(vv, ww, **kw)
    # s-wave point in the bosonic momentum = average over P (:172-200)
    for vv in (v, INF):
        for ww in (w, INF):
            avg = np.mean([orc.eval_channel(F, L, 0, W, vv, ww, (i, j), (0, 0), (0, 0)) for j in range(L) for i in range(L)])
            assert abs(orc.eval_channel(F, L, 0, W, vv, ww, "sw", (0, 0), (0, 0)) - avg) < 1e-14


def test_nl_vertex_channel_maps(orc):
    """test/test_nonlocal_vertex.jl:127-152: only K1 non-zero -> F in channel Ch = U + K1's at the converted arguments"""
    from otypes import pCh, tCh, aCh, pSp
    T, U, L = 0.5, 3.0, 3
    F = _rand_nl(T, 10, (4, 3), (2, 1), L, U=U)
    for g in F.channels():
        g.K2[...] = 0
        g.K3[...] = 0
    W, v, w = 1, 2, -1
    P, k, q = (-1, 1), (1, 1), (0, 1)

    def K1(g, m, mom):
        if abs(m) > 9:
            return 0.0
        return g.K1[m + 9, (mom[0] % L) + L * (mom[1] % L)]
    add = lambda *xs: tuple(sum(c) for c in zip(*xs))
    neg = lambda x: (-x[0], -x[1])
    ev = lambda vv, ww, ch: orc.eval_vertex(F, L, W, vv, ww, P, k, q, ch, pSp)
    for ch, g in ((pCh, F.γp), (tCh, F.γt), (aCh, F.γa)):
        assert abs(ev(INF, INF, ch) - (U + K1(g, W, P))) < 1e-14
    exp_p = U + K1(F.γp, W, P) + K1(F.γt, W - v - w - 1, add(P, neg(k), neg(q))) + K1(F.γa, v - w, add(k, neg(q)))
    exp_t = U + K1(F.γt, W, P) + K1(F.γp, W + v + w + 1, add(P, k, q)) + K1(F.γa, w - v, add(q, neg(k)))
    exp_a = U + K1(F.γa, W, P) + K1(F.γp, W + v + w + 1, add(P, k, q)) + K1(F.γt, v - w, add(k, neg(q)))
    assert abs(ev(v, w, pCh) - exp_p) < 1e-14
    assert abs(ev(v, w, tCh) - exp_t) < 1e-14
    assert abs(ev(v, w, aCh) - exp_a) < 1e-14


@pytest.mark.parametrize("v,w", [(2, -1), (INF, -1), (2, INF), (INF, INF)])
def test_nl_swave_evaluation_equals_explicit_average(orc, v, w):
    """test/test_nonlocal_vertex.jl:203-246: F(Ω, ν, ω, P, kSW, k) and F(Ω, ν, ω, P, k, kSW) == explicit BZ averages for every
    channel, spin component and the reference's five switch combinations"""
    from otypes import pCh, tCh, aCh, pSp, xSp, dSp
    T, L, W = 0.5, 3, 1
    F = _rand_nl(T, 10, (4, 3), (2, 1), L, U=2.0, seed=7)
    pts = [(i, j) for j in range(L) for i in range(L)]
    for P in (pts[0], pts[7]):
        for k in (pts[0], pts[3]):
            for ch in (aCh, pCh, tCh):
                for sp in (pSp, xSp, dSp):
                    for γa, γp, γt, F0 in ((1, 1, 1, 1), (1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1)):
                        kw = dict(F0=bool(F0), γp=bool(γp), γt=bool(γt), γa=bool(γa))
                        full = lambda a, b: orc.eval_vertex(F, L, W, v, w, P, a, b, ch, sp, **kw)
                        val1 = np.mean([full(q, k) for q in pts])
                        val2 = np.mean([full(k, q) for q in pts])
                        assert abs(full("sw", k) - val1) < 1e-13
                        assert abs(full(k, "sw") - val2) < 1e-13


# ------------------------------------------------------------------ test/test_nonlocal_fdPA.jl
def _pa_solver(orc, mu, t2=0.0):
    from otypes import RefVertex
    T, U, t1, nmax, nq, LG = 0.5, 2.0, 1.0, 3, 3, 24
    Gb = orc.hubbard_bare_Green(T, 4 * nmax, LG, μ=mu, t1=t1, t2=t2)
    S = orc.OracleNLSolver(4 * nmax, (nmax, nmax), (nmax, nmax), nq, Gb, np.zeros_like(Gb), np.zeros_like(Gb), RefVertex(T, U), T=T)
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


def test_nl_fdPA_equals_scPA_for_zero_reference(orc, converged_reference):
    """test/test_nonlocal_fdPA.jl:29-41"""
    S0, x0 = converged_reference
    S0_fd = _pa_solver(orc, 0.0)
    x1 = _solve(orc, S0_fd, "fdPA")
    assert np.max(np.abs(x1 - x0)) < 1e-10


def test_nl_converged_fdPA_matches_scPA_of_target(orc, converged_reference):
    """test/test_nonlocal_fdPA.jl:44-86 with the reference's own tolerances, the code as written (for the s-wave solver the SDE uses
    the own reducible vertex of each level, src/nonlocal/SDE.jl:24,61, and the reference's test passes as coded)"""
    from otypes import pCh, tCh, aCh, pSp, xSp
    S0, _ = converged_reference
    orc.Dyson(S0)
    S = _pa_solver(orc, 0.5, -0.3)
    _solve(orc, S, "scPA")
    Gb = orc.hubbard_bare_Green(0.5, 12, 24, μ=0.5, t1=1.0, t2=-0.3)
    Sfd = orc.OracleNLSolver(12, (3, 3), (3, 3), 3, Gb, S0.G, S0.Σ, S0.F, T=0.5)
    Sfd.init_sym_grp()
    _solve(orc, Sfd, "fdPA")
    assert np.max(np.abs(Sfd.Σ - S.Σ)) < 2e-4
    tols = {("K1", pCh): 2e-3, ("K1", aCh): 2e-3, ("K1", tCh): 2e-3, ("K2", pCh): 4e-3, ("K2", aCh): 4e-3, ("K2", tCh): 1e-3,
            ("K3", pCh): 1e-3, ("K3", aCh): 1e-3, ("K3", tCh): 2e-3}
    for (cls, ch), tol in tols.items():
        d = getattr(Sfd.F.channel(ch), cls) + getattr(Sfd.F0.channel(ch), cls) - getattr(S.F.channel(ch), cls)
        assert np.max(np.abs(d)) < tol, (ch, cls, np.max(np.abs(d)))
    L = 3
    P, k, kp = (1, 0), (2, 0), (0, 1)         # value(mK_Γ[2]), [3], [4]
    for ch in (pCh, tCh, aCh):
        for sp in (pSp, xSp):
            a = orc.eval_vertex(Sfd.F, L, 0, 1, -2, P, k, kp, ch, sp)
            b = orc.eval_vertex(S.F, L, 0, 1, -2, P, k, kp, ch, sp)
            assert abs(a - b) < 2e-3


def test_nl_symmetry_errors_of_unsymmetrised_scPA_solution(orc):
    """test/test_nonlocal_symmetry.jl:52-108: solve scPA (update_Σ = false) WITHOUT symmetry groups, then the symmetry error of every
    class under the s-wave solver's groups is below the reference's thresholds"""
    from otypes import RefVertex, pCh, tCh, aCh
    import oracle as o
    T, U, nmax, LG, L = 0.5, 2.0, 4, 6, 3
    Gb = orc.hubbard_bare_Green(T, 4 * nmax, LG, μ=-2.0, t1=1.0, t2=-0.5)
    S = orc.OracleNLSolver(4 * nmax, (nmax, nmax), (nmax, nmax), L, Gb, np.zeros_like(Gb), np.zeros_like(Gb), RefVertex(T, U), T=T)
    nF = len(S.F)

    def fp(x):
        S.F.unflatten(x)
        orc.iterate_solver(S, "scPA", False)
        return S.F.flatten() - x
    x, it, err = anderson(fp, S.F.flatten(), tol=1e-10)
    assert err < 1e-10
    S.F.unflatten(x)
    for ch in (pCh, tCh, aCh):
        for a in S.F.channel(ch).arrays():
            assert np.max(np.abs(a)) > 1e-3
    S.init_sym_grp()

    def sym_err(which, a):
        b = a.copy(order="F")
        orc.lib().orc_symmetrize(orc._p(b), __import__("ctypes").byref(orc.sg_struct(S.sg[which])))
        return np.max(np.abs(b - a))
    thr = {("K1", pCh): 1e-10, ("K1", aCh): 1e-10, ("K1", tCh): 1e-10, ("K2", pCh): 1e-3, ("K2", aCh): 1e-3, ("K2", tCh): 2e-3,
           ("K3", pCh): 2e-2, ("K3", aCh): 2e-3, ("K3", tCh): 2e-3}
    for (cls, ch), t in thr.items():
        which = {"K1": o.SG_K1, "K2": o.SG_PP2 if ch == pCh else o.SG_PH2, "K3": o.SG_PP3 if ch == pCh else o.SG_PH3}[cls]
        assert sym_err(which, getattr(S.F.channel(ch), cls)) < t, (cls, ch)


@pytest.mark.slow
def test_nl_fdPA_from_an_impurity_reference_has_symmetric_left_vertices(orc):
    """test/test_nonlocal_symmetry.jl:111-187: the production configuration of the s-wave solver -- reference system = a converged
    SIAM (local Vertex with a RefVertex core, impurity G and Σ on every lattice momentum), target = the Hubbard model, fdPA solved
    WITHOUT symmetry groups -- gives non-zero left vertices FL whose symmetry errors under the s-wave groups are below the
    reference's thresholds (the symmetries are exact only at convergence)"""
    import ctypes
    from otypes import RefVertex, Vertex, pCh, tCh, aCh
    import oracle as o
    T, U, μ, t1 = 0.5, 2.0, -2.0, 1.0
    D, e, Δ = 4 * t1, μ, 1.5
    nmax0 = 12
    Gb0 = orc.siam_bare_Green(T, 4 * nmax0, e=e, Δ=Δ, D=D)
    S0 = orc.OracleLocalSolver(4 * nmax0, (nmax0, nmax0), (nmax0, nmax0), Gb0, np.zeros_like(Gb0), np.zeros_like(Gb0), RefVertex(T, U), T=T)
    S0.init_sym_grp()
    nF0 = len(S0.F)

    def fp0(x):
        S0.F.unflatten(x[:nF0]); S0.Σ[...] = x[nF0:].reshape(S0.Σ.shape, order="F")
        orc.iterate_solver_local(S0, "scPA", True)
        return np.concatenate([S0.F.flatten(), S0.Σ.ravel(order="F")]) - x
    x, it, err = anderson(fp0, flatten_solver(S0), tol=1e-8)
    assert err < 1e-8
    S0.F.unflatten(x[:nF0]); S0.Σ[...] = x[nF0:].reshape(S0.Σ.shape, order="F")
    orc.Dyson(S0)
    # the converged impurity vertex as a local Vertex, G0 / Σ0 momentum independent
    Floc = Vertex(RefVertex(T, U), T, 4 * nmax0, (nmax0, nmax0), (nmax0, nmax0))
    for gl, g0 in zip(Floc.channels(), S0.F.channels()):
        gl.K1[...] = g0.K1[:, 0]; gl.K2[...] = g0.K2[:, :, 0, 0]; gl.K3[...] = g0.K3[:, :, :, 0]
    nmax, LG, L = 8, 6, 3
    nG = 4 * nmax
    Gbare = orc.hubbard_bare_Green(T, nG, LG, μ=μ, t1=t1)
    sl = slice(4 * nmax0 - nG, 4 * nmax0 + nG)
    G0 = np.asfortranarray(np.repeat(S0.G[sl, :], LG * LG, axis=1))
    Σ0 = np.asfortranarray(np.repeat(S0.Σ[sl, :], LG * LG, axis=1))
    S = orc.OracleNLSolver(4 * nmax, (nmax, nmax), (nmax, nmax), L, Gbare, G0, Σ0, Floc, T=T)
    x, it, err = anderson(oracle_fixed_point(orc, S, "fdPA"), flatten_solver(S), tol=1e-7)
    assert err < 1e-7, (it, err)
    for ch in (pCh, aCh, tCh):
        assert np.max(np.abs(S.FL.channel(ch).K2)) > 1e-4 and np.max(np.abs(S.FL.channel(ch).K3)) > 1e-4
    S.init_sym_grp()

    def sym_err(which, a):
        b = a.copy(order="F")
        orc.lib().orc_symmetrize(orc._p(b), ctypes.byref(orc.sg_struct(S.sg[which])))
        return np.max(np.abs(b - a))
    assert sym_err(o.SG_PP2, S.FL.γp.K2) < 1e-2 and sym_err(o.SG_PH2, S.FL.γa.K2) < 1e-2 and sym_err(o.SG_PH2, S.FL.γt.K2) < 1e-2
    assert sym_err(o.SG_PPL3, S.FL.γp.K3) < 1e-3 and sym_err(o.SG_PHL3, S.FL.γa.K3) < 1e-3 and sym_err(o.SG_PHL3, S.FL.γt.K3) < 1e-3
