"""GPU parity of the multi-boson-exchange parametrisation (MBEVertex / NL2_MBEVertex, src/boson_exchange.jl; nl_method = -2 of
script/run_Wu_point.jl) through the C-ABI against the CPU oracle: MBE K3 caches (explicit s-wave averages), every BSE entry point,
SDE and complete iterations, for the nonlocal solver on a small mesh (nested MBE reference chain) and for the local solver."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-10
_CACHES = ("cache_Γpx", "cache_F0p", "cache_F0a", "cache_F0t", "cache_Γpp", "cache_Γa", "cache_Γt", "cache_Fp", "cache_Fa", "cache_Ft")


def rel(a, b):
    s = max(np.max(np.abs(a)), np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / s)


def compare_vertex(Vg, Vo, what, classes=("K1", "K2", "K3")):
    for ch in range(3):
        for cls in classes:
            a, b = getattr(Vg.channel(ch), cls), getattr(Vo.channel(ch), cls)
            assert rel(a, b) < TOL, f"{what} ch={ch} {cls}: rel dev {rel(a, b):.3e}"


def make_nl2(orc, *, nested, sym=True, seed=1):
    """NL2 solver with S.F an NL2_MBEVertex.  nested: F0 = NL2_MBEVertex over a local MBEVertex over a RefVertex with a core (the
    state of an fd calculation); else the parquet approximation F0 = U"""
    import fddgasolver_jl_b200 as fd
    import oracle as o
    T, U, nmax, L, LG = 0.5, 2.0, 2, 2, 4
    rng = np.random.default_rng(seed)
    Gb = fd.hubbard_bare_Green(T, 4 * nmax, LG, μ=0.3, t1=1.0, t2=-0.2)
    if nested:
        core = fd.RefVertex(T, U, (2, 2), *[0.2 * (rng.random((3, 4, 4)) + 1j * rng.random((3, 4, 4))) for _ in range(4)])
        loc = fd.randomize_vertex(fd.MBEVertex(core, T, 12, (4, 4), (3, 3)), seed + 1, 0.3)
        F0 = fd.randomize_vertex(fd.NL2_MBEVertex(loc, T, 4 * nmax, (nmax, nmax), (nmax, nmax), L), seed + 2, 0.2)
        G0 = fd.hubbard_bare_Green(T, 4 * nmax, LG, μ=0.1, t1=1.0, t2=-0.2)
    else:
        F0, G0 = fd.RefVertex(T, U), np.zeros_like(Gb)
    S = fd.NL2_ParquetSolver(4 * nmax, (nmax, nmax), (nmax, nmax), L, Gb, G0, np.zeros_like(Gb), F0, T=T, VT=fd.NL2_MBEVertex)
    fd.randomize_vertex(S.F, seed + 3, 0.3)
    S.push("F")
    R = orc.OracleSolver(S.nK1, S.nK2, S.nK3, S.L, S.Gbare, S.G0, S.Σ0, S.F0, T=S.T, VT=o.NL2_MBEVertex)
    assert type(R.F0).__name__ == ("ONL2_MBEVertex" if nested else "ORefVertex")
    if sym:
        S.init_sym_grp(); R.init_sym_grp()
    R.F.set(S.F)
    return S, R


@pytest.mark.parametrize("nested,sym", [(True, True), (True, False), (False, True)])
def test_mbe_caches_and_bse_kernels_stepwise(orc, nested, sym):
    import fddgasolver_jl_b200 as fd
    S, R = make_nl2(orc, nested=nested, sym=sym)
    order = (fd.pCh, fd.aCh, fd.tCh)
    fd.build_K3_cache(S); orc.build_K3_cache(R)
    S.pull("cache")
    for n in _CACHES:
        assert rel(getattr(S, n), getattr(R, n)) < TOL, n
    for ch in order:
        fd.BSE_L_K2(S, ch); orc.BSE_L_K2(R, ch)
    for ch in order:
        fd.BSE_L_K3(S, ch); orc.BSE_L_K3(R, ch)
    S.pull("FL")
    compare_vertex(S.FL, R.FL, "FL", ("K2", "K3"))
    for ch in order:
        fd.BSE_K1(S, ch); orc.BSE_K1(R, ch)
    for ch in order:
        fd.BSE_K2(S, ch); orc.BSE_K2(R, ch)
    for ch in order:
        fd.BSE_K3(S, ch); orc.BSE_K3(R, ch)
    S.pull("Fbuff")
    compare_vertex(S.Fbuff, R.Fbuff, "Fbuff")
    S.close()


@pytest.mark.parametrize("strategy,nested", [("fdPA", True), ("scPA", True), ("scPA", False), ("fdPA_1loop", True)])
def test_mbe_iterate_solver_and_sde(orc, strategy, nested):
    import fddgasolver_jl_b200 as fd
    S, R = make_nl2(orc, nested=nested)
    for it in range(2):
        fd.iterate_solver(S, strategy); orc.iterate_solver(R, strategy)
        S.pull("F", "Σ", "G")
        compare_vertex(S.F, R.F, f"F it{it}")
        assert rel(S.Σ, R.Σ) < TOL and rel(S.G, R.G) < TOL
    S.close()


def test_mbe_contexts_have_no_fast_paths(orc):
    import fddgasolver_jl_b200 as fd
    S, _ = make_nl2(orc, nested=False)
    with pytest.raises(fd.FdgaError, match="generic kernels"):
        S.set_option("generic_kernels", 0)
    for fn in (fd.BSE_K1_new, fd.BSE_K2_new):
        with pytest.raises(fd.FdgaError, match="not available for MBE"):
            fn(S, fd.pCh)
    with pytest.raises(fd.FdgaError, match="not available for MBE"):
        fd.iterate_solver(S, "fdPA_new")
    S.close()
    T = 0.5
    Gb = fd.hubbard_bare_Green(T, 8, 4, μ=0.1, t1=1.0)
    bad = fd.NL2_Vertex(fd.NL2_MBEVertex(fd.RefVertex(T, 1.0), T, 8, (2, 2), (2, 2), 2), T, 8, (2, 2), (2, 2), 2)      # MBE below asymptotic
    with pytest.raises(fd.FdgaError, match="head of the chain"):
        fd.NL2_ParquetSolver(8, (2, 2), (2, 2), 2, Gb, Gb, np.zeros_like(Gb), bad, T=T)


def test_local_mbe_solver_matches_oracle_and_the_asymptotic_solution(orc):
    """ParquetSolver(...; VT = MBEVertex) on the device: two iterations against the oracle at 1e-10, then converged scPA solutions in
    the MBE and in the asymptotic parametrisation give the same self-energy and K1 (test/test_boson_exchange_local.jl:87-128 at
    smaller boxes, tolerances scaled accordingly)"""
    import fddgasolver_jl_b200 as fd
    import oracle as o
    from helpers import anderson
    T, U, D, e, Δ, nmax = 0.1, 1.0, 10.0, 0.5, np.pi / 5, 6
    mk = lambda VT: fd.parquet_solver_siam_parquet_approximation(6 * nmax, 4 * nmax, (nmax, nmax), (nmax, nmax), e=e, Δ=Δ, D=D, T=T, U=U, VT=VT)
    S = mk(fd.MBEVertex); S.init_sym_grp()
    Gb = orc.siam_bare_Green(T, 6 * nmax, e=e, Δ=Δ, D=D)
    R = orc.OracleLocalSolver(4 * nmax, (nmax, nmax), (nmax, nmax), Gb, np.zeros_like(Gb), np.zeros_like(Gb), o.RefVertex(T, U), T=T, VT=o.NL2_MBEVertex)
    R.init_sym_grp()
    for strategy in ("scPA", "fdPA"):
        fd.iterate_solver(S, strategy); orc.iterate_solver_local(R, strategy, True)
        S.pull("F", "Σ", "FL")
        assert rel(S.F.flatten(), R.F.flatten()) < TOL and rel(S.Σ, R.Σ) < TOL and rel(S.FL.flatten(), R.FL.flatten()) < TOL

    def converge(X):
        nF = X.length_F()
        x0 = np.concatenate([X.F.flatten() * 0, X.Σ.ravel(order="F") * 0])
        fp = lambda x: fd.fixed_point(np.empty_like(x), x, X, "scPA", True)
        x, it, err = anderson(fp, x0, tol=1e-8)
        assert err < 1e-8
        X.unflatten_F(x[:nF]); X.pull("F")
        return x[nF:]
    Σ1 = converge(S)
    A = mk(None); A.init_sym_grp()
    Σ2 = converge(A)
    assert np.max(np.abs(Σ1 - Σ2)) < 5e-5
    for n in ("γa", "γp", "γt"):
        assert np.max(np.abs(getattr(S.F, n).K1 - getattr(A.F, n).K1)) < 5e-5
    S.close(); A.close()


def _vertex_values(V, device_eval, pts, **kw):
    return np.array([device_eval(*p, **kw) for p in pts])


def test_eval_vertex_entry_point_matches_oracle(orc):
    """fdga_eval_vertex: the chain as a callable, asymptotic and MBE levels, Brillouin and s-wave points, infinite frequencies"""
    import fddgasolver_jl_b200 as fd
    INFo, INFd = (2 ** 31 - 1) // 4, fd.solver.INF_FREQ
    S, R = make_nl2(orc, nested=True)
    L = S.L
    rng = np.random.default_rng(5)
    n = 40
    W, v, w = rng.integers(-3, 4, n), rng.integers(-4, 4, n), rng.integers(-4, 4, n)
    P, k, q = (rng.integers(0, L * L, n) for _ in range(3))
    v[::7] = INFd
    w[3::9] = INFd
    for level in (0, 1, 2):
        for ch in range(3):
            for sp in range(3):
                for kw in (dict(), dict(F0=False), dict(γp=False, γa=False)):
                    for sw in (False, True):
                        got = S.eval_vertex(W, v, w, ch, sp, P, k, q, level=level, swave=sw, **kw)
                        V = [S.F, S.F0, S.F0.F0][level]
                        exp = np.array([orc.eval_vertex(V, L, int(W[i]), INFo if v[i] == INFd else int(v[i]), INFo if w[i] == INFd else int(w[i]),
                                                        (int(P[i]) % L, int(P[i]) // L), "sw" if sw else (int(k[i]) % L, int(k[i]) // L),
                                                        "sw" if sw else (int(q[i]) % L, int(q[i]) // L), ch, sp, **kw) for i in range(n)])
                        assert rel(got, exp) < TOL, (level, ch, sp, kw, sw)
    S.close()


def test_asymptotic_to_mbe_and_back_local(orc):
    """test/test_boson_exchange_local.jl:66-84: K3 branch (vertex with a K3 box) and the core branch (DMFT-like vertex whose reducible
    part has a dummy K3: the SBE contribution goes into the RefVertex core, src/boson_exchange.jl:694-707)"""
    import fddgasolver_jl_b200 as fd
    T, U = 0.5, 2.0
    F = fd.randomize_vertex(fd.Vertex(fd.RefVertex(T, U), T, 10, (5, 5), (3, 3)), 3, 1.0)
    Fm = fd.asymptotic_to_mbe(F)
    assert isinstance(Fm, fd.MBEVertex) and np.array_equal(Fm.γp.K1, F.γp.K1) and not np.array_equal(Fm.γa.K3, F.γa.K3)
    z = (0, 0)
    for ch in range(3):
        for sp in range(3):
            assert abs(orc.eval_vertex(Fm, 1, -1, -1, 1, z, z, z, ch, sp) - orc.eval_vertex(F, 1, -1, -1, 1, z, z, z, ch, sp)) < 1e-12
    Fn = fd.mbe_to_asymptotic(Fm)
    assert isinstance(Fn, fd.Vertex) and not isinstance(Fn, fd.MBEVertex)
    assert rel(Fn.flatten(), F.flatten()) < 1e-13
    # core branch
    rng = np.random.default_rng(8)
    core = fd.RefVertex(T, U, (3, 2), *[0.3 * (rng.random((5, 4, 4)) + 1j * rng.random((5, 4, 4))) for _ in range(4)])
    G = fd.randomize_vertex(fd.Vertex(core, T, 10, (5, 5), (1, 1)), 4, 1.0)
    for g in G.channels():
        g.K3[...] = 0
    Gm = fd.asymptotic_to_mbe(G)
    assert np.array_equal(Gm.γt.K2, G.γt.K2) and not np.array_equal(Gm.F0.Fp_p, core.Fp_p)
    # the parallel-spin components of the p and t channels, whose core arrays the conversion corrects directly, coincide inside the
    # core box (the other components follow by crossing symmetry for physical data; the random core here has none)
    for ch in (fd.pCh, fd.tCh):
        for (W, v, w) in ((0, 0, 1), (-1, -2, 1), (2, 1, -1)):
            assert abs(orc.eval_vertex(Gm, 1, W, v, w, z, z, z, ch, fd.pSp) - orc.eval_vertex(G, 1, W, v, w, z, z, z, ch, fd.pSp)) < 1e-12, (ch, W, v, w)
    # every corrected core array against the formula evaluated by the oracle: Λ -= F_mbe(...; F0 = false) - F(...; F0 = false)
    G0m = fd.MBEVertex(core.copy(), T, 10, (5, 5), (1, 1)); G0m.set(G)
    for name, ch, sp in (("Fp_p", fd.pCh, fd.pSp), ("Fp_x", fd.pCh, fd.xSp), ("Ft_p", fd.tCh, fd.pSp), ("Ft_x", fd.tCh, fd.xSp)):
        for (W, v, w) in ((0, 0, 1), (-1, -2, 1), (2, 1, -2)):
            nabla = orc.eval_vertex(G0m, 1, W, v, w, z, z, z, ch, sp, F0=False) - orc.eval_vertex(G, 1, W, v, w, z, z, z, ch, sp, F0=False)
            assert abs(getattr(Gm.F0, name)[W + 2, v + 2, w + 2] - (getattr(core, name)[W + 2, v + 2, w + 2] - nabla)) < 1e-12, (name, W, v, w)


def test_asymptotic_to_mbe_and_back_nl2(orc):
    """test/test_boson_exchange_NL2.jl:77-100: the two parametrisations share their s-wave component; round trip"""
    import fddgasolver_jl_b200 as fd
    T, U, L = 0.5, 2.0, 3
    F = fd.randomize_vertex(fd.NL2_Vertex(fd.RefVertex(T, U), T, 10, (5, 5), (3, 3), L), 6, 1.0)
    Fm = fd.asymptotic_to_mbe(F)
    assert isinstance(Fm, fd.NL2_MBEVertex)
    for ch in range(3):
        kw = dict(γp=ch == 0, γt=ch == 1, γa=ch == 2)
        for sp in range(3):
            a = orc.eval_vertex(Fm, L, 0, -2, 1, (0, 1), "sw", "sw", ch, sp, **kw)
            b = orc.eval_vertex(F, L, 0, -2, 1, (0, 1), "sw", "sw", ch, sp, **kw)
            assert abs(a - b) < 1e-12, (ch, sp)
    assert rel(fd.mbe_to_asymptotic(Fm).flatten(), F.flatten()) < 1e-13


def test_wu_point_mbe_construction_and_iteration(orc):
    """nl_method = -2 of script/run_Wu_point.jl:95-96 on a small mesh: F0 = NL2_MBEVertex(asymptotic_to_mbe(Γ_local)), one fdPA
    iteration with the self-energy against the oracle"""
    import fddgasolver_jl_b200 as fd
    import oracle as o
    S = fd.wu_point_solver(nmax=2, nq=2, LG=4, small_reference=True, F0_scale=0.02, F_scale=0.1, nl_method=-2)
    chain = fd.vertex_chain(S.F)
    assert [type(V).__name__ for V in chain] == ["NL2_MBEVertex", "NL2_MBEVertex", "MBEVertex", "RefVertex"]
    R = orc.OracleSolver(S.nK1, S.nK2, S.nK3, S.L, S.Gbare, S.G0, S.Σ0, S.F0, T=S.T, VT=o.NL2_MBEVertex)
    R.init_sym_grp()
    R.F.set(S.F)
    fd.iterate_solver(S, "fdPA"); orc.iterate_solver(R, "fdPA")
    S.pull("F", "Σ", "FL")
    compare_vertex(S.F, R.F, "F")
    compare_vertex(S.FL, R.FL, "FL", ("K2", "K3"))
    assert rel(S.Σ, R.Σ) < TOL
    S.close()


def test_mbe_mfrg_cache_kernels_and_linear_map(orc):
    """build_K3_cache_mfRG! is typed for any vertex (src/nonlocal_2/build_K3_cache.jl:97-164) and solve_using_mfRG! is what
    script/run_Wu_point.jl runs for nl_method = -2 too: mfRG caches, the mfRG branches of the BSE kernels, the linear map and a
    DQGMRES solve with MBE vertices against the oracle"""
    import fddgasolver_jl_b200 as fd
    S, R = make_nl2(orc, nested=True)
    order = (fd.pCh, fd.aCh, fd.tCh)
    for first in (True, False):
        fd.build_K3_cache_mfRG(S, first); orc.build_K3_cache_mfRG(R, first)
        S.pull("cache")
        for n in ("cache_Γpx", "cache_Γpp", "cache_Γa", "cache_Γt", "cache_Fp", "cache_Fa", "cache_Ft"):
            assert rel(getattr(S, n), getattr(R, n)) < TOL, (n, first)
    for ch in order:
        fd.BSE_L_K2(S, ch); orc.BSE_L_K2(R, ch)
    for ch in order:
        fd.BSE_K1(S, ch, True); orc.BSE_K1(R, ch, True)
    for ch in order:
        fd.BSE_K2(S, ch, True); orc.BSE_K2(R, ch, True)
    for ch in order:
        fd.BSE_L_K3(S, ch); orc.BSE_L_K3(R, ch)
    for ch in order:
        fd.BSE_K3(S, ch, True); orc.BSE_K3(R, ch, True)
    S.pull("Fbuff", "FL")
    compare_vertex(S.FL, R.FL, "FL", ("K2", "K3"))
    compare_vertex(S.Fbuff, R.Fbuff, "Fbuff(mfRG)")
    x = S.F.flatten() * 2.0
    A, B = fd.mfRGLinearMap(S), orc.mfRGLinearMap(R)
    for _ in range(2):
        yg, yo = A.matvec(x), B.matvec(x)
        assert rel(yg, yo) < TOL
        x = yo * 0.5
    rng = np.random.default_rng(2)
    b = rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size)
    xg, sg = fd.dqgmres(A, b, memory=5, atol=1e-9, rtol=1e-9, itmax=20)
    xo, so = orc.dqgmres(B, b, memory=5, atol=1e-9, rtol=1e-9, itmax=20)
    assert sg["niter"] == so["niter"] and rel(xg, xo) < 1e-8
    S.close()
