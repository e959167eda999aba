"""GPU parity of the s-wave solver (NL_ParquetSolver, src/nonlocal/) through the C-ABI against the CPU oracle on the same seeded
inputs: bubbles, every BSE entry point, caches, SDE, full iterations (fdPA / scPA), the mfRG linear map and its Krylov solve.
Tolerance 1e-10 relative to the largest entry of each array; symmetry tables bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-10
_CACHES = ("cache_Γpx", "cache_F0p", "cache_F0a", "cache_F0t", "cache_Γpp", "cache_Γa", "cache_Γt", "cache_Fp", "cache_Fa", "cache_Ft")


def rel(a, b):
    s = max(np.max(np.abs(a)), np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / s)


def make_pair(orc, *, nmax=2, nq=3, LG=6, sym=True, pa=False, F0_scale=0.03, seed=1, mΠν_factor=None):
    """(GPU NL_ParquetSolver, OracleNLSolver) with identical inputs.  pa: parquet approximation (F0 = RefVertex), else the Wu-point
    construction F0 = NL_Vertex(local DMFT-like vertex) with filled K's (state after an outer mfRG iteration)."""
    import fddgasolver_jl_b200 as fd
    if pa:
        S = fd.parquet_solver_hubbard_parquet_approximation(4 * nmax, 4 * nmax, (nmax, nmax), (nmax, nmax), LG, nq, T=0.5, U=2.0, μ=0.3, t1=1.0, t2=-0.2)
        fd.randomize_vertex(S.F, seed, 0.3)
        S.push("F")
        if sym:
            S.init_sym_grp()
    else:
        S = fd.wu_point_solver(nmax=nmax, nq=nq, LG=LG, small_reference=True, F0_scale=F0_scale, seed=seed, init_sym=sym, F_scale=0.2, nl_method=1)
    R = orc.OracleNLSolver(S.nK1, S.nK2, S.nK3, S.L, S.Gbare, S.G0, S.Σ0, S.F0, T=S.T)
    if sym:
        R.init_sym_grp()
    R.F.set(S.F)
    return S, R


def compare_vertex(Vg, Vo, what, classes=("K1", "K2", "K3")):
    for ch in range(3):
        for cls in classes:
            a, b = getattr(Vg.channel(ch), cls), getattr(Vo.channel(ch), cls)
            assert a.shape == b.shape
            assert rel(a, b) < TOL, f"{what} ch={ch} {cls}: rel dev {rel(a, b):.3e}"


def test_swave_symmetry_tables_bit_exact(orc):
    S, R = make_pair(orc, nmax=2, nq=4, LG=8)
    for which in range(8):
        for a, b in zip(S._sg[which], R.sg[which]):
            assert np.array_equal(a, b), which
    assert S.F.γp.K2.ndim == 3
    S.close()


@pytest.mark.parametrize("nq,LG", [(3, 6), (4, 8), (4, 4), (2, 6)])
def test_swave_bubbles_dyson(orc, nq, LG):
    """bubbles_real_space!(::NL_MF_Pi) with the 1/ν tail, odd / even meshes, LG == L (half weights at the zone edge)"""
    S, R = make_pair(orc, nmax=2, nq=nq, LG=LG)
    S.pull("Π", "G")
    assert S.Πpp.shape == R.Πpp.shape and S.Πpp.ndim == 3
    assert rel(S.G, R.G) < TOL
    for n in ("Π0pp", "Π0ph", "Πpp", "Πph"):
        assert rel(getattr(S, n), getattr(R, n)) < TOL, n
    S.close()


@pytest.mark.parametrize("sym,pa", [(True, False), (False, False), (True, True)])
def test_swave_bse_kernels_stepwise(orc, sym, pa):
    import fddgasolver_jl_b200 as fd
    S, R = make_pair(orc, sym=sym, pa=pa)
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


@pytest.mark.parametrize("sym", [True, False])
def test_swave_mfrg_kernels_and_matvec(orc, sym):
    import fddgasolver_jl_b200 as fd
    S, R = make_pair(orc, sym=sym)
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
    x = S.F.flatten() * 3.0
    A, B = fd.mfRGLinearMap(S), orc.mfRGLinearMap(R)
    for _ in range(2):
        yg, yo = A.matvec(x), B.matvec(x)
        assert rel(yg, yo) < TOL
        x = yo * 0.5
    S.close()


@pytest.mark.parametrize("strategy,pa", [("scPA", False), ("fdPA", False), ("scPA", True), ("fdPA", True)])
def test_swave_sde(orc, strategy, pa):
    import fddgasolver_jl_b200 as fd
    S, R = make_pair(orc, pa=pa)
    fd.SDE(S, strategy); orc.SDE(R, strategy)
    S.pull("Σ")
    assert rel(S.Σ, R.Σ) < TOL
    # the L arrays of the target system alone (SDE_channel_L_pp! / ph!), level 0 of the chain
    fd.SDE_channel_L(S, False, 0)
    S.pull("L")
    Lpp, Lph = np.zeros_like(R.Lpp), np.zeros_like(R.Lph)
    orc.SDE_channel_L(R, Lpp, R.Πpp, R.F, 0, True); orc.SDE_channel_L(R, Lph, R.Πph, R.F, 0, False)
    chain_len = len(fd.vertex_chain(S.F))
    if chain_len == 2:          # parquet approximation: the chain below level 0 is the bare vertex, whose L is zero
        assert rel(S.Lpp, Lpp) < TOL and rel(S.Lph, Lph) < TOL
    S.close()


@pytest.mark.parametrize("strategy,sym,pa", [("fdPA", True, False), ("scPA", True, False), ("fdPA", False, False), ("fdPA", True, True)])
def test_swave_iterate_solver(orc, strategy, sym, pa):
    """two complete iterate_solver! calls (Dyson, bubbles, cache, BSE stages, SDE) on fused lanes vs the oracle"""
    import fddgasolver_jl_b200 as fd
    S, R = make_pair(orc, sym=sym, pa=pa)
    for it in range(2):
        fd.iterate_solver(S, strategy); orc.iterate_solver(R, strategy)
        S.pull("F", "Σ", "G", "FL")
        compare_vertex(S.F, R.F, f"F it{it}")
        assert rel(S.Σ, R.Σ) < TOL and rel(S.G, R.G) < TOL
    S.close()


def test_swave_variant_strategies_are_rejected(orc):
    import fddgasolver_jl_b200 as fd
    S, _ = make_pair(orc)
    for fn in (fd.BSE_K1_new, fd.BSE_K2_new, fd.BSE_K1_1loop, fd.BSE_K2_1loop, fd.BSE_K3_1loop):
        with pytest.raises(fd.FdgaError, match="s-wave"):
            fn(S, fd.pCh)
    with pytest.raises(fd.FdgaError, match="s-wave"):
        fd.bubbles_momentum_space(S)
    S.close()


def test_swave_level_type_mismatch_is_rejected(orc):
    """an NL2_Vertex in the F0 chain of the s-wave solver (and vice versa) is refused by the host mirror and by fdga_create"""
    import fddgasolver_jl_b200 as fd
    T, L = 0.5, 2
    Gb = fd.hubbard_bare_Green(T, 8, 4, μ=0.1, t1=1.0)
    bad = fd.NL2_Vertex(fd.RefVertex(T, 1.0), T, 8, (2, 2), (2, 2), L)
    with pytest.raises(fd.FdgaError):
        fd.NL_ParquetSolver(8, (2, 2), (2, 2), L, Gb, Gb, np.zeros_like(Gb), bad, T=T)
    bad2 = fd.NL_Vertex(fd.RefVertex(T, 1.0), T, 8, (2, 2), (2, 2), L)
    with pytest.raises(fd.FdgaError):
        fd.NL2_ParquetSolver(8, (2, 2), (2, 2), L, Gb, Gb, np.zeros_like(Gb), bad2, T=T)


def test_swave_converged_fdPA_matches_scPA_on_device(orc):
    """test/test_nonlocal_fdPA.jl on the device: scPA reference solution, fdPA of the shifted target from it, the reference's
    tolerances (solved with the package's own solve = NLsolve-style Anderson on the device residual)"""
    import fddgasolver_jl_b200 as fd
    T, U, nmax, nq, LG = 0.5, 2.0, 3, 3, 24
    mk = lambda μ, t2: fd.parquet_solver_hubbard_parquet_approximation(4 * nmax, 4 * nmax, (nmax, nmax), (nmax, nmax), LG, nq, T=T, U=U, μ=μ, t1=1.0, t2=t2)
    S0 = mk(0.0, 0.0); S0.init_sym_grp()
    fd.solve(S0, strategy="scPA", tol=1e-9)
    S = mk(0.5, -0.3); S.init_sym_grp()
    fd.solve(S, strategy="scPA", tol=1e-9)
    S0.pull("F", "Σ", "G"); S.pull("F", "Σ")
    Gb = fd.hubbard_bare_Green(T, 4 * nmax, LG, μ=0.5, t1=1.0, t2=-0.3)
    fd.Dyson(S0); S0.pull("G")
    Sfd = fd.NL_ParquetSolver(4 * nmax, (nmax, nmax), (nmax, nmax), nq, Gb, S0.G, S0.Σ, S0.F, T=T)
    Sfd.init_sym_grp()
    fd.solve(Sfd, strategy="fdPA", tol=1e-9)
    Sfd.pull("F", "Σ")
    assert np.max(np.abs(Sfd.Σ - S.Σ)) < 2e-4
    tols = {("K1", 0): 2e-3, ("K1", 2): 2e-3, ("K1", 1): 2e-3, ("K2", 0): 4e-3, ("K2", 2): 4e-3, ("K2", 1): 1e-3,
            ("K3", 0): 1e-3, ("K3", 2): 1e-3, ("K3", 1): 2e-3}
    for (cls, ch), tol in tols.items():
        d = getattr(Sfd.F.channel(ch), cls) + getattr(Sfd.F0.channel(ch), cls) - getattr(S.F.channel(ch), cls)
        assert np.max(np.abs(d)) < tol, (cls, ch, np.max(np.abs(d)))
    for X in (S0, S, Sfd):
        X.close()


def test_swave_dqgmres_and_preconditioned_fixed_point(orc):
    import fddgasolver_jl_b200 as fd
    S, R = make_pair(orc)
    rng = np.random.default_rng(7)
    n = S.length_F()
    assert n == len(R.F)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    xg, sg = fd.dqgmres(fd.mfRGLinearMap(S), b, memory=5, atol=1e-9, rtol=1e-9, itmax=30)
    xo, so = orc.dqgmres(orc.mfRGLinearMap(R), b, memory=5, atol=1e-9, rtol=1e-9, itmax=30)
    assert sg["niter"] == so["niter"] and sg["solved"] == so["solved"]
    assert np.max(np.abs(np.array(so["residuals"]) - np.array(sg["residuals"])) / so["residuals"][0]) < 1e-9
    assert rel(xg, xo) < 1e-8
    x = S.F.flatten() * (1.0 + 0.05 * rng.standard_normal(n))
    Rg, Ro = np.zeros_like(x), np.zeros_like(x)
    a = fd.fixed_point_preconditioned(Rg, x, S, strategy="fdPA", use_preconditioner=True, krylov_maxiter=25, memory=8)
    c = orc.fixed_point_preconditioned(Ro, x, R, strategy="fdPA", use_preconditioner=True, krylov_maxiter=25, memory=8)
    assert a == c and rel(Rg, Ro) < 1e-8
    S.close()


def test_swave_solve_using_mfRG_outer_loop_and_checkpoint(orc, tmp_path):
    """solve_using_mfRG! with the s-wave solver (the call of script/run_Wu_point.jl:100 for nl_method = 1) against the oracle's
    restatement: bubble mixing, chemical potential, preconditioned Anderson, SDE, reference update; then a checkpoint written by
    save_solver restores the device state of a fresh solver"""
    import fddgasolver_jl_b200 as fd
    T, U, nG, LG, L = 0.5, 2.0, 8, 6, 3
    hp = {"t1": 1.0, "t2": -0.3}
    Gb = fd.hubbard_bare_Green(T, nG, LG, μ=0.3, **hp)
    G0 = fd.hubbard_bare_Green(T, nG, LG, μ=0.1, **hp)
    mk = lambda: fd.NL_Vertex(fd.RefVertex(T, U), T, 8, (2, 2), (2, 2), L)
    new = lambda: fd.NL_ParquetSolver(8, (2, 2), (2, 2), L, Gb, G0, np.zeros_like(G0), mk(), T=T, mΠν_factor=4)
    S = new(); S.init_sym_grp()
    R = orc.OracleNLSolver(8, (2, 2), (2, 2), L, Gb, G0, np.zeros_like(G0), mk(), T=T, mΠν_factor=4)
    R.init_sym_grp()
    kw = dict(occ_target=0.45, hubbard_params=hp, mixing_init=0.5, tol=1e-5, strategy="fdPA", anderson_iterations=30, krylov_maxiter=40, memory=10)
    hg = fd.solve_using_mfRG(S, maxiter=2, **kw)
    ho = orc.solve_using_mfRG(R, maxiter=2, **kw)
    assert len(hg["Σ_err"]) == len(ho["Σ_err"]) == 2 and hg["mixing"] == ho["mixing"]
    assert np.allclose(hg["Σ_err"], ho["Σ_err"], rtol=1e-4) and np.allclose(hg["μ"], ho["μ"], rtol=0, atol=1e-5)
    S.pull("F", "F0", "Σ", "Σ0", "G", "G0", "Π")
    assert S.Πpp.ndim == 3
    assert rel(S.Σ0, R.Σ0) < 1e-4 and rel(S.G, R.G) < 1e-4 and rel(S.Π0pp, R.Π0pp) < 1e-4 and rel(S.Πph, R.Πph) < 1e-4
    for a, b in zip(S.F0.channels(), R.F0.channels()):
        for x, y in zip(a.arrays(), b.arrays()):
            assert rel(x, y) < 1e-4
    assert np.max(np.abs(S.F0.flatten())) > 0.02 and not np.any(S.F.flatten())
    # checkpoint round trip through the reference's file layout
    path = str(tmp_path / "swave.iter2.h5")
    fd.save_solver(S, path, extra={"mixing": hg["mixing"][-1]})
    S2 = new(); S2.init_sym_grp()
    fd.load_solver(S2, path)
    fd.iterate_solver(S, "fdPA"); fd.iterate_solver(S2, "fdPA")
    S.pull("F", "Σ"); S2.pull("F", "Σ")
    assert rel(S.F.flatten(), S2.F.flatten()) < 1e-13 and rel(S.Σ, S2.Σ) < 1e-13
    S.close(); S2.close()


@pytest.mark.slow
def test_swave_fullsize_config3_iteration_and_matvec(orc):
    """BASELINE config 3 sizes with the s-wave solver on the reference's DMFT data (script/run_Wu_point.jl, nl_method = 1):
    nmax = 4, nq = 8, LG = 48, bubble mesh 16 x 512 (mΠν_factor = 32).  One complete fdPA iteration with the self-energy update
    and one mfRG matvec, every array compared un-sampled."""
    import fddgasolver_jl_b200 as fd
    S = fd.wu_point_solver(nmax=4, nq=8, LG=48, nl_method=1, F_scale=0.05, F0_scale=0.02)
    assert S.nΠF == 512 and S.F.γp.K2.shape == (7, 8, 64)
    R = orc.OracleNLSolver(S.nK1, S.nK2, S.nK3, S.L, S.Gbare, S.G0, S.Σ0, S.F0, T=S.T)
    R.init_sym_grp()
    R.F.set(S.F)
    S.pull("Π")
    for n in ("Π0pp", "Π0ph", "Πpp", "Πph"):
        assert rel(getattr(S, n), getattr(R, n)) < TOL, n
    fd.iterate_solver(S, "fdPA"); orc.iterate_solver(R, "fdPA")
    S.pull("F", "FL", "Σ", "G")
    compare_vertex(S.F, R.F, "F")
    compare_vertex(S.FL, R.FL, "FL", ("K2", "K3"))
    assert rel(S.Σ, R.Σ) < TOL and rel(S.G, R.G) < TOL
    x = S.F.flatten() * 2.0
    assert rel(fd.mfRGLinearMap(S).matvec(x), orc.mfRGLinearMap(R).matvec(x)) < TOL
    S.close()
