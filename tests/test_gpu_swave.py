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
