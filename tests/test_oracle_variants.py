"""The `_new` and `_1loop` BSE variants of the oracle (src/nonlocal_2/BSEa/BSEa_K1.jl:62-113, BSEa_K2.jl:142-216,
BSE_1loop.jl) have no test in the reference ("parity unpinned").  They are tied here, by exact
algebraic identities, to the kernels that ARE pinned by the reference's tests (BSE_K1!/K2!/K3!):

 * FL = 0                               =>  BSE_K1_1loop!(fd) == BSE_K1!(fd),  BSE_K2_1loop!(fd) == BSE_K2!(fd)
 * FL = 0 and cache_Γ = 0               =>  BSE_K3_1loop!(fd) == BSE_K3!(fd)
 * mfRG branches                        =>  K1/K2 1loop == BSE_K1!/K2!(mfRG);  K3 1loop == BSE_K3!(mfRG) - FL.K3 post-add
 * F0 = RefVertex(U), Π0 = 0, FL = 0    =>  BSE_K1_new! == BSE_K1!  (the right vertex is the bare U)
 * same, Π zeroed outside the K2 ν-box  =>  BSE_K2_new! == BSE_K2!  (`_new` sums ω over the K2 mesh only)
"""
import numpy as np
import pytest


def _wu_oracle(orc, sym=True, seed=1):
    import fddgasolver_jl_b200 as fd
    inp = fd.wu_point_inputs(2, 3, 6, seed=seed, F_scale=0.2, F0_scale=0.03, small_reference=True)
    R = orc.OracleSolver(inp["nK1"], inp["nK2"], inp["nK3"], inp["L"], inp["Gbare"], inp["G0"], inp["Σ0"], inp["F0"], T=inp["T"])
    if sym:
        R.init_sym_grp()
    R.F.set(inp["F"])
    return R


def _pa_oracle(orc, seed=5):
    import fddgasolver_jl_b200 as fd
    from fddgasolver_jl_b200.types import RefVertex
    T, U, nG, LG, L = 0.5, 2.0, 8, 6, 3
    Gb = orc.hubbard_bare_Green(T, nG, LG, μ=0.3, t1=1.0, t2=-0.2)
    R = orc.OracleSolver(8, (2, 2), (2, 2), L, Gb, np.zeros_like(Gb), np.zeros_like(Gb), RefVertex(T, U), T=T)
    R.init_sym_grp()
    fd.randomize_vertex(R.F, seed, 0.3)
    return R


def _snap(V):
    return [a.copy() for g in V.channels() for a in g.arrays()]


def _maxdiff(A, B):
    return max(float(np.max(np.abs(a - b))) for a, b in zip(A, B))


def _zero(V):
    for g in V.channels():
        for a in g.arrays():
            a[...] = 0


ORDER = None


@pytest.fixture(autouse=True)
def _order():
    global ORDER
    import fddgasolver_jl_b200 as fd
    ORDER = (fd.pCh, fd.aCh, fd.tCh)


@pytest.mark.parametrize("sym", [True, False])
def test_1loop_fd_equals_full_kernels_for_zero_FL(orc, sym):
    R = _wu_oracle(orc, sym)
    orc.build_K3_cache(R)
    _zero(R.FL)
    for n in ("cache_Γpx", "cache_Γa", "cache_Γt"):
        getattr(R, n)[...] = 0
    for ch in ORDER:
        orc.BSE_K1(R, ch)
    for ch in ORDER:
        orc.BSE_K2(R, ch)
    for ch in ORDER:
        orc.BSE_K3(R, ch)
    full = _snap(R.Fbuff)
    _zero(R.Fbuff)
    for ch in ORDER:
        orc.BSE_K3_1loop(R, ch)
    for ch in ORDER:
        orc.BSE_K1_1loop(R, ch)
    for ch in ORDER:
        orc.BSE_K2_1loop(R, ch)
    assert max(np.max(np.abs(a)) for a in full) > 1e-4
    assert _maxdiff(full, _snap(R.Fbuff)) < 1e-14


def test_1loop_mfrg_equals_full_mfrg_kernels(orc):
    import fddgasolver_jl_b200 as fd
    R = _wu_oracle(orc)
    orc.build_K3_cache_mfRG(R, True)
    fd.randomize_vertex(R.FL, 11, 0.1)
    import ctypes
    for ch in ORDER:      # the K3 post-add happens inside the class fill: make FL.K3 symmetric under the K3 groups
        sg = R.sg[orc.SG_PP3 if ch == fd.pCh else orc.SG_PH3]
        orc.lib().orc_symmetrize(orc._p(R.FL.channel(ch).K3), ctypes.byref(orc.sg_struct(sg)))
    for ch in ORDER:
        orc.BSE_K1(R, ch, True)
    for ch in ORDER:
        orc.BSE_K2(R, ch, True)
    for ch in ORDER:
        orc.BSE_K3(R, ch, True)
    full = _snap(R.Fbuff)
    # BSE_K3!(mfRG) adds FL.K3 (a, p) resp. 2 FLt - FLa (t, before the d -> p spin fix, which turns it into FLt);
    # the 1-loop variant does not
    post = {fd.pCh: R.FL.γp.K3, fd.aCh: R.FL.γa.K3, fd.tCh: R.FL.γt.K3}
    _zero(R.Fbuff)
    for ch in ORDER:
        orc.BSE_K3_1loop(R, ch, True)
    for ch in ORDER:
        orc.BSE_K1_1loop(R, ch, True)
    for ch in ORDER:
        orc.BSE_K2_1loop(R, ch, True)
    for ch in ORDER:
        R.Fbuff.channel(ch).K3[...] += post[ch]
    assert _maxdiff(full, _snap(R.Fbuff)) < 1e-13


def test_new_forms_equal_full_kernels_for_bare_reference(orc):
    import fddgasolver_jl_b200 as fd
    R = _pa_oracle(orc)
    assert np.all(R.Π0pp == 0) and np.all(R.Π0ph == 0)
    _zero(R.FL)
    for ch in ORDER:
        orc.BSE_K1(R, ch)
    k1 = [R.Fbuff.channel(ch).K1.copy() for ch in ORDER]
    for ch in ORDER:
        R.Fbuff.channel(ch).K1[...] = 0
    for ch in ORDER:
        orc.BSE_K1_new(R, ch)
    assert max(np.max(np.abs(a)) for a in k1) > 1e-3
    assert _maxdiff(k1, [R.Fbuff.channel(ch).K1 for ch in ORDER]) < 1e-14
    # K2: window the bubbles to the K2 fermionic box
    nΠF, nf = R.nΠF, R.nK2[1]
    for Π in (R.Πpp, R.Πph):
        Π[:, : nΠF - nf] = 0
        Π[:, nΠF + nf:] = 0
    for ch in ORDER:
        orc.BSE_K2(R, ch)
    k2 = [R.Fbuff.channel(ch).K2.copy() for ch in ORDER]
    for ch in ORDER:
        R.Fbuff.channel(ch).K2[...] = 0
    for ch in ORDER:
        orc.BSE_K2_new(R, ch)
    assert max(np.max(np.abs(a)) for a in k2) > 1e-3
    assert _maxdiff(k2, [R.Fbuff.channel(ch).K2 for ch in ORDER]) < 1e-14


@pytest.mark.parametrize("strategy", ["scPA_new", "fdPA_new", "fdPA_1loop"])
def test_iterate_solver_variant_strategies_run_and_keep_symmetry(orc, strategy):
    """iterate_solver!(S; strategy) for the variant strategies (src/solve.jl:26-58): finite, non-trivial, symmetric."""
    R = _wu_oracle(orc)
    orc.iterate_solver(R, strategy, update_Σ=True)
    x = R.F.flatten()
    assert np.isfinite(x).all() and np.max(np.abs(x)) > 1e-6 and np.isfinite(R.Σ).all()
    for which, arrs in ((orc.SG_K1, [g.K1 for g in R.F.channels()]), (orc.SG_PP2, [R.F.γp.K2]), (orc.SG_PH2, [R.F.γa.K2, R.F.γt.K2])):
        for a in arrs:
            b = a.copy(order="F")
            orc.lib().orc_symmetrize(orc._p(b), __import__("ctypes").byref(orc.sg_struct(R.sg[which])))
            assert np.max(np.abs(a - b)) < 1e-13
