"""Parity against the CPU oracle AT THE SIZES THE BENCH QUOTES ("close the parity holes"):

  * BASELINE config 3 (nmax = 4, nq = 8, LG = 48) in FULL, nothing sampled: bubbles, all ten caches, FL, every class of every
    channel of F (p, a AND t), the self-energy, and one mfRG matvec;
  * nq = 16 (config 4 size, lanes off, 1 GB bubbles) and nK1 = 64 (sweep size): every kernel on sampled class representatives,
    the K3 kernels and caches in full / on sampled ranges, and the self-energy through the oracle's real-space contraction of
    the device's L arrays.

Tolerance 1e-10 relative to the largest entry of each array (BASELINE.json north_star)."""
import ctypes as C

import numpy as np
import pytest

from test_gpu_parity import TOL, compare_vertex, rel

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

CACHES = ("cache_Γpx", "cache_F0p", "cache_F0a", "cache_F0t", "cache_Γpp", "cache_Γa", "cache_Γt", "cache_Fp", "cache_Fa", "cache_Ft")


def oracle_twin(orc, fd, S, inp, compute_bubbles):
    R = orc.OracleSolver(inp["nK1"], inp["nK2"], inp["nK3"], inp["L"], inp["Gbare"], inp["G0"], inp["Σ0"], inp["F0"], T=inp["T"],
                         compute_bubbles=compute_bubbles)
    if not compute_bubbles:
        orc.Dyson(R)            # G from (Gbare, Σ0), as the constructor does between the two bubble evaluations
    R.init_sym_grp()
    for which in range(8):      # index / symmetry tables: bit exact
        for a, b in zip(S._sg[which], R.sg[which]):
            assert np.array_equal(a, b), which
    R.F.set(inp["F"])
    return R


def test_config3_full_iteration_sde_and_mfrg_matvec(orc):
    import fddgasolver_jl_b200 as fd
    inp = fd.wu_point_inputs(4, 8, 48, F0_scale=0.02)
    S = fd.wu_point_solver(4, 8, 48, F0_scale=0.02)
    R = oracle_twin(orc, fd, S, inp, True)
    S.pull("Π", "G")
    assert rel(S.G, R.G) < TOL
    for n in ("Π0pp", "Π0ph", "Πpp", "Πph"):
        assert rel(getattr(S, n), getattr(R, n)) < TOL, n
    fd.iterate_solver(S, "fdPA", False); orc.iterate_solver(R, "fdPA", False)
    S.pull("cache", "FL", "F")
    for n in CACHES:
        assert rel(getattr(S, n), getattr(R, n)) < TOL, n
    compare_vertex(S.FL, R.FL, "FL", ("K2", "K3"))
    compare_vertex(S.F, R.F, "F")                     # K1, K2, K3 of p, t and a, every class
    fd.SDE(S, "scPA"); orc.SDE(R, "scPA")
    S.pull("Σ")
    assert rel(S.Σ, R.Σ) < TOL
    x = R.F.flatten() * 3.0
    A, B = fd.mfRGLinearMap(S), orc.mfRGLinearMap(R)
    yg, yo = A.matvec(x), B.matvec(x)
    n = len(x) // 3
    for c in range(3):
        assert rel(yg[c * n:(c + 1) * n], yo[c * n:(c + 1) * n]) < TOL, c
    S.pull("cache")
    for n_ in ("cache_Γpx", "cache_Γpp", "cache_Γa", "cache_Γt", "cache_Fp", "cache_Fa", "cache_Ft"):
        assert rel(getattr(S, n_), getattr(R, n_)) < TOL, ("mfRG", n_)
    S.close()


def _class_members(sg, c):
    return sg[1][sg[0][c]:sg[0][c + 1]]


def _sampled_parity(orc, nmax, nq, LG, seed, cache_samples):
    import fddgasolver_jl_b200 as fd
    inp = fd.wu_point_inputs(nmax, nq, LG, F0_scale=0.02)
    S = fd.wu_point_solver(nmax, nq, LG, F0_scale=0.02)
    R = oracle_twin(orc, fd, S, inp, False)
    S.pull("Π")
    R.Π0pp, R.Π0ph, R.Πpp, R.Πph = S.Π0pp, S.Π0ph, S.Πpp, S.Πph        # 1 GB each: the bubbles are checked at config 3 and below
    rng = np.random.default_rng(seed)
    order = (fd.pCh, fd.aCh, fd.tCh)
    flat = lambda a: a.ravel(order="F")

    fd.iterate_solver(S, "fdPA", False)
    S.pull("cache", "FL", "Fbuff")
    # caches: sampled index ranges (the oracle recomputes every Brillouin-zone mean per element)
    n3 = S.cache_Γpx.size
    for i0 in rng.integers(0, n3 - cache_samples, size=3):
        orc.build_K3_cache(R, int(i0), int(i0) + cache_samples)
        for n in CACHES:
            a, b = flat(getattr(S, n))[i0:i0 + cache_samples], flat(getattr(R, n))[i0:i0 + cache_samples]
            assert rel(a, b) < TOL, (n, i0)
    for n in CACHES:      # the K3 kernels below get the device caches as input
        getattr(R, n)[...] = getattr(S, n)
    # BSE_L_K2!: a before t (the t result is post-fixed with a, BSE_templates.jl:72-73)
    sg = {ch: S._sg[fd._lib.SG_PP2 if ch == fd.pCh else fd._lib.SG_PH2] for ch in order}
    for c in rng.integers(0, len(sg[fd.pCh][0]) - 1, size=3):
        for ch in order:
            orc.BSE_L_K2(R, ch, c0=int(c), c1=int(c) + 1)
            idx = _class_members(sg[ch], c)
            assert rel(flat(S.FL.channel(ch).K2)[idx], flat(R.FL.channel(ch).K2)[idx]) < TOL, ("L_K2", ch, c)
    # BSE_L_K3! in full
    R.FL.set(S.FL)
    dev_FL_K3 = [S.FL.channel(ch).K3.copy() for ch in range(3)]
    for ch in order:
        orc.BSE_L_K3(R, ch)
    for ch in range(3):
        assert rel(dev_FL_K3[ch], R.FL.channel(ch).K3) < TOL, ("L_K3", ch)
    R.FL.set(S.FL)
    # BSE_K1! / BSE_K2!: sampled classes, all three channels (a before t)
    sg1 = S._sg[fd._lib.SG_K1]
    for c in rng.integers(0, len(sg1[0]) - 1, size=3):
        for ch in order:
            orc.BSE_K1(R, ch, c0=int(c), c1=int(c) + 1)
            idx = _class_members(sg1, c)
            assert rel(flat(S.Fbuff.channel(ch).K1)[idx], flat(R.Fbuff.channel(ch).K1)[idx]) < TOL, ("K1", ch, c)
    for c in rng.integers(0, len(sg[fd.pCh][0]) - 1, size=3):
        for ch in order:
            R.Fbuff.channel(ch).K2[...] = 0
        for ch in order:
            orc.BSE_K2(R, ch, c0=int(c), c1=int(c) + 1)
            idx = _class_members(sg[ch], c)
            assert rel(flat(S.Fbuff.channel(ch).K2)[idx], flat(R.Fbuff.channel(ch).K2)[idx]) < TOL, ("K2", ch, c)
    # BSE_K3! in full
    for ch in order:
        orc.BSE_K3(R, ch)
    for ch in range(3):
        assert rel(S.Fbuff.channel(ch).K3, R.Fbuff.channel(ch).K3) < TOL, ("K3", ch)

    # SDE!: the L arrays on sampled classes (all levels of the chain, 1/3 for the RefVertex level), then the full self-energy
    # through the oracle's real-space contraction + U^2 + Hartree of the DEVICE L arrays
    S.pull("F")
    R.F.set(S.F)
    fd.SDE_channel_L(S)
    S.pull("L")
    chain = orc.vertex_chain(R.F)
    for is_pp, Ldev, Π, which in ((True, S.Lpp, R.Πpp, fd._lib.SG_PP2), (False, S.Lph, R.Πph, fd._lib.SG_PH2)):
        sgL = S._sg[which]
        for c in rng.integers(0, len(sgL[0]) - 1, size=3):
            acc = np.zeros_like(Ldev)
            for level in range(len(chain)):
                tmp = np.zeros_like(Ldev)
                orc.SDE_channel_L(R, tmp, Π, R.F, level, is_pp, int(c), int(c) + 1)
                acc += tmp * (1.0 / 3.0 if isinstance(chain[level], orc.RefVertex) else 1.0)
            idx = _class_members(sgL, c)
            assert rel(flat(Ldev)[idx], flat(acc)[idx]) < TOL, ("L", is_pp, c)
    Σ = np.zeros(R.Σ.shape, dtype=np.complex128, order="F")
    sgΣ = orc.sg_struct(R.sg[0])
    orc.lib().orc_sde_real_space(orc._p(Σ), R.nG, R.LG, orc._p(R.G), R.nG, R.LG, orc._p(S.Lpp), orc._p(S.Lph), R.nK2[0], R.nK2[1],
                                 C.byref(sgΣ), C.byref(R.grid))
    U = R.F.bare_vertex()
    ΣU2 = np.zeros_like(Σ)
    orc.lib().orc_sde_U2(orc._p(ΣU2), orc._p(R.G), R.nG, R.LG, C.c_double(U.real), C.c_double(U.imag), C.c_double(R.T), C.byref(sgΣ))
    Σ += ΣU2 + (orc.compute_occupation(R, R.G) - 0.5) * U * 1j
    fd.SDE(S, "scPA")
    S.pull("Σ")
    assert rel(S.Σ, Σ) < TOL
    S.close()


def test_nq16_sampled_classes_and_self_energy(orc):
    """config-4 size: 16 x 16 momentum mesh, 1 GB per bubble: the concurrency lanes switch off above 160 MB (fdga_lib.cu:
    lanes_enabled), 3.67 M K2 elements per channel"""
    _sampled_parity(orc, 4, 16, 48, 11, 64)


def test_nK1_64_sampled_classes_and_self_energy(orc):
    """sweep size nmax = 16 (nK1 = 64, nK2 = nK3 = (16, 16)) on the 8 x 8 mesh: 31-wide K2 bosonic box, 2 M K3 elements"""
    _sampled_parity(orc, 16, 8, 48, 12, 32)
