"""GPU tests of the LOCAL (impurity) solver path (BASELINE configs[0], configs[1]): parity with the CPU oracle and the
reference's own golden numbers (test/test_siam_scPA.jl:28-31,60-63) computed end to end on the device."""
import numpy as np
import pytest

from helpers import anderson

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    s = max(np.max(np.abs(a)), np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / s)


def _random_local_pair(orc, strategy_seed=3):
    """local solver whose reference vertex is itself a local Vertex on a larger box over a RefVertex with core arrays"""
    import fddgasolver_jl_b200 as fd
    T, U, nG = 0.2, 1.3, 20
    core = fd.synthetic_local_vertex(T, U, numK1=5, numK2=(2, 2), numK3=(1, 1), core=(3, 2), seed=9).F0     # RefVertex with core arrays
    F0 = fd.NL2_Vertex(core, T, 14, (5, 6), (3, 3), 1)
    fd.randomize_vertex(F0, 21, 0.3)
    Gb = fd.siam_bare_Green(T, nG, e=0.3, Δ=0.7, D=8.0)
    G0 = 0.8 * Gb
    Σ0 = 0.05 * Gb
    S = fd.ParquetSolver(10, (4, 5), (3, 2), Gb, G0, Σ0, F0, T=T, mΠν_factor=3)
    fd.randomize_vertex(S.F, strategy_seed, 0.2); S.push("F"); S.init_sym_grp()
    R = orc.OracleLocalSolver(10, (4, 5), (3, 2), Gb, G0, Σ0, F0, T=T, mΠν_factor=3)
    R.init_sym_grp(); R.F.set(S.F)
    return S, R


@pytest.mark.parametrize("strategy", ["fdPA", "scPA"])
def test_local_iterate_matches_oracle(orc, strategy):
    import fddgasolver_jl_b200 as fd
    S, R = _random_local_pair(orc)
    S.pull("Π", "G")
    for n in ("Π0pp", "Π0ph", "Πpp", "Πph"):
        assert rel(getattr(S, n), getattr(R, n)) < TOL, n
    for _ in range(2):
        fd.iterate_solver(S, strategy, True); orc.iterate_solver_local(R, strategy, True)
    S.pull("F", "Σ", "G", "FL")
    assert rel(S.F.flatten(), R.F.flatten()) < TOL
    assert rel(S.FL.flatten(), R.FL.flatten()) < TOL
    assert rel(S.Σ, R.Σ) < TOL and rel(S.G, R.G) < TOL
    S.close()


GOLD = {
    0.0: dict(nK2=(6, 6),
              Σ=[-0.052138235296134906, -0.03838544776344314, 0.03838544776344314, 0.052138235296134906],
              γa=[0.13203850929270397, 0.5403615530152339, 0.2333246221064017, 0.09056300899983459],
              γp=[-0.10420799999591804, -0.2403951910434166, -0.15592452265704748, -0.07622568434721624],
              γt=[0.013898648018808482, 0.1499562726081748, 0.03867632419082161, 0.007160400841240708]),
    0.5: dict(nK2=(7, 6),
              Σ=[-0.0389123277075552 - 0.16855090184215607j, -0.025252640312580586 - 0.17429637478745583j,
                 0.025252640312580586 - 0.17429637478745583j, 0.0389123277075552 - 0.16855090184215607j],
              γa=[0.11925962005661812 + 8.57514999021054e-5j, 0.416232811242488 + 3.319936929625957e-5j,
                  0.20353141073439696 - 8.209974062547027e-5j, 0.08259294412067451 - 7.660306021755952e-5j],
              γp=[-0.12570450372739117 + 0.06583917638195431j, -0.24548654160724023 + 0.021014183409764874j,
                  -0.17578586892780296 - 0.05408344507941078j, -0.09544981624337806 - 0.06686768343644132j],
              γt=[0.011016969129875598 + 7.455601769977568e-5j, 0.09533547032272821 + 2.4477973720973318e-5j,
                  0.028843799251846686 - 6.260706942592786e-5j, 0.005828299701446311 - 7.32415771550901e-5j]),
}


@pytest.mark.parametrize("e", [0.0, 0.5])
def test_siam_scPA_golden_numbers_on_gpu(orc, e):
    """SIAM parquet solve (BASELINE configs[0] family) converged ON THE GPU through the C-ABI (fixed_point! + a host
    Anderson loop standing in for NLsolve) reproduces the reference's stored numbers (mΠν_factor = 1)."""
    import fddgasolver_jl_b200 as fd
    T, nmax = 0.1, 6
    nG, nK1 = 6 * nmax, 4 * nmax
    g = GOLD[e]
    S = fd.parquet_solver_siam_parquet_approximation(nG, nK1, g["nK2"], g["nK2"], e=e, Δ=np.pi / 5, D=10.0, T=T, U=1.0, mΠν_factor=1)
    S.init_sym_grp()
    nF = S.length_F()
    x0 = np.concatenate([S.F.flatten(), S.Σ.ravel(order="F")])

    def fp(x):
        R = np.empty_like(x)
        return fd.fixed_point(R, x, S, "scPA", True)
    x, it, err = anderson(fp, x0, tol=1e-10)
    assert err < 1e-10
    S.unflatten_F(x[:nF]); S.pull("F")
    Σ = x[nF:]
    assert np.max(np.abs(np.array([Σ[n + nG] for n in (-2, -1, 0, 1)]) - np.array(g["Σ"]))) < 1e-4
    xs = [-4 * np.pi * T + i for i in range(4)]
    for name, ch in (("γa", S.F.γa), ("γp", S.F.γp), ("γt", S.F.γt)):
        vals = [orc.interp_boson(ch.K1[:, 0], T, nK1, xx) for xx in xs]
        assert np.max(np.abs(np.array(vals) - np.array(g[name]))) < 1e-4, name
    S.close()


def test_siam_fdPA_golden_sigma_on_gpu(orc):
    """test/test_siam_fdPA.jl:72-84 on the device (single Hartree subtraction): fdPA from the converged
    reference (nmax = 12) to the target on a smaller vertex box (nmax = 8) reproduces the golden Σ(πT)."""
    import fddgasolver_jl_b200 as fd
    T, U, nmax = 0.1, 1.0, 12
    nG = nK1 = 8 * nmax

    def solve(S, strategy):
        nF = S.length_F()
        x0 = np.concatenate([S.flatten_F(), S.Σ.ravel(order="F")])
        x, it, err = anderson(lambda x: fd.fixed_point(np.empty_like(x), x, S, strategy, True), x0, tol=1e-10)
        assert err < 1e-10
        S.unflatten_F(x[:nF]); S.Σ[...] = x[nF:].reshape(S.Σ.shape, order="F"); S.push("Σ")
        fd.Dyson(S); S.pull("F", "G")
    S0 = fd.parquet_solver_siam_parquet_approximation(nG, nK1, (nmax, nmax), (nmax, nmax), e=-0.3, Δ=np.pi / 3, D=10.0, T=T, U=U, mΠν_factor=1)
    S0.init_sym_grp(); solve(S0, "scPA")
    Gb2 = fd.siam_bare_Green(T, nG, e=0.5, Δ=np.pi / 5, D=20.0)
    S2 = fd.ParquetSolver(64, (8, 8), (8, 8), Gb2, S0.G, S0.Σ, S0.F, T=T, mΠν_factor=1)
    S2.set_option("fd_hartree_once", 1)
    S2.init_sym_grp(); solve(S2, "fdPA")
    assert abs(S2.Σ[nG, 0] - (0.024643001835742997 - 0.17494219707558506j)) < 5e-5
    S0.close(); S2.close()


def test_local_hubbard_config2_sizes(orc):
    """BASELINE configs[1]: local Hubbard fdΓA iteration on the grid sizes of data/high_temperature_U2.0*.h5 (T = 0.5; G N = 128;
    K1 N = 128; K2 N = (74, 50); core (24, 16); mΠν_factor = 6 => Π 255 x 1536).  The device runs the full iteration; the oracle
    recomputes K1 / K3 / caches in full and sampled class representatives of L_K2 / K2 (inputs taken from the device state), plus
    the exact symmetry of every output."""
    import ctypes
    import fddgasolver_jl_b200 as fd
    T, U, nG = 0.5, 2.0, 128
    F0 = fd.synthetic_local_vertex(T, U, numK1=128, numK2=(74, 50), numK3=(1, 1), core=(24, 16), seed=2)
    for g in F0.channels():
        for a in g.arrays():
            a *= 0.05
    Gb = fd.siam_bare_Green(T, nG, e=0.1, Δ=0.6, D=10.0)
    S = fd.ParquetSolver(128, (74, 50), (24, 16), Gb, 0.9 * Gb, 0.02 * Gb, F0, T=T, mΠν_factor=6)
    fd.randomize_vertex(S.F, 3, 0.05); S.push("F"); S.init_sym_grp()
    fd.symmetrize_solver(S); S.pull("F")
    R = orc.OracleLocalSolver(128, (74, 50), (24, 16), Gb, 0.9 * Gb, 0.02 * Gb, F0, T=T, mΠν_factor=6)
    R.init_sym_grp(); R.F.set(S.F)
    S.pull("Π", "G")
    for n in ("Π0pp", "Π0ph", "Πpp", "Πph"):
        assert rel(getattr(S, n), getattr(R, n)) < TOL, n
    fd.iterate_solver(S, "fdPA", True)
    S.pull("FL", "Fbuff", "cache", "Σ")
    assert np.isfinite(S.Σ).all() and np.isfinite(S.Fbuff.flatten()).all()
    orc.build_K3_cache(R)
    for n in ("cache_Γpx", "cache_F0p", "cache_F0a", "cache_F0t", "cache_Γpp", "cache_Γa", "cache_Γt", "cache_Fp", "cache_Fa", "cache_Ft"):
        assert rel(getattr(S, n), getattr(R, n)) < TOL, n
    rng = np.random.default_rng(5)
    order = (fd.pCh, fd.aCh)
    for ch in order:                                  # L_K2 samples (inputs: S.F, S.F0)
        sg = S._sg[fd._lib.SG_PP2 if ch == fd.pCh else fd._lib.SG_PH2]
        for c in rng.integers(0, len(sg[0]) - 1, size=6):
            orc.BSE_L_K2_local(R, ch, c0=int(c), c1=int(c) + 1)
            idx = sg[1][sg[0][c]:sg[0][c + 1]]
            assert rel(S.FL.channel(ch).K2.ravel(order="F")[idx], R.FL.channel(ch).K2.ravel(order="F")[idx]) < TOL, ("L_K2", ch, c)
    R.FL.set(S.FL)
    for ch in (fd.pCh, fd.aCh, fd.tCh):
        orc.BSE_L_K3(R, ch)
    for a, b in zip(S.FL.channels(), R.FL.channels()):
        assert rel(a.K3, b.K3) < TOL
    for ch in (fd.pCh, fd.aCh, fd.tCh):
        orc.BSE_K1(R, ch)
    for ch in (fd.pCh, fd.aCh, fd.tCh):
        orc.BSE_K3(R, ch)
    for a, b in zip(S.Fbuff.channels(), R.Fbuff.channels()):
        assert rel(a.K1, b.K1) < TOL and rel(a.K3, b.K3) < TOL
    for ch in order:                                  # K2 samples
        sg = S._sg[fd._lib.SG_PP2 if ch == fd.pCh else fd._lib.SG_PH2]
        for c in rng.integers(0, len(sg[0]) - 1, size=6):
            R.Fbuff.channel(ch).K2[...] = 0
            orc.BSE_K2(R, ch, c0=int(c), c1=int(c) + 1)
            idx = sg[1][sg[0][c]:sg[0][c + 1]]
            assert rel(S.Fbuff.channel(ch).K2.ravel(order="F")[idx], R.Fbuff.channel(ch).K2.ravel(order="F")[idx]) < TOL, ("K2", ch, c)
    for ch in order:                                  # outputs are exactly class-constant
        for cls, which in (("K1", fd._lib.SG_K1), ("K2", fd._lib.SG_PP2 if ch == fd.pCh else fd._lib.SG_PH2),
                           ("K3", fd._lib.SG_PP3 if ch == fd.pCh else fd._lib.SG_PH3)):
            flat = getattr(S.Fbuff.channel(ch), cls).ravel(order="F").copy()
            sym = flat.copy()
            orc.lib().orc_symmetrize(orc._p(sym), ctypes.byref(orc.sg_struct(S._sg[which])))
            assert np.array_equal(sym, flat), (ch, cls)
    S.close()


def test_solve_mirror_reproduces_siam_golden_sigma(orc):
    """solve!(S; strategy = :scPA, tol) (src/solve.jl:160-196) as mirrored by fd.solve: the reference's own call sequence of
    test/test_siam_scPA.jl:20-31 end to end on the device."""
    import fddgasolver_jl_b200 as fd
    T, nmax = 0.1, 6
    nG, nK1 = 6 * nmax, 4 * nmax
    g = GOLD[0.0]
    S = fd.parquet_solver_siam_parquet_approximation(nG, nK1, g["nK2"], g["nK2"], e=0.0, Δ=np.pi / 5, D=10.0, T=T, U=1.0, mΠν_factor=1)
    S.init_sym_grp()
    res = fd.solve(S, strategy="scPA", tol=1e-9, maxiter=100)
    assert res.f_converged and res.iterations < 60
    S.pull("Σ")
    Σ = S.Σ.ravel(order="F")
    assert np.max(np.abs(np.array([Σ[n + nG] for n in (-2, -1, 0, 1)]) - np.array(g["Σ"]))) < 1e-4
    S.close()


def test_siam_scPA_real_element_type(orc):
    """test/test_siam_scPA.jl:22-36 with Q = Float64: the half-filled impurity (e = 0) has real i G, i Σ and real vertices; the
    real-typed solver goes through the same device arithmetic, every imaginary part is an exact zero, and the golden numbers are
    reproduced; a doped impurity (complex i G) is refused like Julia's InexactError"""
    import fddgasolver_jl_b200 as fd
    T, nmax = 0.1, 6
    nG, nK1 = 6 * nmax, 4 * nmax
    g = GOLD[0.0]
    S = fd.parquet_solver_siam_parquet_approximation(nG, nK1, g["nK2"], g["nK2"], np.float64, e=0.0, Δ=np.pi / 5, D=10.0, T=T, U=1.0, mΠν_factor=1)
    assert S.eltype == np.float64
    S.init_sym_grp()
    nF = S.length_F()
    x0 = np.concatenate([S.F.flatten(), S.Σ.ravel(order="F")]).real.copy()      # the solver's state is a REAL vector

    def fp(xr):
        x = xr.astype(np.complex128)
        R = fd.fixed_point(np.empty_like(x), x, S, "scPA", True)
        assert not np.any(R.imag)                         # the device never produces an imaginary part from real inputs
        return R.real.copy()
    x, it, err = anderson(fp, x0, tol=1e-10)
    assert err < 1e-10 and x.dtype == np.float64
    S.unflatten_F(x[:nF].astype(np.complex128)); S.pull("F", "Σ", "G")
    Σ = fd.real_array(S, S.Σ)
    assert Σ.dtype == np.float64 and fd.real_array(S, S.G).dtype == np.float64
    for ch in S.F.channels():
        for a in ch.arrays():
            assert fd.real_array(S, a).dtype == np.float64          # raises if any imaginary part is not an exact zero
    assert np.max(np.abs(np.array([Σ[n + nG, 0] for n in (-2, -1, 0, 1)]) - np.real(np.array(g["Σ"])))) < 1e-4
    xs = [-4 * np.pi * T + i for i in range(4)]
    for name, ch in (("γa", S.F.γa), ("γp", S.F.γp), ("γt", S.F.γt)):
        vals = [orc.interp_boson(ch.K1[:, 0], T, nK1, xx) for xx in xs]
        assert np.max(np.abs(np.array(vals) - np.array(g[name]))) < 1e-4, name
    S.close()
    with pytest.raises(fd.FdgaError, match="must be real"):
        fd.parquet_solver_siam_parquet_approximation(nG, nK1, g["nK2"], g["nK2"], np.float64, e=0.5, Δ=np.pi / 5, D=10.0, T=T, U=1.0)
