"""Size-independent properties at BASELINE's full size (config 3: nmax = 4, nq = 8, LG = 48; length(S.F) = 780 096) where the CPU
oracle would need minutes per kernel: linearity of the mfRG map, the DQGMRES residual bound, algebraic identities between the
BSE variants, idempotence of the symmetrisation, Fourier interpolation round trip."""
import numpy as np
import pytest

from test_gpu_parity import rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S3():
    import fddgasolver_jl_b200 as fd
    S = fd.wu_point_solver(nmax=4, nq=8, LG=48, F0_scale=0.02)
    yield S
    S.close()


def test_mfrg_map_is_real_affine(S3):
    """mfRGLinearMap (src/mfRG.jl:34-89) as coded is a real-AFFINE map: A(a x + b y) = a A x + b A y + (1 - a - b) A(0) for real
    a, b.  Not complex-linear because the symmetry-class fill conjugates some members; and A(0) != 0 in the K3 sector when S.F0
    carries nonlocal K's, because build_K3_cache_mfRG! takes S.F(...; γr = false) - S.F.F0(...; γr = false) and the inner F0 call
    does not forward the switch (src/nonlocal_2/build_K3_cache.jl:108-126 with src/nonlocal/vertex.jl:87-89; DESIGN.md "E3").
    The CPU oracle shows the same behaviour; K1 and K2 sectors are exactly linear."""
    import fddgasolver_jl_b200 as fd
    rng = np.random.default_rng(0)
    n = S3.length_F()
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    a, b = 0.7, -1.3
    nK3 = S3.F.γp.K3.size
    per = n // 3
    for strategy in ("fdPA", "fdPA_new"):
        A = fd.mfRGLinearMap(S3, strategy)
        A0 = A.matvec(np.zeros(n, dtype=complex))
        lhs = A.matvec(a * x + b * y)
        rhs = a * A.matvec(x) + b * A.matvec(y) + (1 - a - b) * A0
        assert rel(lhs, rhs) < 1e-11, strategy
        for c in range(3):                                  # the constant lives in the K3 blocks only
            assert not np.any(A0[c * per: (c + 1) * per - nK3]), (strategy, c)


def test_dqgmres_true_residual_obeys_the_quasi_residual_bound():
    """Saad & Wu, Prop. 4.1: ||b - A x_m|| <= sqrt(m + 1) |gamma_{m+1}| for the truncated method, and the estimate tracks the true
    residual while the orthogonalisation is complete.  Needs a genuinely linear operator: S.F0 without nonlocal K's (else the map
    is affine, finding E3) and a right-hand side in the symmetric subspace (the class fill is only real-linear)."""
    import fddgasolver_jl_b200 as fd
    S = fd.wu_point_solver(nmax=4, nq=8, LG=48, F0_scale=0.0)
    rng = np.random.default_rng(1)
    n = S.length_F()
    S.unflatten_F(rng.standard_normal(n) + 1j * rng.standard_normal(n))
    fd.symmetrize_solver(S)
    b = S.flatten_F()
    A = fd.mfRGLinearMap(S)
    assert not np.any(A.matvec(np.zeros(n, dtype=complex)))
    for memory, itmax in ((30, 12), (4, 12)):
        x, st = fd.dqgmres(A, b, memory=memory, atol=0.0, rtol=0.0, itmax=itmax)
        assert st["niter"] == itmax and len(st["residuals"]) == itmax + 1
        true = np.linalg.norm(A.matvec(x) - b)
        est = st["residuals"][-1]
        assert est < st["residuals"][0]
        assert true <= np.sqrt(itmax + 1) * est * (1 + 1e-6), (memory, true, est)
        if memory >= itmax:
            assert abs(true - est) <= 0.1 * est, (true, est)
    S.close()


def test_variant_identities_at_full_size(S3):
    """FL = 0  =>  BSE_K{1,2}_1loop!(fd) == BSE_K{1,2}!(fd): the optimised column path on the RK_1L right factor against the
    same path on RK_FD (tests/test_oracle_variants.py proves the identity for the oracle)."""
    import fddgasolver_jl_b200 as fd
    S = S3
    S.unflatten_F(S.F.flatten())
    for g in S.FL.channels():
        for a in g.arrays():
            a[...] = 0
    S.push("FL")
    order = (fd.pCh, fd.aCh, fd.tCh)
    for ch in order:
        fd.BSE_K1(S, ch)
    for ch in order:
        fd.BSE_K2(S, ch)
    S.pull("Fbuff")
    full = [a.copy() for g in S.Fbuff.channels() for a in (g.K1, g.K2)]
    for ch in order:
        fd.BSE_K1_1loop(S, ch)
    for ch in order:
        fd.BSE_K2_1loop(S, ch)
    S.pull("Fbuff")
    one = [a for g in S.Fbuff.channels() for a in (g.K1, g.K2)]
    assert max(float(np.max(np.abs(a))) for a in full) > 1e-6
    for a, b in zip(full, one):
        assert rel(a, b) < 1e-12


def test_symmetrize_is_idempotent_and_iterate_preserves_symmetry(S3):
    import fddgasolver_jl_b200 as fd
    S = S3
    rng = np.random.default_rng(2)
    n = S.length_F()
    S.unflatten_F(rng.standard_normal(n) + 1j * rng.standard_normal(n))
    fd.symmetrize_solver(S)
    x1 = S.flatten_F()
    fd.symmetrize_solver(S)
    assert np.array_equal(S.flatten_F(), x1)
    fd.iterate_solver(S, "fdPA", False)
    y = S.flatten_F()
    fd.symmetrize_solver(S)
    # K1 and the p, a channels are class-constant by construction; γt = (γt^d + γa) / 2 mixes the ph groups only up to rounding
    assert rel(S.flatten_F(), y) < 1e-13


def test_interpolation_round_trip_at_full_size(S3):
    """refine 8 x 8 -> 16 x 16 and coarsen back: the identity (every Fourier coefficient of the coarse mesh survives, the split
    Nyquist components recombine)"""
    import fddgasolver_jl_b200 as fd
    T, U = S3.T, 5.6
    Fi = fd.NL2_Vertex(fd.RefVertex(T, U), T, 4, (2, 2), (2, 2), 8)
    fd.randomize_vertex(Fi, 12, 1.0)
    Gb = fd.hubbard_bare_Green(T, 6, 16, μ=0.1, t1=1.0)
    Sf = fd.NL2_ParquetSolver(4, (2, 2), (2, 2), 16, Gb, Gb, np.zeros_like(Gb), fd.RefVertex(T, U), T=T)
    fd.interpolate_vertex(Sf, Fi)
    Sf.pull("F")
    # coincident points of the fine mesh carry the coarse values
    k1 = Sf.F.γp.K1.reshape(7, 16, 16, order="F")[:, ::2, ::2].reshape(7, 64, order="F")
    assert rel(k1, Fi.γp.K1) < 1e-12
    Gc = fd.hubbard_bare_Green(T, 6, 8, μ=0.1, t1=1.0)
    Sc = fd.NL2_ParquetSolver(4, (2, 2), (2, 2), 8, Gc, Gc, np.zeros_like(Gc), fd.RefVertex(T, U), T=T)
    fd.interpolate_vertex(Sc, Sf.F)
    Sc.pull("F")
    assert rel(Sc.F.flatten(), Fi.flatten()) < 1e-11
    Sf.close(); Sc.close()
