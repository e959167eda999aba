"""Solver constructors and seeded synthetic inputs.

There is no HDF5 reader in this environment, so the DMFT input files of the reference (data/*.h5) are
replaced by synthetic vertices of the same grid sizes: K1 N=128, K2 N=(64,48), dummy K3 N=(1,1),
core box N=(24,16) (sizes decoded from data/Wu_point*.h5).
"""
import os

import numpy as np

from .models import hubbard_bare_Green, siam_bare_Green
from .solver import NL2_ParquetSolver, NL_ParquetSolver, ParquetSolver
from .types import NL2_MBEVertex, NL2_Vertex, NL_Vertex, RefVertex, Vertex, nB, nF


def parquet_solver_hubbard_parquet_approximation_NL2(nG, nK1, nK2, nK3, LG, L, *, T, U, μ, t1, t2=0.0, t3=0.0,
                                                     mode="threads", device=0):
    """Parquet approximation: G0 = Σ0 = 0, F0 = U (src/nonlocal_2/ParquetSolver.jl:167-196)."""
    Gbare = hubbard_bare_Green(T, nG, LG, μ=μ, t1=t1, t2=t2, t3=t3)
    G0 = np.zeros_like(Gbare)
    Σ0 = np.zeros_like(Gbare)
    F0 = RefVertex(T, U)
    return NL2_ParquetSolver(nK1, nK2, nK3, L, Gbare, G0, Σ0, F0, T=T, mode=mode, device=device)


def parquet_solver_hubbard_parquet_approximation(nG, nK1, nK2, nK3, LG, L, *, T, U, μ, t1, t2=0.0, t3=0.0, mode="threads", device=0):
    """Parquet approximation with the s-wave solver: G0 = Σ0 = 0, F0 = U (src/nonlocal/ParquetSolver.jl:167-194)."""
    Gbare = hubbard_bare_Green(T, nG, LG, μ=μ, t1=t1, t2=t2, t3=t3)
    return NL_ParquetSolver(nK1, nK2, nK3, L, Gbare, np.zeros_like(Gbare), np.zeros_like(Gbare), RefVertex(T, U), T=T, mode=mode, device=device)


def parquet_solver_siam_parquet_approximation(nG, nK1, nK2, nK3, Q=np.complex128, *, e, Δ, D, T, U, mode="threads", mΠν_factor=6, device=0, VT=None):
    """Parquet approximation for the SIAM: G0 = Σ0 = 0, F0 = U (src/ParquetSolver.jl:170-200).  Q = np.float64: the reference's
    real-typed solver (only meaningful at e = 0, where i G is real)."""
    Gbare = siam_bare_Green(T, nG, e=e, Δ=Δ, D=D)
    if np.dtype(Q) == np.float64:
        Gbare = Gbare.real.astype(np.complex128) if np.all(Gbare.imag == 0) else Gbare      # a complex i G is refused by the solver
    z = np.zeros_like(Gbare)
    return ParquetSolver(nK1, nK2, nK3, Gbare, z, z, RefVertex(T, U), T=T, mode=mode, mΠν_factor=mΠν_factor, device=device, Q=Q, VT=VT)


def _decay_b(N):
    m = np.arange(-(N - 1), N)
    return 1.0 / (1.0 + m.astype(float) ** 2)


def _decay_f(N):
    n = np.arange(-N, N)
    return 1.0 / (1.0 + (n + 0.5) ** 2)


def _rand(rng, shape):
    return np.asfortranarray(rng.random(shape) - 0.5 + 1j * (rng.random(shape) - 0.5))


def randomize_vertex(V, seed, scale=1.0):
    """Fill K1, K2, K3 of a Vertex / NL2_Vertex with seeded, frequency-decaying random numbers."""
    rng = np.random.default_rng(seed)
    for g in V.channels():
        db1, db2, df2 = _decay_b(g.numK1), _decay_b(g.numK2[0]), _decay_f(g.numK2[1])
        db3, df3 = _decay_b(g.numK3[0]), _decay_f(g.numK3[1])
        ex1 = (slice(None),) + (None,) * (g.K1.ndim - 1)
        g.K1[...] = scale * _rand(rng, g.K1.shape) * db1[ex1]
        ex = (None,) * (g.K2.ndim - 2)
        g.K2[...] = scale * _rand(rng, g.K2.shape) * db2[(slice(None), None) + ex] * df2[(None, slice(None)) + ex]
        ex = (None,) * (g.K3.ndim - 3)
        g.K3[...] = scale * _rand(rng, g.K3.shape) * (db3[(slice(None), None, None) + ex]
                                                       * df3[(None, slice(None), None) + ex]
                                                       * df3[(None, None, slice(None)) + ex])
    return V


def synthetic_local_vertex(T, U, numK1=128, numK2=(64, 48), numK3=(1, 1), core=(24, 16), seed=2):
    """Stand-in for load_triqs_data(...).Γ: local Vertex{RefVertex} with the DMFT-file grid sizes."""
    rng = np.random.default_rng(seed)
    shp = (nB(core[0]), nF(core[1]), nF(core[1]))
    dec = (_decay_b(core[0])[:, None, None] * _decay_f(core[1])[None, :, None] * _decay_f(core[1])[None, None, :])
    arrs = [U * U * _rand(rng, shp) * dec for _ in range(4)]
    ref = RefVertex(T, U, core, *arrs)
    V = Vertex(ref, T, numK1, numK2, numK3)
    randomize_vertex(V, seed + 1000, scale=U * U)
    for g in V.channels():
        if numK3 == (1, 1):
            g.K3[...] = 0.0        # dummy K3 of the packaged DMFT data
    return V


def synthetic_local_sigma(T, nG, U, n=0.48, seed=3):
    """i*Σ(iν) of an impurity-like self-energy: Hartree-free tail U² n(1-n)/(iν) regularised at low ν."""
    nu = (2 * np.arange(-nG, nG) + 1) * np.pi * T
    a = U * U * n * (1 - n)
    sig = a / (1j * nu + 1.5 * np.sign(nu) * 1j)
    return 1j * sig


DMFT_FIXTURE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "wu_point_dmft.npz")


def load_dmft_fixture(path=DMFT_FIXTURE):
    """The reference's packaged DMFT data of the Wu point (data/Wu_point.h5 as load_triqs_data returns it, src/utility/
    load_triqs.jl:298-308), from the fixture written by tests/golden/make_wu_point_fixture.py: dict G, G0, Σ (i G, i Σ on the
    fermionic mesh of size nG), Γ (local Vertex with its RefVertex core), occ, params, T."""
    z = np.load(path)
    T, U = float(z["T"]), complex(z["U"])
    core = RefVertex(T, U, tuple(int(x) for x in z["core_N"]), z["Fp_p"], z["Fp_x"], z["Ft_p"], z["Ft_x"])
    Γ = Vertex(core, T, int(z["numK1"]), tuple(int(x) for x in z["numK2"]), tuple(int(x) for x in z["numK3"]))
    for ch, g in zip("pta", Γ.channels()):
        g.K1[...], g.K2[...], g.K3[...] = z[f"K1_{ch}"], z[f"K2_{ch}"], z[f"K3_{ch}"]
    return {"G": z["G"], "G0": z["G0"], "Σ": z["Sigma"], "nG": int(z["nG"]), "T": T, "Γ": Γ, "occ": float(z["occ"]),
            "params": dict(zip([str(k) for k in z["param_names"]], [float(v) for v in z["param_values"]])), "source": str(z["source"])}


def wu_point_inputs(nmax=4, nq=8, LG=48, *, seed=1, F_scale=1e-2, F0_scale=0.0, small_reference=False, data="auto", nl_method=2):
    """Pure-numpy inputs of the Wu-point NL2 problem (no device needed): dict with the constructor arguments
    of NL2_ParquetSolver plus the seeded start vertex `F` (an NL2_Vertex whose F0 is the reference vertex).

    Mirrors script/benchmark_Wu.jl:7-44: T=0.2, U=5.6, μ=2.1800201007694464-U/2, t1=1, t2=-0.3,
    nG = nK1 = 4 nmax, nK2 = nK3 = (nmax, nmax), G mesh LG x LG, vertex mesh nq x nq,
    F0 = NL2_Vertex(Γ_local), S.F seeded random * F_scale.  F0_scale != 0 also fills S.F0's own K's
    (state after the first outer iteration).  nl_method = 1: the s-wave solver's inputs (script/run_Wu_point.jl:88-90),
    F0 = NL_Vertex(Γ_local), F an NL_Vertex.
    """
    T, U = 0.2, 5.6
    μ = 2.1800201007694464 - U / 2
    t1, t2 = 1.0, -0.3
    nG = nK1 = 4 * nmax
    nK2 = nK3 = (nmax, nmax)
    Gbare = hubbard_bare_Green(T, nG, LG, μ=μ, t1=t1, t2=t2)
    use_ref = (data == "reference") or (data == "auto" and not small_reference and os.path.exists(DMFT_FIXTURE))
    if use_ref:
        d = load_dmft_fixture()
        assert abs(d["T"] - T) < 1e-12 and nG <= d["nG"], "the packaged data is on the T = 0.2 mesh with N = 128"
        sl = slice(d["nG"] - nG, d["nG"] + nG)             # same temperature: the mesh points coincide, data_triqs.G(value(ν)) is exact
        G0 = np.asfortranarray(np.repeat(d["G"][sl][:, None], LG * LG, axis=1))
        Σ0 = np.asfortranarray(np.repeat(d["Σ"][sl][:, None], LG * LG, axis=1))
        Γ = d["Γ"]
    else:
        Σloc = synthetic_local_sigma(T, nG, U)
        Σ0 = np.asfortranarray(np.repeat(Σloc[:, None], LG * LG, axis=1))
        Glat = 1.0 / (1.0 / Gbare + Σ0)
        G0 = np.asfortranarray(np.repeat(Glat.mean(axis=1)[:, None], LG * LG, axis=1))   # impurity G = local lattice G
        if small_reference:
            Γ = synthetic_local_vertex(T, U, numK1=2 * nK1, numK2=(2 * nmax, 2 * nmax), numK3=(1, 1), core=(nmax + 1, nmax), seed=2)
        else:
            Γ = synthetic_local_vertex(T, U, seed=2)
    VT = {1: NL_Vertex, 2: NL2_Vertex, -2: NL2_MBEVertex}[nl_method]
    if nl_method == -2:      # script/run_Wu_point.jl:95: F0 = NL2_MBEVertex(asymptotic_to_mbe(data_triqs.Γ), ...); the conversion runs on the device
        from .mbe import asymptotic_to_mbe
        Γ = asymptotic_to_mbe(Γ)
    F0 = VT(Γ, T, nK1, nK2, nK3, nq)
    if F0_scale:
        randomize_vertex(F0, seed + 7, scale=F0_scale)
    F = VT(F0, T, nK1, nK2, nK3, nq)
    randomize_vertex(F, seed, scale=F_scale)
    return dict(T=T, U=U, nK1=nK1, nK2=nK2, nK3=nK3, L=nq, Gbare=Gbare, G0=G0, Σ0=Σ0, F0=F0, F=F,
                data=("reference file data/Wu_point.h5 (local DMFT vertex, impurity G and Σ, via tests/golden/wu_point_dmft.npz) + seeded nonlocal start vertex"
                      if use_ref else "synthetic"))


def wu_point_solver(nmax=4, nq=8, LG=48, *, seed=1, device=0, init_sym=True, F_scale=1e-2, F0_scale=0.0,
                    small_reference=False, data="auto", nl_method=2):
    """NL2_ParquetSolver (nl_method = 2) or NL_ParquetSolver (nl_method = 1) on the GPU for wu_point_inputs(...) (see there)."""
    inp = wu_point_inputs(nmax, nq, LG, seed=seed, F_scale=F_scale, F0_scale=F0_scale, small_reference=small_reference, data=data, nl_method=nl_method)
    kw = dict(VT=NL2_MBEVertex) if nl_method == -2 else {}
    S = {1: NL_ParquetSolver, 2: NL2_ParquetSolver, -2: NL2_ParquetSolver}[nl_method](inp["nK1"], inp["nK2"], inp["nK3"], inp["L"], inp["Gbare"], inp["G0"], inp["Σ0"], inp["F0"],
                          T=inp["T"], device=device, **kw)
    S.F.set(inp["F"])
    S.push("F")
    if init_sym:
        S.init_sym_grp()
    return S
