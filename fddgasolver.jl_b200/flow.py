"""Ω-flow of the bare Green function between the impurity and the lattice problem (src/flow.jl:1-61; Eqs. (20), (23), (24) of
Phys. Rev. Research 4, 013034 (2022)).  Host-side preprocessing of a run (one complex root per Matsubara frequency), like the
reference's.

    G0_Λ(ν, k) = Θ_Λ(ν) G0_lat(ν, k) + Ξ_Λ(ν) G0_imp(ν),      Θ_Λ(ν) = ν² / (ν² + Λ²)

with Ξ_Λ(ν) fixed by the DMFT self-consistency  mean_k [ G0_Λ(ν, k)⁻¹ + Σ_imp(ν) ]⁻¹ = G_imp(ν)  (stored quantities are i G and
i Σ, so Dyson reads G = 1 / (1 / G0 + Σ), src/dyson.jl:20-31).  Λ = ∞ gives the impurity function, Λ = 0 the lattice one.
"""
import numpy as np


def bare_Green_Ω_flow(Λ, G0_imp, Σ_imp, G0_lat, T, *, ftol=1e-12, maxiter=100):
    """bare_Green_Ω_flow(Λ, G0_imp, Σ_imp, G0_lat) => G0_Λ.

    G0_lat: [2 nG, Nk] lattice bare Green function on the fermionic mesh of size nG at temperature T; G0_imp, Σ_imp: impurity
    functions on a fermionic mesh of the SAME temperature and size >= nG (the packaged data has N = 128), taken at the coinciding
    Matsubara frequencies as the reference does with `G0_imp[ν]`.  The root is found by a complex Newton iteration from Ξ = 0 (the
    map is holomorphic in Ξ; the reference calls NLsolve on (Re Ξ, Im Ξ) from the same start with the same ftol)."""
    G0_lat = np.asarray(G0_lat, dtype=np.complex128)
    nG = G0_lat.shape[0] // 2
    ni = len(G0_imp) // 2
    assert len(G0_imp) == len(Σ_imp) and ni >= nG, "impurity mesh must contain the lattice mesh"
    sl = slice(ni - nG, ni + nG)
    g0i, si = np.asarray(G0_imp, dtype=np.complex128)[sl], np.asarray(Σ_imp, dtype=np.complex128)[sl]
    g_imp = 1.0 / (1.0 / g0i + si)                                   # Dyson!(G_imp, Σ_imp, G0_imp)
    ν = (2 * np.arange(-nG, nG) + 1) * np.pi * T
    Θ = ν ** 2 / (ν ** 2 + Λ ** 2) if np.isfinite(Λ) else np.zeros_like(ν)
    out = np.zeros_like(G0_lat, order="F")
    for i in range(2 * nG):
        Ξ = 0.0 + 0.0j
        for _ in range(maxiter):
            g = Θ[i] * G0_lat[i, :] + Ξ * g0i[i]
            d = 1.0 + si[i] * g
            f = np.mean(g / d) - g_imp[i]
            if max(abs(f.real), abs(f.imag)) <= ftol:
                break
            Ξ -= f / (g0i[i] * np.mean(1.0 / d ** 2))
        else:
            raise RuntimeError(f"bare_Green_Ω_flow: no convergence at frequency index {i}")
        out[i, :] = Θ[i] * G0_lat[i, :] + Ξ * g0i[i]
    # sanity check of the DMFT self-consistency (src/flow.jl:46-56)
    G_Λ = 1.0 / (1.0 / out + si[:, None])
    assert np.max(np.abs(G_Λ.mean(axis=1) - g_imp)) < 1e-10
    return out
