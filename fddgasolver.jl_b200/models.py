"""Model functions used to build inputs (src/models/hubbard.jl:8-44). Elementwise host code (input generation)."""
import numpy as np


def hubbard_band(k1, k2, t1, t2=0.0, t3=0.0):
    ek = -2 * t1 * (np.cos(k1) + np.cos(k2))
    ek = ek + -4 * t2 * np.cos(k1) * np.cos(k2)
    ek = ek + -2 * t3 * (np.cos(2 * k1) + np.cos(2 * k2))
    return ek


def hubbard_bare_Green(T, nG, LG, *, μ, t1, t2=0.0, t3=0.0):
    """i * G0(ν, k) on MatsubaraMesh(T, nG, Fermion) x BrillouinZoneMesh(LG): array (2 nG, LG*LG), x fastest."""
    n = np.arange(-nG, nG)
    nu = (2 * n + 1) * np.pi * T
    ix, iy = np.meshgrid(np.arange(LG), np.arange(LG), indexing="ij")      # [ix, iy]
    ek = hubbard_band(2 * np.pi * ix / LG, 2 * np.pi * iy / LG, t1, t2, t3)  # [ix, iy]
    ek = ek.reshape(LG * LG, order="F")                                   # linear index ix + LG * iy
    G = 1.0 / (1j * nu[:, None] + μ - ek[None, :]) * 1j
    return np.asfortranarray(G, dtype=np.complex128)


def siam_bare_Green(T, nG, *, e, Δ, D):
    """i * G0(ν) of the single-impurity Anderson model (src/models/siam.jl:9-28): array of length 2 nG."""
    nu = (2 * np.arange(-nG, nG) + 1) * np.pi * T
    if np.isinf(D):
        return 1.0 / (nu + 1j * e + Δ * np.sign(nu))
    return 1.0 / (nu + 1j * e + 2 * Δ / np.pi * np.arctan(D / nu))
