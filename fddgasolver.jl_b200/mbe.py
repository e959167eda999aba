"""Conversions between the asymptotic and the multi-boson-exchange parametrisation of a vertex.

  asymptotic_to_mbe(F)     src/boson_exchange.jl:672-710 (Vertex), :794-813 (NL2_Vertex)
  mbe_to_asymptotic(F)     src/boson_exchange.jl:713-735 (MBEVertex), :816-837 (NL2_MBEVertex)

The reference does this on the host by calling the vertices at every grid point.  Here the calls go to the device evaluators
(fdga_eval_vertex through a small helper context that holds the vertex as its reference chain), the array arithmetic is numpy.
Used once per calculation, e.g. F0 = NL2_MBEVertex(asymptotic_to_mbe(data_triqs.Γ), ...) of script/run_Wu_point.jl:95.
"""
import numpy as np

from . import _lib as L
from .types import (MBEVertex, NL2_MBEVertex, NL2_Vertex, RefVertex, Vertex, aCh, pCh, pSp, tCh, xSp)


class _Evaluator:
    """the vertex V as a callable on the device: a helper solver whose reference chain S.F0 is V (S.F itself stays zero)"""

    def __init__(self, V, device=0):
        from .solver import NL2_ParquetSolver
        self.V = V
        nl2 = isinstance(V, NL2_Vertex)
        self.L = V.L if nl2 else 1
        nK2, nK3 = V.numK2, V.numK3
        nK1 = max(V.numK1, nK2[0] + 1, nK2[1] + 1)
        z = np.zeros((2, self.L * self.L), dtype=np.complex128, order="F")
        VT = NL2_MBEVertex if getattr(V, "mbe", False) else None
        self.S = NL2_ParquetSolver(nK1, nK2, nK3, self.L, z, z, z, V, T=V.T, compute_bubbles=False, device=device, VT=VT)

    def refresh(self):
        self.S.push("F0")

    def __call__(self, W, v, w, Ch, Sp, P=0, **kw):
        return self.S.eval_vertex(W, v, w, Ch, Sp, P, 0, 0, level=1, swave=isinstance(self.V, NL2_Vertex), **kw)

    def close(self):
        self.S.close()


def _k3_grid(g, nonlocal_):
    """index arrays (Ω, ν, ω[, P]) of a K3 array, shaped for broadcasting against it"""
    nb, nf = g.numK3
    W = np.arange(-(nb - 1), nb)
    f = np.arange(-nf, nf)
    if nonlocal_:
        P = np.arange(g.K3.shape[3])
        return W[:, None, None, None], f[None, :, None, None], f[None, None, :, None], P[None, None, None, :]
    return W[:, None, None], f[None, :, None], f[None, None, :], 0


def _asymptotic_part(g, nonlocal_):
    """K1(Ω[, P]) + K2(Ω, ν[, P, kSW]) + K2(Ω, ω[, P, kSW]) on the K3 grid of the channel g (the K3 box lies inside the others)"""
    nb, nf = g.numK3
    o1, o2b, o2f = g.numK1 - nb, g.numK2[0] - nb, g.numK2[1] - nf
    sb1, sb2, sf2 = slice(o1, o1 + 2 * nb - 1), slice(o2b, o2b + 2 * nb - 1), slice(o2f, o2f + 2 * nf)
    if nonlocal_:
        K2sw = g.K2.mean(axis=3)                                   # γ.K2[Ω, ν, P, kSW]
        K1 = g.K1[sb1][:, None, None, :]
        return K1 + K2sw[sb2, sf2][:, :, None, :] + K2sw[sb2, sf2][:, None, :, :]
    K1 = g.K1[sb1].reshape(-1)[:, None, None]
    K2 = g.K2[sb2, sf2].reshape(2 * nb - 1, 2 * nf)
    return K1 + K2[:, :, None] + K2[:, None, :]


def _own(ch):
    return dict(γp=ch == pCh, γt=ch == tCh, γa=ch == aCh)


def asymptotic_to_mbe(F, device=0):
    """MBE vertex with the same full vertex as the asymptotic vertex F (for NL2 vertices: the same s-wave component).
    F: Vertex or NL2_Vertex over a RefVertex (or a deeper chain); returns MBEVertex / NL2_MBEVertex (new arrays)."""
    nonlocal_ = isinstance(F, NL2_Vertex)
    if not isinstance(F, (Vertex, NL2_Vertex)) or getattr(F, "mbe", False):
        raise L.FdgaError("asymptotic_to_mbe: needs an asymptotic Vertex or NL2_Vertex")
    Fm = (NL2_MBEVertex(F.F0.copy(), F.T, F.numK1, F.numK2, F.numK3, F.L) if nonlocal_
          else MBEVertex(F.F0.copy(), F.T, F.numK1, F.numK2, F.numK3))
    Fm.set(F)
    U = Fm.bare_vertex()
    core = Fm.F0
    if nonlocal_ or not isinstance(core, RefVertex) or tuple(F.γa.numK3) >= tuple(core.numK3):
        # subtract the SBE contribution from K3, channel by channel IN THIS ORDER (a, p, t): the t channel is assembled from the a
        # classes, which are already converted when it is processed (src/boson_exchange.jl:680-692, 800-811)
        ev = _Evaluator(Fm, device)
        try:
            for ch in (aCh, pCh, tCh):
                g = Fm.channel(ch)
                W, v, w, P = _k3_grid(g, nonlocal_)
                nabla = ev(W, v, w, ch, pSp, P, **_own(ch)) - g.K3
                g.K3 -= nabla - (U + _asymptotic_part(g, nonlocal_))
                ev.refresh()
        finally:
            ev.close()
        return Fm
    # the reducible part has (almost) no K3: subtract the SBE contribution from the RefVertex core instead, in all channels (:694-707)
    nb, nf = core.numK3
    W, v, w = np.arange(-(nb - 1), nb)[:, None, None], np.arange(-nf, nf)[None, :, None], np.arange(-nf, nf)[None, None, :]
    em, ea = _Evaluator(Fm, device), _Evaluator(F, device)
    try:
        for name, ch, sp in (("Fp_p", pCh, pSp), ("Fp_x", pCh, xSp), ("Ft_p", tCh, pSp), ("Ft_x", tCh, xSp)):
            nabla = em(W, v, w, ch, sp, F0=False) - ea(W, v, w, ch, sp, F0=False)
            getattr(core, name)[...] -= nabla          # the F0 = false evaluations do not depend on the core: no refresh needed
    finally:
        em.close(); ea.close()
    return Fm


def mbe_to_asymptotic(Fm, device=0):
    """asymptotic vertex whose K3 class absorbs the SBE term of the MBE vertex Fm (inverse of the K3 branch of asymptotic_to_mbe)"""
    nonlocal_ = isinstance(Fm, NL2_MBEVertex)
    if not getattr(Fm, "mbe", False):
        raise L.FdgaError("mbe_to_asymptotic: needs an MBEVertex or NL2_MBEVertex")
    F = (NL2_Vertex(Fm.F0.copy(), Fm.T, Fm.numK1, Fm.numK2, Fm.numK3, Fm.L) if nonlocal_
         else Vertex(Fm.F0.copy(), Fm.T, Fm.numK1, Fm.numK2, Fm.numK3))
    F.set(Fm)
    U = Fm.bare_vertex()
    ev = _Evaluator(Fm, device)
    try:
        for ch in (aCh, pCh, tCh):
            g = F.channel(ch)
            W, v, w, P = _k3_grid(g, nonlocal_)
            g.K3[...] = ev(W, v, w, ch, pSp, P, **_own(ch)) - (U + _asymptotic_part(g, nonlocal_))
    finally:
        ev.close()
    return F
