"""The oracle's OWN data model (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Deliberately independent of fddgasolver.jl_b200/types.py: mesh lengths, array shapes and the flatten order are restated here
from the reference, so that a layout mistake in the product's containers shows up as a parity failure instead of being
shared by both sides.

  mesh lengths    MatsubaraMesh(T, N, Boson) has the indices -(N-1)..N-1, MatsubaraMesh(T, N, Fermion) -N..N-1
                  (test/test_channel.jl:16, script/check_triqs.jl:55-57)
  array shapes    src/types.jl:105-129: K1[Ω(,P)], K2[Ω,ν(,P,k)], K3[Ω,ν,ν'(,P)]
  flatten order   per channel [K1; K2; K3] (src/channel.jl:155-176), per vertex [γp; γt; γa] (src/vertex.jl:153-167),
                  every array in Julia's column-major element order
Foreign containers (anything with the same attribute names) are accepted wherever a vertex is read.
"""
import copy as _copy

import numpy as np

pCh, tCh, aCh = 0, 1, 2
pSp, xSp, dSp = 0, 1, 2


def n_boson(N):
    return len(range(-(N - 1), N))


def n_fermion(N):
    return len(range(-N, N))


# names used by oracle.py
nB, nF = n_boson, n_fermion


def zeros(shape):
    return np.zeros(shape, dtype=np.complex128, order="F")


def _colmajor(a):
    """elements of `a` in Julia's vec(a) order: first index fastest (= row-major order of the reversed-axes view)"""
    return np.ascontiguousarray(np.transpose(a)).reshape(-1)


def _fill_colmajor(a, x):
    np.transpose(a)[...] = np.asarray(x).reshape(a.shape[::-1])


class OChannel:
    def __init__(self, T, numK1, numK2, numK3, L=None, swave=False):
        # swave: NL_Channel of src/nonlocal/channel.jl:3-51, K2[Ω, ν, P] (bosonic momentum only)
        self.T, self.numK1, self.numK2, self.numK3 = float(T), int(numK1), (int(numK2[0]), int(numK2[1])), (int(numK3[0]), int(numK3[1]))
        self.L = None if L is None else int(L)
        self.swave = bool(swave)
        mom1 = () if L is None else (self.L ** 2,)
        mom2 = () if L is None else ((self.L ** 2,) if swave else (self.L ** 2, self.L ** 2))
        self.K1 = zeros((n_boson(self.numK1),) + mom1)
        self.K2 = zeros((n_boson(self.numK2[0]), n_fermion(self.numK2[1])) + mom2)
        self.K3 = zeros((n_boson(self.numK3[0]), n_fermion(self.numK3[1]), n_fermion(self.numK3[1])) + mom1)

    @property
    def nonlocal_(self):
        return self.L is not None

    def arrays(self):
        return (self.K1, self.K2, self.K3)

    def __len__(self):
        return sum(a.size for a in self.arrays())

    def flatten(self):
        return np.concatenate([_colmajor(self.K1), _colmajor(self.K2), _colmajor(self.K3)])

    def unflatten(self, x):
        x = np.asarray(x)
        assert x.size == len(self)
        o = 0
        for a in (self.K1, self.K2, self.K3):
            _fill_colmajor(a, x[o:o + a.size])
            o += a.size

    def set(self, other):
        if np.isscalar(other):
            for a in self.arrays():
                a[...] = other
            return
        for name in ("K1", "K2", "K3"):
            src = getattr(other, name)
            dst = getattr(self, name)
            assert dst.shape == src.shape, (name, dst.shape, src.shape)
            dst[...] = src

    def copy(self):
        return _copy.deepcopy(self)


class ORefVertex:
    def __init__(self, T, U, numK3=None, Fp_p=None, Fp_x=None, Ft_p=None, Ft_x=None):
        self.T, self.U = float(T), complex(U)
        self.numK3 = (1, 1) if numK3 is None else (int(numK3[0]), int(numK3[1]))     # null vertices, src/refvertex.jl:20-35
        shp = (n_boson(self.numK3[0]), n_fermion(self.numK3[1]), n_fermion(self.numK3[1]))
        for name, a in (("Fp_p", Fp_p), ("Fp_x", Fp_x), ("Ft_p", Ft_p), ("Ft_x", Ft_x)):
            arr = zeros(shp) if a is None else np.asfortranarray(np.array(a, dtype=np.complex128))
            assert arr.shape == shp, (name, arr.shape, shp)
            setattr(self, name, arr)

    def arrays(self):
        return (self.Fp_p, self.Fp_x, self.Ft_p, self.Ft_x)

    def copy(self):
        return _copy.deepcopy(self)


class _OVertexBase:
    ORDER = ("γp", "γt", "γa")      # flatten order of src/vertex.jl:153-167; also the integer channel tags 0, 1, 2

    def channels(self):
        return tuple(getattr(self, n) for n in self.ORDER)

    def channel(self, ch):
        return getattr(self, self.ORDER[ch])

    T = property(lambda self: self.γp.T)
    numK1 = property(lambda self: self.γp.numK1)
    numK2 = property(lambda self: self.γp.numK2)
    numK3 = property(lambda self: self.γp.numK3)

    def __len__(self):
        return sum(len(g) for g in self.channels())

    def flatten(self):
        return np.concatenate([getattr(self, n).flatten() for n in self.ORDER])

    def unflatten(self, x):
        x = np.asarray(x)
        o = 0
        for n in self.ORDER:
            g = getattr(self, n)
            g.unflatten(x[o:o + len(g)])
            o += len(g)
        assert o == x.size

    def set(self, other):
        # set!(F1, F2) copies the three reducible vertices only, not F0 (src/vertex.jl:80-101)
        for n in self.ORDER:
            getattr(self, n).set(other if np.isscalar(other) else getattr(other, n))

    def add(self, other):
        for n in self.ORDER:
            for a, b in zip(getattr(self, n).arrays(), getattr(other, n).arrays()):
                a += b

    def bare_vertex(self):
        v = self.F0
        while not isinstance(v, ORefVertex):
            v = v.F0
        return v.U

    def copy(self):
        return _copy.deepcopy(self)


class OVertex(_OVertexBase):
    def __init__(self, F0, T, numK1, numK2, numK3):
        self.F0 = F0
        for n in self.ORDER:
            setattr(self, n, OChannel(T, numK1, numK2, numK3))


class ONL2_Vertex(_OVertexBase):
    def __init__(self, F0, T, numK1, numK2, numK3, L):
        self.F0 = F0
        self.L = int(L)
        for n in self.ORDER:
            setattr(self, n, OChannel(T, numK1, numK2, numK3, L))


class ONL_Vertex(_OVertexBase):
    """NL_Vertex (src/nonlocal/vertex.jl:1-37): bosonic momentum dependence only"""
    def __init__(self, F0, T, numK1, numK2, numK3, L):
        self.F0 = F0
        self.L = int(L)
        for n in self.ORDER:
            setattr(self, n, OChannel(T, numK1, numK2, numK3, L, swave=True))


class OMBEVertex(OVertex):
    """MBEVertex (src/boson_exchange.jl:237-266): the arrays of a Vertex, evaluated as screened interaction / Hedin vertices /
    multi-boson part"""


class ONL2_MBEVertex(ONL2_Vertex):
    """NL2_MBEVertex (src/boson_exchange.jl:738-770)"""


# the names oracle.py uses
RefVertex, Vertex, NL2_Vertex, NL_Vertex, MBEVertex, NL2_MBEVertex = ORefVertex, OVertex, ONL2_Vertex, ONL_Vertex, OMBEVertex, ONL2_MBEVertex


def adopt(V):
    """deep copy of a vertex chain held in foreign containers (same attribute names) into the oracle's own types"""
    if isinstance(V, (ORefVertex, OVertex, ONL2_Vertex, ONL_Vertex)):
        return V
    if hasattr(V, "Fp_p"):
        return ORefVertex(V.T, V.U, V.numK3, V.Fp_p, V.Fp_x, V.Ft_p, V.Ft_x)
    F0 = adopt(V.F0)
    mbe = bool(getattr(V, "mbe", False)) or type(V).__name__.endswith("MBEVertex")
    if getattr(V, "L", None) is not None and V.γp.K1.ndim == 2:
        cls = ONL_Vertex if V.γp.K2.ndim == 3 else (ONL2_MBEVertex if mbe else ONL2_Vertex)
        out = cls(F0, V.T, V.numK1, V.numK2, V.numK3, V.L)
    else:
        out = (OMBEVertex if mbe else OVertex)(F0, V.T, V.numK1, V.numK2, V.numK3)
    out.set(V)
    return out


def vertex_chain(F):
    out = [F]
    while not isinstance(out[-1], ORefVertex):
        out.append(out[-1].F0)
    return out
