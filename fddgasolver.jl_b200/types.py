"""Host-side data model mirroring the reference's vertex structs (containers of numpy arrays).

Reference: src/types.jl (tags, array aliases), src/channel.jl:7-48 (Channel), src/nonlocal_2/channel.jl:4-41
(NL2_Channel), src/nonlocal/channel.jl:3-51 (NL_Channel), src/refvertex.jl:1-35 (RefVertex), src/vertex.jl:7-36 (Vertex),
src/nonlocal_2/vertex.jl:1-32 (NL2_Vertex), src/nonlocal/vertex.jl:1-37 (NL_Vertex).  All arrays are complex128 in Fortran (column-major) order, i.e. byte-identical to the
Julia ``Array{ComplexF64,N}`` the C-ABI receives.  Evaluation happens on the device (libfdga), not here.
"""
import numpy as np

# channel / spin tags (src/types.jl:16-30, 79-93); the integer values are the C-ABI's
pCh, tCh, aCh = 0, 1, 2
pSp, xSp, dSp = 0, 1, 2
CHANNELS = (pCh, tCh, aCh)
CH_NAME = {pCh: "γp", tCh: "γt", aCh: "γa"}


def nB(N):
    """length of MatsubaraMesh(T, N, Boson): indices -(N-1)..N-1"""
    return 2 * N - 1


def nF(N):
    """length of MatsubaraMesh(T, N, Fermion): indices -N..N-1"""
    return 2 * N


def zeros(shape):
    return np.zeros(shape, dtype=np.complex128, order="F")


class Channel:
    """local reducible vertex K1[Ω], K2[Ω,ν], K3[Ω,ν,ν'] (src/channel.jl:7-48)"""
    nonlocal_ = False

    def __init__(self, T, numK1, numK2, numK3):
        assert numK1 >= numK2[0] and numK1 >= numK2[1], "K1 mesh must contain the K2 meshes"
        assert numK2[0] >= numK3[0] and numK2[1] >= numK3[1], "K2 meshes must contain the K3 meshes"
        self.T, self.numK1, self.numK2, self.numK3 = float(T), int(numK1), tuple(numK2), tuple(numK3)
        self.K1 = zeros((nB(numK1),))
        self.K2 = zeros((nB(numK2[0]), nF(numK2[1])))
        self.K3 = zeros((nB(numK3[0]), nF(numK3[1]), nF(numK3[1])))

    def arrays(self):
        return (self.K1, self.K2, self.K3)

    def __len__(self):
        return self.K1.size + self.K2.size + self.K3.size

    def flatten(self):
        # [K1; K2; K3], each in column-major order (src/channel.jl:155-176)
        return np.concatenate([a.ravel(order="F") for a in self.arrays()])

    def unflatten(self, x):
        off = 0
        for a in self.arrays():
            a[...] = np.asarray(x[off:off + a.size]).reshape(a.shape, order="F")
            off += a.size
        assert off == len(x)

    def set(self, other):
        for a, b in zip(self.arrays(), other.arrays() if not np.isscalar(other) else (other,) * 3):
            a[...] = b

    def copy(self):
        import copy as _c
        return _c.deepcopy(self)


class NL2_Channel(Channel):
    """K1[Ω,P], K2[Ω,ν,P,k], K3[Ω,ν,ν',P] on an L x L momentum mesh (src/nonlocal_2/channel.jl:4-41)"""
    nonlocal_ = True

    def __init__(self, T, numK1, numK2, numK3, L):
        assert numK1 > numK2[0] and numK1 > numK2[1], "K1 mesh must be strictly larger than the K2 meshes"
        assert numK2[0] >= numK3[0] and numK2[1] >= numK3[1], "K2 meshes must contain the K3 meshes"
        self.T, self.numK1, self.numK2, self.numK3, self.L = float(T), int(numK1), tuple(numK2), tuple(numK3), int(L)
        NP = L * L
        self.K1 = zeros((nB(numK1), NP))
        self.K2 = zeros((nB(numK2[0]), nF(numK2[1]), NP, NP))
        self.K3 = zeros((nB(numK3[0]), nF(numK3[1]), nF(numK3[1]), NP))


class NL_Channel(Channel):
    """K1[Ω,P], K2[Ω,ν,P], K3[Ω,ν,ν',P]: bosonic momentum dependence only (src/nonlocal/channel.jl:3-51)"""
    nonlocal_ = True

    def __init__(self, T, numK1, numK2, numK3, L):
        assert numK1 >= numK2[0] and numK1 >= numK2[1], "K1 mesh must contain the K2 meshes"
        assert numK2[0] >= numK3[0] and numK2[1] >= numK3[1], "K2 meshes must contain the K3 meshes"
        self.T, self.numK1, self.numK2, self.numK3, self.L = float(T), int(numK1), tuple(numK2), tuple(numK3), int(L)
        NP = L * L
        self.K1 = zeros((nB(numK1), NP))
        self.K2 = zeros((nB(numK2[0]), nF(numK2[1]), NP))
        self.K3 = zeros((nB(numK3[0]), nF(numK3[1]), nF(numK3[1]), NP))


class RefVertex:
    """bare U + the four core arrays Fp_p, Fp_x, Ft_p, Ft_x (src/refvertex.jl:1-35)"""

    def __init__(self, T, U, numK3=None, Fp_p=None, Fp_x=None, Ft_p=None, Ft_x=None):
        self.T, self.U = float(T), complex(U)
        if numK3 is None:
            numK3 = (1, 1)         # "null vertices" of RefVertex(T, U), src/refvertex.jl:20-35
        self.numK3 = tuple(numK3)
        shp = (nB(numK3[0]), nF(numK3[1]), nF(numK3[1]))
        self.Fp_p = zeros(shp) if Fp_p is None else np.asfortranarray(Fp_p, dtype=np.complex128)
        self.Fp_x = zeros(shp) if Fp_x is None else np.asfortranarray(Fp_x, dtype=np.complex128)
        self.Ft_p = zeros(shp) if Ft_p is None else np.asfortranarray(Ft_p, dtype=np.complex128)
        self.Ft_x = zeros(shp) if Ft_x is None else np.asfortranarray(Ft_x, dtype=np.complex128)
        for a in self.arrays():
            assert a.shape == shp

    def arrays(self):
        return (self.Fp_p, self.Fp_x, self.Ft_p, self.Ft_x)

    def copy(self):
        import copy as _c
        return _c.deepcopy(self)


class _VertexBase:
    def channels(self):
        return (self.γp, self.γt, self.γa)

    def channel(self, ch):
        return self.channels()[ch]

    @property
    def T(self):
        return self.γp.T

    @property
    def numK1(self):
        return self.γp.numK1

    @property
    def numK2(self):
        return self.γp.numK2

    @property
    def numK3(self):
        return self.γp.numK3

    def __len__(self):
        return 3 * len(self.γp)

    def flatten(self):
        # [γp; γt; γa] (src/vertex.jl:153-167)
        return np.concatenate([g.flatten() for g in self.channels()])

    def unflatten(self, x):
        n = len(self.γp)
        for i, g in enumerate(self.channels()):
            g.unflatten(x[i * n:(i + 1) * n])

    def set(self, other):
        # set!(F1, F2) copies the three reducible vertices, not F0 (src/vertex.jl:80-101)
        for i, g in enumerate(self.channels()):
            g.set(other if np.isscalar(other) else other.channels()[i])

    def add(self, other):
        for g, h in zip(self.channels(), other.channels()):
            for a, b in zip(g.arrays(), h.arrays()):
                a += b

    def bare_vertex(self):
        F0 = self.F0
        while not isinstance(F0, RefVertex):
            F0 = F0.F0
        return F0.U

    def copy(self):
        import copy as _c
        return _c.deepcopy(self)


class Vertex(_VertexBase):
    """local vertex F0 + γp, γt, γa (src/vertex.jl:7-36)"""

    def __init__(self, F0, T, numK1, numK2, numK3):
        self.F0 = F0
        self.γp = Channel(T, numK1, numK2, numK3)
        self.γt = Channel(T, numK1, numK2, numK3)
        self.γa = Channel(T, numK1, numK2, numK3)


class NL2_Vertex(_VertexBase):
    """nonlocal vertex with K2(k) momentum dependence (src/nonlocal_2/vertex.jl:1-32)"""

    def __init__(self, F0, T, numK1, numK2, numK3, L):
        self.F0 = F0
        self.L = int(L)
        self.γp = NL2_Channel(T, numK1, numK2, numK3, L)
        self.γt = NL2_Channel(T, numK1, numK2, numK3, L)
        self.γa = NL2_Channel(T, numK1, numK2, numK3, L)


class NL_Vertex(_VertexBase):
    """nonlocal vertex with bosonic momentum dependence only, the s-wave solver's vertex (src/nonlocal/vertex.jl:1-37)"""

    def __init__(self, F0, T, numK1, numK2, numK3, L):
        self.F0 = F0
        self.L = int(L)
        self.γp = NL_Channel(T, numK1, numK2, numK3, L)
        self.γt = NL_Channel(T, numK1, numK2, numK3, L)
        self.γa = NL_Channel(T, numK1, numK2, numK3, L)


class MBEVertex(Vertex):
    """local multi-boson-exchange vertex (src/boson_exchange.jl:237-266): the arrays of a Vertex; K1 = screened interaction - U,
    K2 = Hedin-vertex part, K3 = multi-boson part M.  Evaluated on the device as U + K1 + K2 + K2' + K2 K2' / (U + K1) + K3."""
    mbe = True


class NL2_MBEVertex(NL2_Vertex):
    """nonlocal multi-boson-exchange vertex (src/boson_exchange.jl:738-770)"""
    mbe = True


def vertex_chain(F):
    """[F, F.F0, F.F0.F0, ..., RefVertex]"""
    out = [F]
    while not isinstance(out[-1], RefVertex):
        out.append(out[-1].F0)
    return out
