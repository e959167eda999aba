"""Python adapter of the CPU oracle (oracle/fdga_oracle.cpp).  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
It restates the reference's host-level call structure (BSE_templates.jl wrappers, iterate_solver!, SDE!,
mfRGLinearMap) on top of the C++ restatement of the kernels, operating on plain numpy arrays.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
# the oracle keeps its OWN data model (mesh lengths, shapes, flatten order restated from the reference): oracle/otypes.py.
# Vertices handed in by tests in the product's containers are copied into it (otypes.adopt), never used in place.
from otypes import (MBEVertex, NL2_MBEVertex, NL2_Vertex, NL_Vertex, RefVertex, Vertex, aCh, adopt, dSp, nB, nF, pCh, pSp, tCh,  # noqa: E402,F401
                    vertex_chain, xSp, zeros)

LIB = os.path.join(_HERE, "_build", "libfdga_oracle.so")
INF = (2 ** 31 - 1) // 4


def build(force=False):
    src = os.path.join(_HERE, "fdga_oracle.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-fopenmp", "-shared", "-fPIC", "-std=c++17", "-o", LIB, src])
    return LIB


class _Level(C.Structure):
    _fields_ = [("type", C.c_int), ("nK1", C.c_int), ("nK2b", C.c_int), ("nK2f", C.c_int), ("nK3b", C.c_int), ("nK3f", C.c_int),
                ("U_re", C.c_double), ("U_im", C.c_double),
                ("K1", C.c_void_p * 3), ("K2", C.c_void_p * 3), ("K3", C.c_void_p * 3), ("core", C.c_void_p * 4)]


class _Vertex(C.Structure):
    _fields_ = [("nlev", C.c_int), ("lev", _Level * 8)]


class _Grid(C.Structure):
    _fields_ = [("T", C.c_double), ("L", C.c_int), ("nPiB", C.c_int), ("nPiF", C.c_int), ("swave", C.c_int)]


class _SG(C.Structure):
    _fields_ = [("nclasses", C.c_int64), ("offsets", C.c_void_p), ("index", C.c_void_p), ("op", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_occupation.restype = C.c_double
        _lib.orc_build_symmetry_group.restype = C.c_int64
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    assert a.flags["F_CONTIGUOUS"] or a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def vertex_struct(V):
    """ctypes descriptor of the nested vertex chain V -> V.F0 -> ... (arrays are referenced, not copied)"""
    out = _Vertex()
    chain = vertex_chain(adopt(V))
    out.nlev = len(chain)
    for i, X in enumerate(chain):
        lv = out.lev[i]
        if isinstance(X, RefVertex):
            lv.type = 2
            lv.nK3b, lv.nK3f = X.numK3
            lv.U_re, lv.U_im = X.U.real, X.U.imag
            for j, a in enumerate(X.arrays()):
                lv.core[j] = a.ctypes.data
        else:
            lv.type = (4 if isinstance(X, NL2_MBEVertex) else 0) if isinstance(X, NL2_Vertex) else (3 if isinstance(X, NL_Vertex) else (5 if isinstance(X, MBEVertex) else 1))
            lv.nK1 = X.numK1
            lv.nK2b, lv.nK2f = X.numK2
            lv.nK3b, lv.nK3f = X.numK3
            for r, g in enumerate(X.channels()):
                lv.K1[r], lv.K2[r], lv.K3[r] = g.K1.ctypes.data, g.K2.ctypes.data, g.K3.ctypes.data
    out._keep = chain
    return out


def sg_struct(tbl):
    offsets, index, ops = tbl
    s = _SG()
    s.nclasses = len(offsets) - 1
    s.offsets, s.index, s.op = offsets.ctypes.data, index.ctypes.data, ops.ctypes.data
    s._keep = tbl
    return s


def build_symmetry_group(kind, n0, n1, L, length):
    offsets = np.zeros(length + 1, dtype=np.int64)
    index = np.zeros(length, dtype=np.int64)
    ops = np.zeros(length, dtype=np.uint8)
    n = lib().orc_build_symmetry_group(kind, n0, n1, L, _p(offsets), _p(index), _p(ops))
    return offsets[: n + 1].copy(), index, ops


def trivial_group(length):
    return (np.arange(length + 1, dtype=np.int64), np.arange(length, dtype=np.int64), np.zeros(length, dtype=np.uint8))


def eval_vertex(V, L, W, v, w, P, k, q, Ch, Sp, F0=True, γp=True, γt=True, γa=True, level=0):
    """V(Ω, ν, ω, P, k, q, Ch, Sp; F0, γp, γt, γa); k / q may be the string 'sw'; ν / ω may be INF."""
    vs = vertex_struct(V)
    out = np.zeros(1, dtype=np.complex128)
    ksw, qsw = isinstance(k, str), isinstance(q, str)
    Pa = (C.c_int * 2)(*P)
    ka = (C.c_int * 2)(*(k if not ksw else (0, 0)))
    qa = (C.c_int * 2)(*(q if not qsw else (0, 0)))
    lib().orc_eval_vertex(C.byref(vs), L, level, W, v, w, Pa, ka, qa, int(ksw), int(qsw), Ch, Sp,
                          int(F0), int(γp), int(γt), int(γa), _p(out))
    return complex(out[0])


K1Cl, K2Cl, K2pCl, K3Cl, ΛCl = range(5)


def eval_class(V, L, W, v, w, P, k, q, Ch, Cl, level=0):
    """V(Ω, ν, ω, P, k, q, Ch, Cl): one asymptotic class summed down the chain (src/boson_exchange.jl:270-330)"""
    vs = vertex_struct(V)
    out = np.zeros(1, dtype=np.complex128)
    ksw, qsw = isinstance(k, str), isinstance(q, str)
    Pa = (C.c_int * 2)(*P)
    ka = (C.c_int * 2)(*(k if not ksw else (0, 0)))
    qa = (C.c_int * 2)(*(q if not qsw else (0, 0)))
    lib().orc_eval_class(C.byref(vs), L, level, W, v, w, Pa, ka, qa, int(ksw), int(qsw), Ch, Cl, _p(out))
    return complex(out[0])


def eval_channel(V, L, r, W, v, w, P, k, q, K1=True, K2=True, K3=True, level=0):
    vs = vertex_struct(V)
    out = np.zeros(1, dtype=np.complex128)
    sw = (1 if isinstance(P, str) else 0) | (2 if isinstance(k, str) else 0) | (4 if isinstance(q, str) else 0)
    arr = [(C.c_int * 2)(*(x if not isinstance(x, str) else (0, 0))) for x in (P, k, q)]
    lib().orc_eval_channel(C.byref(vs), L, level, r, W, v, w, arr[0], arr[1], arr[2], sw, int(K1), int(K2), int(K3), _p(out))
    return complex(out[0])


_CACHE_NAMES = ["cache_Γpx", "cache_F0p", "cache_F0a", "cache_F0t", "cache_Γpp", "cache_Γa", "cache_Γt",
                "cache_Fp", "cache_Fa", "cache_Ft"]
SG_SIGMA, SG_K1, SG_PP2, SG_PH2, SG_PP3, SG_PH3, SG_PPL3, SG_PHL3 = range(8)
SG_NL_PP2, SG_NL_PH2 = 8, 9          # builder kinds of the s-wave solver's K2[Ω, ν, P] groups (stored under SG_PP2 / SG_PH2)


class OracleSolver:
    """CPU restatement of NL2_ParquetSolver (src/nonlocal_2/ParquetSolver.jl:1-154)."""

    def __init__(self, nK1, nK2, nK3, L_, Gbare, G0, Σ0, F0, *, T, mΠν_factor=1, compute_bubbles=True, VT=NL2_Vertex):
        # VT = NL2_MBEVertex: the solver's own vertex S.F is a multi-boson-exchange vertex (src/nonlocal_2/ParquetSolver.jl:86,113);
        # Fbuff and FL stay asymptotic NL2_Vertex'es (:114)
        self.T, self.L, self.NP = float(T), int(L_), int(L_) ** 2
        self.nK1, self.nK2, self.nK3 = int(nK1), tuple(nK2), tuple(nK3)
        self.Gbare = np.asfortranarray(Gbare, dtype=np.complex128)
        self.nG = self.Gbare.shape[0] // 2
        self.LG = int(round(np.sqrt(self.Gbare.shape[1])))
        self.nΠB, self.nΠF = self.nK1, self.nK1 * int(mΠν_factor)
        self.G0 = np.array(G0, dtype=np.complex128, order="F")
        self.Σ0 = np.array(Σ0, dtype=np.complex128, order="F")
        self.G = self.G0.copy(order="F")
        self.Σ = self.Σ0.copy(order="F")
        self.F0 = adopt(F0)
        self.F = VT(self.F0, self.T, nK1, nK2, nK3, self.L)
        self.Fbuff = NL2_Vertex(RefVertex(self.T, 0.0), self.T, nK1, nK2, nK3, self.L)
        self.FL = NL2_Vertex(RefVertex(self.T, 0.0), self.T, nK1, nK2, nK3, self.L)
        shpΠ = (nB(self.nΠB), nF(self.nΠF), self.NP, self.NP)
        self.Π0pp, self.Π0ph, self.Πpp, self.Πph = zeros(shpΠ), zeros(shpΠ), zeros(shpΠ), zeros(shpΠ)
        self.Lpp, self.Lph = zeros(self.F.γp.K2.shape), zeros(self.F.γp.K2.shape)
        for n in _CACHE_NAMES:
            setattr(self, n, zeros(self.F.γp.K3.shape))
        self.grid = _Grid(self.T, self.L, self.nΠB, self.nΠF, 0)
        self._finish_init(compute_bubbles)

    swave = False

    def _finish_init(self, compute_bubbles):
        self.reset_sym_grp()
        if compute_bubbles:
            bubbles_real_space(self, self.Π0pp, self.Π0ph, self.G0)
            Dyson(self)
            bubbles_real_space(self, self.Πpp, self.Πph, self.G)

    def _sg_len(self, which):
        return {SG_SIGMA: self.Σ.size, SG_K1: self.F.γp.K1.size, SG_PP2: self.F.γp.K2.size, SG_PH2: self.F.γp.K2.size}.get(which, self.F.γp.K3.size)

    def reset_sym_grp(self):
        self.sg = {w: trivial_group(self._sg_len(w)) for w in range(8)}

    def init_sym_grp(self):
        n = {SG_SIGMA: (self.nG, 0), SG_K1: (self.nK1, 0), SG_PP2: self.nK2, SG_PH2: self.nK2,
             SG_PP3: self.nK3, SG_PH3: self.nK3, SG_PPL3: self.nK3, SG_PHL3: self.nK3}
        kind = {SG_PP2: SG_NL_PP2, SG_PH2: SG_NL_PH2} if self.swave else {}
        for w, (n0, n1) in n.items():
            self.sg[w] = build_symmetry_group(kind.get(w, w), n0, n1, self.LG if w == SG_SIGMA else self.L, self._sg_len(w))

    def set_symmetry_classes(self, which, offsets, index, ops):
        self.sg[which] = (np.ascontiguousarray(offsets, dtype=np.int64), np.ascontiguousarray(index, dtype=np.int64),
                          np.ascontiguousarray(ops, dtype=np.uint8))

    def caches(self):
        return [getattr(self, n) for n in _CACHE_NAMES]


class OracleNLSolver(OracleSolver):
    """CPU restatement of NL_ParquetSolver (src/nonlocal/ParquetSolver.jl:1-154): the s-wave solver, vertices with bosonic
    momentum dependence only (K2[Ω, ν, P]), bubbles Π[Ω, ν, P] (mΠν_factor = 32 by default, :86)."""
    swave = True

    def __init__(self, nK1, nK2, nK3, L_, Gbare, G0, Σ0, F0, *, T, mΠν_factor=32, compute_bubbles=True):
        self.T, self.L, self.NP = float(T), int(L_), int(L_) ** 2
        self.nK1, self.nK2, self.nK3 = int(nK1), tuple(nK2), tuple(nK3)
        self.Gbare = np.asfortranarray(Gbare, dtype=np.complex128)
        self.nG = self.Gbare.shape[0] // 2
        self.LG = int(round(np.sqrt(self.Gbare.shape[1])))
        self.nΠB, self.nΠF = self.nK1, self.nK1 * int(mΠν_factor)
        self.G0 = np.array(G0, dtype=np.complex128, order="F")
        self.Σ0 = np.array(Σ0, dtype=np.complex128, order="F")
        self.G = self.G0.copy(order="F")
        self.Σ = self.Σ0.copy(order="F")
        self.F0 = adopt(F0)
        self.F = NL_Vertex(self.F0, self.T, nK1, nK2, nK3, self.L)
        self.Fbuff = NL_Vertex(RefVertex(self.T, 0.0), self.T, nK1, nK2, nK3, self.L)
        self.FL = NL_Vertex(RefVertex(self.T, 0.0), self.T, nK1, nK2, nK3, self.L)
        shpΠ = (nB(self.nΠB), nF(self.nΠF), self.NP)
        self.Π0pp, self.Π0ph, self.Πpp, self.Πph = zeros(shpΠ), zeros(shpΠ), zeros(shpΠ), zeros(shpΠ)
        self.Lpp, self.Lph = zeros(self.F.γp.K2.shape), zeros(self.F.γp.K2.shape)
        for n in _CACHE_NAMES:
            setattr(self, n, zeros(self.F.γp.K3.shape))
        self.grid = _Grid(self.T, self.L, self.nΠB, self.nΠF, 1)
        self._finish_init(compute_bubbles)


# ------------------------------------------------------------------------------- reference-named operations
def Dyson(S):
    lib().orc_dyson(_p(S.G), _p(S.Σ), _p(S.Gbare), C.c_int64(S.G.size))


def compute_occupation(S, G=None):
    G = S.G if G is None else G
    return lib().orc_occupation(_p(G), S.nG, S.LG, C.c_double(S.T))


def hubbard_bare_Green(T, nG, LG, *, μ, t1, t2=0.0, t3=0.0):
    G = zeros((2 * nG, LG * LG))
    lib().orc_hubbard_bare_green(_p(G), nG, LG, C.c_double(T), C.c_double(μ), C.c_double(t1), C.c_double(t2), C.c_double(t3))
    return G


def compute_hubbard_chemical_potential(occ_target, S, hubbard_params):
    """compute_hubbard_chemical_potential(occ_target, Σ, hubbard_params): src/dyson.jl:45-57; Roots.find_zero on a bracket =
    bisection down to neighbouring floating-point numbers."""
    t1, t2, t3 = hubbard_params["t1"], hubbard_params.get("t2", 0.0), hubbard_params.get("t3", 0.0)

    class Tmp:
        pass
    tmp = Tmp()
    tmp.Σ, tmp.G = S.Σ, np.zeros_like(S.Σ)

    def f(μ):
        tmp.Gbare = hubbard_bare_Green(S.T, S.nG, S.LG, μ=μ, t1=t1, t2=t2, t3=t3)
        Dyson(tmp)
        return compute_occupation(S, tmp.G) - occ_target
    a, b = -4 * abs(t1), 4 * abs(t1)
    fa, fb = f(a), f(b)
    if fa == 0:
        return a
    if fb == 0:
        return b
    if (fa < 0) == (fb < 0):
        raise ValueError("The interval [a,b] is not a bracketing interval")
    for _ in range(200):
        m = a + 0.5 * (b - a)
        if m <= a or m >= b:
            break
        fm = f(m)
        if fm == 0:
            return m
        if (fm < 0) == (fa < 0):
            a, fa = m, fm
        else:
            b, fb = m, fm
    return a if abs(fa) <= abs(fb) else b


def bubbles_real_space(S, Πpp, Πph, G, use_G_tail=True):
    if S.swave:     # src/nonlocal/bubble.jl:87-158
        lib().orc_nl_bubbles_real_space(_p(Πpp), _p(Πph), _p(G), S.nG, S.LG, C.byref(S.grid), int(use_G_tail))
        return
    lib().orc_bubbles_real_space(_p(Πpp), _p(Πph), _p(G), S.nG, S.LG, C.byref(S.grid))


def bubbles_momentum_space(S, Πpp, Πph, G):
    assert not S.swave, "the s-wave solver's bubbles! is bubbles_real_space! (src/nonlocal/ParquetSolver.jl:294-297)"
    lib().orc_bubbles_momentum_space(_p(Πpp), _p(Πph), _p(G), S.nG, S.LG, C.byref(S.grid))


def bubbles(S):
    bubbles_real_space(S, S.Πpp, S.Πph, S.G)


def _pi(S, ch, reference):
    if ch == pCh:
        return S.Π0pp if reference else S.Πpp
    return S.Π0ph if reference else S.Πph


def _sign_sp(ch):
    # (sign, Sp) of BSE_templates.jl:17,25,33
    return (-1, dSp) if ch == tCh else (+1, pSp)


def _tfix(Xt, Xa):
    # γt = (γt^d + γa) / 2, BSE_templates.jl:35-38
    Xt += Xa
    Xt /= 2


def build_K3_cache(S, i0=0, i1=-1):
    cs = S.caches()
    if i0 == 0 and i1 < 0:
        for c in cs:
            c[...] = 0
    arr = (C.c_void_p * 10)(*[c.ctypes.data for c in cs])
    getattr(lib(), "orc_build_K3_cache_mbe" if isinstance(S.F, NL2_MBEVertex) else "orc_build_K3_cache")(arr, S.nK3[0], S.nK3[1], C.byref(vertex_struct(S.F0)), C.byref(vertex_struct(S.F)),
                             C.byref(S.grid), C.c_int64(i0), C.c_int64(i1))


def build_K3_cache_mfRG(S, is_first_iteration):
    cs = S.caches()
    arr = (C.c_void_p * 10)(*[c.ctypes.data for c in cs])
    lib().orc_build_K3_cache_mfRG(arr, S.nK3[0], S.nK3[1], C.byref(vertex_struct(S.F0)), C.byref(vertex_struct(S.F)),
                                  C.byref(sg_struct(S.sg[SG_PP3])), C.byref(sg_struct(S.sg[SG_PH3])), int(is_first_iteration), C.byref(S.grid))


def BSE_K1(S, ch, is_mfRG=False, c0=0, c1=-1):
    sign, Sp = _sign_sp(ch)
    K1 = S.Fbuff.channel(ch).K1
    getattr(lib(), "orc_nl_bse_K1" if S.swave else "orc_bse_K1")(_p(K1), S.nK1, C.byref(vertex_struct(S.F0)), C.byref(vertex_struct(S.F)), C.byref(vertex_struct(S.FL)),
                     _p(_pi(S, ch, True)), _p(_pi(S, ch, False)), C.byref(sg_struct(S.sg[SG_K1])), sign, ch, Sp, int(is_mfRG),
                     C.byref(S.grid), C.c_int64(c0), C.c_int64(c1))
    if ch == tCh:
        _tfix(S.Fbuff.γt.K1, S.Fbuff.γa.K1)


def BSE_L_K2(S, ch, is_mfRG=False, c0=0, c1=-1):
    sign, Sp = _sign_sp(ch)
    K2 = S.FL.channel(ch).K2
    sg = S.sg[SG_PP2 if ch == pCh else SG_PH2]
    getattr(lib(), "orc_nl_bse_L_K2" if S.swave else "orc_bse_L_K2")(_p(K2), S.nK2[0], S.nK2[1], C.byref(vertex_struct(S.F0)), C.byref(vertex_struct(S.F)),
                       _p(_pi(S, ch, True)), C.byref(sg_struct(sg)), sign, ch, Sp, C.byref(S.grid), C.c_int64(c0), C.c_int64(c1))
    if ch == tCh:
        _tfix(S.FL.γt.K2, S.FL.γa.K2)


def BSE_K2(S, ch, is_mfRG=False, c0=0, c1=-1):
    sign, Sp = _sign_sp(ch)
    K2 = S.Fbuff.channel(ch).K2
    sg = S.sg[SG_PP2 if ch == pCh else SG_PH2]
    getattr(lib(), "orc_nl_bse_K2" if S.swave else "orc_bse_K2")(_p(K2), S.nK2[0], S.nK2[1], C.byref(vertex_struct(S.F0)), C.byref(vertex_struct(S.F)), C.byref(vertex_struct(S.FL)),
                     _p(_pi(S, ch, True)), _p(_pi(S, ch, False)), C.byref(sg_struct(sg)), sign, ch, Sp, int(is_mfRG),
                     C.byref(S.grid), C.c_int64(c0), C.c_int64(c1))
    if ch == tCh:       # BSEa_K2.jl:130-135
        K2 += 2 * S.FL.γt.K2
        K2 += -1 * S.FL.γa.K2
        _tfix(S.Fbuff.γt.K2, S.Fbuff.γa.K2)
    else:
        K2 += S.FL.channel(ch).K2


def BSE_L_K3(S, ch, is_mfRG=False):
    # (cache_Γ, cache_F0, sign): BSE_templates.jl:122,130,138
    cG, cF0, sign = {aCh: (S.cache_Γa, S.cache_F0a, +1), pCh: (S.cache_Γpp, S.cache_F0p, -1), tCh: (S.cache_Γt, S.cache_F0t, -1)}[ch]
    sg = S.sg[SG_PPL3 if ch == pCh else SG_PHL3]
    K3 = S.FL.channel(ch).K3
    lib().orc_bse_L_K3(_p(K3), S.nK3[0], S.nK3[1], _p(cG), _p(cF0), _p(_pi(S, ch, True)), C.byref(sg_struct(sg)), sign, C.byref(S.grid))
    if ch == tCh:
        _tfix(S.FL.γt.K3, S.FL.γa.K3)


def BSE_K3(S, ch, is_mfRG=False):
    # BSE_templates.jl:156,164,172
    cG, cF, cF0, s1, s2 = {aCh: (S.cache_Γa, S.cache_Fa, S.cache_F0a, +1, +1), pCh: (S.cache_Γpx, S.cache_Fp, S.cache_F0p, -1, +1),
                           tCh: (S.cache_Γt, S.cache_Ft, S.cache_F0t, -1, -1)}[ch]
    sg = S.sg[SG_PP3 if ch == pCh else SG_PH3]
    K3 = S.Fbuff.channel(ch).K3
    lib().orc_bse_K3(_p(K3), S.nK3[0], S.nK3[1], _p(S.FL.channel(ch).K3), _p(S.FL.γt.K3), _p(S.FL.γa.K3), _p(cG), _p(cF), _p(cF0),
                     _p(_pi(S, ch, True)), _p(_pi(S, ch, False)), C.byref(sg_struct(sg)), s1, s2, ch, int(is_mfRG), C.byref(S.grid))
    if ch == tCh:
        _tfix(S.Fbuff.γt.K3, S.Fbuff.γa.K3)


def BSE_K1_new(S, ch, is_mfRG=False, c0=0, c1=-1):
    """BSE_K1_new!(S, Ch, is_mfRG): src/BSE_templates.jl:188-218 -> src/nonlocal_2/BSEa/BSEa_K1.jl:62-113"""
    sign, Sp = _sign_sp(ch)
    K1 = S.Fbuff.channel(ch).K1
    lib().orc_bse_K1_new(_p(K1), S.nK1, C.byref(vertex_struct(S.F0)), C.byref(vertex_struct(S.F)),
                         _p(_pi(S, ch, True)), _p(_pi(S, ch, False)), C.byref(sg_struct(S.sg[SG_K1])), sign, ch, Sp, int(is_mfRG),
                         C.byref(S.grid), C.c_int64(c0), C.c_int64(c1))
    if ch == tCh:
        _tfix(S.Fbuff.γt.K1, S.Fbuff.γa.K1)


def BSE_K2_new(S, ch, is_mfRG=False, c0=0, c1=-1):
    """BSE_K2_new!(S, Ch, is_mfRG): src/BSE_templates.jl:223-253 -> src/nonlocal_2/BSEa/BSEa_K2.jl:142-216"""
    sign, Sp = _sign_sp(ch)
    K2 = S.Fbuff.channel(ch).K2
    sg = S.sg[SG_PP2 if ch == pCh else SG_PH2]
    lib().orc_bse_K2_new(_p(K2), S.nK2[0], S.nK2[1], C.byref(vertex_struct(S.F0)), C.byref(vertex_struct(S.F)),
                         _p(_pi(S, ch, True)), _p(_pi(S, ch, False)), C.byref(sg_struct(sg)), sign, ch, Sp, int(is_mfRG),
                         C.byref(S.grid), C.c_int64(c0), C.c_int64(c1))
    if ch == tCh:
        _tfix(S.Fbuff.γt.K2, S.Fbuff.γa.K2)


def BSE_K1_1loop(S, ch, is_mfRG=False, c0=0, c1=-1):
    """BSE_K1_1loop!(S, Ch, is_mfRG): src/BSE_templates.jl:261-291 -> src/nonlocal_2/BSEa/BSE_1loop.jl:2-56"""
    sign, Sp = _sign_sp(ch)
    K1 = S.Fbuff.channel(ch).K1
    lib().orc_bse_K1_1loop(_p(K1), S.nK1, C.byref(vertex_struct(S.F0)), C.byref(vertex_struct(S.F)), C.byref(vertex_struct(S.FL)),
                           _p(_pi(S, ch, True)), _p(_pi(S, ch, False)), C.byref(sg_struct(S.sg[SG_K1])), sign, ch, Sp, int(is_mfRG),
                           C.byref(S.grid), C.c_int64(c0), C.c_int64(c1))
    if ch == tCh:
        _tfix(S.Fbuff.γt.K1, S.Fbuff.γa.K1)


def BSE_K2_1loop(S, ch, is_mfRG=False, c0=0, c1=-1):
    """BSE_K2_1loop!(S, Ch, is_mfRG): src/BSE_templates.jl:297-327 -> src/nonlocal_2/BSEa/BSE_1loop.jl:59-124"""
    sign, Sp = _sign_sp(ch)
    K2 = S.Fbuff.channel(ch).K2
    sg = S.sg[SG_PP2 if ch == pCh else SG_PH2]
    lib().orc_bse_K2_1loop(_p(K2), S.nK2[0], S.nK2[1], C.byref(vertex_struct(S.F0)), C.byref(vertex_struct(S.F)), C.byref(vertex_struct(S.FL)),
                           _p(_pi(S, ch, True)), _p(_pi(S, ch, False)), C.byref(sg_struct(sg)), sign, ch, Sp, int(is_mfRG),
                           C.byref(S.grid), C.c_int64(c0), C.c_int64(c1))
    if ch == tCh:       # BSE_1loop.jl:114-119
        K2 += 2 * S.FL.γt.K2
        K2 += -1 * S.FL.γa.K2
        _tfix(S.Fbuff.γt.K2, S.Fbuff.γa.K2)
    else:
        K2 += S.FL.channel(ch).K2


def BSE_K3_1loop(S, ch, is_mfRG=False):
    """BSE_K3_1loop!(S, Ch, is_mfRG): src/BSE_templates.jl:332-358 -> src/nonlocal_2/BSEa/BSE_1loop.jl:123-199"""
    cG, cF, cF0, s1, s2 = {aCh: (S.cache_Γa, S.cache_Fa, S.cache_F0a, +1, +1), pCh: (S.cache_Γpx, S.cache_Fp, S.cache_F0p, -1, +1),
                           tCh: (S.cache_Γt, S.cache_Ft, S.cache_F0t, -1, -1)}[ch]
    sg = S.sg[SG_PP3 if ch == pCh else SG_PH3]
    K3 = S.Fbuff.channel(ch).K3
    lib().orc_bse_K3_1loop(_p(K3), S.nK3[0], S.nK3[1], _p(S.FL.channel(ch).K3), _p(S.FL.γt.K3), _p(S.FL.γa.K3), _p(cG), _p(cF), _p(cF0),
                           _p(_pi(S, ch, True)), _p(_pi(S, ch, False)), C.byref(sg_struct(sg)), s1, s2, ch, int(is_mfRG), C.byref(S.grid))
    if ch == tCh:
        _tfix(S.Fbuff.γt.K3, S.Fbuff.γa.K3)


def SDE_channel_L(S, Lout, Π, V, level, is_pp, c0=0, c1=-1):
    sg = S.sg[SG_PP2 if is_pp else SG_PH2]
    getattr(lib(), "orc_nl_sde_L" if S.swave else "orc_sde_L")(_p(Lout), S.nK2[0], S.nK2[1], C.byref(vertex_struct(V)), level, _p(Π), C.byref(sg_struct(sg)), int(is_pp),
                    C.byref(S.grid), C.c_int64(c0), C.c_int64(c1))


def SDE_compute(S, G, Πpp, Πph, V, level, include_U2=True, include_Hartree=True):
    """SDE_compute!(copy(Σ), G, Πpp, Πph, Lpp, Lph, F = chain(V)[level:], ...): src/nonlocal_2/SDE.jl:154-324"""
    chain = vertex_chain(V)
    U = chain[-1].U
    SDE_channel_L(S, S.Lpp, Πpp, V, level, True)
    SDE_channel_L(S, S.Lph, Πph, V, level, False)
    Σ = zeros(S.Σ.shape)
    sgΣ = sg_struct(S.sg[SG_SIGMA])
    getattr(lib(), "orc_nl_sde_inner" if S.swave else "orc_sde_real_space")(_p(Σ), S.nG, S.LG, _p(G), S.nG, S.LG, _p(S.Lpp), _p(S.Lph), S.nK2[0], S.nK2[1], C.byref(sgΣ), C.byref(S.grid))
    if isinstance(chain[level], RefVertex):
        Σ *= 1 / 3
    if include_U2:
        ΣU2 = zeros(S.Σ.shape)
        lib().orc_sde_U2(_p(ΣU2), _p(G), S.nG, S.LG, C.c_double(U.real), C.c_double(U.imag), C.c_double(S.T), C.byref(sgΣ))
        Σ += ΣU2
    if include_Hartree:
        n = compute_occupation(S, G)
        Σ += (n - 1 / 2) * U * 1j
    return Σ


def _SDE_chain(S, Σ, G, Πpp, Πph, V, level, include_U2, include_Hartree):
    # SDE!(Σ, G, ..., F): src/SDE.jl:35-48
    chain = vertex_chain(V)
    for l in range(level, len(chain)):
        top = l == level
        Σ += SDE_compute(S, G, Πpp, Πph, V, l, include_U2 and top, include_Hartree and top)
    return Σ


# Quirk toggle: True (default) = as coded in src/SDE.jl:13-24: the reference Hartree term is
# subtracted inside SDE!(..., G0, ..., F0; include_Hartree) AND once more explicitly.  False = subtracted once, which is
# what reproduces the reference's own golden number test/test_siam_fdPA.jl:84 (see tests/test_oracle_siam_golden.py).
QUIRK_E1 = True


def SDE(S, strategy="scPA", include_U2=True, include_Hartree=True):
    """SDE!(S; strategy): src/SDE.jl:3-33"""
    S.Σ[...] = 0
    _SDE_chain(S, S.Σ, S.G, S.Πpp, S.Πph, S.F, 0, include_U2, include_Hartree)
    if strategy in ("fdPA", "fdPA_new", "fdPA_1loop"):
        Σ0part = _SDE_chain(S, zeros(S.Σ.shape), S.G0, S.Π0pp, S.Π0ph, S.F, 1, include_U2, include_Hartree)
        S.Σ += -1 * Σ0part
        S.Σ += S.Σ0
        if include_Hartree and QUIRK_E1:
            n0 = compute_occupation(S, S.G0)
            S.Σ -= (n0 - 1 / 2) * S.F.bare_vertex() * 1j


def iterate_solver(S, strategy="fdPA", update_Σ=True, compute_Hartree=True):
    """iterate_solver!(S; strategy, update_Σ, compute_Hartree): src/solve.jl:4-116"""
    assert strategy in ("fdPA", "scPA", "scPA_new", "fdPA_new", "fdPA_1loop"), "Calculation strategy unknown"
    assert not S.swave or strategy in ("fdPA", "scPA"), "s-wave solver: only the fdPA / scPA strategies are restated"
    order = (pCh, aCh, tCh)
    if update_Σ:
        Dyson(S)
        bubbles(S)
    build_K3_cache(S)
    if strategy in ("fdPA_new", "scPA_new"):          # src/solve.jl:26-45
        if strategy == "fdPA_new":
            for ch in order:
                BSE_L_K3(S, ch)
        for ch in order:
            BSE_K3(S, ch)
        for ch in order:
            BSE_K1_new(S, ch)
        for ch in order:
            BSE_K2_new(S, ch)
    elif strategy == "fdPA_1loop":                    # src/solve.jl:47-58
        for ch in order:
            BSE_K3_1loop(S, ch)
        for ch in order:
            BSE_K1_1loop(S, ch)
        for ch in order:
            BSE_K2_1loop(S, ch)
    else:
        if strategy == "fdPA":
            for ch in order:
                BSE_L_K2(S, ch)
            for ch in order:
                BSE_L_K3(S, ch)
        for ch in order:
            BSE_K1(S, ch)
        for ch in order:
            BSE_K2(S, ch)
        for ch in order:
            BSE_K3(S, ch)
    S.F.set(S.Fbuff)
    if update_Σ:
        SDE(S, strategy, include_Hartree=compute_Hartree)      # src/solve.jl:99


def _fourier_interpolate(yi, Lo, Li):
    """_fourier_interpolate!(yi, Lo, Li): src/interpolate.jl:1-58, vector (Li^2) and matrix (Li^2 x Li^2) methods, restated
    literally: fft / Li^d, coefficients R in [-Li/2, Li/2] copied with half weights at |R_c| = Li/2 (even Li), bfft."""
    d = 2 * yi.ndim
    yR = np.fft.fftn(yi.reshape((Li,) * d, order="F")) / Li ** d
    yo = np.zeros((Lo,) * d, dtype=np.complex128)
    Rs = range(-(Li // 2), Li // 2 + 1)
    import itertools
    for R in itertools.product(Rs, repeat=d):
        w = 1.0
        if Li % 2 == 0:
            for c in R:
                if abs(c) == Li // 2:
                    w /= 2
        yo[tuple(c % Lo for c in R)] += yR[tuple(c % Li for c in R)] * w
    yo = np.fft.ifftn(yo) * Lo ** d
    return yo.reshape((Lo * Lo,) * yi.ndim, order="F")


def interpolate_array(Ko, Ki, nfreq, Lo, Li, shifts, clamp=False):
    """interpolate_vertex!(Ko, Ki) for any class: the first `nfreq` axes are Matsubara meshes (index shift i_in = i_out + shift),
    the rest momentum axes (src/interpolate.jl:62-165).  clamp: edge values outside the input box (Σ in interpolate_solver!)."""
    Ko[...] = 0
    import itertools
    for io in itertools.product(*[range(n) for n in Ko.shape[:nfreq]]):
        ii = []
        ok = True
        for d, i in enumerate(io):
            j = i + shifts[d]
            if j < 0 or j >= Ki.shape[d]:
                if clamp:
                    j = 0 if j < 0 else Ki.shape[d] - 1
                else:
                    ok = False
            ii.append(j)
        if ok:
            Ko[io] = _fourier_interpolate(np.ascontiguousarray(Ki[tuple(ii)]), Lo, Li)


def interpolate_vertex(Fo, Fi):
    """the channel / class loop of interpolate_solver! (src/interpolate.jl:199-206) on host NL2 vertices"""
    for go, gi in zip(Fo.channels(), Fi.channels()):
        interpolate_array(go.K1, gi.K1, 1, Fo.L, Fi.L, (Fi.numK1 - Fo.numK1,))
        interpolate_array(go.K2, gi.K2, 2, Fo.L, Fi.L, (Fi.numK2[0] - Fo.numK2[0], Fi.numK2[1] - Fo.numK2[1]))
        interpolate_array(go.K3, gi.K3, 3, Fo.L, Fi.L, (Fi.numK3[0] - Fo.numK3[0],) + (Fi.numK3[1] - Fo.numK3[1],) * 2)


def interpolate_solver(So, Si, *, occ_target=None, hubbard_params=None):
    """interpolate_solver!(So, Si; occ_target, hubbard_params): src/interpolate.jl:168-213"""
    interpolate_array(So.Σ, Si.Σ, 1, So.LG, Si.LG, (Si.nG - So.nG,), clamp=True)
    if occ_target is not None:
        μ = compute_hubbard_chemical_potential(occ_target, So, hubbard_params)
        So.Gbare[...] = hubbard_bare_Green(So.T, So.nG, So.LG, μ=μ, **hubbard_params)
    Dyson(So)
    bubbles(So)
    interpolate_vertex(So.F, Si.F)
    symmetrize_solver(So)


def symmetrize_solver(S):
    """symmetrize_solver!(S): src/ParquetSolver.jl:246-259"""
    def sym(which, a):
        lib().orc_symmetrize(_p(a), C.byref(sg_struct(S.sg[which])))
    sym(SG_SIGMA, S.Σ)
    for ch in (aCh, pCh, tCh):
        g = S.F.channel(ch)
        sym(SG_K1, g.K1)
        sym(SG_PP2 if ch == pCh else SG_PH2, g.K2)
        sym(SG_PP3 if ch == pCh else SG_PH3, g.K3)


def fixed_point_preconditioned(R, x, S, *, strategy="fdPA", use_preconditioner=True, krylov_maxiter=400, memory=100):
    """fixed_point_preconditioned!(R, x, S; strategy, update_Σ = false): src/mfRG.jl:93-171"""
    nF_ = len(S.F)
    S.F.unflatten(np.asarray(x[:nF_]))
    symmetrize_solver(S)
    iterate_solver(S, strategy, False)
    R_F = S.F.flatten() - x[:nF_]
    stats = {"niter": 0, "solved": True}
    if use_preconditioner:
        xsol, stats = dqgmres(mfRGLinearMap(S, strategy), R_F, atol=1e-6, rtol=1e-6, itmax=krylov_maxiter, memory=memory)
        R[:nF_] = xsol
    else:
        R[:nF_] = R_F
    return stats["niter"], stats["solved"]


def _anderson(fixed_point, x0, *, m=50, beta=0.85, ftol=1e-4, iterations=40):
    """nlsolve(...; method = :anderson) stand-in (same update rule as the product's host driver, written out independently)"""
    x = np.array(x0, dtype=np.complex128, copy=True)
    Xs, Rs = [], []
    err = np.inf
    for it in range(1, iterations + 1):
        R = np.asarray(fixed_point(x))
        err = float(np.max(np.abs(R)))
        if err <= ftol:
            return x, True, it
        Xs.append(x.copy()); Rs.append(R.copy())
        if len(Xs) > m + 1:
            Xs.pop(0); Rs.pop(0)
        if len(Xs) == 1:
            x = x + beta * R
        else:
            dR = np.stack([Rs[i + 1] - Rs[i] for i in range(len(Rs) - 1)], axis=1)
            dX = np.stack([Xs[i + 1] - Xs[i] for i in range(len(Xs) - 1)], axis=1)
            gamma, *_ = np.linalg.lstsq(dR, R, rcond=None)
            x = x + beta * R - (dX + beta * dR) @ gamma
    return x, False, iterations


def solve_using_mfRG(S, *, maxiter=100, occ_target=None, hubbard_params=None, mixing_init=1.0, tol=1e-4, strategy="fdPA",
                     anderson_iterations=40, anderson_m=50, krylov_maxiter=400, memory=100):
    """solve_using_mfRG!(S; ...): src/mfRG.jl:217-372 on the oracle's host arrays (no checkpoint files)"""
    mixing, it = float(mixing_init), 0
    hist = {"mixing": [], "Σ_err": [], "μ": [], "anderson_iterations": [], "converged": False}
    for _ in range(maxiter):
        it += 1
        Πpp_mixed = S.Πpp * mixing + S.Π0pp * (1 - mixing)
        Πph_mixed = S.Πph * mixing + S.Π0ph * (1 - mixing)
        S.Πpp[...] = Πpp_mixed
        S.Πph[...] = Πph_mixed
        nF_ = len(S.F)

        def fp(x):
            R = np.empty(nF_, dtype=np.complex128)
            fixed_point_preconditioned(R, x, S, strategy=strategy, krylov_maxiter=krylov_maxiter, memory=memory)
            return R
        zero, ok, nit = _anderson(fp, S.F.flatten(), m=anderson_m, beta=0.85, ftol=tol, iterations=anderson_iterations)
        hist["anderson_iterations"].append(nit)
        if not ok:
            mixing /= 2.0
            it -= 1
            bubbles(S)
            continue
        used = mixing
        mixing = min(1.0, mixing * 1.2)
        S.F.unflatten(zero)
        bubbles_real_space(S, S.Π0pp, S.Π0ph, S.G0)
        bubbles_real_space(S, S.Πpp, S.Πph, S.G)
        SDE(S, "scPA")
        Σ_err = float(np.max(np.abs(S.Σ - S.Σ0))) / mixing
        S.Π0pp[...] = Πpp_mixed
        S.Π0ph[...] = Πph_mixed
        S.G0[...] = S.G
        S.Σ0[...] = S.Σ
        for g0, g in zip(S.F0.channels(), S.F.channels()):        # add!(S.F0, S.F); set!(S.F, 0)
            for a0, a in zip(g0.arrays(), g.arrays()):
                a0 += a
                a[...] = 0
        if occ_target is not None:
            μ = compute_hubbard_chemical_potential(occ_target, S, hubbard_params)
            S.Gbare[...] = hubbard_bare_Green(S.T, S.nG, S.LG, μ=μ, **hubbard_params)
            hist["μ"].append(μ)
        Dyson(S)
        bubbles(S)
        hist["mixing"].append(used)
        hist["Σ_err"].append(Σ_err)
        if Σ_err < tol:
            hist["converged"] = True
            break
    return hist


class mfRGLinearMap:
    """src/mfRG.jl:20-89"""

    def __init__(self, S, strategy="fdPA"):
        if strategy not in ("fdPA", "fdPA_new", "fdPA_1loop"):
            raise ValueError(f"Invalid strategy {strategy}. Must be fdPA or fdPA_new or fdPA_1loop.")      # src/mfRG.jl:26-28
        self.S = S
        self.strategy = strategy
        self.is_first_iteration = True
        n = len(S.F)
        self.shape = (n, n)

    def matvec(self, x):
        S = self.S
        factor = 1e-2
        S.F.unflatten(np.asarray(x) * factor)
        build_K3_cache_mfRG(S, self.is_first_iteration)
        self.is_first_iteration = False
        order = (pCh, aCh, tCh)
        if self.strategy in ("fdPA", "fdPA_1loop"):        # src/mfRG.jl:51-64
            for ch in order:
                BSE_L_K2(S, ch)
            for ch in order:
                BSE_K1(S, ch, True)
            for ch in order:
                BSE_K2(S, ch, True)
        else:                                              # :fdPA_new, src/mfRG.jl:65-74
            for ch in order:
                BSE_K1_new(S, ch, True)
            for ch in order:
                BSE_K2_new(S, ch, True)
        for ch in order:
            BSE_L_K3(S, ch)
        for ch in order:
            BSE_K3(S, ch, True)
        S.F.set(S.Fbuff)
        y = S.F.flatten()
        return x - y / factor

    __matmul__ = matvec


def sym_givens(a, b):
    """Complex Givens rotation [c s; -conj(s) c] [a; b] = [rho; 0] with real c (b real, >= 0)."""
    if b == 0:
        return 1.0, 0.0 + 0.0j, a
    if a == 0:
        return 0.0, 1.0 + 0.0j, b + 0.0j
    t = np.hypot(abs(a), b)
    ph = a / abs(a)
    return abs(a) / t, ph * b / t, ph * t


def dqgmres(A, b, *, memory=20, atol=1e-6, rtol=1e-6, itmax=0):
    """DQGMRES (Saad & Wu, "DQGMRES: a direct quasi-minimal residual algorithm based on incomplete orthogonalization",
    Numer. Linear Algebra Appl. 3 (1996) 329): the solver the reference calls as Krylov.dqgmres(mfRGLinearMap(S, strategy), R_F;
    atol = 1e-6, rtol = 1e-6, itmax, memory = 100) (src/mfRG.jl:147-151; Krylov.jl is a dependency, not in the tree).
    No preconditioner, x0 = 0, modified Gram-Schmidt over the last `memory` Krylov vectors; the residual estimate of the
    stopping test |gamma_{m+1}| <= atol + rtol * ||b|| is the quasi-residual norm.  Returns (x, stats)."""
    b = np.asarray(b, dtype=np.complex128)
    n = b.size
    itmax = itmax if itmax > 0 else 2 * n
    x = np.zeros(n, dtype=np.complex128)
    beta = float(np.linalg.norm(b))
    stats = {"niter": 0, "solved": beta == 0.0, "residuals": [beta]}
    if beta == 0.0:
        return x, stats
    eps = atol + rtol * beta
    k = int(memory)
    V, P, c, s = {}, {}, {}, {}
    V[1] = b / beta
    gamma = beta + 0.0j
    for m in range(1, itmax + 1):
        w = np.asarray(A.matvec(V[m]), dtype=np.complex128).copy()
        lo = max(1, m - k + 1)
        t = {}
        for i in range(lo, m + 1):
            t[i] = np.vdot(V[i], w)
            w -= t[i] * V[i]
        hnext = float(np.linalg.norm(w))
        plo = max(1, m - k)
        if plo < lo:
            t[plo] = 0.0 + 0.0j
        tn = {m + 1: hnext + 0.0j}
        for i in range(plo, m):                   # previous rotations on rows (i, i + 1)
            ti, tj = t[i], t[i + 1]
            t[i] = c[i] * ti + s[i] * tj
            t[i + 1] = -np.conj(s[i]) * ti + c[i] * tj
        c[m], s[m], rmm = sym_givens(t[m], hnext)
        gamma_next = -np.conj(s[m]) * gamma
        gamma = c[m] * gamma
        p = V[m].copy()
        for i in range(plo, m):
            p -= t[i] * P[i]
        p /= rmm
        P[m] = p
        x += gamma * p
        gamma = gamma_next
        rnorm = abs(gamma)
        stats["niter"] = m
        stats["residuals"].append(rnorm)
        if rnorm <= eps:
            stats["solved"] = True
            break
        if hnext == 0.0:
            break
        V[m + 1] = w / hnext
        V.pop(m - k + 1, None)                    # iteration m + 1 orthogonalises against V[m-k+2 .. m+1] ...
        for d in (P, c, s):                       # ... and uses P, c, s of m-k+1 .. m
            d.pop(m - k, None)
    return x, stats


# ================================================================================================ local solver
class OracleLocalSolver(OracleSolver):
    """CPU restatement of the local ParquetSolver (src/ParquetSolver.jl:5-157): an OracleSolver on a 1 x 1 momentum mesh
    with the local bubbles (1/ν tail, mΠν_factor = 6), the local BSE_L_K2! and the local SDE L kernels (own γ only)."""

    def __init__(self, nK1, nK2, nK3, Gbare, G0, Σ0, F0, *, T, mΠν_factor=6, VT=NL2_Vertex):
        # VT = NL2_MBEVertex: ParquetSolver(...; VT = MBEVertex), carried on the 1 x 1 mesh like everything local
        col = lambda a: np.asfortranarray(np.asarray(a, dtype=np.complex128).reshape(-1, 1))
        super().__init__(nK1, nK2, nK3, 1, col(Gbare), col(G0), col(Σ0), F0, T=T, mΠν_factor=mΠν_factor, compute_bubbles=False, VT=VT)
        self.local = True
        bubbles_local(self, self.Π0pp, self.Π0ph, self.G0)
        Dyson(self)
        bubbles_local(self, self.Πpp, self.Πph, self.G)


def siam_bare_Green(T, nG, *, e, Δ, D):
    G = zeros((2 * nG, 1))
    lib().orc_siam_bare_green(_p(G), nG, C.c_double(T), C.c_double(e), C.c_double(Δ), C.c_double(D))
    return G


def bubbles_local(S, Πpp, Πph, G):
    lib().orc_bubbles_local(_p(Πpp), _p(Πph), _p(G), S.nG, C.byref(S.grid))


def BSE_L_K2_local(S, ch, c0=0, c1=-1):
    sign, Sp = _sign_sp(ch)
    K2 = S.FL.channel(ch).K2
    sg = S.sg[SG_PP2 if ch == pCh else SG_PH2]
    lib().orc_bse_L_K2_local(_p(K2), S.nK2[0], S.nK2[1], C.byref(vertex_struct(S.F0)), C.byref(vertex_struct(S.F)),
                             _p(_pi(S, ch, True)), C.byref(sg_struct(sg)), sign, ch, Sp, C.byref(S.grid), C.c_int64(c0), C.c_int64(c1))
    if ch == tCh and c1 < 0:
        _tfix(S.FL.γt.K2, S.FL.γa.K2)


def iterate_solver_local(S, strategy="fdPA", update_Σ=True):
    """iterate_solver!(S::ParquetSolver): src/solve.jl:4-116 with the local kernels"""
    if update_Σ:
        Dyson(S)
        bubbles_local(S, S.Πpp, S.Πph, S.G)
    build_K3_cache(S)
    if strategy == "fdPA":
        for ch in (pCh, aCh, tCh):
            BSE_L_K2_local(S, ch)
        for ch in (pCh, aCh, tCh):
            BSE_L_K3(S, ch)
    for ch in (pCh, aCh, tCh):
        BSE_K1(S, ch)
    for ch in (pCh, aCh, tCh):
        BSE_K2(S, ch)
    for ch in (pCh, aCh, tCh):
        BSE_K3(S, ch)
    S.F.set(S.Fbuff)
    if update_Σ:
        lib().orc_set_quirk_E2(0)          # local SDE L kernels: F(...; F0 = false, own γ) (src/SDE.jl:102-103, 138-140)
        try:
            SDE(S, strategy)
        finally:
            lib().orc_set_quirk_E2(1)


def interp_boson(K1col, T, N, x):
    """MeshFunction call with a Float64 argument: linear interpolation on the bosonic Matsubara mesh"""
    m = x / (2 * np.pi * T)
    m0 = int(np.floor(m))
    if m0 >= N - 1:
        m0 = N - 2
    w = m - m0
    return (1 - w) * K1col[m0 + N - 1] + w * K1col[m0 + 1 + N - 1]
