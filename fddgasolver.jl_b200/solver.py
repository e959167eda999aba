"""Host-side mirror of the reference's NL2_ParquetSolver API, driving libfdga through the C-ABI.

Same names, argument meaning and call order as the reference (Julia ``!`` dropped):
  NL2_ParquetSolver                     src/nonlocal_2/ParquetSolver.jl:1-154
  parquet_solver_hubbard_parquet_approximation_NL2   :167-196
  init_sym_grp!                         :200-291
  Dyson!, compute_occupation            src/dyson.jl
  bubbles!                              src/nonlocal_2/ParquetSolver.jl:309-312 -> nonlocal_2/bubble.jl:42-122
  build_K3_cache!, build_K3_cache_mfRG! src/nonlocal_2/build_K3_cache.jl
  BSE_L_K2!, BSE_L_K3!, BSE_K1!, BSE_K2!, BSE_K3!    src/BSE_templates.jl:12-180
  SDE!                                  src/SDE.jl:3-48
  iterate_solver!, fixed_point!         src/solve.jl:4-157
  mfRGLinearMap                         src/mfRG.jl:20-89
All compute happens on the GPU inside libfdga; this module only moves arrays and sequences calls.
The solver keeps host copies of its arrays (like the Julia struct); ``push``/``pull`` synchronise them.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from .types import (CH_NAME, CHANNELS, MBEVertex, NL2_MBEVertex, NL2_Vertex, NL_Vertex, RefVertex, Vertex, aCh, nB, nF, pCh, tCh, vertex_chain, zeros)

INF_FREQ = 1 << 28           # νInf: an InfiniteMatsubaraFrequency argument of eval_vertex
STRATEGY = {"scPA": L.SCPA, "fdPA": L.FDPA, "scPA_new": L.SCPA_NEW, "fdPA_new": L.FDPA_NEW, "fdPA_1loop": L.FDPA_1LOOP}
_G_NAMES = {"G": L.G, "G0": L.G0, "Gbare": L.GBARE, "Σ": L.SIGMA, "Σ0": L.SIGMA0}
_PI_NAMES = {"Π0pp": L.PI0PP, "Π0ph": L.PI0PH, "Πpp": L.PIPP, "Πph": L.PIPH}
_CACHE_NAMES = ["cache_Γpx", "cache_F0p", "cache_F0a", "cache_F0t", "cache_Γpp", "cache_Γa", "cache_Γt",
                "cache_Fp", "cache_Fa", "cache_Ft"]


class NL2_ParquetSolver:
    """State of one nonlocal (NL2) parquet / fdDΓA problem, resident on one GPU.

    Gbare, G0, Σ0: complex128 arrays of shape (2 nG, LG*LG) holding i*G (src/models/hubbard.jl:23-24).
    F0: reference vertex: RefVertex | Vertex | NL2_Vertex (arbitrarily nested).
    L: linear size of the vertex momentum mesh mK_Γ.
    """

    _vertex_cls = NL2_Vertex
    swave = False

    def _pi_shape(self):
        return (nB(self.nΠB), nF(self.nΠF), self.NP, self.NP)

    def __init__(self, nK1, nK2, nK3, L_, Gbare, G0, Σ0, F0, *, T, mode="threads", mΠν_factor=1, device=0,
                 compute_bubbles=True, VT=None):
        # VT = NL2_MBEVertex: S.F is a multi-boson-exchange vertex (NL2_ParquetSolver(..., F0, NL2_MBEVertex), src/nonlocal_2/
        # ParquetSolver.jl:86,113); Fbuff and FL stay asymptotic.  Such a solver runs the generic per-term kernels: scPA / fdPA only.
        lib = L.load()
        self._lib = lib
        self.mode = mode          # accepted and ignored: parallelism is the GPU's
        self.T = float(T)
        self.L = int(L_)
        self.NP = self.L * self.L
        self.nK1, self.nK2, self.nK3 = int(nK1), tuple(nK2), tuple(nK3)
        Gbare = np.asfortranarray(Gbare, dtype=np.complex128)
        assert Gbare.ndim == 2 and Gbare.shape == G0.shape == Σ0.shape
        self.nG = Gbare.shape[0] // 2
        self.LG = int(round(np.sqrt(Gbare.shape[1])))
        assert self.LG * self.LG == Gbare.shape[1]
        self.nΠB, self.nΠF = self.nK1, self.nK1 * int(mΠν_factor)     # ParquetSolver.jl:95-97

        self.Gbare = Gbare
        self.G0 = np.asfortranarray(G0, dtype=np.complex128).copy(order="F")
        self.Σ0 = np.asfortranarray(Σ0, dtype=np.complex128).copy(order="F")
        self.G = self.G0.copy(order="F")
        self.Σ = self.Σ0.copy(order="F")
        self.F0 = F0
        self.F = (VT or self._vertex_cls)(F0, self.T, nK1, nK2, nK3, self.L)
        if not isinstance(self.F, self._vertex_cls):
            raise L.FdgaError(f"VT must be a {self._vertex_cls.__name__} type")
        null = RefVertex(self.T, 0.0)
        self.Fbuff = self._vertex_cls(null, self.T, nK1, nK2, nK3, self.L)
        self.FL = self._vertex_cls(null.copy(), self.T, nK1, nK2, nK3, self.L)
        shpΠ = self._pi_shape()
        self.Π0pp, self.Π0ph, self.Πpp, self.Πph = (None,) * 4     # pulled on demand (large)
        self._shpΠ = shpΠ
        self.Lpp = zeros(self.F.γp.K2.shape)
        self.Lph = zeros(self.F.γp.K2.shape)
        for n in _CACHE_NAMES:
            setattr(self, n, zeros(self.F.γp.K3.shape))

        # ---- device context
        self._chain = vertex_chain(self.F)
        d = L.Dims()
        d.T, d.nq, d.LG, d.nG, d.nPiB, d.nPiF = self.T, self.L, self.LG, self.nG, self.nΠB, self.nΠF
        d.nlev = len(self._chain)
        if d.nlev > L.FDGA_MAX_LEVELS:
            raise L.FdgaError("vertex chain too deep")
        for i, V in enumerate(self._chain):
            lv = d.lev[i]
            if isinstance(V, RefVertex):
                lv.type = L.LV_CORE
                lv.nK3[0], lv.nK3[1] = V.numK3
                lv.U_re, lv.U_im = V.U.real, V.U.imag
            else:
                lv.type = L.LV_NL2 if isinstance(V, NL2_Vertex) else (L.LV_NL if isinstance(V, NL_Vertex) else L.LV_LOCAL)
                if getattr(V, "mbe", False):
                    lv.type = {L.LV_NL2: L.LV_NL2_MBE, L.LV_LOCAL: L.LV_LOCAL_MBE}[lv.type]
                if isinstance(V, (NL2_Vertex, NL_Vertex)) and not isinstance(V, self._vertex_cls):
                    raise L.FdgaError(f"the momentum-dependent levels of the F0 chain must be {self._vertex_cls.__name__}s")
                if isinstance(V, (NL2_Vertex, NL_Vertex)) and V.L != self.L:
                    raise L.FdgaError("all momentum-dependent levels must share the momentum mesh")
                lv.nK1 = V.numK1
                lv.nK2[0], lv.nK2[1] = V.numK2
                lv.nK3[0], lv.nK3[1] = V.numK3
        self._dims = d
        ctx = C.c_void_p()
        rc = lib.fdga_create(C.byref(d), int(device), C.byref(ctx))
        if rc != 0:
            raise L.FdgaError("fdga_create failed: " + lib.fdga_last_error(None).decode())
        self._ctx = ctx
        self._sg = {}
        self.reset_sym_grp()

        self.push("Gbare", "G0", "Σ0", "G", "Σ", "F0")
        # reference bubbles, Dyson, target bubbles (ParquetSolver.jl:100-111)
        if compute_bubbles:
            self._call("fdga_bubbles_real_space", 1)
            self._call("fdga_dyson")
            self._call("fdga_bubbles_real_space", 0)
            self.pull("G")

    # ------------------------------------------------------------------ plumbing
    def _call(self, name, *args):
        rc = getattr(self._lib, name)(self._ctx, *args)
        L.check(self._ctx, rc, name)

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.fdga_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        """options of include/fdga.h; 'sde_own_gamma' toggles the SURVEY-E2 reading of the NL2 SDE L kernels"""
        self._call("fdga_set_option", {"sde_own_gamma": 0, "generic_kernels": 1, "fd_hartree_once": 2, "local_solver": 3, "direct_k1": 4, "serial": 5, "qlane": 6}[name], int(value))

    def sync(self):
        self._call("fdga_sync")

    def _vertex_io(self, which, V, put):
        fn = "fdga_set_vertex" if put else "fdga_get_vertex"
        for ch in CHANNELS:
            g = V.channel(ch)
            for cls, a in enumerate(g.arrays()):
                self._call(fn, which, ch, cls, L.ptr(a), a.size)

    def push(self, *names):
        """host -> device for the named groups: F, F0 (whole reference chain), FL, Fbuff, G, G0, Gbare, Σ, Σ0, Π*, cache"""
        for n in names:
            if n == "F":
                self._vertex_io(0, self.F, True)
            elif n == "F0":
                for i, V in enumerate(self._chain[1:], start=1):
                    if isinstance(V, RefVertex):
                        for j, a in enumerate(V.arrays()):
                            self._call("fdga_set_core", i, j, L.ptr(a), a.size)
                    else:
                        self._vertex_io(i, V, True)
            elif n == "FL":
                self._vertex_io(L.V_FL, self.FL, True)
            elif n == "Fbuff":
                self._vertex_io(L.V_FBUFF, self.Fbuff, True)
            elif n in _G_NAMES:
                a = getattr(self, n)
                self._call("fdga_set_green", _G_NAMES[n], L.ptr(a), a.size)
            elif n in _PI_NAMES:
                a = getattr(self, n)
                self._call("fdga_set_bubble", _PI_NAMES[n], L.ptr(a), a.size)
            elif n == "cache":
                for i, cn in enumerate(_CACHE_NAMES):
                    a = getattr(self, cn)
                    self._call("fdga_set_cache", i, L.ptr(a), a.size)
            else:
                raise KeyError(n)

    def pull(self, *names):
        """device -> host for the named groups"""
        for n in names:
            if n == "F":
                self._vertex_io(0, self.F, False)
            elif n == "F0":            # K tables of the reference chain (e.g. after add!(S.F0, S.F)); the core never changes
                for i, V in enumerate(self._chain[1:], start=1):
                    if not isinstance(V, RefVertex):
                        self._vertex_io(i, V, False)
            elif n == "FL":
                self._vertex_io(L.V_FL, self.FL, False)
            elif n == "Fbuff":
                self._vertex_io(L.V_FBUFF, self.Fbuff, False)
            elif n in _G_NAMES:
                a = getattr(self, n)
                self._call("fdga_get_green", _G_NAMES[n], L.ptr(a), a.size)
            elif n in _PI_NAMES:
                if getattr(self, n) is None:
                    setattr(self, n, zeros(self._shpΠ))
                a = getattr(self, n)
                self._call("fdga_get_bubble", _PI_NAMES[n], L.ptr(a), a.size)
            elif n == "Π":
                self.pull(*_PI_NAMES)
            elif n == "cache":
                for i, cn in enumerate(_CACHE_NAMES):
                    a = getattr(self, cn)
                    self._call("fdga_get_cache", i, L.ptr(a), a.size)
            elif n == "L":
                self._call("fdga_get_L", 1, L.ptr(self.Lpp), self.Lpp.size)
                self._call("fdga_get_L", 0, L.ptr(self.Lph), self.Lph.size)
            else:
                raise KeyError(n)

    # ------------------------------------------------------------------ symmetry groups
    def _sg_len(self, which):
        if which == L.SG_SIGMA:
            return self.Σ.size
        if which == L.SG_K1:
            return self.F.γp.K1.size
        if which in (L.SG_PP2, L.SG_PH2):
            return self.F.γp.K2.size
        return self.F.γp.K3.size

    def set_symmetry_classes(self, which, offsets, index, ops):
        """Register SG.classes (flattened) for one symmetry group; what a Julia shim passes in drop-in use."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        index = np.ascontiguousarray(index, dtype=np.int64)
        ops = np.ascontiguousarray(ops, dtype=np.uint8)
        self._call("fdga_set_symmetry_classes", which, len(offsets) - 1, L.ptr(offsets), L.ptr(index), L.ptr(ops))
        self._sg[which] = (offsets, index, ops)

    def reset_sym_grp(self):
        """trivial groups (every element its own class), ParquetSolver.jl:124-132, 293-303"""
        for which in range(8):
            self.set_symmetry_classes(which, *L.trivial_symmetry_group(self._sg_len(which)))

    def init_sym_grp(self):
        """init_sym_grp!(S): src/nonlocal_2/ParquetSolver.jl:200-291"""
        n = {L.SG_SIGMA: (self.nG, 0), L.SG_K1: (self.nK1, 0),
             L.SG_PP2: self.nK2, L.SG_PH2: self.nK2,
             L.SG_PP3: self.nK3, L.SG_PH3: self.nK3, L.SG_PPL3: self.nK3, L.SG_PHL3: self.nK3}
        kind = {L.SG_PP2: L.SG_NL_PP2, L.SG_PH2: L.SG_NL_PH2} if self.swave else {}      # src/nonlocal/ParquetSolver.jl:215-250
        for which, (n0, n1) in n.items():
            nq = self.LG if which == L.SG_SIGMA else self.L
            self.set_symmetry_classes(which, *L.build_symmetry_group(kind.get(which, which), n0, n1, nq, self._sg_len(which)))

    def num_classes(self, which):
        return len(self._sg[which][0]) - 1

    # ------------------------------------------------------------------ flatten / unflatten (device resident F)
    def length_F(self):
        return int(self._lib.fdga_length_F(self._ctx))

    def flatten_F(self, out=None):
        out = np.empty(self.length_F(), dtype=np.complex128) if out is None else out
        self._call("fdga_flatten_F", L.ptr(out))
        return out

    def flatten_F_async(self, out):
        """flatten(S.F) into `out` (ideally pinned) on a side stream, overlapping whatever is issued next; valid after sync()"""
        assert out.dtype == np.complex128 and out.size == self.length_F()
        self._call("fdga_flatten_F_async", L.ptr(out))

    def unflatten_F(self, x, scale=1.0):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        assert x.size == self.length_F()
        self._call("fdga_unflatten_F", L.ptr(x), float(scale))
        self.sync()

    def unflatten_F_async(self, x, scale=1.0):
        """unflatten!(S.F, x * scale) without a host synchronisation (x must stay alive, ideally pinned)"""
        assert x.dtype == np.complex128 and x.size == self.length_F()
        self._call("fdga_unflatten_F", L.ptr(x), float(scale))

    def unflatten_F_from_root(self, x, scale=1.0, root=0, rank=0):
        """collective unflatten!(S.F, x * scale): x crosses PCIe on the root only, the other ranks get it over NVLink"""
        self._call("fdga_unflatten_F_from_root", L.ptr(x) if rank == root else None, float(scale), int(root))

    def stash_F(self):
        self._call("fdga_stash_F")

    def unstash_F(self):
        self._call("fdga_unstash_F")

    def get_green_into(self, name, out):
        self._call("fdga_get_green", _G_NAMES[name], L.ptr(out), out.size)

    # ------------------------------------------------------------------ multi-GPU (one process per GPU)
    def comm_unique_id(self):
        """128-byte ncclUniqueId (call on rank 0, broadcast with the host's own transport)"""
        uid = np.zeros(128, dtype=np.uint8)
        rc = self._lib.fdga_comm_unique_id(L.ptr(uid))
        if rc != 0:
            raise L.FdgaError("fdga_comm_unique_id failed: " + self._lib.fdga_last_error(None).decode())
        return uid

    def comm_init(self, nranks, rank, uid):
        uid = np.ascontiguousarray(uid, dtype=np.uint8)
        assert uid.size == 128
        self._call("fdga_comm_init", int(nranks), int(rank), L.ptr(uid))
        self.nranks, self.rank = int(nranks), int(rank)

    # ------------------------------------------------------------------ the vertex as a callable
    def eval_vertex(self, W, v, w, Ch, Sp, P=0, k=0, q=0, *, level=0, F0=True, γp=True, γt=True, γa=True, swave=False):
        """F(Ω, ν, ω, P, k, q, Ch, Sp; F0, γp, γt, γa) of the chain S.F (level 0), S.F0 (level 1), ... evaluated on the device at the
        broadcast of the integer arrays W, v, w (Matsubara indices; v, w may be INF_FREQ) and P, k, q (linear momentum indices
        ix + L iy, see kidx); swave: k = q = kSW.  Returns a complex array of the broadcast shape."""
        arrs = np.broadcast_arrays(*[np.asarray(x, dtype=np.int64) for x in (W, v, w, P, k, q)])
        shape = arrs[0].shape
        flat = [np.ascontiguousarray(a.reshape(-1), dtype=np.int32) for a in arrs]
        out = np.empty(flat[0].size, dtype=np.complex128)
        flags = (1 if F0 else 0) | (2 if γp else 0) | (4 if γt else 0) | (8 if γa else 0)
        self._call("fdga_eval_vertex", int(level), int(Ch), int(Sp), flags, int(bool(swave)), out.size, *[L.ptr(a) for a in flat], L.ptr(out))
        return out.reshape(shape)

    def kidx(self, x, y):
        """linear index of the (folded) Brillouin point (x, y) of the vertex momentum mesh"""
        return (np.asarray(x) % self.L) + self.L * (np.asarray(y) % self.L)

    # ------------------------------------------------------------------ CUDA graphs (include/fdga.h)
    def record(self, fn):
        """Record the asynchronous library calls made by fn() as a CUDA graph (nothing executes) and return its id for replay().
        fn must have run once eagerly before (lazy tables current) and leave the solver in the lazy state it started from."""
        self._call("fdga_graph_begin")
        try:
            fn()
        except Exception:
            gid = C.c_int(-1)
            self._lib.fdga_graph_end(self._ctx, C.byref(gid))       # abandon the recording
            raise
        gid = C.c_int(-1)
        self._call("fdga_graph_end", C.byref(gid))
        return gid.value

    def replay(self, graph_id):
        self._call("fdga_graph_launch", int(graph_id))

    def drop_graph(self, graph_id):
        self._call("fdga_graph_destroy", int(graph_id))

    # ------------------------------------------------------------------ profiling
    def profile(self, on=True):
        self._call("fdga_profile_enable", int(on))

    def profile_reset(self):
        self._call("fdga_profile_reset")

    def kernel_times(self):
        out = {}
        for i, n in enumerate(L.T_NAMES):
            ms, cnt = C.c_double(0), C.c_int64(0)
            self._call("fdga_kernel_time_ms", i, C.byref(ms), C.byref(cnt))
            out[n] = (ms.value, cnt.value)
        return out

    def measure_fp64_peak(self):
        """FP64 FMA throughput of the device in TFLOP/s (DFMA micro-benchmark inside the library)"""
        v = C.c_double(0.0)
        self._call("fdga_measure_fp64_peak", C.byref(v))
        return v.value

    def total_launches(self):
        return int(self._lib.fdga_total_launches(self._ctx))

    def stream(self):
        return self._lib.fdga_stream(self._ctx)


class NL_ParquetSolver(NL2_ParquetSolver):
    """The s-wave solver NL_ParquetSolver (src/nonlocal/ParquetSolver.jl:1-154; nl_method = 1 of script/run_Wu_point.jl): vertices
    with bosonic momentum dependence only (NL_Vertex: K2[Ω,ν,P]), bubbles Π[Ω,ν,P] with the 1/ν tail of G at R = 0
    (src/nonlocal/bubble.jl:87-158), mΠν_factor = 32.  F0: RefVertex | Vertex | NL_Vertex (nested).  Strategies: scPA, fdPA (and
    the mfRG maps built on them); the `_new` / `_1loop` variants exist for the NL2 solver only."""
    _vertex_cls = NL_Vertex
    swave = True

    def _pi_shape(self):
        return (nB(self.nΠB), nF(self.nΠF), self.NP)

    def __init__(self, nK1, nK2, nK3, L_, Gbare, G0, Σ0, F0, *, T, mode="threads", mΠν_factor=32, device=0, compute_bubbles=True):
        super().__init__(nK1, nK2, nK3, L_, Gbare, G0, Σ0, F0, T=T, mode=mode, mΠν_factor=mΠν_factor, device=device,
                         compute_bubbles=compute_bubbles)


class ParquetSolver(NL2_ParquetSolver):
    """Local (impurity) solver, src/ParquetSolver.jl:5-157, carried as an NL2 solver on a 1 x 1 momentum mesh.

    With a single momentum the NL2 kernels coincide term by term with the local ones (src/BSEa/*.jl, src/build_K3_cache.jl,
    src/SDE.jl); the pieces that differ are switched by FDGA_OPT_LOCAL_SOLVER: BSE_L_K2! in its local form,
    bubbles! with the 1/ν tail, SDE L kernels with the own-channel γ only.  Gbare, G0, Σ0: arrays of length 2 nG holding
    i*G; F0: RefVertex or the F of another ParquetSolver.  The local Vertex arrays are K1[Ω,1], K2[Ω,ν,1,1], K3[Ω,ν,ν',1]."""

    local = True

    def __init__(self, nK1, nK2, nK3, Gbare, G0, Σ0, F0, *, T, mode="threads", mΠν_factor=6, device=0, Q=np.complex128, VT=None):
        # Q = Float64 of the reference (test/test_siam_scPA.jl:22): the particle-hole symmetric impurity has real i G, i Σ and real
        # vertices.  The device arithmetic is complex128 throughout; with real inputs every imaginary part is an exact zero, and a
        # real-typed solver insists on that whenever arrays come back (the InexactError Julia would raise) -- see real_array().
        self.eltype = np.dtype(Q)
        if self.eltype == np.float64:
            for a in (Gbare, G0, Σ0):
                if np.iscomplexobj(a) and np.any(np.asarray(a).imag != 0):
                    raise L.FdgaError("ParquetSolver(Q = Float64): inputs must be real")
        col = lambda a: np.asfortranarray(np.asarray(a, dtype=np.complex128).reshape(-1, 1))
        if VT is MBEVertex:
            VT = NL2_MBEVertex                       # the local solver lives on the 1 x 1 mesh
        super().__init__(nK1, nK2, nK3, 1, col(Gbare), col(G0), col(Σ0), F0, T=T, mode=mode, mΠν_factor=mΠν_factor,
                         device=device, compute_bubbles=False, VT=VT)
        self.set_option("local_solver", 1)
        self.set_option("sde_own_gamma", 1)          # src/SDE.jl:102-103, 138-140: F(...; F0 = false, own γ)
        self._call("fdga_bubbles_local", 1)
        self._call("fdga_dyson")
        self._call("fdga_bubbles_local", 0)
        self.pull("G")


def real_array(S, a):
    """the array `a` of a solver in the solver's element type: for ParquetSolver(Q = Float64) the real part, after checking that the
    imaginary part vanishes identically (what `eltype(S.Σ.data) == Float64` guarantees in the reference)"""
    if getattr(S, "eltype", np.dtype(np.complex128)) != np.float64:
        return a
    if np.any(a.imag != 0):
        raise L.FdgaError("ParquetSolver(Q = Float64): a complex value appeared in a real-typed solver (InexactError)")
    return np.ascontiguousarray(a.real)


# ---------------------------------------------------------------------- reference-named operations
def init_sym_grp(S):
    S.init_sym_grp()


def Dyson(S):
    S._call("fdga_dyson")


def compute_occupation(S, which="G"):
    occ = C.c_double(0)
    S._call("fdga_occupation", _G_NAMES[which], C.byref(occ))
    return occ.value


def set_hubbard_bare_Green(S, *, μ, t1, t2=0.0, t3=0.0):
    """set!(S.Gbare, hubbard_bare_Green(meshes(S.Gbare)...; μ, hubbard_params...)) on the device (src/mfRG.jl:113, 355)"""
    S._call("fdga_set_hubbard_bare_green", float(μ), float(t1), float(t2), float(t3))


def compute_hubbard_chemical_potential(occ_target, S, hubbard_params):
    """compute_hubbard_chemical_potential(occ_target, S.Σ, hubbard_params): src/dyson.jl:45-57, Σ = the device copy of S.Σ"""
    mu = C.c_double(0.0)
    S._call("fdga_hubbard_chemical_potential", float(occ_target), float(hubbard_params["t1"]), float(hubbard_params.get("t2", 0.0)),
            float(hubbard_params.get("t3", 0.0)), C.byref(mu))
    return mu.value


def bubbles(S):
    """bubbles!(S) = bubbles_real_space!(S.Πpp, S.Πph, S.G)  (local solver: src/bubble.jl:9-36)"""
    S._call("fdga_bubbles_local" if getattr(S, "local", False) else "fdga_bubbles_real_space", 0)


def bubbles_real_space(S, reference=False):
    S._call("fdga_bubbles_real_space", int(reference))


def bubbles_momentum_space(S, reference=False):
    S._call("fdga_bubbles_momentum_space", int(reference))


def build_K3_cache(S):
    S._call("fdga_build_K3_cache", 0, 0)


def build_K3_cache_mfRG(S, is_first_iteration):
    S._call("fdga_build_K3_cache", 1, int(is_first_iteration))


def BSE_L_K2(S, Ch, is_mfRG=False):
    S._call("fdga_bse_L_K2", Ch)


def BSE_L_K3(S, Ch, is_mfRG=False):
    S._call("fdga_bse_L_K3", Ch)


def BSE_K1(S, Ch, is_mfRG=False):
    S._call("fdga_bse_K1", Ch, int(is_mfRG))


def BSE_K2(S, Ch, is_mfRG=False):
    S._call("fdga_bse_K2", Ch, int(is_mfRG))


def BSE_K3(S, Ch, is_mfRG=False):
    S._call("fdga_bse_K3", Ch, int(is_mfRG))


def BSE_K1_new(S, Ch, is_mfRG=False):
    """BSE_K1_new!(S, Ch, is_mfRG): src/BSE_templates.jl:188-218"""
    S._call("fdga_bse_K1_new", Ch, int(is_mfRG))


def BSE_K2_new(S, Ch, is_mfRG=False):
    """BSE_K2_new!(S, Ch, is_mfRG): src/BSE_templates.jl:223-253"""
    S._call("fdga_bse_K2_new", Ch, int(is_mfRG))


def BSE_K1_1loop(S, Ch, is_mfRG=False):
    """BSE_K1_1loop!(S, Ch, is_mfRG): src/BSE_templates.jl:261-291"""
    S._call("fdga_bse_K1_1loop", Ch, int(is_mfRG))


def BSE_K2_1loop(S, Ch, is_mfRG=False):
    """BSE_K2_1loop!(S, Ch, is_mfRG): src/BSE_templates.jl:297-327"""
    S._call("fdga_bse_K2_1loop", Ch, int(is_mfRG))


def BSE_K3_1loop(S, Ch, is_mfRG=False):
    """BSE_K3_1loop!(S, Ch, is_mfRG): src/BSE_templates.jl:332-358"""
    S._call("fdga_bse_K3_1loop", Ch, int(is_mfRG))


def SDE(S, strategy="scPA", include_U2=True, include_Hartree=True):
    if strategy not in STRATEGY:
        raise ValueError(f"Calculation strategy {strategy} unknown")      # src/SDE.jl:31
    S._call("fdga_sde", STRATEGY[strategy], int(include_U2), int(include_Hartree))


def SDE_channel_L(S, reference=False, level=0):
    """SDE_channel_L_pp! / SDE_channel_L_ph! (src/nonlocal_2/SDE.jl:16-33, 54-73) for the chain starting at `level`, summed over
    its levels with SDE!'s weights; S.pull("L") fetches S.Lpp, S.Lph"""
    S._call("fdga_sde_channel_L", int(reference), int(level))


def iterate_solver(S, strategy="fdPA", update_Σ=True, compute_Hartree=True):
    """iterate_solver!(S; strategy, update_Σ, compute_Hartree): src/solve.jl:4-116 (fused inside the library)"""
    assert strategy in STRATEGY, "Calculation strategy unknown"      # src/solve.jl:10
    S._call("fdga_iterate_solver", STRATEGY[strategy], int(update_Σ), int(compute_Hartree))


def iterate_solver_stepwise(S, strategy="fdPA", update_Σ=True, compute_Hartree=True):
    """Same sequence as iterate_solver, issued call by call (used by the tests to check the fused driver)."""
    if update_Σ:
        Dyson(S)
        bubbles(S)
    build_K3_cache(S)
    order = (pCh, aCh, tCh)
    if strategy in ("fdPA_new", "scPA_new"):
        stages = ([BSE_L_K3] if strategy == "fdPA_new" else []) + [BSE_K3, BSE_K1_new, BSE_K2_new]
    elif strategy == "fdPA_1loop":
        stages = [BSE_K3_1loop, BSE_K1_1loop, BSE_K2_1loop]
    else:
        stages = ([BSE_L_K2, BSE_L_K3] if strategy == "fdPA" else []) + [BSE_K1, BSE_K2, BSE_K3]
    for stage in stages:
        for ch in order:
            stage(S, ch)
    S._call("fdga_set_F_from_Fbuff")
    if update_Σ:
        SDE(S, strategy, include_Hartree=compute_Hartree)


def fixed_point(R, x, S, strategy="fdPA", update_Σ=True, compute_Hartree=True):
    """fixed_point!(R, x, S): R = iterate(x) - x on the flattened [F; Σ] (src/solve.jl:119-157, src/ParquetSolver.jl:277-306)"""
    nF_ = S.length_F()
    S.unflatten_F(x[:nF_])
    if update_Σ:
        S.Σ[...] = np.asarray(x[nF_:]).reshape(S.Σ.shape, order="F")
        S.push("Σ")
    iterate_solver(S, strategy, update_Σ, compute_Hartree)
    S.flatten_F(R[:nF_])
    if update_Σ:
        S.pull("Σ")
        R[nF_:] = S.Σ.ravel(order="F")
    R -= x
    return R


def solve(S, *, maxiter=100, tol=1e-4, δ=0.85, mem=8, verbose=False, strategy="fdPA", update_Σ=True, compute_Hartree=True):
    """solve!(S; maxiter, tol, δ, mem, kwargs_solver...): src/solve.jl:160-196 -- nlsolve(:anderson) on fixed_point! over the
    flattened [F; Σ] (or F alone).  Returns the nlsolve-like result (zero, f_converged, iterations, residual_norm).  As in the
    reference, S is left in the state of the LAST iterate_solver! call (F, Σ = its outputs; G, Π = its inputs' Dyson / bubbles):
    nothing is restored from res.zero."""
    from .nlsolve import anderson
    nF_ = S.length_F()
    x0 = S.flatten_F()
    if update_Σ:
        S.pull("Σ")
        x0 = np.concatenate([x0, S.Σ.ravel(order="F")])
    res = anderson(lambda x: fixed_point(np.empty_like(x), x, S, strategy, update_Σ, compute_Hartree), x0, m=mem, beta=δ, ftol=tol,
                   iterations=maxiter, show_trace=verbose)
    S.pull("F")
    if update_Σ:
        S.pull("Σ")
    return res


class mfRGLinearMap:
    """mfRGLinearMap(S, strategy): y = x - BSE_lin(1e-2 x)/1e-2 (src/mfRG.jl:20-89)"""

    def __init__(self, S, strategy="fdPA"):
        if strategy not in ("fdPA", "fdPA_new", "fdPA_1loop"):
            raise ValueError(f"Invalid strategy {strategy}. Must be fdPA or fdPA_new or fdPA_1loop.")    # src/mfRG.jl:26-28
        self.S = S
        self.strategy = strategy
        self.is_first_iteration = True
        n = S.length_F()
        self.shape = (n, n)

    def __matmul__(self, x):
        return self.matvec(x)

    def matvec(self, x, out=None, root=None):
        """y = A x.  `out` (optional) receives y: pass pinned host buffers for x and out to move them at full PCIe speed
        (pageable numpy arrays are staged by the driver at a fraction of it).  root (multi-rank jobs, collective): the vectors live
        on that rank's host only; x is uploaded once and broadcast over NVLink, y comes back on the root (None elsewhere)."""
        if root is not None:
            me = getattr(self.S, "rank", 0) == root
            if me:
                x = np.ascontiguousarray(x, dtype=np.complex128)
                y = np.empty_like(x) if out is None else out
            self.S._call("fdga_mfrg_matvec_from_root", L.ptr(x) if me else None, L.ptr(y) if me else None, int(self.is_first_iteration),
                         STRATEGY[self.strategy], int(root))
            self.is_first_iteration = False
            return y if me else None
        x = np.ascontiguousarray(x, dtype=np.complex128)
        y = np.empty_like(x) if out is None else out
        assert y.dtype == np.complex128 and y.size == x.size and y.flags["C_CONTIGUOUS"]
        self.S._call("fdga_mfrg_matvec_strategy", L.ptr(x), L.ptr(y), int(self.is_first_iteration), STRATEGY[self.strategy])
        self.is_first_iteration = False
        return y


def dqgmres(A, b, *, memory=20, atol=1e-6, rtol=1e-6, itmax=0, history=True):
    """Krylov.dqgmres(A::mfRGLinearMap, b; atol, rtol, itmax, memory) as called at src/mfRG.jl:147-151, device resident:
    returns (x, stats) with stats = dict(niter, solved, residuals)."""
    if not isinstance(A, mfRGLinearMap):
        raise TypeError("dqgmres: A must be an mfRGLinearMap (the device-resident Krylov solver is tied to this operator)")
    b = np.ascontiguousarray(b, dtype=np.complex128)
    x = np.empty_like(b)
    niter, solved = C.c_int(0), C.c_int(0)
    nres = (itmax if itmax > 0 else 4096) + 1
    res = np.zeros(nres, dtype=np.float64)
    A.S._call("fdga_mfrg_dqgmres", L.ptr(b), L.ptr(x), STRATEGY[A.strategy], int(memory), float(atol), float(rtol), int(itmax),
              C.byref(niter), C.byref(solved), res.ctypes.data_as(C.POINTER(C.c_double)), nres)
    A.is_first_iteration = False
    return x, {"niter": niter.value, "solved": bool(solved.value), "residuals": res[: min(nres, niter.value + 1)].tolist()}


def symmetrize_solver(S):
    """symmetrize_solver!(S): src/ParquetSolver.jl:246-259 (on the device state)"""
    S._call("fdga_symmetrize_solver")


def fixed_point_preconditioned(R, x, S, *, strategy="fdPA", use_preconditioner=True, krylov_maxiter=400, memory=100):
    """fixed_point_preconditioned!(R, x, S; strategy, update_Σ = false): src/mfRG.jl:93-171.  One call: unflatten, symmetrise,
    iterate_solver!, residual and the DQGMRES preconditioning all stay on the device.  Returns (niter, solved) of the Krylov solve."""
    x = np.ascontiguousarray(x, dtype=np.complex128)
    out = np.empty_like(x)
    niter, solved = C.c_int(0), C.c_int(0)
    S._call("fdga_fixed_point_preconditioned", L.ptr(x), L.ptr(out), STRATEGY[strategy], int(use_preconditioner), int(krylov_maxiter), int(memory),
            C.byref(niter), C.byref(solved))
    R[: x.size] = out
    return niter.value, bool(solved.value)


def mix_bubbles(S, mixing):
    S._call("fdga_mix_bubbles", float(mixing))


def update_reference(S):
    S._call("fdga_update_reference")


def save_solver(S, filename, extra=None):
    """save!(f, "S", S): src/ParquetSolver.jl:309-330 (HDF5 in the reference's MeshFunction layout, fddgasolver.jl_b200/io.py)"""
    from . import io
    S.pull("Gbare", "G0", "Σ0", "G", "Σ", "F", "F0", "Π")
    io.save_solver(filename, S, extra=extra)


def load_solver(S, filename):
    """load_solver!(S, filename): src/nonlocal/ParquetSolver.jl:332-346"""
    from . import io
    for n in ("Π0pp", "Π0ph", "Πpp", "Πph"):
        if getattr(S, n) is None:
            setattr(S, n, zeros(S._shpΠ))
    io.load_solver(S, filename)
    S.push("Gbare", "G0", "Σ0", "G", "Σ", "F0", "F", "Π0pp", "Π0ph", "Πpp", "Πph")


def solve_using_mfRG(S, *, maxiter=100, verbose=False, occ_target=None, hubbard_params=None, mixing_init=1.0, tol=1e-4,
                     strategy="fdPA", anderson_iterations=40, anderson_m=50, krylov_maxiter=400, memory=100, debug_single_iter=False,
                     filename_log=None, iter_restart=0, auto_restart=False):
    """solve_using_mfRG!(S; maxiter, occ_target, hubbard_params, mixing_init, tol, strategy, filename_log, iter_restart,
    auto_restart): src/mfRG.jl:217-372 (adaptive mixing of the target bubble, vertex solve by Anderson iteration of the
    DQGMRES-preconditioned fixed point, SDE, reference update).  Everything between two vertex solves stays on the device.
    filename_log: after every accepted iteration the solver and the scalar `mixing` are written to `$filename_log.iter$i.h5`
    (:363-370); iter_restart / auto_restart resume from such a file (:240-275).  Returns a dict with the history (mixing, Σ error,
    μ per accepted iteration)."""
    from .nlsolve import anderson
    mixing, it = float(mixing_init), 0
    if auto_restart and filename_log is not None:
        from . import io
        _, last = io.last_checkpoint(filename_log, 100)
        if last:
            iter_restart = last
    if iter_restart:
        from . import h5min
        fn = f"{filename_log}.iter{iter_restart}.h5"
        load_solver(S, fn)
        mixing = float(h5min.File(fn)["mixing"].read())
        it = int(iter_restart)
    hist = {"mixing": [], "Σ_err": [], "μ": [], "anderson_iterations": [], "converged": False, "iterations": it}
    for _ in range(maxiter):
        it += 1
        mix_bubbles(S, mixing)                                                   # :271-276
        nF_ = S.length_F()

        def fp(x):
            R = np.empty(nF_, dtype=np.complex128)
            fixed_point_preconditioned(R, x, S, strategy=strategy, krylov_maxiter=krylov_maxiter, memory=memory)
            return R
        res = anderson(fp, S.flatten_F(), m=anderson_m, beta=0.85, ftol=tol, iterations=anderson_iterations, show_trace=verbose)   # :287-294
        hist["anderson_iterations"].append(res.iterations)
        if debug_single_iter:
            S.unflatten_F(res.zero)
            return hist
        if not res.f_converged:                                                  # :300-308
            mixing /= 2.0
            it -= 1
            bubbles(S)
            continue
        used = mixing
        mixing = min(1.0, mixing * 1.2)                                          # :310
        S.unflatten_F(res.zero)
        S._call("fdga_bubbles_real_space", 1)                                    # restore the unmixed bubbles for the SDE, :318-319
        S._call("fdga_bubbles_real_space", 0)
        SDE(S, "scPA")                                                           # :323
        S.pull("Σ", "Σ0")
        Σ_err = float(np.max(np.abs(S.Σ - S.Σ0))) / mixing                       # :324 (the reference divides by the UPDATED mixing)
        update_reference(S)                                                      # :328-347
        if occ_target is not None:                                               # :350-355
            μ = compute_hubbard_chemical_potential(occ_target, S, hubbard_params)
            set_hubbard_bare_Green(S, μ=μ, **hubbard_params)
            hist["μ"].append(μ)
        Dyson(S)                                                                 # :357-358
        bubbles(S)
        hist["mixing"].append(used)
        hist["Σ_err"].append(Σ_err)
        hist["iterations"] = it
        if filename_log is not None:                                             # :363-370
            save_solver(S, f"{filename_log}.iter{it}.h5", extra={"mixing": mixing})
        if Σ_err < tol:                                                          # :377-379
            hist["converged"] = True
            break
    return hist


def interpolate_vertex(So, Fi, which=0):
    """interpolate_vertex!(γo.K, γi.K) for the three channels and classes of an NL2 vertex (src/interpolate.jl:62-165, 199-206):
    Fourier interpolation of the host vertex Fi (its own momentum mesh and frequency boxes) into vertex `which` of So's context."""
    Li = Fi.L
    for ch in CHANNELS:
        g = Fi.channel(ch)
        for cls, (a, n) in enumerate(((g.K1, (Fi.numK1,)), (g.K2, Fi.numK2), (g.K3, Fi.numK3))):
            nk = (C.c_int32 * 2)(*(tuple(n) + (0,))[:2])
            So._call("fdga_interpolate_vertex", which, ch, cls, L.ptr(a), nk, Li)


def interpolate_solver(So, Si, *, occ_target=None, hubbard_params=None):
    """interpolate_solver!(So, Si; occ_target, hubbard_params): src/interpolate.jl:168-213.  Si only provides HOST arrays (Si.Σ,
    Si.F: pull them first if Si is a device solver); everything is evaluated in So's context."""
    nGi = Si.Σ.shape[0] // 2
    LGi = int(round(np.sqrt(Si.Σ.shape[1])))
    So._call("fdga_interpolate_green", L.SIGMA, L.ptr(np.asfortranarray(Si.Σ)), nGi, LGi, 1)
    if occ_target is not None:
        μ = compute_hubbard_chemical_potential(occ_target, So, hubbard_params)
        set_hubbard_bare_Green(So, μ=μ, **hubbard_params)
    Dyson(So)
    bubbles(So)
    interpolate_vertex(So, Si.F, 0)
    symmetrize_solver(So)
