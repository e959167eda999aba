"""ctypes binding of libfdga.so (the C-ABI declared in include/fdga.h).

This is the same boundary a Julia ``ccall`` shim binds.  There is no CPU
fallback: if the shared library is missing, or no CUDA device is usable, everything that
touches the device raises ``FdgaError``.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FDGA_LIB_PATH", os.path.join(_HERE, "libfdga.so"))   # override only for A/B experiments

FDGA_MAX_LEVELS = 6
PCH, TCH, ACH = 0, 1, 2
K1, K2, K3 = 0, 1, 2
LV_NL2, LV_LOCAL, LV_CORE, LV_NL, LV_NL2_MBE, LV_LOCAL_MBE = 0, 1, 2, 3, 4, 5
V_FL, V_FBUFF = 100, 101
G, G0, GBARE, SIGMA, SIGMA0 = 0, 1, 2, 3, 4
PI0PP, PI0PH, PIPP, PIPH = 0, 1, 2, 3
SG_SIGMA, SG_K1, SG_PP2, SG_PH2, SG_PP3, SG_PH3, SG_PPL3, SG_PHL3 = range(8)
SG_NL_PP2, SG_NL_PH2 = 8, 9       # builder-only ids: K2[Ω, ν, P] groups of the s-wave solver
SCPA, FDPA, SCPA_NEW, FDPA_NEW, FDPA_1LOOP = 0, 1, 2, 3, 4
T_NAMES = ["cache", "L_K2", "L_K3", "K1", "K2", "K3", "sde_L", "sde_rs", "sde_U2", "bubble",
           "right", "swave", "expand", "misc", "comm", "column_K2", "krylov"]

# every symbol include/fdga.h declares (checked by tests/test_abi.py without a GPU)
EXPORTS = [
    "fdga_create", "fdga_destroy", "fdga_last_error", "fdga_sync", "fdga_set_option", "fdga_comm_unique_id", "fdga_comm_init", "fdga_partition",
    "fdga_set_vertex", "fdga_get_vertex", "fdga_set_core", "fdga_set_green", "fdga_get_green",
    "fdga_set_bubble", "fdga_get_bubble", "fdga_set_cache", "fdga_get_cache", "fdga_get_L",
    "fdga_set_symmetry_classes", "fdga_build_symmetry_group", "fdga_length_F", "fdga_flatten_F", "fdga_flatten_F_async",
    "fdga_unflatten_F", "fdga_unflatten_F_from_root", "fdga_stash_F", "fdga_unstash_F", "fdga_dyson", "fdga_occupation", "fdga_bubbles_real_space",
    "fdga_set_hubbard_bare_green", "fdga_hubbard_chemical_potential", "fdga_bubbles_momentum_space", "fdga_bubbles_local", "fdga_build_K3_cache", "fdga_bse_L_K2", "fdga_bse_L_K3", "fdga_bse_K1",
    "fdga_bse_K2", "fdga_bse_K3", "fdga_bse_K1_new", "fdga_bse_K2_new", "fdga_bse_K1_1loop", "fdga_bse_K2_1loop", "fdga_bse_K3_1loop",
    "fdga_set_F_from_Fbuff", "fdga_sde", "fdga_sde_channel_L", "fdga_iterate_solver",
    "fdga_mfrg_matvec", "fdga_mfrg_matvec_strategy", "fdga_mfrg_matvec_from_root", "fdga_mfrg_dqgmres", "fdga_symmetrize_solver", "fdga_fixed_point_preconditioned", "fdga_mix_bubbles", "fdga_update_reference", "fdga_interpolate_vertex", "fdga_interpolate_green",
    "fdga_measure_fp64_peak", "fdga_profile_enable", "fdga_profile_reset", "fdga_kernel_time_ms",
    "fdga_total_launches", "fdga_stream", "fdga_eval_vertex", "fdga_graph_begin", "fdga_graph_end", "fdga_graph_launch", "fdga_graph_destroy",
]


class FdgaError(RuntimeError):
    pass


class LevelDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("nK1", C.c_int32), ("nK2", C.c_int32 * 2), ("nK3", C.c_int32 * 2),
                ("U_re", C.c_double), ("U_im", C.c_double)]


class Dims(C.Structure):
    _fields_ = [("T", C.c_double), ("nq", C.c_int32), ("LG", C.c_int32), ("nG", C.c_int32),
                ("nPiB", C.c_int32), ("nPiF", C.c_int32), ("nlev", C.c_int32),
                ("lev", LevelDesc * FDGA_MAX_LEVELS)]


_lib = None


def load():
    """Load libfdga.so (raises FdgaError if it has not been built; see __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FdgaError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    lib.fdga_last_error.restype = C.c_char_p
    lib.fdga_last_error.argtypes = [vp]
    lib.fdga_create.argtypes = [C.POINTER(Dims), i32, C.POINTER(vp)]
    lib.fdga_destroy.argtypes = [vp]
    lib.fdga_sync.argtypes = [vp]
    lib.fdga_set_option.argtypes = [vp, i32, i32]
    lib.fdga_comm_unique_id.argtypes = [vp]
    lib.fdga_comm_init.argtypes = [vp, i32, i32, vp]
    lib.fdga_partition.argtypes = [i64, i32, i32, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    lib.fdga_set_vertex.argtypes = [vp, i32, i32, i32, vp, i64]
    lib.fdga_get_vertex.argtypes = [vp, i32, i32, i32, vp, i64]
    lib.fdga_set_core.argtypes = [vp, i32, i32, vp, i64]
    for n in ("green", "bubble", "cache"):
        getattr(lib, f"fdga_set_{n}").argtypes = [vp, i32, vp, i64]
        getattr(lib, f"fdga_get_{n}").argtypes = [vp, i32, vp, i64]
    lib.fdga_get_L.argtypes = [vp, i32, vp, i64]
    lib.fdga_set_symmetry_classes.argtypes = [vp, i32, i64, vp, vp, vp]
    lib.fdga_build_symmetry_group.argtypes = [i32, i32, i32, i32, vp, vp, vp, C.POINTER(i64)]
    lib.fdga_length_F.restype = i64
    lib.fdga_length_F.argtypes = [vp]
    lib.fdga_flatten_F.argtypes = [vp, vp]
    lib.fdga_flatten_F_async.argtypes = [vp, vp]
    lib.fdga_unflatten_F.argtypes = [vp, vp, dbl]
    lib.fdga_stash_F.argtypes = [vp]
    lib.fdga_unstash_F.argtypes = [vp]
    lib.fdga_dyson.argtypes = [vp]
    lib.fdga_occupation.argtypes = [vp, i32, C.POINTER(dbl)]
    lib.fdga_bubbles_real_space.argtypes = [vp, i32]
    lib.fdga_bubbles_momentum_space.argtypes = [vp, i32]
    lib.fdga_bubbles_local.argtypes = [vp, i32]
    lib.fdga_build_K3_cache.argtypes = [vp, i32, i32]
    lib.fdga_bse_L_K2.argtypes = [vp, i32]
    lib.fdga_bse_L_K3.argtypes = [vp, i32]
    lib.fdga_bse_K1.argtypes = [vp, i32, i32]
    lib.fdga_bse_K2.argtypes = [vp, i32, i32]
    lib.fdga_bse_K3.argtypes = [vp, i32, i32]
    for _n in ("fdga_bse_K1_new", "fdga_bse_K2_new", "fdga_bse_K1_1loop", "fdga_bse_K2_1loop", "fdga_bse_K3_1loop"):
        getattr(lib, _n).argtypes = [vp, i32, i32]
    lib.fdga_set_F_from_Fbuff.argtypes = [vp]
    lib.fdga_sde.argtypes = [vp, i32, i32, i32]
    lib.fdga_sde_channel_L.argtypes = [vp, i32, i32]
    lib.fdga_iterate_solver.argtypes = [vp, i32, i32, i32]
    lib.fdga_set_hubbard_bare_green.argtypes = [vp, dbl, dbl, dbl, dbl]
    lib.fdga_hubbard_chemical_potential.argtypes = [vp, dbl, dbl, dbl, dbl, C.POINTER(dbl)]
    lib.fdga_mfrg_matvec.argtypes = [vp, vp, vp, i32]
    lib.fdga_mfrg_matvec_strategy.argtypes = [vp, vp, vp, i32, i32]
    lib.fdga_mfrg_matvec_from_root.argtypes = [vp, vp, vp, i32, i32, i32]
    lib.fdga_mfrg_dqgmres.argtypes = [vp, vp, vp, i32, i32, dbl, dbl, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(dbl), i32]
    lib.fdga_symmetrize_solver.argtypes = [vp]
    lib.fdga_interpolate_vertex.argtypes = [vp, i32, i32, i32, vp, C.POINTER(C.c_int32), i32]
    lib.fdga_interpolate_green.argtypes = [vp, i32, vp, i32, i32, i32]
    lib.fdga_measure_fp64_peak.argtypes = [vp, C.POINTER(dbl)]
    lib.fdga_unflatten_F_from_root.argtypes = [vp, vp, dbl, i32]
    lib.fdga_mix_bubbles.argtypes = [vp, dbl]
    lib.fdga_update_reference.argtypes = [vp]
    lib.fdga_fixed_point_preconditioned.argtypes = [vp, vp, vp, i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    lib.fdga_profile_enable.argtypes = [vp, i32]
    lib.fdga_profile_reset.argtypes = [vp]
    lib.fdga_kernel_time_ms.argtypes = [vp, i32, C.POINTER(dbl), C.POINTER(i64)]
    lib.fdga_eval_vertex.argtypes = [vp, i32, i32, i32, i32, i32, i64, vp, vp, vp, vp, vp, vp, vp]
    lib.fdga_graph_begin.argtypes = [vp]
    lib.fdga_graph_end.argtypes = [vp, C.POINTER(C.c_int)]
    lib.fdga_graph_launch.argtypes = [vp, i32]
    lib.fdga_graph_destroy.argtypes = [vp, i32]
    lib.fdga_total_launches.restype = i64
    lib.fdga_total_launches.argtypes = [vp]
    lib.fdga_stream.restype = vp
    lib.fdga_stream.argtypes = [vp]
    _lib = lib
    return lib


def ptr(a):
    """Pointer to the memory of a complex128 / int64 / uint8 array (must be contiguous in F or C order)."""
    assert a.flags["F_CONTIGUOUS"] or a.flags["C_CONTIGUOUS"], "array must be contiguous"
    return C.c_void_p(a.ctypes.data)


def check(ctx, rc, what=""):
    if rc != 0:
        msg = load().fdga_last_error(ctx)
        raise FdgaError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")


def build_symmetry_group(which, n0, n1, nq, length):
    """Class tables of SymmetryGroup(symmetries, f) for the generator list `which` (host, integer only).

    Returns (offsets[ncls+1], index[length], ops[length]) as int64/int64/uint8 numpy arrays.
    """
    lib = load()
    offsets = np.zeros(length + 1, dtype=np.int64)
    index = np.zeros(length, dtype=np.int64)
    ops = np.zeros(length, dtype=np.uint8)
    ncls = C.c_int64(0)
    rc = lib.fdga_build_symmetry_group(which, n0, n1, nq, ptr(offsets), ptr(index), ptr(ops), C.byref(ncls))
    if rc != 0:
        raise FdgaError("fdga_build_symmetry_group failed")
    return offsets[: ncls.value + 1].copy(), index, ops


def partition(nclasses, nranks, rank):
    """(c0, c1, chunk): class representatives [c0, c1) computed by `rank`; `chunk` all-gather slots per rank"""
    c0, c1, ch = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    if load().fdga_partition(nclasses, nranks, rank, C.byref(c0), C.byref(c1), C.byref(ch)) != 0:
        raise FdgaError("fdga_partition: bad arguments")
    return c0.value, c1.value, ch.value


def trivial_symmetry_group(length):
    """SymmetryGroup(f): every element its own class (src/nonlocal_2/ParquetSolver.jl:124-132)."""
    return (np.arange(length + 1, dtype=np.int64), np.arange(length, dtype=np.int64), np.zeros(length, dtype=np.uint8))
