"""HDF5 input / output in the reference's on-disk layout (pure Python, fddgasolver.jl_b200/h5min.py).

  load_triqs_data(filename)            src/utility/load_triqs.jl:298-308   the packaged DMFT / DCA data (data/*.h5)
  save_vertex / load_vertex            src/vertex.jl:380-420, src/channel.jl:377-402, src/refvertex.jl:219-250
  save_solver(filename, S)             save!(f, label, S)                   src/ParquetSolver.jl:309-330
  load_solver(S, filename)             load_solver!(S, filename)            src/nonlocal/ParquetSolver.jl:332-346
  last_checkpoint(filename_log)        restart scan of solve_using_mfRG!    src/mfRG.jl:240-275

MeshFunction layout (decoded from the reference's data files): a group with the attribute type = "MeshFunction", a sub-group
meshes/mesh_i per axis with the attributes tag, type, temperature, N for a MatsubaraMesh, and the dataset `data` of the compound
{r: f64, i: f64}; HDF5 lists the dimensions of a Julia array in reverse order.  Brillouin-zone meshes do not occur in the packaged
files and MatsubaraFunctions.jl is not vendored: they are written as tag = "BrillouinZoneMesh" with the attribute L (that part of
the layout is this package's own, to be aligned with MatsubaraFunctions.jl where Julia is available).
"""
import os

import numpy as np

from . import h5min
from .types import NL2_Vertex, NL_Vertex, RefVertex, Vertex


def _mats_mesh(kind, T, N):
    return {"tag": "MatsubaraMesh", "type": kind, "temperature": float(T), "N": int(N)}


def _bz_mesh(L):
    return {"tag": "BrillouinZoneMesh", "L": int(L)}


def mesh_function_spec(data, meshes):
    """GroupSpec of one MeshFunction: `data` in Julia (column-major) axis order, one mesh dict per axis"""
    data = np.asarray(data, dtype=np.complex128)
    assert data.ndim == len(meshes)
    ms = h5min.GroupSpec({f"mesh_{i + 1}": h5min.GroupSpec({}, attrs=m) for i, m in enumerate(meshes)})
    stored = np.ascontiguousarray(np.transpose(data))          # HDF5 dimension order = reversed Julia order
    return h5min.GroupSpec({"meshes": ms, "data": h5min.DatasetSpec(stored)}, attrs={"type": "MeshFunction"})


def load_mesh_function(group):
    """(data as a Fortran-ordered complex array in Julia axis order, [mesh attribute dicts in axis order])"""
    if group.attrs.get("type") != "MeshFunction":
        raise h5min.H5Error(f"{group.name}: not a MeshFunction group")
    raw = group["data"].read()
    data = np.asfortranarray(np.transpose(raw)).astype(np.complex128)
    mg = group["meshes"]
    meshes = [dict(mg[f"mesh_{i + 1}"].attrs) for i in range(data.ndim)]
    for ax, m in enumerate(meshes):
        if m.get("tag") == "MatsubaraMesh":
            want = 2 * m["N"] - 1 if m["type"] == "Boson" else 2 * m["N"]
            if data.shape[ax] != want:
                raise h5min.H5Error(f"{group.name}: axis {ax + 1} has {data.shape[ax]} points, mesh says {want}")
    return data, meshes


# ---- vertices -------------------------------------------------------------------------------------------------------------
def _channel_spec(g, nonlocal_L=None):
    T = g.T
    bz = [] if nonlocal_L is None else [_bz_mesh(nonlocal_L)]
    bz2 = bz * (g.K2.ndim - 2)          # NL2_Channel: K2[Ω,ν,P,k]; NL_Channel: K2[Ω,ν,P]
    return h5min.GroupSpec({
        "K1": mesh_function_spec(g.K1, [_mats_mesh("Boson", T, g.numK1)] + bz),
        "K2": mesh_function_spec(g.K2, [_mats_mesh("Boson", T, g.numK2[0]), _mats_mesh("Fermion", T, g.numK2[1])] + bz2),
        "K3": mesh_function_spec(g.K3, [_mats_mesh("Boson", T, g.numK3[0]), _mats_mesh("Fermion", T, g.numK3[1]),
                                        _mats_mesh("Fermion", T, g.numK3[1])] + bz)})


def vertex_spec(F):
    """save!(file, label, F) for RefVertex / Vertex / NL2_Vertex / NL_Vertex (recursively through F.F0)"""
    if isinstance(F, RefVertex):
        ms = [_mats_mesh("Boson", F.T, F.numK3[0]), _mats_mesh("Fermion", F.T, F.numK3[1]), _mats_mesh("Fermion", F.T, F.numK3[1])]
        return h5min.GroupSpec({n: mesh_function_spec(getattr(F, n), ms) for n in ("Fp_p", "Fp_x", "Ft_p", "Ft_x")},
                               attrs={"U": np.complex128(F.U)})
    L = getattr(F, "L", None) if isinstance(F, (NL2_Vertex, NL_Vertex)) else None
    return h5min.GroupSpec({"F0": vertex_spec(F.F0), "γp": _channel_spec(F.γp, L), "γt": _channel_spec(F.γt, L), "γa": _channel_spec(F.γa, L)})


def _load_channel_into(g, group):
    for n in ("K1", "K2", "K3"):
        data, _ = load_mesh_function(group[n])
        dst = getattr(g, n)
        if dst.shape != data.shape:
            raise h5min.H5Error(f"{group.name}/{n}: shape {data.shape} on disk, {dst.shape} expected")
        dst[...] = data


def load_refvertex(group):
    U = group.attrs["U"]
    arrs, N = {}, None
    for n in ("Fp_p", "Fp_x", "Ft_p", "Ft_x"):
        arrs[n], meshes = load_mesh_function(group[n])
        N = (meshes[0]["N"], meshes[1]["N"])
        T = meshes[0]["temperature"]
    return RefVertex(T, complex(U), N, arrs["Fp_p"], arrs["Fp_x"], arrs["Ft_p"], arrs["Ft_x"])


def load_vertex(group):
    """load_vertex(T, file, label): the vertex type is read off the stored meshes (local Vertex / NL2_Vertex / RefVertex)"""
    if "Fp_p" in group.keys():
        return load_refvertex(group)
    F0 = load_vertex(group["F0"])
    _, m1 = load_mesh_function(group["γp/K1"])
    _, m2 = load_mesh_function(group["γp/K2"])
    _, m3 = load_mesh_function(group["γp/K3"])
    T = m1[0]["temperature"]
    n1, n2, n3 = m1[0]["N"], (m2[0]["N"], m2[1]["N"]), (m3[0]["N"], m3[1]["N"])
    if len(m1) == 1:
        F = Vertex(F0, T, n1, n2, n3)
    elif len(m2) == 4:
        F = NL2_Vertex(F0, T, n1, n2, n3, m1[1]["L"])
    elif len(m2) == 3:
        F = NL_Vertex(F0, T, n1, n2, n3, m1[1]["L"])
    else:
        raise h5min.H5Error(f"{group.name}: vertex type with a {len(m2)}-axis K2 is outside this package's scope (NL3 / MBE)")
    for n in ("γp", "γt", "γa"):
        _load_channel_into(getattr(F, n), group[n])
    return F


def load_triqs_data(filename):
    """load_triqs_data(filename) of src/utility/load_triqs.jl:298-308: dict with G, G0, Σ (arrays on the fermionic mesh, stored
    as i G / i Σ like everything in the reference), their mesh size nG, Γ (local Vertex with its RefVertex core), occ, params"""
    f = h5min.File(filename)
    out = {}
    for n in ("G", "G0", "Σ"):
        out[n], meshes = load_mesh_function(f[n])
        out["nG"], out["T"] = meshes[0]["N"], meshes[0]["temperature"]
    out["Γ"] = load_vertex(f["Γ"])
    out["occ"] = float(f["occ"].read())
    out["params"] = {k: float(f["params"][k].read()) for k in f["params"].keys()}
    return out


# ---- solver checkpoints ------------------------------------------------------------------------------------------------------
_G_NAMES = ("Gbare", "G0", "Σ0", "G", "Σ")
_PI_NAMES = ("Π0pp", "Π0ph", "Πpp", "Πph")


def solver_spec(S):
    """tree written by save!(f, label, S): Gbare, G0, Σ0, F0, Π0pp, Π0ph, G, Σ, F, Πpp, Πph at the top level (the label is
    ignored by the reference too, src/ParquetSolver.jl:315-327)"""
    T, L, LG = S.T, S.L, S.LG
    tree = h5min.GroupSpec({})
    for n in _G_NAMES:
        tree[n] = mesh_function_spec(getattr(S, n), [_mats_mesh("Fermion", T, S.nG), _bz_mesh(LG)])
    for n in _PI_NAMES:
        a = getattr(S, n)
        tree[n] = mesh_function_spec(a, [_mats_mesh("Boson", T, (a.shape[0] + 1) // 2), _mats_mesh("Fermion", T, a.shape[1] // 2)] + [_bz_mesh(L)] * (a.ndim - 2))
    tree["F0"] = vertex_spec(S.F0)
    tree["F"] = vertex_spec(S.F)
    return tree


def save_solver(filename, S, extra=None):
    """save!(f, "S", S) (+ optional scalars such as `mixing`, src/mfRG.jl:363-370).  S: an NL2_ParquetSolver whose host mirrors are
    current (S.pull("F", "Σ", "G", "Π") first), or any object with the same attributes."""
    tree = solver_spec(S)
    for k, v in (extra or {}).items():
        tree[k] = h5min.DatasetSpec(np.asarray(v))
    h5min.write_file(filename, tree)


def load_solver(S, filename):
    """load_solver!(S, filename): copies the stored arrays into S's host mirrors (shapes must match); returns the names loaded.
    The caller pushes them to the device (S.push(...))."""
    f = h5min.File(filename)
    for n in _G_NAMES + _PI_NAMES:
        data, _ = load_mesh_function(f[n])
        dst = getattr(S, n)
        if dst is None:
            setattr(S, n, data)
            continue
        if dst.shape != data.shape:
            raise h5min.H5Error(f"{n}: shape {data.shape} on disk, {dst.shape} in the solver")
        dst[...] = data
    for name in ("F0", "F"):
        V = load_vertex(f[name])
        dst = getattr(S, name)
        lvl_src, lvl_dst = V, dst
        while True:       # load_vertex!(S.F0, f, "F0"): set! level by level
            if isinstance(lvl_dst, RefVertex):
                for a, b in zip(lvl_dst.arrays(), lvl_src.arrays()):
                    a[...] = b
                lvl_dst.U = lvl_src.U
                break
            lvl_dst.set(lvl_src)
            lvl_src, lvl_dst = lvl_src.F0, lvl_dst.F0
    return _G_NAMES + _PI_NAMES + ("F0", "F")


def last_checkpoint(filename_log, maxiter=10000):
    """restart scan of solve_using_mfRG! (src/mfRG.jl:240-275): the last existing `$filename_log.iter$i.h5`; (None, 0) if none"""
    last, it = None, 0
    for i in range(1, maxiter + 1):
        p = f"{filename_log}.iter{i}.h5"
        if not os.path.exists(p):
            break
        last, it = p, i
    return last, it
