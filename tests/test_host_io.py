"""HDF5 reader / writer (fddgasolver.jl_b200/h5min.py, io.py) without libhdf5: the reference's packaged data files, the
MeshFunction / vertex / solver layouts of save! and load_solver!, round trips.  CPU only."""
import glob
import os
import types

import numpy as np
import pytest

import fddgasolver_jl_b200 as fd
from fddgasolver_jl_b200 import h5min, io

REF_DATA = "/root/reference/data"


def test_h5min_round_trip_of_every_supported_type(tmp_path):
    rng = np.random.default_rng(0)
    tree = {"f64": rng.standard_normal((3, 4)), "c128": rng.standard_normal((2, 3, 5)) + 1j * rng.standard_normal((2, 3, 5)),
            "i64": np.arange(7, dtype=np.int64), "scalar": h5min.DatasetSpec(np.float64(0.25), {"unit": "eV", "n": 3, "z": 1 - 2j, "v": np.arange(3.0)}),
            "empty": np.zeros((0,), dtype=np.float64),
            "grp": h5min.GroupSpec({"inner": {"x": np.float32(1.5) * np.ones(4, dtype=np.float32)}}, attrs={"type": "MeshFunction", "N": 12}),
            "many": {f"d{i:02d}": np.full(2, float(i)) for i in range(20)}}      # more than one symbol-table node
    p = str(tmp_path / "t.h5")
    h5min.write_file(p, tree)
    f = h5min.File(p)
    assert sorted(f.keys()) == sorted(tree)
    for k in ("f64", "c128", "i64", "empty"):
        got = f[k].read()
        assert got.dtype == np.asarray(tree[k]).dtype and np.array_equal(got, tree[k]), k
    assert f["scalar"].read() == 0.25 and f["scalar"].attrs["unit"] == "eV" and f["scalar"].attrs["n"] == 3
    assert f["scalar"].attrs["z"] == 1 - 2j and np.array_equal(f["scalar"].attrs["v"], np.arange(3.0))
    assert f["grp"].attrs == {"type": "MeshFunction", "N": 12}
    assert np.array_equal(f["grp/inner/x"].read(), 1.5 * np.ones(4, dtype=np.float32))
    assert sorted(f["many"].keys()) == [f"d{i:02d}" for i in range(20)]
    assert all(np.array_equal(f["many"][f"d{i:02d}"].read(), np.full(2, float(i))) for i in range(20))
    with pytest.raises(KeyError):
        f["nope"]
    open(str(tmp_path / "bad.h5"), "wb").write(b"not hdf5 at all")
    with pytest.raises(h5min.H5Error):
        h5min.File(str(tmp_path / "bad.h5"))


def test_vertex_round_trip_through_the_reference_layout(tmp_path):
    T = 0.3
    core = fd.RefVertex(T, 2.5, (3, 2), *[np.random.default_rng(i).standard_normal((5, 4, 4)) + 0j for i in range(4)])
    loc = fd.Vertex(core, T, 9, (5, 4), (1, 1))
    V = fd.NL2_Vertex(loc, T, 6, (3, 2), (2, 2), 4)
    fd.randomize_vertex(loc, 3, 1.0); fd.randomize_vertex(V, 4, 0.5)
    p = str(tmp_path / "v.h5")
    h5min.write_file(p, {"F": io.vertex_spec(V)})
    f = h5min.File(p)
    # the layout the reference reads: F/γp/K2 is a MeshFunction group, dimensions stored reversed, meshes tagged
    k2 = f["F/γp/K2"]
    assert k2.attrs["type"] == "MeshFunction" and k2["data"].shape == V.γp.K2.shape[::-1]
    m = k2["meshes/mesh_2"].attrs
    assert (m["tag"], m["type"], m["N"], m["temperature"]) == ("MatsubaraMesh", "Fermion", 2, T)
    assert f["F/F0/F0"].attrs["U"] == 2.5
    W = io.load_vertex(f["F"])
    assert isinstance(W, fd.NL2_Vertex) and W.L == 4 and isinstance(W.F0, fd.Vertex) and isinstance(W.F0.F0, fd.RefVertex)
    assert np.array_equal(W.flatten(), V.flatten()) and np.array_equal(W.F0.flatten(), loc.flatten())
    for a, b in zip(W.F0.F0.arrays(), core.arrays()):
        assert np.array_equal(a, b)


def test_swave_vertex_round_trip(tmp_path):
    """test/test_nonlocal_vertex.jl:154-166, test/test_nonlocal_solver.jl:47-62: save! / load_vertex(NL_Vertex, ...)"""
    T = 0.3
    V = fd.NL_Vertex(fd.RefVertex(T, 3.0), T, 10, (4, 3), (2, 1), 3)
    fd.randomize_vertex(V, 5, 0.5)
    p = str(tmp_path / "nl.h5")
    h5min.write_file(p, {"f": io.vertex_spec(V)})
    f = h5min.File(p)
    assert f["f/γt/K2/data"].shape == V.γt.K2.shape[::-1] and len(f["f/γt/K2/meshes"].keys()) == 3
    W = io.load_vertex(f["f"])
    assert isinstance(W, fd.NL_Vertex) and W.L == 3 and np.array_equal(W.flatten(), V.flatten())


def _fake_solver(seed):
    rng = np.random.default_rng(seed)
    T, L, LG, nG = 0.25, 3, 6, 4
    c = lambda *shp: np.asfortranarray(rng.standard_normal(shp) + 1j * rng.standard_normal(shp))
    S = types.SimpleNamespace(T=T, L=L, LG=LG, nG=nG)
    for n in ("Gbare", "G0", "Σ0", "G", "Σ"):
        setattr(S, n, c(2 * nG, LG * LG))
    for n in ("Π0pp", "Π0ph", "Πpp", "Πph"):
        setattr(S, n, c(2 * 4 - 1, 2 * 4, L * L, L * L))
    S.F0 = fd.NL2_Vertex(fd.RefVertex(T, 1.5), T, 4, (2, 2), (2, 2), L)
    S.F = fd.NL2_Vertex(S.F0, T, 4, (2, 2), (2, 2), L)
    fd.randomize_vertex(S.F0, seed + 1, 0.2); fd.randomize_vertex(S.F, seed + 2, 0.3)
    return S


def test_solver_checkpoint_round_trip_and_restart_scan(tmp_path):
    A, B = _fake_solver(1), _fake_solver(2)
    log = str(tmp_path / "run")
    assert io.last_checkpoint(log) == (None, 0)
    io.save_solver(log + ".iter1.h5", A, extra={"mixing": 0.35})
    io.save_solver(log + ".iter2.h5", A, extra={"mixing": 0.5})
    path, it = io.last_checkpoint(log)
    assert it == 2 and path.endswith(".iter2.h5")
    f = h5min.File(path)
    assert sorted(f.keys()) == sorted(["Gbare", "G0", "Σ0", "F0", "Π0pp", "Π0ph", "G", "Σ", "F", "Πpp", "Πph", "mixing"])   # src/ParquetSolver.jl:315-327
    assert f["mixing"].read() == 0.5
    io.load_solver(B, path)
    for n in ("Gbare", "G0", "Σ0", "G", "Σ", "Π0pp", "Π0ph", "Πpp", "Πph"):
        assert np.array_equal(getattr(A, n), getattr(B, n)), n
    assert np.array_equal(A.F.flatten(), B.F.flatten()) and np.array_equal(A.F0.flatten(), B.F0.flatten())


def test_committed_dmft_fixture():
    d = fd.synthetic.load_dmft_fixture()
    Γ = d["Γ"]
    assert (d["nG"], d["T"]) == (128, 0.2) and abs(d["occ"] - 0.47995214595886937) < 1e-15
    assert (Γ.numK1, Γ.numK2, Γ.numK3, Γ.F0.numK3) == (128, (74, 50), (1, 1), (24, 16)) and Γ.F0.U == 5.6
    assert d["params"]["U"] == 5.6 and d["params"]["t2"] == -0.3
    inp = fd.wu_point_inputs(4, 8, 48)
    assert inp["data"].startswith("reference file") and inp["F0"].F0 is not None
    assert np.array_equal(inp["G0"][:, 0], d["G"][112:144]) and np.array_equal(inp["Σ0"][:, 5], d["Σ"][112:144])


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference tree not present")
def test_reads_every_packaged_reference_file():
    """load_triqs_data on all data/*.h5 of the reference: mesh sizes against the MeshFunction attributes, U against params"""
    files = sorted(glob.glob(os.path.join(REF_DATA, "*.h5")))
    assert len(files) >= 30
    for p in files:
        d = io.load_triqs_data(p)
        Γ = d["Γ"]
        assert d["G"].shape == d["Σ"].shape == (2 * d["nG"],)
        assert Γ.γp.K1.shape == (2 * Γ.numK1 - 1,) and Γ.γa.K2.shape == (2 * Γ.numK2[0] - 1, 2 * Γ.numK2[1])
        assert Γ.F0.Fp_p.shape == (2 * Γ.F0.numK3[0] - 1, 2 * Γ.F0.numK3[1], 2 * Γ.F0.numK3[1])
        assert abs(Γ.F0.U - d["params"]["U"]) < 1e-12 and 0.0 < d["occ"] < 1.0
        assert abs(d["T"] - d["params"]["T"]) < 1e-5 * d["T"]          # Krien_point.h5: mesh temperature 0.149999 vs params T 0.15
    # the committed fixture is a faithful conversion of data/Wu_point.h5
    d, z = io.load_triqs_data(os.path.join(REF_DATA, "Wu_point.h5")), fd.synthetic.load_dmft_fixture()
    assert np.array_equal(d["G"], z["G"]) and np.array_equal(d["Γ"].flatten(), z["Γ"].flatten())
    assert all(np.array_equal(a, b) for a, b in zip(d["Γ"].F0.arrays(), z["Γ"].F0.arrays()))
