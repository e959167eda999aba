"""Replays a dump of the REAL reference (tools/reference_dump.jl, run where Julia + fdDGAsolver.jl are installed) through the CPU
oracle: inputs from the dump, outputs compared at 1e-10.  Skipped when no dump is present (the build container has no Julia);
when present it pins what DESIGN.md lists as "parity unpinned": SymmetryGroup class tables, bubbles_real_space! for L < LG,
the fd iteration, the SDE and the mfRG linear map."""
import os

import numpy as np
import pytest

DUMP = os.path.join(os.path.dirname(__file__), "golden", "julia")
_DT = {"ComplexF64": np.complex128, "Int64": np.int64, "UInt8": np.uint8, "Float64": np.float64}


def _load():
    man = os.path.join(DUMP, "manifest.txt")
    if not os.path.exists(man):
        pytest.skip("no reference dump (run tools/reference_dump.jl with the Julia reference installed)")
    arrs, scal = {}, None
    for line in open(man):
        if line.startswith("#"):
            scal = [float(x) for x in line.split("=")[1].split()]
            continue
        name, et, *dims = line.split()
        a = np.fromfile(os.path.join(DUMP, name + ".bin"), dtype=_DT[et])
        arrs[name] = a.reshape([int(d) for d in dims], order="F") if dims else a
    return arrs, scal


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(a)), np.max(np.abs(b)), 1e-300))


def test_oracle_against_reference_dump(orc):
    import fddgasolver_jl_b200 as fd
    d, sc = _load()
    T, U, nG, nK1 = sc[0], sc[1], int(sc[2]), int(sc[3])
    nK2, nK3, L, LG = (int(sc[4]), int(sc[5])), (int(sc[6]), int(sc[7])), int(sc[8]), int(sc[9])
    flat2 = lambda a: np.asfortranarray(a.reshape(a.shape[0], -1, order="F"))
    R = orc.OracleSolver(nK1, nK2, nK3, L, flat2(d["in_Gbare"]), flat2(d["in_G0"]), flat2(d["in_Sigma0"]), fd.RefVertex(T, U), T=T)
    R.init_sym_grp()
    names = ["SGsigma", "SGK1", "SGpp2", "SGph2", "SGpp3", "SGph3", "SGppL3", "SGphL3"]
    for w, n in enumerate(names):          # class tables: bit-exact, including the order of classes and members
        offs, idx, ops = R.sg[w]
        assert np.array_equal(offs, d[n + "_offsets"]) and np.array_equal(idx[: offs[-1]], d[n + "_index"]) and np.array_equal(ops[: offs[-1]], d[n + "_ops"]), n
    for g, c in zip(R.F.channels(), "pta"):
        for cls, a in zip(("K1", "K2", "K3"), g.arrays()):
            a[...] = d[f"in_F_{c}_{cls}"].reshape(a.shape, order="F")
    assert _rel(R.G, flat2(d["out_G"])) < 1e-10
    for n, a in (("out_Pi0pp", R.Π0pp), ("out_Pipp", R.Πpp), ("out_Piph", R.Πph)):
        assert _rel(a, d[n].reshape(a.shape, order="F")) < 1e-10, n
    orc.iterate_solver(R, "fdPA", True)
    for V, pre in ((R.F, "out_F"), (R.FL, "out_FL")):
        for g, c in zip(V.channels(), "pta"):
            for cls, a in zip(("K1", "K2", "K3"), g.arrays()):
                if pre == "out_FL" and cls == "K1":
                    continue
                assert _rel(a, d[f"{pre}_{c}_{cls}"].reshape(a.shape, order="F")) < 1e-10, (pre, c, cls)
    assert _rel(R.Σ, flat2(d["out_Sigma"])) < 1e-10
    y = orc.mfRGLinearMap(R).matvec(d["mfrg_x"])
    assert _rel(y, d["mfrg_y"]) < 1e-10
