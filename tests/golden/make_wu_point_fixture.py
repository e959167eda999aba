"""Convert the reference's packaged DMFT data of the Wu point (data/Wu_point.h5, read by script/benchmark_Wu.jl:18 through
load_triqs_data) into the compact fixture tests/golden/wu_point_dmft.npz, using this package's own HDF5 reader
(fddgasolver.jl_b200/h5min.py + io.py).  Run in the build container, where /root/reference exists:

    python tests/golden/make_wu_point_fixture.py [/root/reference/data/Wu_point.h5]

The GPU box has no reference tree: bench.py and the tests read the .npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fddgasolver_jl_b200 import io  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data/Wu_point.h5"
d = io.load_triqs_data(src)
Γ = d["Γ"]
out = {"source": os.path.basename(src), "T": d["T"], "nG": d["nG"], "occ": d["occ"],
       "param_names": np.array(sorted(d["params"])), "param_values": np.array([d["params"][k] for k in sorted(d["params"])]),
       "G": d["G"], "G0": d["G0"], "Sigma": d["Σ"], "numK1": Γ.numK1, "numK2": np.array(Γ.numK2), "numK3": np.array(Γ.numK3),
       "core_N": np.array(Γ.F0.numK3), "U": Γ.F0.U}
for ch, g in zip("pta", Γ.channels()):
    out[f"K1_{ch}"], out[f"K2_{ch}"], out[f"K3_{ch}"] = g.K1, g.K2, g.K3
for n in ("Fp_p", "Fp_x", "Ft_p", "Ft_x"):
    out[n] = getattr(Γ.F0, n)
dst = os.path.join(ROOT, "tests", "golden", "wu_point_dmft.npz")
np.savez_compressed(dst, **out)
print(dst, os.path.getsize(dst), "bytes")
