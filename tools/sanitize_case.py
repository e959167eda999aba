"""Small fdPA workload for compute-sanitizer (tools/sanitize.sh): fused iterations with the self-energy update (enough of them that
the automatic CUDA graphs record and replay), one mfRG matvec and a DQGMRES solve, concurrency lanes forced ON
(FDGA_OPT_SERIAL = 2), once per contraction kernel (q-lane / column) of the NL2 solver and once for the s-wave solver; then an
explicitly recorded step graph is replayed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fddgasolver_jl_b200 as fd  # noqa: E402

for qlane, nl in ((1, 2), (0, 2), (-1, 1)):
    S = fd.wu_point_solver(nmax=2, nq=4, LG=8, small_reference=True, F0_scale=0.02, F_scale=0.2, nl_method=nl)
    S.set_option("serial", 2)
    if nl == 2:
        S.set_option("qlane", qlane)
    for _ in range(4):
        fd.iterate_solver(S, "fdPA", True)
    A = fd.mfRGLinearMap(S)
    x = S.flatten_F()
    y = A.matvec(x)
    fd.dqgmres(A, y, memory=5, atol=0.0, rtol=0.0, itmax=3)
    fd.SDE(S, "scPA")
    S.pull("F", "Σ")
    step = lambda: (fd.iterate_solver(S, "fdPA", update_Σ=False), fd.SDE(S, "scPA"))
    step(); step()
    gid = S.record(step)
    S.replay(gid); S.replay(gid); S.sync()
    S.pull("F", "Σ")
    print("nl_method", nl, "qlane", qlane, "checksum", float(np.abs(S.F.flatten()).sum() + np.abs(S.Σ).sum()), "launches", S.total_launches(), flush=True)
    S.close()
