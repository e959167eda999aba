"""Generates tests/golden/*.npz with the CPU oracle (run here, on CPU; the vectors travel to the GPU box).

The reference (Julia) cannot be executed in this container, so these are ORACLE-generated vectors: they pin the
CUDA path to the oracle's numbers across rounds; the oracle itself is pinned by tests/test_oracle_*.py.
Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fddgasolver_jl_b200 as fd  # noqa: E402
import oracle as o  # noqa: E402


def main():
    seed = 1
    inp = fd.wu_point_inputs(2, 3, 6, seed=seed, F_scale=0.2, F0_scale=0.03, small_reference=True)
    R = o.OracleSolver(inp["nK1"], inp["nK2"], inp["nK3"], inp["L"], inp["Gbare"], inp["G0"], inp["Σ0"], inp["F0"], T=inp["T"])
    R.init_sym_grp()
    R.F.set(inp["F"])
    o.iterate_solver(R, "fdPA", True)
    np.savez_compressed(os.path.join(HERE, "nl2_fdpa_small.npz"), seed=seed, F=R.F.flatten(), Sigma=R.Σ.ravel(order="F"))
    print("wrote nl2_fdpa_small.npz", R.F.flatten().shape, np.abs(R.F.flatten()).max(), np.abs(R.Σ).max())


if __name__ == "__main__":
    main()
