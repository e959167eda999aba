"""Step time of the s-wave solver (NL_ParquetSolver, nl_method = 1 of script/run_Wu_point.jl) at BASELINE config 3 sizes:
python tools/swave_bench.py [steps]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import fddgasolver_jl_b200 as fd

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
S = fd.wu_point_solver(nmax=4, nq=8, LG=48, nl_method=1)
for _ in range(5):
    fd.iterate_solver(S, "fdPA")
S.sync()
t = time.perf_counter()
for _ in range(steps):
    fd.iterate_solver(S, "fdPA")
S.sync()
dt = (time.perf_counter() - t) / steps
print(f"s-wave solver, config 3: {dt * 1e3:.3f} ms per iterate_solver!(fdPA) = {1 / dt:.0f} iterations/s, launches/step {S.total_launches() / (steps + 5):.0f}")
A = fd.mfRGLinearMap(S)
x = S.F.flatten()
for _ in range(3):
    A.matvec(x)
t = time.perf_counter()
for _ in range(steps):
    A.matvec(x)
print(f"mfRG matvec (host vectors): {(time.perf_counter() - t) / steps * 1e3:.3f} ms")
S.profile(True); S.profile_reset()
for _ in range(20):
    fd.iterate_solver(S, "fdPA")
S.sync()
for k, (ms, n) in S.kernel_times().items():
    if n:
        print(f"  {k:10s} {ms / 20:8.4f} ms/step {n / 20:5.1f} launches/step")
