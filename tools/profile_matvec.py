"""Per-kernel device times of one mfRG matvec / one DQGMRES iteration at a bench configuration (CUDA events, serialised lanes)."""
import sys, time, json
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np
import fddgasolver_jl_b200 as fd

nmax, nq = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4, 8)
S = fd.wu_point_solver(nmax=nmax, nq=nq, LG=48, F0_scale=0.02)
A = fd.mfRGLinearMap(S)
x = S.flatten_F()
for _ in range(3):
    A.matvec(x)
S.sync(); t0 = time.perf_counter()
for _ in range(10):
    A.matvec(x)
S.sync(); t_mv = (time.perf_counter() - t0) / 10
S.profile(True); S.profile_reset()
for _ in range(5):
    A.matvec(x)
kt = S.kernel_times(); S.profile(False)
print(json.dumps({"matvec_ms_e2e": round(t_mv * 1e3, 3), "kernels_ms": {k: round(v[0] / 5, 3) for k, v in kt.items() if v[1]}, "launches": {k: v[1] / 5 for k, v in kt.items() if v[1]}}))
for mem in (20, 100):
    fd.dqgmres(A, x, memory=mem, atol=0.0, rtol=0.0, itmax=3)
    S.sync(); t0 = time.perf_counter()
    _, st = fd.dqgmres(A, x, memory=mem, atol=0.0, rtol=0.0, itmax=60)
    S.sync(); dt = time.perf_counter() - t0
    S.profile(True); S.profile_reset()
    fd.dqgmres(A, x, memory=mem, atol=0.0, rtol=0.0, itmax=60)
    kt = S.kernel_times(); S.profile(False)
    print(json.dumps({"dqgmres_memory": mem, "ms_per_iteration": round(dt / st["niter"] * 1e3, 3), "krylov_ms_per_iteration_profiled": round(kt["krylov"][0] / 60, 3),
                      "krylov_launches_per_iteration": kt["krylov"][1] / 60}))
S.close()
