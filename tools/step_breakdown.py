"""Step time (lanes on / off) vs the sum of the per-kernel device times at a given configuration: python tools/step_breakdown.py nmax nq"""
import sys, time, json
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np
import fddgasolver_jl_b200 as fd

nmax, nq = int(sys.argv[1]), int(sys.argv[2])
S = fd.wu_point_solver(nmax=nmax, nq=nq, LG=48, F0_scale=0.02)
S.stash_F()
def step():
    S.unstash_F(); fd.iterate_solver(S, "fdPA", update_Σ=False); fd.SDE(S, "scPA")
out = {"nmax": nmax, "nq": nq}
for serial in (2, 1):
    S.set_option("serial", serial)
    for _ in range(2):
        step()
    S.sync(); t0 = time.perf_counter()
    for _ in range(4):
        step()
    S.sync(); out["ms_per_step_%s" % {1: "one_stream", 2: "lanes"}[serial]] = round((time.perf_counter() - t0) / 4 * 1e3, 3)
S.set_option("serial", 0)
# host-side enqueue time (no synchronisation inside the loop): if it approaches the step time the CPU, not the GPU, is the limit
S.sync(); t0 = time.perf_counter()
for _ in range(10):
    step()
t_enq = (time.perf_counter() - t0) / 10
S.sync(); out["enqueue_ms_per_step"] = round(t_enq * 1e3, 3); out["ms_per_step_default"] = round((time.perf_counter() - t0) / 10 * 1e3, 3)
S.profile(True); S.profile_reset()
S.sync(); t0 = time.perf_counter()
step(); S.sync(); out["ms_profiled_step_wall"] = round((time.perf_counter() - t0) * 1e3, 3)
kt = S.kernel_times(); S.profile(False)
out["kernels_ms"] = {k: round(v[0], 3) for k, v in kt.items() if v[1]}
out["kernel_sum_ms"] = round(sum(v[0] for k, v in kt.items() if k != "column_K2"), 3)
print(json.dumps(out))
S.close()
