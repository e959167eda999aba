"""How long does the HOST take to issue one bench step (all launches, no sync) compared with the device time of the step?"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fddgasolver_jl_b200 as fd

S = fd.wu_point_solver(4, 8, 48, F0_scale=0.02)
S.stash_F()
def step():
    S.unstash_F(); fd.iterate_solver(S, "fdPA", update_Σ=False); fd.SDE(S, "scPA")
for _ in range(10): step()
S.sync()
K = int(os.environ.get('K', '200'))
t0 = time.perf_counter()
for _ in range(K): step()
t1 = time.perf_counter()
S.sync()
t2 = time.perf_counter()
print(f"host issue {1e3*(t1-t0)/K:.3f} ms/step, until device done {1e3*(t2-t0)/K:.3f} ms/step, launches/step {S.total_launches()/(K+10):.1f}")
S.set_option("serial", 1)
for _ in range(10): step()
S.sync()
t0 = time.perf_counter()
for _ in range(K): step()
t1 = time.perf_counter(); S.sync(); t2 = time.perf_counter()
print(f"one stream: host issue {1e3*(t1-t0)/K:.3f} ms/step, until device done {1e3*(t2-t0)/K:.3f} ms/step")
