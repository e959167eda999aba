"""CUDA-graph replay of a step vs eager issue: identical state, step time of both (python tools/graph_check.py [nl_method])"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import fddgasolver_jl_b200 as fd

nl = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = fd.wu_point_solver(nmax=4, nq=8, LG=48, nl_method=nl, F0_scale=0.02)
x = S.F.flatten()
S.unflatten_F(x); S.stash_F()


def step():
    S.unstash_F()
    fd.iterate_solver(S, "fdPA", update_Σ=False)
    fd.SDE(S, "scPA")


def timeit(f, n=200):
    for _ in range(5):
        f()
    S.sync()
    t = time.perf_counter()
    for _ in range(n):
        f()
    S.sync()
    return (time.perf_counter() - t) / n * 1e3


for _ in range(3):
    step()
S.sync()
ya = S.flatten_F().copy(); S.pull("Σ"); sa = S.Σ.copy()
l0 = S.total_launches()
gid = S.record(step)
assert S.total_launches() == l0
S.replay(gid); S.sync()
per = S.total_launches() - l0
yb = S.flatten_F().copy(); S.pull("Σ"); sb = S.Σ.copy()
print("replay == eager:", np.array_equal(ya, yb), np.array_equal(sa, sb), "launches per step", per)
te = timeit(step)
tg = timeit(lambda: S.replay(gid))
print(f"nl_method {nl}: eager {te:.3f} ms/step ({1e3 / te:.0f} it/s), graph replay {tg:.3f} ms/step ({1e3 / tg:.0f} it/s)")
# a stale graph is refused
S.set_option("serial", 1)
try:
    S.replay(gid)
    print("ERROR: stale graph accepted")
except fd.FdgaError as e:
    print("stale graph refused:", str(e)[:90])
