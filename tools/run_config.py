import sys, time, json
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, torch
import fddgasolver_jl_b200 as fd
def run(nmax, nq, steps=3):
    t0=time.time()
    S = fd.wu_point_solver(nmax=nmax, nq=nq, LG=48, F0_scale=0.02)
    S.sync(); t1=time.time()
    S.stash_F()
    for _ in range(2):
        S.unstash_F(); fd.iterate_solver(S,'fdPA',update_Σ=False); fd.SDE(S,'scPA')
    S.sync(); t2=time.time()
    for _ in range(steps):
        S.unstash_F(); fd.iterate_solver(S,'fdPA',update_Σ=False); fd.SDE(S,'scPA')
    S.sync(); t3=time.time()
    S.profile(True); S.profile_reset()
    S.unstash_F(); fd.iterate_solver(S,'fdPA',update_Σ=False); fd.SDE(S,'scPA')
    kt=S.kernel_times(); S.profile(False)
    x = S.flatten_F(); S.pull("Σ")
    print(json.dumps(dict(nmax=nmax, nq=nq, setup_s=round(t1-t0,2), ms_per_step=round((t3-t2)/steps*1e3,2), lenF=int(x.size), classes=dict(K1=S.num_classes(1),K2pp=S.num_classes(2),K3pp=S.num_classes(4)),
          finite=bool(np.isfinite(x).all() and np.isfinite(S.Σ).all()), mem_GB=round(torch.cuda.mem_get_info()[1]/1e9-torch.cuda.mem_get_info()[0]/1e9,1),
          kernels={k:round(v[0],2) for k,v in kt.items() if v[1]})), flush=True)
    S.close()
for nmax,nq in [(6,8),(4,16),(8,8)]:
    run(nmax,nq)
