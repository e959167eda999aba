#!/usr/bin/env python
"""bench.py -- parquet BSE iterations/sec at the nonlocal Wu point (BASELINE.json metric).

One "step" = one pass of the hot path over one synthetic vertex:
    iterate_solver!(S; strategy = :fdPA, update_Σ = false)   (script/benchmark_Wu.jl:53, src/solve.jl:4-116)
  + SDE!(S; strategy = :scPA)                                 (src/mfRG.jl:335)
on an NL2_ParquetSolver with the grid sizes of script/benchmark_Wu.jl:78 (nmax = 4, nq = 8, LG = 48; "config 3").

  python bench.py --gpus N --steps K --warmup W            our arm (libfdga, CUDA sm_100a)
  python bench.py --impl reference ...                     CPU arm: the oracle restatement of the reference's
                                                           algorithm on the host cores (Julia is not installed)
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "parquet_bse_iterations_per_sec"
UNIT = "iterations/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--nmax", type=int, default=4)
    p.add_argument("--nq", type=int, default=8)
    p.add_argument("--LG", type=int, default=48)
    p.add_argument("--nl-method", type=int, default=2, choices=[1, 2, -2],
                   help="2 (default, the headline): NL2_ParquetSolver; 1: the s-wave NL_ParquetSolver of script/run_Wu_point.jl (side workload); "
                        "-2: NL2_ParquetSolver with multi-boson-exchange vertices (generic kernels; at most 10 steps, no CPU arm, no mfRG extras)")
    p.add_argument("--no-graph", action="store_true", help="issue every step eagerly instead of replaying its CUDA graph (1 GPU)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-extras", action="store_true", help="skip the mfRG matvec / DQGMRES side measurements (large sweep sizes)")
    a = p.parse_args()
    if a.nl_method == -2:
        a.steps, a.warmup, a.no_cpu_baseline, a.no_extras, a.no_graph = min(a.steps, 10), min(a.warmup, 3), True, True, True
    return a


def workload_name(a):
    if a.nl_method == -2:
        return (f"NL2_ParquetSolver with NL2_MBEVertex, Wu point U=5.6 T=0.2 (script/run_Wu_point.jl, nl_method=-2): nmax={a.nmax} (nK1={4 * a.nmax}, "
                f"nK2=nK3=({a.nmax},{a.nmax})), nq={a.nq}, LG={a.LG}; generic per-term kernels")
    if a.nl_method == 1:
        return (f"s-wave NL_ParquetSolver, Wu point U=5.6 T=0.2 (script/run_Wu_point.jl, nl_method=1): nmax={a.nmax} (nK1={4 * a.nmax}, "
                f"nK2=nK3=({a.nmax},{a.nmax})), nq={a.nq}, LG={a.LG}, bubble mesh {4 * a.nmax} x {128 * a.nmax} (m_Pi_nu_factor=32)")
    return f"NL2 Wu point U=5.6 T=0.2 (script/benchmark_Wu.jl): nmax={a.nmax} (nK1={4 * a.nmax}, nK2=nK3=({a.nmax},{a.nmax})), nq={a.nq}, LG={a.LG}"


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU arm
def host_threads(o):
    """Use every host core whatever the launcher exported: torch.distributed.run sets OMP_NUM_THREADS=1 for its workers, which
    silently pinned the CPU arm to one core in round 1."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    o.lib().orc_set_num_threads(int(n))
    used = int(o.lib().orc_num_threads())
    if used == 1 and n > 1:
        raise SystemExit(f"bench.py: the CPU arm would run on 1 of {n} host cores; refusing to print a line")
    return used


def cpu_full_iteration(o, R, F_start):
    """ONE complete, un-sampled pass of the hot path on the host: iterate_solver!(fdPA, update_Σ = false) + SDE!(scPA) of the
    oracle restatement (reference loop structure, OpenMP over class representatives = the reference's Threads.@threads over the
    symmetry classes).  Returns the measured wall time and its split."""
    R.F.set(F_start)
    t0 = time.perf_counter()
    o.iterate_solver(R, "fdPA", False)
    t1 = time.perf_counter()
    o.SDE(R, "scPA")
    t2 = time.perf_counter()
    return t2 - t0, {"iterate_solver": round(t1 - t0, 3), "SDE": round(t2 - t1, 3)}


def make_oracle_solver(o, inp, share_bubbles_from=None):
    cls = o.OracleNLSolver if inp["F"].γp.K2.ndim == 3 else o.OracleSolver
    R = cls(inp["nK1"], inp["nK2"], inp["nK3"], inp["L"], inp["Gbare"], inp["G0"], inp["Σ0"], inp["F0"], T=inp["T"],
                       compute_bubbles=share_bubbles_from is None)
    if share_bubbles_from is not None:
        S = share_bubbles_from
        S.pull("Π", "G")
        R.Π0pp, R.Π0ph, R.Πpp, R.Πph, R.G = S.Π0pp, S.Π0ph, S.Πpp, S.Πph, S.G.copy(order="F")
    R.init_sym_grp()
    R.F.set(inp["F"])
    return R


def run_reference(a):
    """--impl reference: the reference's CPU path (its restatement under oracle/; Julia + MatsubaraFunctions.jl are not installed on
    any box) on all host cores.  Every timed step is one COMPLETE iteration, nothing is sampled or extrapolated; because a step
    takes seconds, the number of timed steps is capped by a wall-clock budget and the line reports the steps actually run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if a.nl_method == -2:
        raise SystemExit("bench.py --impl reference: the MBE side workload has no CPU arm (its inputs are converted on the device and one "
                         "oracle iteration at config 3 takes hours); use --nl-method 2 or 1")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fddgasolver_jl_b200 as fd
    import oracle as o
    o.build()
    cores = host_threads(o)
    inp = fd.wu_point_inputs(a.nmax, a.nq, a.LG, F0_scale=0.02, nl_method=a.nl_method)
    R = make_oracle_solver(o, inp)
    budget = float(os.environ.get("FDGA_REF_BUDGET_S", "150"))
    t_warm, _ = cpu_full_iteration(o, R, inp["F"]) if a.warmup > 0 else (0.0, None)      # one full warm-up pass (threads, page faults)
    ts, parts = [], None
    while len(ts) < max(a.steps, 1) and (not ts or sum(ts) + ts[-1] <= budget):
        t, parts = cpu_full_iteration(o, R, inp["F"])
        ts.append(t)
    sec = float(np.mean(ts))
    val = 1.0 / sec
    sample = (f"{len(ts)} complete un-sampled iteration(s) (iterate_solver!(fdPA, update_Σ=false) + SDE!(scPA)), measured wall time; "
              f"{a.steps} requested, capped by a {budget:.0f} s budget")
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": len(ts), "steps_requested": a.steps,
           "warmup": 1 if a.warmup > 0 else 0, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64 (complex)", "data": inp["data"],
           "config": {"workload": workload_name(a), "timing": "host wall clock of complete iterations of the CPU restatement (oracle/) of the reference algorithm; Julia + MatsubaraFunctions.jl are not installed"},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "seconds_per_step": [round(t, 3) for t in ts], "parts_s": parts, "timed_wall_s": round(sum(ts), 3), "warmup_wall_s": round(t_warm, 3)}
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------- our arm
def run_ours(a):
    import torch
    import fddgasolver_jl_b200 as fd
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        # NCCL's communicator lines go to stderr (stdout carries the one JSON line): the driver counts the ranks from them
        os.environ.setdefault("NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    inp = fd.wu_point_inputs(a.nmax, a.nq, a.LG, F0_scale=0.02, nl_method=a.nl_method)
    Solver = fd.NL_ParquetSolver if a.nl_method == 1 else fd.NL2_ParquetSolver
    skw = dict(VT=fd.NL2_MBEVertex) if a.nl_method == -2 else {}
    S = Solver(inp["nK1"], inp["nK2"], inp["nK3"], inp["L"], inp["Gbare"], inp["G0"], inp["Σ0"], inp["F0"], T=inp["T"], device=local, **skw)
    S.F.set(inp["F"]); S.push("F"); S.init_sym_grp()
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(S.comm_unique_id())
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        S.comm_init(world, rank, uid.cpu().numpy())
    stream = torch.cuda.ExternalStream(S.stream())
    nF = S.length_F()
    x_host = torch.empty(nF, dtype=torch.complex128).pin_memory()
    y_host = torch.empty(nF, dtype=torch.complex128).pin_memory()
    s_host = torch.empty(S.Σ.size, dtype=torch.complex128).pin_memory()
    x_np, y_np, s_np = x_host.numpy(), y_host.numpy(), s_host.numpy()
    x_np[:] = inp["F"].flatten()

    def barrier():
        torch.cuda.synchronize(); S.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(); S.sync()

    def step_resident():
        S.unstash_F()
        fd.iterate_solver(S, "fdPA", update_Σ=False)
        fd.SDE(S, "scPA")

    def e2e_device_part():
        if world == 1:
            S.unflatten_F_async(x_np)                   # H2D of this step's input vertex
        else:
            S.unflatten_F_from_root(x_np, 1.0, 0, rank)  # one PCIe upload on rank 0, NVLink broadcast to the other ranks
        fd.iterate_solver(S, "fdPA", update_Σ=False)
        if rank == 0:
            S.flatten_F_async(y_np)                     # D2H of the updated vertex (the job's result leaves through rank 0),
        fd.SDE(S, "scPA")                               # ... overlapping the SDE

    def step_e2e():
        e2e_device_part()
        if rank == 0:
            S.get_green_into("Σ", s_np)                 # D2H of the self-energy
        S.sync()                                        # both copies have landed

    # one GPU: a step is ~80 small dependent kernels on three lanes; it is recorded once as a CUDA graph (fdga_graph_begin / _end)
    # and every timed step is one replay of it -- same kernels, same order, bit-identical state (tools/graph_check.py)
    use_graph = (world == 1 or os.environ.get("FDGA_GRAPH_MULTIRANK", "0") == "1") and not a.no_graph
    step_resident_eager, step_e2e_eager = step_resident, step_e2e

    def make_graph_steps():
        step_resident_eager(); step_resident_eager()
        g1 = S.record(step_resident_eager)
        step_e2e_eager(); step_e2e_eager()
        g2 = S.record(e2e_device_part)

        def e2e_graph():
            S.replay(g2)
            S.get_green_into("Σ", s_np)
            S.sync()
        return (lambda: S.replay(g1)), e2e_graph

    def timed(step, K, W):
        for _ in range(W):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = S.total_launches()
        e0.record(stream)
        for _ in range(K):
            step()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, S.total_launches() - n0

    S.unflatten_F(x_np); S.stash_F()
    if use_graph:
        try:
            step_resident, step_e2e = make_graph_steps()
        except Exception as e:      # never lose the measurement to the replay path: fall back to eager issue and say so
            print(f"bench.py: CUDA-graph recording failed ({e}); issuing eagerly", file=sys.stderr)
            use_graph = False
            step_resident, step_e2e = step_resident_eager, step_e2e_eager
    W = max(a.warmup, 3)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, launches = timed(step_resident, a.steps, W)
    ms_e2e, _ = timed(step_e2e, a.steps, 1)
    clk = clocks.stop() if rank == 0 else None
    value = a.steps / (ms * 1e-3)
    e2e = a.steps / (ms_e2e * 1e-3)

    # per-kernel device times (CUDA events on the library's stream) over a few profiled steps
    S.profile(True); S.profile_reset()
    PS = min(a.steps, 5)
    for _ in range(PS):
        step_resident_eager()
    kt = S.kernel_times()
    S.profile(False)
    kernels = {k: {"ms_per_step": v[0] / PS, "launches_per_step": v[1] / PS} for k, v in kt.items() if v[1]}

    # roofline of the dominant kernel: the contraction kernel of BSE_K2! (qlane_kernel<JOB_K2>; column_kernel on small meshes), one
    # launch per channel, timed alone with CUDA events on the library's stream
    n2cls = [S.num_classes(fd._lib.SG_PP2), S.num_classes(fd._lib.SG_PH2), S.num_classes(fd._lib.SG_PH2)]
    nB2, nF2, NP, nFΠ = 2 * S.nK2[0] - 1, 2 * S.nK2[1], S.NP, 2 * S.nΠF
    chunk = [(-(-n // world)) for n in n2cls]           # representatives per rank
    tables = sum(sum(arr.size for arr in V.γp.arrays()) * 3 * 16 for V in fd.vertex_chain(S.F)[:-1]) + 4 * fd.vertex_chain(S.F)[-1].Fp_p.size * 16
    bytes_per_launch = nB2 * NP * nFΠ * NP * 16 / world + tables + np.mean(chunk) * 16      # R slabs of the K2 bosonic box + vertex tables + outputs
    flop_per_launch = 26.0 * np.mean(chunk) * nFΠ * NP                                       # 26 flop per (representative, w, q) term
    k2 = kernels.get("column_K2", {"ms_per_step": float("nan"), "launches_per_step": 3})
    if a.nl_method == -2:       # MBE: the generic kernel of BSE_K2! (no column sub-timer)
        k2 = kernels.get("K2", k2)
    if a.nl_method == 1:        # s-wave solver: no inner momentum sum; one launch reads its representatives' two bubble rows + the tables
        bytes_per_launch = np.mean(chunk) * nFΠ * 2 * 16 + tables + np.mean(chunk) * 16
        flop_per_launch = 40.0 * np.mean(chunk) * nFΠ
        k2 = kernels.get("K2", k2)
    k2_ms_launch = k2["ms_per_step"] / max(k2["launches_per_step"], 1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_hbm = float(peaks.get("hbm_gbs", 6650.0))
    fp64_peak = S.measure_fp64_peak()                     # DFMA micro-benchmark, measured live on this device
    traffic = None                                        # measured DRAM bytes per launch, from the committed ncu capture of this workload
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"nmax{a.nmax}_nq{a.nq}_LG{a.LG}_world{world}")
        traffic = None if tr is None else {"dram_bytes_per_launch": tr["dram_bytes_per_launch"], "l2_to_l1_bytes_per_launch": tr.get("l2_to_l1_bytes_per_launch"), "source": tr["source"]}
    except Exception:
        pass
    ach = bytes_per_launch / (k2_ms_launch * 1e-3) / 1e9
    roofline = {"kernel": ("sw_bse_k2_kernel (BSE_K2!, src/nonlocal/BSEa/BSEa_K2.jl:44-106), mean over the p, t, a launches" if a.nl_method == 1 else
                           "qlane_kernel<JOB_K2> (BSE_K2!, src/nonlocal_2/BSEa/BSEa_K2.jl:55-138), mean over the p, t, a launches"), "bound": "hbm", "achieved": ach, "peak": peak_hbm,
                "unit": "GB/s", "frac": ach / peak_hbm,
                "traffic": None if traffic is None else traffic["dram_bytes_per_launch"], "traffic_detail": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": bytes_per_launch, "ms_per_launch": k2_ms_launch,
                "note": "latency-bound: ~500 class representatives x 1024 inner frequencies per launch, a few MB of L2-resident tables" if a.nl_method == 1 else "not HBM-bound: a gather contraction whose tables and slabs are L2-resident (DRAM traffic below the algorithmic bytes); bounded by L2 -> L1 latency at the occupancy its registers allow, see fp64 and DESIGN.md section 4",
                "l2": None if (traffic is None or not traffic.get("l2_to_l1_bytes_per_launch")) else {
                    "bytes_per_launch": traffic["l2_to_l1_bytes_per_launch"], "achieved_tbs": traffic["l2_to_l1_bytes_per_launch"] / (k2_ms_launch * 1e-3) / 1e12,
                    "peak_tbs": 6300 * 1.965e9 / 1e12, "frac": traffic["l2_to_l1_bytes_per_launch"] / (k2_ms_launch * 1e-3) / (6300 * 1.965e9),
                    "peak_source": "LTS throughput cap ~6300 B/clk (B300_MICROARCH.md, measured on B300) x 1.965 GHz; bytes = lts__t_sectors_srcunit_tex_op_read x 32 from the committed ncu capture"},
                "fp64": {"algorithmic_gflop_per_launch": flop_per_launch / 1e9, "achieved_tflops": flop_per_launch / (k2_ms_launch * 1e-3) / 1e12,
                         "peak_tflops": fp64_peak, "peak_source": "DFMA micro-benchmark in libfdga (fdga_measure_fp64_peak), measured in this run",
                         "frac": flop_per_launch / (k2_ms_launch * 1e-3) / 1e12 / fp64_peak}}

    # second hot entry of the reference: the mfRG linear map y = A x (src/mfRG.jl:34-89, script/benchmark_Wu.jl:60-64),
    # host vectors in and out exactly as Krylov.dqgmres calls it
    mfrg_per_s = krylov_per_s = None
    if not a.no_extras:
        A = fd.mfRGLinearMap(S)
        xm = x_np                                            # pinned host vectors in and out (x_host / y_host above)
        mroot = 0 if world > 1 else None                     # multi-rank: the vectors live on rank 0's host (one upload + NVLink broadcast)
        A.matvec(xm, out=y_np, root=mroot); A.matvec(xm, out=y_np, root=mroot)
        barrier()
        t0 = time.perf_counter()
        nmv = max(3, min(a.steps, 50))
        for _ in range(nmv):
            A.matvec(xm, out=y_np, root=mroot)
        barrier()
        mfrg_per_s = nmv / (time.perf_counter() - t0)
        if world > 1:                                        # the root-based map and the every-rank map agree bit for bit
            y_all = A.matvec(xm)
            if rank == 0 and not np.array_equal(y_all, y_np):
                raise SystemExit("bench.py: fdga_mfrg_matvec_from_root disagrees with fdga_mfrg_matvec_strategy")
        # the same operator inside the device-resident DQGMRES (Krylov.dqgmres(...; memory = 100) of src/mfRG.jl:147-151): one Krylov
        # iteration = one matvec + the incomplete orthogonalisation and direction update, all vectors in HBM
        nk = 30
        kmem = int(max(2, min(100, 8e9 // (2 * nF * 16))))    # the rings of basis / direction vectors stay under 8 GB
        fd.dqgmres(A, xm, memory=kmem, atol=0.0, rtol=0.0, itmax=2)
        barrier()
        t0 = time.perf_counter()
        _, kst = fd.dqgmres(A, xm, memory=kmem, atol=0.0, rtol=0.0, itmax=nk)
        barrier()
        krylov_per_s = kst["niter"] / (time.perf_counter() - t0)

    # state fingerprint after one more iteration from the stashed vertex: identical on every rank and for every N
    import hashlib
    step_resident_eager()
    S.flatten_F(y_np); S.get_green_into("Σ", s_np)
    digest = hashlib.sha1(y_np.tobytes() + s_np.tobytes()).hexdigest()[:16]
    checksum = float(np.abs(y_np).sum() + np.abs(s_np).sum())
    if dist is not None:
        objs = [None] * world
        dist.all_gather_object(objs, digest)
        if rank == 0 and len(set(objs)) != 1:
            raise SystemExit(f"bench.py: ranks disagree on the iterated state: {objs}")

    # device memory in use on every rank (the slab-shaped arrays are sharded: each rank holds only the (W, P) slabs it reads)
    free_b, total_b = torch.cuda.mem_get_info()
    mem_gb = round((total_b - free_b) / 1e9, 3)
    mem_all = [mem_gb]
    if dist is not None:
        mem_all = [None] * world
        dist.all_gather_object(mem_all, mem_gb)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as o
        o.build()
        cores = host_threads(o)
        R = make_oracle_solver(o, inp, share_bubbles_from=S)
        sec, parts = cpu_full_iteration(o, R, inp["F"])
        cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "ONE complete un-sampled iteration (iterate_solver!(fdPA, update_Σ=false) + SDE!(scPA)) of the oracle restatement, measured wall time",
               "s_per_iteration": sec, "parts_s": parts}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": W, "ms_per_step": ms / a.steps,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64 (complex)", "data": inp["data"],
               "config": {"workload": workload_name(a), "l2": "inputs larger than L2, nothing flushed between steps: K2 tables of S.F, S.F0, S.F + S.F0, FL, Fbuff and their momentum-fastest copies %.0f MB + compact bubble slabs and right factors (only the (W,P) slabs with class representatives) + scratch tables, against 126 MB of L2" % (16e-6 * S.F.γp.K2.size * 3 * 14),
                          "symmetry_classes": {"K1": S.num_classes(fd._lib.SG_K1), "K2pp": n2cls[0], "K2ph": n2cls[1], "K3pp": S.num_classes(fd._lib.SG_PP3), "K3ph": S.num_classes(fd._lib.SG_PH3)},
                          "parallelism": f"class representatives sharded over {world} rank(s), NCCL all-gather per kernel" if world > 1 else "1 GPU",
                          "issue": "every timed step is one replay of the step's CUDA graph (recorded once from the same library calls)" if use_graph else "eager launches"},
               "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(nF * 16), "d2h_bytes_per_step": int(nF * 16 + S.Σ.size * 16), "ms_per_step": ms_e2e / a.steps},
               "gpu_launches": int(launches), "mfrg_matvecs_per_sec_e2e": mfrg_per_s, "mfrg_dqgmres_iterations_per_sec_device_resident": krylov_per_s, "state_sha1": digest, "state_checksum": checksum, "device_memory_gb_per_rank": mem_all, "kernels": kernels, "roofline": roofline, "cpu_baseline": cpu, "clocks": clk}
        print(json.dumps(out))
    S.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
