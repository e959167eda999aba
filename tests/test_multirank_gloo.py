"""N > 1 host logic on CPU: world_size-2 `gloo` run of the sharded class-representative scheme.

Each rank evaluates its block of class representatives of BSE_K2! / BSE_K1! (with the CPU oracle standing in for the
kernels), the blocks are all-gathered (padded to `chunk` slots per rank, exactly what libfdga does with ncclAllGather)
and expanded to all class members; the result must equal the single-rank evaluation (to rounding of the test's own FL bookkeeping)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import fddgasolver_jl_b200 as fd
    import oracle as o
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    inp = fd.wu_point_inputs(2, 3, 6, small_reference=True, F0_scale=0.03, F_scale=0.2)
    R = o.OracleSolver(inp["nK1"], inp["nK2"], inp["nK3"], inp["L"], inp["Gbare"], inp["G0"], inp["Σ0"], inp["F0"], T=inp["T"])
    R.init_sym_grp(); R.F.set(inp["F"]); R.FL.set(inp["F"])
    out = {}
    for name, fn, sgk in (("K2a", lambda c0, c1: o.BSE_K2(R, fd.aCh, c0=c0, c1=c1), o.SG_PH2), ("K1p", lambda c0, c1: o.BSE_K1(R, fd.pCh, c0=c0, c1=c1), o.SG_K1)):
        offsets, index, ops = R.sg[sgk]
        ncls = len(offsets) - 1
        c0, c1, chunk = fd._lib.partition(ncls, world, rank)
        arr = R.Fbuff.γa.K2 if name == "K2a" else R.Fbuff.γp.K1
        arr[...] = 0
        fn(c0, c1)                                              # this rank's representatives only
        flat = arr.reshape(-1, order="F").copy()
        if name == "K2a":
            flat -= R.FL.γa.K2.reshape(-1, order="F")           # the oracle wrapper already added FL.K2: take the bare SG(...) values
        mine = np.zeros(chunk, dtype=np.complex128)
        mine[: c1 - c0] = flat[index[offsets[c0:c1]]]
        gathered = [torch.zeros(chunk, dtype=torch.complex128) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(mine))
        repvals = torch.cat(gathered).numpy()[:ncls]
        full = np.zeros(flat.size, dtype=np.complex128)         # expansion (expand_kernel)
        cls_of = np.repeat(np.arange(ncls), np.diff(offsets))
        v = repvals[cls_of]
        v = np.where(ops & 2, np.conj(v), v); v = np.where(ops & 1, -v, v)
        full[index] = v
        if name == "K2a":
            full += R.FL.γa.K2.reshape(-1, order="F")           # BSEa_K2.jl:134 add!(K2, FL.K2)
        out[name] = full
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_covers_all_classes():
    import fddgasolver_jl_b200 as fd
    for ncls in (1, 7, 240, 10262):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                c0, c1, chunk = fd._lib.partition(ncls, world, r)
                assert 0 <= c0 <= c1 <= ncls and c1 - c0 <= chunk and chunk * world >= ncls
                got += list(range(c0, c1))
            assert got == list(range(ncls))


def test_two_rank_gloo_matches_single_rank(orc):
    import torch.multiprocessing as mp
    import fddgasolver_jl_b200 as fd
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    inp = fd.wu_point_inputs(2, 3, 6, small_reference=True, F0_scale=0.03, F_scale=0.2)
    R = orc.OracleSolver(inp["nK1"], inp["nK2"], inp["nK3"], inp["L"], inp["Gbare"], inp["G0"], inp["Σ0"], inp["F0"], T=inp["T"])
    R.init_sym_grp(); R.F.set(inp["F"]); R.FL.set(inp["F"])
    orc.BSE_K2(R, fd.aCh); orc.BSE_K1(R, fd.pCh)
    ref = {"K2a": R.Fbuff.γa.K2.reshape(-1, order="F"), "K1p": R.Fbuff.γp.K1.reshape(-1, order="F")}
    for rank in (0, 1):
        for k in ref:
            assert np.max(np.abs(res[rank][k] - ref[k])) <= 1e-14 * np.max(np.abs(ref[k])), (rank, k)
