"""Error behaviour of the C-ABI (status code + fdga_last_error, never a crash) and a 2-rank NCCL run (skipped on a single GPU)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from test_gpu_parity import make_pair

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_error_paths_return_status_and_message(orc):
    import fddgasolver_jl_b200 as fd
    from fddgasolver_jl_b200 import _lib as L
    S, _ = make_pair(orc, nmax=2, nq=3, LG=6, sym=False)
    S.reset_sym_grp()
    lib = L.load()
    z = np.zeros(4, dtype=np.complex128)
    cases = [
        lambda: S._call("fdga_bse_K2", 7, 0),                                    # bad channel
        lambda: S._call("fdga_set_vertex", 0, 0, 1, L.ptr(z), 4),                # length mismatch
        lambda: S._call("fdga_set_vertex", 55, 0, 0, L.ptr(z), 4),               # bad vertex selector
        lambda: S._call("fdga_set_bubble", 9, L.ptr(z), 4),                      # bad bubble selector
        lambda: S._call("fdga_iterate_solver", 17, 0, 1),                           # unknown strategy
        lambda: S._call("fdga_set_option", 99, 1),                               # unknown option
        lambda: S._call("fdga_interpolate_green", 9, L.ptr(z), 1, 1, 0),         # bad selector
        lambda: S._call("fdga_update_reference"),                                # before fdga_mix_bubbles
    ]
    for i, c in enumerate(cases):
        with pytest.raises(fd.FdgaError) as e:
            c()
        assert "failed (status 1)" in str(e.value) and len(str(e.value)) > 30, (i, str(e.value))
    # the context is still usable after every failure
    S.init_sym_grp()
    fd.iterate_solver(S, "fdPA", True)
    S.pull("Σ")
    assert np.isfinite(S.Σ).all()
    # malformed class table: repeated index
    offs = np.array([0, 2], dtype=np.int64); idx = np.array([0, 0], dtype=np.int64); ops = np.zeros(2, dtype=np.uint8)
    with pytest.raises(fd.FdgaError):
        S.set_symmetry_classes(L.SG_K1, offs, idx, ops)
    S.close()
    # no device behind the ordinal -> create fails with a message, no context
    import ctypes as C
    ctx = C.c_void_p()
    d = S._dims if hasattr(S, "_dims") else None
    if d is not None:
        assert lib.fdga_create(C.byref(d), 4096, C.byref(ctx)) != 0 and lib.fdga_last_error(None)


def test_two_ranks_nccl_match_one_rank():
    """one process per GPU, NCCL all-gather of the sharded class representatives: same state as a single rank"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")

    def run(n, extra=()):
        cmd = [sys.executable]
        if n > 1:
            cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", "29577"]
        cmd += [os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--steps", "3", "--warmup", "3", "--nmax", "3", "--nq", "4", "--LG", "8", "--no-cpu-baseline"] + list(extra)
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
        assert out.returncode == 0, out.stderr[-2000:]
        return json.loads(out.stdout.strip().splitlines()[-1])
    a, b = run(1), run(2)
    assert abs(a["state_checksum"] - b["state_checksum"]) <= 1e-10 * abs(a["state_checksum"])
    assert b["n_gpus"] == 2 and "comm" in b["kernels"]
    # the s-wave solver shards the same way (class representatives of K1 / K2[W,v,P] / K3, one batched all-gather per stage)
    a, b = run(1, ("--nl-method", "1")), run(2, ("--nl-method", "1"))
    assert "s-wave" in a["config"]["workload"] and abs(a["state_checksum"] - b["state_checksum"]) <= 1e-10 * abs(a["state_checksum"])
    assert b["n_gpus"] == 2 and "comm" in b["kernels"]
    # and so do the generic per-term kernels of a solver with MBE vertices
    a, b = run(1, ("--nl-method", "-2")), run(2, ("--nl-method", "-2"))
    assert "MBE" in a["config"]["workload"] and a["state_sha1"] == b["state_sha1"] and b["n_gpus"] == 2
