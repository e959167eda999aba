"""CPU-only unit tests of the CUDA evaluators' index machinery: the device headers are compiled for the HOST with nvcc
(the evaluators are __host__ __device__) and compared term by term with the straightforward per-term evaluator.
No GPU is used; these run in the `-m "not gpu"` suite."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("src", ["host_eval_test.cu", "host_column_test.cu"])
def test_host_side_kernel_arithmetic(src, tmp_path):
    nvcc = shutil.which("nvcc")
    if nvcc is None:
        pytest.skip("nvcc not available")
    exe = str(tmp_path / src.replace(".cu", ""))
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(ROOT, "tests", src)])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "WORST" in out.stdout
