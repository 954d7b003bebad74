"""The kernels' packed integer arithmetic, unit-tested on the CPU: tests/host/kernel_math.cu includes the
__host__ __device__ helpers of the CUDA sources (interp_fast, sat_add_bgr, tile_row_word, lane_*) and checks them
against the scalar definitions of cv2.remap / BlendMask / cv2.add.  nvcc compiles it; only host code runs."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def test_packed_kernel_arithmetic_on_the_host(tmp_path):
    nvcc = _nvcc()
    if nvcc is None:
        pytest.skip("nvcc not found")
    exe = tmp_path / "kernel_math"
    src = os.path.join(ROOT, "tests", "host", "kernel_math.cu")
    build = subprocess.run([nvcc, "-O2", "-std=c++17", "--fmad=false", "-gencode", "arch=compute_100a,code=sm_100a",
                            "-o", str(exe), src], capture_output=True, text=True, timeout=600)
    assert build.returncode == 0, build.stdout + build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "fails=0" in run.stdout
