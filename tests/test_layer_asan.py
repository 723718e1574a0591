"""Memory safety of the layer-epilogue kernels' index arithmetic without a GPU: tests/native/layer_asan_driver.cu includes
csrc/sc_layer.cu, is built with AddressSanitizer and runs the kernels' host checks (the very __host__ __device__ tile functions the
GPU executes) on exact-size heap buffers over a sweep of extents around every tile edge.  Any out-of-bounds access aborts the run.
(The full sweep, 2847 shape cases, was run clean on the committed sc_layer.cu; the tier runs the "quick" subset.)"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_layer_tile_functions_are_asan_clean(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "layer_asan")
    build = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-g", "-std=c++17", "-Xcompiler", "-fsanitize=address",
                            "-Xcompiler", "-fno-omit-frame-pointer", "-I", os.path.join(ROOT, "include"),
                            "-I", os.path.join(ROOT, "neuraloperator_b200", "csrc"),
                            os.path.join(ROOT, "tests", "native", "layer_asan_driver.cu"), "-o", exe],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert build.returncode == 0, build.stdout[-2000:]
    env = dict(os.environ, ASAN_OPTIONS="protect_shadow_gap=0:detect_leaks=0")
    run = subprocess.run([exe, "quick"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, env=env)
    assert run.returncode == 0, run.stdout[-3000:]
    assert "0 failures" in run.stdout
