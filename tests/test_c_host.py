"""The C ABI from a plain C host (examples/c_host.c): no Python, no torch -- device buffers from cudaMalloc, the env
driven through include/skyjo_b200.h.  Compiling it as C99 with -Wall -Wextra -Werror also proves that the header is
valid C (not only C++) and that every symbol it uses links against the shipped library."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build_c_host(tmp_path):
    exe = str(tmp_path / "c_host")
    lib_dir = os.path.join(ROOT, "skyjo_rl_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(CUDA, "include"), os.path.join(ROOT, "examples", "c_host.c"), "-o", exe,
           "-L" + lib_dir, "-lskyjo_b200", "-L" + os.path.join(CUDA, "lib64"), "-lcudart", "-Wl,-rpath," + lib_dir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


needs_toolchain = pytest.mark.skipif(shutil.which("gcc") is None or not os.path.isdir(os.path.join(CUDA, "include")),
                                     reason="gcc and the CUDA headers are needed")


@needs_toolchain
def test_c_host_compiles_links_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe = build_c_host(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the run is test_c_host_runs_on_the_gpu")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "no CPU fallback" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@needs_toolchain
def test_c_host_runs_on_the_gpu(tmp_path):
    exe = build_c_host(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "c_host ok" in r.stdout, r.stdout + r.stderr
    assert "0 illegal" in r.stdout
