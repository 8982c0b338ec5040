"""In-tree build of libskyjo_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m skyjo_rl_b200.build [--force]

The fused step kernel is instantiated once per player count (csrc/skyjo_step_inst.cu with
-DSKYJO_N=1..12); the 13 translation units compile in parallel and link into
skyjo_rl_b200/libskyjo_b200.so next to this file, so the library travels with the tree.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libskyjo_b200.so")
NVCC = os.environ.get("NVCC", "nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--fmad=false",
] + os.environ.get("SKYJO_NVCC_EXTRA", "").split()


def _sources():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "skyjo_b200.h")]
    return deps


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(job):
    src, obj, extra = job
    cmd = [NVCC, *FLAGS, *extra, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return obj


def build(force=False, verbose=False):
    deps = _sources()
    if not force and not _stale(LIB, deps):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    jobs = [(os.path.join(CSRC, "skyjo_capi.cu"), os.path.join(OBJ, "skyjo_capi.o"), []),
            (os.path.join(CSRC, "skyjo_hostsimd.cpp"), os.path.join(OBJ, "skyjo_hostsimd.o"), []),
            (os.path.join(CSRC, "skyjo_policy.cu"), os.path.join(OBJ, "skyjo_policy.o"), [])]
    for n in range(1, 13):
        jobs.append((os.path.join(CSRC, "skyjo_step_inst.cu"), os.path.join(OBJ, f"skyjo_step_n{n}.o"), [f"-DSKYJO_N={n}"]))
    return _run(jobs, LIB, verbose)


def build_variant(out, defines=(), players=(4,), verbose=False):
    """Development aid (tools/variants.py): an experimental build of the library with extra -D
    flags, only the listed player counts instantiated (the others fail with SKYJO_E_INVALID)."""
    tag = os.path.splitext(os.path.basename(out))[0]
    obj = os.path.join(OBJ, "variants", tag)
    os.makedirs(obj, exist_ok=True)
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    dflags = [f"-D{d}" for d in defines]
    mask = sum(1 << (n - 1) for n in players)
    jobs = [(os.path.join(CSRC, "skyjo_capi.cu"), os.path.join(obj, "skyjo_capi.o"), dflags + [f"-DSKYJO_ONLY_PLAYERS_MASK={mask}"]),
            (os.path.join(CSRC, "skyjo_hostsimd.cpp"), os.path.join(obj, "skyjo_hostsimd.o"), []),
            (os.path.join(CSRC, "skyjo_policy.cu"), os.path.join(obj, "skyjo_policy.o"), dflags)]
    for n in players:
        jobs.append((os.path.join(CSRC, "skyjo_step_inst.cu"), os.path.join(obj, f"skyjo_step_n{n}.o"), dflags + [f"-DSKYJO_N={n}"]))
    return _run(jobs, out, verbose)


def _run(jobs, LIB, verbose):
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(_compile, jobs))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed: {r.stdout}\n{r.stderr}")
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
