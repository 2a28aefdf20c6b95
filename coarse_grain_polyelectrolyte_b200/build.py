"""Builds libcgpe.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the tree)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcgpe.so")
SOURCES = ["cgpe_api.cu", "chain_engine.cu", "duo_engine.cu", "wide_engine.cu", "diag_kernels.cu"]
HEADERS = ["cgpe_device.cuh", "chain_engine.cuh", "duo_engine.cuh", "wide_engine.cuh", os.path.join("..", "..", "include", "cgpe.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--ftz=true",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _obj_stale(src, obj):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src] + [os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, extra_flags=()):
    """Compile csrc/*.cu into libcgpe.so (stale objects only, one nvcc per source in parallel); returns the path."""
    if not force and not is_stale():
        return LIB
    objs, jobs = [], []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(CSRC, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or extra_flags or _obj_stale(src, obj):
            cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, "-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            jobs.append((cmd, subprocess.Popen(cmd)))
    for cmd, proc in jobs:
        if proc.wait() != 0:
            raise subprocess.CalledProcessError(proc.returncode, cmd)
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC", "-cudart", "static"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True,
                extra_flags=("-Xptxas", "-v") if "--ptxas" in sys.argv else ()))
