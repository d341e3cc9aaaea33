"""In-tree build of libocs_b200.so (nvcc, sm_100a only).

    python optimal-control-solvers_b200/build.py [--force]

The library carries the built-in problem's sweep kernels twice (FMA-contracted and
strict arithmetic), the C ABI of include/ocs_b200.h, the NVRTC front end for generated
problems and the roofline micro-benchmarks.  The .so is git-ignored but travels to the
GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libocs_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
                 "-Xptxas", "-v"] + os.environ.get("OCS_EXTRA_NVCC_FLAGS", "").split()

# (object name, source, extra flags)
UNITS = [
    ("kernels_fast.o", "kernels_builtin.cu", ["-DOCS_STRICT=0", "-DOCS_UNIFORM_H=1", "-fmad=true"]),
    ("kernels_fastv.o", "kernels_builtin.cu", ["-DOCS_STRICT=0", "-DOCS_UNIFORM_H=0", "-fmad=true"]),
    ("kernels_strict.o", "kernels_builtin.cu", ["-DOCS_STRICT=1", "-fmad=false"]),
    ("ocs_api.o", "ocs_api.cu", []),
    ("ocs_multi.o", "ocs_multi.cu", []),
    ("host_copy.o", "host_copy.cpp", []),
    ("nvrtc_problem.o", "nvrtc_problem.cu", []),
    ("microbench.o", "microbench.cu", []),
]
HEADERS = ["rk4_kernels.cuh", "solution_interp.cuh", "problem_test.cuh", "kernel_table.h", "ocs_internal.h", "host_stager.h", os.path.join("..", "..", "include", "ocs_b200.h")]


def _newest_input() -> float:
    files = [os.path.join(CSRC, u[1]) for u in UNITS] + [os.path.join(CSRC, h) for h in HEADERS] + [__file__]
    return max(os.path.getmtime(f) for f in files)


def up_to_date() -> bool:
    if os.environ.get("OCS_B200_LIB"):
        return True     # an explicitly selected build is used as is
    return os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_input()


def _compile(unit):
    obj, src, extra = unit
    cmd = [NVCC] + COMMON + extra + ["-c", os.path.join(CSRC, src), "-o", os.path.join(OBJ, obj)]
    p = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJ, obj + ".ptxas.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{p.stdout}\n{p.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(_compile, UNITS))
    cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + [os.path.join(OBJ, o) for o in objs] + ["-ldl", "-lpthread"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    if verbose:
        print("built", LIB)
    return LIB


def build_variant(name: str, flags: list[str]) -> str:
    """Kernel tuning experiments: recompile only the FMA-contracted kernel unit with extra -D flags and link it
    with the objects of the regular build into build/variants/libocs_<name>.so (select it with OCS_B200_LIB)."""
    build()
    vdir = os.path.join(OBJ, "variants")
    os.makedirs(vdir, exist_ok=True)
    obj = os.path.join(vdir, f"kernels_fast_{name}.o")
    cmd = [NVCC] + COMMON + ["-DOCS_STRICT=0", "-DOCS_UNIFORM_H=1", "-fmad=true"] + flags + ["-c", os.path.join(CSRC, "kernels_builtin.cu"), "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    with open(obj + ".ptxas.log", "w") as f:
        f.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
    if p.returncode != 0:
        raise RuntimeError(p.stdout + p.stderr)
    lib = os.path.join(vdir, f"libocs_{name}.so")
    others = [os.path.join(OBJ, u[0]) for u in UNITS if u[0] != "kernels_fast.o"]
    p = subprocess.run([NVCC] + ARCH + ["-shared", "-o", lib, obj] + others + ["-ldl", "-lpthread"], capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(p.stdout + p.stderr)
    return lib


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":      # build.py --variant NAME -DFOO=1 ...
        print(build_variant(sys.argv[2], sys.argv[3:]))
    else:
        build(force="--force" in sys.argv, verbose=True)
