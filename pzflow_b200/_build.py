"""Build libpzflow_b200.so in-tree with nvcc for sm_100a (no torch involved)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpzflow_b200.so")
SOURCES = ["pzb_api.cu", "nsf_simt.cu", "nsf_tc.cu", "nsf_wide.cu", "nsf_train.cu", "nsf_train_step.cu"]
HEADERS = ["pzb_common.cuh", "pzb_tc_common.cuh", "nsf_tc_pipe3.cuh", os.path.join("..", "..", "include", "pzflow_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # the reference evaluates op by op in fp32: no implicit mul+add contraction;
    # the GEMM inner loops call fmaf() explicitly.
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fopenmp", "-shared",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libpzflow_b200.so")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_variant(name: str, defines) -> str:
    """An experimental build with extra -D flags -> pzflow_b200/libpzflow_b200_<name>.so (select it with the
    PZFLOW_B200_LIB environment variable)."""
    out = os.path.join(HERE, f"libpzflow_b200_{name}.so")
    cmd = [_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        print(res.stdout + res.stderr, file=sys.stderr)
        raise RuntimeError(f"nvcc failed building variant {name}")
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(CSRC, "build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if verbose or res.returncode != 0:
        print(log, file=sys.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libpzflow_b200.so (see csrc/build.log)")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
