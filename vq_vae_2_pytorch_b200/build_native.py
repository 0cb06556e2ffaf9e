"""Build libvqb200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvqb200.so")
# experiments: VQB200_NVCC_DEFS="-DVQB200_ROLEMAP=1" VQB200_LIB_OUT=/path/libvqb200_exp.so python build_native.py --force
EXTRA_DEFS = os.environ.get("VQB200_NVCC_DEFS", "").split()
LIB_OUT = os.environ.get("VQB200_LIB_OUT", LIB)
SOURCES = ["vqb200_abi.cu"]
HEADERS = ["common.cuh", "simt_kernels.cuh", "tc_kernel.cuh", "tc_wide_kernel.cuh", "fused_kernels.cuh", os.path.join("..", "..", "include", "vqb200.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the vqb200 CUDA library cannot be built")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not is_stale() and LIB_OUT == LIB:
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared"] + (["-DVQB200_P2P_TRACE"] if os.environ.get("VQB200_P2P_TRACE") else []) + [ "-Xptxas", "-v" if verbose else "-O3",
           "-o", LIB_OUT] + EXTRA_DEFS + [os.path.join(CSRC, s) for s in SOURCES] + ["-lcuda"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libvqb200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
