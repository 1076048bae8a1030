"""Build the in-tree CUDA library (sm_100a only).

    python -m maximizingacquisitionfunctions_b200.build

nvcc cross-compiles without a GPU; the resulting ``libmaf_b200.so`` sits next to this
file (git-ignored, shipped to the GPU box with the working tree).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmaf_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    hdr = os.path.join(os.path.dirname(HERE), "include", "maf.h")
    return any(os.path.getmtime(s) > t for s in sources() + [hdr, os.path.abspath(__file__)])


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    # no CUDA math libraries: every kernel, including the training-set factorisation, is in csrc/
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB, os.path.join(CSRC, "maf_api.cu")]
    for extra in os.environ.get("MAF_NVCC_DEFS", "").split():   # tuning experiments: -DNAME=value overrides of kernel macros
        cmd.insert(1, extra)
    if os.environ.get("MAF_OZ_DIAG"):
        cmd.insert(1, "-DMAF_OZ_DIAG")      # extra diagnostic instantiations of the INT8 kernel (MAF_OZ_DBG)
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
