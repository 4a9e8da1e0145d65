"""Build libchalte.so in-tree with nvcc for sm_100a (python -m cha1_mcmc_b200.build)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libchalte.so")
SOURCES = ["chalte.cu"]
HEADERS = ["lte_common.cuh", "lte_kernels.cuh", "lte_sampler.cuh", os.path.join("..", "..", "include", "chalte.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libchalte.so cannot be built (there is no CPU fallback)")
    return exe


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force=False, verbose=False, out=None, defines=()):
    """Compile every CUDA source into csrc/libchalte.so (or `out`, with extra -D defines: kernel experiments loaded
    through CHALTE_LIB).  Returns the path."""
    target = out or LIB
    if out is None and not force and not is_stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [f"-D{d}" for d in defines] + \
          ["-o", target] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libchalte.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return target


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
