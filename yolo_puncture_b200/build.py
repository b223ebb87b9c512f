"""In-tree build of libypb200.so (hand-written sm_100a CUDA, single translation unit).

    python -m yolo_puncture_b200.build [--force]

nvcc cross-compiles without a GPU.  The built library is git-ignored but travels to the GPU box with
the snapshot; `ensure_built()` only rebuilds when a source is newer than the library.
"""

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libypb200.so")
SOURCES = ["ypb200.cu", "common.cuh", "conv_tc.cuh", "conv_plan.cuh", "head_kernels.cuh", "mask_kernels.cuh",
           "misc_kernels.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-diag-suppress", "177"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "ypb200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


LIB_PROF = os.path.join(HERE, "libypb200_prof.so")


def build(force=False, verbose=False, prof=False):
    """prof=True builds libypb200_prof.so: the same library with the per-role wait-cycle / epilogue-phase accounting
    compiled in (-DYPB_PROF=1); tools/conv_layers.py loads it through YPB_LIB."""
    if prof:
        nvcc = _nvcc()
        res = subprocess.run([nvcc, *NVCC_FLAGS, "-DYPB_PROF=1", "-o", LIB_PROF, os.path.join(CSRC, "ypb200.cu")],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        return LIB_PROF
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libypb200.so")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, os.path.join(CSRC, "ypb200.cu")]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


def ensure_built():
    return build(force=False)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, prof="--prof" in sys.argv))
