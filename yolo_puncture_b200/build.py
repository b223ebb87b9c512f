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
           "misc_kernels.cuh", "v10_kernels.cuh", "conv_halo2.cuh", "host_stage.cpp", "jpeg_source.cpp"]  # + tma_bench.cuh in the diagnostics build
UNITS = [os.path.join(CSRC, "ypb200.cu"), os.path.join(CSRC, "host_stage.cpp"), os.path.join(CSRC, "jpeg_source.cpp")]
# device TU + host-only staging pool + nvJPEG frame source (dlopen: no link-time dependency)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-diag-suppress", "177", "-ldl"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in os.listdir(CSRC) if s != "tma_bench.cuh"] + [os.path.join(ROOT, "include", "ypb200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


LIB_PROF = os.path.join(HERE, "libypb200_prof.so")
LIB_DIAG = os.path.join(HERE, "libypb200_diag.so")
LIB_EXACT = os.path.join(HERE, "libypb200_exact.so")


def _stale(path):
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    deps = [os.path.join(CSRC, s) for s in os.listdir(CSRC)] + [os.path.join(ROOT, "include", h) for h in ("ypb200.h", "ypb200_diag.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, prof=False, diag=False, exact=False):
    """Product library: libypb200.so (the product kernels and the C ABI of include/ypb200.h, nothing else).
    diag=True builds libypb200_diag.so: the same translation unit with -DYPB_DIAG=1, i.e. plus the debugging twins of
    the conv kernel and the micro-benchmarks of include/ypb200_diag.h (tests/test_gpu_conv.py twins, tools/).
    prof=True builds libypb200_prof.so: the diagnostics build with the per-role wait-cycle / epilogue-phase accounting
    compiled in (-DYPB_PROF=1); tools/conv_layers.py loads it through YPB_LIB."""
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libypb200.so")
    if prof or diag or exact:
        out = LIB_EXACT if exact else LIB_PROF if prof else LIB_DIAG
        if not force and not _stale(out):
            return out
        # exact: the product build with full-precision SiLU (A/B of the tanh.approx epilogue against the oracle)
        flags = ["-DYPB_EXACT_SILU=1"] if exact else ["-DYPB_DIAG=1"] + (["-DYPB_PROF=1"] if prof else [])
        res = subprocess.run([nvcc, *NVCC_FLAGS, *flags, "-o", out, *UNITS],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        return out
    if not force and not needs_build():
        return LIB
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, *UNITS]
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
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, prof="--prof" in sys.argv, diag="--diag" in sys.argv,
                exact="--exact" in sys.argv))
