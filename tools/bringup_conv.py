"""GPU bring-up driver: runs every conv parity case in its own subprocess (a faulting kernel poisons
the CUDA context) with a timeout, and prints a mismatch map for each.  Usage: python tools/bringup_conv.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from test_gpu_conv import CASES, run_case, RTOL, ATOL
from gpu_util import describe_mismatch
from yolo_puncture_b200._lib import lib
import ctypes
i, impl = int(sys.argv[1]), int(sys.argv[2])
got, ref, clean = run_case(CASES[i], impl)
w = ctypes.c_uint32(); lib().ypb_device_error(None, ctypes.byref(w))
ok = bool(((got - ref).abs() <= ATOL + RTOL * ref.abs()).all())
print('CASE', i, 'impl', impl, CASES[i], 'OK' if ok else 'FAIL', 'clean' if clean else 'DIRTY', 'deverr', w.value)
if not ok: print(describe_mismatch(got, ref, RTOL, ATOL))
"""


def main():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_conv import CASES
    impls = [int(a) for a in sys.argv[1:]] or [1, 0]
    for impl in impls:
        for i in range(len(CASES)):
            try:
                r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, ROOT), str(i), str(impl)], capture_output=True,
                                   text=True, timeout=120)
                print(r.stdout.strip())
                if r.returncode != 0:
                    print("  rc", r.returncode, r.stderr.strip()[-1500:])
            except subprocess.TimeoutExpired:
                print("CASE", i, "impl", impl, "TIMEOUT")
            sys.stdout.flush()


if __name__ == "__main__":
    main()
