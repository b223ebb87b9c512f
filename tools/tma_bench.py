"""TMA operand-fetch microbenchmark over access patterns / box sizes / number of producer threads."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_puncture_b200._lib import check, diag_lib, lib  # noqa: E402


def run(buf, mode, stages, iters, rows=0, W=0, H=0, B=0):
    ms, by = C.c_float(), C.c_double()
    check(diag_lib().ypb_tma_bench(C.c_void_p(buf.data_ptr()), mode, stages, iters, rows, W, H, B, C.byref(ms), C.byref(by)))
    return by.value / (ms.value * 1e-3) / 1e9, ms.value


def main():
    big = torch.zeros(2 << 30, dtype=torch.uint8, device="cuda")
    small = big[: 64 << 20]
    for name, buf in (("HBM 2GiB", big), ("L2 64MiB", small)):
        rows = buf.numel() // 128
        for box in (128, 256):
            for pairs in (1, 2, 4):
                for stages in (4, 8):
                    if stages * box * 128 > 200 * 1024:
                        continue
                    gbs, ms = run(buf, 0, stages, 4000, rows=rows, W=pairs, H=box)
                    print(f"stream {name} box {box:3d} rows, {pairs} producer(s), {stages:2d} stages: {gbs:8.0f} GB/s ({gbs / 148:6.1f} /SM)")
    for pairs in (1, 2):
        gbs, ms = run(small, 1, 8, 4000, rows=small.numel() // 128, W=pairs, H=128)
        print(f"same box, {pairs} producer(s): {gbs:8.0f} GB/s ({gbs / 148:6.1f} /SM)")
    for (W, H, B) in ((80, 80, 64),):
        gbs, ms = run(big, 2, 8, 3996, W=W, H=H, B=B)
        print(f"3x3 taps {W}x{H}x{B}: {gbs:8.0f} GB/s ({gbs / 148:6.1f} /SM)")


if __name__ == "__main__":
    main()
