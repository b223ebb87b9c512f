"""Run the TMA operand-fetch microbenchmark over a few access patterns.  python tools/tma_bench.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_puncture_b200._lib import check, lib  # noqa: E402


def run(buf, mode, stages, iters, rows=0, W=0, H=0, B=0):
    ms, by = C.c_float(), C.c_double()
    check(lib().ypb_tma_bench(C.c_void_p(buf.data_ptr()), mode, stages, iters, rows, W, H, B, C.byref(ms), C.byref(by)))
    return by.value / (ms.value * 1e-3) / 1e9, ms.value


def main():
    big = torch.zeros(2 << 30, dtype=torch.uint8, device="cuda")      # 2 GiB: streams from HBM
    for stages in (2, 4, 8, 12):
        gbs, ms = run(big, 0, stages, 4000, rows=(2 << 30) // 128)
        print(f"mode0 stream HBM   stages {stages:2d}: {gbs:8.0f} GB/s  ({gbs / 148:6.1f} GB/s/SM)  {ms:.3f} ms")
    small = big[: 64 << 20]                                              # 64 MiB: L2 resident
    for stages in (2, 4, 8, 12):
        gbs, ms = run(small, 0, stages, 4000, rows=(64 << 20) // 128)
        print(f"mode0 stream L2    stages {stages:2d}: {gbs:8.0f} GB/s  ({gbs / 148:6.1f} GB/s/SM)  {ms:.3f} ms")
    for stages in (2, 8):
        gbs, ms = run(small, 1, stages, 4000, rows=(64 << 20) // 128)
        print(f"mode1 same box     stages {stages:2d}: {gbs:8.0f} GB/s  ({gbs / 148:6.1f} GB/s/SM)  {ms:.3f} ms")
    for (W, H, B) in ((80, 80, 64), (160, 160, 64)):
        for stages in (2, 8):
            gbs, ms = run(big, 2, stages, 3996, W=W, H=H, B=B)
            print(f"mode2 3x3 taps {W}x{H}x{B} stages {stages:2d}: {gbs:8.0f} GB/s  ({gbs / 148:6.1f} GB/s/SM)  {ms:.3f} ms")


if __name__ == "__main__":
    main()
