"""Tensor-pipe ceiling for M=128 x N smem-fed MMAs, alone and with concurrent TMA streaming.  python tools/mma_bench.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_puncture_b200._lib import check, diag_lib, lib  # noqa: E402

buf = torch.zeros(1 << 30, dtype=torch.uint8, device="cuda")
rows = (1 << 30) // 128
iters = 20000
for n in (64, 128, 256):
    for shifted in (0, 1):
        for tma in (0, 1):
            ms = C.c_float()
            # TMA stream sized to last about as long as the MMAs (16 KB per ~0.35 us)
            tma_iters = 0 if not tma else int(iters * (n / 2) / 1965 / 0.35)
            check(diag_lib().ypb_mma_bench(C.c_void_p(buf.data_ptr()), rows, n, iters, shifted, tma_iters, C.byref(ms)))
            tf = 148 * iters * 2.0 * 128 * n * 16 / (ms.value * 1e-3) / 1e12
            cyc = ms.value * 1e-3 * 1.965e9 / iters
            print(f"N {n:3d} shifted {shifted} tma_iters {tma_iters:6d}: {ms.value:7.3f} ms  {tf:7.1f} TFLOP/s  {cyc:6.1f} cyc/MMA"
                  + (f"  tma {148 * tma_iters * 16384 / (ms.value * 1e-3) / 1e9:6.0f} GB/s" if tma else ""))
