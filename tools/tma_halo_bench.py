"""Halo-box fetch ceiling: what conv3_halo_kernel's producer can get for a (B, H, W, C_tot) bf16 map.

    python tools/tma_halo_bench.py
Each CTA streams {64 ch, 10, R} boxes (R = 16*msub + 2 rows) for tiles strided over the grid; a consumer thread
frees the stage as soon as it lands.  Columns: us per box per SM, GB/s of shared-memory fill, chip-wide.
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_puncture_b200._lib import check, diag_lib, lib  # noqa: E402


def run(buf, ctot, R, prods, stages, iters, W, H, B):
    ms, by = C.c_float(), C.c_double()
    check(diag_lib().ypb_tma_bench(C.c_void_p(buf.data_ptr()), 4, stages, iters, ctot | (R << 16) | (prods << 24), W, H, B,
                              C.byref(ms), C.byref(by)))
    return ms.value, by.value


def main():
    B = 64
    for (ctot, W, H) in ((32, 160, 160), (96, 160, 160), (64, 80, 80), (128, 160, 160), (64, 160, 160)):
        buf = torch.zeros(B * H * W * ctot, dtype=torch.bfloat16, device="cuda")
        for R in (18, 34):
            tiles = (W // 8) * ((H + R - 3) // (R - 2)) * B
            iters = (tiles + 147) // 148
            for prods in (1,):
                for stages in (2, 3, 4):
                    if stages * R * 1280 > 200 * 1024:
                        continue
                    ms, by = run(buf, ctot, R, prods, stages, iters, W, H, B)
                    print(f"C_tot {ctot:3d} {W}x{H} R {R} producers {prods} stages {stages}: {ms:7.4f} ms  "
                          f"{ms * 1e3 / iters:6.3f} us/box  fill {by / ms / 1e6:7.0f} GB/s  "
                          f"unique input {B * H * W * min(ctot, 64) * 2 / ms / 1e6:6.0f} GB/s")
        del buf


if __name__ == "__main__":
    main()
