"""Does tcgen05.mma read a SWIZZLE_128B tile correctly through a descriptor whose start is advanced by whole rows
and whose 8-row groups are 1280 B apart (halo reuse for 3x3 convs)?  python tools/halo_experiment.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import conv_reference, describe_mismatch  # noqa: E402
from yolo_puncture_b200.engine import conv2d_bf16, gemm_weight  # noqa: E402

for (B, H, W, cin, cout) in [(1, 16, 8, 64, 64), (2, 32, 24, 64, 64), (2, 40, 40, 128, 128), (1, 20, 20, 32, 16)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((B, H, W, cin), device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn((cout, cin, 3, 3), device="cuda", generator=g) / (cin * 9) ** 0.5
    bias = torch.randn((cout,), device="cuda", generator=g)
    out = conv2d_bf16(x, gemm_weight(w), bias, 3, 1, 1, impl=3)
    torch.cuda.synchronize()
    ref = conv_reference(x, w, bias, 3, 1, 1)
    ok = bool(((out.float() - ref).abs() <= 2e-2 + 2 ** -7 * ref.abs()).all())
    print((B, H, W, cin, cout), "OK" if ok else "FAIL")
    if not ok:
        print(describe_mismatch(out.float(), ref, 2 ** -7, 2e-2))
