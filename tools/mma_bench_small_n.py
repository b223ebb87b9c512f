import ctypes as C, os, sys, torch
sys.path.insert(0, "/root/repo")
from yolo_puncture_b200._lib import check, lib, diag_lib
buf = torch.zeros(1 << 28, dtype=torch.uint8, device="cuda")
rows = (1 << 28) // 128
iters = 20000
for n in (16, 32, 48, 64, 96, 128):
    ms = C.c_float()
    check(diag_lib().ypb_mma_bench(C.c_void_p(buf.data_ptr()), rows, n, iters, 1, 0, C.byref(ms)))
    print(f"N {n:3d}: {ms.value * 1e-3 * 1.965e9 / iters:6.1f} cyc/MMA")
