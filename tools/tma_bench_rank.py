import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_puncture_b200._lib import check, lib, diag_lib
def run(buf, mode, stages, iters, rows=0, W=0, H=0, B=0):
    ms, by = C.c_float(), C.c_double()
    check(diag_lib().ypb_tma_bench(C.c_void_p(buf.data_ptr()), mode, stages, iters, rows, W, H, B, C.byref(ms), C.byref(by)))
    return by.value / (ms.value * 1e-3) / 1e9
buf = torch.zeros(64 << 20, dtype=torch.uint8, device="cuda")
rows = buf.numel() // 128
for variant, name in ((0, "expect_tx then copy"), (1, "relaxed expect_tx then copy"), (2, "copy then expect_tx")):
    for stages in (2, 4, 8):
        g = run(buf, 0, stages, 4000, rows=rows, W=1, H=128, B=variant)
        print(f"{name:30s} stages {stages}: {g:8.0f} GB/s ({g/148:6.1f} /SM)")
