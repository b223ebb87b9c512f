import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_puncture_b200._lib import check, lib, diag_lib
out = torch.zeros(8, dtype=torch.int64, device="cuda")
for _ in range(2):
    check(diag_lib().ypb_latency_probe(C.c_void_p(out.data_ptr())))
names = ["mbarrier.arrive", "arrive.expect_tx", "test_wait (done)", "tcgen05.commit->visible", "2-warp ping-pong round trip",
         "tcgen05.ld x16 + wait", "1 MMA 128x64x16 + commit->visible"]
for n, v in zip(names, out.tolist()):
    print(f"{n:36s} {v:6d} cycles")
