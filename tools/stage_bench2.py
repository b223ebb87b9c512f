"""In-situ staging benchmark: COLD sources (several sets of frames cycled, larger than the last-level cache), per-chunk
calls as predict() issues them, pinned destination, followed by the H2D copy; compares the native pool (memcpy /
non-temporal, thread counts), torch's own pinned copy_ and a direct cudaMemcpy from pageable memory."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_puncture_b200._lib import lib  # noqa: E402

n, sz, sets = 64, 640 * 640 * 3, 6
print("host cores", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
src_sets = [[np.random.randint(0, 255, (sz,), dtype=np.uint8) for _ in range(n)] for _ in range(sets)]
dst = torch.empty((n, sz), dtype=torch.uint8).pin_memory()
dev = torch.empty((n, sz), dtype=torch.uint8, device="cuda")
by = (C.c_size_t * n)(*([sz] * n))
dp = (C.c_void_p * n)(*[dst[i].data_ptr() for i in range(n)])
vp, sp_sz = C.sizeof(C.c_void_p), C.sizeof(C.c_size_t)
L = lib()


def run(label, fn, reps=12):
    fn(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(reps):
        fn(r % sets)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{label:60s} {dt * 1e3:7.2f} ms / 64 frames = {n * sz / dt / 1e9:6.1f} GB/s")


for mode in (0, 1):
    for th in (4, 8, 16, 24):
        for chunk in (8, 64):
            def f(s, mode=mode, th=th, chunk=chunk):
                sp = (C.c_void_p * n)(*[a.ctypes.data for a in src_sets[s]])
                for c0 in range(0, n, chunk):
                    L.ypb_stage_frames_ex(C.byref(dp, c0 * vp), C.byref(sp, c0 * vp), C.byref(by, c0 * sp_sz), chunk, th, mode)
            run(f"pool mode={mode} threads={th} chunk={chunk} (host copy only)", f)


def with_h2d(s, chunk=8, th=16):
    sp = (C.c_void_p * n)(*[a.ctypes.data for a in src_sets[s]])
    for c0 in range(0, n, chunk):
        L.ypb_stage_frames_ex(C.byref(dp, c0 * vp), C.byref(sp, c0 * vp), C.byref(by, c0 * sp_sz), chunk, th, 1)
        dev[c0:c0 + chunk].copy_(dst[c0:c0 + chunk], non_blocking=True)


run("pool NT 16 threads chunk 8 + async H2D per chunk", with_h2d)


def torch_copy(s):
    for i in range(n):
        dst[i].copy_(torch.from_numpy(src_sets[s][i]))


run("torch copy_ into pinned, frame by frame (1 thread)", torch_copy, reps=4)


def pageable_h2d(s):
    for i in range(n):
        dev[i].copy_(torch.from_numpy(src_sets[s][i]), non_blocking=True)


run("direct H2D from pageable memory (driver staging)", pageable_h2d, reps=6)


def pinned_h2d(s):
    dev.copy_(dst, non_blocking=True)


run("H2D from pinned memory (PCIe rate)", pinned_h2d)
