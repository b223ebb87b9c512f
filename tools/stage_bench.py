"""Host staging micro-benchmark: pageable frames -> pinned ring through ypb_stage_frames_ex (memcpy vs non-temporal stores, 1-32 threads)."""
import ctypes as C, time, numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_puncture_b200._lib import lib
import torch
n=64; sz=640*640*3
src=[np.random.randint(0,255,(sz,),dtype=np.uint8) for _ in range(n)]
dst=torch.empty((n,sz),dtype=torch.uint8)
try: dst=dst.pin_memory()
except Exception as e: print('no pin', e)
sp=(C.c_void_p*n)(*[s.ctypes.data for s in src]); dp=(C.c_void_p*n)(*[dst[i].data_ptr() for i in range(n)]); by=(C.c_size_t*n)(*([sz]*n))
for mode in (0,1):
  for th in (1,2,4,8,16,32):
    lib().ypb_stage_frames_ex(dp,sp,by,n,th,mode)
    t0=time.perf_counter()
    for _ in range(5): lib().ypb_stage_frames_ex(dp,sp,by,n,th,mode)
    dt=(time.perf_counter()-t0)/5
    ok = all(np.array_equal(dst[i].numpy(), src[i]) for i in (0,n-1))
    print(f"mode {mode} threads {th:2d}: {dt*1e3:.2f} ms for {n} frames = {n*sz/dt/1e9:.1f} GB/s ok={ok}")
