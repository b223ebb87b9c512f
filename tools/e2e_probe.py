"""How engine time scales with the pass size, and what predict() makes of it.  python tools/e2e_probe.py [spec]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_puncture_b200 import YOLO, synth
from yolo_puncture_b200.model import box_xform

name = sys.argv[1] if len(sys.argv) > 1 else "yolov8s-seg"
yolo = YOLO(name, device=0)
N = 64
pin = torch.empty((N, 640, 640, 3), dtype=torch.uint8).pin_memory()
for i in range(N):
    pin[i] = torch.from_numpy(synth.synth_frame(i))
frames = [pin[i].numpy() for i in range(N)]
dev = pin.cuda()
eng = yolo.engine
for B in (8, 16, 24, 32, 48, 64):
    eng.plan(B, 640, 640)
    xf = torch.tensor([box_xform((640, 640), (640, 640))] * B, dtype=torch.float32, device="cuda")
    for _ in range(3):
        eng.infer(dev[:B], xf, 0.25, 0.7)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.infer(dev[:B], xf, 0.25, 0.7)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"engine pass B={B:2d}: {ms:6.3f} ms  ({ms / B * 1e3:6.1f} us/frame)")
for mb, hp in ((None, 16), (32, 16), (64, 16), (None, 8), (None, 12), (None, 16), (None, 24)):
    yolo.micro_batch = mb
    yolo.head_pass = hp
    for _ in range(3):
        yolo.predict(frames, conf=0.25, retina_masks=True, batch=N)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        yolo.predict(frames, conf=0.25, retina_masks=True, batch=N)
    torch.cuda.synchronize()
    print(f"predict(64 pinned frames) micro_batch {mb} head {hp}: {(time.perf_counter() - t0) * 100:.3f} ms")
