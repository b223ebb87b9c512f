"""predict() on pageable frames under different pass schedules.  python tools/e2e_probe2.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_puncture_b200 import YOLO, synth

yolo = YOLO("yolov8s-seg", device=0)
N = 64
frames = [synth.synth_frame(i) for i in range(N)]
for mb, hp in ((None, 16), (32, 16), (16, 16), (8, 16), (24, 16), (None, 16)):
    yolo.micro_batch = mb
    yolo.head_pass = hp
    for _ in range(3):
        yolo.predict(frames, conf=0.25, retina_masks=True, batch=N)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        yolo.predict(frames, conf=0.25, retina_masks=True, batch=N)
    torch.cuda.synchronize()
    print(f"predict(64 pageable frames) micro_batch {mb} head {hp}: {(time.perf_counter() - t0) * 100:.3f} ms")
