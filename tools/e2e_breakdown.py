"""Host wall-clock split of YOLO.predict() on 64 pinned frames.  python tools/e2e_breakdown.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_puncture_b200 import YOLO, synth

yolo = YOLO("yolov8s-seg", device=0)
N = 64
pin = torch.empty((N, 640, 640, 3), dtype=torch.uint8).pin_memory()
for i in range(N):
    pin[i] = torch.from_numpy(synth.synth_frame(i))
frames = [pin[i].numpy() for i in range(N)]
for _ in range(3):
    yolo.predict(frames, conf=0.25, retina_masks=True, batch=N)
rows = []
for _ in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = yolo.predict(frames, conf=0.25, retina_masks=True, batch=N)
    t1 = time.perf_counter()
    boxes = torch.cat([r.boxes.data for r in res]).cpu()
    t2 = time.perf_counter()
    sp = res[0].speed
    rows.append(((t1 - t0) * 1e3, sp["preprocess"] * N, sp["inference"] * N, sp["postprocess"] * N, (t2 - t1) * 1e3))
a = np.median(np.array(rows), 0)
print(f"predict() wall {a[0]:.3f} ms = prologue {a[1]:.3f} + passes {a[2]:.3f} + error-word sync {a[3]:.3f} + rest (source handling, Results) "
      f"{a[0] - a[1] - a[2] - a[3]:.3f};  cat+cpu of the boxes {a[4]:.3f} ms")
