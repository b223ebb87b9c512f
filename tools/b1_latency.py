"""End-to-end latency of YOLO.predict(frame) for single frames (what the reference's video loop does, yolo_seg/app.py:85-91)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_puncture_b200 import YOLO, synth

name = sys.argv[1] if len(sys.argv) > 1 else "yolov8s-seg"
yolo = YOLO(name, device=0)
frames = [synth.synth_frame(i) for i in range(8)]
for f in frames[:4]:
    yolo.predict(f, conf=0.25, retina_masks=True)
lat = []
for i in range(200):
    f = frames[i % 8]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = yolo.predict(source=f, conf=0.25, retina_masks=True)
    pb = r[0].boxes.cpu().numpy()
    lat.append((time.perf_counter() - t0) * 1e3)
print(name, "predict(frame)+boxes.cpu().numpy(): p50 %.3f ms  p90 %.3f ms  min %.3f ms" % (np.percentile(lat, 50), np.percentile(lat, 90), min(lat)))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(100):
    r = yolo.predict(source=frames[i % 8], conf=0.25, retina_masks=True); r[0].boxes.cpu().numpy()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(16)
