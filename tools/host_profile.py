"""Where does YOLO.predict() spend host wall-clock time?  python tools/host_profile.py [micro_batch]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_puncture_b200 import YOLO, synth
from yolo_puncture_b200 import model as M

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
yolo = YOLO("yolov8s-seg", device=0)
yolo.micro_batch = mb
frames = [synth.synth_frame(i) for i in range(64)]
if len(sys.argv) > 2 and sys.argv[2] == "pinned":
    pin = torch.empty((64, 640, 640, 3), dtype=torch.uint8).pin_memory().numpy()
    for i, f in enumerate(frames):
        pin[i] = f
    frames = [pin[i] for i in range(64)]
for _ in range(3):
    yolo.predict(frames, conf=0.25, retina_masks=True, batch=64)
torch.cuda.synchronize()
# 1. raw staging speed
host = torch.empty((64, 640, 640, 3), dtype=torch.uint8).pin_memory()
hn = host.numpy()
t0 = time.perf_counter()
futs = [M._pool().submit(np.copyto, hn[i], f) for i, f in enumerate(frames)]
[f.result() for f in futs]
t1 = time.perf_counter()
print(f"staging 64 frames via pool: {(t1 - t0) * 1e3:.2f} ms  ({os.cpu_count()} cpus)")
t0 = time.perf_counter()
for i, f in enumerate(frames):
    np.copyto(hn[i], f)
t1 = time.perf_counter()
print(f"staging 64 frames single thread: {(t1 - t0) * 1e3:.2f} ms")
dev = torch.empty_like(host, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
t1 = time.perf_counter()
print(f"H2D 79 MB pinned: {(t1 - t0) * 1e3:.2f} ms = {host.numel() / (t1 - t0) / 1e9:.1f} GB/s")
# 2. predict wall time split
for _ in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = yolo.predict(frames, conf=0.25, retina_masks=True, batch=64)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"predict: {(t1 - t0) * 1e3:.2f} ms (+{(t2 - t1) * 1e3:.2f} ms to drain) speed/frame {res[0].speed}")
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    yolo.predict(frames, conf=0.25, retina_masks=True, batch=64)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
