"""Where index_masks() spends its time on a workload (host stages vs device).  usage: handoff_probe.py [workload]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from yolo_puncture_b200 import YOLO, handoff, index_masks  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "yolov8m-seg-1080p-b16"
model, B, hw, imgsz = bench.WORKLOADS[wl]
yolo = YOLO(model, device=0, synth_geometry=bench.synth_geometry(hw))
frames = bench.make_frames(B, hw, 0)
res = yolo.predict(frames, conf=bench.CONF, iou=bench.IOU, retina_masks=True, imgsz=imgsz, batch=B)
torch.cuda.synchronize()
raws = [r.masks.raw for r in res if r.masks is not None]
print("frames with masks", len(raws), "detections", sum(int(m.shape[0]) for m in raws), "cropped", [getattr(r.masks, "cropped", None) for r in res][:3])
print("contiguous chain", all(b.data_ptr() == a.data_ptr() + a.numel() for a, b in zip(raws, raws[1:])), "aligned", raws[0].data_ptr() % 16)
for k in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = index_masks(res, suppress_small_mask=True, min_area=100)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"call {k}: index_masks returned after {(t1 - t0) * 1e3:.3f} ms, device drained after {(t2 - t0) * 1e3:.3f} ms, kept {sum(len(i) for _, i in out)}")
# stage timing of one call
t = time.perf_counter()
hb = [handoff._host_boxes(r) for r in res if r.masks is not None]
print(f"host boxes {(time.perf_counter() - t) * 1e3:.3f} ms")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
out = index_masks(res, suppress_small_mask=True, min_area=100)
ev1.record()
torch.cuda.synchronize()
print(f"device time between events {ev0.elapsed_time(ev1):.3f} ms")
