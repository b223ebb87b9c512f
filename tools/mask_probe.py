"""One engine pass + mask decode of a bench workload, a few times - the short program to put under ncu when only the
selection / mask kernels are of interest (bench.py launches thousands of kernels).  usage: mask_probe.py [workload] [reps] [nograph]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from yolo_puncture_b200 import YOLO  # noqa: E402
from yolo_puncture_b200.model import box_xform, letterbox_geometry, letterbox_into  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "yolov8x-seg-640-b32"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
model, B, hw, imgsz = bench.WORKLOADS[wl]
yolo = YOLO(model, device=0, synth_geometry=bench.synth_geometry(hw))
eng = yolo.engine
if len(sys.argv) > 3 and sys.argv[3] == "nograph":
    eng.set_graph(False)
frames = bench.make_frames(B, hw, 0)
new_unpad, top, bottom, left, right = letterbox_geometry(hw, (imgsz, imgsz), auto=True)
H, W = new_unpad[1] + top + bottom, new_unpad[0] + left + right
eng.plan(B, H, W)
host = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
for i, f in enumerate(frames):
    letterbox_into(host[i].numpy(), f, new_unpad, top, left)
fr = host.cuda()
xf = torch.tensor([box_xform((H, W), hw)] * B, dtype=torch.float32, device="cuda")
eng.infer(fr, xf, bench.CONF, bench.IOU)
torch.cuda.synchronize()
n = int(eng.count.sum().item())
masks = torch.empty((max(n, 1), hw[0], hw[1]), dtype=torch.uint8, device="cuda")
for _ in range(reps):
    eng.infer(fr, xf, bench.CONF, bench.IOU)
    eng.masks(masks, True, hw[0], hw[1])
torch.cuda.synchronize()
d = eng.det[:, :, :4]
cnt = eng.count.cpu()
area = sum(float(((d[b, :int(cnt[b]), 2] - d[b, :int(cnt[b]), 0]) * (d[b, :int(cnt[b]), 3] - d[b, :int(cnt[b]), 1])).sum()) for b in range(B))
print(f"{wl}: {n} detections, mean box area {area / max(n, 1):.0f} px = {100 * area / max(n, 1) / (hw[0] * hw[1]):.1f} % of the frame, "
      f"masks set {float(masks[:n].float().mean()) * 100:.2f} %")
