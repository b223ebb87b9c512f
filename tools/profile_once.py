"""Minimal driver for ncu captures: one warm-up forward pass, then one profiled forward pass (+ masks) of a
workload, plain launches (no CUDA graph) so that kernels appear individually.  Usage:
    python tools/profile_once.py [model] [batch] [h] [w]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yolo_puncture_b200 import YOLO, synth  # noqa: E402
from yolo_puncture_b200.model import box_xform  # noqa: E402


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "yolov8s-seg"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 640
    W = int(sys.argv[4]) if len(sys.argv) > 4 else 640
    yolo = YOLO(model, device=0)
    eng = yolo.engine
    eng.set_graph(False)
    eng.plan(B, H, W)
    frames = torch.from_numpy(np.stack([synth.synth_frame(i % 8, H, W) for i in range(B)])).cuda()
    xf = torch.tensor([box_xform((H, W), (H, W))] * B, dtype=torch.float32).cuda()
    masks = None
    for it in range(2):
        eng.infer(frames, xf, 0.25, 0.7)
        torch.cuda.synchronize()
        if yolo.task == "segment":
            n = max(int(eng.count.sum().item()), 1)
            masks = torch.empty((n, H, W), dtype=torch.uint8, device="cuda") if masks is None else masks
            eng.masks(masks, True, H, W)
            torch.cuda.synchronize()
    print("ok", eng.count.sum().item(), "detections, device error", eng.device_error())


if __name__ == "__main__":
    main()
