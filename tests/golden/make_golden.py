"""Generates the committed golden fixtures from the CPU oracle (fp32).  The reference ships no golden
vectors (parity unpinned, see oracle/__init__.py), so these pin the ORACLE against regressions and give the
GPU tests fixed, reference-independent inputs.     python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import OracleYOLO  # noqa: E402
from oracle.model import build_model  # noqa: E402
from yolo_puncture_b200 import synth  # noqa: E402


def predict_fixture():
    name = "yolov8n-seg"
    net = build_model(name)
    sd = synth.synth_state_dict([(k, v.shape) for k, v in net.state_dict().items()], name)
    yolo = OracleYOLO(name, state_dict=sd)
    frames = synth.synth_frames(3)
    res = yolo.predict(frames, conf=0.25, iou=0.7, retina_masks=True)
    out = {}
    for i, r in enumerate(res):
        out[f"boxes{i}"] = r.boxes.data.numpy().astype(np.float32)
        out[f"keep{i}"] = yolo.last["kept_idx"][i].numpy().astype(np.int32)
        out[f"area{i}"] = (r.masks.data.sum((1, 2)).numpy().astype(np.int64) if r.masks is not None else np.zeros(0, np.int64))
    np.savez_compressed(os.path.join(HERE, "predict_yolov8n-seg.npz"), **out)
    print({k: v.shape for k, v in out.items()})


def nms_fixture():
    g = torch.Generator().manual_seed(7)
    B, N = 3, 700
    ctr = torch.rand(B, N, 2, generator=g) * 600 + 20
    wh = torch.rand(B, N, 2, generator=g) * 150 + 10
    boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], -1)
    scores = (torch.rand(B, N, generator=g) * 200).round() / 256 + 0.2  # quantised: plenty of exact ties
    cls = torch.randint(0, 4, (B, N), generator=g, dtype=torch.int32)
    keeps = []
    for b in range(B):
        k = torchvision.ops.nms(boxes[b] + cls[b].float()[:, None] * 7680, scores[b], 0.7)[:300]
        keeps.append(np.pad(k.numpy().astype(np.int32), (0, 300 - len(k)), constant_values=-1))
    np.savez_compressed(os.path.join(HERE, "nms_case.npz"), boxes=boxes.numpy(), scores=scores.numpy(), cls=cls.numpy(),
                        keep=np.stack(keeps))
    print("nms kept", [(k >= 0).sum() for k in keeps])


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    predict_fixture()
    nms_fixture()
