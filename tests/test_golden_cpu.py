"""The oracle reproduces the committed golden fixtures (guards the checker itself against regressions)."""
import os

import numpy as np
import torch
import torchvision

from oracle import OracleYOLO
from oracle.model import build_model
from yolo_puncture_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_predict_reproduces_golden():
    gold = np.load(os.path.join(G, "predict_yolov8n-seg.npz"))
    name = "yolov8n-seg"
    net = build_model(name)
    sd = synth.synth_state_dict([(k, v.shape) for k, v in net.state_dict().items()], name)
    yolo = OracleYOLO(name, state_dict=sd)
    res = yolo.predict(synth.synth_frames(3), conf=0.25, iou=0.7, retina_masks=True)
    for i, r in enumerate(res):
        assert yolo.last["kept_idx"][i].tolist() == gold[f"keep{i}"].tolist()
        np.testing.assert_allclose(r.boxes.data.numpy(), gold[f"boxes{i}"], rtol=1e-4, atol=1e-3)
        if len(gold[f"area{i}"]):
            area = r.masks.data.sum((1, 2)).numpy()
            assert np.abs(area - gold[f"area{i}"]).max() <= 2
        else:
            assert r.masks is None


def test_torchvision_nms_reproduces_golden():
    gold = np.load(os.path.join(G, "nms_case.npz"))
    boxes, scores, cls = (torch.from_numpy(gold[k]) for k in ("boxes", "scores", "cls"))
    for b in range(boxes.shape[0]):
        k = torchvision.ops.nms(boxes[b] + cls[b].float()[:, None] * 7680, scores[b], 0.7)[:300]
        assert k.tolist() == [int(x) for x in gold["keep"][b] if x >= 0]
