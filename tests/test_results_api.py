"""The Results / Boxes / Masks surface the reference consumes, on CPU tensors."""
import numpy as np
import torch

from yolo_puncture_b200.results import Results


def _make():
    img = np.zeros((100, 200, 3), np.uint8)
    boxes = torch.tensor([[10.0, 20.0, 110.0, 60.0, 0.9, 3.0], [0.0, 0.0, 50.0, 50.0, 0.5, 1.0]])
    masks = torch.zeros((2, 100, 200), dtype=torch.uint8)
    masks[0, 20:60, 10:110] = 1
    masks[1, 0:50, 0:50] = 1
    return Results(img, None, {i: str(i) for i in range(80)}, boxes=boxes, masks=masks)


def test_boxes_surface_like_app_py():
    r = _make()
    pb = r.boxes.cpu().numpy()  # reference yolo_seg/app.py:92
    assert len(pb.cls) == 2 and isinstance(pb.conf, np.ndarray)
    best = int(np.argmax(pb.conf))
    assert list(map(int, pb.xyxy[best].squeeze())) == [10, 20, 110, 60]
    assert abs(r.boxes.conf[0].item() - 0.9) < 1e-6 and r.boxes.cls[0].item() == 3.0  # reference yolo_with_deva.py:82-83
    xywhn = r.boxes.xywhn[0]  # reference cls_bbox_dataset_generate.py:52
    assert torch.allclose(xywhn, torch.tensor([60 / 200, 40 / 100, 100 / 200, 40 / 100]))
    assert r.orig_shape == (100, 200) and len(r) == 2


def test_masks_surface_like_yolo_with_deva():
    r = _make()
    assert len(r.masks) == 2
    m = r.masks.data[0]  # reference yolo_with_deva.py:64
    assert m.dtype == torch.float32 and m.shape == (100, 200) and m.sum() == 4000 and bool(((m > 0.5) == (m == 1)).all())
    assert r.masks.raw.dtype == torch.uint8
    poly = r.masks.xy[0]  # reference yolo_seg/app.py:50,101
    assert poly.dtype == np.float32 and poly.shape[1] == 2
    assert poly[:, 0].min() == 10 and poly[:, 0].max() == 109 and poly[:, 1].min() == 20 and poly[:, 1].max() == 59
    assert r.masks.xyn[0].max() <= 1.0


def test_empty_results():
    img = np.zeros((10, 10, 3), np.uint8)
    r = Results(img, None, {}, boxes=torch.zeros((0, 6)), masks=None)
    assert r.masks is None and len(r.boxes.cls) == 0 and len(r) == 0  # reference yolo_with_deva.py:61
