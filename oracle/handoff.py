"""Oracle of the index-mask hand-off (test infrastructure; see oracle/__init__.py).

Restates reference yolo_seg/yolo_with_deva.py:54-88 (`auto_segment`, masks already at frame size): torch CPU, the same
loop, `.sum()` filter and boolean scatter as the reference."""

import torch


def auto_segment_index_mask(masks, conf, cls, suppress_small_mask=True, min_area=100):
    """masks: (n, H, W) float {0,1} or None; conf, cls: (n,).  Returns (int64 (H, W) index mask, list of (id, score, category_id))."""
    if masks is None or len(masks) == 0:
        return None, []
    h, w = masks.shape[1:]
    output_mask = torch.zeros((h, w), dtype=torch.int64)
    segments_info = []
    curr_id = 1
    for i in range(len(masks)):
        mask = masks[i].float()
        if suppress_small_mask and mask.sum() < min_area:
            continue
        output_mask[mask > 0.5] = curr_id
        segments_info.append((curr_id, float(conf[i]), int(cls[i])))
        curr_id += 1
    return output_mask, segments_info


def auto_segment_index_mask_resized(masks, conf, cls, out_hw, suppress_small_mask=True, min_area=100):
    """The `min_side` branch (reference yolo_seg/yolo_with_deva.py:44-48,71-72): masks (n, h1, w1) from a resized frame,
    each brought back to (h, w) with torchvision's `F.resize(mask.unsqueeze(0), size=[h, w])[0]` exactly as the reference does."""
    from torchvision.transforms import functional as F
    if masks is None or len(masks) == 0:
        return None, []
    h, w = out_hw
    output_mask = torch.zeros((h, w), dtype=torch.int64)
    segments_info = []
    curr_id = 1
    for i in range(len(masks)):
        mask = masks[i].float()
        if mask.shape != (h, w):
            mask = F.resize(mask.unsqueeze(0), size=[h, w])[0]
        if suppress_small_mask and mask.sum() < min_area:
            continue
        output_mask[mask > 0.5] = curr_id
        segments_info.append((curr_id, float(conf[i]), int(cls[i])))
        curr_id += 1
    return output_mask, segments_info


def coord_min_rect_len(mask):
    """Reference yolo_seg/app.py:101-102 + utils/mask_tools.py:12-22 for one (H, W) {0,1} mask: external contours
    (cv2.findContours, as `Masks.xy` does), int32 points, cv2.minAreaRect -> (long side, long / max(short, 1))."""
    import cv2
    import numpy as np
    m = np.ascontiguousarray(np.asarray(mask).astype(np.uint8))
    c = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0]
    if not c:
        return 0.0, 0.0
    pts = np.concatenate([p.reshape(-1, 2) for p in c]).astype(np.float32)
    points = np.array(pts, dtype=np.int32).reshape((-1, 2))
    if len(points) < 3:
        return 0.0, 0.0
    (_, (width, height), _) = cv2.minAreaRect(points)
    length = max(width, height)
    width = min(height, width)
    if width == 0:
        width = 1
    return float(length), float(length / width)
