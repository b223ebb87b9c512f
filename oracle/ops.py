"""Oracle pre/post-processing (test infrastructure; see oracle/__init__.py).

Restates UPSTREAM ultralytics 8.3.x `data/augment.py::LetterBox`, `engine/predictor.py::preprocess`,
`utils/ops.py::{non_max_suppression, xywh2xyxy, scale_boxes, clip_boxes, process_mask,
process_mask_native, scale_masks, crop_mask, masks2segments, scale_coords}`.
These run inside every `model.predict(...)` at reference yolo_seg/app.py:49,91,
yolo_seg/yolo_with_deva.py:51 and dev_tools/auto_speed_calc.py:62.
"""

import cv2
import numpy as np
import torch
import torch.nn.functional as F
import torchvision


# ------------------------------------------------------------------------------------------------
# preprocessing
# ------------------------------------------------------------------------------------------------
def letterbox_params(shape, new_shape=(640, 640), auto=False, stride=32, scaleup=True):
    """Geometry of LetterBox for a (h, w) frame: returns (new_unpad (w,h), top, bottom, left, right)."""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    if not scaleup:
        r = min(r, 1.0)
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_unpad, top, bottom, left, right


def letterbox(img, new_shape=(640, 640), auto=False, stride=32):
    """LetterBox(center=True, scaleup=True): resize (cv2 INTER_LINEAR) + pad with 114."""
    shape = img.shape[:2]
    new_unpad, top, bottom, left, right = letterbox_params(shape, new_shape, auto, stride)
    if shape[::-1] != new_unpad:
        img = cv2.resize(img, new_unpad, interpolation=cv2.INTER_LINEAR)
    return cv2.copyMakeBorder(img, top, bottom, left, right, cv2.BORDER_CONSTANT, value=(114, 114, 114))


def preprocess(frames, imgsz=640, stride=32):
    """list of uint8 BGR HWC frames -> fp32 (B,3,H,W) RGB in [0,1] (BasePredictor.preprocess).
    `auto` (minimal rectangle) is on iff all frames share one shape, as upstream does for PyTorch models."""
    same = len({f.shape for f in frames}) == 1
    lb = [letterbox(f, (imgsz, imgsz) if isinstance(imgsz, int) else imgsz, auto=same, stride=stride) for f in frames]
    im = np.stack(lb)
    im = im[..., ::-1].transpose((0, 3, 1, 2))
    im = np.ascontiguousarray(im)
    return torch.from_numpy(im).float() / 255


# ------------------------------------------------------------------------------------------------
# selection
# ------------------------------------------------------------------------------------------------
def xywh2xyxy(x):
    y = torch.empty_like(x)
    xy = x[..., :2]
    wh = x[..., 2:] / 2
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        max_det=300, nc=0, max_nms=30000, max_wh=7680, end2end=False, return_idx=False):
    """UPSTREAM ops.non_max_suppression without the wall-clock time limit.
    prediction: (B, 4+nc+nm, A) for NMS models, (B, n, 6) for end2end models.
    With return_idx also returns, per image, the anchor index of every kept row."""
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    if classes is not None:
        classes = torch.tensor(classes, device=prediction.device)
    if prediction.shape[-1] == 6 or end2end:
        out = [pred[pred[:, 4] > conf_thres][:max_det] for pred in prediction]
        if classes is not None:
            out = [pred[(pred[:, 5:6] == classes).any(1)] for pred in out]
        return (out, None) if return_idx else out

    bs = prediction.shape[0]
    nc = nc or (prediction.shape[1] - 4)
    nm = prediction.shape[1] - nc - 4
    mi = 4 + nc
    xc = prediction[:, 4:mi].amax(1) > conf_thres

    prediction = prediction.transpose(-1, -2).clone()
    prediction[..., :4] = xywh2xyxy(prediction[..., :4])

    output = [torch.zeros((0, 6 + nm), device=prediction.device)] * bs
    kept_idx = [torch.zeros((0,), dtype=torch.long)] * bs
    for xi, x in enumerate(prediction):
        aidx = torch.nonzero(xc[xi]).view(-1)
        x = x[xc[xi]]
        if not x.shape[0]:
            continue
        box, cls, mask = x.split((4, nc, nm), 1)
        conf, j = cls.max(1, keepdim=True)
        sel = conf.view(-1) > conf_thres
        x = torch.cat((box, conf, j.float(), mask), 1)[sel]
        aidx = aidx[sel]
        if classes is not None:
            sel = (x[:, 5:6] == classes).any(1)
            x, aidx = x[sel], aidx[sel]
        n = x.shape[0]
        if not n:
            continue
        if n > max_nms:
            order = x[:, 4].argsort(descending=True)[:max_nms]
            x, aidx = x[order], aidx[order]
        c = x[:, 5:6] * (0 if agnostic else max_wh)
        scores = x[:, 4]
        boxes = x[:, :4] + c
        i = torchvision.ops.nms(boxes, scores, iou_thres)
        i = i[:max_det]
        output[xi] = x[i]
        kept_idx[xi] = aidx[i]
    return (output, kept_idx) if return_idx else output


# ------------------------------------------------------------------------------------------------
# geometry
# ------------------------------------------------------------------------------------------------
def clip_boxes(boxes, shape):
    boxes[..., 0] = boxes[..., 0].clamp(0, shape[1])
    boxes[..., 1] = boxes[..., 1].clamp(0, shape[0])
    boxes[..., 2] = boxes[..., 2].clamp(0, shape[1])
    boxes[..., 3] = boxes[..., 3].clamp(0, shape[0])
    return boxes


def scale_boxes(img1_shape, boxes, img0_shape, padding=True):
    """Undo the letterbox on xyxy boxes, in place (UPSTREAM ops.scale_boxes)."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1),
           round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    if padding:
        boxes[..., 0] -= pad[0]
        boxes[..., 1] -= pad[1]
        boxes[..., 2] -= pad[0]
        boxes[..., 3] -= pad[1]
    boxes[..., :4] /= gain
    return clip_boxes(boxes, img0_shape)


def crop_mask(masks, boxes):
    """Zero everything outside each box (UPSTREAM ops.crop_mask): keep r>=x1 & r<x2 & c>=y1 & c<y2."""
    _, h, w = masks.shape
    x1, y1, x2, y2 = torch.chunk(boxes[:, :, None], 4, 1)
    r = torch.arange(w, device=masks.device, dtype=x1.dtype)[None, None, :]
    c = torch.arange(h, device=masks.device, dtype=x1.dtype)[None, :, None]
    return masks * ((r >= x1) * (r < x2) * (c >= y1) * (c < y2))


def scale_masks(masks, shape, padding=True, round_pad=False):
    """(N,C,mh,mw) -> (N,C,*shape): slice off the letterbox pad in mask space, bilinear upsample.
    round_pad=False: `int()` truncation (8.1-8.3 early); True: later 8.3.x `round(pad -/+ 0.1)`."""
    mh, mw = masks.shape[2:]
    gain = min(mh / shape[0], mw / shape[1])
    pad = [mw - shape[1] * gain, mh - shape[0] * gain]
    if padding:
        pad[0] /= 2
        pad[1] /= 2
    if round_pad:
        top, left = (int(round(pad[1] - 0.1)), int(round(pad[0] - 0.1))) if padding else (0, 0)
        bottom, right = mh - int(round(pad[1] + 0.1)), mw - int(round(pad[0] + 0.1))
    else:
        top, left = (int(pad[1]), int(pad[0])) if padding else (0, 0)
        bottom, right = int(mh - pad[1]), int(mw - pad[0])
    masks = masks[..., top:bottom, left:right]
    return F.interpolate(masks, shape, mode="bilinear", align_corners=False)


def process_mask_native(protos, masks_in, bboxes, shape, round_pad=False):
    """Retina masks (UPSTREAM ops.process_mask_native): logits at frame resolution, crop, > 0."""
    c, mh, mw = protos.shape
    masks = (masks_in @ protos.float().view(c, -1)).view(-1, mh, mw)
    masks = scale_masks(masks[None], shape, round_pad=round_pad)[0]
    masks = crop_mask(masks, bboxes)
    return masks.gt_(0.0)


def process_mask(protos, masks_in, bboxes, shape, upsample=True):
    """Non-retina masks (UPSTREAM ops.process_mask): crop in proto space, upsample to net input."""
    c, mh, mw = protos.shape
    ih, iw = shape
    masks = (masks_in @ protos.float().view(c, -1)).view(-1, mh, mw)
    width_ratio, height_ratio = mw / iw, mh / ih
    db = bboxes.clone()
    db[:, 0] *= width_ratio
    db[:, 2] *= width_ratio
    db[:, 3] *= height_ratio
    db[:, 1] *= height_ratio
    masks = crop_mask(masks, db)
    if upsample:
        masks = F.interpolate(masks[None], shape, mode="bilinear", align_corners=False)[0]
    return masks.gt_(0.0)


def scale_coords(img1_shape, coords, img0_shape, normalize=False, padding=True):
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    if padding:
        coords[..., 0] -= pad[0]
        coords[..., 1] -= pad[1]
    coords[..., 0] /= gain
    coords[..., 1] /= gain
    coords[..., 0] = coords[..., 0].clip(0, img0_shape[1])
    coords[..., 1] = coords[..., 1].clip(0, img0_shape[0])
    if normalize:
        coords[..., 0] /= img0_shape[1]
        coords[..., 1] /= img0_shape[0]
    return coords


def masks2segments(masks, strategy="all"):
    """(n,h,w) {0,1} masks -> list of (k,2) float32 polygons via cv2.findContours
    (UPSTREAM ops.masks2segments; 'all' concatenates every external contour, 'largest' keeps one)."""
    segments = []
    for x in masks.int().cpu().numpy().astype("uint8"):
        c = cv2.findContours(x, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0]
        if c:
            if strategy == "all":
                c = np.concatenate([x.reshape(-1, 2) for x in c])
            else:
                c = np.array(c[np.array([len(x) for x in c]).argmax()]).reshape(-1, 2)
        else:
            c = np.zeros((0, 2))
        segments.append(c.astype("float32"))
    return segments
