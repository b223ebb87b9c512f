"""Oracle predict() (test infrastructure; see oracle/__init__.py).

Restates UPSTREAM `engine/model.py::Model.predict`, `engine/predictor.py::BasePredictor`,
`models/yolo/{detect,segment}/predict.py::postprocess` and `engine/results.py` as
consumed at reference yolo_seg/app.py:49-50,91-101 and yolo_seg/yolo_with_deva.py:51-83.
"""

import time

import numpy as np
import torch

from . import ops
from .model import build_model


class _Base:
    def __init__(self, data, orig_shape):
        self.data, self.orig_shape = data, orig_shape

    def __len__(self):
        return len(self.data)

    def __getitem__(self, i):
        return self.__class__(self.data[i], self.orig_shape)

    def cpu(self):
        return self if isinstance(self.data, np.ndarray) else self.__class__(self.data.cpu(), self.orig_shape)

    def numpy(self):
        return self if isinstance(self.data, np.ndarray) else self.__class__(self.data.numpy(), self.orig_shape)

    def to(self, *a, **k):
        return self.__class__(torch.as_tensor(self.data).to(*a, **k), self.orig_shape)


class Boxes(_Base):
    """(n,6) [x1,y1,x2,y2,conf,cls] in original-frame pixels (UPSTREAM results.Boxes)."""

    @property
    def xyxy(self):
        return self.data[:, :4]

    @property
    def conf(self):
        return self.data[:, -2]

    @property
    def cls(self):
        return self.data[:, -1]

    @property
    def xywh(self):
        b = self.xyxy
        lib = np if isinstance(b, np.ndarray) else torch
        y = lib.empty_like(b)
        y[..., 0] = (b[..., 0] + b[..., 2]) / 2
        y[..., 1] = (b[..., 1] + b[..., 3]) / 2
        y[..., 2] = b[..., 2] - b[..., 0]
        y[..., 3] = b[..., 3] - b[..., 1]
        return y

    @property
    def xyxyn(self):
        b = self.xyxy.copy() if isinstance(self.data, np.ndarray) else self.xyxy.clone()
        b[..., [0, 2]] /= self.orig_shape[1]
        b[..., [1, 3]] /= self.orig_shape[0]
        return b

    @property
    def xywhn(self):
        b = self.xywh
        b[..., [0, 2]] /= self.orig_shape[1]
        b[..., [1, 3]] /= self.orig_shape[0]
        return b


class Masks(_Base):
    """(n,h,w) float {0,1} masks (UPSTREAM results.Masks)."""

    @property
    def xy(self):
        return [ops.scale_coords(self.data.shape[1:], x, self.orig_shape, normalize=False)
                for x in ops.masks2segments(torch.as_tensor(self.data))]

    @property
    def xyn(self):
        return [ops.scale_coords(self.data.shape[1:], x, self.orig_shape, normalize=True)
                for x in ops.masks2segments(torch.as_tensor(self.data))]


class Results:
    def __init__(self, orig_img, path, names, boxes=None, masks=None, speed=None):
        self.orig_img, self.path, self.names = orig_img, path, names
        self.orig_shape = orig_img.shape[:2]
        self.boxes = Boxes(boxes, self.orig_shape) if boxes is not None else None
        self.masks = Masks(masks, self.orig_shape) if masks is not None else None
        self.speed = speed or {}

    def __len__(self):
        return len(self.boxes) if self.boxes is not None else 0


class OracleYOLO:
    """`YOLO(spec)`-alike around the oracle model; predict() returns list[Results]."""

    def __init__(self, model="yolov8n-seg", nc=80, state_dict=None):
        self.net = build_model(model, nc) if isinstance(model, str) else model
        if state_dict is not None:
            self.net.load_state_dict(state_dict)
        self.net.fuse()
        self.names = self.net.names
        self.task = self.net.task
        self.last = {}

    @property
    def model(self):
        return self.net

    @torch.no_grad()
    def predict(self, source=None, conf=0.25, iou=0.7, retina_masks=False, imgsz=640, max_det=300,
                classes=None, agnostic_nms=False, round_pad=False, **ignored):
        frames = source if isinstance(source, (list, tuple)) else [source]
        frames = [np.asarray(f) for f in frames]
        t0 = time.perf_counter()
        im = ops.preprocess(frames, imgsz)
        t1 = time.perf_counter()
        out = self.net(im)
        t2 = time.perf_counter()
        results = self.postprocess(out, im, frames, conf, iou, retina_masks, max_det, classes, agnostic_nms, round_pad)
        t3 = time.perf_counter()
        n = len(frames)
        for r in results:
            r.speed = {"preprocess": (t1 - t0) * 1e3 / n, "inference": (t2 - t1) * 1e3 / n,
                       "postprocess": (t3 - t2) * 1e3 / n}
        return results

    __call__ = predict

    def postprocess(self, out, im, frames, conf, iou, retina_masks, max_det, classes, agnostic_nms, round_pad=False):
        head = self.net.model[-1]
        results = []
        if self.task == "segment":
            pred, (maps, mc, proto) = out
            self.last = {"pred": pred, "proto": proto}
            dets, kept = ops.non_max_suppression(pred, conf, iou, classes, agnostic_nms, max_det, nc=head.nc,
                                                 return_idx=True)
            self.last["kept_idx"] = kept
            for i, (det, frame) in enumerate(zip(dets, frames)):
                det = det.clone()
                if not len(det):
                    masks = None
                elif retina_masks:
                    det[:, :4] = ops.scale_boxes(im.shape[2:], det[:, :4], frame.shape)
                    masks = ops.process_mask_native(proto[i], det[:, 6:], det[:, :4], frame.shape[:2], round_pad)
                else:
                    masks = ops.process_mask(proto[i], det[:, 6:], det[:, :4], im.shape[2:], upsample=True)
                    det[:, :4] = ops.scale_boxes(im.shape[2:], det[:, :4], frame.shape)
                results.append(Results(frame, None, self.names, boxes=det[:, :6], masks=masks))
        else:
            pred = out[0] if isinstance(out, (list, tuple)) else out
            self.last = {"pred": pred}
            dets = ops.non_max_suppression(pred, conf, iou, classes, agnostic_nms, max_det, nc=head.nc,
                                           end2end=getattr(head, "end2end", False))
            for det, frame in zip(dets, frames):
                det = det.clone()
                det[:, :4] = ops.scale_boxes(im.shape[2:], det[:, :4], frame.shape)
                results.append(Results(frame, None, self.names, boxes=det[:, :6]))
        return results
