"""`Results` / `Boxes` / `Masks` containers with the ultralytics surface the reference consumes.

Mirrors UPSTREAM engine/results.py for exactly the attributes yolo_seg touches:
  results[0].boxes.cpu().numpy() -> .cls .conf .xyxy            reference yolo_seg/app.py:92-98
  results[0].masks.xy[i]                                        reference yolo_seg/app.py:50,101; dev_tools/auto_speed_calc.py:71
  len(masks), masks.data[i] (float-compatible), boxes.conf[i].item(), boxes.cls[i].item()
                                                                reference yolo_seg/yolo_with_deva.py:61-83
  boxes.xywhn[0]                                                reference dev_tools/classify/cls_bbox_dataset_generate.py:52
Masks stay uint8 {0,1} on the device (1 B/pixel instead of upstream's 4) and are converted to float32
only when `.data` is read — `F.resize(mask)` / `mask > 0.5` / `mask.sum()` in auto_segment need a float tensor.
"""

import numpy as np
import torch


class BaseTensor:
    def __init__(self, data, orig_shape, lazy=None, n=None, host=None):
        """data: tensor / ndarray, or None with `lazy` = a zero-argument callable producing it on first use (predict()
        hands out views of batch-level tensors: slicing 64 frames x {boxes, masks} eagerly costs more host time than a
        B200 needs for the frames themselves).  n: row count when known without materialising.  host: zero-argument
        callable returning an already-fetched CPU copy (predict() reads every box of a pass back in ONE copy)."""
        self._d = data
        self._lazy = lazy
        self._n = n
        self._host = host
        self.orig_shape = orig_shape

    @property
    def _data(self):
        if self._d is None and self._lazy is not None:
            self._d = self._lazy()
            self._lazy = None
        return self._d

    @property
    def data(self):
        return self._data

    @property
    def shape(self):
        return self.data.shape

    def __len__(self):
        return self._n if self._n is not None else len(self._data)

    def __getitem__(self, idx):
        d = self._data[idx]
        if d.ndim < self._data.ndim:
            d = d[None]
        return self.__class__(d, self.orig_shape)

    def cpu(self):
        if self._host is not None:  # fetched with the rest of its engine pass: no device round trip
            return self.__class__(self._host(), self.orig_shape)
        return self if isinstance(self._data, np.ndarray) else self.__class__(self._data.cpu(), self.orig_shape)

    def numpy(self):
        if self._host is not None:
            return self.__class__(self._host().numpy(), self.orig_shape)
        return self if isinstance(self._data, np.ndarray) else self.__class__(self._data.cpu().numpy(), self.orig_shape)

    def cuda(self):
        return self.__class__(torch.as_tensor(self._data).cuda(), self.orig_shape)

    def to(self, *args, **kwargs):
        return self.__class__(torch.as_tensor(self._data).to(*args, **kwargs), self.orig_shape)


class Boxes(BaseTensor):
    """(n,6) fp32 rows [x1,y1,x2,y2,conf,cls] in original-frame pixels, descending confidence."""

    @property
    def xyxy(self):
        return self._data[:, :4]

    @property
    def conf(self):
        return self._data[:, -2]

    @property
    def cls(self):
        return self._data[:, -1]

    @property
    def xywh(self):
        b = self.xyxy
        y = np.empty_like(b) if isinstance(b, np.ndarray) else torch.empty_like(b)
        y[..., 0] = (b[..., 0] + b[..., 2]) / 2
        y[..., 1] = (b[..., 1] + b[..., 3]) / 2
        y[..., 2] = b[..., 2] - b[..., 0]
        y[..., 3] = b[..., 3] - b[..., 1]
        return y

    def _norm(self, b):
        b = b.copy() if isinstance(b, np.ndarray) else b.clone()
        b[..., [0, 2]] /= self.orig_shape[1]
        b[..., [1, 3]] /= self.orig_shape[0]
        return b

    @property
    def xyxyn(self):
        return self._norm(self.xyxy)

    @property
    def xywhn(self):
        return self._norm(self.xywh)


class Masks(BaseTensor):
    """(n,h,w) masks.  Stored uint8 {0,1}; `.data` yields float32 {0.,1.} like upstream."""

    cropped = False  # True on the masks predict() hands out: mask i is zero outside box i (what index_masks exploits)

    @property
    def data(self):
        """float32 view of the masks, converted ONCE and cached: the reference loop `masks.data[i]` per detection
        (yolo_seg/yolo_with_deva.py:61-83) must not convert the whole (n,H,W) tensor on every access."""
        d = self._data
        f = getattr(self, "_f32", None)
        if f is None:
            if isinstance(d, np.ndarray):
                f = d if d.dtype == np.float32 else d.astype(np.float32)
            else:
                f = d if d.dtype == torch.float32 else d.to(torch.float32)
            self._f32 = f
        return f

    @property
    def raw(self):
        """The packed uint8 masks as produced by the kernel (no conversion)."""
        return self._data

    def _segments(self, normalize):
        import cv2
        d = self._data
        m = d if isinstance(d, np.ndarray) else d.cpu().numpy()
        m = np.ascontiguousarray(m.astype(np.uint8))
        mh, mw = m.shape[1:]
        h0, w0 = self.orig_shape
        out = []
        for x in m:
            c = cv2.findContours(x, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0]
            if c:
                c = np.concatenate([p.reshape(-1, 2) for p in c]).astype(np.float32)
            else:
                c = np.zeros((0, 2), np.float32)
            # ops.scale_coords(mask_shape -> orig_shape): identity + clip for retina masks
            gain = min(mh / h0, mw / w0)
            pad = (mw - w0 * gain) / 2, (mh - h0 * gain) / 2
            c[:, 0] = np.clip((c[:, 0] - pad[0]) / gain, 0, w0)
            c[:, 1] = np.clip((c[:, 1] - pad[1]) / gain, 0, h0)
            if normalize:
                c[:, 0] /= w0
                c[:, 1] /= h0
            out.append(c)
        return out

    @property
    def xy(self):
        """Per mask, the (k,2) float32 pixel polygon of its external contours (cv2.findContours on the host)."""
        return self._segments(False)

    @property
    def xyn(self):
        return self._segments(True)


class Results:
    def __init__(self, orig_img, path, names, boxes=None, masks=None, speed=None):
        self.orig_img = orig_img
        self.orig_shape = tuple(orig_img.shape[:2])
        self.path = path
        self.names = names
        self.boxes = boxes if isinstance(boxes, Boxes) else Boxes(boxes, self.orig_shape) if boxes is not None else None
        self.masks = masks if isinstance(masks, Masks) else Masks(masks, self.orig_shape) if masks is not None else None
        self.probs = None
        self.keypoints = None
        self.obb = None
        self.speed = speed or {"preprocess": None, "inference": None, "postprocess": None}

    def __len__(self):
        return len(self.boxes) if self.boxes is not None else 0

    def _apply(self, fn, *a, **k):
        r = Results(self.orig_img, self.path, self.names, speed=self.speed)
        r.boxes = getattr(self.boxes, fn)(*a, **k) if self.boxes is not None else None
        r.masks = getattr(self.masks, fn)(*a, **k) if self.masks is not None else None
        return r

    def cpu(self):
        return self._apply("cpu")

    def numpy(self):
        return self._apply("numpy")

    def cuda(self):
        return self._apply("cuda")

    def to(self, *a, **k):
        return self._apply("to", *a, **k)
