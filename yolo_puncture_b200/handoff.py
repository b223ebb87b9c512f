"""Index-mask hand-off to the tracker.

Drop-in for the per-detection loop of reference yolo_seg/yolo_with_deva.py:54-88 (`auto_segment`): from the `Results` of
`YOLO.predict(..., retina_masks=True)` build, per frame, the int64 (H, W) index mask DEVA consumes (pixel = 1-based id of
the last kept detection covering it, 0 = background) and the `(id, score, category_id)` list — three kernel launches
for a whole batch of frames (`ypb_index_masks`) instead of a `.sum()` host sync and a boolean scatter per detection.
"""

import ctypes as C

import numpy as np
import torch

from ._lib import check, lib


def index_masks(results, suppress_small_mask=True, min_area=100, out_shape=None):
    """results: list of `Results` (masks at frame size, i.e. predict(retina_masks=True)).
    Returns a list of (index_mask int64 (H, W) device tensor, segments_info list of dicts {id, score, category_id}),
    one per frame, with the reference's semantics: detections in order; `mask.sum() < min_area` ones are skipped when
    suppress_small_mask; kept ones get ids 1, 2, ...; later detections overwrite earlier ones.
    out_shape=(h, w): the reference's `min_side` branch - predict() ran on a resized copy of the frame, every mask is
    resized back to (h, w) the way `F.resize(mask[None], [h, w])` does before the area filter and the paint."""
    results = list(results)
    if not results:
        return []
    if out_shape is not None and any(tuple(out_shape) != tuple(r.orig_shape) for r in results):
        return _index_masks_resized(results, suppress_small_mask, min_area, tuple(int(v) for v in out_shape))
    raws, counts = [], []
    dev = None
    for r in results:
        m = r.masks.raw if r.masks is not None else None
        if m is not None:
            if not torch.is_tensor(m) or not m.is_cuda or m.dtype != torch.uint8:
                raise ValueError("index_masks needs the device-resident uint8 masks of YOLO.predict()")
            if tuple(m.shape[1:]) != tuple(r.orig_shape):
                raise ValueError("index_masks needs masks at frame size: call predict(retina_masks=True) on unresized frames")
            dev = m.device
        raws.append(m)
        counts.append(0 if m is None else int(m.shape[0]))
    shapes = {tuple(r.orig_shape) for r in results}
    if len(shapes) != 1:
        raise ValueError("index_masks: all frames of a call must have one size")
    H, W = shapes.pop()
    if dev is None:  # no detection in any frame
        ref = results[0].boxes.data if results[0].boxes is not None and torch.is_tensor(results[0].boxes.data) else None
        dev = ref.device if ref is not None and ref.is_cuda else torch.device("cuda")
        return [(torch.zeros((H, W), dtype=torch.int64, device=dev), []) for _ in results]
    B, n_total = len(results), sum(counts)
    # masks of one predict() call are consecutive slices of one buffer: use it in place, otherwise gather once
    present = [m for m in raws if m is not None]
    contiguous = all(m.is_contiguous() for m in present) and all(
        b.data_ptr() == a.data_ptr() + a.numel() for a, b in zip(present, present[1:]))
    base = present[0] if contiguous else torch.cat(present)
    # (conf, cls) of every detection from the host copy predict() already fetched with the pass (no device gather), and -
    # when every mask is known to be cropped to its box (the masks of predict()) - the boxes as pixel rectangles, so that
    # the area sum and the paint only touch the boxes
    hb = [_host_boxes(r) for r, c in zip(results, counts) if c]
    allb = np.concatenate(hb) if len(hb) > 1 else hb[0]
    if allb.shape[0] != n_total:
        raise ValueError("index_masks: boxes and masks of a frame differ in length")
    cropped = all(getattr(r.masks, "cropped", False) for r, c in zip(results, counts) if c)
    r_off = (B + 1 + 3) & ~3  # the rectangles are read as int4: 16-byte aligned behind the offsets
    stage = _pinned_i32(r_off + 4 * n_total)
    sn = stage.numpy()
    sn[0] = 0
    np.cumsum(counts, out=sn[1:B + 1])
    if cropped:  # mask on  <=>  bx1 <= x < bx2 and by1 <= y < by2 (mask_decode_kernel): [floor(b1), ceil(b2)) covers it
        rc = sn[r_off:r_off + 4 * n_total].reshape(n_total, 4)
        rc[:, 0] = np.clip(np.floor(allb[:, 0]), 0, W)
        rc[:, 1] = np.clip(np.floor(allb[:, 1]), 0, H)
        rc[:, 2] = np.clip(np.ceil(allb[:, 2]), 0, W)
        rc[:, 3] = np.clip(np.ceil(allb[:, 3]), 0, H)
    with torch.cuda.device(dev):
        n_stage = r_off + 4 * n_total if cropped else B + 1
        stage_d = torch.empty(n_stage, dtype=torch.int32, device=dev)
        stage_d.copy_(stage[:n_stage], non_blocking=True)
        area = torch.empty(n_total, dtype=torch.int32, device=dev)
        ids = torch.empty(n_total, dtype=torch.int32, device=dev)
        index_map = torch.empty((B, H, W), dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        rects_p = stage_d.data_ptr() + 4 * r_off if cropped else None
        check(lib().ypb_index_masks_boxed(C.c_void_p(st), C.c_void_p(base.data_ptr()), C.c_void_p(stage_d.data_ptr()),
                                          C.c_void_p(rects_p), B, n_total, H, W,
                                          int(min_area) if suppress_small_mask else -1, C.c_void_p(area.data_ptr()),
                                          C.c_void_p(ids.data_ptr()), C.c_void_p(index_map.data_ptr())))
        ids_h = ids.cpu().numpy()  # the one host sync of the hand-off
    return _pack_infos(index_map, ids_h, allb, counts)


def _host_boxes(r):
    """(n, 6) float32 numpy rows of a frame's boxes; predict()'s Results carry a host copy fetched with their pass."""
    b = r.boxes
    h = b._host() if getattr(b, "_host", None) is not None else b.data
    if torch.is_tensor(h):
        h = h.detach().cpu().numpy()
    return np.asarray(h, dtype=np.float32).reshape(-1, 6)


_PINNED = {}


def _pinned_i32(n):
    """Page-locked int32 scratch, grown on demand; every use is followed by a host sync before the next one."""
    buf = _PINNED.get("i32")
    if buf is None or buf.numel() < n:
        buf = _PINNED["i32"] = torch.empty(max(n, 4096), dtype=torch.int32).pin_memory()
    return buf


def _pack_infos(index_map, ids_h, allb, counts):
    """[(index_map[b], [{id, score, category_id, index}, ...])]: kept detections only, in order.
    "index" = the detection's row in its frame's Results (beyond the reference's ObjectInfo fields)."""
    keep = np.flatnonzero(ids_h)
    starts = np.zeros(len(counts) + 1, dtype=np.int64)
    np.cumsum(counts, out=starts[1:])
    frame_of = np.searchsorted(starts, keep, side="right") - 1
    cut = np.searchsorted(frame_of, np.arange(len(counts) + 1))
    kid = ids_h[keep].tolist()
    ksc = allb[keep, 4].tolist()
    kcl = allb[keep, 5].astype(np.int64).tolist()
    krow = (keep - starts[frame_of]).tolist()
    infos = [{"id": i, "score": s, "category_id": c, "index": r} for i, s, c, r in zip(kid, ksc, kcl, krow)]
    return [(index_map[b], infos[cut[b]:cut[b + 1]]) for b in range(len(counts))]


def _index_masks_resized(results, suppress_small_mask, min_area, out_shape):
    H, W = out_shape
    shapes = {tuple(r.orig_shape) for r in results}
    if len(shapes) != 1:
        raise ValueError("index_masks: all frames of a call must have one size")
    h1, w1 = shapes.pop()
    raws = [r.masks.raw if r.masks is not None else None for r in results]
    counts = [0 if m is None else int(m.shape[0]) for m in raws]
    present = [m for m in raws if m is not None]
    if any((not torch.is_tensor(m)) or (not m.is_cuda) or m.dtype != torch.uint8 or tuple(m.shape[1:]) != (h1, w1) for m in present):
        raise ValueError("index_masks needs the device-resident uint8 retina masks of YOLO.predict()")
    B, n_total = len(results), sum(counts)
    if not present:
        ref = results[0].boxes.data if results[0].boxes is not None and torch.is_tensor(results[0].boxes.data) else None
        dev = ref.device if ref is not None and ref.is_cuda else torch.device("cuda")
        return [(torch.zeros((H, W), dtype=torch.int64, device=dev), []) for _ in results]
    dev = present[0].device
    contiguous = all(m.is_contiguous() for m in present) and all(
        b.data_ptr() == a.data_ptr() + a.numel() for a, b in zip(present, present[1:]))
    base = present[0] if contiguous else torch.cat(present)
    offsets = torch.zeros(B + 1, dtype=torch.int32)
    offsets[1:] = torch.tensor(counts, dtype=torch.int32).cumsum(0)
    with torch.cuda.device(dev):
        offs_d = offsets.to(dev, non_blocking=True)
        bins = torch.empty((n_total, H, W), dtype=torch.uint8, device=dev)
        area = torch.empty(n_total, dtype=torch.float32, device=dev)
        ids = torch.empty(n_total, dtype=torch.int32, device=dev)
        index_map = torch.empty((B, H, W), dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        check(lib().ypb_index_masks_resized(C.c_void_p(st), C.c_void_p(base.data_ptr()), C.c_void_p(offs_d.data_ptr()), B, n_total,
                                            h1, w1, H, W, float(min_area) if suppress_small_mask else -1.0,
                                            C.c_void_p(bins.data_ptr()), C.c_void_p(area.data_ptr()), C.c_void_p(ids.data_ptr()),
                                            C.c_void_p(index_map.data_ptr())))
        ids_h = ids.cpu().numpy()
    hb = [_host_boxes(r) for r, c in zip(results, counts) if c]
    return _pack_infos(index_map, ids_h, np.concatenate(hb) if len(hb) > 1 else hb[0], counts)


def auto_segment(config, image, yolo_model, min_side, suppress_small_mask):
    """Drop-in for reference yolo_seg/yolo_with_deva.py:35-88 `auto_segment(config, image, yolo_model, min_side,
    suppress_small_mask)`: optional `min_side` resize of the frame (cv2.resize, as the reference), predict(retina_masks=True,
    conf=0.9), then the index mask and the segment list in three kernel launches.  Returns (int64 (h, w) device tensor,
    [{"id", "score", "category_id"}, ...]) - the fields of the reference's ObjectInfo."""
    h, w = image.shape[:2]
    if min_side > 0:
        import cv2
        scale = min_side / min(h, w)
        image = cv2.resize(image, (int(w * scale), int(h * scale)))
    results = yolo_model.predict(image, retina_masks=True, conf=0.9)
    (index_mask, info), = index_masks(results[:1], suppress_small_mask, config.get("MIN_AREA_THRESHOLD", 100), out_shape=(h, w))
    return index_mask, info


def min_rect_len(masks):
    """Device replacement of reference utils/mask_tools.py:12-22 `get_coord_min_rect_len(masks.xy[i])` for every mask of
    a `Masks` object (or an (n, H, W) uint8 CUDA tensor): returns an (n, 2) float32 device tensor of
    (length = long side of the minimum-area rectangle, length / max(short side, 1)) without copying the masks to the
    host or running cv2.findContours (consumer: yolo_seg/app.py:97-105)."""
    m = masks.raw if hasattr(masks, "raw") else masks
    if not torch.is_tensor(m) or not m.is_cuda or m.dtype != torch.uint8 or m.dim() != 3:
        raise ValueError("min_rect_len needs device-resident (n, H, W) uint8 masks")
    m = m.contiguous()
    n, H, W = m.shape
    with torch.cuda.device(m.device):
        out = torch.empty((n, 2), dtype=torch.float32, device=m.device)
        ext = torch.empty((max(n, 1), H, 2), dtype=torch.int32, device=m.device)
        st = torch.cuda.current_stream(m.device).cuda_stream
        check(lib().ypb_mask_min_rect(C.c_void_p(st), C.c_void_p(m.data_ptr()), n, H, W, C.c_void_p(ext.data_ptr()),
                                      C.c_void_p(out.data_ptr())))
    return out
