"""Hand-derived known-answer vectors for the oracle (the reference ships no golden vectors and its
arithmetic lives in the un-vendored `ultralytics` package, reference pyproject.toml:23 / yolo_seg/app.py:45-50).
Every expected value below was worked out on paper from the published definitions (DFL = softmax expectation over 16
bins, dist2bbox, bilinear align_corners=False, crop r >= x1 & r < x2, threshold > 0) - none is produced by the oracle."""
import math

import torch

from oracle import ops
from oracle.modules import Detect

NEG = -1.0e4  # exp(NEG - max) underflows to exactly 0 in fp32


def _side(**bins):
    """16 DFL logits: NEG everywhere except the given {bin: logit}."""
    v = [NEG] * 16
    for k, x in bins.items():
        v[int(k[1:])] = x
    return v


def test_dfl_dist2bbox_two_anchor_known_answer():
    # one level, a 1 x 2 map, stride 8, nc = 1: anchors at (0.5, 0.5) and (1.5, 0.5)
    head = Detect(nc=1, ch=(16,))
    assert head.stride.tolist() == [8.0]
    a0 = (_side(b2=0.0, b4=math.log(3.0))      # l: p = (1/4 @ 2, 3/4 @ 4)        -> 3.5
          + _side(b1=0.0)                      # t: one-hot @ 1                   -> 1
          + _side(b2=5.0)                      # r: one-hot @ 2                   -> 2
          + [0.25] * 16)                       # b: uniform over 0..15            -> 7.5
    a1 = (_side(b0=1.0)                        # l -> 0
          + _side(b0=2.0, b1=2.0)              # t: (1/2 @ 0, 1/2 @ 1)            -> 0.5
          + _side(b15=0.0)                     # r -> 15
          + _side(b1=-3.0, b2=-3.0))           # b: (1/2 @ 1, 1/2 @ 2)            -> 1.5
    cls = [0.0, math.log(3.0)]                 # sigmoid -> 0.5, 0.75
    m = torch.tensor([a0 + [cls[0]], a1 + [cls[1]]], dtype=torch.float32).t().reshape(1, 65, 1, 2)
    with torch.no_grad():
        y = head._inference([m])
    assert y.shape == (1, 5, 2)
    # anchor 0: x1 = 0.5 - 3.5 = -3, y1 = 0.5 - 1 = -0.5, x2 = 0.5 + 2 = 2.5, y2 = 0.5 + 7.5 = 8
    #           xywh = ((-3 + 2.5)/2, (-0.5 + 8)/2, 5.5, 8.5) * 8 = (-2, 30, 44, 68)
    # anchor 1: x1 = 1.5 - 0 = 1.5, y1 = 0.5 - 0.5 = 0, x2 = 1.5 + 15 = 16.5, y2 = 0.5 + 1.5 = 2
    #           xywh = (9, 1, 15, 2) * 8 = (72, 8, 120, 16)
    exp = torch.tensor([[-2.0, 30.0, 44.0, 68.0, 0.5], [72.0, 8.0, 120.0, 16.0, 0.75]]).t()
    assert torch.allclose(y[0], exp, rtol=0, atol=2e-5), y[0]
    # and through xywh2xyxy (the first step of NMS): (-24, -4, 20, 64) and (12, 0, 132, 16)
    xyxy = ops.xywh2xyxy(y[0, :4].t())
    assert torch.allclose(xyxy, torch.tensor([[-24.0, -4.0, 20.0, 64.0], [12.0, 0.0, 132.0, 16.0]]), rtol=0, atol=4e-5)


def test_dfl_end2end_xyxy_known_answer():
    # v10 heads decode with xywh=False: box = (anchor - lt, anchor + rb) * stride, here stride 16 at level 1 of 2
    head = Detect(nc=1, ch=(16, 16))
    head.end2end = True
    lvl0 = torch.zeros(1, 65, 1, 1)            # uniform bins: every side 7.5 -> (0.5 -/+ 7.5) * 8 = (-56, -56, 64, 64)
    a = _side(b3=0.0) + _side(b0=0.0) + _side(b1=0.0, b2=0.0) + _side(b10=0.0) + [0.0]   # l 3, t 0, r 1.5, b 10
    lvl1 = torch.tensor(a, dtype=torch.float32).reshape(1, 65, 1, 1)
    with torch.no_grad():
        y = head._inference([lvl0, lvl1])
    exp = torch.tensor([[-56.0, -56.0, 64.0, 64.0, 0.5], [(0.5 - 3) * 16, 0.5 * 16, (0.5 + 1.5) * 16, (0.5 + 10) * 16, 0.5]]).t()
    assert torch.allclose(y[0], exp, rtol=0, atol=3e-5), y[0]


def test_process_mask_native_4x4_proto_known_answer():
    # logits L = 1 * P0 - 2 * P1 on a 4 x 4 proto, upsampled x2 (bilinear, align_corners=False), cropped, > 0.
    # L rows: (1, 1, -1, -1) twice, then (-3, -3, 3, 3) twice.  Output pixel i samples source 0.5 i - 0.25 (clamped):
    # weights (1,0) (.75,.25) (.25,.75) | (.75,.25) (.25,.75) | (.75,.25) (.25,.75) (0,1) over source pairs (0,1) (1,2) (2,3).
    #   horizontal: (a, a, b, b) -> (a, a, a, .75a+.25b, .25a+.75b, b, b, b):
    #       R0 = ( 1,  1,  1,  0.5, -0.5, -1, -1, -1)      R1 = (-3, -3, -3, -1.5,  1.5,  3,  3,  3)
    #   vertical: rows (R0, R0, R0, .75 R0 + .25 R1, .25 R0 + .75 R1, R1, R1, R1):
    #       row 3 = 0 everywhere EXACTLY (0.75 - 0.75, 0.375 - 0.375, ...) -> not > 0
    #       row 4 = (-2, -2, -2, -1, 1, 2, 2, 2)
    #   > 0: rows 0-2 -> columns 0..3; row 3 -> none; rows 4-7 -> columns 4..7
    #   crop box (x1, y1, x2, y2) = (1, 0.5, 6.5, 7): columns 1..6 (r >= 1, r < 6.5), rows 1..6 (c >= 0.5, c < 7)
    L = torch.tensor([[1.0, 1, -1, -1], [1, 1, -1, -1], [-3, -3, 3, 3], [-3, -3, 3, 3]])
    p1 = torch.arange(16, dtype=torch.float32).view(4, 4) / 4
    protos = torch.stack([L + 2 * p1, p1])
    out = ops.process_mask_native(protos, torch.tensor([[1.0, -2.0]]), torch.tensor([[1.0, 0.5, 6.5, 7.0]]), (8, 8))
    exp = torch.zeros(8, 8)
    exp[1:3, 1:4] = 1
    exp[4:7, 4:7] = 1
    assert out.shape == (1, 8, 8) and torch.equal(out[0], exp), out[0]
    assert int(out.sum()) == 15
    # un-cropped logits, checked value by value on the two interesting rows
    lg = ops.scale_masks((torch.tensor([[1.0, -2.0]]) @ protos.view(2, -1)).view(1, 1, 4, 4), (8, 8))[0, 0]
    assert torch.equal(lg[3], torch.zeros(8))
    assert torch.equal(lg[4], torch.tensor([-2.0, -2, -2, -1, 1, 2, 2, 2]))
    assert torch.equal(lg[0], torch.tensor([1.0, 1, 1, 0.5, -0.5, -1, -1, -1]))


def test_process_mask_non_retina_crops_in_proto_space_known_answer():
    # process_mask: crop at proto resolution with boxes scaled by (mw / iw, mh / ih), then upsample and threshold.
    # 4 x 4 proto of ones for an 8 x 8 input, box (2, 2, 6, 6) px -> (1, 1, 3, 3) in proto cells -> cells 1..2 kept.
    # Upsampling the cropped map [0, 1, 1, 0] by 2: (0, .25, .75, 1, 1, .75, .25, 0) per axis; the product is > 0 on 1..6.
    protos = torch.ones(1, 4, 4)
    out = ops.process_mask(protos, torch.tensor([[1.0]]), torch.tensor([[2.0, 2.0, 6.0, 6.0]]), (8, 8), upsample=True)
    exp = torch.zeros(8, 8)
    exp[1:7, 1:7] = 1
    assert torch.equal(out[0], exp)


def test_nms_class_offset_and_order_known_answer():
    # rows (cx, cy, w, h, s0, s1): A and B overlap at IoU 0.6 (area 100 each, intersection 75) and share class 0 ->
    # B suppressed at iou 0.5, kept at iou 0.7; C is the same box as A but class 1 -> the 7680 px class offset separates them.
    pred = torch.tensor([[10.0, 10, 10, 10, 0.9, 0.1],     # A  xyxy (5, 5, 15, 15)   class 0  0.9
                         [12.5, 10, 10, 10, 0.8, 0.2],     # B  xyxy (7.5, 5, 17.5, 15) class 0  0.8
                         [10.0, 10, 10, 10, 0.3, 0.7],     # C  xyxy (5, 5, 15, 15)   class 1  0.7
                         [50.0, 50, 4, 4, 0.2, 0.1]]).t()[None]   # D  below conf
    out, idx = ops.non_max_suppression(pred, 0.25, 0.5, nc=2, return_idx=True)
    assert idx[0].tolist() == [0, 2]
    assert torch.equal(out[0], torch.tensor([[5.0, 5, 15, 15, 0.9, 0], [5.0, 5, 15, 15, 0.7, 1]]))
    out, idx = ops.non_max_suppression(pred, 0.25, 0.7, nc=2, return_idx=True)
    assert idx[0].tolist() == [0, 1, 2]
    out, idx = ops.non_max_suppression(pred, 0.25, 0.5, nc=2, agnostic=True, return_idx=True)
    assert idx[0].tolist() == [0]                               # without the offset C (IoU 1 with A) goes too
    out, idx = ops.non_max_suppression(pred, 0.25, 0.5, nc=2, classes=[1], return_idx=True)
    assert idx[0].tolist() == [2]
