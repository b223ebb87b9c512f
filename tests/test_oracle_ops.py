"""Known-answer micro-cases for the oracle's pre/post-processing."""
import numpy as np
import torch

from oracle import ops
from oracle.modules import Detect


def test_letterbox_1080p_at_1280_rect():
    # 1920x1080 @ imgsz 1280, auto -> 1280x720 resized, pad 8 + 8 -> (736, 1280)
    new_unpad, top, bottom, left, right = ops.letterbox_params((1080, 1920), (1280, 1280), auto=True)
    assert new_unpad == (1280, 720) and (top, bottom, left, right) == (8, 8, 0, 0)
    img = np.zeros((1080, 1920, 3), np.uint8)
    out = ops.letterbox(img, (1280, 1280), auto=True)
    assert out.shape == (736, 1280, 3) and out[0, 0, 0] == 114 and out[8, 0, 0] == 0 and out[727, 0, 0] == 0
    assert out[728, 0, 0] == 114


def test_letterbox_square_pad_rounding():
    # odd total padding: dh = 1.5 -> top = round(1.4) = 1, bottom = round(1.6) = 2
    new_unpad, top, bottom, left, right = ops.letterbox_params((637, 640), (640, 640), auto=False)
    assert new_unpad == (640, 637) and (top, bottom) == (1, 2) and (left, right) == (0, 0)


def test_preprocess_bgr_to_rgb_chw_unit_range():
    f = np.zeros((640, 640, 3), np.uint8)
    f[..., 0], f[..., 1], f[..., 2] = 255, 128, 0  # B, G, R
    im = ops.preprocess([f], 640)
    assert im.shape == (1, 3, 640, 640) and im.dtype == torch.float32
    assert im[0, 0, 0, 0] == 0.0 and abs(im[0, 1, 0, 0] - 128 / 255) < 1e-7 and im[0, 2, 0, 0] == 1.0


def test_scale_boxes_undoes_letterbox():
    boxes = torch.tensor([[0.0, 8.0, 1280.0, 728.0], [640.0, 368.0, 2000.0, 900.0]])
    out = ops.scale_boxes((736, 1280), boxes.clone(), (1080, 1920))
    assert torch.allclose(out[0], torch.tensor([0.0, 0.0, 1920.0, 1080.0]))
    assert torch.allclose(out[1], torch.tensor([960.0, 540.0, 1920.0, 1080.0]))  # clipped


def test_crop_mask_edge_inclusivity():
    m = torch.ones(1, 6, 6)
    out = ops.crop_mask(m, torch.tensor([[1.0, 2.0, 4.0, 5.0]]))[0]
    assert out[:, 1:4].sum() == 9 and out.sum() == 9 and out[2, 1] == 1 and out[2, 4] == 0 and out[5, 1] == 0
    out = ops.crop_mask(m, torch.tensor([[0.5, 0.5, 3.5, 3.5]]))[0]  # r >= 0.5 -> 1.., r < 3.5 -> ..3
    assert out.sum() == 9 and out[1, 1] == 1 and out[0, 0] == 0 and out[3, 3] == 1 and out[4, 4] == 0


def test_scale_masks_slices_letterbox_pad_with_int_truncation():
    # proto 184x320 for a 736x1280 input of a 1080x1920 frame: gain 1/6, pad_h = 2 -> rows [2, 182)
    masks = torch.zeros(1, 1, 184, 320)
    masks[..., 2:182, :] = 1.0
    out = ops.scale_masks(masks, (1080, 1920))
    assert out.shape == (1, 1, 1080, 1920) and bool((out == 1).all())


def test_nms_semantics_of_torchvision_backend():
    import torchvision
    # IoU exactly == thr is kept (strict >); equal scores: lower index first and survives
    b = torch.tensor([[0.0, 0.0, 10.0, 10.0], [0.0, 0.0, 10.0, 5.0]])  # IoU = 0.5
    assert torchvision.ops.nms(b, torch.tensor([0.9, 0.8]), 0.5).tolist() == [0, 1]
    assert torchvision.ops.nms(b, torch.tensor([0.9, 0.8]), 0.49).tolist() == [0]
    assert torchvision.ops.nms(b, torch.tensor([0.7, 0.7]), 0.4).tolist() == [0]
    z = torch.tensor([[5.0, 5.0, 5.0, 5.0], [5.0, 5.0, 5.0, 5.0]])  # zero area -> NaN IoU -> never suppressed
    assert torchvision.ops.nms(z, torch.tensor([0.9, 0.8]), 0.5).tolist() == [0, 1]


def _pred(boxes_xywh, scores):
    """Build a (1, 4+nc, A) prediction tensor from xywh boxes and (A, nc) scores."""
    return torch.cat([boxes_xywh.T, scores.T], 0)[None]


def test_non_max_suppression_class_offset_and_order():
    boxes = torch.tensor([[50.0, 50.0, 20.0, 20.0], [51.0, 50.0, 20.0, 20.0], [50.0, 51.0, 20.0, 20.0], [300.0, 300.0, 10.0, 10.0]])
    scores = torch.zeros(4, 3)
    scores[0, 0], scores[1, 0], scores[2, 1], scores[3, 2] = 0.9, 0.8, 0.7, 0.1
    out, idx = ops.non_max_suppression(_pred(boxes, scores), 0.25, 0.5, nc=3, return_idx=True)
    # anchor 1 is suppressed by 0 (same class), anchor 2 survives (other class), anchor 3 below conf
    assert idx[0].tolist() == [0, 2]
    assert out[0][:, 5].tolist() == [0.0, 1.0] and torch.allclose(out[0][0, :4], torch.tensor([40.0, 40.0, 60.0, 60.0]))
    out, idx = ops.non_max_suppression(_pred(boxes, scores), 0.25, 0.5, nc=3, agnostic=True, return_idx=True)
    assert idx[0].tolist() == [0]
    out = ops.non_max_suppression(_pred(boxes, scores), 0.95, 0.5, nc=3)
    assert out[0].shape == (0, 6)


def test_non_max_suppression_max_det_truncation():
    n = 500
    g = torch.Generator().manual_seed(0)
    boxes = torch.cat([torch.arange(n)[:, None] * 30.0 + 10, torch.full((n, 1), 10.0), torch.full((n, 2), 8.0)], 1)
    scores = torch.rand(n, 1, generator=g) * 0.5 + 0.4
    out, idx = ops.non_max_suppression(_pred(boxes, scores), 0.25, 0.7, nc=1, max_det=300, return_idx=True)
    assert len(idx[0]) == 300
    assert idx[0].tolist() == scores[:, 0].argsort(descending=True, stable=True)[:300].tolist()


def test_v10_two_stage_topk_equals_flat_topk_on_tie_free_scores():
    g = torch.Generator().manual_seed(1)
    preds = torch.rand(2, 8400, 84, generator=g)
    # distinct scores (a random permutation of k/N): torch.topk's tie order is unspecified
    n = 8400 * 80
    preds[..., 4:] = torch.stack([torch.randperm(n, generator=g) for _ in range(2)]).view(2, 8400, 80).float() / n
    out = Detect.postprocess(preds, 300, 80)
    flat, fi = preds[..., 4:].flatten(1).topk(300)
    assert torch.equal(out[..., 4], flat)
    assert torch.equal(out[..., 5], (fi % 80).float())
    assert torch.equal(out[..., :4], preds[..., :4][torch.arange(2)[:, None], fi // 80])
