"""Selection parity (through the C ABI): the batched NMS kernel keeps exactly the indices torchvision.ops.nms
keeps when run on the oracle's boxes/scores — bit-exact, including ties, class offsets, zero-area boxes, the
max_det cap, the >4096-candidate global-memory sort path and the max_nms=30000 truncation."""
import os

import numpy as np
import pytest
import torch
import torchvision

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def ref_keep(boxes, scores, cls, iou, max_det, agnostic, max_nms=30000):
    idx = torch.arange(len(scores))
    if len(scores) > max_nms:
        idx = scores.argsort(descending=True, stable=True)[:max_nms]
        boxes, scores, cls = boxes[idx], scores[idx], cls[idx]
    off = cls.float()[:, None] * (0.0 if agnostic else 7680.0)
    k = torchvision.ops.nms(boxes + off, scores, iou)[:max_det]
    return idx[k]


def run(boxes, scores, cls, n_valid, iou=0.7, max_det=300, agnostic=False):
    from yolo_puncture_b200.engine import nms
    keep, count = nms(boxes.cuda(), scores.cuda(), cls.cuda(), n_valid.cuda(), iou, max_det, agnostic)
    torch.cuda.synchronize()
    return keep.cpu(), count.cpu()


def random_case(B, N, seed, nclass=80, tie_levels=None, spread=600.0):
    g = torch.Generator().manual_seed(seed)
    ctr = torch.rand(B, N, 2, generator=g) * spread + 20
    wh = torch.rand(B, N, 2, generator=g) * 150 + 10
    boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], -1)
    scores = torch.rand(B, N, generator=g) * 0.7 + 0.25
    if tie_levels:
        scores = (scores * tie_levels).round() / tie_levels
    cls = torch.randint(0, nclass, (B, N), generator=g, dtype=torch.int32)
    return boxes, scores, cls


@pytest.mark.parametrize("N,nclass,ties", [(1, 80, None), (37, 3, 64), (300, 80, None), (1000, 2, 512), (5000, 4, None),
                                           (8400, 1, 4096), (2500, 300, 256), (3000, 16, None)])
@pytest.mark.parametrize("agnostic", [False, True])
def test_nms_bit_exact_vs_torchvision(N, nclass, ties, agnostic):
    B = 3
    boxes, scores, cls = random_case(B, N, seed=N, nclass=nclass, tie_levels=ties)
    n_valid = torch.tensor([N, max(N // 2, 1), max(N - 1, 1)], dtype=torch.int32)
    keep, count = run(boxes, scores, cls, n_valid, agnostic=agnostic)
    for b in range(B):
        n = int(n_valid[b])
        ref = ref_keep(boxes[b, :n], scores[b, :n], cls[b, :n], 0.7, 300, agnostic)
        assert int(count[b]) == len(ref)
        assert keep[b, :len(ref)].tolist() == ref.tolist()


def test_nms_class_buckets_and_window_fallback():
    """The class-aware scan chains kept boxes per (class & 127) bucket and only runs while every coordinate sits in a
    window narrower than the class offset; dense same-class clusters, bucket collisions (classes 5 / 133 / 261) and
    coordinates outside the window (plain scan) must all keep exactly what torchvision keeps."""
    g = torch.Generator().manual_seed(11)
    N = 2400
    ctr = torch.rand(1, N, 2, generator=g) * 120 + 200           # one dense cluster
    wh = torch.rand(1, N, 2, generator=g) * 60 + 40
    boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], -1)
    scores = torch.rand(1, N, generator=g) * 0.7 + 0.25
    cls = torch.tensor([5, 133, 261], dtype=torch.int32)[torch.randint(0, 3, (1, N), generator=g)]
    nv = torch.tensor([N], dtype=torch.int32)
    for iou in (0.7, 0.3):
        keep, count = run(boxes, scores, cls, nv, iou=iou)
        ref = ref_keep(boxes[0], scores[0], cls[0], iou, 300, False)
        assert int(count[0]) == len(ref) and keep[0, :len(ref)].tolist() == ref.tolist()
    far = boxes.clone()
    far[0, ::7] += 6000.0                                         # outside the window: the plain scan takes over
    keep, count = run(far, scores, cls, nv)
    ref = ref_keep(far[0], scores[0], cls[0], 0.7, 300, False)
    assert int(count[0]) == len(ref) and keep[0, :len(ref)].tolist() == ref.tolist()


def test_nms_golden_fixture():
    gold = np.load(os.path.join(G, "nms_case.npz"))
    boxes, scores, cls = (torch.from_numpy(gold[k]) for k in ("boxes", "scores", "cls"))
    B, N = scores.shape
    keep, count = run(boxes, scores, cls, torch.full((B,), N, dtype=torch.int32))
    for b in range(B):
        ref = [int(x) for x in gold["keep"][b] if x >= 0]
        assert int(count[b]) == len(ref) and keep[b, :len(ref)].tolist() == ref


def test_nms_edge_cases():
    # empty image, zero-area duplicates (NaN IoU never suppresses), IoU exactly at the threshold is kept
    boxes = torch.tensor([[[0.0, 0.0, 10.0, 10.0], [0.0, 0.0, 10.0, 5.0], [5.0, 5.0, 5.0, 5.0], [5.0, 5.0, 5.0, 5.0]]]).repeat(2, 1, 1)
    scores = torch.tensor([[0.9, 0.8, 0.7, 0.6]]).repeat(2, 1)
    cls = torch.zeros((2, 4), dtype=torch.int32)
    keep, count = run(boxes, scores, cls, torch.tensor([4, 0], dtype=torch.int32), iou=0.5)
    assert count.tolist() == [4, 0] and keep[0, :4].tolist() == [0, 1, 2, 3]
    keep, count = run(boxes, scores, cls, torch.tensor([4, 4], dtype=torch.int32), iou=0.49)
    assert count.tolist() == [3, 3] and keep[0, :3].tolist() == [0, 2, 3]


def test_nms_max_det_and_max_nms_truncation():
    N = 33600  # anchors of a 1280x1280 input: more candidates than max_nms = 30000
    g = torch.Generator().manual_seed(3)
    boxes, _, cls = random_case(1, N, seed=5, nclass=80, spread=1200.0)
    scores = (torch.randperm(N, generator=g).float() / N * 0.7 + 0.25)[None]  # distinct scores
    keep, count = run(boxes, scores, cls, torch.tensor([N], dtype=torch.int32), max_det=300)
    ref = ref_keep(boxes[0], scores[0], cls[0], 0.7, 300, False)
    assert int(count[0]) == len(ref) == 300 and keep[0].tolist() == ref.tolist()
    keep, count = run(boxes[:, :2000], scores[:, :2000], cls[:, :2000], torch.tensor([2000], dtype=torch.int32), max_det=7)
    assert int(count[0]) == 7 and keep[0, :7].tolist() == ref_keep(boxes[0, :2000], scores[0, :2000], cls[0, :2000], 0.7, 7, False).tolist()
