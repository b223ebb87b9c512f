"""Index-mask hand-off (reference yolo_seg/yolo_with_deva.py:54-88): oracle micro-case on the CPU,
kernel vs oracle on the GPU (random masks and real predict() output) — integer output, bit-exact."""
import pytest
import torch

from oracle.handoff import auto_segment_index_mask, coord_min_rect_len


def test_oracle_overwrite_order_and_small_mask_suppression():
    m = torch.zeros(3, 20, 20)
    m[0, 0:15, 0:15] = 1       # 225 px  -> id 1
    m[1, 2:5, 2:5] = 1         # 9 px    -> suppressed (< 100), id not consumed
    m[2, 5:18, 5:18] = 1       # 169 px  -> id 2, overwrites the overlap with detection 0
    out, info = auto_segment_index_mask(m, torch.tensor([0.9, 0.8, 0.7]), torch.tensor([3.0, 4.0, 5.0]), True, 100)
    assert [i[0] for i in info] == [1, 2] and [i[2] for i in info] == [3, 5]
    assert out[0, 0] == 1 and out[3, 3] == 1 and out[10, 10] == 2 and out[19, 19] == 0
    out2, info2 = auto_segment_index_mask(m, torch.tensor([0.9, 0.8, 0.7]), torch.tensor([3.0, 4.0, 5.0]), False, 100)
    assert [i[0] for i in info2] == [1, 2, 3] and out2[3, 3] == 2 and out2[10, 10] == 3


class _FakeBoxes:
    def __init__(self, conf, cls):
        self.conf, self.cls = conf, cls
        self.data = torch.cat([torch.zeros(len(conf), 4, device=conf.device), conf[:, None], cls[:, None]], 1)


class _FakeMasks:
    def __init__(self, raw):
        self.raw = raw


class _FakeResults:
    def __init__(self, raw, conf, cls, shape):
        self.masks = _FakeMasks(raw) if raw is not None else None
        self.boxes = _FakeBoxes(conf, cls)
        self.orig_shape = shape


@pytest.mark.gpu
@pytest.mark.parametrize("H,W", [(640, 640), (1080, 1920), (37, 53)])
@pytest.mark.parametrize("suppress", [True, False])
def test_index_masks_kernel_matches_oracle_on_random_masks(H, W, suppress):
    from yolo_puncture_b200 import index_masks
    g = torch.Generator().manual_seed(H * 7 + W)
    counts = [5, 0, 1, 17]
    res, refs = [], []
    for n in counts:
        if n == 0:
            res.append(_FakeResults(None, torch.zeros(0), torch.zeros(0), (H, W)))
            refs.append((None, []))
            continue
        m = torch.zeros(n, H, W, dtype=torch.uint8)
        for i in range(n):  # rectangles of very different areas, some tiny
            y0, x0 = int(torch.randint(0, H - 2, (1,), generator=g)), int(torch.randint(0, W - 2, (1,), generator=g))
            hh, ww = int(torch.randint(1, max(2, H // 2), (1,), generator=g)), int(torch.randint(1, max(2, W // 2), (1,), generator=g))
            if i % 3 == 0:
                hh, ww = min(hh, 6), min(ww, 9)
            m[i, y0:y0 + hh, x0:x0 + ww] = 1
        conf, cls = torch.rand(n, generator=g), torch.randint(0, 80, (n,), generator=g).float()
        res.append(_FakeResults(m.cuda(), conf.cuda(), cls.cuda(), (H, W)))
        refs.append(auto_segment_index_mask(m.float(), conf, cls, suppress, 100))
    out = index_masks(res, suppress_small_mask=suppress, min_area=100)
    for (imap, info), (rmap, rinfo) in zip(out, refs):
        if rmap is None:
            assert int(imap.abs().sum()) == 0 and info == []
            continue
        assert torch.equal(imap.cpu(), rmap)
        assert [(d["id"], d["category_id"]) for d in info] == [(i[0], i[2]) for i in rinfo]
        assert all(abs(d["score"] - i[1]) < 1e-6 for d, i in zip(info, rinfo))


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,counts", [(640, 640, [40, 0, 300, 3]), (1080, 1920, [25, 1]), (100, 104, [7, 2])])
def test_index_masks_boxed_path_matches_oracle(H, W, counts):
    """Masks cropped to their (fractional) boxes, as predict() produces them: the rectangle-restricted area / paint kernels
    (ypb_index_masks_boxed) against the reference loop.  300 detections in a frame = two rounds of the paint's list."""
    from yolo_puncture_b200 import index_masks
    g = torch.Generator().manual_seed(H + 3 * W)
    res, refs = [], []
    for n in counts:
        if n == 0:
            res.append(_FakeResults(None, torch.zeros(0), torch.zeros(0), (H, W)))
            refs.append((None, []))
            continue
        m = torch.zeros(n, H, W, dtype=torch.uint8)
        boxes = torch.zeros(n, 4)
        for i in range(n):
            bw, bh = float(torch.rand(1, generator=g)) * W * (0.05 if i % 3 == 0 else 0.6) + 2, float(torch.rand(1, generator=g)) * H * (0.05 if i % 3 == 0 else 0.6) + 2
            x1, y1 = float(torch.rand(1, generator=g)) * (W - bw), float(torch.rand(1, generator=g)) * (H - bh)
            boxes[i] = torch.tensor([x1, y1, x1 + bw, y1 + bh])
            ys, xs = torch.arange(H).float()[:, None], torch.arange(W).float()[None, :]
            inside = (xs >= x1) & (xs < x1 + bw) & (ys >= y1) & (ys < y1 + bh)
            pattern = torch.rand(H, W, generator=g) < 0.7
            m[i] = (inside & pattern).to(torch.uint8)
        conf, cls = torch.rand(n, generator=g), torch.randint(0, 80, (n,), generator=g).float()
        r = _FakeResults(m.cuda(), conf.cuda(), cls.cuda(), (H, W))
        r.boxes.data = torch.cat([boxes, conf[:, None], cls[:, None]], 1).cuda()
        r.masks.cropped = True
        res.append(r)
        refs.append(auto_segment_index_mask(m.float(), conf, cls, True, 100))
    out = index_masks(res, suppress_small_mask=True, min_area=100)
    for (imap, info), (rmap, rinfo) in zip(out, refs):
        if rmap is None:
            assert int(imap.abs().sum()) == 0 and info == []
            continue
        assert torch.equal(imap.cpu(), rmap)
        assert [(d["id"], d["category_id"]) for d in info] == [(i[0], i[2]) for i in rinfo]
    # the full-scan path on the same masks gives the same answer
    for r in res:
        if r.masks is not None:
            r.masks.cropped = False
    out2 = index_masks(res, suppress_small_mask=True, min_area=100)
    for (a, ia), (b, ib) in zip(out, out2):
        assert torch.equal(a, b) and ia == ib


@pytest.mark.gpu
def test_index_masks_on_predict_output():
    from yolo_puncture_b200 import YOLO, index_masks, synth
    yolo = YOLO("yolov8n-seg", device=0)
    frames = synth.synth_frames(6)
    res = yolo.predict(frames, conf=0.25, retina_masks=True)
    out = index_masks(res, suppress_small_mask=True, min_area=100)
    assert len(out) == len(frames)
    for r, (imap, info) in zip(res, out):
        assert imap.shape == (640, 640) and imap.dtype == torch.int64
        if r.masks is None:
            assert info == [] and int(imap.max()) == 0
            continue
        rmap, rinfo = auto_segment_index_mask(r.masks.data.cpu(), r.boxes.conf.cpu(), r.boxes.cls.cpu(), True, 100)
        assert torch.equal(imap.cpu(), rmap)
        assert [(d["id"], d["category_id"]) for d in info] == [(i[0], i[2]) for i in rinfo]


def _shape_masks(H, W):
    """Rotated rectangles, an ellipse, two separate blobs, an L shape: masks whose minimum-area rectangle is unambiguous."""
    import cv2
    import numpy as np
    ms = []
    for (cx, cy, w, h, ang) in [(300, 200, 220, 40, 17.0), (150, 400, 90, 60, -33.0), (400, 420, 300, 12, 71.0), (320, 320, 50, 50, 45.0)]:
        m = np.zeros((H, W), np.uint8)
        box = cv2.boxPoints(((cx, cy), (w, h), ang)).astype(np.int32)
        cv2.fillPoly(m, [box], 1)
        ms.append(m)
    m = np.zeros((H, W), np.uint8)
    cv2.ellipse(m, (250, 300), (120, 35), 28.0, 0, 360, 1, -1)
    ms.append(m)
    m = np.zeros((H, W), np.uint8)
    cv2.circle(m, (100, 100), 30, 1, -1)
    cv2.circle(m, (400, 180), 22, 1, -1)
    ms.append(m)
    m = np.zeros((H, W), np.uint8)
    m[100:400, 100:140] = 1
    m[360:400, 100:300] = 1
    ms.append(m)
    return np.stack(ms)


def test_oracle_min_rect_len_of_an_axis_aligned_bar():
    import numpy as np
    m = np.zeros((100, 100), np.uint8)
    m[10:20, 5:85] = 1  # 80 x 10 pixels -> pixel-centre extents 79 x 9
    length, ratio = coord_min_rect_len(m)
    assert abs(length - 79.0) < 1e-3 and abs(ratio - 79.0 / 9.0) < 1e-3


@pytest.mark.gpu
def test_min_rect_len_matches_cv2_on_shapes():
    from yolo_puncture_b200 import min_rect_len
    ms = _shape_masks(480, 640)
    got = min_rect_len(torch.from_numpy(ms).cuda()).cpu()
    for i, m in enumerate(ms):
        length, ratio = coord_min_rect_len(m)
        assert abs(float(got[i, 0]) - length) <= 2e-2 + 1e-4 * length, (i, float(got[i, 0]), length)
        assert abs(float(got[i, 1]) - ratio) <= 1e-2 * ratio, (i, float(got[i, 1]), ratio)


@pytest.mark.gpu
def test_min_rect_len_on_predict_output():
    from yolo_puncture_b200 import YOLO, min_rect_len, synth
    yolo = YOLO("yolov8n-seg", device=0)
    res = yolo.predict(synth.synth_frames(4), conf=0.25, retina_masks=True)
    checked = 0
    for r in res:
        if r.masks is None:
            continue
        got = min_rect_len(r.masks).cpu()
        m = r.masks.raw.cpu().numpy()
        for i in range(len(m)):
            if int(m[i].sum()) < 50:
                continue
            length, ratio = coord_min_rect_len(m[i])
            # the minimum-area rectangle can be ambiguous (two orientations of nearly equal area): compare areas,
            # and the length when the areas pin the same rectangle
            assert abs(float(got[i, 0]) - length) <= 0.05 + 2e-3 * length, (i, float(got[i, 0]), length)
            checked += 1
    assert checked > 0


def test_oracle_min_side_branch_resizes_like_torchvision():
    """CPU: the restated `min_side` branch on a hand-checkable case - a 2x up-scaled frame, masks brought back by a 2x
    antialiased bilinear down-scale (weights 1/8, 3/8, 3/8, 1/8): a 20x20 block of ones at even offsets comes back as a
    10x10 block (interior 1, the border row at 0.5+ stays on only where the kernel's weight > 0.5)."""
    from oracle.handoff import auto_segment_index_mask_resized
    m = torch.zeros(1, 40, 40)
    m[0, 10:30, 10:30] = 1
    out, info = auto_segment_index_mask_resized(m, torch.tensor([0.9]), torch.tensor([2.0]), (20, 20), True, 50)
    assert len(info) == 1 and info[0][0] == 1
    assert out[5:15, 5:15].eq(1).all() and int(out.sum()) == 100


def _aa_resize_reference(masks_u8, H, W):
    from torchvision.transforms import functional as F
    return torch.stack([F.resize(m.float().unsqueeze(0), size=[H, W])[0] for m in masks_u8])


@pytest.mark.gpu
@pytest.mark.parametrize("h1,w1,H,W", [(960, 1280, 480, 640), (720, 1280, 1080, 1920), (333, 500, 480, 721), (1000, 750, 400, 300),
                                         (640, 640, 640, 640)])
def test_index_masks_resized_matches_torchvision_resize(h1, w1, H, W):
    """GPU: `index_masks(out_shape=...)` (ypb_index_masks_resized) against the reference loop with torchvision's F.resize
    (antialiased bilinear), down- and up-scaling, on blob masks: painted ids equal except on pixels whose resized value is
    within 1e-4 of the 0.5 threshold; kept / suppressed sets equal."""
    from oracle.handoff import auto_segment_index_mask_resized
    from yolo_puncture_b200 import index_masks
    g = torch.Generator().manual_seed(h1 + W)
    n = 6
    m = torch.zeros(n, h1, w1, dtype=torch.uint8)
    yy, xx = torch.meshgrid(torch.arange(h1), torch.arange(w1), indexing="ij")
    for i in range(n):
        cy, cx = float(torch.rand(1, generator=g)) * h1, float(torch.rand(1, generator=g)) * w1
        r = (3.0 if i == 2 else 20.0 + 100.0 * float(torch.rand(1, generator=g)))  # one tiny blob: suppressed
        m[i] = (((yy - cy) ** 2 + ((xx - cx) * 0.7) ** 2) < r * r).to(torch.uint8)
    conf, cls = torch.rand(n, generator=g), torch.randint(0, 80, (n,), generator=g).float()
    res = [_FakeResults(m.cuda(), conf.cuda(), cls.cuda(), (h1, w1))]
    (got, info), = index_masks(res, True, 100, out_shape=(H, W))
    ref, ref_info = auto_segment_index_mask_resized(m, conf, cls, (H, W), True, 100)
    assert [d["id"] for d in info] == [i[0] for i in ref_info]
    assert [d["category_id"] for d in info] == [i[2] for i in ref_info]
    got = got.cpu()
    assert got.shape == (H, W) and got.dtype == torch.int64
    diff = got != ref
    if (h1, w1) == (H, W):
        assert not diff.any()
    else:
        r = _aa_resize_reference(m, H, W)
        near = ((r - 0.5).abs() < 1e-4).any(0)          # the value sits on the threshold: either side is right
        assert not (diff & ~near).any(), int((diff & ~near).sum())
        assert int(diff.sum()) <= 0.001 * H * W


@pytest.mark.gpu
def test_auto_segment_drop_in_with_min_side():
    """The reference's entry point itself (yolo_with_deva.py:35-88), min_side > 0: frame resized, predict, masks back."""
    from oracle.handoff import auto_segment_index_mask_resized
    from yolo_puncture_b200 import YOLO, auto_segment, synth
    yolo = YOLO("yolov8n-seg", device=0)
    frame = synth.synth_frame(0, 480, 640)
    out, info = auto_segment({"MIN_AREA_THRESHOLD": 100}, frame, yolo, 720, True)
    assert out.shape == (480, 640) and out.dtype == torch.int64 and out.is_cuda
    import cv2
    big = cv2.resize(frame, (int(640 * 1.5), int(480 * 1.5)))
    r = yolo.predict(big, retina_masks=True, conf=0.9)[0]
    if r.masks is None:
        assert info == [] and int(out.sum()) == 0
        return
    ref, ref_info = auto_segment_index_mask_resized(r.masks.data.cpu(), r.boxes.conf.cpu(), r.boxes.cls.cpu(), (480, 640), True, 100)
    assert [d["id"] for d in info] == [i[0] for i in ref_info]
    assert int((out.cpu() != ref).sum()) <= 0.001 * 480 * 640


def test_pack_infos_matches_the_per_frame_loop():
    """Host side of the hand-off: kept detections only, in order, with their row inside the frame (pure numpy, no GPU)."""
    import numpy as np
    from yolo_puncture_b200.handoff import _pack_infos
    rng = np.random.default_rng(3)
    counts = [3, 0, 5, 1, 0, 4]
    n = sum(counts)
    ids = np.zeros(n, np.int32)
    k = 0
    for c in counts:
        cur = 0
        for j in range(c):
            if rng.random() < 0.6:
                cur += 1
                ids[k + j] = cur
        k += c
    allb = rng.random((n, 6)).astype(np.float32)
    allb[:, 5] = rng.integers(0, 80, n)
    maps = torch.zeros((len(counts), 2, 2), dtype=torch.int64)
    out = _pack_infos(maps, ids, allb, counts)
    k = 0
    for b, c in enumerate(counts):
        exp = [{"id": int(ids[k + r]), "score": float(allb[k + r, 4]), "category_id": int(allb[k + r, 5]), "index": r}
               for r in range(c) if ids[k + r]]
        assert out[b][1] == exp and out[b][0] is not None
        k += c
    assert all(info == [] for _, info in _pack_infos(maps, np.zeros(n, np.int32), allb, counts))
