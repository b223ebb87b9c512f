"""Cross-check of the oracle against the real `ultralytics` package the reference imports (reference
pyproject.toml:23, yolo_seg/app.py:7,45-50) - runs wherever that package is importable, skipped otherwise
(it is neither vendored under /root/reference nor installable in the build container: no network).

Same synthetic state_dict into upstream's own SegmentationModel / DetectionModel (the key names are upstream's,
SURVEY.md A.6), same synthetic frames: raw head output, NMS rows and retina masks must agree."""
import numpy as np
import pytest
import torch

ultralytics = pytest.importorskip("ultralytics")


def _upstream_model(name):
    from ultralytics.nn.tasks import DetectionModel, SegmentationModel
    cls = SegmentationModel if name.endswith("-seg") else DetectionModel
    return cls(f"{name}.yaml", ch=3, nc=80, verbose=False).eval()


@pytest.mark.parametrize("name", ["yolov8n-seg", "yolo11n-seg", "yolov10n"])
def test_raw_head_output_matches_upstream(name):
    from oracle import ops as oops
    from oracle.model import build_model
    from yolo_puncture_b200 import synth
    net = build_model(name)
    sd = synth.synth_state_dict([(k, v.shape) for k, v in net.state_dict().items()], name)
    net.load_state_dict(sd)
    up = _upstream_model(name)
    missing, unexpected = up.load_state_dict(sd, strict=False)
    assert not [k for k in missing if "num_batches_tracked" not in k], missing
    assert not unexpected, unexpected
    im = oops.preprocess(synth.synth_frames(2), 640)
    with torch.no_grad():
        a, b = net(im), up(im)
    pa = a[0] if isinstance(a, (list, tuple)) else a
    pb = b[0] if isinstance(b, (list, tuple)) else b
    if isinstance(pb, dict):  # newer upstream end2end heads return {"one2many", "one2one"}
        pb = pb["one2one"]
    assert pa.shape == pb.shape
    assert torch.allclose(pa, pb, rtol=1e-4, atol=1e-4)
    if name.endswith("-seg"):
        proto_a, proto_b = a[1][-1], b[1][-1]
        assert torch.allclose(proto_a, proto_b, rtol=1e-4, atol=1e-4)


def test_nms_rows_and_retina_masks_match_upstream():
    from ultralytics.utils import ops as uops
    from oracle import ops as oops
    from oracle.model import build_model
    from yolo_puncture_b200 import synth
    name = "yolov8n-seg"
    net = build_model(name)
    sd = synth.synth_state_dict([(k, v.shape) for k, v in net.state_dict().items()], name)
    net.load_state_dict(sd)
    frames = synth.synth_frames(2, 480, 640)
    im = oops.preprocess(frames, 640)
    with torch.no_grad():
        pred, (maps, mc, proto) = net(im)
    mine = oops.non_max_suppression(pred, 0.25, 0.7, nc=80)
    theirs = uops.non_max_suppression(pred, 0.25, 0.7, nc=80)
    for m, t, p in zip(mine, theirs, proto):
        assert torch.equal(m, t)
        if len(m) == 0:
            continue
        d = m.clone()
        d[:, :4] = oops.scale_boxes(im.shape[2:], d[:, :4], frames[0].shape)
        t2 = t.clone()
        t2[:, :4] = uops.scale_boxes(im.shape[2:], t2[:, :4], frames[0].shape)
        assert torch.equal(d[:, :4], t2[:, :4])
        ma = oops.process_mask_native(p, d[:, 6:], d[:, :4], frames[0].shape[:2])
        mb = uops.process_mask_native(p, t2[:, 6:], t2[:, :4], frames[0].shape[:2])
        # upstream changed the pad rounding of scale_masks inside 8.3.x: accept either variant of the oracle
        if not torch.equal(ma.bool(), mb.bool()):
            ma = oops.process_mask_native(p, d[:, 6:], d[:, :4], frames[0].shape[:2], round_pad=True)
        assert torch.equal(ma.bool(), mb.bool())


def test_letterbox_matches_upstream():
    from ultralytics.data.augment import LetterBox
    from oracle import ops as oops
    rng = np.random.default_rng(0)
    for shape, auto in (((1080, 1920), True), ((480, 640), False), ((637, 640), False), ((300, 200), True)):
        img = rng.integers(0, 256, (*shape, 3), dtype=np.uint8)
        a = oops.letterbox(img, (1280, 1280) if shape[0] > 640 else (640, 640), auto=auto)
        b = LetterBox((1280, 1280) if shape[0] > 640 else (640, 640), auto=auto, stride=32)(image=img)
        assert a.shape == b.shape and np.array_equal(a, b)
