"""YOLOv10n (NMS-free, one-to-one head + top-k) parity on the GPU — BASELINE config C2's model.
Layers vs the bf16-emulating oracle; selection strict: the oracle's Detect.postprocess + conf filter run on
the ENGINE's head tensor must give the same (anchor, class) picks in the same order."""
import pytest
import torch

from gpu_util import oracle_with_synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def v10():
    from oracle import ops as oops
    from yolo_puncture_b200 import YOLO, synth
    net, sd = oracle_with_synth("yolov10n", emulate=True)
    yolo = YOLO("yolov10n", state_dict=sd, device=0)
    return {"net": net, "yolo": yolo, "frames": synth.synth_frames(3), "oops": oops}


def test_v10_layers_track_bf16_emulating_oracle(v10):
    net, yolo, frames, oops = v10["net"], v10["yolo"], v10["frames"], v10["oops"]
    res = yolo.predict(frames, conf=0.25)
    assert all(r.masks is None for r in res)
    eng = yolo.engine
    assert eng.device_error() == 0
    with torch.no_grad():
        feats = net.features(oops.preprocess(frames, 640), upto=len(net.model) - 1)
        head = net.model[-1]
        maps = head.head_maps([feats[j] for j in net.froms[-1]], one2one=True)
    for vname in eng.view_table():
        if not vname.startswith("model."):
            continue
        ref = feats[int(vname.split(".")[1])]
        got, ref = eng.view(vname).float().cpu(), ref.permute(0, 2, 3, 1)
        rel = float((got - ref).abs().mean() / ref.abs().mean())
        assert rel < 0.03, f"{vname}: mean relative error {rel:.4f}"
    raw = torch.cat([m.flatten(2) for m in maps], 2).permute(0, 2, 1)
    got = eng.view("head")[:, 0].float().cpu()
    assert float((got - raw).abs().mean() / raw.abs().mean()) < 0.03


@pytest.mark.parametrize("conf,max_det,classes", [(0.25, 300, None), (0.5, 300, None), (0.25, 10, None), (0.25, 300, [0, 7, 22, 44]), (0.25, 10, [0, 7, 22, 44]), (0.15, 300, [3, 5]),
                                                  (0.15, 300, None), (0.15, 50, None)])  # low conf: > 4096 pairs -> select-then-sort path
def test_v10_topk_selection_strict_on_engine_tensors(v10, conf, max_det, classes):
    net, yolo, frames, oops = v10["net"], v10["yolo"], v10["frames"], v10["oops"]
    res = yolo.predict(frames, conf=conf, max_det=max_det, classes=classes)
    eng, B, nc = yolo.engine, len(frames), 80
    hd = eng.view("head")[:, 0].float().cpu()
    maps, off = [], 0
    for (h, w) in [(80, 80), (40, 40), (20, 20)]:
        maps.append(hd[:, off:off + h * w].permute(0, 2, 1).reshape(B, 64 + nc, h, w))
        off += h * w
    head = net.model[-1]
    with torch.no_grad():
        y = head._inference(maps)                                   # (B, 4+nc, A) xyxy + sigmoid scores
        sel = head.postprocess(y.permute(0, 2, 1), head.max_det, nc)       # (B, 300, 6) descending score
    # upstream order: top-k -> conf -> [:max_det] -> classes (oracle/ops.py, end2end branch); nothing is bent here
    dets = oops.non_max_suppression(sel, conf, 0.7, classes=classes, max_det=max_det, end2end=True)
    for b in range(B):
        d = dets[b]
        n = len(res[b])
        assert n == len(d), f"frame {b}: {n} vs oracle {len(d)}"
        if n == 0:
            continue
        got = res[b].boxes.data.cpu()
        assert torch.equal(got[:, 5], d[:, 5])                      # class ids, in order: bit-exact
        assert (got[:, 4] - d[:, 4]).abs().max() <= 1e-6
        db = oops.scale_boxes((640, 640), d[:, :4].clone(), (640, 640))
        assert (got[:, :4] - db).abs().max() <= 1e-2
