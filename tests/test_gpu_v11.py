"""YOLO11-seg (the app's default weights, reference yolo_seg/app.py:216-223 and
yolo_with_deva.py:226) on the GPU: C3k2 / C3k / C2PSA blocks and the depthwise class branch.
Layers vs the bf16-emulating oracle; selection and masks strict on the engine's own head / proto tensors."""
import pytest
import torch

from gpu_util import mask_iou, oracle_select_on_engine_tensors, oracle_with_synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["yolo11n-seg", "yolo11s-seg"])
def v11(request):
    from oracle import ops as oops
    from yolo_puncture_b200 import YOLO, synth
    net, sd = oracle_with_synth(request.param, emulate=True)
    yolo = YOLO(request.param, state_dict=sd, device=0)
    return {"name": request.param, "net": net, "yolo": yolo, "frames": synth.synth_frames(3), "oops": oops}


def test_v11_layers_track_bf16_emulating_oracle(v11):
    net, yolo, frames, oops = v11["net"], v11["yolo"], v11["frames"], v11["oops"]
    yolo.predict(frames, conf=0.25, retina_masks=True)
    eng = yolo.engine
    assert eng.device_error() == 0
    with torch.no_grad():
        feats = net.features(oops.preprocess(frames, 640))
    for vname in eng.view_table():
        if not vname.startswith("model."):
            continue
        ref = feats[int(vname.split(".")[1])]
        if not torch.is_tensor(ref):
            continue
        got, ref = eng.view(vname).float().cpu(), ref.permute(0, 2, 3, 1)
        rel = float((got - ref).abs().mean() / ref.abs().mean())
        assert rel < 0.03, f"{v11['name']} {vname}: mean relative error {rel:.4f}"
        assert not torch.isnan(got).any()
    pred, (maps, mc, proto) = feats[-1]
    rel = float((eng.view("proto").float().cpu() - proto.permute(0, 2, 3, 1)).abs().mean() / proto.abs().mean())
    assert rel < 0.04
    raw = torch.cat([torch.cat([m.flatten(2) for m in maps], 2), mc], 1).permute(0, 2, 1)
    got = eng.view("head")[:, 0].float().cpu()
    assert float((got - raw).abs().mean() / raw.abs().mean()) < 0.04


@pytest.mark.parametrize("conf,iou,max_det", [(0.25, 0.7, 300), (0.5, 0.45, 300), (0.25, 0.7, 5)])
def test_v11_selection_and_masks_strict_on_engine_tensors(v11, conf, iou, max_det):
    net, yolo, frames, oops = v11["net"], v11["yolo"], v11["frames"], v11["oops"]
    res = yolo.predict(frames, conf=conf, iou=iou, max_det=max_det, retina_masks=True)
    eng, B = yolo.engine, len(frames)
    dets, kept, proto = oracle_select_on_engine_tensors(eng, net, B, [(80, 80), (40, 40), (20, 20)], conf, iou, max_det)
    for b in range(B):
        n = len(res[b])
        assert n == len(dets[b])
        assert eng.keep[b, :n].cpu().long().tolist() == kept[b].tolist()  # bit-exact kept anchors
        if n == 0:
            assert res[b].masks is None
            continue
        d = dets[b].clone()
        d[:, :4] = oops.scale_boxes((640, 640), d[:, :4], (640, 640))
        got = res[b].boxes.data.cpu()
        assert torch.equal(got[:, 5], d[:, 5])
        assert (got[:, :4] - d[:, :4]).abs().max() <= 1e-2
        assert (got[:, 4] - d[:, 4]).abs().max() <= 1e-6
        mo = oops.process_mask_native(proto[b], d[:, 6:], d[:, :4], (640, 640))
        assert mask_iou(mo, res[b].masks.data.cpu()).min() >= 0.99
