"""Whole-path parity on the GPU, through YOLO.predict() / the C ABI, against the CPU oracle.

Three levels (SURVEY.md §7 "bf16 vs the 1e-2 px box bar"):
  (i)  layer level: every named activation tracks the bf16-EMULATING oracle (same storage roundings);
  (ii) stage level, strict: decode/NMS/mask kernels vs the oracle's post-processing fed the ENGINE's own
       head/proto tensors -> kept anchors and class ids bit-exact, boxes <= 1e-2 px, masks IoU >= 0.99;
  (iii) end to end vs the fp32 oracle on matched detections, drift reported and bounded.
"""
import json
import os

import numpy as np
import pytest
import torch

from gpu_util import (box_iou_matrix, drift_stats, mask_iou, oracle_select_on_engine_tensors, oracle_with_synth,
                      ulp_report)

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")


def report(**kw):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(json.dumps(kw) + "\n")


@pytest.fixture(scope="module")
def nseg():
    from oracle import ops as oops
    from yolo_puncture_b200 import YOLO, synth
    net, sd = oracle_with_synth("yolov8n-seg", emulate=True)
    yolo = YOLO("yolov8n-seg", state_dict=sd, device=0)
    frames = synth.synth_frames(3)
    return {"net": net, "sd": sd, "yolo": yolo, "frames": frames, "oops": oops}


def test_layers_track_bf16_emulating_oracle(nseg):
    net, yolo, frames, oops = nseg["net"], nseg["yolo"], nseg["frames"], nseg["oops"]
    yolo.predict(frames, conf=0.25, retina_masks=True)
    eng = yolo.engine
    assert eng.device_error() == 0
    with torch.no_grad():
        feats = net.features(oops.preprocess(frames, 640))
    worst = 0.0
    for vname in eng.view_table():
        if not vname.startswith("model."):
            continue
        ref = feats[int(vname.split(".")[1])]
        if not torch.is_tensor(ref):
            continue
        got, ref = eng.view(vname).float().cpu(), ref.permute(0, 2, 3, 1)
        rel = float((got - ref).abs().mean() / ref.abs().mean())
        worst = max(worst, rel)
        assert rel < 0.02, f"{vname}: mean relative error {rel:.4f}"
        assert not torch.isnan(got).any()
    # the stem and first conv have no accumulated drift: elementwise within one bf16 ulp
    got, ref = eng.view("model.1").float().cpu(), feats[1].permute(0, 2, 3, 1)
    assert ((got - ref).abs() <= 2.0 ** -7 * ref.abs() + 1e-2).all()
    pred, (maps, mc, proto) = feats[-1]
    rel = float((eng.view("proto").float().cpu() - proto.permute(0, 2, 3, 1)).abs().mean() / proto.abs().mean())
    assert rel < 0.03
    # the fp32 head rows [64 box logits | nc class logits | 32 mask coefficients] that decode / NMS / masks consume
    raw = torch.cat([torch.cat([m.flatten(2) for m in maps], 2), mc], 1).permute(0, 2, 1)
    head = eng.view("head")[:, 0].float().cpu()
    assert head.shape == raw.shape
    for name, lo, hi in (("box", 0, 64), ("cls", 64, 144), ("coef", 144, 176)):
        r = float((head[..., lo:hi] - raw[..., lo:hi]).abs().mean() / raw[..., lo:hi].abs().mean())
        assert r < 0.03, f"head rows [{name}]: mean relative error {r:.4f}"
    report(test="layers", model="yolov8n-seg", worst_mean_rel_err=worst, proto_mean_rel_err=rel)


@pytest.mark.parametrize("conf,iou,max_det,classes,agnostic", [(0.25, 0.7, 300, None, False), (0.5, 0.45, 300, None, False),
                                                              (0.25, 0.7, 5, None, False), (0.25, 0.7, 300, [7, 22, 44], False),
                                                              (0.25, 0.7, 300, None, True), (0.9, 0.7, 300, None, False)])
def test_selection_and_masks_strict_on_engine_tensors(nseg, conf, iou, max_det, classes, agnostic):
    net, yolo, frames, oops = nseg["net"], nseg["yolo"], nseg["frames"], nseg["oops"]
    res = yolo.predict(frames, conf=conf, iou=iou, max_det=max_det, classes=classes, agnostic_nms=agnostic, retina_masks=True)
    eng, B = yolo.engine, len(frames)
    dets, kept, proto = oracle_select_on_engine_tensors(eng, net, B, [(80, 80), (40, 40), (20, 20)], conf, iou, max_det,
                                                        classes, agnostic)
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
        assert torch.equal(got[:, 5], d[:, 5])                         # class ids bit-exact
        assert (got[:, :4] - d[:, :4]).abs().max() <= 1e-2              # boxes within 1e-2 px
        assert (got[:, 4] - d[:, 4]).abs().max() <= 1e-6
        mo = oops.process_mask_native(proto[b], d[:, 6:], d[:, :4], (640, 640))
        iou_m = mask_iou(mo, res[b].masks.data.cpu())
        assert iou_m.min() >= 0.99
        report(test="strict", conf=conf, frame=b, n=n, mask_mismatch_px=int((mo != res[b].masks.data.cpu()).sum()))


def test_non_retina_masks_strict(nseg):
    net, yolo, frames, oops = nseg["net"], nseg["yolo"], nseg["frames"], nseg["oops"]
    res = yolo.predict(frames, conf=0.25, iou=0.7, retina_masks=False)
    dets, kept, proto = oracle_select_on_engine_tensors(yolo.engine, net, 3, [(80, 80), (40, 40), (20, 20)], 0.25, 0.7)
    for b in range(3):
        if len(dets[b]) == 0:
            assert res[b].masks is None
            continue
        mo = oops.process_mask(proto[b], dets[b][:, 6:], dets[b][:, :4], (640, 640), upsample=True)
        me = res[b].masks.data.cpu()
        assert me.shape == mo.shape and mask_iou(mo, me).min() >= 0.99


def test_rect_1080p_retina_strict():
    """BASELINE config C4 geometry: 1920x1080 frames, imgsz 1280 -> net input 736x1280, masks at 1080p."""
    from oracle import ops as oops
    from yolo_puncture_b200 import YOLO, synth
    net, sd = oracle_with_synth("yolov8n-seg", emulate=True)
    yolo = YOLO("yolov8n-seg", state_dict=sd, device=0)
    frames = synth.synth_frames(2, 1080, 1920, start=40)
    res = yolo.predict(frames, conf=0.25, iou=0.7, retina_masks=True, imgsz=1280)
    eng = yolo.engine
    assert eng.shape == (2, 736, 1280) and eng.anchors == 19320
    dets, kept, proto = oracle_select_on_engine_tensors(eng, net, 2, [(92, 160), (46, 80), (23, 40)], 0.25, 0.7)
    for b in range(2):
        n = len(res[b])
        assert n == len(dets[b]) and eng.keep[b, :n].cpu().long().tolist() == kept[b].tolist()
        if n == 0:
            continue
        d = dets[b].clone()
        d[:, :4] = oops.scale_boxes((736, 1280), d[:, :4], (1080, 1920))
        assert (res[b].boxes.data.cpu()[:, :4] - d[:, :4]).abs().max() <= 1e-2
        mo = oops.process_mask_native(proto[b], d[:, 6:], d[:, :4], (1080, 1920))
        me = res[b].masks.data.cpu()
        assert me.shape == (n, 1080, 1920) and mask_iou(mo, me).min() >= 0.99
        report(test="rect1080p", frame=b, n=n, mask_mismatch_px=int((mo != me).sum()))
    # layer check on the rect input as well (letterbox resize happens on the host exactly like the oracle's)
    with torch.no_grad():
        feats = net.features(oops.preprocess(frames, 1280))
    got, ref = eng.view("model.15").float().cpu(), feats[15].permute(0, 2, 3, 1)
    assert float((got - ref).abs().mean() / ref.abs().mean()) < 0.02


def _drift(ref, got):
    """Match detections of `got` to `ref` by class and IoU > 0.9; return (match rate, box errors, mask IoUs)."""
    tot, matched, errs, ious = 0, 0, [], []
    for r, g in zip(ref, got):
        if len(r) == 0:
            continue
        rb = r.boxes.data
        gb = g.boxes.data.cpu() if torch.is_tensor(g.boxes.data) else torch.as_tensor(g.boxes.data)
        if len(gb) == 0:
            tot += len(rb)
            continue
        m = box_iou_matrix(rb[:, :4], gb[:, :4]) * (rb[:, 5:6] == gb[None, :, 5]).float()
        best, j = m.max(1)
        ok = best > 0.9
        tot += len(rb)
        matched += int(ok.sum())
        errs += (rb[ok, :4] - gb[j[ok], :4]).abs().max(1).values.tolist()
        ious += mask_iou(r.masks.data[ok], torch.as_tensor(g.masks.data).cpu()[j[ok]]).tolist()
    return matched / max(tot, 1), errs, ious


def test_end_to_end_vs_fp32_oracle_matched_detections(nseg):
    """bf16 network vs the fp32 oracle cannot be bit-compared and random-init networks amplify
    rounding noise, so the yardstick is the oracle itself run with bf16 storage emulation: the engine's drift
    from fp32 must not exceed the drift ANY faithful bf16 implementation shows (plus a small margin)."""
    from oracle import OracleYOLO
    yolo, frames = nseg["yolo"], nseg["frames"]
    net32, _ = oracle_with_synth("yolov8n-seg", emulate=False)
    ref = OracleYOLO(net32).predict(frames, conf=0.25, iou=0.7, retina_masks=True)
    emu = OracleYOLO(nseg["net"]).predict(frames, conf=0.25, iou=0.7, retina_masks=True)
    got = yolo.predict(frames, conf=0.25, iou=0.7, retina_masks=True)
    rate_e, errs_e, ious_e = _drift(ref, emu)
    rate_g, errs_g, ious_g = _drift(ref, got)
    report(test="e2e_fp32", engine={"match_rate": rate_g, "box_err_median": float(np.median(errs_g)),
                                    "box_err_p99": float(np.percentile(errs_g, 99)), "mask_iou_median": float(np.median(ious_g))},
           bf16_emulating_oracle={"match_rate": rate_e, "box_err_median": float(np.median(errs_e)),
                                  "box_err_p99": float(np.percentile(errs_e, 99)), "mask_iou_median": float(np.median(ious_e))})
    assert rate_g >= rate_e - 0.05
    assert np.median(errs_g) <= 1.25 * np.median(errs_e) + 0.05
    assert np.median(ious_g) >= np.median(ious_e) - 0.02
    # and the engine agrees with the emulating oracle far better than either agrees with fp32
    rate_x, errs_x, ious_x = _drift(emu, got)
    report(test="e2e_vs_emulating_oracle", match_rate=rate_x, box_err_median=float(np.median(errs_x)),
           mask_iou_median=float(np.median(ious_x)))
    assert rate_x >= rate_g


def test_predict_api_sources_order_and_batch_invariance(nseg):
    from PIL import Image
    from yolo_puncture_b200 import synth
    yolo, frames = nseg["yolo"], nseg["frames"]
    one = yolo.predict(frames[0], conf=0.25, retina_masks=True)          # single ndarray (reference app.py:91)
    assert isinstance(one, list) and len(one) == 1 and one[0].orig_shape == (640, 640)
    allr = yolo(frames, conf=0.25, retina_masks=True)                    # __call__ == predict, list source
    assert torch.equal(one[0].boxes.data, allr[0].boxes.data)            # batch-size invariant, bit for bit
    assert torch.equal(one[0].masks.raw, allr[0].masks.raw)
    pil = Image.fromarray(frames[2][:, :, ::-1].copy())                  # PIL is RGB -> same result as the BGR array
    pr = yolo.predict(pil, conf=0.25, retina_masks=True)
    assert torch.equal(pr[0].boxes.data, allr[2].boxes.data)
    small = synth.synth_frame(9, 480, 640)                               # ragged batch: shapes differ -> square letterbox
    mixed = yolo.predict([frames[0], small, frames[2]], conf=0.25, retina_masks=True)
    assert [r.orig_shape for r in mixed] == [(640, 640), (480, 640), (640, 640)]
    assert mixed[1].masks is None or mixed[1].masks.data.shape[1:] == (480, 640)
    pb = allr[0].boxes.cpu().numpy()
    assert pb.xyxy.shape[1] == 4 and pb.conf.dtype == np.float32 and (np.diff(pb.conf) <= 0).all()
    poly = allr[0].masks.xy[int(np.argmax(pb.conf))]
    assert poly.ndim == 2 and poly.shape[1] == 2
    assert next(yolo.model.parameters()).device.type == "cuda"
    sp = allr[0].speed
    assert set(sp) == {"preprocess", "inference", "postprocess"} and all(v >= 0 for v in sp.values())


def test_pass_schedule_invariance(nseg):
    """predict() splits a large group into a 16-frame head pass (second engine plan) + the rest; any schedule must
    give the results of one frame at a time, bit for bit, pinned or pageable frames alike."""
    from yolo_puncture_b200 import synth
    yolo = nseg["yolo"]
    frames = synth.synth_frames(40)
    assert [hi - lo for (_, lo, hi, _, _) in yolo._schedule(40)] == [16, 24]
    auto = yolo.predict(frames, conf=0.25, retina_masks=True, batch=64)
    pin = torch.empty((40, 640, 640, 3), dtype=torch.uint8).pin_memory()
    for i, f in enumerate(frames):
        pin[i] = torch.from_numpy(f)
    auto_pin = yolo.predict([pin[i].numpy() for i in range(40)], conf=0.25, retina_masks=True, batch=64)
    try:
        yolo.micro_batch = 7  # uniform passes with a ragged tail
        assert [hi - lo for (_, lo, hi, _, _) in yolo._schedule(40)] == [7, 7, 7, 7, 7, 5]
        uni = yolo.predict(frames, conf=0.25, retina_masks=True, batch=64)
    finally:
        yolo.micro_batch = None
    n = 0
    for i in (0, 15, 16, 39):
        one = yolo.predict(frames[i], conf=0.25, retina_masks=True)[0]
        for r in (auto[i], auto_pin[i], uni[i]):
            assert torch.equal(one.boxes.data, r.boxes.data)
            assert (one.masks is None) == (r.masks is None)
            if one.masks is not None:
                assert torch.equal(one.masks.raw, r.masks.raw)
        n += len(one.boxes)
    for a, b, c in zip(auto, auto_pin, uni):
        assert torch.equal(a.boxes.data, b.boxes.data) and torch.equal(a.boxes.data, c.boxes.data)
    assert n > 0


@pytest.mark.parametrize("name", ["yolov8s-seg", "yolov8m-seg"])
def test_other_scales_strict(name):
    from oracle import ops as oops
    from yolo_puncture_b200 import YOLO, synth
    net, sd = oracle_with_synth(name, emulate=True)
    yolo = YOLO(name, state_dict=sd, device=0)
    frames = synth.synth_frames(2, start=3)
    res = yolo.predict(frames, conf=0.25, iou=0.7, retina_masks=True)
    eng = yolo.engine
    with torch.no_grad():
        feats = net.features(oops.preprocess(frames, 640))
    for vname in ("model.9", "model.15", "model.21"):
        got, ref = eng.view(vname).float().cpu(), feats[int(vname.split(".")[1])].permute(0, 2, 3, 1)
        assert float((got - ref).abs().mean() / ref.abs().mean()) < 0.03, vname
    dets, kept, proto = oracle_select_on_engine_tensors(eng, net, 2, [(80, 80), (40, 40), (20, 20)], 0.25, 0.7)
    for b in range(2):
        n = len(res[b])
        assert n == len(dets[b]) and eng.keep[b, :n].cpu().long().tolist() == kept[b].tolist()
        if n:
            d = dets[b].clone()
            d[:, :4] = oops.scale_boxes((640, 640), d[:, :4], (640, 640))
            mo = oops.process_mask_native(proto[b], d[:, 6:], d[:, :4], (640, 640))
            assert mask_iou(mo, res[b].masks.data.cpu()).min() >= 0.99


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configs: the C5 model (yolov8x-seg), the C4 model AND geometry (yolov8m-seg, 1080p -> 736x1280),
# and the C3 bench shape (yolov8s-seg, B = 64).  Layers vs the bf16-emulating oracle, head rows included; kept anchors
# and class ids bit-exact, boxes <= 1e-2 px and masks IoU >= 0.99 against the oracle's post-processing.
# ---------------------------------------------------------------------------------------------------
def _strict_selection_and_masks(yolo, net, res, net_hw, orig_hw, tag):
    from oracle import ops as oops
    eng, B = yolo.engine, len(res)
    H, W = net_hw
    shapes = [(H // s, W // s) for s in (8, 16, 32)]
    dets, kept, proto = oracle_select_on_engine_tensors(eng, net, B, shapes, 0.25, 0.7)
    ndet = 0
    for b in range(B):
        n = len(res[b])
        assert n == len(dets[b]), f"{tag} frame {b}: {n} detections vs oracle {len(dets[b])}"
        assert eng.keep[b, :n].cpu().long().tolist() == kept[b].tolist(), f"{tag} frame {b}: kept anchors differ"
        if n == 0:
            assert res[b].masks is None
            continue
        ndet += n
        d = dets[b].clone()
        d[:, :4] = oops.scale_boxes(net_hw, d[:, :4], orig_hw)
        got = res[b].boxes.data.cpu()
        assert torch.equal(got[:, 5], d[:, 5])
        assert (got[:, :4] - d[:, :4]).abs().max() <= 1e-2
        assert (got[:, 4] - d[:, 4]).abs().max() <= 1e-6
        mo = oops.process_mask_native(proto[b], d[:, 6:], d[:, :4], orig_hw)
        me = res[b].masks.data.cpu()
        assert me.shape == mo.shape and mask_iou(mo, me).min() >= 0.99
    return ndet


def _layers_and_head(eng, net, frames, imgsz, rows, tol, tag):
    """Engine activations of batch rows `rows` vs the bf16-emulating oracle run on just those frames."""
    from oracle import ops as oops
    with torch.no_grad():
        feats = net.features(oops.preprocess([frames[i] for i in rows], imgsz))
    worst = 0.0
    for vname in eng.view_table():
        if not vname.startswith("model."):
            continue
        ref = feats[int(vname.split(".")[1])]
        if not torch.is_tensor(ref):
            continue
        got, ref = eng.view(vname)[rows].float().cpu(), ref.permute(0, 2, 3, 1)
        rel = float((got - ref).abs().mean() / ref.abs().mean())
        worst = max(worst, rel)
        assert rel < tol, f"{tag} {vname}: mean relative error {rel:.4f}"
    pred, (maps, mc, proto) = feats[-1]
    raw = torch.cat([torch.cat([m.flatten(2) for m in maps], 2), mc], 1).permute(0, 2, 1)
    head = eng.view("head")[rows, 0].float().cpu()
    r_head = float((head - raw).abs().mean() / raw.abs().mean())
    r_proto = float((eng.view("proto")[rows].float().cpu() - proto.permute(0, 2, 3, 1)).abs().mean() / proto.abs().mean())
    assert r_head < tol and r_proto < tol, f"{tag}: head {r_head:.4f} proto {r_proto:.4f}"
    return worst, r_head, r_proto


def test_config_c5_model_yolov8x_seg_b4_strict():
    """Cin 80 / 160 / 320 / 640: ragged 64-channel chunks in every layer."""
    from yolo_puncture_b200 import YOLO, synth
    net, sd = oracle_with_synth("yolov8x-seg", emulate=True)
    yolo = YOLO("yolov8x-seg", state_dict=sd, device=0)
    frames = synth.synth_frames(4, start=11)
    res = yolo.predict(frames, conf=0.25, iou=0.7, retina_masks=True)
    assert yolo.engine.device_error() == 0 and yolo.engine.shape == (4, 640, 640)
    worst, r_head, r_proto = _layers_and_head(yolo.engine, net, frames, 640, [0, 1, 2, 3], 0.03, "x-seg")
    n = _strict_selection_and_masks(yolo, net, res, (640, 640), (640, 640), "x-seg")
    report(test="c5_model", model="yolov8x-seg", B=4, detections=n, worst_layer=worst, head=r_head, proto=r_proto)
    assert n > 0


def test_config_c4_yolov8m_seg_1080p_b2_strict():
    """The C4 model on the C4 geometry: 1920x1080 frames, imgsz 1280 -> 736x1280 net input (device LetterBox), 1080p masks.
    Weights: the class shift calibrated on 1080p frames (what bench.py's C4 workload runs), ~50 detections per frame."""
    from yolo_puncture_b200 import YOLO, synth
    net, sd = oracle_with_synth("yolov8m-seg", emulate=True, geometry=(1080, 1920))
    yolo = YOLO("yolov8m-seg", state_dict=sd, device=0)
    frames = synth.synth_frames(2, 1080, 1920, start=40)
    res = yolo.predict(frames, conf=0.25, iou=0.7, retina_masks=True, imgsz=1280)
    eng = yolo.engine
    assert eng.device_error() == 0 and eng.shape == (2, 736, 1280) and eng.anchors == 19320
    worst, r_head, r_proto = _layers_and_head(eng, net, frames, 1280, [0, 1], 0.03, "m-seg 1080p")
    n = _strict_selection_and_masks(yolo, net, res, (736, 1280), (1080, 1920), "m-seg 1080p")
    assert all(r.masks is None or r.masks.data.shape[1:] == (1080, 1920) for r in res)
    report(test="c4", model="yolov8m-seg", B=2, detections=n, worst_layer=worst, head=r_head, proto=r_proto)
    assert n >= 20


def test_config_c3_bench_shape_yolov8s_seg_b64_strict():
    """The bench shape itself: one engine pass of 64 frames (the [16, 48] schedule is covered by
    test_pass_schedule_invariance); selection and masks strict on all 64 frames, layers on first / middle / last row."""
    from yolo_puncture_b200 import YOLO, synth
    net, sd = oracle_with_synth("yolov8s-seg", emulate=True)
    yolo = YOLO("yolov8s-seg", state_dict=sd, device=0)
    yolo.micro_batch = 64
    frames = synth.synth_frames(64)
    res = yolo.predict(frames, conf=0.25, iou=0.7, retina_masks=True, batch=64)
    assert yolo.engine.device_error() == 0 and yolo.engine.shape == (64, 640, 640)
    worst, r_head, r_proto = _layers_and_head(yolo.engine, net, frames, 640, [0, 31, 63], 0.03, "s-seg b64")
    n = _strict_selection_and_masks(yolo, net, res, (640, 640), (640, 640), "s-seg b64")
    report(test="c3_bench_shape", model="yolov8s-seg", B=64, detections=n, worst_layer=worst, head=r_head, proto=r_proto)
    assert n > 500


# ---------------------------------------------------------------------------------------------------
# The north_star bar itself, asserted outright against the FP32 oracle: boxes of matched detections within 1e-2 px,
# masks at IoU >= 0.99.  A bf16 pipeline can only meet it on a network whose outputs are as decisive as a trained one's
# and which does not amplify rounding noise from layer to layer, hence the "damped" weight recipe (synth.py RECIPES,
# DESIGN.md section 4; the default recipe's drift is reported and bounded by the test further up).
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,nframes", [("yolov8n-seg", 6), ("yolov8s-seg", 4)])
def test_north_star_bar_vs_fp32_oracle_on_damped_recipe(name, nframes):
    from oracle import OracleYOLO
    from yolo_puncture_b200 import YOLO, synth
    net32, sd = oracle_with_synth(name, emulate=False, recipe="damped")
    yolo = YOLO(name, state_dict=sd, device=0)
    frames = synth.synth_frames(nframes, structured=synth.RECIPES["damped"]["structured"])
    ref = OracleYOLO(net32).predict(frames, conf=0.25, iou=0.7, retina_masks=True)
    got = yolo.predict(frames, conf=0.25, iou=0.7, retina_masks=True)
    assert yolo.engine.device_error() == 0
    rate, errs, ious, tot = drift_stats(ref, got)
    errs, ious = np.array(errs), np.array(ious)
    report(test="north_star_damped", model=name, reference_detections=tot, match_rate=rate,
           box_err_px={"median": float(np.median(errs)), "p99": float(np.percentile(errs, 99)), "max": float(errs.max())},
           mask_iou={"median": float(np.median(ious)), "p1": float(np.percentile(ious, 1)), "min": float(ious.min())})
    assert tot >= 100 * nframes // 2 and rate >= 0.95
    if name == "yolov8n-seg":  # every matched detection
        assert errs.max() <= 1e-2
        assert ious.min() >= 0.99
    else:  # wider nets: a handful of outliers in ~1000 detections (one boundary row of a mask, one DFL side)
        assert np.percentile(errs, 99) <= 1e-2 and errs.max() <= 5e-2
        assert np.percentile(ious, 1) >= 0.99 and ious.min() >= 0.97
    # layer-level view of the same run: elementwise error of the engine against the bf16-emulating oracle, in bf16 ulps
    net_e, _ = oracle_with_synth(name, emulate=True, recipe="damped")
    from oracle import ops as oops
    with torch.no_grad():
        feats = net_e.features(oops.preprocess(frames[:2], 640))
    for vname in ("model.1", "model.4", "model.9", "model.15", "model.21"):
        mx, p999, mean, share = ulp_report(yolo.engine.view(vname)[:2].cpu(), feats[int(vname.split(".")[1])].permute(0, 2, 3, 1))
        report(test="ulp_vs_emulating_oracle", model=name, recipe="damped", layer=vname, max_ulp=mx, p999_ulp=p999,
               mean_ulp=mean, share_gt_1ulp=share)


def test_small_max_det_on_a_fresh_engine_addresses_later_images_correctly():
    """The output arrays keep 300 rows per image whatever max_det asks for (include/ypb200.h).  Regression: the NMS
    kernel once strode its outputs by max_det while the mask decode and the host wrappers strode by 300, which only
    showed on a FRESH engine (the shared fixture's buffers still held the right rows from earlier max_det=300 calls)."""
    from oracle import ops as oops
    from yolo_puncture_b200 import YOLO, synth
    net, sd = oracle_with_synth("yolov8n-seg", emulate=True)
    yolo = YOLO("yolov8n-seg", state_dict=sd, device=0)
    frames = [synth.synth_frame(i) for i in (0, 2, 0, 2, 2)]  # frames with detections in every slot (B >= 4: two batch halves)
    res = yolo.predict(frames, conf=0.25, iou=0.7, max_det=5, retina_masks=True)
    eng = yolo.engine
    dets, kept, proto = oracle_select_on_engine_tensors(eng, net, len(frames), [(80, 80), (40, 40), (20, 20)], 0.25, 0.7, 5)
    for b in range(len(frames)):
        n = len(res[b])
        assert n == len(dets[b]) == 5
        assert eng.keep[b, :n].cpu().long().tolist() == kept[b].tolist()
        d = dets[b].clone()
        d[:, :4] = oops.scale_boxes((640, 640), d[:, :4], (640, 640))
        got = res[b].boxes.data.cpu()
        assert torch.equal(got[:, 5], d[:, 5]) and (got[:, :4] - d[:, :4]).abs().max() <= 1e-2
        mo = oops.process_mask_native(proto[b], d[:, 6:], d[:, :4], (640, 640))
        assert mask_iou(mo, res[b].masks.data.cpu()).min() >= 0.99


def test_retina_masks_of_a_frame_smaller_than_the_proto_window(nseg):
    """A 100x100 crop at imgsz 640 (proto 160x160 > frame): upstream's scale_masks down-samples; so does the engine."""
    from yolo_puncture_b200 import synth
    net, yolo, oops = nseg["net"], nseg["yolo"], nseg["oops"]
    frames = [synth.synth_frame(i)[100:200, 300:400].copy() for i in (0, 2)]
    res = yolo.predict(frames, conf=0.25, iou=0.7, retina_masks=True)
    eng = yolo.engine
    assert eng.shape == (2, 640, 640)
    dets, kept, proto = oracle_select_on_engine_tensors(eng, net, 2, [(80, 80), (40, 40), (20, 20)], 0.25, 0.7)
    n_tot = 0
    for b in range(2):
        n = len(res[b])
        assert n == len(dets[b])
        if n == 0:
            continue
        n_tot += n
        d = dets[b].clone()
        d[:, :4] = oops.scale_boxes((640, 640), d[:, :4], (100, 100))
        assert (res[b].boxes.data.cpu()[:, :4] - d[:, :4]).abs().max() <= 1e-2
        mo = oops.process_mask_native(proto[b], d[:, 6:], d[:, :4], (100, 100))
        me = res[b].masks.data.cpu()
        assert me.shape == (n, 100, 100) and int((mo != me).sum()) <= max(2, int(0.001 * mo.numel()))
    assert n_tot > 0
