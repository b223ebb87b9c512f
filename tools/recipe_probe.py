"""CPU probe (test infrastructure): how far a bf16-storage forward drifts from the fp32 oracle under a synthetic
weight recipe, per layer and at the outputs (boxes of matched detections, mask IoU).  Used to design the
non-chaotic "damped" recipe of yolo_puncture_b200/synth.py.

    python tools/recipe_probe.py [model] [recipe] [n_frames]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gpu_util import box_iou_matrix, mask_iou  # noqa: E402
from oracle import OracleYOLO, ops as oops  # noqa: E402
from oracle.model import build_model  # noqa: E402
from yolo_puncture_b200 import synth  # noqa: E402


def nets(name, recipe):
    out = []
    calib = None
    if recipe != "default" and f"{name}:0:{recipe}" not in synth.load_calibration():
        from oracle.calibrate_synth import calibrate
        calib = calibrate(name, recipe=recipe)  # not committed yet: calibrate in-process
        print("cls_shift", calib["cls_shift"])
    for emu in (False, True):
        net = build_model(name)
        sd = synth.synth_state_dict([(k, v.shape) for k, v in net.state_dict().items()], name, recipe=recipe, calib=calib)
        net.load_state_dict(sd)
        net.fuse()
        if emu:
            net.set_emulation(True)
        out.append(net)
    return out


def drift(ref, got):
    tot, matched, errs, ious = 0, 0, [], []
    for r, g in zip(ref, got):
        if len(r) == 0:
            continue
        rb, gb = r.boxes.data, g.boxes.data
        tot += len(rb)
        if len(gb) == 0:
            continue
        m = box_iou_matrix(rb[:, :4], gb[:, :4]) * (rb[:, 5:6] == gb[None, :, 5]).float()
        best, j = m.max(1)
        ok = best > 0.9
        matched += int(ok.sum())
        errs += (rb[ok, :4] - gb[j[ok], :4]).abs().max(1).values.tolist()
        if r.masks is not None and g.masks is not None:
            ious += mask_iou(r.masks.data[ok], g.masks.data[j[ok]]).tolist()
    return matched / max(tot, 1), np.array(errs), np.array(ious), tot


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "yolov8n-seg"
    recipe = sys.argv[2] if len(sys.argv) > 2 else "default"
    nf = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    torch.set_num_threads(os.cpu_count() or 1)
    n32, nemu = nets(name, recipe)
    frames = synth.synth_frames(nf, structured=synth.RECIPES[recipe]["structured"])
    with torch.no_grad():
        im = oops.preprocess(frames, 640)
        f32 = n32.features(im)
        femu = nemu.features(im)
    for i, (a, b) in enumerate(zip(f32, femu)):
        if torch.is_tensor(a):
            rel = float((a - b).abs().mean() / a.abs().mean())
            print(f"layer {i:2d} mean-rel-err {rel:.5f}  |x| {float(a.abs().mean()):.3f}")
    ref = OracleYOLO(n32).predict(frames, conf=0.25, iou=0.7, retina_masks=True)
    emu = OracleYOLO(nemu).predict(frames, conf=0.25, iou=0.7, retina_masks=True)
    rate, errs, ious, tot = drift(ref, emu)
    print(f"{name} recipe={recipe}: dets/frame {[len(r) for r in ref]} match {rate:.3f} of {tot}")
    if len(errs):
        print(f"  box err px: median {np.median(errs):.4f} p90 {np.percentile(errs, 90):.4f} p99 {np.percentile(errs, 99):.4f} max {errs.max():.4f}")
    if len(ious):
        print(f"  mask IoU: median {np.median(ious):.4f} p10 {np.percentile(ious, 10):.4f} min {ious.min():.4f}")


if __name__ == "__main__":
    main()
