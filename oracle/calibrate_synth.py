"""Offline calibration of the synthetic-weight recipe (test infrastructure; see oracle/__init__.py).

Writes yolo_puncture_b200/synth_calibration.json: per model, (i) for every Conv module the mean and
variance of its pre-BatchNorm output on synthetic frame 0 (two scalars per layer, so the random BN
statistics are centred the way trained ones are), and (ii) the per-level class-bias shift that puts
the 98th percentile of the per-anchor max class logit at logit(0.25).

    python -m oracle.calibrate_synth [--recipe=damped] [model ...]
    python -m oracle.calibrate_synth [--recipe=damped] --geometry=1080x1920@1280 model ...   (class shift for another frame size)
"""

import json
import math
import sys

import torch

from yolo_puncture_b200 import synth
from . import ops
from .model import MODEL_SPECS, build_model
from .modules import Conv


@torch.no_grad()
def calibrate(name, seed=0, pct=0.98, conf=0.25, recipe="default"):
    net = build_model(name)
    specs = [(k, v.shape) for k, v in net.state_dict().items()]
    mod_names = {id(m): n for n, m in net.named_modules()}
    bn = {}

    def conv_forward(self, x):
        y = self.conv(x)
        mname = mod_names[id(self)]
        mu, v = float(y.mean()), float(y.var())
        bn[mname] = [mu, max(v, 1e-12)]
        part = synth.synth_state_dict(
            [(f"{mname}.bn.{leaf}", self.bn.weight.shape) for leaf in ("weight", "bias", "running_mean", "running_var")],
            name, seed, calib={"bn": bn}, recipe=recipe)
        self.bn.weight.copy_(part[f"{mname}.bn.weight"])
        self.bn.bias.copy_(part[f"{mname}.bn.bias"])
        self.bn.running_mean.copy_(part[f"{mname}.bn.running_mean"])
        self.bn.running_var.copy_(part[f"{mname}.bn.running_var"])
        return self.act(self.bn(y))

    net.load_state_dict(synth.synth_state_dict(specs, name, seed, calib={}, recipe=recipe))
    im = ops.preprocess([synth.synth_frame(0, structured=synth.RECIPES[recipe]["structured"])], 640)
    orig = Conv.forward
    Conv.forward = conv_forward
    try:
        net(im)
    finally:
        Conv.forward = orig
    # class-bias shift per level, on the calibrated net
    calib = {"bn": bn, "cls_shift": [0.0, 0.0, 0.0]}
    net.load_state_dict(synth.synth_state_dict(specs, name, seed, calib=calib, recipe=recipe))
    feats = net.features(im, upto=len(net.model) - 1)
    head = net.model[-1]
    maps = head.head_maps([feats[j] for j in net.froms[-1]])
    target = math.log(conf / (1 - conf))
    shifts = []
    for mp in maps:
        amax = mp[:, 64:].amax(1).flatten()
        q = torch.quantile(amax, pct).item()
        shifts.append(target - q)
    calib["cls_shift"] = shifts
    if synth.RECIPES[recipe]["bias_proto"] and hasattr(head, "proto"):
        # coefficient of the constant prototype, per level: mean mask logit = 2.5 x its spatial standard deviation
        net.load_state_dict(synth.synth_state_dict(specs, name, seed, calib=calib, recipe=recipe))
        xs = [feats[j] for j in net.froms[-1]]
        proto = head.proto(xs[0])[0]                                   # (32, mh, mw); channel 0 is the constant plane
        p0 = float(proto[0].mean())
        coef0 = []
        for i, x in enumerate(xs):
            mc = head.cv4[i](x)[0].flatten(1)                          # (32, h*w)
            mc = mc[:, :: max(1, mc.shape[1] // 128)]
            lg = torch.einsum("ka,khw->ahw", mc[1:], proto[1:])
            m, s = float(lg.mean()), float(lg.std((1, 2)).mean())
            coef0.append((2.5 * s - m) / p0)
        calib["coef0_bias"] = coef0
    return calib


@torch.no_grad()
def calibrate_geometry(name, hw, imgsz, base, seed=0, pct=0.98, conf=0.25, recipe="default", n_frames=2):
    """Class-bias shift for frames of another size (e.g. 1080p letterboxed to 736x1280, BASELINE config C4): same
    weights and BN statistics as the base entry, only the per-level shift is re-derived on synthetic frames of that size
    (the 640x640 shift leaves such frames with no candidate at all, i.e. nothing for NMS and the mask decode to do)."""
    net = build_model(name)
    specs = [(k, v.shape) for k, v in net.state_dict().items()]
    calib = {"bn": base["bn"], "cls_shift": [0.0, 0.0, 0.0]}
    if "coef0_bias" in base:
        calib["coef0_bias"] = base["coef0_bias"]
    net.load_state_dict(synth.synth_state_dict(specs, name, seed, calib=calib, recipe=recipe))
    structured = synth.RECIPES[recipe]["structured"]
    im = ops.preprocess([synth.synth_frame(i, hw[0], hw[1], structured=structured) for i in range(n_frames)], imgsz)
    feats = net.features(im, upto=len(net.model) - 1)
    head = net.model[-1]
    maps = head.head_maps([feats[j] for j in net.froms[-1]])
    target = math.log(conf / (1 - conf))
    return {"cls_shift": [target - torch.quantile(mp[:, 64:].amax(1).flatten(), pct).item() for mp in maps]}


def main(argv):
    recipe = "default"
    if argv and argv[0].startswith("--recipe="):
        recipe, argv = argv[0].split("=", 1)[1], argv[1:]
    if argv and argv[0].startswith("--geometry="):  # --geometry=1080x1920@1280 model ...
        geo, argv = argv[0].split("=", 1)[1], argv[1:]
        hw_s, imgsz = geo.split("@")
        hw = tuple(int(v) for v in hw_s.split("x"))
        table = synth.load_calibration()
        for n in argv:
            base_key = f"{n}:0" if recipe == "default" else f"{n}:0:{recipe}"
            key = f"{base_key}@{hw[0]}x{hw[1]}"
            table[key] = calibrate_geometry(n, hw, int(imgsz), table[base_key], recipe=recipe)
            print(key, table[key])
        with open(synth._CALIB_PATH, "w") as f:
            json.dump(table, f, separators=(",", ":"))
        return
    names = argv or list(MODEL_SPECS)
    table = synth.load_calibration()
    for n in names:
        key = f"{n}:0" if recipe == "default" else f"{n}:0:{recipe}"
        table[key] = calibrate(n, recipe=recipe)
        print(key, "layers", len(table[key]["bn"]), "cls_shift", table[key]["cls_shift"], table[key].get("coef0_bias"))
    with open(synth._CALIB_PATH, "w") as f:
        json.dump(table, f, separators=(",", ":"))


if __name__ == "__main__":
    main(sys.argv[1:])
