"""Offline calibration of the synthetic-weight recipe (test infrastructure; see oracle/__init__.py).

Writes yolo_puncture_b200/synth_calibration.json: per model, (i) for every Conv module the mean and
variance of its pre-BatchNorm output on synthetic frame 0 (two scalars per layer, so the random BN
statistics are centred the way trained ones are), and (ii) the per-level class-bias shift that puts
the 98th percentile of the per-anchor max class logit at logit(0.25).

    python -m oracle.calibrate_synth [model ...]
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
def calibrate(name, seed=0, pct=0.98, conf=0.25):
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
            name, seed, calib={"bn": bn})
        self.bn.weight.copy_(part[f"{mname}.bn.weight"])
        self.bn.bias.copy_(part[f"{mname}.bn.bias"])
        self.bn.running_mean.copy_(part[f"{mname}.bn.running_mean"])
        self.bn.running_var.copy_(part[f"{mname}.bn.running_var"])
        return self.act(self.bn(y))

    net.load_state_dict(synth.synth_state_dict(specs, name, seed, calib={}))
    im = ops.preprocess([synth.synth_frame(0)], 640)
    orig = Conv.forward
    Conv.forward = conv_forward
    try:
        net(im)
    finally:
        Conv.forward = orig
    # class-bias shift per level, on the calibrated net
    calib = {"bn": bn, "cls_shift": [0.0, 0.0, 0.0]}
    net.load_state_dict(synth.synth_state_dict(specs, name, seed, calib=calib))
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
    return calib


def main(argv):
    names = argv or list(MODEL_SPECS)
    table = synth.load_calibration()
    for n in names:
        table[f"{n}:0"] = calibrate(n)
        print(n, "layers", len(table[f"{n}:0"]["bn"]), "cls_shift", table[f"{n}:0"]["cls_shift"])
    with open(synth._CALIB_PATH, "w") as f:
        json.dump(table, f, separators=(",", ":"))


if __name__ == "__main__":
    main(sys.argv[1:])
