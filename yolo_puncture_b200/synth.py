"""Synthetic weights and frames (SURVEY.md §8d recipe).

There is no network for checkpoints or datasets, and the reference ships neither weights nor sample
videos (reference .gitignore:148,165), so benchmarks and parity tests run on deterministic synthetic
inputs.  Every tensor is seeded by its own *name*, so the engine (which enumerates weights from its
native weight table) and the test oracle (which enumerates its `state_dict`) obtain bit-identical
tensors independent of enumeration order.  Key names follow the upstream checkpoint layout
 — the same `state_dict` a real `YOLO(path)` (reference yolo_seg/app.py:45) carries.
"""

import json
import math
import os
import zlib

import numpy as np
import torch

_CALIB_PATH = os.path.join(os.path.dirname(__file__), "synth_calibration.json")
# Open-loop random init either collapses onto the BN biases or explodes within ~60 layers, so the
# recipe is closed per layer the way trained BatchNorm statistics are: each Conv's running_mean /
# running_var are centred on that layer's measured pre-BN statistics (two scalars per Conv module,
# measured once on frame 0 by oracle/calibrate_synth.py and committed in synth_calibration.json).


# Weight recipes.  "default": SURVEY.md 8d.  "damped": the same tensors with every BatchNorm gain scaled by `gamma`
# (pre-activations stay in SiLU's near-linear range, so a rounding perturbation is not amplified faster than the signal
# itself from layer to layer: a random-init net at full gain sits in the chaotic phase, a trained one does not), see
# tools/recipe_probe.py and DESIGN.md section 4.
# "damped" also gives the head the decisiveness of a trained one, without which no bf16 pipeline can meet a 1e-2 px /
# IoU 0.99 bar against fp32 (a logit error of 0.3 % moves a flat DFL expectation by ~1 px):
#   * DFL prior: the box branch's final bias is a peak -dfl_alpha*(k - k0)^2 per side (k0 per level and side below), so
#     the softmax over the 16 bins is as peaked as a trained model's; bin k0+1 keeps a prior mass of DFL_P1 so that the
#     expectation sits ~0.002 bin above k0 and the features move it (box edges then stay clear of integer pixel
#     coordinates, where a 1e-4 px difference would flip a whole row of the cropped mask);
#   * bias prototype: proto channel 0 is the constant plane SiLU(2) (BN gain 0, beta 2), its coefficient is a per-level
#     constant calibrated so that the mean mask logit sits 2.5 sigma above zero (masks fill most of their box; few pixels
#     have a logit within rounding noise of the threshold), like the saturated mask logits of a trained Proto head.
RECIPES = {"default": {"gamma": 1.0, "dfl_alpha": 0.0, "bias_proto": False, "structured": True},
           "damped": {"gamma": 0.5, "dfl_alpha": 8.0, "bias_proto": True, "structured": False}}
DFL_K0 = [[3, 4, 5, 4], [4, 3, 4, 5], [2, 3, 3, 2]]  # [level][side l,t,r,b], in bins (x stride = pixels)
BIAS_PROTO_BETA = 2.0
DFL_P1 = 0.002


def _gen(name, seed):
    g = torch.Generator()
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _uniform(shape, lo, hi, g):
    return torch.rand(shape, generator=g) * (hi - lo) + lo


def load_calibration():
    if os.path.exists(_CALIB_PATH):
        with open(_CALIB_PATH) as f:
            return json.load(f)
    return {}


def synth_state_dict(specs, model_name, seed=0, calib=None, nc=80, recipe="default", geometry=None):
    """specs: iterable of (name, shape).  Returns {name: fp32 tensor} following the recipe:
    conv weights U(+-sqrt(3/fan_in)); BN gamma~U(.5,1.5), beta~N(0,.1), running_mean = mu_l +
    N(0,.1)*sqrt(v_l), running_var = v_l*U(.5,1.5) with (mu_l, v_l) the layer's calibrated pre-BN
    statistics; box-branch final bias 1.0; class-branch final bias log(5/nc/(640/s)^2) + per-level
    calibrated shift (so a few hundred candidates per frame pass conf=0.25); mask-coefficient final
    bias ~N(0,1).  calib: {"bn": {conv_module_name: [mu, v]}, "cls_shift": [s0, s1, s2]}.
    geometry=(h, w): frames of that size - if the calibration file has a class shift derived on such frames
    (key "<model>:<seed>[:recipe]@<h>x<w>", oracle/calibrate_synth.py --geometry) it replaces the 640x640 one; every
    other tensor is unchanged."""
    if calib is None:
        key = f"{model_name}:{seed}" if recipe == "default" else f"{model_name}:{seed}:{recipe}"
        table = load_calibration()
        calib = table.get(key, {})
        if geometry is not None:
            over = table.get(f"{key}@{int(geometry[0])}x{int(geometry[1])}")
            if over:
                calib = dict(calib, **over)
    rp = RECIPES[recipe]
    cls_bias_shift = calib.get("cls_shift", [0.0, 0.0, 0.0])
    bn_stats = calib.get("bn", {})
    strides = [8.0, 16.0, 32.0]
    sd = {}
    for name, shape in specs:
        shape = tuple(int(s) for s in shape)
        g = _gen(name, seed)
        leaf = name.rsplit(".", 1)[-1]
        parts = name.split(".")
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.zeros(shape, dtype=torch.long)
            continue
        if name.endswith("dfl.conv.weight"):
            sd[name] = torch.arange(16, dtype=torch.float32).view(shape)
            continue
        is_bn = ".bn." in name
        if is_bn:
            mu_l, v_l = bn_stats.get(name[: name.index(".bn.")], (0.0, 1.0))
            if leaf == "weight":
                t = _uniform(shape, 0.5, 1.5, g) * rp["gamma"]
            elif leaf == "bias":
                t = torch.randn(shape, generator=g) * 0.1
            elif leaf == "running_mean":
                t = torch.randn(shape, generator=g) * (0.1 * math.sqrt(v_l)) + mu_l
            elif leaf == "running_var":
                t = _uniform(shape, 0.5, 1.5, g) * v_l
            else:
                raise ValueError(name)
            if rp["bias_proto"] and ".proto.cv3.bn." in name and leaf in ("weight", "bias"):
                t[0] = 0.0 if leaf == "weight" else BIAS_PROTO_BETA
            sd[name] = t
            continue
        if leaf == "weight":
            if "upsample" in parts:  # ConvTranspose2d (cin, cout, 2, 2): each output pixel sees cin taps
                fan_in = shape[0]
            else:
                fan_in = shape[1] * shape[2] * shape[3]
            b = math.sqrt(3.0 / fan_in)
            t = _uniform(shape, -b, b, g)
            if rp["bias_proto"] and "cv4" in parts and parts[-2] == "2":
                t[0] = 0.0  # the bias prototype's coefficient is the calibrated constant alone
            sd[name] = t
            continue
        if leaf == "bias":
            # final 1x1 convs of the head branches and the ConvTranspose
            branch = next((p for p in parts if p in ("cv2", "cv3", "cv4", "one2one_cv2", "one2one_cv3", "upsample")), None)
            if branch in ("cv2", "one2one_cv2"):
                t = torch.full(shape, 1.0)
                if rp["dfl_alpha"] > 0:
                    lvl = int(parts[parts.index(branch) + 1])
                    k = torch.arange(16, dtype=torch.float32)
                    sides = []
                    for k0 in DFL_K0[lvl]:
                        side = -rp["dfl_alpha"] * (k - k0) ** 2
                        side[k0 + 1] = math.log(DFL_P1)
                        sides.append(side)
                    t = torch.cat(sides)
            elif branch in ("cv3", "one2one_cv3"):
                lvl = int(parts[parts.index(branch) + 1])
                t = torch.full(shape, math.log(5 / nc / (640 / strides[lvl]) ** 2) + float(cls_bias_shift[lvl]))
            elif branch == "cv4":
                t = torch.randn(shape, generator=g)
                if rp["bias_proto"]:
                    lvl = int(parts[parts.index(branch) + 1])
                    t[0] = float(calib.get("coef0_bias", [0.0, 0.0, 0.0])[lvl])
            else:
                t = torch.randn(shape, generator=g) * 0.1
            sd[name] = t
            continue
        raise ValueError(f"unhandled tensor {name}")
    return sd


def _box_blur(a, k):
    c = np.cumsum(np.cumsum(np.pad(a, ((k, 0), (k, 0), (0, 0)), mode="wrap"), 0), 1)
    return (c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]) / (k * k)


def synth_frame(idx, h=640, w=640, structured=True):
    """uint8 BGR HWC frame #idx: three octaves of box-blurred uniform noise (low-frequency structure
    for the convs to see) or plain uniform noise."""
    rng = np.random.default_rng(1234 + idx)
    if not structured:
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    acc = np.zeros((h, w, 3), np.float64)
    for k, wgt in ((31, 0.5), (9, 0.3), (3, 0.2)):
        n = rng.random((h, w, 3))
        b = _box_blur(n, k)
        b = (b - b.mean()) / (b.std() + 1e-9)
        acc += wgt * b
    acc = (acc - acc.min()) / (acc.max() - acc.min() + 1e-9)
    return np.clip(acc * 255.0 + 0.5, 0, 255).astype(np.uint8)


def synth_frames(n, h=640, w=640, start=0, structured=True):
    return [synth_frame(start + i, h, w, structured) for i in range(n)]
