"""Oracle model builder (test infrastructure; see oracle/__init__.py).

Restates UPSTREAM `cfg/models/v8/yolov8-seg.yaml`, `cfg/models/v10/yolov10n.yaml` and
`nn/tasks.py::parse_model / BaseModel.fuse / _predict_once`.
The reference constructs these through `YOLO(path)` at yolo_seg/app.py:45 and
yolo_seg/yolo_with_deva.py:226.
"""

import math

import torch
import torch.nn as nn

from .modules import (
    C2f, C2fCIB, C2PSA, C3k2, Concat, Conv, Detect, PSA, RepVGGDW, SCDown, SPPF, Segment, v10Detect,
)

# (depth, width, max_channels)
_V8_SCALES = {"n": (0.33, 0.25, 1024), "s": (0.33, 0.50, 1024), "m": (0.67, 0.75, 768),
              "l": (1.00, 1.00, 512), "x": (1.00, 1.25, 512)}
_V10_SCALES = {"n": (0.33, 0.25, 1024), "s": (0.33, 0.50, 1024), "m": (0.67, 0.75, 768),
               "b": (0.67, 1.00, 512), "l": (1.00, 1.00, 512), "x": (1.00, 1.25, 512)}

# rows: (from, repeats, module, args) exactly as in the upstream yaml files
_V8_SEG = [
    (-1, 1, "Conv", [64, 3, 2]), (-1, 1, "Conv", [128, 3, 2]), (-1, 3, "C2f", [128, True]),
    (-1, 1, "Conv", [256, 3, 2]), (-1, 6, "C2f", [256, True]), (-1, 1, "Conv", [512, 3, 2]),
    (-1, 6, "C2f", [512, True]), (-1, 1, "Conv", [1024, 3, 2]), (-1, 3, "C2f", [1024, True]),
    (-1, 1, "SPPF", [1024, 5]),
    (-1, 1, "Upsample", [None, 2, "nearest"]), ([-1, 6], 1, "Concat", [1]), (-1, 3, "C2f", [512]),
    (-1, 1, "Upsample", [None, 2, "nearest"]), ([-1, 4], 1, "Concat", [1]), (-1, 3, "C2f", [256]),
    (-1, 1, "Conv", [256, 3, 2]), ([-1, 12], 1, "Concat", [1]), (-1, 3, "C2f", [512]),
    (-1, 1, "Conv", [512, 3, 2]), ([-1, 9], 1, "Concat", [1]), (-1, 3, "C2f", [1024]),
    ([15, 18, 21], 1, "Segment", ["nc", 32, 256]),
]
_V10N = [
    (-1, 1, "Conv", [64, 3, 2]), (-1, 1, "Conv", [128, 3, 2]), (-1, 3, "C2f", [128, True]),
    (-1, 1, "Conv", [256, 3, 2]), (-1, 6, "C2f", [256, True]), (-1, 1, "SCDown", [512, 3, 2]),
    (-1, 6, "C2f", [512, True]), (-1, 1, "SCDown", [1024, 3, 2]), (-1, 3, "C2f", [1024, True]),
    (-1, 1, "SPPF", [1024, 5]), (-1, 1, "PSA", [1024]),
    (-1, 1, "Upsample", [None, 2, "nearest"]), ([-1, 6], 1, "Concat", [1]), (-1, 3, "C2f", [512]),
    (-1, 1, "Upsample", [None, 2, "nearest"]), ([-1, 4], 1, "Concat", [1]), (-1, 3, "C2f", [256]),
    (-1, 1, "Conv", [256, 3, 2]), ([-1, 13], 1, "Concat", [1]), (-1, 3, "C2f", [512]),
    (-1, 1, "SCDown", [512, 3, 2]), ([-1, 10], 1, "Concat", [1]), (-1, 3, "C2fCIB", [1024, True, True]),
    ([16, 19, 22], 1, "v10Detect", ["nc"]),
]

# UPSTREAM cfg/models/11/yolo11-seg.yaml; parse_model forces c3k=True for the m/l/x scales and builds
# the head with the non-legacy (depthwise) class branch
_V11_SCALES = {"n": (0.50, 0.25, 1024), "s": (0.50, 0.50, 1024), "m": (0.50, 1.00, 512),
               "l": (1.00, 1.00, 512), "x": (1.00, 1.50, 512)}
_V11_SEG = [
    (-1, 1, "Conv", [64, 3, 2]), (-1, 1, "Conv", [128, 3, 2]), (-1, 2, "C3k2", [256, False, 0.25]),
    (-1, 1, "Conv", [256, 3, 2]), (-1, 2, "C3k2", [512, False, 0.25]), (-1, 1, "Conv", [512, 3, 2]),
    (-1, 2, "C3k2", [512, True]), (-1, 1, "Conv", [1024, 3, 2]), (-1, 2, "C3k2", [1024, True]),
    (-1, 1, "SPPF", [1024, 5]), (-1, 2, "C2PSA", [1024]),
    (-1, 1, "Upsample", [None, 2, "nearest"]), ([-1, 6], 1, "Concat", [1]), (-1, 2, "C3k2", [512, False]),
    (-1, 1, "Upsample", [None, 2, "nearest"]), ([-1, 4], 1, "Concat", [1]), (-1, 2, "C3k2", [256, False]),
    (-1, 1, "Conv", [256, 3, 2]), ([-1, 13], 1, "Concat", [1]), (-1, 2, "C3k2", [512, False]),
    (-1, 1, "Conv", [512, 3, 2]), ([-1, 10], 1, "Concat", [1]), (-1, 2, "C3k2", [1024, True]),
    ([16, 19, 22], 1, "Segment", ["nc", 32, 256]),
]

MODEL_SPECS = {f"yolov8{s}-seg": (_V8_SEG, _V8_SCALES[s]) for s in "nsmlx"}
MODEL_SPECS["yolov10n"] = (_V10N, _V10_SCALES["n"])
MODEL_SPECS.update({f"yolo11{s}-seg": (_V11_SEG, _V11_SCALES[s]) for s in "nsmlx"})

_MODULES = {"Conv": Conv, "C2f": C2f, "SPPF": SPPF, "SCDown": SCDown, "PSA": PSA, "C2fCIB": C2fCIB, "C3k2": C3k2,
            "C2PSA": C2PSA}


def make_divisible(x, divisor):
    return math.ceil(x / divisor) * divisor


class OracleModel(nn.Module):
    """Sequential-with-skips container (UPSTREAM nn/tasks.py::BaseModel)."""

    def __init__(self, name, nc=80):
        super().__init__()
        rows, (depth, width, max_ch) = MODEL_SPECS[name]
        self.name, self.nc = name, nc
        self.task = "segment" if name.endswith("-seg") else "detect"
        ch = [3]
        legacy = True  # v8 full-conv class branch; C3k2 models (YOLO11) use the depthwise one
        layers, self.froms, save = [], [], set()
        for i, (f, n, m, args) in enumerate(rows):
            args = [nc if a == "nc" else a for a in args]
            n = max(round(n * depth), 1) if n > 1 else n
            if m in _MODULES:
                c1, c2 = ch[f], args[0]
                c2 = make_divisible(min(c2, max_ch) * width, 8)
                a = [c1, c2, *args[1:]]
                if m in ("C2f", "C2fCIB", "C3k2", "C2PSA"):
                    a.insert(2, n)
                if m == "C3k2":
                    legacy = False
                    if len(a) < 4:
                        a.append(False)
                    if name[len("yolo11")] in "mlx":
                        a[3] = True
                mod = _MODULES[m](*a)
            elif m == "Upsample":
                mod, c2 = nn.Upsample(None, args[1], args[2]), ch[f]
            elif m == "Concat":
                mod, c2 = Concat(args[0]), sum(ch[x] for x in f)
            elif m == "Segment":
                npr = make_divisible(min(args[2], max_ch) * width, 8)
                mod, c2 = Segment(args[0], args[1], npr, [ch[x] for x in f], legacy=legacy), None
            elif m == "v10Detect":
                mod, c2 = v10Detect(args[0], [ch[x] for x in f]), None
            else:
                raise ValueError(m)
            layers.append(mod)
            self.froms.append(f)
            for x in ([f] if isinstance(f, int) else f):
                if x != -1:
                    save.add(x % (i + 1) if x < 0 else x)
            if i == 0:
                ch = []
            ch.append(c2)
        self.model = nn.Sequential(*layers)
        self.save = sorted(save)
        self.names = {i: f"class{i}" for i in range(nc)}
        self.stride = torch.tensor([8.0, 16.0, 32.0])
        self.model[-1].bias_init()
        self.fused = False

    # ------------------------------------------------------------------ forward
    def features(self, x, upto=None):
        """Run layers [0, upto) and return the list of per-layer outputs (None where not kept)."""
        y = []
        n = len(self.model) if upto is None else upto
        for i in range(n):
            f, m = self.froms[i], self.model[i]
            if f != -1:
                x = y[f] if isinstance(f, int) else [x if j == -1 else y[j] for j in f]
            x = m(x)
            y.append(x)
        return y

    def forward(self, x):
        y = []
        for f, m in zip(self.froms, self.model):
            if f != -1:
                x = y[f] if isinstance(f, int) else [x if j == -1 else y[j] for j in f]
            x = m(x)
            y.append(x if len(y) in self.save else None)
        return x

    # ------------------------------------------------------------------ fuse
    @torch.no_grad()
    def fuse(self):
        """Fold every BatchNorm into its conv (UPSTREAM BaseModel.fuse / fuse_conv_and_bn):
        w' = w * gamma / sqrt(var + eps),  b' = beta - mean * gamma / sqrt(var + eps)."""
        if self.fused:
            return self
        for m in self.modules():
            if isinstance(m, Conv) and hasattr(m, "bn"):
                conv, bn = m.conv, m.bn
                fused = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride,
                                  conv.padding, conv.dilation, conv.groups, bias=True).requires_grad_(False)
                w_conv = conv.weight.view(conv.out_channels, -1)
                w_bn = torch.diag(bn.weight.div(torch.sqrt(bn.eps + bn.running_var)))
                fused.weight.copy_(torch.mm(w_bn, w_conv).view(fused.weight.shape))
                b_conv = torch.zeros(conv.out_channels) if conv.bias is None else conv.bias
                b_bn = bn.bias - bn.weight.mul(bn.running_mean).div(torch.sqrt(bn.running_var + bn.eps))
                fused.bias.copy_(torch.mm(w_bn, b_conv.reshape(-1, 1)).reshape(-1) + b_bn)
                m.conv = fused
                del m.bn
        for m in self.modules():
            if isinstance(m, RepVGGDW):
                m.fuse()
        self.fused = True
        return self

    # ------------------------------------------------------------------ bf16 emulation
    @torch.no_grad()
    def set_emulation(self, on=True):
        """Mirror the engine's storage precision (see modules.py docstring).  Requires fuse().
        Conv weights are rounded to bf16, except the stem's (the engine feeds raw uint8 pixel values, exact in bf16, to
        the tensor core against a hi+lo bf16 split of w/255, i.e. fp32-grade products); activations are rounded to bf16 after each block op; the
        final 1x1 head convs (plain nn.Conv2d) and Proto.cv3 keep fp32 outputs."""
        assert self.fused
        mode = "bf16" if on else None
        stem = self.model[0]
        for m in self.modules():
            if hasattr(m, "emu"):
                m.emu = mode
            if on and isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                if m is stem.conv:
                    continue  # the engine splits the stem weights into two bf16 terms: fp32-grade products
                if m.weight.shape[1:] == (16, 1, 1) and m.weight.shape[0] == 1:
                    continue  # DFL arange
                m.weight.copy_(m.weight.to(torch.bfloat16).to(torch.float32))
        head = self.model[-1]
        if isinstance(head, Segment):
            head.proto.cv3.emu = "fp32" if on else None
        # the fused RepVGGDW inner conv feeds a SiLU inside the same kernel: no rounding in between
        for m in self.modules():
            if isinstance(m, RepVGGDW):
                m.conv.emu = None
        return self


def build_model(name, nc=80):
    return OracleModel(name, nc).eval()


def count_parameters(model, exclude_one2many=False):
    tot = 0
    for k, p in model.named_parameters():
        if exclude_one2many and (".cv2." in k or ".cv3." in k) and k.startswith(f"model.{len(model.model) - 1}.") \
                and "one2one" not in k:
            continue
        tot += p.numel()
    return tot


@torch.no_grad()
def conv_flops(model, h=640, w=640, one2one_only=True):
    """Algorithmic conv FLOPs (2*MAC over Conv2d / ConvTranspose2d) for one (1,3,h,w) frame
    (SURVEY.md B.2).  With one2one_only the v10 one-to-many head branches are not counted."""
    total = [0]
    hooks = []
    head = model.model[-1]
    skip = set()
    if one2one_only and getattr(head, "end2end", False):
        for mod in list(head.cv2.modules()) + list(head.cv3.modules()):
            skip.add(id(mod))

    def hook(m, inp, out):
        if id(m) in skip:
            return
        if isinstance(m, nn.ConvTranspose2d):
            macs = inp[0].shape[2] * inp[0].shape[3] * m.in_channels * m.out_channels * m.kernel_size[0] * m.kernel_size[1]
        else:
            if tuple(m.weight.shape) == (1, 16, 1, 1):
                return  # DFL
            macs = out.shape[2] * out.shape[3] * m.out_channels * (m.in_channels // m.groups) * m.kernel_size[0] * m.kernel_size[1]
        total[0] += 2 * macs * inp[0].shape[0]

    for m in model.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            hooks.append(m.register_forward_hook(hook))
    x = torch.zeros(1, 3, h, w)
    if getattr(head, "end2end", False) and one2one_only:
        feats = model.features(x, upto=len(model.model) - 1)
        head.head_maps([feats[j] for j in model.froms[-1]], one2one=True)
    else:
        model(x)
    for hk in hooks:
        hk.remove()
    return total[0]
