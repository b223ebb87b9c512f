"""Oracle building blocks (test infrastructure; see oracle/__init__.py).

Restates UPSTREAM ultralytics 8.3.x `nn/modules/{conv,block,head}.py` and `utils/tal.py`
(SURVEY.md Appendix A.1-A.3).  Attribute names follow upstream so that `state_dict()` keys are
interchangeable with real checkpoints.  These are the blocks behind the
`YOLO(...).predict(...)` calls at reference yolo_seg/app.py:45,91 and yolo_seg/yolo_with_deva.py:51,226.

`emu` (bf16 emulation): when a fused model is put in emulation mode (model.set_emulation(True)),
every Conv rounds its activated output to bf16 (fp32 accumulate), mirroring the storage precision of
the sm_100a engine, so GPU-vs-oracle comparisons isolate real bugs from expected bf16 drift.
"""

import copy
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def autopad(k, p=None, d=1):
    """'same' padding (UPSTREAM nn/modules/conv.py::autopad)."""
    if d > 1:
        k = d * (k - 1) + 1 if isinstance(k, int) else [d * (x - 1) + 1 for x in k]
    if p is None:
        p = k // 2 if isinstance(k, int) else [x // 2 for x in k]
    return p


def _r16(x):
    """Round an fp32 tensor to bf16 precision and return it as fp32."""
    return x.to(torch.bfloat16).to(torch.float32)


class Conv(nn.Module):
    """Conv2d(bias=False) -> BatchNorm2d(eps=1e-3) -> SiLU (UPSTREAM conv.py::Conv)."""

    default_act = nn.SiLU()

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p, d), groups=g, dilation=d, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
        self.act = self.default_act if act is True else act if isinstance(act, nn.Module) else nn.Identity()
        self.emu = None  # None | "bf16" | "fp32"

    def forward(self, x):
        if hasattr(self, "bn"):
            y = self.act(self.bn(self.conv(x)))
        else:  # fused
            y = self.act(self.conv(x))
        return _r16(y) if self.emu == "bf16" else y


class Concat(nn.Module):
    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def forward(self, x):
        return torch.cat(x, self.d)


class Bottleneck(nn.Module):
    """UPSTREAM block.py::Bottleneck."""

    def __init__(self, c1, c2, shortcut=True, g=1, k=(3, 3), e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, k[0], 1)
        self.cv2 = Conv(c_, c2, k[1], 1, g=g)
        self.add = shortcut and c1 == c2
        self.emu = None

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        if self.add:
            y = x + y
            if self.emu == "bf16":
                y = _r16(y)
        return y


class C2f(nn.Module):
    """UPSTREAM block.py::C2f."""

    def __init__(self, c1, c2, n=1, shortcut=False, g=1, e=0.5):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, self.c, shortcut, g, k=((3, 3), (3, 3)), e=1.0) for _ in range(n))

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        y.extend(m(y[-1]) for m in self.m)
        return self.cv2(torch.cat(y, 1))


class SPPF(nn.Module):
    """UPSTREAM block.py::SPPF."""

    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def forward(self, x):
        y = [self.cv1(x)]
        y.extend(self.m(y[-1]) for _ in range(3))
        return self.cv2(torch.cat(y, 1))


class DFL(nn.Module):
    """Distribution focal loss integral (UPSTREAM block.py::DFL)."""

    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1)
        self.c1 = c1

    def forward(self, x):
        b, _, a = x.shape
        return self.conv(x.view(b, 4, self.c1, a).transpose(2, 1).softmax(1)).view(b, 4, a)


class Proto(nn.Module):
    """Mask prototypes (UPSTREAM block.py::Proto)."""

    def __init__(self, c1, c_=256, c2=32):
        super().__init__()
        self.cv1 = Conv(c1, c_, k=3)
        self.upsample = nn.ConvTranspose2d(c_, c_, 2, 2, 0, bias=True)
        self.cv2 = Conv(c_, c_, k=3)
        self.cv3 = Conv(c_, c2)
        self.emu = None

    def forward(self, x):
        y = self.upsample(self.cv1(x))
        if self.emu == "bf16":
            y = _r16(y)
        return self.cv3(self.cv2(y))


# ------------------------------------------------------------------------------------------------
# YOLOv10 blocks
# ------------------------------------------------------------------------------------------------
class SCDown(nn.Module):
    def __init__(self, c1, c2, k, s):
        super().__init__()
        self.cv1 = Conv(c1, c2, 1, 1)
        self.cv2 = Conv(c2, c2, k=k, s=s, g=c2, act=False)

    def forward(self, x):
        return self.cv2(self.cv1(x))


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, attn_ratio=0.5):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.key_dim = int(self.head_dim * attn_ratio)
        self.scale = self.key_dim ** -0.5
        nh_kd = self.key_dim * num_heads
        h = dim + nh_kd * 2
        self.qkv = Conv(dim, h, 1, act=False)
        self.proj = Conv(dim, dim, 1, act=False)
        self.pe = Conv(dim, dim, 3, 1, g=dim, act=False)
        self.emu = None

    def forward(self, x):
        B, C, H, W = x.shape
        N = H * W
        qkv = self.qkv(x)
        q, k, v = qkv.view(B, self.num_heads, self.key_dim * 2 + self.head_dim, N).split(
            [self.key_dim, self.key_dim, self.head_dim], dim=2
        )
        attn = (q.transpose(-2, -1) @ k) * self.scale
        attn = attn.softmax(dim=-1)
        x = (v @ attn.transpose(-2, -1)).view(B, C, H, W) + self.pe(v.reshape(B, C, H, W))
        if self.emu == "bf16":
            x = _r16(x)
        return self.proj(x)


class PSA(nn.Module):
    def __init__(self, c1, c2, e=0.5):
        super().__init__()
        assert c1 == c2
        self.c = int(c1 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv(2 * self.c, c1, 1)
        self.attn = Attention(self.c, attn_ratio=0.5, num_heads=self.c // 64)
        self.ffn = nn.Sequential(Conv(self.c, self.c * 2, 1), Conv(self.c * 2, self.c, 1, act=False))
        self.emu = None

    def forward(self, x):
        a, b = self.cv1(x).split((self.c, self.c), dim=1)
        b = b + self.attn(b)
        if self.emu == "bf16":
            b = _r16(b)
        b = b + self.ffn(b)
        if self.emu == "bf16":
            b = _r16(b)
        return self.cv2(torch.cat((a, b), 1))


class RepVGGDW(nn.Module):
    def __init__(self, ed):
        super().__init__()
        self.conv = Conv(ed, ed, 7, 1, 3, g=ed, act=False)
        self.conv1 = Conv(ed, ed, 3, 1, 1, g=ed, act=False)
        self.dim = ed
        self.act = nn.SiLU()
        self.emu = None

    def forward(self, x):
        if hasattr(self, "conv1"):
            y = self.act(self.conv(x) + self.conv1(x))
        else:  # fused: the 3x3 (zero-padded to 7x7) has been summed into the 7x7
            y = self.act(self.conv(x))
        return _r16(y) if self.emu == "bf16" else y

    @torch.no_grad()
    def fuse(self):
        """Merge the two (already BN-folded) depthwise branches into one 7x7 depthwise conv."""
        if not hasattr(self, "conv1"):
            return
        w7, b7 = self.conv.conv.weight.data, self.conv.conv.bias.data
        w3, b3 = self.conv1.conv.weight.data, self.conv1.conv.bias.data
        w7 = w7 + F.pad(w3, [2, 2, 2, 2])
        dw = nn.Conv2d(self.dim, self.dim, 7, 1, 3, groups=self.dim, bias=True)
        dw.weight.data.copy_(w7)
        dw.bias.data.copy_(b7 + b3)
        dw.requires_grad_(False)
        fused = Conv(self.dim, self.dim, 7, 1, 3, g=self.dim, act=False)
        del fused.bn
        fused.conv = dw
        # the summed branch output is NOT rounded before the SiLU: keep emu off on the inner conv
        self.conv = fused
        del self.conv1


class CIB(nn.Module):
    def __init__(self, c1, c2, shortcut=True, e=0.5, lk=False):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = nn.Sequential(
            Conv(c1, c1, 3, g=c1),
            Conv(c1, 2 * c_, 1),
            RepVGGDW(2 * c_) if lk else Conv(2 * c_, 2 * c_, 3, g=2 * c_),
            Conv(2 * c_, c2, 1),
            Conv(c2, c2, 3, g=c2),
        )
        self.add = shortcut and c1 == c2
        self.emu = None

    def forward(self, x):
        y = self.cv1(x)
        if self.add:
            y = x + y
            if self.emu == "bf16":
                y = _r16(y)
        return y


class C2fCIB(C2f):
    def __init__(self, c1, c2, n=1, shortcut=False, lk=False, g=1, e=0.5):
        super().__init__(c1, c2, n, shortcut, g, e)
        self.m = nn.ModuleList(CIB(self.c, self.c, shortcut, e=1.0, lk=lk) for _ in range(n))


# ------------------------------------------------------------------------------------------------
# YOLO11 blocks (UPSTREAM block.py::C3k / C3k2 / PSABlock / C2PSA)
# ------------------------------------------------------------------------------------------------
class C3k(nn.Module):
    """C3 with n Bottleneck(c_, c_, k=(k,k), e=1.0): cv3(cat(m(cv1(x)), cv2(x)))."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5, k=3):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, g, k=(k, k), e=1.0) for _ in range(n)))

    def forward(self, x):
        return self.cv3(torch.cat((self.m(self.cv1(x)), self.cv2(x)), 1))


class C3k2(C2f):
    """C2f whose inner blocks are C3k(c, c, 2) (c3k=True) or Bottleneck(c, c) with the default e=0.5."""

    def __init__(self, c1, c2, n=1, c3k=False, e=0.5, g=1, shortcut=True):
        super().__init__(c1, c2, n, shortcut, g, e)
        self.m = nn.ModuleList(
            C3k(self.c, self.c, 2, shortcut, g) if c3k else Bottleneck(self.c, self.c, shortcut, g) for _ in range(n)
        )


class PSABlock(nn.Module):
    def __init__(self, c, attn_ratio=0.5, num_heads=4, shortcut=True):
        super().__init__()
        self.attn = Attention(c, attn_ratio=attn_ratio, num_heads=num_heads)
        self.ffn = nn.Sequential(Conv(c, c * 2, 1), Conv(c * 2, c, 1, act=False))
        self.add = shortcut
        self.emu = None

    def forward(self, x):
        x = x + self.attn(x) if self.add else self.attn(x)
        if self.emu == "bf16":
            x = _r16(x)
        x = x + self.ffn(x) if self.add else self.ffn(x)
        if self.emu == "bf16":
            x = _r16(x)
        return x


class C2PSA(nn.Module):
    def __init__(self, c1, c2, n=1, e=0.5):
        super().__init__()
        assert c1 == c2
        self.c = int(c1 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv(2 * self.c, c1, 1)
        self.m = nn.Sequential(*(PSABlock(self.c, attn_ratio=0.5, num_heads=self.c // 64) for _ in range(n)))

    def forward(self, x):
        a, b = self.cv1(x).split((self.c, self.c), dim=1)
        b = self.m(b)
        return self.cv2(torch.cat((a, b), 1))


# ------------------------------------------------------------------------------------------------
# Heads
# ------------------------------------------------------------------------------------------------
def make_anchors(feats, strides, grid_cell_offset=0.5):
    """Cell-centre anchors, row-major, levels concatenated (UPSTREAM utils/tal.py::make_anchors)."""
    anchor_points, stride_tensor = [], []
    dtype, device = feats[0].dtype, feats[0].device
    for i, stride in enumerate(strides):
        h, w = feats[i].shape[2:]
        sx = torch.arange(end=w, device=device, dtype=dtype) + grid_cell_offset
        sy = torch.arange(end=h, device=device, dtype=dtype) + grid_cell_offset
        sy, sx = torch.meshgrid(sy, sx, indexing="ij")
        anchor_points.append(torch.stack((sx, sy), -1).view(-1, 2))
        stride_tensor.append(torch.full((h * w, 1), stride, dtype=dtype, device=device))
    return torch.cat(anchor_points), torch.cat(stride_tensor)


def dist2bbox(distance, anchor_points, xywh=True, dim=-1):
    """(l,t,r,b) distances -> box (UPSTREAM utils/tal.py::dist2bbox)."""
    lt, rb = distance.chunk(2, dim)
    x1y1 = anchor_points - lt
    x2y2 = anchor_points + rb
    if xywh:
        c_xy = (x1y1 + x2y2) / 2
        wh = x2y2 - x1y1
        return torch.cat((c_xy, wh), dim)
    return torch.cat((x1y1, x2y2), dim)


class Detect(nn.Module):
    """UPSTREAM head.py::Detect (v8 'legacy' class branch unless legacy=False)."""

    end2end = False
    max_det = 300

    def __init__(self, nc=80, ch=(), legacy=True):
        super().__init__()
        self.nc = nc
        self.nl = len(ch)
        self.reg_max = 16
        self.no = nc + self.reg_max * 4
        self.stride = torch.tensor([8.0, 16.0, 32.0])[: self.nl]
        c2, c3 = max((16, ch[0] // 4, self.reg_max * 4)), max(ch[0], min(self.nc, 100))
        self.cv2 = nn.ModuleList(
            nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for x in ch
        )
        if legacy:
            self.cv3 = nn.ModuleList(
                nn.Sequential(Conv(x, c3, 3), Conv(c3, c3, 3), nn.Conv2d(c3, self.nc, 1)) for x in ch
            )
        else:
            self.cv3 = nn.ModuleList(
                nn.Sequential(
                    nn.Sequential(Conv(x, x, 3, g=x), Conv(x, c3, 1)),
                    nn.Sequential(Conv(c3, c3, 3, g=c3), Conv(c3, c3, 1)),
                    nn.Conv2d(c3, self.nc, 1),
                )
                for x in ch
            )
        self.dfl = DFL(self.reg_max)
        if self.end2end:
            self.one2one_cv2 = copy.deepcopy(self.cv2)
            self.one2one_cv3 = copy.deepcopy(self.cv3)

    def bias_init(self):
        """UPSTREAM Detect.bias_init: box bias 1.0, class bias log(5/nc/(640/s)^2)."""
        branches = [(self.cv2, self.cv3)]
        if self.end2end:
            branches.append((self.one2one_cv2, self.one2one_cv3))
        for cv2, cv3 in branches:
            for a, b, s in zip(cv2, cv3, self.stride):
                a[-1].bias.data[:] = 1.0
                b[-1].bias.data[: self.nc] = math.log(5 / self.nc / (640 / float(s)) ** 2)

    def head_maps(self, x, one2one=None):
        """Raw per-level maps cat(box logits 64, class logits nc) -> list of (B, 64+nc, Hi, Wi)."""
        if one2one is None:
            one2one = self.end2end
        cv2, cv3 = (self.one2one_cv2, self.one2one_cv3) if one2one else (self.cv2, self.cv3)
        return [torch.cat((cv2[i](x[i]), cv3[i](x[i])), 1) for i in range(self.nl)]

    def _inference(self, maps):
        """Decode raw maps to (B, 4+nc, A): DFL expectation, dist2bbox, *stride; class sigmoid."""
        shape = maps[0].shape
        x_cat = torch.cat([xi.view(shape[0], self.no, -1) for xi in maps], 2)
        anchors, strides = (t.transpose(0, 1) for t in make_anchors(maps, self.stride.tolist(), 0.5))
        box, cls = x_cat.split((self.reg_max * 4, self.nc), 1)
        dbox = dist2bbox(self.dfl(box), anchors.unsqueeze(0), xywh=not self.end2end, dim=1) * strides
        return torch.cat((dbox, cls.sigmoid()), 1)

    def forward(self, x):
        maps = self.head_maps(list(x))
        y = self._inference(maps)
        if self.end2end:
            return self.postprocess(y.permute(0, 2, 1), self.max_det, self.nc), maps
        return y, maps

    @staticmethod
    def postprocess(preds, max_det, nc=80):
        """Two-stage top-k of the one-to-one head (UPSTREAM Detect.postprocess)."""
        batch_size, anchors, _ = preds.shape
        boxes, scores = preds.split([4, nc], dim=-1)
        index = scores.amax(dim=-1).topk(min(max_det, anchors))[1].unsqueeze(-1)
        boxes = boxes.gather(dim=1, index=index.repeat(1, 1, 4))
        scores = scores.gather(dim=1, index=index.repeat(1, 1, nc))
        scores, index = scores.flatten(1).topk(min(max_det, anchors))
        i = torch.arange(batch_size)[..., None]
        return torch.cat([boxes[i, index // nc], scores[..., None], (index % nc)[..., None].float()], dim=-1)


class Segment(Detect):
    """UPSTREAM head.py::Segment."""

    def __init__(self, nc=80, nm=32, npr=256, ch=(), legacy=True):
        super().__init__(nc, ch, legacy)
        self.nm = nm
        self.npr = npr
        self.proto = Proto(ch[0], self.npr, self.nm)
        c4 = max(ch[0] // 4, self.nm)
        self.cv4 = nn.ModuleList(nn.Sequential(Conv(x, c4, 3), Conv(c4, c4, 3), nn.Conv2d(c4, self.nm, 1)) for x in ch)

    def forward(self, x):
        x = list(x)
        p = self.proto(x[0])
        bs = p.shape[0]
        mc = torch.cat([self.cv4[i](x[i]).view(bs, self.nm, -1) for i in range(self.nl)], 2)
        y, maps = Detect.forward(self, x)
        return torch.cat([y, mc], 1), (maps, mc, p)


class v10Detect(Detect):
    """UPSTREAM head.py::v10Detect: NMS-free one-to-one head, depthwise class branch."""

    end2end = True

    def __init__(self, nc=80, ch=()):
        super().__init__(nc, ch, legacy=False)
