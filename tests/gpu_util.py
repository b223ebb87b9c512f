"""Helpers shared by the -m gpu parity tests."""
import torch
import torch.nn.functional as F


def conv_reference(x_nhwc, w, bias, k, stride, act, res=None):
    """fp32 torch reference of the fused conv on bf16-rounded operands.
    x_nhwc: (B,H,W,Cin) bf16; w: (cout,cin,k,k) fp32 (will be bf16-rounded); returns fp32 NHWC."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x = x_nhwc.float().permute(0, 3, 1, 2).contiguous()
    wq = w.to(torch.bfloat16).float()
    y = F.conv2d(x, wq, bias.float(), stride=stride, padding=k // 2)
    if act:
        y = F.silu(y)
    y = y.permute(0, 2, 3, 1).contiguous()
    if res is not None:
        y = y.to(torch.bfloat16).float() + res.float()
    return y


def describe_mismatch(got, ref, rtol, atol, max_items=12):
    """Human-readable summary of where `got` deviates from `ref` (both (B,H,W,C) float tensors)."""
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    bad = err > (atol + rtol * ref.abs())
    nbad = int(bad.sum())
    lines = [f"mismatch {nbad}/{bad.numel()} ({100.0 * nbad / bad.numel():.3f}%), max abs err {float(err.max()):.4g}, "
             f"ref absmax {float(ref.abs().max()):.4g}, got absmax {float(got.abs().max()):.4g}, "
             f"nan {int(torch.isnan(got).sum())}"]
    if nbad:
        B, H, W, C = bad.shape
        lines.append("bad per image: " + str(bad.sum((1, 2, 3)).tolist()))
        lines.append("bad per row h (first 24): " + str(bad.sum((0, 2, 3)).tolist()[:24]))
        lines.append("bad per col w (first 24): " + str(bad.sum((0, 1, 3)).tolist()[:24]))
        lines.append("bad per channel (first 32): " + str(bad.sum((0, 1, 2)).tolist()[:32]))
        idx = bad.nonzero()[:max_items]
        for b, h, w, c in idx.tolist():
            lines.append(f"  [{b},{h},{w},{c}] got {float(got[b, h, w, c]):.5g} ref {float(ref[b, h, w, c]):.5g}")
    return "\n".join(lines)


def oracle_with_synth(name, emulate, recipe="default", geometry=None):
    """(oracle net fused [+bf16 emulation], state_dict) with the deterministic synthetic weights."""
    from oracle.model import build_model
    from yolo_puncture_b200 import synth
    net = build_model(name)
    sd = synth.synth_state_dict([(k, v.shape) for k, v in net.state_dict().items()], name, recipe=recipe, geometry=geometry)
    net.load_state_dict(sd)
    net.fuse()
    if emulate:
        net.set_emulation(True)
    return net, sd


def oracle_select_on_engine_tensors(eng, net, B, level_shapes, conf, iou, max_det=300, classes=None, agnostic=False):
    """Run the ORACLE's decode + NMS on the engine's own fp32 head rows.  Returns (dets, kept_idx, proto)."""
    from oracle import ops as oops
    nc = net.nc
    head = eng.view("head")[:, 0].float().cpu()
    maps, off = [], 0
    for (h, w) in level_shapes:
        maps.append(head[:, off:off + h * w, :64 + nc].permute(0, 2, 1).reshape(B, 64 + nc, h, w))
        off += h * w
    with torch.no_grad():
        pred = torch.cat([net.model[-1]._inference(maps), head[..., 64 + nc:].permute(0, 2, 1)], 1)
    dets, kept = oops.non_max_suppression(pred, conf, iou, classes=classes, agnostic=agnostic, max_det=max_det, nc=nc,
                                          return_idx=True)
    proto = eng.view("proto").float().cpu().permute(0, 3, 1, 2)
    return dets, kept, proto


def mask_iou(a, b):
    a, b = a.float(), b.float()
    inter, union = (a * b).sum((1, 2)), ((a + b) > 0).float().sum((1, 2))
    return torch.where(union > 0, inter / union.clamp(min=1), torch.ones_like(union))  # two empty masks agree


def box_iou_matrix(a, b):
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:4], b[None, :, 2:4])
    inter = (rb - lt).clamp(min=0).prod(2)
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (aa[:, None] + ab[None] - inter).clamp(min=1e-9)


def drift_stats(ref, got):
    """Match detections of `got` to the reference `ref` (lists of Results) by class and IoU > 0.9.
    Returns (match rate, per-match max box-coordinate error in px, per-match mask IoU, reference detections)."""
    tot, matched, errs, ious = 0, 0, [], []
    for r, g in zip(ref, got):
        if len(r) == 0:
            continue
        rb = torch.as_tensor(r.boxes.data).cpu()
        gb = torch.as_tensor(g.boxes.data).cpu()
        tot += len(rb)
        if len(gb) == 0:
            continue
        m = box_iou_matrix(rb[:, :4], gb[:, :4]) * (rb[:, 5:6] == gb[None, :, 5]).float()
        best, j = m.max(1)
        ok = best > 0.9
        matched += int(ok.sum())
        errs += (rb[ok, :4] - gb[j[ok], :4]).abs().max(1).values.tolist()
        if r.masks is not None and g.masks is not None:
            ious += mask_iou(torch.as_tensor(r.masks.data).cpu()[ok], torch.as_tensor(g.masks.data).cpu()[j[ok]]).tolist()
    return matched / max(tot, 1), errs, ious, tot


def ulp_report(got, ref):
    """Elementwise |got - ref| in units of the bf16 ulp of the reference value: (max, p99.9, mean, share > 1 ulp)."""
    got, ref = got.float().flatten(), ref.float().flatten()
    ulp = torch.exp2(torch.floor(torch.log2(ref.abs().clamp(min=2.0 ** -20))) - 7)
    e = (got - ref).abs() / ulp
    k = max(1, int(e.numel() * 0.999))
    return float(e.max()), float(e.kthvalue(k).values), float(e.mean()), float((e > 1).float().mean())
