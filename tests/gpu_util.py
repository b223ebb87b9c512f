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
