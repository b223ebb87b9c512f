"""GPU bring-up: layer-by-layer comparison of the native engine against the bf16-emulating oracle,
then head decode / NMS / masks against the oracle's post-processing fed with the ENGINE's tensors.
Usage: python tools/bringup_engine.py [model] [impl...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ops as oops  # noqa: E402
from oracle.model import build_model  # noqa: E402
from yolo_puncture_b200 import synth  # noqa: E402
from yolo_puncture_b200.engine import Engine  # noqa: E402
from yolo_puncture_b200.model import box_xform  # noqa: E402


def stats(name, got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    denom = ref.abs().mean().item() + 1e-9
    bad = (err > 0.05 + 0.03 * ref.abs()).float().mean().item()
    print(f"  {name:28s} shape {tuple(ref.shape)} mean|ref| {denom:.4f} mean|err| {err.mean().item():.5f} "
          f"max|err| {err.max().item():.4f} frac_bad {bad:.5f} nan {int(torch.isnan(got).sum())}")
    return bad


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "yolov8n-seg"
    impls = [int(a) for a in sys.argv[2:]] or [1, 0]
    B, H, W = 2, 640, 640
    torch.set_num_threads(os.cpu_count())
    net = build_model(name)
    sd = synth.synth_state_dict([(k, v.shape) for k, v in net.state_dict().items()], name)
    net.load_state_dict(sd)
    net.fuse().set_emulation(True)
    frames = synth.synth_frames(B, H, W)
    im = oops.preprocess(frames, 640)
    with torch.no_grad():
        feats = net.features(im)
    pred_o, (maps_o, mc_o, proto_o) = feats[-1]

    eng = Engine(name)
    eng.load_state_dict(sd)
    eng.finalize(0)
    eng.plan(B, H, W)
    fr = torch.from_numpy(np.stack(frames)).cuda()
    xf = torch.tensor([box_xform((H, W), (H, W))] * B, dtype=torch.float32).cuda()
    for impl in impls:
        print(f"=== {name} conv impl {impl} ({'tcgen05' if impl == 0 else 'simt twin'}) ===")
        eng.set_conv_impl(impl)
        eng.ws.zero_()
        eng.infer(fr, xf, conf=0.25, iou=0.7)
        torch.cuda.synchronize()
        print("  device error word:", eng.device_error())
        for vname in eng.view_table():
            if vname.startswith("model."):
                li = int(vname.split(".")[1])
                ref = feats[li]
                if not torch.is_tensor(ref):
                    continue
                stats(vname, eng.view(vname), ref.permute(0, 2, 3, 1))
        stats("proto", eng.view("proto"), proto_o.permute(0, 2, 3, 1))
        head = eng.view("head")[:, 0]  # (B, A, no)
        raw_o = torch.cat([m.flatten(2) for m in maps_o], 2).permute(0, 2, 1)  # (B, A, 64+nc)
        stats("head.box_logits", head[..., :64], raw_o[..., :64])
        stats("head.cls_logits", head[..., 64:144], raw_o[..., 64:])
        stats("head.coefs", head[..., 144:], mc_o.permute(0, 2, 1))
        # ---- selection on the ENGINE's head tensor, oracle post-processing as the checker ----
        hd = head.float().cpu()
        seg = net.model[-1]
        maps_e, off = [], 0
        for m in maps_o:
            n = m.shape[2] * m.shape[3]
            maps_e.append(hd[:, off:off + n, :144].permute(0, 2, 1).reshape(B, 144, m.shape[2], m.shape[3]))
            off += n
        with torch.no_grad():
            y = seg._inference(maps_e)
        pred_e = torch.cat([y, hd[..., 144:].permute(0, 2, 1)], 1)
        dets, kept = oops.non_max_suppression(pred_e, 0.25, 0.7, max_det=300, nc=80, return_idx=True)
        cnt = eng.count.cpu().tolist()
        print("  counts engine", cnt, "oracle", [len(d) for d in dets])
        proto_e = eng.view("proto").float().cpu().permute(0, 3, 1, 2)
        tot = sum(cnt)
        masks = torch.zeros((max(tot, 1), H, W), dtype=torch.uint8, device="cuda")
        eng.masks(masks, True, H, W)
        torch.cuda.synchronize()
        print("  mask status", eng.mask_status.cpu().tolist(), "device error word:", eng.device_error())
        o = 0
        for b in range(B):
            n = min(cnt[b], len(dets[b]))
            ke = eng.keep[b, :cnt[b]].cpu().long()
            same_keep = cnt[b] == len(dets[b]) and bool((ke == kept[b]).all())
            de = eng.det[b, :n].cpu()
            do = dets[b][:n].clone()
            do[:, :4] = oops.scale_boxes((H, W), do[:, :4], (H, W))
            berr = (de[:, :4] - do[:, :4]).abs().max().item() if n else 0.0
            serr = (de[:, 4] - do[:, 4]).abs().max().item() if n else 0.0
            cls_same = bool((de[:, 5] == do[:, 5]).all()) if n else True
            print(f"  img {b}: n {cnt[b]} keep identical {same_keep} box max err {berr:.4g} score err {serr:.3g} cls same {cls_same}")
            if n:
                mo = oops.process_mask_native(proto_e[b], do[:, 6:], do[:, :4], (H, W))
                me = masks[o:o + n].cpu().float()
                inter = (mo * me).sum((1, 2))
                union = ((mo + me) > 0).float().sum((1, 2))
                iou = torch.where(union > 0, inter / union.clamp(min=1), torch.ones_like(union))
                print(f"         mask IoU min {iou.min().item():.5f} mean {iou.mean().item():.5f} "
                      f"mismatch px {(mo != me).sum().item()} of {mo.numel()}")
            o += cnt[b]


if __name__ == "__main__":
    main()
