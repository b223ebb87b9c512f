"""Per-layer conv microbenchmark: every distinct conv shape of yolov8s-seg at B=64, 640x640, timed in isolation
(planned once, `iters` back-to-back launches between CUDA events) under the YPB_DBG experiment masks.

    python tools/conv_layers.py [--dbg 0,7] [--only substr] [--iters 20] [--impl 0] [--csv out.csv]

dbg bits: 1 = no bias/SiLU math, 2 = no output stores, 4 = no MMA issue (7 = operand fetch + hand-offs only).
Columns: ms, TFLOP/s, GB/s of algorithmic HBM traffic (input slice + output slice + weights).
"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if "--prof" in sys.argv:  # wait-cycle accounting lives in the profiling build: python -m yolo_puncture_b200.build --prof
    sys.argv.remove("--prof")
    os.environ["YPB_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "yolo_puncture_b200",
                                         "libypb200_prof.so")
from yolo_puncture_b200._lib import check, diag_lib, lib  # noqa: E402

B = 64
# name, Hin, Win, in_ctot, in_c_off, cin, cout, k, stride, res, out_mode(0 bf16,1 f32,2 shuffle), out_ctot, act
LAYERS = [
    ("model.1       3x3s2  32->64  @320", 320, 320, 32, 0, 32, 64, 3, 2, 0, 0, 64, 1),
    ("model.2.cv1   1x1    64->64  @160", 160, 160, 64, 0, 64, 64, 1, 1, 0, 0, 96, 1),
    ("model.2.m.cv1 3x3    32->32  @160", 160, 160, 96, 32, 32, 32, 3, 1, 0, 0, 32, 1),
    ("model.2.m.cv2 3x3    32->32+r@160", 160, 160, 32, 0, 32, 32, 3, 1, 1, 0, 96, 1),
    ("model.2.cv2   1x1    96->64  @160", 160, 160, 96, 0, 96, 64, 1, 1, 0, 0, 64, 1),
    ("model.3       3x3s2  64->128 @160", 160, 160, 64, 0, 64, 128, 3, 2, 0, 0, 128, 1),
    ("model.4.cv1   1x1   128->128 @80", 80, 80, 128, 0, 128, 128, 1, 1, 0, 0, 256, 1),
    ("model.4.m.cv1 3x3    64->64  @80", 80, 80, 256, 64, 64, 64, 3, 1, 0, 0, 64, 1),
    ("model.4.m.cv2 3x3    64->64+r@80", 80, 80, 64, 0, 64, 64, 3, 1, 1, 0, 256, 1),
    ("model.4.cv2   1x1   256->128 @80", 80, 80, 256, 0, 256, 128, 1, 1, 0, 0, 384, 1),
    ("model.5       3x3s2 128->256 @80", 80, 80, 384, 256, 128, 256, 3, 2, 0, 0, 256, 1),
    ("model.6.cv1   1x1   256->256 @40", 40, 40, 256, 0, 256, 256, 1, 1, 0, 0, 512, 1),
    ("model.6.m.cv1 3x3   128->128 @40", 40, 40, 512, 128, 128, 128, 3, 1, 0, 0, 128, 1),
    ("model.6.m.cv2 3x3   128->128+r@40", 40, 40, 128, 0, 128, 128, 3, 1, 1, 0, 512, 1),
    ("model.6.cv2   1x1   512->256 @40", 40, 40, 512, 0, 512, 256, 1, 1, 0, 0, 768, 1),
    ("model.7       3x3s2 256->512 @40", 40, 40, 768, 512, 256, 512, 3, 2, 0, 0, 512, 1),
    ("model.8.cv1   1x1   512->512 @20", 20, 20, 512, 0, 512, 512, 1, 1, 0, 0, 768, 1),
    ("model.8.m.cv1 3x3   256->256 @20", 20, 20, 768, 256, 256, 256, 3, 1, 0, 0, 256, 1),
    ("model.8.cv2   1x1   768->512 @20", 20, 20, 768, 0, 768, 512, 1, 1, 0, 0, 512, 1),
    ("model.9.cv1   1x1   512->256 @20", 20, 20, 512, 0, 512, 256, 1, 1, 0, 0, 1024, 1),
    ("model.9.cv2   1x1  1024->512 @20", 20, 20, 1024, 0, 1024, 512, 1, 1, 0, 0, 1536, 1),
    ("model.12.cv1  1x1   768->256 @40", 40, 40, 768, 0, 768, 256, 1, 1, 0, 0, 384, 1),
    ("model.15.cv1  1x1   384->128 @80", 80, 80, 384, 0, 384, 128, 1, 1, 0, 0, 192, 1),
    ("model.15.cv2  1x1   192->128 @80", 80, 80, 192, 0, 192, 128, 1, 1, 0, 0, 128, 1),
    ("model.16      3x3s2 128->128 @80", 80, 80, 128, 0, 128, 128, 3, 2, 0, 0, 384, 1),
    ("model.19      3x3s2 256->256 @40", 40, 40, 256, 0, 256, 256, 3, 2, 0, 0, 768, 1),
    ("head0.s0      3x3   128->224 @80", 80, 80, 128, 0, 128, 224, 3, 1, 0, 0, 224, 1),
    ("head0.cv3.1   3x3   128->128 @80", 80, 80, 224, 64, 128, 128, 3, 1, 0, 0, 128, 1),
    ("head0.cv4.1   3x3    32->32  @80", 80, 80, 224, 192, 32, 32, 3, 1, 0, 0, 32, 1),
    ("head0.cv2.2   1x1    64->64 f32@80", 80, 80, 64, 0, 64, 64, 1, 1, 0, 1, 176, 0),
    ("head0.cv3.2   1x1   128->80 f32@80", 80, 80, 128, 0, 128, 80, 1, 1, 0, 1, 176, 0),
    ("head1.s0      3x3   256->224 @40", 40, 40, 256, 0, 256, 224, 3, 1, 0, 0, 224, 1),
    ("head2.s0      3x3   512->224 @20", 20, 20, 512, 0, 512, 224, 3, 1, 0, 0, 224, 1),
    ("head2.cv4.2   1x1    32->32 f32@20", 20, 20, 32, 0, 32, 32, 1, 1, 0, 1, 176, 0),
    ("proto.cv1     3x3   128->128 @80", 80, 80, 128, 0, 128, 128, 3, 1, 0, 0, 128, 1),
    ("proto.up      1x1T  128->512 @80", 80, 80, 128, 0, 128, 512, 1, 1, 0, 2, 128, 0),
    ("proto.cv2     3x3   128->128 @160", 160, 160, 128, 0, 128, 128, 3, 1, 0, 0, 128, 1),
    ("proto.cv3     1x1   128->32 f32@160", 160, 160, 128, 0, 128, 32, 1, 1, 0, 1, 32, 1),
]


# yolov8x-seg at B=32 (BASELINE config C5's model): Cin 80 / 160 / 320 / 640 -> ragged 64-channel chunks, streamed weights
LAYERS_X = [
    ("x.1           3x3s2  80->160 @320", 320, 320, 80, 0, 80, 160, 3, 2, 0, 0, 160, 1),
    ("x.2.m.cv1     3x3    80->80  @160", 160, 160, 400, 80, 80, 80, 3, 1, 0, 0, 80, 1),
    ("x.2.m.cv2     3x3    80->80+r@160", 160, 160, 80, 0, 80, 80, 3, 1, 1, 0, 400, 1),
    ("x.3           3x3s2 160->320 @160", 160, 160, 160, 0, 160, 320, 3, 2, 0, 0, 320, 1),
    ("x.4.m.cv1     3x3   160->160 @80", 80, 80, 1280, 160, 160, 160, 3, 1, 0, 0, 160, 1),
    ("x.4.m.cv2     3x3   160->160+r@80", 80, 80, 160, 0, 160, 160, 3, 1, 1, 0, 1280, 1),
    ("x.4.cv2       1x1  1280->320 @80", 80, 80, 1280, 0, 1280, 320, 1, 1, 0, 0, 960, 1),
    ("x.6.m.cv1     3x3   320->320 @40", 40, 40, 2560, 320, 320, 320, 3, 1, 0, 0, 320, 1),
    ("x.6.m.cv2     3x3   320->320+r@40", 40, 40, 320, 0, 320, 320, 3, 1, 1, 0, 2560, 1),
    ("x.8.m.cv1     3x3   320->320 @20", 20, 20, 1600, 320, 320, 320, 3, 1, 0, 0, 320, 1),
    ("x.head0.s0    3x3   320->480 @80", 80, 80, 320, 0, 320, 480, 3, 1, 0, 0, 480, 1),
    ("x.head0.cv3.1 3x3   320->320 @80", 80, 80, 480, 80, 320, 320, 3, 1, 0, 0, 320, 1),
    ("x.proto.cv2   3x3   320->320 @160", 160, 160, 320, 0, 320, 320, 3, 1, 0, 0, 320, 1),
]
# yolov8m-seg at 736x1280, B=16 (config C4): c = 48 / 96 / 192 / 288
LAYERS_M = [
    ("m.2.m.cv1     3x3    48->48  @184x320", 184, 320, 192, 48, 48, 48, 3, 1, 0, 0, 48, 1),
    ("m.4.m.cv1     3x3    96->96  @92x160", 92, 160, 576, 96, 96, 96, 3, 1, 0, 0, 96, 1),
    ("m.6.m.cv1     3x3   192->192 @46x80", 46, 80, 1152, 192, 192, 192, 3, 1, 0, 0, 192, 1),
    ("m.8.m.cv1     3x3   288->288 @23x40", 23, 40, 1152, 288, 288, 288, 3, 1, 0, 0, 288, 1),
    ("m.head0.s0    3x3   192->320 @92x160", 92, 160, 192, 0, 192, 320, 3, 1, 0, 0, 320, 1),
    ("m.proto.cv2   3x3   192->192 @184x320", 184, 320, 192, 0, 192, 192, 3, 1, 0, 0, 192, 1),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="s", choices=["s", "x", "m"], help="layer table: yolov8s-seg B=64, yolov8x-seg B=32, yolov8m-seg 736x1280 B=16")
    ap.add_argument("--dbg", default="0,1,2,4,7")
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--impl", type=int, default=0)
    ap.add_argument("--batch", type=int, default=B)
    ap.add_argument("--csv", default=None)
    a = ap.parse_args()
    layers = {"s": LAYERS, "x": LAYERS_X, "m": LAYERS_M}[a.set]
    if a.set != "s" and a.batch == B:
        a.batch = 32 if a.set == "x" else 16
    dbgs = [int(x) for x in a.dbg.split(",")]
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    print(f"{'layer':36s} " + " ".join(f"{'dbg' + str(d) + ' ms':>10s}" for d in dbgs) + "   TFLOP/s    GB/s  plan")
    for (name, H, W, ictot, ioff, cin, cout, k, s, res, omode, octot, act) in layers:
        if a.only and a.only not in name:
            continue
        nb = a.batch
        x = (torch.randn(nb, H, W, ictot, device=dev) * 0.5).to(torch.bfloat16)
        w = (torch.randn(k * k, cout, cin, device=dev) * (1.0 / (cin * k * k) ** 0.5)).to(torch.bfloat16)
        bias = torch.randn(cout, device=dev) * 0.1
        oH, oW = H // s, W // s
        if omode == 2:
            out = torch.empty(nb, 2 * oH, 2 * oW, octot, device=dev, dtype=torch.bfloat16)
        elif omode == 1:
            out = torch.empty(nb, oH, oW, octot, device=dev, dtype=torch.float32)
        else:
            out = torch.empty(nb, oH, oW, octot, device=dev, dtype=torch.bfloat16)
        r = (torch.randn(nb, oH, oW, octot, device=dev) * 0.5).to(torch.bfloat16) if res else None
        flops = 2.0 * nb * oH * oW * cout * cin * k * k
        oe = 4 if omode == 1 else 2
        byts = nb * H * W * cin * 2 + nb * oH * oW * cout * oe + w.numel() * 2 + (nb * oH * oW * cout * 2 if res else 0)
        ms_l, desc = [], C.create_string_buffer(640)
        for d in dbgs:
            ms = C.c_float()
            check(diag_lib().ypb_conv_bench(C.c_void_p(st), C.c_void_p(x.data_ptr()), nb, H, W, ictot, ioff, cin,
                                       C.c_void_p(w.data_ptr()), C.c_void_p(bias.data_ptr()), cout, k, s, act,
                                       C.c_void_p(r.data_ptr()) if res else None, C.c_void_p(out.data_ptr()), octot, 0, omode,
                                       a.impl, d, a.iters, C.byref(ms), desc, 640))
            ms_l.append(ms.value)
        t0 = ms_l[0] * 1e-3
        print(f"{name:36s} " + " ".join(f"{m:10.4f}" for m in ms_l) + f" {flops / t0 / 1e12:9.1f} {byts / t0 / 1e9:7.0f}  {desc.value.decode()}")
        rows.append((name, ms_l, flops, byts, desc.value.decode()))
        del x, w, out, r
    print("total " + " ".join(f"{sum(r[1][i] for r in rows):10.4f}" for i in range(len(dbgs))))
    if a.csv:
        with open(a.csv, "w") as f:
            f.write("layer," + ",".join(f"ms_dbg{d}" for d in dbgs) + ",gflop,mbytes,plan\n")
            for name, ms_l, flops, byts, desc in rows:
                f.write(name.replace(",", ";") + "," + ",".join(f"{m:.5f}" for m in ms_l) + f",{flops / 1e9:.3f},{byts / 1e6:.2f},{desc}\n")


if __name__ == "__main__":
    main()
