"""Which phase of nms_kernel the slowest image of the bench batch spends its time in (profiling build).
    python -m yolo_puncture_b200.build --prof && python tools/nms_probe.py"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["YPB_LIB"] = os.path.join(ROOT, "yolo_puncture_b200", "libypb200_prof.so")
from yolo_puncture_b200 import YOLO, synth
from yolo_puncture_b200._lib import check, lib, diag_lib
from yolo_puncture_b200.model import box_xform

yolo = YOLO("yolov8s-seg", device=0)
eng = yolo.engine
eng.set_graph(False)
B = 64
fr = torch.from_numpy(np.stack([synth.synth_frame(i) for i in range(B)])).cuda()
xf = torch.tensor([box_xform((640, 640), (640, 640))] * B, dtype=torch.float32, device="cuda")
eng.plan(B, 640, 640)
z = (C.c_ulonglong * 16)()
for it in range(3):
    check(diag_lib().ypb_debug_prof(z, 1))
    eng.infer(fr, xf, 0.25, 0.7)
    check(diag_lib().ypb_debug_prof(z, 1))
    print(f"slowest image: candidates {z[0]} kept {z[3]} sort {z[1]} cycles, greedy+rest {z[2]} cycles path {z[4]}")
cc = eng.candidate_counts().cpu().numpy()
print("candidates per image: max", cc.max(), "top5", np.sort(cc)[-5:], "kept", eng.count.cpu().numpy()[np.argsort(cc)[-5:]])
