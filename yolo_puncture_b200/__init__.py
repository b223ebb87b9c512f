"""yolo_puncture_b200 — B200-native (sm_100a) per-frame YOLO-seg / YOLOv10 detector.

Drop-in for the `ultralytics` predict path that daisy9542/yolo-puncture runs on every video frame
(reference yolo_seg/app.py:45,91; yolo_seg/yolo_with_deva.py:51,226):

    from yolo_puncture_b200 import YOLO          # or: install_ultralytics_shim(); from ultralytics import YOLO
    results = YOLO("yolov8s-seg").predict(frame, conf=0.9, retina_masks=True)

Hand-written CUDA (tcgen05/TMEM/TMA implicit-GEMM convs, warp-level decode/NMS, fused mask decode)
behind the C ABI of include/ypb200.h; PyTorch only moves bytes.  No CPU fallback.
"""

from .frames import decode_jpegs  # noqa: F401
from .handoff import auto_segment, index_masks, min_rect_len  # noqa: F401
from .model import YOLO  # noqa: F401
from .results import Boxes, Masks, Results  # noqa: F401

__all__ = ["YOLO", "Results", "Boxes", "Masks", "install_ultralytics_shim", "index_masks", "min_rect_len", "auto_segment", "decode_jpegs"]


def install_ultralytics_shim():
    """Make `from ultralytics import YOLO` resolve to this package (the reference imports it that way:
    yolo_seg/app.py:7, yolo_seg/yolo_with_deva.py:12, dev_tools/auto_speed_calc.py:10)."""
    import sys
    import types

    mod = types.ModuleType("ultralytics")
    mod.YOLO = YOLO
    mod.__version__ = "8.3.0+ypb200"
    sys.modules["ultralytics"] = mod
    return mod
