"""CPU oracle for the per-frame YOLO detector hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on PyTorch-CPU fp32, the third-party `ultralytics` predict path that
daisy9542/yolo-puncture calls on every video frame (reference call sites:
yolo_seg/app.py:45,49-50,91-101; yolo_seg/yolo_with_deva.py:42,51-83,226;
dev_tools/auto_speed_calc.py:40,62-71; dev_tools/classify/cls_bbox_dataset_generate.py:48-52,66).

The arithmetic of that path lives in `ultralytics` (declared *unpinned* at reference
pyproject.toml:23, must be >= 8.3.0 because yolo_seg/app.py:219-223 loads YOLO11 checkpoints).
The package source is not under /root/reference and is not installable here (no network), and the
reference ships no tests, fixtures, weights or golden vectors (reference .gitignore:144,148,165).

    ==>  PARITY UNPINNED by the reference.  <==

What pins this restatement instead (tests/test_oracle_*.py):
  * exact parameter totals of the upstream model zoo (yolov8{n,s,m,l,x}-seg, yolov10n),
  * conv FLOPs of yolov10n's one-to-one path = 6.70 G (reference README.md:48 "6.7G"),
  * output shapes, NMS / top-k / letterbox / scale_boxes / crop_mask known-answer micro-cases,
  * torchvision.ops.nms as the arithmetic backend of NMS so library semantics are inherited,
  * hand-derived known-answer vectors for DFL + dist2bbox, process_mask_native / process_mask and the class-offset
    NMS (tests/test_oracle_known_answers.py: expected values worked out on paper, not produced by this package),
  * a cross-check against a real `ultralytics` install wherever one is importable
    (tests/test_oracle_vs_ultralytics.py, pytest.importorskip; skipped in the build container).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package, and only as the checker or the timed CPU baseline.  The product path
(yolo_puncture_b200/) never imports it and has no CPU fallback.
"""

from .model import build_model, count_parameters, conv_flops, MODEL_SPECS  # noqa: F401
from .predict import OracleYOLO  # noqa: F401
