"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/ypb200.h declares, the native graph builder's weight table equals the oracle's state_dict
(names and shapes), planning reproduces the survey's FLOP totals, and the product
path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle.model import build_model
from yolo_puncture_b200 import YOLO, _lib
from yolo_puncture_b200.engine import Engine, YpbError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEG = [f"yolov8{s}-seg" for s in "nsmlx"] + ["yolov10n"] + [f"yolo11{s}-seg" for s in "nsmlx"]


def _declared(header_name):
    header = open(os.path.join(ROOT, "include", header_name)).read()
    return set(re.findall(r"\b(ypb_[a-z0-9_]+)\s*\(", header))


def test_library_exports_every_declared_symbol():
    declared = _declared("ypb200.h")
    assert len(declared) >= 20
    handle = ctypes.CDLL(_lib._build.ensure_built())
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in ypb200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes prototypes out of sync with the header"
    assert _lib.lib().ypb_version() == 100


def test_product_library_carries_no_diagnostics():
    """Debugging twins and micro-benchmarks live in libypb200_diag.so (include/ypb200_diag.h), not in the product."""
    diag = _declared("ypb200_diag.h") - _declared("ypb200.h")
    assert diag == set(_lib.DIAG_SIGNATURES) and len(diag) >= 5
    handle = ctypes.CDLL(_lib._build.ensure_built())
    if os.environ.get("YPB_LIB"):
        pytest.skip("YPB_LIB override in effect")
    assert handle.ypb_is_diag_build() == 0
    for name in diag:
        assert not hasattr(handle, name), f"{name} is a diagnostic but the product library exports it"
    import subprocess
    syms = subprocess.run(["cuobjdump", "-elf", _lib._build.LIB], capture_output=True, text=True).stdout
    for twin in ("conv_simt_kernel", "conv_halo_test_kernel", "tma_bench_kernel", "mma_bench_kernel", "latency_probe_kernel"):
        assert twin not in syms, f"{twin} compiled into the product library"
    e = Engine("yolov8n-seg")
    with pytest.raises(YpbError, match="libypb200_diag"):
        e.set_conv_impl(2)


def test_diag_library_exports_its_header():
    handle = ctypes.CDLL(_lib._build.build(diag=True))
    assert handle.ypb_is_diag_build() == 1
    for name in _declared("ypb200_diag.h") | _declared("ypb200.h"):
        assert hasattr(handle, name), f"{name} missing from libypb200_diag.so"


@pytest.mark.parametrize("name", SEG)
def test_weight_table_matches_upstream_state_dict(name):
    specs = {n: s for n, s, _ in Engine(name).weight_specs()}
    sd = {k: tuple(v.shape) for k, v in build_model(name).state_dict().items() if not k.endswith("num_batches_tracked")}
    assert specs == sd


@pytest.mark.parametrize("name,gflop", [("yolov8n-seg", 12.00), ("yolov8s-seg", 40.09), ("yolov8m-seg", 104.54),
                                        ("yolov8x-seg", 328.36)])
def test_plan_reproduces_survey_flops(name, gflop):
    e = Engine(name)
    assert e.plan_only(1, 640, 640) > 0
    assert abs(_lib.lib().ypb_conv_flops(e._h) / 1e9 - gflop) < 0.01
    assert _lib.lib().ypb_num_anchors(e._h) == 8400


def test_plan_rect_1080p():
    e = Engine("yolov8m-seg")
    e.plan_only(1, 736, 1280)
    assert _lib.lib().ypb_num_anchors(e._h) == 19320
    assert abs(_lib.lib().ypb_conv_flops(e._h) / 1e9 - 240.44) < 0.01


def test_error_behaviour():
    with pytest.raises(YpbError, match="unknown model spec"):
        Engine("yolov99-seg")
    e = Engine("yolov8n-seg")
    with pytest.raises(YpbError, match="multiples of 32"):
        e.plan_only(1, 650, 640)
    with pytest.raises(YpbError, match="unknown weight"):
        e.load_state_dict({"model.0.conv.bogus": np.zeros(3, np.float32)}, strict=False) or \
            _lib.check(_lib.lib().ypb_load_weight(e._h, b"nope", ctypes.c_void_p(0), 0) if False else
                       _lib.lib().ypb_load_weight(e._h, b"nope", np.zeros(1, np.float32).ctypes.data_as(ctypes.c_void_p), 1))
    import torch
    with pytest.raises(YpbError, match="shape mismatch"):
        e.load_state_dict({"model.0.conv.weight": torch.zeros(1, 3, 3, 3)}, strict=False)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = YOLO("yolov8n-seg")
    with pytest.raises(YpbError, match="no CPU fallback|only runs on CUDA"):
        m.predict(np.zeros((64, 64, 3), np.uint8))
    with pytest.raises(YpbError):
        m.predict(np.zeros((64, 64, 3), np.uint8), device="cpu")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "yolo_puncture_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{fn} imports the oracle"
            assert "from .. import oracle" not in src
