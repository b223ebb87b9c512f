"""Known-answer checks that pin the oracle's model restatement:
published parameter totals, published FLOPs, output shapes, BN-fold equivalence."""
import pytest
import torch

from oracle.model import build_model, conv_flops, count_parameters

PARAMS = {"yolov8n-seg": 3409968, "yolov8s-seg": 11821056, "yolov8m-seg": 27285968,
          "yolov8l-seg": 45997728, "yolov8x-seg": 71827888, "yolov10n": 2775520,
          # YOLO11-seg: upstream's model summaries / model-zoo table give 2,876,848 parameters for
          # yolo11n-seg and 10.1 / 22.4 / 27.6 / 62.1 M for s / m / l / x
          "yolo11n-seg": 2876848, "yolo11s-seg": 10113248, "yolo11m-seg": 22420896, "yolo11l-seg": 27678368,
          "yolo11x-seg": 62142656}


@pytest.mark.parametrize("name,total", sorted(PARAMS.items()))
def test_parameter_totals_match_upstream_zoo(name, total):
    assert count_parameters(build_model(name)) == total


def test_yolov10n_inference_path_matches_readme():
    # reference README.md:48: YOLOv10-N 2.3 M params, 6.7 GFLOPs (one-to-one path only)
    m = build_model("yolov10n")
    assert count_parameters(m, exclude_one2many=True) == 2310624
    assert abs(conv_flops(m) / 1e9 - 6.70) < 0.01


@pytest.mark.parametrize("name,gflop", [("yolov8n-seg", 12.00), ("yolov8s-seg", 40.09)])
def test_conv_flops(name, gflop):
    assert abs(conv_flops(build_model(name)) / 1e9 - gflop) < 0.01


def test_output_shapes_and_fuse_equivalence():
    torch.manual_seed(0)
    m = build_model("yolov8n-seg")
    for mod in m.modules():  # non-trivial BN statistics so folding is exercised
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.1)
    x = torch.rand(1, 3, 640, 640)
    with torch.no_grad():
        y, (maps, mc, proto) = m(x)
        assert y.shape == (1, 116, 8400) and proto.shape == (1, 32, 160, 160) and mc.shape == (1, 32, 8400)
        assert [tuple(t.shape[1:]) for t in maps] == [(144, 80, 80), (144, 40, 40), (144, 20, 20)]
        m.fuse()
        y2, (_, _, proto2) = m(x)
    assert (y - y2).abs().max() < 1e-3 and (proto - proto2).abs().max() < 1e-4


def test_rect_input_anchor_count():
    m = build_model("yolov8n-seg").fuse()
    with torch.no_grad():
        y, (_, _, proto) = m(torch.rand(1, 3, 736, 1280))
    assert y.shape == (1, 116, 19320) and proto.shape == (1, 32, 184, 320)


def test_v10_head_output():
    m = build_model("yolov10n").fuse()
    with torch.no_grad():
        y, _ = m(torch.rand(2, 3, 640, 640))
    assert y.shape == (2, 300, 6)
    assert bool((y[:, :-1, 4] >= y[:, 1:, 4]).all())  # descending scores


def test_yolo11_seg_shapes_and_fuse_equivalence():
    torch.manual_seed(0)
    m = build_model("yolo11n-seg")
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.1)
    x = torch.rand(1, 3, 640, 640)
    with torch.no_grad():
        y, (maps, mc, proto) = m(x)
        assert y.shape == (1, 116, 8400) and proto.shape == (1, 32, 160, 160) and mc.shape == (1, 32, 8400)
        m.fuse()
        y2, (_, _, proto2) = m(x)
    assert (y - y2).abs().max() < 1e-3 and (proto - proto2).abs().max() < 1e-4


def test_geometry_class_shift_only_touches_the_class_bias():
    """synth_state_dict(geometry=(1080, 1920)) (BASELINE config C4): the class-branch final biases move by the calibrated
    shift, every other tensor is the one of the 640x640 recipe."""
    import torch
    from oracle.model import build_model
    from yolo_puncture_b200 import synth
    net = build_model("yolov8m-seg")
    specs = [(k, v.shape) for k, v in net.state_dict().items()]
    a = synth.synth_state_dict(specs, "yolov8m-seg")
    b = synth.synth_state_dict(specs, "yolov8m-seg", geometry=(1080, 1920))
    table = synth.load_calibration()
    base, over = table["yolov8m-seg:0"]["cls_shift"], table["yolov8m-seg:0@1080x1920"]["cls_shift"]
    changed = [k for k in a if not torch.equal(a[k], b[k])]
    assert sorted(changed) == sorted(f"model.22.cv3.{i}.2.bias" for i in range(3))
    for i in range(3):
        d = (b[f"model.22.cv3.{i}.2.bias"] - a[f"model.22.cv3.{i}.2.bias"])
        assert torch.allclose(d, torch.full_like(d, over[i] - base[i]), atol=1e-5)
    # a size without an entry falls back to the base recipe
    c = synth.synth_state_dict(specs, "yolov8m-seg", geometry=(123, 457))
    assert all(torch.equal(a[k], c[k]) for k in a)
