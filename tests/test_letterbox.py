"""Device LetterBox.  CPU: the coefficient tables reproduce cv2.resize(INTER_LINEAR) bit for bit
when down-scaling (numpy restatement of OpenCV's fixed-point arithmetic = what the kernel computes).  GPU: the kernel's
letterboxed frames equal the host path (cv2.resize + 114 padding) exactly, and predict() gives identical results with
the device and the host LetterBox."""
import numpy as np
import pytest
import torch

from yolo_puncture_b200.model import cv2_linear_tables, letterbox_geometry, letterbox_into


def _resize_fixed_point(src, dw, dh):
    """What letterbox_u8_kernel computes, in numpy."""
    sh, sw = src.shape[:2]
    xi, xa = cv2_linear_tables(sw, dw)
    yi, ya = cv2_linear_tables(sh, dh, vertical=True)
    s = src.astype(np.int32)
    x1 = np.minimum(xi + 1, sw - 1)
    y0, y1 = np.clip(yi, 0, sh - 1), np.clip(yi + 1, 0, sh - 1)  # cv2 clamps the ROWS it fetches, not the coefficients
    hrows = s[:, xi, :] * xa[:, 0].astype(np.int32)[None, :, None] + s[:, x1, :] * xa[:, 1].astype(np.int32)[None, :, None]
    b0, b1 = ya[:, 0].astype(np.int32)[:, None, None], ya[:, 1].astype(np.int32)[:, None, None]
    out = (((b0 * (hrows[y0] >> 4)) >> 16) + ((b1 * (hrows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("sh,sw,dh,dw", [(1080, 1920, 720, 1280), (720, 1280, 360, 640), (1080, 1920, 608, 1080),
                                          (2160, 3840, 720, 1280), (719, 1279, 640, 1138), (100, 200, 37, 53), (64, 48, 64, 48)])
def test_fixed_point_tables_reproduce_cv2_downscale(sh, sw, dh, dw):
    import cv2
    src = np.random.default_rng(sh + sw).integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    assert np.array_equal(_resize_fixed_point(src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("sh,sw,dh,dw", [(480, 640, 960, 1280), (360, 640, 720, 1280), (300, 200, 640, 427), (100, 100, 640, 640),
                                          (540, 960, 720, 1280), (37, 53, 100, 200), (479, 641, 640, 856), (64, 64, 65, 65)])
def test_fixed_point_tables_reproduce_cv2_upscale(sh, sw, dh, dw):
    """Frames smaller than the network input (the project trains at imgsz 1280, reference docs/quickstart.md:57,63)."""
    import cv2
    src = np.random.default_rng(sh + sw).integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    assert np.array_equal(_resize_fixed_point(src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,imgsz", [((1080, 1920), 1280), ((720, 1280), 640), ((720, 1280), 1280), ((900, 700), 640),
                                         ((480, 640), 1280), ((360, 640), 1280), ((300, 200), 640), ((100, 100), 640)])  # + up-scaling
def test_device_letterbox_equals_host_letterbox(shape, imgsz):
    import ctypes as C
    from yolo_puncture_b200._lib import check, lib
    from yolo_puncture_b200 import synth
    frames = synth.synth_frames(3, shape[0], shape[1], start=7)
    new_unpad, top, bottom, left, right = letterbox_geometry(shape, (imgsz, imgsz), True)
    H, W = new_unpad[1] + top + bottom, new_unpad[0] + left + right
    ref = np.empty((3, H, W, 3), np.uint8)
    for i, f in enumerate(frames):
        letterbox_into(ref[i], f, new_unpad, top, left)
    xofs, xa = cv2_linear_tables(shape[1], new_unpad[0])
    yofs, ya = cv2_linear_tables(shape[0], new_unpad[1], vertical=True)
    src = torch.from_numpy(np.stack(frames)).cuda()
    dst = torch.empty((3, H, W, 3), dtype=torch.uint8, device="cuda")
    t = [torch.from_numpy(a).cuda() for a in (xofs, xa, yofs, ya)]
    check(lib().ypb_letterbox_u8(C.c_void_p(torch.cuda.current_stream().cuda_stream), C.c_void_p(src.data_ptr()), 3, shape[0],
                                 shape[1], C.c_void_p(dst.data_ptr()), H, W, new_unpad[0], new_unpad[1], top, left,
                                 C.c_void_p(t[0].data_ptr()), C.c_void_p(t[1].data_ptr()), C.c_void_p(t[2].data_ptr()),
                                 C.c_void_p(t[3].data_ptr()), 114))
    torch.cuda.synchronize()
    assert np.array_equal(dst.cpu().numpy(), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,imgsz", [((1080, 1920), 1280), ((480, 640), 1280), ((240, 320), 640)])  # down- and up-scaling
def test_predict_same_results_with_device_and_host_letterbox(shape, imgsz):
    from yolo_puncture_b200 import YOLO, synth
    yolo = YOLO("yolov8n-seg", device=0)
    frames = synth.synth_frames(3, shape[0], shape[1], start=40)
    yolo.device_letterbox = True
    a = yolo.predict(frames, conf=0.25, retina_masks=True, imgsz=imgsz)
    yolo.device_letterbox = False
    b = yolo.predict(frames, conf=0.25, retina_masks=True, imgsz=imgsz)
    for ra, rb in zip(a, b):
        assert torch.equal(ra.boxes.data, rb.boxes.data)
        assert (ra.masks is None) == (rb.masks is None)
        if ra.masks is not None:
            assert torch.equal(ra.masks.raw, rb.masks.raw)
