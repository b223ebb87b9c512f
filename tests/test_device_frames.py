"""Device-resident frame sources: predict() on CUDA uint8 frames equals predict() on the same pixels
handed over as numpy arrays, bit for bit; nvJPEG-decoded frames (frames.decode_jpegs, replacing the host JPEG reads of reference
yolo_seg/utils/video_reader.py:91-99) are within JPEG-decoder tolerance of cv2's decode of the same bitstream."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,imgsz", [((640, 640), 640), ((480, 640), 640), ((1080, 1920), 1280), ((240, 320), 640)])
def test_predict_on_device_frames_equals_predict_on_host_frames(shape, imgsz):
    from yolo_puncture_b200 import YOLO, synth
    yolo = YOLO("yolov8n-seg", device=0)
    frames = synth.synth_frames(3, shape[0], shape[1], start=5)
    host = yolo.predict(frames, conf=0.25, iou=0.7, retina_masks=True, imgsz=imgsz)
    dev = yolo.predict(torch.from_numpy(np.stack(frames)).cuda(), conf=0.25, iou=0.7, retina_masks=True, imgsz=imgsz)
    assert len(dev) == 3
    if shape == (640, 640):
        assert sum(len(r) for r in host) > 0
    for a, b in zip(host, dev):
        assert b.orig_shape == shape and torch.is_tensor(b.orig_img) and b.orig_img.is_cuda
        assert torch.equal(a.boxes.data, b.boxes.data)
        assert np.array_equal(a.boxes.cpu().numpy().data, b.boxes.cpu().numpy().data)
        assert (a.masks is None) == (b.masks is None)
        if a.masks is not None:
            assert torch.equal(a.masks.raw, b.masks.raw)
    one = yolo.predict(torch.from_numpy(frames[1]).cuda(), conf=0.25, retina_masks=True, imgsz=imgsz)  # a single (H,W,3) tensor
    assert len(one) == 1 and torch.equal(one[0].boxes.data, host[1].boxes.data)


def test_nvjpeg_frames_decode_on_the_device(tmp_path):
    import cv2
    from yolo_puncture_b200 import YOLO, decode_jpegs, synth
    from yolo_puncture_b200._lib import YpbError
    from yolo_puncture_b200.frames import jpeg_size
    frames = synth.synth_frames(4, 480, 640, start=11)
    paths, blobs = [], []
    # 4:4:4 sampling: the two decoders then differ by IDCT rounding only (with 4:2:0 the chroma up-sampling FILTERS differ -
    # libjpeg-turbo's "fancy" triangle filter vs nvJPEG's - which on noise-like frames is several LSB on average)
    enc_args = [cv2.IMWRITE_JPEG_QUALITY, 95]
    if hasattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR"):
        enc_args += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]
    for i, f in enumerate(frames):
        ok, enc = cv2.imencode(".jpg", f, enc_args)
        assert ok
        p = tmp_path / f"frame_{i}.jpg"
        p.write_bytes(enc.tobytes())
        paths.append(str(p))
        blobs.append(enc.tobytes())
    try:
        assert jpeg_size(blobs[0]) == (480, 640)
    except YpbError as e:
        if "libnvjpeg" in str(e):
            pytest.skip("no libnvjpeg on this machine")
        raise
    dev = decode_jpegs(paths, device=0)
    assert dev.shape == (4, 480, 640, 3) and dev.dtype == torch.uint8 and dev.is_cuda
    dev2 = decode_jpegs(blobs, device=0)
    assert torch.equal(dev, dev2)
    ref = np.stack([cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR) for b in blobs])
    diff = np.abs(dev.cpu().numpy().astype(np.int16) - ref.astype(np.int16))
    # two IDCT / chroma up-sampling implementations of the same bitstream: same image up to a few LSB (BGR order included)
    if hasattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR"):
        assert diff.mean() < 1.0 and np.percentile(diff, 99.9) <= 4, (diff.mean(), diff.max())
    else:
        assert diff.mean() < 6.0, (diff.mean(), diff.max())
    yolo = YOLO("yolov8n-seg", device=0)
    res = yolo.predict(dev, conf=0.25, retina_masks=True)
    assert len(res) == 4 and all(r.orig_shape == (480, 640) for r in res)
    with pytest.raises(YpbError):
        jpeg_size(b"not a jpeg at all")
