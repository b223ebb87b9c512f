"""Frame sources that never materialise pixels on the host.

`decode_jpegs(paths_or_bytes, device)` -> (B, H, W, 3) uint8 **BGR** CUDA tensor: the JPEG files the reference's tracker path
writes and re-reads per frame (`yolo_seg/utils/video_reader.py:57-99`, `Image.open(...).convert('RGB')` on the host) are
decoded by nvJPEG straight into device memory; `YOLO.predict()` takes such a tensor as `source` and letterboxes it on the
device, so only the bitstream crosses PCIe.  (H.264 / HEVC through NVDEC would slot in at the same place; this image ships
neither libnvcuvid nor the Video Codec SDK headers.)
"""

import ctypes as C
import os

import torch

from ._lib import check, lib


def jpeg_size(data):
    """(height, width) of a JPEG bitstream (bytes)."""
    h, w = C.c_int(), C.c_int()
    buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
    check(lib().ypb_jpeg_info(buf, len(data), C.byref(h), C.byref(w)))
    return h.value, w.value


def decode_jpegs(sources, device=0, out=None):
    """sources: iterable of file paths or `bytes` objects, all of one frame size.  Returns a (B, H, W, 3) uint8 BGR tensor
    on `device` (the layout `cap.read()` / `cv2.imread` hand the reference).  `out`: optional preallocated destination."""
    blobs = []
    for s in sources:
        if isinstance(s, (bytes, bytearray, memoryview)):
            blobs.append(bytes(s))
        else:
            with open(os.fspath(s), "rb") as f:
                blobs.append(f.read())
    if not blobs:
        raise ValueError("decode_jpegs: no sources")
    dev = torch.device("cuda", device) if not isinstance(device, torch.device) else device
    with torch.cuda.device(dev):
        H, W = jpeg_size(blobs[0])
        if out is None:
            out = torch.empty((len(blobs), H, W, 3), dtype=torch.uint8, device=dev)
        elif tuple(out.shape) != (len(blobs), H, W, 3) or out.dtype != torch.uint8 or not out.is_cuda or not out.is_contiguous():
            raise ValueError("decode_jpegs: `out` must be a contiguous (B, H, W, 3) uint8 CUDA tensor")
        st = torch.cuda.current_stream(dev).cuda_stream
        for i, b in enumerate(blobs):
            buf = (C.c_ubyte * len(b)).from_buffer_copy(b)
            check(lib().ypb_jpeg_decode_bgr(C.c_void_p(st), buf, len(b), C.c_void_p(out[i].data_ptr()), H, W))
    return out
