"""`YOLO(...).predict(...) -> list[Results]` drop-in over the native sm_100a engine.

Host-side mirror of the UPSTREAM surface the reference calls:
  YOLO(path)                                   reference yolo_seg/app.py:45; yolo_seg/yolo_with_deva.py:226;
                                               dev_tools/auto_speed_calc.py:40
  model.predict(source=frame, conf=…, retina_masks=True, device=…)
                                               reference yolo_seg/app.py:49,91; yolo_seg/yolo_with_deva.py:51
  next(model.model.parameters()).device / model.model.to(device)
                                               reference yolo_seg/yolo_with_deva.py:42,130
Everything numeric — the fused-BN network, DFL decode, NMS, mask decode — runs in libypb200.so on the
GPU; this module only letterboxes frames on the host (cv2, exactly like UPSTREAM data/augment.py::
LetterBox), moves bytes, and wraps outputs.  There is no CPU execution path: without a CUDA device
(or without the compiled library) predict() raises.
"""

import ctypes
import os
import time

import numpy as np
import torch

from ._lib import check, lib
from .engine import MAX_DET, Engine, YpbError
from .results import Boxes, Masks, Results
from .synth import synth_state_dict

KNOWN_SPECS = tuple(f"yolov8{s}-seg" for s in "nsmlx") + ("yolov10n",) + tuple(f"yolo11{s}-seg" for s in "nsmlx")


# ---------------------------------------------------------------------------------------------------
# host-side geometry (UPSTREAM data/augment.py::LetterBox, utils/ops.py::scale_boxes)
# ---------------------------------------------------------------------------------------------------
def letterbox_geometry(shape, new_shape, auto, stride=32):
    """(h0,w0) -> (resized (w,h), top, bottom, left, right) for LetterBox(center=True, scaleup=True)."""
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = dw % stride, dh % stride
    dw, dh = dw / 2, dh / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_unpad, top, bottom, left, right


def letterbox_into(dst, img, new_unpad, top, left):
    """Resize `img` (cv2 INTER_LINEAR, as upstream) and paste it into the 114-padded canvas `dst`."""
    if (img.shape[1], img.shape[0]) != new_unpad:
        import cv2
        img = cv2.resize(img, new_unpad, interpolation=cv2.INTER_LINEAR)
    h, w = img.shape[:2]
    if (h, w) == dst.shape[:2]:
        np.copyto(dst, img)  # no padding: one pass over the pixels
        return
    dst[:top] = 114
    dst[top + h:] = 114
    dst[top:top + h, :left] = 114
    dst[top:top + h, left + w:] = 114
    dst[top:top + h, left:left + w] = img


def cv2_linear_tables(src_n, dst_n, vertical=False):
    """Source index and 11-bit fixed-point coefficient pair of every output column (or row) of OpenCV's 8-bit
    INTER_LINEAR resize, built the way cv2 builds them (resize.cpp): scale = 1/(dst/src) in double,
    f = float((d+0.5)*scale - 0.5), s = floor(f), f -= s, coefficients saturate_cast<short>((1-f, f) * 2048).
    Columns: at both borders cv2 clamps the index AND zeroes f.  Rows (vertical=True): cv2 keeps f and clamps the two
    source ROWS instead when it fetches them (s may be -1 or src_n-1 here; the kernel clamps s and s+1) - the two
    border conventions differ by one LSB when up-scaling, which is where rows above/below the first/last source row
    centre exist.  Bit-exact against cv2.resize, down- and up-scaling (tests/test_letterbox.py)."""
    scale = 1.0 / (dst_n / src_n)
    d = np.arange(dst_n, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if not vertical:
        lo, hi = s < 0, s >= src_n - 1
        f[lo | hi] = 0.0
        s[lo] = 0
        s[hi] = src_n - 1
    coef = np.stack([np.rint((np.float32(1.0) - f) * np.float32(2048.0)), np.rint(f * np.float32(2048.0))], 1).astype(np.int16)
    return s, coef


_POOL = None


def _pool():
    """Host threads for frame staging (cv2.resize and numpy copies release the GIL)."""
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1))
    return _POOL


def box_xform(net_hw, orig_hw):
    """[pad_w, pad_h, gain, W0, H0] of ops.scale_boxes for one frame."""
    gain = min(net_hw[0] / orig_hw[0], net_hw[1] / orig_hw[1])
    pad_w = round((net_hw[1] - orig_hw[1] * gain) / 2 - 0.1)
    pad_h = round((net_hw[0] - orig_hw[0] * gain) / 2 - 0.1)
    return [float(pad_w), float(pad_h), float(gain), float(orig_hw[1]), float(orig_hw[0])]


def _to_bgr_array(src):
    if isinstance(src, np.ndarray):
        if src.ndim != 3 or src.shape[2] != 3 or src.dtype != np.uint8:
            raise ValueError("ndarray sources must be (H,W,3) uint8 BGR")
        return src, None
    if isinstance(src, (str, os.PathLike)):
        import cv2
        im = cv2.imread(str(src))
        if im is None:
            raise FileNotFoundError(f"cannot read image {src}")
        return im, str(src)
    if hasattr(src, "convert"):  # PIL.Image: RGB -> BGR (UPSTREAM LoadPilAndNumpy._single_check)
        im = np.asarray(src.convert("RGB"))[:, :, ::-1]
        return np.ascontiguousarray(im), None
    raise TypeError(f"unsupported source type {type(src)}")


class _ModelProxy:
    """What `yolo.model` must be for the reference: `.parameters()` on the engine's device and `.to()`."""

    def __init__(self, owner):
        self._owner = owner

    def parameters(self):
        yield self._owner._device_token

    def to(self, device=None, *a, **k):
        if device is not None:
            self._owner._set_device(device)
        return self

    def eval(self):
        return self

    @property
    def names(self):
        return self._owner.names



def _addr(a):
    """Address of a C-contiguous numpy array's first byte (a.ctypes.data costs ~2 us per frame; this path 0.7 us)."""
    try:
        return ctypes.addressof(ctypes.c_char.from_buffer(a))
    except (TypeError, ValueError, BufferError):  # read-only arrays do not export a writable buffer
        return a.ctypes.data

class YOLO:
    """Drop-in for `ultralytics.YOLO` on the predict path.

    model: a spec name ("yolov8s-seg"; synthetic weights), or a file written by
    `torch.save({"spec": ..., "nc": ..., "names": ..., "state_dict": ...})` / a bare upstream-named
    state_dict (real `.pt` pickles need the upstream classes to unpickle — export their
    `model.state_dict()` on a box that has ultralytics)."""

    def __init__(self, model="yolov8n-seg", task=None, verbose=False, nc=None, state_dict=None, device=None, seed=0,
                 synth_geometry=None):
        names = None
        spec = model
        if isinstance(model, (str, os.PathLike)) and os.path.exists(str(model)):
            blob = torch.load(str(model), map_location="cpu", weights_only=True)
            if "state_dict" in blob:
                spec, state_dict = blob["spec"], blob["state_dict"]
                nc = blob.get("nc", nc)
                names = blob.get("names")
            else:
                state_dict = blob
                spec = os.path.splitext(os.path.basename(str(model)))[0]
        spec = str(spec)
        for ext in (".pt", ".yaml", ".pth"):
            if spec.endswith(ext):
                spec = spec[: -len(ext)]
        if nc is None and state_dict is not None:
            key = next((k for k in state_dict if k.endswith(".cv3.0.2.bias")), None)
            nc = int(state_dict[key].numel()) if key else 80
        self.spec = spec
        self.nc = int(nc or 80)
        self.task = task or ("segment" if spec.endswith("-seg") else "detect")
        self.names = names or {i: f"class{i}" for i in range(self.nc)}
        self.engine = Engine(spec, self.nc)
        if state_dict is None:
            # synthetic weights (no checkpoint given); synth_geometry=(h, w) selects the class shift calibrated on frames
            # of that size when the calibration file has one (synth.synth_state_dict)
            state_dict = synth_state_dict([(n, s) for n, s, _ in self.engine.weight_specs()], spec, seed, nc=self.nc,
                                          geometry=synth_geometry)
        self.engine.load_state_dict(state_dict)
        self._state_dict = state_dict  # kept for the head-pass engine (second plan size), built on first use
        self._head_engine = None
        self._device = None
        self._device_token = torch.zeros(1)
        self._staging = {}
        self.overrides = {}
        # Frames per engine pass inside predict().  None: automatic schedule - up to 32 frames run as one pass; larger
        # groups start with a HEAD pass of 16 frames so the engine starts after a quarter of the first upload, and the
        # rest follows in passes of up to 48 (fixed per-pass cost ~0.5 ms on yolov8s-seg: fewer, larger passes win).
        # An int forces uniform passes of that size.  H2D of pass k+1 always overlaps compute of pass k.
        self.micro_batch = None
        self.head_pass = 16
        self.head_stage_chunk = 4  # frames per staging chunk of the first pass (later passes: stage_chunk)
        self.stage_chunk = 16     # frames per staging call / H2D copy when the caller's frames are pageable
        # True: staging is queued on the native pool behind a stream gate (cudaLaunchHostFunc) and the host never blocks.
        # Measured slower on B200 hosts (9.7 vs 7.4 ms per 64-frame call: the gate's callback latency and 16 pool threads
        # next to the CUDA callback thread on 16 vCPUs), so the blocking, spin-hot pool is the default.
        self.async_staging = False
        self.stage_threads = None  # host threads of the staging pool (None: min(8, this rank's share of the cores - 2))
        self.device_letterbox = True  # resize + pad on the GPU (False: cv2 on host threads, exactly upstream's LetterBox)
        if device is not None:
            self._set_device(device)

    # ------------------------------------------------------------------ device handling
    @staticmethod
    def _parse_device(device):
        if device is None or device == "":
            return torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        if isinstance(device, int):
            return torch.device("cuda", device)
        if isinstance(device, str) and device.isdigit():
            return torch.device("cuda", int(device))
        d = torch.device(device)
        if d.type != "cuda":
            raise YpbError(f"device '{device}': this engine only runs on CUDA sm_100a GPUs (no CPU path)")
        return torch.device("cuda", d.index or 0)

    def _set_device(self, device):
        d = self._parse_device(device)
        if self._device == d:
            return
        if not torch.cuda.is_available():
            raise YpbError("no CUDA device available: the B200 engine has no CPU fallback")
        self.engine.finalize(d)
        if self._head_engine is not None:
            self._head_engine.finalize(d)
        self._device = d
        self._device_token = torch.zeros(1, device=d)
        self._staging = {}

    def to(self, device):
        self._set_device(device)
        return self

    @property
    def model(self):
        return _ModelProxy(self)

    @property
    def device(self):
        return self._device

    # ------------------------------------------------------------------ predict
    def __call__(self, source=None, **kwargs):
        return self.predict(source, **kwargs)

    def predict(self, source=None, stream=False, conf=0.25, iou=0.7, retina_masks=False, device=None, imgsz=None,
                max_det=MAX_DET, classes=None, agnostic_nms=False, half=False, batch=64, verbose=False, **ignored):
        """One `Results` per frame, in input order.  Unknown kwargs are accepted and ignored like upstream."""
        if source is None:
            raise ValueError("predict() needs a source")
        if device is not None or self._device is None:
            self._set_device(device)
        if torch.is_tensor(source):
            # frames that already live on the GPU (frames.decode_jpegs, a capture / NVDEC ring): (H,W,3) or (B,H,W,3) uint8 BGR
            if not source.is_cuda or source.dtype != torch.uint8 or source.shape[-1] != 3 or source.dim() not in (3, 4):
                raise ValueError("tensor sources must be CUDA uint8 (H,W,3) or (B,H,W,3) BGR frames")
            res = self._predict_device_frames(source if source.dim() == 4 else source[None], imgsz or self.overrides.get("imgsz", 640),
                                              conf, iou, retina_masks, min(int(max_det), MAX_DET), classes, agnostic_nms, batch)
            return iter(res) if stream else res
        srcs = list(source) if isinstance(source, (list, tuple)) else [source]
        frames, paths = zip(*[_to_bgr_array(s) for s in srcs])
        imgsz = imgsz or self.overrides.get("imgsz", 640)
        imgsz = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
        max_det = min(int(max_det), MAX_DET)
        same_shape = len({f.shape for f in frames}) == 1
        results = [None] * len(frames)
        # frames of one original shape share letterbox geometry and mask size: batch them together
        groups = {}
        for i, f in enumerate(frames):
            groups.setdefault(f.shape[:2], []).append(i)
        for shape, idxs in groups.items():
            for s in range(0, len(idxs), batch):
                chunk = idxs[s:s + batch]
                out = self._predict_batch([frames[i] for i in chunk], shape, imgsz, same_shape, conf, iou,
                                          retina_masks, max_det, classes, agnostic_nms)
                for i, r in zip(chunk, out):
                    r.path = paths[i]
                    results[i] = r
        return iter(results) if stream else results

    def _predict_device_frames(self, frames, imgsz, conf, iou, retina, max_det, classes, agnostic, batch):
        """Device-resident frames (B, H0, W0, 3): LetterBox on the device when needed, one engine pass per chunk; nothing
        but counts and boxes crosses PCIe.  `Results.orig_img` is the frame's device tensor."""
        if frames.device != self._device:
            frames = frames.to(self._device)
        frames = frames.contiguous()
        imgsz = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
        n_all, shape = frames.shape[0], (int(frames.shape[1]), int(frames.shape[2]))
        new_unpad, top, bottom, left, right = letterbox_geometry(shape, imgsz, True)
        H, W = new_unpad[1] + top + bottom, new_unpad[0] + left + right
        if H % 32 or W % 32:
            raise YpbError(f"letterboxed size {H}x{W} is not a multiple of 32 (imgsz={imgsz})")
        seg = self.task == "segment"
        mh, mw = shape if retina else (H, W)
        direct = shape == (H, W)
        out = []
        t0 = time.perf_counter()
        with torch.cuda.device(self._device):
            main = torch.cuda.current_stream(self._device)
            cmask = None
            if classes is not None:
                words = np.zeros(((self.nc + 31) // 32,), np.uint32)
                for c in classes:
                    words[int(c) >> 5] |= np.uint32(1) << np.uint32(int(c) & 31)
                cmask = torch.from_numpy(words.view(np.int32)).to(self._device)
            for lo in range(0, n_all, batch):
                chunk = frames[lo:lo + batch]
                B = int(chunk.shape[0])
                eng = self.engine
                eng.plan(B, H, W)
                eng.use_outputs(0)
                if direct:
                    net_in = chunk
                else:
                    key = ("devlb", B, shape, H, W)
                    st = self._staging.get(key)
                    if st is None:
                        xofs, xa = cv2_linear_tables(shape[1], new_unpad[0])
                        yofs, ya = cv2_linear_tables(shape[0], new_unpad[1], vertical=True)
                        st = self._staging[key] = {"dst": torch.empty((B, H, W, 3), dtype=torch.uint8, device=self._device),
                                                   "t": [torch.from_numpy(a).to(self._device) for a in (xofs, xa, yofs, ya)]}
                    net_in = st["dst"]
                    check(lib().ypb_letterbox_u8(ctypes.c_void_p(main.cuda_stream), ctypes.c_void_p(chunk.data_ptr()), B, shape[0], shape[1],
                                                 ctypes.c_void_p(net_in.data_ptr()), H, W, new_unpad[0], new_unpad[1], top, left,
                                                 *[ctypes.c_void_p(t.data_ptr()) for t in st["t"]], 114))
                xf = torch.tensor([box_xform((H, W), shape)] * B, dtype=torch.float32).to(self._device)
                eng.infer(net_in, xf, conf, iou, max_det, agnostic, cmask)
                o = eng.out_sets[0]
                counts = o.count[:B].cpu()
                det_h = o.det[:B].cpu()
                n_k = int(counts.sum())
                det = o.det[:B].clone()
                masks = None
                if seg and n_k:
                    masks = torch.empty((n_k, mh, mw), dtype=torch.uint8, device=self._device)
                    eng.masks(masks, retina, mh, mw)
                err = eng.device_error()
                if err:
                    raise YpbError(f"device pipeline error word 0x{err:x}")
                speed = {"preprocess": 0.0, "inference": (time.perf_counter() - t0) * 1e3 / max(B, 1), "postprocess": 0.0}
                off = 0
                for j, n in enumerate(counts.tolist()):
                    boxes = Boxes(det[j, :n], shape, n=n, host=(lambda d=det_h, j=j, n=n: d[j, :n]))
                    m = Masks(masks[off:off + n], shape, n=n) if (masks is not None and n) else None
                    if m is not None:
                        m.cropped = bool(retina)
                    off += n
                    out.append(Results(chunk[j], None, self.names, boxes=boxes, masks=m, speed=speed))
        return out

    def _head(self):
        """Second engine instance (same weights) so that two pass sizes stay planned side by side."""
        if self._head_engine is None:
            e = Engine(self.spec, self.nc)
            e.load_state_dict(self._state_dict)
            e.finalize(self._device)
            self._head_engine = e
        return self._head_engine

    def _schedule(self, B, pinned=True):
        """[(engine, lo, hi, capacity, slot)]: the engine passes of a group of B frames (slot = input / output set).
        The same schedule serves frames in pageable and in page-locked memory: pageable frames are staged into the pinned
        ring in chunks of `stage_chunk` frames by the native thread pool (csrc/host_stage.cpp), every chunk's H2D copy is
        issued as soon as it is staged, so staging, PCIe transfer and the previous engine pass overlap."""
        if self.micro_batch:
            mb = min(B, int(self.micro_batch))
            return [(self.engine, lo, min(lo + mb, B), mb, k & 1) for k, lo in enumerate(range(0, B, mb))]
        if B <= 32 or B - self.head_pass < self.head_pass:
            return [(self.engine, 0, B, B, 0)]
        passes = [(self._head(), 0, self.head_pass, self.head_pass, 0)]
        rest = B - self.head_pass
        n = (rest + 47) // 48
        mb = (rest + n - 1) // n
        for k, lo in enumerate(range(self.head_pass, B, mb)):
            passes.append((self.engine, lo, min(lo + mb, B), mb, (k + 1) & 1))
        return passes

    def _buffers(self, B, caps, H, W):
        key = (B, caps, H, W)
        if key not in self._staging:
            self._staging = {key: {
                "host": torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory(),
                "dev": [torch.empty((c, H, W, 3), dtype=torch.uint8, device=self._device) for c in caps],
                "xf_dev": torch.empty((max(caps), 5), dtype=torch.float32, device=self._device),
                "copy_stream": torch.cuda.Stream(device=self._device),
                "in_free": [None, None],
            }}
        return self._staging[key]

    def _letterbox_buffers(self, buf, B, caps, shape, new_unpad, need_host=True):
        """Raw-frame staging and the cv2 coefficient tables of the device LetterBox for one source shape."""
        mb = caps
        key = (B, mb, tuple(shape[:2]), tuple(new_unpad))
        lb = buf.get("letterbox")
        if lb is not None and lb["key"] == key and need_host and lb["raw_host"] is None:
            lb["raw_host"] = torch.empty((B, shape[0], shape[1], 3), dtype=torch.uint8).pin_memory()
        if lb is None or lb["key"] != key:
            xofs, xa = cv2_linear_tables(shape[1], new_unpad[0])
            yofs, ya = cv2_linear_tables(shape[0], new_unpad[1], vertical=True)
            dev = self._device
            lb = {"key": key,
                  "raw_host": torch.empty((B, shape[0], shape[1], 3), dtype=torch.uint8).pin_memory() if need_host else None,
                  "raw_dev": [torch.empty((c, shape[0], shape[1], 3), dtype=torch.uint8, device=dev) for c in caps],
                  "xofs": torch.from_numpy(xofs).to(dev), "xa": torch.from_numpy(xa).to(dev),
                  "yofs": torch.from_numpy(yofs).to(dev), "ya": torch.from_numpy(ya).to(dev)}
            buf["letterbox"] = lb
        return lb

    def _predict_batch(self, frames, shape, imgsz, auto, conf, iou, retina, max_det, classes, agnostic):
        """One group of same-shape frames.  The group runs as the engine passes of `_schedule()` through static engine
        plans: frames are letterboxed into pinned memory by host threads, each pass's H2D copy is issued on a copy
        stream as soon as its frames are staged, and it overlaps the engine pass before it."""
        B = len(frames)
        t0 = time.perf_counter()
        tm = self.last_timing = {"stage_ms": 0.0, "wait_ms": 0.0}  # host-side breakdown of the last group (diagnostics)
        new_unpad, top, bottom, left, right = letterbox_geometry(shape, imgsz, auto)
        H, W = new_unpad[1] + top + bottom, new_unpad[0] + left + right
        if H % 32 or W % 32:
            raise YpbError(f"letterboxed size {H}x{W} is not a multiple of 32 (imgsz={imgsz})")
        seg = self.task == "segment"
        mh, mw = (shape[0], shape[1]) if retina else (H, W)
        with torch.cuda.device(self._device):
            direct = (shape[0], shape[1]) == (H, W)  # frames already have the network size: staging is a plain copy
            # LetterBox on the device (bit-exact with cv2's 8-bit INTER_LINEAR, down- and up-scaling): the raw frames are
            # uploaded and resized + padded by one kernel instead of cv2.resize on host threads
            dev_lb = (not direct) and self.device_letterbox
            futs = None
            pinned = False
            if direct or dev_lb:
                frames_c = [f if f.flags["C_CONTIGUOUS"] else np.ascontiguousarray(f) for f in frames]
                nbytes = shape[0] * shape[1] * 3  # == H * W * 3 when direct
                src_ptrs = (ctypes.c_void_p * B)(*[_addr(f) for f in frames_c])
                sizes = (ctypes.c_size_t * B)(*([nbytes] * B))
                # staging threads: the rank's share of the host cores minus two (the Python thread and the CUDA driver's)
                share = len(os.sched_getaffinity(0)) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))
                # (8: measured on 16-vCPU B200 hosts, 60 calls of 64 frames each - median call 6.3 ms with 8 threads, 7.3 ms
                # with 14, 8.4 ms with 4; more threads only add scheduling jitter on a shared host)
                nthreads = self.stage_threads or max(1, min(8, share - 2))
                # frames that already live in page-locked memory go to the device from where they are (no staging copy)
                flag = ctypes.c_int(0)
                check(lib().ypb_hosts_are_pinned(src_ptrs, B, ctypes.byref(flag)))
                pinned = bool(flag.value)
            passes = self._schedule(B, pinned)
            n_mb = len(passes)
            caps = tuple(max([c for (_, _, _, c, sl) in passes if sl == slot] or [1]) for slot in range(2))
            for (e, _, _, cap, _) in passes:
                e.plan(cap, H, W)
            buf = self._buffers(B, caps, H, W)
            host = buf["host"].numpy()
            main = torch.cuda.current_stream(self._device)
            cs = buf["copy_stream"]
            xf_key = ((H, W), tuple(shape[:2]))
            if buf.get("xf_key") != xf_key:  # same geometry as the last call: the device copy is still valid
                buf["xf_dev"].copy_(torch.tensor([box_xform((H, W), shape)] * max(caps), dtype=torch.float32))
                buf["xf_key"] = xf_key
            cmask = None
            if classes is not None:
                words = np.zeros(((self.nc + 31) // 32,), np.uint32)
                for c in classes:
                    words[int(c) >> 5] |= np.uint32(1) << np.uint32(int(c) & 31)
                cmask = torch.from_numpy(words.view(np.int32)).to(self._device)
            if direct or dev_lb:
                if dev_lb:
                    lb = self._letterbox_buffers(buf, B, caps, shape, new_unpad, need_host=not pinned)
                if not pinned:
                    stage_host = lb["raw_host"] if dev_lb else buf["host"]
                    d0, dstep = stage_host.data_ptr(), stage_host.stride(0) * stage_host.element_size()
                    dst_ptrs = (ctypes.c_void_p * B)(*range(d0, d0 + B * dstep, dstep))
            else:
                futs = [_pool().submit(letterbox_into, host[i], f, new_unpad, top, left) for i, f in enumerate(frames)]
            cs.wait_stream(main)
            h2d_done = []
            t1 = time.perf_counter()
            dets, cnts, mask_parts = [], [], []

            def enqueue_h2d(k):
                _, lo, hi, _, slot = passes[k]
                vp = ctypes.sizeof(ctypes.c_void_p)
                with torch.cuda.stream(cs):
                    if buf["in_free"][slot] is not None:
                        cs.wait_event(buf["in_free"][slot])
                    target = lb["raw_dev"][slot] if dev_lb else buf["dev"][slot]
                    if pinned:  # frames already live in page-locked memory: copy from where they are
                        check(lib().ypb_h2d_frames(ctypes.c_void_p(cs.cuda_stream), ctypes.c_void_p(target.data_ptr()),
                                                   ctypes.byref(src_ptrs, lo * vp), nbytes, hi - lo))
                    elif direct or dev_lb:
                        # pageable frames: native multi-threaded copy into the pinned ring (ctypes drops the GIL), a chunk
                        # at a time, each chunk's H2D issued right behind it: chunk i+1 is staged while chunk i crosses PCIe
                        stage_host = lb["raw_host"] if dev_lb else buf["host"]
                        # the first pass's H2D is on the call's critical path (nothing overlaps it): smaller chunks there
                        step = max(1, int(self.stage_chunk if k else self.head_stage_chunk))
                        for c0 in range(lo, hi, step):
                            c1 = min(c0 + step, hi)
                            ts = time.perf_counter()
                            if self.async_staging:  # queue the copy and gate the copy stream on it: the host moves on
                                check(lib().ypb_stage_frames_gated(ctypes.c_void_p(cs.cuda_stream), ctypes.byref(dst_ptrs, c0 * vp),
                                                                   ctypes.byref(src_ptrs, c0 * vp),
                                                                   ctypes.byref(sizes, c0 * ctypes.sizeof(ctypes.c_size_t)), c1 - c0, nthreads))
                            else:
                                check(lib().ypb_stage_frames(ctypes.byref(dst_ptrs, c0 * vp), ctypes.byref(src_ptrs, c0 * vp),
                                                             ctypes.byref(sizes, c0 * ctypes.sizeof(ctypes.c_size_t)), c1 - c0, nthreads))
                            tm["stage_ms"] += (time.perf_counter() - ts) * 1e3
                            # the staged slots are adjacent: one cudaMemcpyAsync per chunk, issued like the pinned-frame path
                            # (torch's copy_ from a pinned tensor adds a pointer query and a host-allocator event per copy;
                            # with it the call time had 10-180 ms outliers that the pinned-frame arm never showed)
                            check(lib().ypb_h2d_frames(ctypes.c_void_p(cs.cuda_stream),
                                                       ctypes.c_void_p(target.data_ptr() + (c0 - lo) * nbytes),
                                                       ctypes.byref(dst_ptrs, c0 * vp), nbytes, c1 - c0))
                    else:
                        for fu in futs[lo:hi]:
                            fu.result()
                        target[: hi - lo].copy_(buf["host"][lo:hi], non_blocking=True)
                    if dev_lb:
                        check(lib().ypb_letterbox_u8(
                            ctypes.c_void_p(cs.cuda_stream), ctypes.c_void_p(target.data_ptr()), hi - lo, shape[0], shape[1],
                            ctypes.c_void_p(buf["dev"][slot].data_ptr()), H, W, new_unpad[0], new_unpad[1], top, left,
                            ctypes.c_void_p(lb["xofs"].data_ptr()), ctypes.c_void_p(lb["xa"].data_ptr()),
                            ctypes.c_void_p(lb["yofs"].data_ptr()), ctypes.c_void_p(lb["ya"].data_ptr()), 114))
                    ev = torch.cuda.Event()
                    ev.record(cs)
                h2d_done.append(ev)

            # Software pipeline: pass k+1 is enqueued BEFORE the host blocks on pass k's detection counts, so the GPU
            # never idles between passes; pass k's outputs live in output set k&1 and its prototypes are copied aside
            # (device-to-device) because the workspace is reused by pass k+1.  Counts AND boxes of a pass come back in
            # one copy on a side stream that only waits for THAT pass, and the pass's mask decode runs on the side
            # stream too: masks of pass k overlap the network of pass k+1 instead of queueing behind it.
            protos = buf.setdefault("proto_copy", [None, None])
            side = buf.setdefault("side_stream", torch.cuda.Stream(device=self._device))
            hostbuf = buf.setdefault("host_out", {})
            inf_done, dets_h, side_done = [], [], []

            def launch(k):
                eng, lo, hi, cap, slot = passes[k]
                main.wait_event(h2d_done[k])
                if k >= 2:  # output set / proto copy k&1 are still being read by the side stream's work of pass k-2
                    main.wait_event(side_done[k - 2])
                eng.use_outputs(slot)
                eng.infer(buf["dev"][slot][:cap], buf["xf_dev"][:cap], conf, iou, max_det, agnostic, cmask)
                fr = torch.cuda.Event()
                fr.record(main)
                buf["in_free"][slot] = fr
                if seg:
                    pv = eng.proto_view()
                    if protos[k & 1] is None or protos[k & 1].numel() != pv.numel():
                        protos[k & 1] = torch.empty_like(pv)
                    protos[k & 1].copy_(pv, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(main)
                o = eng.out_sets[slot]
                n = hi - lo
                hb = hostbuf.get((k & 1, cap))
                if hb is None:
                    hb = hostbuf[(k & 1, cap)] = (torch.empty((cap,), dtype=torch.int32).pin_memory(),
                                                  torch.empty((cap, MAX_DET, 6), dtype=torch.float32).pin_memory(),
                                                  torch.zeros((1,), dtype=torch.int32).pin_memory())
                with torch.cuda.stream(side):
                    side.wait_event(ev)
                    hb[0][:n].copy_(o.count[:n], non_blocking=True)
                    hb[1][:n].copy_(o.det[:n], non_blocking=True)
                    eng.device_error_async(side, hb[2])  # the kernels' error word rides along: no separate host sync
                    done = torch.cuda.Event()
                    done.record(side)
                inf_done.append(done)

            def finish(k):
                eng, lo, hi, cap, slot = passes[k]
                o = eng.out_sets[slot]
                n = hi - lo
                ts = time.perf_counter()
                inf_done[k].synchronize()  # the host sync of this pass: its counts and boxes are in pinned memory now
                tm["wait_ms"] += (time.perf_counter() - ts) * 1e3
                hb = hostbuf[(k & 1, cap)]
                if int(hb[2][0]):
                    raise YpbError(f"device pipeline error word 0x{eng.device_error():x}")
                counts = hb[0][:n].tolist()
                n_k = sum(counts)
                cnts.append(counts)
                dets.append(o.det[:n].clone() if n_k else None)
                dets_h.append(hb[1][:n].clone() if n_k else None)
                if seg and n_k:
                    with torch.cuda.stream(side):
                        m = torch.empty((n_k, mh, mw), dtype=torch.uint8, device=self._device)
                        if hi - lo < cap:
                            o.count[hi - lo:].zero_()  # padded tail of the last pass
                        eng.masks(m, retina, mh, mw, proto=protos[k & 1], outputs=o, stream=side)
                    m.record_stream(main)
                    mask_parts.append(m)
                else:
                    mask_parts.append(None)
                sd = torch.cuda.Event()
                sd.record(side)
                side_done.append(sd)

            enqueue_h2d(0)
            for k in range(n_mb):
                launch(k)
                if k + 1 < n_mb:
                    enqueue_h2d(k + 1)  # its staging + PCIe transfer overlap this pass
                if k > 0:
                    finish(k - 1)
            finish(n_mb - 1)
            # The masks of the last pass may still be in flight when predict() returns (as upstream's are): they are ordered
            # on the caller's stream, and every pass's error word came back with its boxes.
            main.wait_stream(side)
            t2 = time.perf_counter()
            tm["enqueue_to_done_ms"] = (t2 - t1) * 1e3
            tm["prologue_ms"] = (t1 - t0) * 1e3
            t3 = t2
        speed = {"preprocess": (t1 - t0) * 1e3 / B, "inference": (t2 - t1) * 1e3 / B, "postprocess": (t3 - t2) * 1e3 / B}
        out = []
        empty = None
        for k in range(n_mb):
            _, lo, hi, _, _ = passes[k]
            counts, det, det_h, masks, off = cnts[k], dets[k], dets_h[k], mask_parts[k], 0
            for j in range(hi - lo):
                n = counts[j]
                shp = frames[lo + j].shape[:2]
                if n:
                    boxes = Boxes(None, shp, lazy=(lambda d=det, j=j, n=n: d[j, :n]), n=n,
                                  host=(lambda d=det_h, j=j, n=n: d[j, :n]))
                    m = Masks(None, shp, lazy=(lambda mm=masks, a=off, b=off + n: mm[a:b]), n=n) if masks is not None else None
                    if m is not None:
                        m.cropped = bool(retina)  # retina masks are cropped to the frame-space boxes of `boxes`
                else:
                    if empty is None:
                        empty = torch.zeros((0, 6), device=self._device)
                    boxes, m = Boxes(empty, shp, n=0), None
                off += n
                out.append(Results(frames[lo + j], None, self.names, boxes=boxes, masks=m, speed=speed))
        tm["results_ms"] = (time.perf_counter() - t3) * 1e3
        tm["threads"] = nthreads if (direct or dev_lb) else 0
        return out
