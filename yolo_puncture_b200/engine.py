"""Thin Python handle on the native engine (C ABI in include/ypb200.h).

PyTorch is used for device memory, streams and host<->device copies only; every FLOP of the
detector runs in libypb200.so.  One Engine = one model on one GPU (upstream holds a lock around
inference; concurrency = one engine per GPU).
"""

import ctypes as C

import numpy as np
import torch

from ._lib import InferParams, YpbError, check, diag_lib, lib

MAX_DET = 300


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class OutputSet:
    """Caller-owned output buffers of one engine pass."""

    def __init__(self, B, nm, device):
        self.det = torch.zeros((B, MAX_DET, 6), dtype=torch.float32, device=device)
        self.det_lb = torch.zeros((B, MAX_DET, 4), dtype=torch.float32, device=device)
        self.keep = torch.zeros((B, MAX_DET), dtype=torch.int32, device=device)
        self.coef = torch.zeros((B, MAX_DET, nm), dtype=torch.float32, device=device)
        self.count = torch.zeros((B,), dtype=torch.int32, device=device)
        self.offsets = torch.zeros((B + 1,), dtype=torch.int32, device=device)


class Engine:
    def __init__(self, spec, nc=80):
        self._lib = lib()
        self._h = C.c_void_p()
        check(self._lib.ypb_engine_create(spec.encode(), int(nc), C.byref(self._h)))
        self.spec, self.nc = spec, nc
        self.nm = self._lib.ypb_num_mask_coefs(self._h)
        self.device = None
        self.shape = None  # (B, H, W)
        self.ws = None
        self._views = {}

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.ypb_engine_destroy(h)
            except Exception:
                pass

    # ------------------------------------------------------------------ weights
    def weight_specs(self, include_unused=True):
        """[(name, shape, used)] in upstream state_dict naming."""
        out = []
        for i in range(self._lib.ypb_weight_count(self._h)):
            name, ndim, shape, used = C.c_char_p(), C.c_int(), (C.c_int64 * 4)(), C.c_int()
            check(self._lib.ypb_weight_info(self._h, i, C.byref(name), C.byref(ndim), C.byref(shape), C.byref(used)))
            if used.value or include_unused:
                out.append((name.value.decode(), tuple(shape[d] for d in range(ndim.value)), bool(used.value)))
        return out

    def load_state_dict(self, sd, strict=True):
        """Copy an upstream-named fp32 state_dict into the engine (BN folding happens in finalize())."""
        specs = {n: (s, u) for n, s, u in self.weight_specs()}
        for name, t in sd.items():
            if name.endswith("num_batches_tracked"):
                continue
            if name not in specs:
                if strict:
                    raise YpbError(f"unexpected weight '{name}' for {self.spec}")
                continue
            t = torch.as_tensor(t).detach().to("cpu", torch.float32).contiguous()
            if tuple(t.shape) != specs[name][0]:
                raise YpbError(f"shape mismatch for '{name}': {tuple(t.shape)} vs {specs[name][0]}")
            check(self._lib.ypb_load_weight(self._h, name.encode(), C.c_void_p(t.data_ptr()), t.numel()))
        missing = [n for n, (s, u) in specs.items() if u and n not in sd]
        if missing and strict:
            raise YpbError(f"missing weights: {missing[:5]}{'...' if len(missing) > 5 else ''}")

    def finalize(self, device=0):
        dev = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        check(self._lib.ypb_finalize_weights(self._h, dev.index or 0))
        self.device = dev
        self.shape = None

    # ------------------------------------------------------------------ planning
    def plan_only(self, B, H, W):
        """Shape inference + workspace size without touching CUDA (works on a CPU-only box)."""
        n = C.c_size_t()
        check(self._lib.ypb_plan(self._h, B, H, W, C.byref(n)))
        return n.value

    def plan(self, B, H, W):
        if self.shape == (B, H, W):
            return
        if self.device is None:
            raise YpbError("finalize() weights before planning")
        nbytes = self.plan_only(B, H, W)
        with torch.cuda.device(self.device):
            self.ws = torch.zeros(nbytes + 1024, dtype=torch.uint8, device=self.device)
            off = (-self.ws.data_ptr()) % 1024
            self._ws_off = off
            check(self._lib.ypb_bind_workspace(self._h, C.c_void_p(self.ws.data_ptr() + off), nbytes))
            # two sets of output buffers: predict() double-buffers engine passes (set 0 is the default one)
            self.out_sets = [OutputSet(B, max(self.nm, 1), self.device) for _ in range(2)]
            self.use_outputs(0)
            self.mask_status = torch.zeros((2,), dtype=torch.int32, device=self.device)
        self.shape = (B, H, W)
        self._views = {}
        self.anchors = self._lib.ypb_num_anchors(self._h)
        self.launches = self._lib.ypb_kernel_launches(self._h)
        self.conv_flops = self._lib.ypb_conv_flops(self._h)

    # ------------------------------------------------------------------ run
    def infer(self, frames, xform, conf=0.25, iou=0.7, max_det=MAX_DET, agnostic_nms=False, class_mask=None):
        """frames: cuda uint8 (B,H,W,3) letterboxed BGR; xform: cuda fp32 (B,5).  Enqueues on the current stream."""
        B, H, W = self.shape
        assert frames.is_cuda and frames.dtype == torch.uint8 and tuple(frames.shape) == (B, H, W, 3) and frames.is_contiguous()
        assert xform.is_cuda and xform.dtype == torch.float32 and tuple(xform.shape) == (B, 5) and xform.is_contiguous()
        prm = InferParams(float(conf), float(iou), int(max_det), int(bool(agnostic_nms)),
                          class_mask.data_ptr() if class_mask is not None else None)
        st = torch.cuda.current_stream(self.device).cuda_stream
        check(self._lib.ypb_infer(self._h, C.c_void_p(st), _ptr(frames), _ptr(xform), C.byref(prm), _ptr(self.det),
                                  _ptr(self.det_lb), _ptr(self.keep), _ptr(self.coef), _ptr(self.count)))

    def ops(self):
        """[(name, kind, flops, bytes)] of the planned layer ops (+ decode_filter, nms)."""
        out = []
        for i in range(self._lib.ypb_op_count(self._h)):
            name, kind, fl, by = C.c_char_p(), C.c_int(), C.c_double(), C.c_double()
            check(self._lib.ypb_op_info(self._h, i, C.byref(name), C.byref(kind), C.byref(fl), C.byref(by)))
            out.append((name.value.decode(), kind.value, fl.value, by.value))
        return out

    def infer_profile(self, frames, xform, conf=0.25, iou=0.7, max_det=MAX_DET):
        """Like infer() but returns the device milliseconds of every op (CUDA events around each launch)."""
        n = self._lib.ypb_op_count(self._h)
        ms = (C.c_float * n)()
        prm = InferParams(float(conf), float(iou), int(max_det), 0, None)
        st = torch.cuda.current_stream(self.device).cuda_stream
        check(self._lib.ypb_infer_profile(self._h, C.c_void_p(st), _ptr(frames), _ptr(xform), C.byref(prm),
                                          _ptr(self.det), _ptr(self.det_lb), _ptr(self.keep), _ptr(self.coef),
                                          _ptr(self.count), ms, n))
        return list(ms)

    def use_outputs(self, i):
        """Select which output set the next infer()/masks() calls write/read."""
        o = self.out_sets[i]
        self.det, self.det_lb, self.keep, self.coef, self.count = o.det, o.det_lb, o.keep, o.coef, o.count
        self._cur_out = o

    def candidate_counts(self):
        """Per-image candidate counts of the last infer() (diagnostics)."""
        off, cs = C.c_size_t(), C.c_int()
        check(self._lib.ypb_select_info(self._h, C.byref(off), C.byref(cs)))
        start = self._ws_off + off.value
        return self.ws[start:start + 4 * self.shape[0]].view(torch.int32)

    def proto_view(self):
        """The (B, H/4, W/4, 32) fp32 proto buffer of the last infer(), as a view of the workspace."""
        off, nbytes = C.c_size_t(), C.c_size_t()
        check(self._lib.ypb_proto_info(self._h, C.byref(off), C.byref(nbytes)))
        start = self._ws_off + off.value
        return self.ws[start:start + nbytes.value]

    def masks(self, out, retina, out_h=0, out_w=0, proto=None, outputs=None, stream=None):
        """out: cuda uint8 (capacity, h, w).  Decodes masks of an infer() in detection order.  proto / outputs: a copy
        of that pass's proto buffer and its OutputSet when the workspace has already moved on to the next pass."""
        o = outputs or self._cur_out
        st = (stream or torch.cuda.current_stream(self.device)).cuda_stream
        check(self._lib.ypb_masks_ex(self._h, C.c_void_p(st), int(bool(retina)), int(out_h), int(out_w), _ptr(o.det),
                                     _ptr(o.det_lb), _ptr(o.coef), _ptr(o.count), _ptr(out), int(out.shape[0]),
                                     _ptr(self.mask_status), _ptr(proto), _ptr(o.offsets)))

    def device_error(self):
        w = C.c_uint32()
        check(self._lib.ypb_device_error(self._h, C.byref(w)))
        return w.value

    def device_error_async(self, stream, host_word):
        """Copy the error word into `host_word` (pinned int32 tensor, >= 1 element) in `stream` order; no host sync."""
        check(self._lib.ypb_device_error_async(self._h, C.c_void_p(stream.cuda_stream), C.c_void_p(host_word.data_ptr())))

    def set_conv_impl(self, impl):
        check(self._lib.ypb_set_conv_impl(self._h, int(impl)))

    def set_graph(self, on):
        check(self._lib.ypb_set_graph(self._h, int(bool(on))))

    # ------------------------------------------------------------------ introspection
    def view_table(self):
        out = {}
        for i in range(self._lib.ypb_view_count(self._h)):
            name, off = C.c_char_p(), C.c_size_t()
            H, W, Ct, co, Cc, dt = (C.c_int() for _ in range(6))
            check(self._lib.ypb_view_info(self._h, i, C.byref(name), C.byref(off), C.byref(H), C.byref(W), C.byref(Ct),
                                          C.byref(co), C.byref(Cc), C.byref(dt)))
            out[name.value.decode()] = (off.value, H.value, W.value, Ct.value, co.value, Cc.value, dt.value)
        return out

    def view(self, name):
        """Activation `name` ('model.4', 'proto', 'head', ...) as a (B,H,W,C) torch view of the workspace."""
        if not self._views:
            self._views = self.view_table()
        off, H, W, Ct, co, Cc, dt = self._views[name]
        B = self.shape[0]
        esz, dtype = (4, torch.float32) if dt else (2, torch.bfloat16)
        start = self._ws_off + off
        flat = self.ws[start:start + B * H * W * Ct * esz].view(dtype)
        return flat.view(B, H, W, Ct)[..., co:co + Cc]


# ---------------------------------------------------------------------------------------------------
# stand-alone kernels (parity tests)
# ---------------------------------------------------------------------------------------------------
def conv2d_bf16(x, w_gemm, bias, k, stride, act, cin=None, in_c_off=0, res=None, out=None, out_c_off=0, out_fp32=False,
                impl=0):
    """x: cuda bf16 (B,H,W,Ctot); w_gemm: cuda bf16 (k*k, cout, cin); bias: cuda fp32 (cout)."""
    B, H, W, ctot = x.shape
    cin = cin or w_gemm.shape[2]
    cout = w_gemm.shape[1]
    oH, oW = H // stride, W // stride
    if out is None:
        out = torch.zeros((B, oH, oW, cout), dtype=torch.float32 if out_fp32 else torch.bfloat16, device=x.device)
    st = torch.cuda.current_stream(x.device).cuda_stream
    handle = lib() if impl == 0 else diag_lib()  # impl 1-3: debugging twins, libypb200_diag.so only
    check(handle.ypb_conv2d_bf16(C.c_void_p(st), _ptr(x), B, H, W, ctot, in_c_off, cin, _ptr(w_gemm), _ptr(bias), cout, k,
                                stride, int(act), _ptr(res), _ptr(out), out.shape[3], out_c_off, int(out_fp32), impl))
    return out


def nms(boxes, scores, cls, n_valid, iou=0.7, max_det=MAX_DET, agnostic=False):
    """boxes (B,N,4) xyxy fp32, scores (B,N), cls (B,N) int32, n_valid (B) int32 -> keep (B,max_det) int32, count (B)."""
    B, N = scores.shape
    scratch = torch.zeros(lib().ypb_nms_scratch_bytes(B, N), dtype=torch.uint8, device=boxes.device)
    keep = torch.full((B, max_det), -1, dtype=torch.int32, device=boxes.device)
    count = torch.zeros((B,), dtype=torch.int32, device=boxes.device)
    st = torch.cuda.current_stream(boxes.device).cuda_stream
    check(lib().ypb_nms(C.c_void_p(st), _ptr(boxes), _ptr(scores), _ptr(cls), _ptr(n_valid), B, N, float(iou), max_det,
                        int(agnostic), _ptr(scratch), _ptr(keep), _ptr(count)))
    return keep, count


def gemm_weight(w):
    """(cout, cin, k, k) fp32 conv weight -> bf16 implicit-GEMM layout (k*k, cout, cin)."""
    cout, cin, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin).contiguous().to(torch.bfloat16)
