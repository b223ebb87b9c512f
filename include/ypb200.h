/* ypb200 — C ABI of the B200-native (sm_100a) YOLO-seg / YOLOv10 per-frame detector engine.
 *
 * The reference (daisy9542/yolo-puncture) has no FFI of its own: its seam is the Python API of the
 * third-party `ultralytics` package,
 *     model = YOLO(weights)                                   yolo_seg/app.py:45, yolo_with_deva.py:226
 *     results = model.predict(source=frame, conf=..., retina_masks=True, device=...)
 *                                                             yolo_seg/app.py:49,91; yolo_with_deva.py:51;
 *                                                             dev_tools/auto_speed_calc.py:62
 * and everything below `predict` (letterboxed frame -> fused Conv/BN/SiLU network -> DFL decode ->
 * NMS / top-k -> proto-mask decode) is what this library replaces.
 * yolo_puncture_b200/model.py binds these entry points with ctypes and rebuilds the unchanged
 * `YOLO.predict() -> list[Results]` surface on top (INTEGRATION.md shows the stub).
 *
 * Conventions: every entry point returns 0 on success or a negative ypb_status; nothing throws across
 * the boundary; ypb_last_error() returns a thread-local message.  An engine is single-threaded and
 * bound to one CUDA device (mirrors upstream's `with self._lock:` around inference); use one engine
 * per GPU.  All device work is enqueued on the caller's stream; there is no hidden synchronisation
 * and, after ypb_finalize_weights(), no hidden allocation: the activation workspace and every output
 * buffer are allocated by the caller (PyTorch tensors).  There is no CPU fallback: on a device that is
 * not sm_100 ypb_bind_workspace() fails.
 */
#ifndef YPB200_H_
#define YPB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ypb_engine ypb_engine;

typedef enum {
  YPB_OK = 0,
  YPB_ERR_ARG = -1,        /* bad argument / unknown model spec / shape mismatch */
  YPB_ERR_STATE = -2,      /* call order violated (weights missing, not planned, ...) */
  YPB_ERR_CUDA = -3,       /* CUDA runtime / driver error, or not an sm_100 device */
  YPB_ERR_DEVICE = -4      /* a kernel flagged an internal pipeline error */
} ypb_status;

/* ---- library ---------------------------------------------------------------------------- */
int ypb_version(void);
int ypb_is_diag_build(void); /* 0 for libypb200.so; 1 for libypb200_diag.so (debugging twins + micro-benchmarks built in) */
const char* ypb_last_error(void);

/* ---- engine lifetime: replaces YOLO(path) + AutoBackend(fuse=True) ----------------------- */
/* model_spec: "yolov8{n,s,m,l,x}-seg" | "yolov10n" | "yolo11{n,s,m,l,x}-seg"; nc = number of classes. */
int ypb_engine_create(const char* model_spec, int nc, ypb_engine** out);
void ypb_engine_destroy(ypb_engine* e);

/* Weight table in upstream state_dict naming, e.g. "model.0.conv.weight",
 * "model.0.bn.running_var", "model.22.cv2.0.2.bias".  used==0: present in checkpoints but not on the
 * inference path (v10 one-to-many branches). */
int ypb_weight_count(const ypb_engine* e);
int ypb_weight_info(const ypb_engine* e, int index, const char** name, int* ndim, int64_t shape[4], int* used);
/* Copies `numel` fp32 values from host memory; the caller keeps ownership of `data`. */
int ypb_load_weight(ypb_engine* e, const char* name, const float* data, int64_t numel);
/* Folds BatchNorm (w' = w*g/sqrt(v+1e-3), b' = beta - mu*g/sqrt(v+1e-3)), converts to bf16, repacks to the
 * implicit-GEMM layout [tap][Cout][Cin] and uploads to `device`.  Allocates the (engine-owned) weight arena. */
int ypb_finalize_weights(ypb_engine* e, int device);

/* ---- planning: static per-(B,H,W) layer plan, one workspace arena ------------------------ */
/* H, W: letterboxed network input size (multiples of 32).  Returns the workspace size the caller must
 * allocate (device memory, 1024-byte aligned) and pass to ypb_bind_workspace(). */
int ypb_plan(ypb_engine* e, int batch, int height, int width, size_t* workspace_bytes);
int ypb_bind_workspace(ypb_engine* e, void* workspace, size_t bytes);
int ypb_num_anchors(const ypb_engine* e);      /* A for the current plan */
int ypb_num_classes(const ypb_engine* e);
int ypb_num_mask_coefs(const ypb_engine* e);   /* 32 for -seg models, 0 otherwise */
int ypb_kernel_launches(const ypb_engine* e);  /* kernels one ypb_infer() enqueues */
double ypb_conv_flops(const ypb_engine* e);    /* algorithmic conv FLOPs of one ypb_infer() (whole batch) */

/* ---- inference: replaces BasePredictor.inference + postprocess ---------------------------- */
typedef struct {
  float conf;          /* predict(conf=...)      default 0.25 */
  float iou;           /* predict(iou=...)       default 0.7  */
  int max_det;         /* predict(max_det=...)   default 300 (<= 300) */
  int agnostic_nms;    /* predict(agnostic_nms=) default 0 */
  const uint32_t* class_mask; /* device ptr to ceil(nc/32) words, bit c set = class c allowed; NULL = all */
} ypb_infer_params;

/* frames : device, (B,H,W,3) uint8 BGR, already letterboxed (resize + 114 pad) to the planned H x W
 * xform  : device, (B,5) fp32 [pad_w, pad_h, gain, W0, H0] per frame (ops.scale_boxes geometry)
 * det    : device, (B,300,6) fp32 [x1,y1,x2,y2,conf,cls] in original-frame pixels, descending conf
 * det_lb : device, (B,300,4) fp32 same boxes in letterboxed-input pixels
 * keep   : device, (B,300) int32 anchor index of each detection
 * coef   : device, (B,300,nm) fp32 mask coefficients (ignored when nm == 0; may be NULL)
 * count  : device, (B) int32 detections per frame (<= params->max_det)
 * The output arrays ALWAYS have YPB_MAX_DET = 300 rows per image, whatever params->max_det asks for: image b starts at
 * row b * 300 (ypb_masks reads them with the same stride); rows at or beyond count[b] are not written. */
#define YPB_MAX_DET 300
int ypb_infer(ypb_engine* e, void* cuda_stream, const uint8_t* frames, const float* xform,
              const ypb_infer_params* params, float* det, float* det_lb, int32_t* keep, float* coef, int32_t* count);

/* Profiling twin of ypb_infer(): brackets every layer op with CUDA events on the caller's stream,
 * synchronises the stream, and returns the device time of each op in milliseconds (op order =
 * ypb_op_info order; the last two entries are decode_filter and nms).  Not for the hot loop. */
int ypb_infer_profile(ypb_engine* e, void* cuda_stream, const uint8_t* frames, const float* xform,
                      const ypb_infer_params* params, float* det, float* det_lb, int32_t* keep, float* coef,
                      int32_t* count, float* op_ms, int capacity);
int ypb_op_count(const ypb_engine* e);
/* kind: 0 stem, 1 tensor-core conv, 2 upsample, 3 sppf pool, 100 decode_filter, 101 nms;
 * flops / bytes: algorithmic work of the op for the planned batch. */
int ypb_op_info(const ypb_engine* e, int index, const char** name, int* kind, double* flops, double* bytes);

/* Fused proto-mask decode for the detections of the last ypb_infer() on this engine (reads count/det/coef
 * on the device; no host sync).  Masks are packed in detection order: frame 0's detections first.
 * retina=1: ops.process_mask_native, masks (n,H0,W0) for frames that all have original size H0 x W0;
 * retina=0: ops.process_mask, masks (n,H,W) at the letterboxed input size (out_h/out_w ignored).
 * masks  : device, (capacity, out_h, out_w) uint8 {0,1}
 * status : device, int32[2] = {total detections in batch, 1 if total > capacity (excess not written)} */
int ypb_masks(ypb_engine* e, void* cuda_stream, int retina, int out_h, int out_w, const float* det,
              const float* det_lb, const float* coef, const int32_t* count, uint8_t* masks, int capacity,
              int32_t* status);

/* Same, reading the prototypes from `proto` (a copy of the engine's (B,H/4,W/4,32) fp32 proto buffer taken after the
 * matching ypb_infer()) and using `offsets_scratch` (device, B+1 int32) instead of workspace scratch, so that the next
 * ypb_infer() may already be running on the workspace.  NULL for either = the engine's own buffers. */
int ypb_masks_ex(ypb_engine* e, void* cuda_stream, int retina, int out_h, int out_w, const float* det,
                 const float* det_lb, const float* coef, const int32_t* count, uint8_t* masks, int capacity,
                 int32_t* status, const float* proto, int32_t* offsets_scratch);
/* Byte offset and size of the proto buffer inside the bound workspace. */
int ypb_proto_info(const ypb_engine* e, size_t* offset, size_t* bytes);
/* Diagnostics: byte offset of the (B) int32 candidate counters of the selection stage inside the workspace. */
int ypb_select_info(const ypb_engine* e, size_t* count_offset, int* cand_stride);

/* Device-side error word (0 = ok); nonzero means a bounded pipeline wait inside a kernel gave up. */
int ypb_device_error(ypb_engine* e, uint32_t* word);
/* The same word copied to `host_word` (page-locked) in stream order, without a host synchronisation: the caller reads it
 * after the event / stream it already waits on for the pass's results, and calls ypb_device_error() to clear a nonzero
 * word. */
int ypb_device_error_async(ypb_engine* e, void* cuda_stream, uint32_t* host_word);

/* ---- introspection for tests / profiling --------------------------------------------------- */
/* Named activation views of the bound workspace, e.g. "model.4", "head", "proto".
 * dtype: 0 = bf16, 1 = fp32.  Layout is (B, H, W, Ctot) with the view on channels [c_off, c_off+C). */
int ypb_view_count(const ypb_engine* e);
int ypb_view_info(const ypb_engine* e, int index, const char** name, size_t* offset, int* H, int* W, int* Ctot,
                  int* c_off, int* C, int* dtype);
/* 0: persistent tcgen05 tensor-core convs (the product path, the only implementation in libypb200.so).  1 (CUDA-core
 * twin), 2 (first-generation one-tile-per-CTA tcgen05 kernel) and 3 (halo experiment) exist only in libypb200_diag.so
 * (include/ypb200_diag.h); the product library rejects them. */
int ypb_set_conv_impl(ypb_engine* e, int impl);
/* ypb_infer() replays a CUDA graph captured on first use of an argument set (default on); 0 = plain launches. */
int ypb_set_graph(ypb_engine* e, int on);

/* Stand-alone kernel entry points used by the parity tests (device pointers, caller's stream). */
int ypb_conv2d_bf16(void* cuda_stream, const void* in_nhwc_bf16, int B, int H, int W, int in_ctot, int in_c_off,
                    int cin, const void* w_gemm_bf16 /*[k*k][cout][cin]*/, const float* bias, int cout, int k,
                    int stride, int act, const void* res_nhwc_bf16 /*nullable, same layout as out*/, void* out,
                    int out_ctot, int out_c_off, int out_fp32, int impl);
int ypb_nms(void* cuda_stream, const float* boxes_xyxy /*(B,N,4)*/, const float* scores /*(B,N)*/,
            const int32_t* cls /*(B,N)*/, const int32_t* n_valid /*(B)*/, int B, int N, float iou, int max_det,
            int agnostic, void* scratch /* >= B*nextpow2(N)*8 + B*4 bytes */, int32_t* keep /*(B,max_det)*/,
            int32_t* count /*(B)*/);
size_t ypb_nms_scratch_bytes(int B, int N);
/* Needle length on the device (replaces D2H of the mask + cv2.findContours + cv2.minAreaRect of reference
   yolo_seg/app.py:97-105, utils/mask_tools.py:12-22): for every (H,W) uint8 {0,1} mask the long side of its
   minimum-area bounding rectangle and long / max(short, 1).  row_extents: scratch of n*H*2 int32.  Device pointers. */
int ypb_mask_min_rect(void* cuda_stream, const uint8_t* masks, int n, int H, int W, int32_t* row_extents, float* out);
/* LetterBox on the device (UPSTREAM data/augment.py::LetterBox for frames at least as large as the network input):
   src (B,H0,W0,3) uint8 -> dst (B,H,W,3): cv2.INTER_LINEAR resize to (new_h,new_w) pasted at (top,left), rest = pad_value.
   xofs/yofs (new_w / new_h int32 source indices) and xa/ya (2 int16 11-bit coefficients per output column / row) are
   built by the caller exactly as OpenCV builds them (model.py::cv2_linear_tables); device pointers, caller's stream. */
int ypb_letterbox_u8(void* cuda_stream, const uint8_t* src, int B, int H0, int W0, uint8_t* dst, int H, int W, int new_w,
                     int new_h, int top, int left, const int32_t* xofs, const int16_t* xa, const int32_t* yofs,
                     const int16_t* ya, int pad_value);
/* Index-mask hand-off to the tracker (replaces the per-detection loop of reference yolo_seg/yolo_with_deva.py:54-88).
   masks: (n_total, H, W) uint8 {0,1} in detection order, frame b owning rows [offsets[b], offsets[b+1]) (offsets on the
   device).  area (n_total) receives the pixel count of every mask, ids (n_total) the 1-based id of every kept detection
   inside its frame (0 = suppressed: area < min_area; min_area < 0 keeps everything), index_map (B, H, W) int64 the id
   of the LAST kept detection covering each pixel, else 0.  All device pointers, caller's stream. */
int ypb_index_masks(void* cuda_stream, const uint8_t* masks, const int32_t* offsets, int B, int n_total, int H, int W,
                    int min_area, int32_t* area, int32_t* ids, int64_t* index_map);
/* The same hand-off for masks known to be zero outside a rectangle each (predict() crops every mask to its box):
 * rects (n_total, 4) int32 device array of (x0, y0, x1, y1), x1 / y1 exclusive.  The area sum and the paint then touch
 * ~the box areas instead of n_total x H x W bytes twice.  rects == NULL (or W % 16 != 0) falls back to ypb_index_masks. */
int ypb_index_masks_boxed(void* cuda_stream, const uint8_t* masks, const int32_t* offsets, const int32_t* rects, int B,
                          int n_total, int H, int W, int min_area, int32_t* area, int32_t* ids, int64_t* index_map);
/* The same hand-off for the reference's `min_side` branch (yolo_with_deva.py:44-48,71-72): predict() ran on a resized
   frame, masks are (n_total, h1, w1) and go back to (H, W) as torchvision's F.resize does (antialiased bilinear) before
   the float `mask.sum() < min_area` filter (min_area < 0: keep all) and the `mask > 0.5` paint.
   bins: scratch (n_total, H, W) uint8; area_f: (n_total) fp32 sums of the resized masks. */
int ypb_index_masks_resized(void* cuda_stream, const uint8_t* masks, const int32_t* offsets, int B, int n_total, int h1,
                            int w1, int H, int W, float min_area, uint8_t* bins, float* area_f, int32_t* ids,
                            int64_t* index_map);
/* JPEG frames decoded straight into device memory (replaces the host decode of reference yolo_seg/utils/video_reader.py:
   91-99 `Image.open(path).convert("RGB")` and of `cv2.imread`): jpeg_host = the file's bytes in host memory; dst_dev =
   device (height, width, 3) uint8 interleaved BGR, the layout ypb_letterbox_u8 / ypb_infer consume.  nvJPEG (loaded
   lazily; YPB_ERR_CUDA if the machine has no libnvjpeg).  Enqueued on the caller's stream after a host-side Huffman pass. */
int ypb_jpeg_info(const uint8_t* jpeg_host, size_t nbytes, int* height, int* width);
int ypb_jpeg_decode_bgr(void* cuda_stream, const uint8_t* jpeg_host, size_t nbytes, uint8_t* dst_dev, int height, int width);
/* Point-to-point mask hand-off between the GPUs of one box (BASELINE config C5): the tracker that consumes
   the index masks (reference yolo_seg/yolo_with_deva.py:133-159) runs on ONE GPU; detector replicas on the other GPUs
   copy their index masks into a mailbox in that GPU's memory over NVLink / NVSwitch.  create: cudaMalloc on `device` +
   CUDA IPC handle (64 bytes) to hand to the producer processes; open / close: map it in a producer process;
   ypb_peer_copy: cudaMemcpyAsync(dst, src) on the caller's stream (peer access is enabled by the IPC open). */
int ypb_mailbox_create(int device, size_t bytes, void** dev_ptr, unsigned char handle[64]);
int ypb_mailbox_destroy(int device, void* dev_ptr);
int ypb_mailbox_open(int device, const unsigned char handle[64], void** dev_ptr);
int ypb_mailbox_close(int device, void* dev_ptr);
int ypb_peer_copy(void* cuda_stream, void* dst, const void* src, size_t bytes);
/* Zero-staging path of predict(): is this host pointer page-locked (cudaHostAlloc / cudaHostRegister / pinned torch
   tensor)?  and: copy n such frames to consecutive device slots on `cuda_stream`, merging adjacent sources. */
int ypb_host_is_pinned(const void* p, int* pinned);
int ypb_hosts_are_pinned(const void* const* ptrs, int n, int* all_pinned);  /* the same for n pointers: 1 iff all are */
int ypb_h2d_frames(void* cuda_stream, void* dst_dev, const void* const* src, size_t bytes_each, int n);
/* Host helper of the predict() pipeline: copy n frames into pinned staging memory with nthreads host threads. */
int ypb_stage_frames(void* const* dst, const void* const* src, const size_t* bytes, int n, int nthreads);
/* Same with the copy flavour chosen by the caller: mode 0 = memcpy, 1 = non-temporal stores (the default of
   ypb_stage_frames unless YPB_STAGE_MEMCPY is set).  Persistent thread pool, 256 KB pieces. */
int ypb_stage_frames_ex(void* const* dst, const void* const* src, const size_t* bytes, int n, int nthreads, int mode);
/* Asynchronous flavour: queues the copy on the staging pool, returns at once, and gates `cuda_stream` (a host function):
   whatever is enqueued on that stream afterwards - the chunk's H2D copy - runs only once the frames are staged.  The
   caller keeps enqueuing engine passes meanwhile; sources and destinations must outlive the gate. */
int ypb_stage_frames_gated(void* cuda_stream, void* const* dst, const void* const* src, const size_t* bytes, int n, int nthreads);
#ifdef __cplusplus
}
#endif
#endif /* YPB200_H_ */
