// Detection-head selection kernels (HBM / latency bound, warp-shuffle + shared memory):
//   decode_filter_kernel : DFL softmax-expectation box decode + class max + conf filter + candidate push
//   nms_kernel           : batched (one CTA per image) sort + greedy class-aware NMS + scale_boxes + coef gather
// UPSTREAM sites replaced: head.py::Detect._inference (DFL, dist2bbox, sigmoid), utils/ops.py::
// non_max_suppression (+ torchvision.ops.nms), ops.scale_boxes / clip_boxes  (SURVEY.md §8 a6, a7, a9);
// all of it runs inside `model.predict(...)` at reference yolo_seg/app.py:91 and yolo_with_deva.py:51.
//
// Bit-exactness contract (tests/test_gpu_select.py): fed the oracle's decoded boxes/scores, the kept
// anchor indices and class ids are identical to torchvision's.  Hence: IEEE *_rn ops in the oracle's
// order (no FMA contraction), strict `>` on IoU, inter/(a+b-inter) with no +1, fp32 class offset
// box + cls*7680 applied before IoU, ordering by (score desc, anchor index asc).
#pragma once
#include "common.cuh"

namespace ypb {

struct HeadGeom {
  int A;              // anchors per image
  int no;             // floats per anchor row in the head buffer: 64 + nc + nm
  int nc, nm;
  int cand_stride;    // slots per image in the candidate key list (next_pow2(A): room for sort padding)
  int lvl_start[4];   // first anchor of each level, [3] = A
  int lvl_w[3];       // map width per level
  float lvl_stride[3];
};

__device__ __forceinline__ float sigmoid_f(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// One warp per anchor.  head: (B, A, no) fp32 rows [64 box logits | nc class logits | nm coefs].
// Candidates (max class score > conf) get their decoded xyxy box and class written at the dense
// slot [b][anchor] and a 64-bit sort key pushed on the image's candidate list.
// key = score_bits << 32 | ~anchor   (descending key order == score desc, anchor asc)
__global__ void __launch_bounds__(256)
decode_filter_kernel(const float* __restrict__ head, HeadGeom g, int nB, float conf, int xyxy_direct,
                     const unsigned* __restrict__ cls_mask, float4* __restrict__ dbox, int* __restrict__ dcls, unsigned long long* __restrict__ cand_keys,
                     int* __restrict__ cand_count) {
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= (long long)nB * g.A) return;
  const int b = (int)(wid / g.A), a = (int)(wid - (long long)b * g.A);
  const float* row = head + wid * g.no;

  // ---- class max / argmax (first index wins ties, like torch.max) ----
  float best = -INFINITY;
  int bidx = 0x7fffffff;
  for (int c = lane; c < g.nc; c += 32) {
    const float v = row[64 + c];
    if (v > best) { best = v; bidx = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
  }
  const float score = sigmoid_f(best);
  if (!(score > conf)) return;  // warp-uniform
  // predict(classes=[...]): upstream filters on the arg-max class after the confidence test
  if (!xyxy_direct && cls_mask != nullptr && !((cls_mask[bidx >> 5] >> (bidx & 31)) & 1u)) return;

  // ---- DFL: softmax over 16 bins per side, expectation; lane l holds side l/16 (and +2) bin l%16 ----
  float d[2];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const float x = row[half * 32 + lane];
    float m = x;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float e = expf(x - m);
    float s = e, ws = e * (float)(lane & 15);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      ws += __shfl_xor_sync(0xffffffffu, ws, o);
    }
    d[half] = __fdiv_rn(ws, s);
  }
  // lanes 0..15: (l, r)   lanes 16..31: (t, b)
  const float dl = __shfl_sync(0xffffffffu, d[0], 0), dt = __shfl_sync(0xffffffffu, d[0], 16);
  const float dr = __shfl_sync(0xffffffffu, d[1], 0), db = __shfl_sync(0xffffffffu, d[1], 16);
  if (lane != 0) return;

  int lvl = 0;
  if (a >= g.lvl_start[1]) lvl = 1;
  if (a >= g.lvl_start[2]) lvl = 2;
  const int local = a - g.lvl_start[lvl];
  const int iy = local / g.lvl_w[lvl], ix = local - iy * g.lvl_w[lvl];
  const float ax = (float)ix + 0.5f, ay = (float)iy + 0.5f, st = g.lvl_stride[lvl];
  const float x1 = __fsub_rn(ax, dl), y1 = __fsub_rn(ay, dt), x2 = __fadd_rn(ax, dr), y2 = __fadd_rn(ay, db);
  float4 box;
  if (xyxy_direct) {  // end2end heads: dist2bbox(xywh=False) * stride
    box = make_float4(__fmul_rn(x1, st), __fmul_rn(y1, st), __fmul_rn(x2, st), __fmul_rn(y2, st));
  } else {            // dist2bbox(xywh=True) * stride, then ops.xywh2xyxy inside NMS
    const float cx = __fmul_rn(__fdiv_rn(__fadd_rn(x1, x2), 2.0f), st), cy = __fmul_rn(__fdiv_rn(__fadd_rn(y1, y2), 2.0f), st);
    const float w = __fmul_rn(__fsub_rn(x2, x1), st), h = __fmul_rn(__fsub_rn(y2, y1), st);
    const float hw = __fdiv_rn(w, 2.0f), hh = __fdiv_rn(h, 2.0f);
    box = make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
  }
  dbox[wid] = box;
  dcls[wid] = bidx;
  const int slot = atomicAdd(cand_count + b, 1);
  cand_keys[(long long)b * g.cand_stride + slot] =
      ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)a);
}

// Same contract as decode_filter_kernel for nc % 4 == 0: a warp scans EIGHT anchors at a time, four lanes per anchor,
// each lane streaming its share of the class logits with independent 16-byte loads (the one-warp-per-anchor version
// has three dependent-latency loads in flight per warp and reaches ~1 TB/s); the rare candidates (score > conf) are
// then decoded one after the other by the whole warp with exactly the arithmetic of decode_filter_kernel.
__global__ void __launch_bounds__(256)
decode_filter8_kernel(const float* __restrict__ head, HeadGeom g, int nB, float conf, int xyxy_direct,
                      const unsigned* __restrict__ cls_mask, float4* __restrict__ dbox, int* __restrict__ dcls,
                      unsigned long long* __restrict__ cand_keys, int* __restrict__ cand_count) {
  const int lane = threadIdx.x & 31, grp = lane >> 2, sub = lane & 3;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long total = (long long)nB * g.A;
  const long long a0 = warp_id * 8;
  if (a0 >= total) return;
  const long long mine = a0 + grp;
  float best = -INFINITY;
  int bidx = 0x7fffffff;
  // sigmoid(x) > conf  =>  x > logit(conf); a margin keeps the exact float comparison on sigmoid_f authoritative
  const float pair_logit_lo = (conf > 0.f && conf < 1.f) ? __logf(conf / (1.0f - conf)) - 0.05f : -INFINITY;
  if (mine < total) {
    const float4* cl = reinterpret_cast<const float4*>(head + mine * g.no + 64);
    const int n4 = g.nc >> 2;
    // eight independent 16-byte loads per lane in flight (nc <= 128: the whole row in one batch), then the scan in
    // ascending class order: the kernel is latency-bound on these loads (ncu: 21 warps per issue on long scoreboard)
    for (int base = 0; base < n4; base += 32) {
    float4 vq[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c4 = base + sub + 4 * u;
      vq[u] = c4 < n4 ? __ldg(cl + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u8 = 0; u8 < 8; ++u8) {
      const int c4 = base + sub + 4 * u8;
      if (c4 >= n4) break;
      const float4 v = vq[u8];
      const int c = c4 * 4;
      if (v.x > best) { best = v.x; bidx = c; }
      if (v.y > best) { best = v.y; bidx = c + 1; }
      if (v.z > best) { best = v.z; bidx = c + 2; }
      if (v.w > best) { best = v.w; bidx = c + 3; }
      if (xyxy_direct) {
        // end-to-end heads: every (anchor, class) pair with score > conf is a top-k candidate (what
        // pair_candidates_kernel does in a second pass over the head); logits far below the threshold skip the sigmoid
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (!(vv[u] > pair_logit_lo)) continue;
          const float sc = sigmoid_f(vv[u]);
          if (!(sc > conf)) continue;
          const int cc = c + u;  // predict(classes=[...]) filters AFTER the top-k / conf / max_det cut (nms_kernel, e2e)
          const int bimg = (int)(mine / g.A), an = (int)(mine - (long long)bimg * g.A);
          const int slot = atomicAdd(cand_count + bimg, 1);
          if (slot < g.cand_stride)
            cand_keys[(long long)bimg * g.cand_stride + slot] =
                ((unsigned long long)__float_as_uint(sc) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(an * g.nc + cc));
          else
            atomicOr(&g_dev_error, 0x100u);  // candidate list overflow (conf far below any sensible value)
        }
      }
    }
    }
  }
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {  // first index wins ties, like torch.max
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
  }
  const float score = sigmoid_f(best);
  bool cand = (mine < total) && (score > conf);
  // predict(classes=[...]): upstream filters on the arg-max class after the confidence test
  if (cand && !xyxy_direct && cls_mask != nullptr && !((cls_mask[bidx >> 5] >> (bidx & 31)) & 1u)) cand = false;
  unsigned todo = __ballot_sync(0xffffffffu, cand && sub == 0);
  while (todo != 0u) {  // warp-uniform
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const long long wid = a0 + (src >> 2);
    const float c_score = __shfl_sync(0xffffffffu, score, src);
    const int c_idx = __shfl_sync(0xffffffffu, bidx, src);
    const int b = (int)(wid / g.A), a = (int)(wid - (long long)b * g.A);
    const float* row = head + wid * g.no;
    // ---- DFL: softmax over 16 bins per side, expectation; lane l holds side l/16 (and +2) bin l%16 ----
    float d[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const float x = row[half * 32 + lane];
      float m = x;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      const float e = expf(x - m);
      float s = e, ws = e * (float)(lane & 15);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ws += __shfl_xor_sync(0xffffffffu, ws, o);
      }
      d[half] = __fdiv_rn(ws, s);
    }
    const float dl = __shfl_sync(0xffffffffu, d[0], 0), dt = __shfl_sync(0xffffffffu, d[0], 16);
    const float dr = __shfl_sync(0xffffffffu, d[1], 0), db = __shfl_sync(0xffffffffu, d[1], 16);
    if (lane == 0) {
      int lvl = 0;
      if (a >= g.lvl_start[1]) lvl = 1;
      if (a >= g.lvl_start[2]) lvl = 2;
      const int local = a - g.lvl_start[lvl];
      const int iy = local / g.lvl_w[lvl], ix = local - iy * g.lvl_w[lvl];
      const float ax = (float)ix + 0.5f, ay = (float)iy + 0.5f, st = g.lvl_stride[lvl];
      const float x1 = __fsub_rn(ax, dl), y1 = __fsub_rn(ay, dt), x2 = __fadd_rn(ax, dr), y2 = __fadd_rn(ay, db);
      float4 box;
      if (xyxy_direct) {
        box = make_float4(__fmul_rn(x1, st), __fmul_rn(y1, st), __fmul_rn(x2, st), __fmul_rn(y2, st));
      } else {
        const float cx = __fmul_rn(__fdiv_rn(__fadd_rn(x1, x2), 2.0f), st), cy = __fmul_rn(__fdiv_rn(__fadd_rn(y1, y2), 2.0f), st);
        const float w = __fmul_rn(__fsub_rn(x2, x1), st), h = __fmul_rn(__fsub_rn(y2, y1), st);
        const float hw = __fdiv_rn(w, 2.0f), hh = __fdiv_rn(h, 2.0f);
        box = make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
      }
      dbox[wid] = box;
      dcls[wid] = c_idx;
      if (!xyxy_direct) {  // v8 / 11: one candidate per anchor (end-to-end heads pushed their (anchor, class) pairs above)
        const int slot = atomicAdd(cand_count + b, 1);
        cand_keys[(long long)b * g.cand_stride + slot] =
            ((unsigned long long)__float_as_uint(c_score) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)a);
      }
    }
  }
}

// End-to-end (YOLOv10, NMS-free) heads: every (anchor, class) pair with score > conf is a candidate of the
// top-k; key = score_bits << 32 | ~(anchor*nc + class).  Runs after decode_filter_kernel(xyxy_direct=1), which
// has already written the xyxy box of every anchor whose best class passes conf.  One warp per anchor.
__global__ void __launch_bounds__(256)
pair_candidates_kernel(const float* __restrict__ head, HeadGeom g, int nB, float conf, const unsigned* __restrict__ cls_mask,
                       unsigned long long* __restrict__ cand_keys, int* __restrict__ cand_count) {
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= (long long)nB * g.A) return;
  const int b = (int)(wid / g.A), a = (int)(wid - (long long)b * g.A);
  const float* row = head + wid * g.no;
  for (int c = lane; c < g.nc; c += 32) {
    const float score = sigmoid_f(row[64 + c]);
    if (!(score > conf)) continue;
    (void)cls_mask;  // predict(classes=[...]) filters AFTER the top-k / conf / max_det cut (nms_kernel, e2e)
    const int slot = atomicAdd(cand_count + b, 1);
    if (slot < g.cand_stride)
      cand_keys[(long long)b * g.cand_stride + slot] =
          ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(a * g.nc + c));
    else
      atomicOr(&g_dev_error, 0x100u);  // candidate list overflow (conf far below any sensible value)
  }
}

// In-place bitonic sort, descending, of n_pow2 keys (any address space), by one CTA.
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* k, int n_pow2) {
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < (n_pow2 >> 1); i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = k[lo], c = k[hi];
        if ((a < c) == desc) { k[lo] = c; k[hi] = a; }
      }
    }
  }
  __syncthreads();
}

struct FrameXform {   // per-image letterbox undo (ops.scale_boxes): x = clamp((x - pad) / gain, 0, size)
  float pad_w, pad_h, gain, W0, H0;
};

__device__ __forceinline__ float iou_tv(float ix1, float iy1, float ix2, float iy2, float iarea, float jx1, float jy1,
                                        float jx2, float jy2, float jarea) {
  const float xx1 = fmaxf(ix1, jx1), yy1 = fmaxf(iy1, jy1), xx2 = fminf(ix2, jx2), yy2 = fminf(iy2, jy2);
  const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1)), h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter));
}

// iou_tv(...) > thr with an exact early-out: boxes that do not overlap have inter = 0, IoU = 0 (or NaN for two
// zero-area boxes) and are never suppressed; with the per-class offset that is nearly every pair, and it skips the
// IEEE division.
__device__ __forceinline__ bool iou_gt(float ix1, float iy1, float ix2, float iy2, float iarea, float jx1, float jy1,
                                       float jx2, float jy2, float jarea, float thr) {
  const float xx1 = fmaxf(ix1, jx1), yy1 = fmaxf(iy1, jy1), xx2 = fminf(ix2, jx2), yy2 = fminf(iy2, jy2);
  const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1)), h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
  if (!(w > 0.0f && h > 0.0f)) return thr < 0.0f;  // IoU is exactly 0 (or NaN): only a negative threshold is exceeded by 0
  const float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter)) > thr;
}

constexpr int kNmsSmemKeys = 4096;
constexpr int kNmsMaxDet = 300;

// One CTA per image.  Outputs (row i < count[b], descending score):
//   det     (B, out_stride, 6)  [x1,y1,x2,y2,conf,cls] boxes mapped back to the original frame
//   det_lb  (B, out_stride, 4)  the same boxes in letterboxed-input pixels (what non-retina masks crop with)
//   keep    (B, out_stride)     anchor index of each kept row
//   coef    (B, out_stride, nm) mask coefficients gathered from the head rows
// out_stride = rows per image of the output arrays: always kNmsMaxDet (300) for the engine, whatever max_det (<= 300)
// the call asks for, so that the mask decode and the host wrappers address image b at b * 300 rows.
constexpr int kNmsThreads = 1024;
__global__ void __launch_bounds__(kNmsThreads)
nms_kernel(const float* __restrict__ head, HeadGeom g, const float4* __restrict__ dbox, const int* __restrict__ dcls,
           unsigned long long* __restrict__ cand_keys, const int* __restrict__ cand_count, float iou_thr, int max_det,
           int max_nms, float max_wh, int e2e, const unsigned* __restrict__ cls_mask, int out_stride,
           const FrameXform* __restrict__ xf,
           float* __restrict__ det, float* __restrict__ det_lb, int* __restrict__ keep, float* __restrict__ coef,
           int* __restrict__ count) {
  __shared__ unsigned long long s_keys[kNmsSmemKeys];
  __shared__ float s_kept[kNmsMaxDet][5];  // offset box + area
  __shared__ float s_score[kNmsMaxDet];
  __shared__ int s_nkept;
  const int b = blockIdx.x;
  int n = cand_count[b];
  if (n > g.cand_stride) n = g.cand_stride;
  unsigned long long* gk = cand_keys + (long long)b * g.cand_stride;
  int n_pow2 = 1;
  while (n_pow2 < n) n_pow2 <<= 1;
  unsigned long long* keys;
  __shared__ int s_sel[2];
  __shared__ int s_path;  // profiling build: 1 class-aware scan, 2 plain scan, 0 end-to-end top-k
  if (kProf && threadIdx.x == 0) s_path = 0;
  bool selected = false;
  if (e2e && n_pow2 > kNmsSmemKeys) {
    // End-to-end top-k over more pairs than the shared-memory sort holds: only the best max_det matter, so select
    // them first.  Histogram of the scores (float bits are monotonic for positive floats) over 4096 bins, the bin
    // where the count from the top reaches max_det, then every key at or above that bin is compacted into shared
    // memory and sorted there (exactly the keys a full sort would put first, in the same order).
    int* hist = reinterpret_cast<int*>(s_keys);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) hist[i] = 0;
    unsigned lo_bits = 0xffffffffu, hi_bits = 0u;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned sb = (unsigned)(gk[i] >> 32);
      lo_bits = min(lo_bits, sb);
      hi_bits = max(hi_bits, sb);
    }
    __shared__ unsigned s_lo, s_hi;
    if (threadIdx.x == 0) { s_lo = 0xffffffffu; s_hi = 0u; }
    __syncthreads();
    atomicMin(&s_lo, lo_bits);
    atomicMax(&s_hi, hi_bits);
    __syncthreads();
    const unsigned base = s_lo;
    int shift = 0;
    while (((s_hi - base) >> shift) >= 4096u) ++shift;
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[((unsigned)(gk[i] >> 32) - base) >> shift], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      int acc = 0, t = 4095;
      for (; t > 0; --t) {
        acc += hist[t];
        if (acc >= max_det) break;
      }
      if (t == 0) acc += hist[0];
      s_sel[0] = t;
      s_sel[1] = acc;  // keys in bins >= t
    }
    __syncthreads();
    const int tbin = s_sel[0], nsel = s_sel[1];
    __syncthreads();  // everyone has read the histogram results: its memory is reused for the keys
    if (nsel <= kNmsSmemKeys) {
      if (threadIdx.x == 0) s_sel[0] = 0;
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long k = gk[i];
        if ((int)((((unsigned)(k >> 32)) - base) >> shift) >= tbin) s_keys[atomicAdd(&s_sel[0], 1)] = k;
      }
      __syncthreads();
      n = nsel;
      n_pow2 = 1;
      while (n_pow2 < n) n_pow2 <<= 1;
      for (int i = n + threadIdx.x; i < n_pow2; i += blockDim.x) s_keys[i] = 0ull;
      selected = true;
    }
  }
  if (selected) {
    keys = s_keys;
  } else if (n_pow2 <= kNmsSmemKeys) {
    keys = s_keys;
    for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) keys[i] = i < n ? gk[i] : 0ull;
  } else {  // rare: sort in place in global memory (cand_stride = next_pow2(A) leaves room for the padding)
    keys = gk;
    for (int i = n + threadIdx.x; i < n_pow2; i += blockDim.x) keys[i] = 0ull;
  }
  const long long _np0 = kProf ? clock64() : 0;
  bitonic_sort_desc(keys, n_pow2);
  const long long _np1 = kProf ? clock64() : 0;
  if (n > max_nms) n = max_nms;

  __shared__ int s_cls[kNmsMaxDet];
  if (e2e) {
    // UPSTREAM Detect.postprocess + `pred[pred[:, 4] > conf][:max_det]`: the top-max_det pairs by score, no suppression;
    // predict(classes=[...]) is applied to THOSE rows afterwards (ops.non_max_suppression, end2end branch), so a
    // filtered class frees no slot for a lower-scoring pair.  Order-preserving compaction (max_det <= blockDim.x).
    const int nk = n < max_det ? n : max_det;
    __shared__ int s_wcnt[kNmsThreads / 32];
    const int i = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long key = 0ull;
    int cls = 0;
    bool take = false;
    if (i < nk) {
      key = keys[i];
      const unsigned pair = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull);
      cls = (int)(pair % (unsigned)g.nc);
      take = cls_mask == nullptr || ((cls_mask[cls >> 5] >> (cls & 31)) & 1u) != 0u;
    }
    const unsigned tk = __ballot_sync(0xffffffffu, take);
    if (lane == 0) s_wcnt[warp] = __popc(tk);
    __syncthreads();
    int pos = __popc(tk & ((1u << lane) - 1u)), total = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      if (w < warp) pos += s_wcnt[w];
      total += s_wcnt[w];
    }
    if (take) {
      const unsigned pair = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull);
      s_score[pos] = __uint_as_float((unsigned)(key >> 32));
      s_cls[pos] = cls;
      keep[(long long)b * out_stride + pos] = (int)(pair / (unsigned)g.nc);
    }
    if (threadIdx.x == 0) { s_nkept = total; count[b] = total; }
  } else {
    // Greedy scan, blockDim.x candidates (one per thread) per round, in descending score order:
    //  A. every thread tests its candidate against the boxes kept in earlier rounds (all warps in parallel);
    //  B. the warps in turn resolve their own 32 candidates with ballots/shuffles (a live lane is kept and
    //     suppresses the later lanes it overlaps), publish the newly kept boxes, and the later warps test
    //     their candidates against just those.  Equivalent to torchvision's sequential loop.
    // 1024 threads: an image with thousands of candidates needs few rounds (phase A costs nkept IoUs per round).
    __shared__ int s_range[kNmsThreads / 32][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    int nkept = 0;  // CTA-uniform
    // Class-aware fast path.  With the per-class offset (cls * max_wh) boxes of different classes cannot overlap as
    // long as every coordinate lies in a window narrower than max_wh: then a candidate only has to be tested against
    // kept boxes of ITS class.  Kept boxes are chained per class bucket (s_head / s_next, newest first), the in-warp
    // resolution uses match.any on the class id, and the greedy order inside a warp is recovered with a ballot fixed
    // point.  Same decisions as the plain scan below (IoU is symmetric bit for bit: max/min, a*b, a+b commute) -
    // an image with thousands of same-class-sparse candidates drops from O(candidates x kept) IoU tests to a few
    // per candidate.  Class-agnostic NMS (max_wh = 0) and out-of-window coordinates take the plain scan.
    __shared__ int s_head[128];
    __shared__ short s_next[kNmsMaxDet];
    __shared__ float s_wbox[32][5];
    bool in_window = true;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int anchor = (int)(0xFFFFFFFFu - (unsigned)(keys[i] & 0xFFFFFFFFull));
      const float4 bx = dbox[(long long)b * g.A + anchor];
      in_window = in_window && bx.x > -2000.f && bx.y > -2000.f && bx.z < 5000.f && bx.w < 5000.f && bx.x <= bx.z && bx.y <= bx.w;
    }
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_head[i] = -1;
    const bool class_sep = __syncthreads_and(in_window) && max_wh >= 7680.0f && max_det <= kNmsMaxDet && iou_thr >= 0.0f;
    if (kProf && threadIdx.x == 0) s_path = class_sep ? 1 : 2;
    for (int i0 = 0; class_sep && i0 < n && nkept < max_det; i0 += blockDim.x) {
      const int i = i0 + threadIdx.x;
      bool alive = i < n;
      float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f, area = 0.f, score = 0.f;
      int anchor = 0, cls = -1 - lane;  // dead lanes: distinct pseudo classes, never matched
      if (alive) {
        const unsigned long long key = keys[i];
        anchor = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
        score = __uint_as_float((unsigned)(key >> 32));
        const float4 bx = dbox[(long long)b * g.A + anchor];
        cls = dcls[(long long)b * g.A + anchor];
        const float off = __fmul_rn((float)cls, max_wh);
        x1 = __fadd_rn(bx.x, off); y1 = __fadd_rn(bx.y, off); x2 = __fadd_rn(bx.z, off); y2 = __fadd_rn(bx.w, off);
        area = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
      }
      const int bkt = cls & 127;
      // A. kept boxes of earlier rounds, this class only
      if (alive) {
        for (int k = s_head[bkt]; k >= 0; k = s_next[k]) {
          if (s_cls[k] == cls && iou_gt(s_kept[k][0], s_kept[k][1], s_kept[k][2], s_kept[k][3], s_kept[k][4], x1, y1, x2, y2, area, iou_thr)) {
            alive = false;
            break;
          }
        }
      }
      const int nw_round = min(nwarps, (n - i0 + 31) >> 5);  // warps that hold candidates this round (CTA-uniform)
      for (int ws = 0; ws < nw_round; ++ws) {
        if (warp == ws) {
          // B. this warp's 32 candidates among themselves
          s_wbox[lane][0] = x1; s_wbox[lane][1] = y1; s_wbox[lane][2] = x2; s_wbox[lane][3] = y2; s_wbox[lane][4] = area;
          __syncwarp();
          const unsigned alive_in = __ballot_sync(0xffffffffu, alive);
          unsigned cand = __match_any_sync(0xffffffffu, cls) & ((1u << lane) - 1u) & alive_in;  // live earlier lanes of my class
          unsigned sup = 0u;  // those that overlap me beyond the threshold
          while (alive && cand != 0u) {
            const int j = __ffs(cand) - 1;
            cand &= cand - 1u;
            if (iou_gt(s_wbox[j][0], s_wbox[j][1], s_wbox[j][2], s_wbox[j][3], s_wbox[j][4], x1, y1, x2, y2, area, iou_thr)) sup |= 1u << j;
          }
          // greedy order: lane i is kept iff it is alive and no KEPT earlier lane suppresses it; iterating
          // K <- {alive, sup & K == 0} from K = alive fixes lane t after t+1 steps, and stops at the (unique) solution
          unsigned K = alive_in;
          for (;;) {
            const unsigned Kn = __ballot_sync(0xffffffffu, alive && (sup & K) == 0u);
            if (Kn == K) break;
            K = Kn;
          }
          const int rank = __popc(K & ((1u << lane) - 1u));
          const int room = max_det - nkept;
          const bool kept_here = ((K >> lane) & 1u) != 0u && rank < room;
          if (kept_here) {
            const int idx = nkept + rank;
            s_kept[idx][0] = x1; s_kept[idx][1] = y1; s_kept[idx][2] = x2; s_kept[idx][3] = y2; s_kept[idx][4] = area;
            s_score[idx] = score;
            s_cls[idx] = cls;
            keep[(long long)b * out_stride + idx] = anchor;
            s_next[idx] = (short)atomicExch(&s_head[bkt], idx);  // newest first
          }
          if (lane == 0) { s_range[ws][0] = nkept; s_range[ws][1] = nkept + min(__popc(K), room); }
        }
        __syncthreads();
        const int kb = s_range[ws][0], ke = s_range[ws][1];
        if (warp > ws && alive) {  // C. later candidates against the boxes this warp just kept (front of the chain)
          for (int k = s_head[bkt]; k >= kb; k = s_next[k]) {
            if (s_cls[k] == cls && iou_gt(s_kept[k][0], s_kept[k][1], s_kept[k][2], s_kept[k][3], s_kept[k][4], x1, y1, x2, y2, area, iou_thr)) {
              alive = false;
              break;
            }
          }
        }
        nkept = ke;
        if (nkept >= max_det) break;  // CTA-uniform
      }
      __syncthreads();  // s_range / s_wbox are rewritten next round
    }
    for (int i0 = 0; !class_sep && i0 < n && nkept < max_det; i0 += blockDim.x) {
      const int i = i0 + threadIdx.x;
      bool alive = i < n;
      float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f, area = 0.f, score = 0.f;
      int anchor = 0, cls = 0;
      if (alive) {
        const unsigned long long key = keys[i];
        anchor = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
        score = __uint_as_float((unsigned)(key >> 32));
        const float4 bx = dbox[(long long)b * g.A + anchor];
        cls = dcls[(long long)b * g.A + anchor];
        const float off = __fmul_rn((float)cls, max_wh);
        x1 = __fadd_rn(bx.x, off); y1 = __fadd_rn(bx.y, off); x2 = __fadd_rn(bx.z, off); y2 = __fadd_rn(bx.w, off);
        area = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
      }
      for (int k = 0; k < nkept; ++k) {
        if (!__any_sync(0xffffffffu, alive)) break;  // the whole warp is already suppressed (or past the end)
        if (alive && iou_gt(s_kept[k][0], s_kept[k][1], s_kept[k][2], s_kept[k][3], s_kept[k][4], x1, y1, x2, y2, area, iou_thr))
          alive = false;
      }
      const int nw_round = min(nwarps, (n - i0 + 31) >> 5);  // warps that hold candidates this round (CTA-uniform)
      for (int ws = 0; ws < nw_round; ++ws) {
        if (warp == ws) {
          int nk = nkept;
          unsigned live = __ballot_sync(0xffffffffu, alive);
          while (live != 0u && nk < max_det) {
            const int l = __ffs(live) - 1;
            const float kx1 = __shfl_sync(0xffffffffu, x1, l), ky1 = __shfl_sync(0xffffffffu, y1, l);
            const float kx2 = __shfl_sync(0xffffffffu, x2, l), ky2 = __shfl_sync(0xffffffffu, y2, l);
            const float ka = __shfl_sync(0xffffffffu, area, l);
            if (lane == l) {
              s_kept[nk][0] = x1; s_kept[nk][1] = y1; s_kept[nk][2] = x2; s_kept[nk][3] = y2; s_kept[nk][4] = area;
              s_score[nk] = score;
              s_cls[nk] = cls;
              keep[(long long)b * out_stride + nk] = anchor;
            }
            ++nk;
            if (alive && lane > l && iou_gt(kx1, ky1, kx2, ky2, ka, x1, y1, x2, y2, area, iou_thr)) alive = false;
            live = __ballot_sync(0xffffffffu, alive) & ~((2u << l) - 1u);
          }
          if (lane == 0) { s_range[ws][0] = nkept; s_range[ws][1] = nk; }
        }
        __syncthreads();
        const int kb = s_range[ws][0], ke = s_range[ws][1];
        if (warp > ws) {
          for (int k = kb; k < ke; ++k) {
            if (alive && iou_gt(s_kept[k][0], s_kept[k][1], s_kept[k][2], s_kept[k][3], s_kept[k][4], x1, y1, x2, y2, area, iou_thr))
              alive = false;
          }
        }
        nkept = ke;
        if (nkept >= max_det) break;  // CTA-uniform
      }
      __syncthreads();  // s_range is rewritten next round
    }
    if (threadIdx.x == 0) { s_nkept = nkept; count[b] = nkept; }
  }
  __syncthreads();
  if (kProf && threadIdx.x == 0) {  // slowest image of the launch: [0] candidates, [1] sort, [2] greedy scan cycles, [3] kept
    const long long _np2 = clock64();
    if (atomicMax(&g_conv_prof[2], (unsigned long long)(_np2 - _np1)) < (unsigned long long)(_np2 - _np1)) {
      g_conv_prof[0] = (unsigned long long)n;
      g_conv_prof[1] = (unsigned long long)(_np1 - _np0);
      g_conv_prof[3] = (unsigned long long)s_nkept;
      g_conv_prof[4] = (unsigned long long)s_path;
    }
  }
  const int nk = s_nkept;
  const FrameXform t = xf[b];
  for (int i = threadIdx.x; i < nk; i += blockDim.x) {
    const int anchor = keep[(long long)b * out_stride + i];
    const long long ga = (long long)b * g.A + anchor;
    const float4 bx = dbox[ga];
    const int cls = s_cls[i];
    const float score = s_score[i];
    if (det == nullptr) continue;  // selection-only call (ypb_nms)
    float* lb = det_lb + ((long long)b * out_stride + i) * 4;
    lb[0] = bx.x; lb[1] = bx.y; lb[2] = bx.z; lb[3] = bx.w;
    float* o = det + ((long long)b * out_stride + i) * 6;
    o[0] = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.x, t.pad_w), t.gain), 0.0f), t.W0);
    o[1] = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.y, t.pad_h), t.gain), 0.0f), t.H0);
    o[2] = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.z, t.pad_w), t.gain), 0.0f), t.W0);
    o[3] = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.w, t.pad_h), t.gain), 0.0f), t.H0);
    o[4] = score;
    o[5] = (float)cls;
  }
  if (g.nm > 0 && coef != nullptr) {
    for (int i = threadIdx.x; i < nk * g.nm; i += blockDim.x) {
      const int r = i / g.nm, c = i - r * g.nm;
      const int anchor = keep[(long long)b * out_stride + r];
      coef[((long long)b * out_stride + r) * g.nm + c] = head[((long long)b * g.A + anchor) * g.no + 64 + g.nc + c];
    }
  }
}

}  // namespace ypb
