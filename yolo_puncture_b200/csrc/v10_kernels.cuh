// YOLOv10-only layer kernels: depthwise convs (SCDown.cv2, CIB, RepVGGDW 7x7, Attention.pe, v10Detect
// class branch) and the PSA multi-head self-attention.  Bandwidth / latency bound; CUDA cores.
// UPSTREAM sites replaced: block.py::{SCDown, CIB, RepVGGDW, Attention, PSA}, head.py::v10Detect
//  — the NMS-free detector variant behind the same model.predict() call.
#pragma once
#include "common.cuh"

namespace ypb {

struct DwParams {
  const __nv_bfloat16* in;   // NHWC (B, H, W, in_ctot), channels [in_c_off, +C)
  int H, W, in_ctot, in_c_off;
  const __nv_bfloat16* wk;   // [k*k][C] bf16 (BN folded)
  const float* bias;         // [C]
  int C, k, stride, act;
  __nv_bfloat16* out;        // NHWC (B, oH, oW, out_ctot), channels [out_c_off, +C)
  int oH, oW, out_ctot, out_c_off;
  const __nv_bfloat16* res;  // optional residual (same indexing rule as out, own ctot/c_off), added after act
  int res_ctot, res_c_off;
  int nB;
};

// One thread per (output pixel, 8-channel vector).  w_row = channels per weight row ([tap][w_row] layout), so a
// launch may cover a channel sub-range of a wider depthwise conv (Attention.pe runs once per head).
__global__ void dwconv_strided_kernel(const DwParams p, int w_row) {
  const int vec = p.C >> 3;
  const long long total = (long long)p.nB * p.oH * p.oW * vec;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int v = (int)(idx % vec);
  long long pix = idx / vec;
  const int ow = (int)(pix % p.oW);
  pix /= p.oW;
  const int oh = (int)(pix % p.oH);
  const int b = (int)(pix / p.oH);
  const int pad = p.k >> 1;
  float acc[8];
  {
    const float4 b0 = *reinterpret_cast<const float4*>(p.bias + v * 8), b1 = *reinterpret_cast<const float4*>(p.bias + v * 8 + 4);
    acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
  }
  for (int kh = 0; kh < p.k; ++kh) {
    const int ih = oh * p.stride - pad + kh;
    if (ih < 0 || ih >= p.H) continue;
    for (int kw = 0; kw < p.k; ++kw) {
      const int iw = ow * p.stride - pad + kw;
      if (iw < 0 || iw >= p.W) continue;
      const uint4 xi = *reinterpret_cast<const uint4*>(p.in + (((long long)b * p.H + ih) * p.W + iw) * p.in_ctot + p.in_c_off + v * 8);
      const uint4 wi = *reinterpret_cast<const uint4*>(p.wk + (long long)(kh * p.k + kw) * w_row + v * 8);
      const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&xi);
      const __nv_bfloat162* wp = reinterpret_cast<const __nv_bfloat162*>(&wi);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 xf = __bfloat1622float2(xp[j]), wf = __bfloat1622float2(wp[j]);
        acc[2 * j] = fmaf(xf.x, wf.x, acc[2 * j]);
        acc[2 * j + 1] = fmaf(xf.y, wf.y, acc[2 * j + 1]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = p.act ? silu_f(acc[j]) : acc[j];
  const long long opix = ((long long)b * p.oH + oh) * p.oW + ow;
  if (p.res != nullptr) {
    const uint4 r = *reinterpret_cast<const uint4*>(p.res + opix * p.res_ctot + p.res_c_off + v * 8);
    const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 rf = __bfloat1622float2(rp[j]);
      acc[2 * j] = bf16_round(acc[2 * j]) + rf.x;
      acc[2 * j + 1] = bf16_round(acc[2 * j + 1]) + rf.y;
    }
  }
  *reinterpret_cast<uint4*>(p.out + opix * p.out_ctot + p.out_c_off + v * 8) =
      make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
}

// ------------------------------------------------------------------------------------------------
// PSA attention core: out = V softmax(Q^T K * scale)^T + pe, per (image, head).
// qkv: NHWC (B, N tokens, qkv_ctot); head h owns channels [h*(2*KD+HD), ...) laid out [q(KD) | k(KD) | v(HD)]
// pe : NHWC (B, N, pe_ctot) slice [pe_c_off + h*HD, +HD) = depthwise 3x3 of v (already computed, bf16)
// out: NHWC (B, N, out_ctot) slice [out_c_off + h*HD, +HD)
// One thread per query token; keys/values streamed through shared memory in chunks with an online softmax.
// ------------------------------------------------------------------------------------------------
constexpr int kAttnKD = 32, kAttnHD = 64, kAttnChunk = 64, kAttnThreads = 128;

__global__ void __launch_bounds__(kAttnThreads)
psa_attention_kernel(const __nv_bfloat16* __restrict__ qkv, int qkv_ctot, int qkv_c_off, const __nv_bfloat16* __restrict__ pe,
                     int pe_ctot, int pe_c_off, __nv_bfloat16* __restrict__ out, int out_ctot, int out_c_off, int N,
                     float scale) {
  __shared__ float s_k[kAttnChunk][kAttnKD];
  __shared__ float s_v[kAttnChunk][kAttnHD];
  const int h = blockIdx.y, b = blockIdx.z;
  const int qi = blockIdx.x * kAttnThreads + threadIdx.x;
  const int hc = 2 * kAttnKD + kAttnHD;
  const __nv_bfloat16* base = qkv + (long long)b * N * qkv_ctot + qkv_c_off + h * hc;
  float q[kAttnKD];
  if (qi < N) {
#pragma unroll
    for (int d = 0; d < kAttnKD; ++d) q[d] = __bfloat162float(base[(long long)qi * qkv_ctot + d]) * scale;
  } else {
#pragma unroll
    for (int d = 0; d < kAttnKD; ++d) q[d] = 0.f;
  }
  float m = -INFINITY, l = 0.f, acc[kAttnHD];
#pragma unroll
  for (int d = 0; d < kAttnHD; ++d) acc[d] = 0.f;
  for (int j0 = 0; j0 < N; j0 += kAttnChunk) {
    __syncthreads();
    for (int i = threadIdx.x; i < kAttnChunk * (kAttnKD + kAttnHD); i += kAttnThreads) {
      const int r = i / (kAttnKD + kAttnHD), c = i - r * (kAttnKD + kAttnHD);
      const int j = j0 + r;
      const float val = j < N ? __bfloat162float(base[(long long)j * qkv_ctot + kAttnKD + c]) : 0.f;
      if (c < kAttnKD) s_k[r][c] = val; else s_v[r][c - kAttnKD] = val;
    }
    __syncthreads();
    const int jn = min(kAttnChunk, N - j0);
    for (int r = 0; r < jn; ++r) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < kAttnKD; ++d) s = fmaf(q[d], s_k[r][d], s);
      if (s > m) {
        const float corr = __expf(m - s);
        l *= corr;
#pragma unroll
        for (int d = 0; d < kAttnHD; ++d) acc[d] *= corr;
        m = s;
      }
      const float pr = __expf(s - m);
      l += pr;
#pragma unroll
      for (int d = 0; d < kAttnHD; ++d) acc[d] = fmaf(pr, s_v[r][d], acc[d]);
    }
  }
  if (qi >= N) return;
  const float inv = 1.0f / l;
  const __nv_bfloat16* pp = pe + ((long long)b * N + qi) * pe_ctot + pe_c_off + h * kAttnHD;
  __nv_bfloat16* op = out + ((long long)b * N + qi) * out_ctot + out_c_off + h * kAttnHD;
#pragma unroll
  for (int d = 0; d < kAttnHD; d += 8) {
    const uint4 pv = *reinterpret_cast<const uint4*>(pp + d);
    const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&pv);
    float y[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 pf = __bfloat1622float2(p2[j]);
      y[2 * j] = acc[d + 2 * j] * inv + pf.x;
      y[2 * j + 1] = acc[d + 2 * j + 1] * inv + pf.y;
    }
    *reinterpret_cast<uint4*>(op + d) = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]),
                                                   pack_bf16x2(y[6], y[7]));
  }
}

}  // namespace ypb
