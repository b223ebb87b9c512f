// Bandwidth-bound layer kernels around the tensor-core convs: fused preprocess + stem conv,
// nearest 2x upsample into a concat slice, SPPF chained max-pools.
// UPSTREAM sites replaced: engine/predictor.py::preprocess (BGR->RGB, /255) + model.0 Conv,
// nn.Upsample(2,'nearest') + Concat, block.py::SPPF's three MaxPool2d(5,1,2)  (SURVEY.md §2.2).
#pragma once
#include "common.cuh"

namespace ypb {

// ------------------------------------------------------------------------------------------------
// Stem: uint8 BGR HWC letterboxed frame -> RGB/255 -> Conv3x3 s2 p1 (+folded BN) -> SiLU -> bf16 NHWC.
// The frame is read as bytes straight from HBM (3 B/pixel instead of a 12 B/pixel fp32 CHW tensor),
// K = 27 is too thin for the tensor pipe so this one runs on the FMA pipe in fp32.
// Weights: wk[(kh*3+kw)*3 + c_rgb][C0] fp32, bias[C0].
// ------------------------------------------------------------------------------------------------
constexpr int kStemTile = 16;  // 16x16 output pixels per CTA, 256 threads

__global__ void __launch_bounds__(256)
stem_conv_kernel(const uint8_t* __restrict__ frames, int H, int W, const float* __restrict__ wk,
                 const float* __restrict__ bias, int C0, __nv_bfloat16* __restrict__ out, int out_ctot) {
  extern __shared__ float stem_smem[];
  const int IT = 2 * kStemTile + 1;               // 33 input rows/cols per tile
  float* s_in = stem_smem;                        // [IT][IT][3] RGB/255
  float* s_w = s_in + IT * IT * 3;                // [27][C0]
  float* s_b = s_w + 27 * C0;                     // [C0]
  const int oH = H >> 1, oW = W >> 1;
  const int b = blockIdx.z;
  const int oh0 = blockIdx.y * kStemTile, ow0 = blockIdx.x * kStemTile;
  const int ih0 = 2 * oh0 - 1, iw0 = 2 * ow0 - 1;
  const uint8_t* img = frames + (size_t)b * H * W * 3;
  for (int i = threadIdx.x; i < IT * IT; i += 256) {
    const int r = i / IT, c = i - r * IT;
    const int ih = ih0 + r, iw = iw0 + c;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
      const uint8_t* px = img + ((size_t)ih * W + iw) * 3;
      v2 = __fdiv_rn((float)px[0], 255.f);  // B -> channel 2
      v1 = __fdiv_rn((float)px[1], 255.f);
      v0 = __fdiv_rn((float)px[2], 255.f);  // R -> channel 0
    }
    s_in[i * 3 + 0] = v0;
    s_in[i * 3 + 1] = v1;
    s_in[i * 3 + 2] = v2;
  }
  for (int i = threadIdx.x; i < 27 * C0; i += 256) s_w[i] = wk[i];
  for (int i = threadIdx.x; i < C0; i += 256) s_b[i] = bias[i];
  __syncthreads();
  const int ty = threadIdx.x / kStemTile, tx = threadIdx.x - ty * kStemTile;
  const int oh = oh0 + ty, ow = ow0 + tx;
  if (oh >= oH || ow >= oW) return;
  float x[27];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
      for (int c = 0; c < 3; ++c) x[(kh * 3 + kw) * 3 + c] = s_in[((2 * ty + kh) * IT + 2 * tx + kw) * 3 + c];
  __nv_bfloat16* o = out + ((size_t)(b * oH + oh) * oW + ow) * out_ctot;
  for (int n0 = 0; n0 < C0; n0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = s_b[n0 + j];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(x[k], s_w[k * C0 + n0 + j], acc[j]);
    }
    uint4 s = make_uint4(pack_bf16x2(silu_f(acc[0]), silu_f(acc[1])), pack_bf16x2(silu_f(acc[2]), silu_f(acc[3])),
                         pack_bf16x2(silu_f(acc[4]), silu_f(acc[5])), pack_bf16x2(silu_f(acc[6]), silu_f(acc[7])));
    *reinterpret_cast<uint4*>(o + n0) = s;
  }
}

// ------------------------------------------------------------------------------------------------
// Nearest 2x upsample of a channel slice into a channel slice of the (2h, 2w) concat buffer.
// One thread per (output pixel, 8-channel vector).
// ------------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const __nv_bfloat16* __restrict__ in, int in_ctot, int in_c_off,
                                  __nv_bfloat16* __restrict__ out, int out_ctot, int out_c_off, int nB, int h, int w,
                                  int C) {
  const int vec = C >> 3;
  const long long total = (long long)nB * 4 * h * w * vec;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int v = (int)(idx % vec);
  long long pix = idx / vec;
  const int ow = (int)(pix % (2 * w));
  pix /= (2 * w);
  const int oh = (int)(pix % (2 * h));
  const int b = (int)(pix / (2 * h));
  const uint4 val = *reinterpret_cast<const uint4*>(in + (((long long)b * h + (oh >> 1)) * w + (ow >> 1)) * in_ctot +
                                                    in_c_off + v * 8);
  *reinterpret_cast<uint4*>(out + (((long long)b * 2 * h + oh) * (2 * w) + ow) * out_ctot + out_c_off + v * 8) = val;
}

// ------------------------------------------------------------------------------------------------
// SPPF pools: y1 = maxpool5(x), y2 = maxpool5(y1) (== 9x9 window), y3 = maxpool5(y2) (== 13x13),
// windows clipped at the border (MaxPool2d pads with -inf).  x is channel slice 0 of a (.., 4c)
// buffer; y1..y3 are written to slices 1..3 of the same buffer, so SPPF.cv2 reads the concat as is.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

// One CTA per (image, 8-channel vector): the h x w map of that vector lives in shared memory (2 x h*w*16 B) and each
// pool is two separable 5-tap passes (horizontal into T, vertical back into A), chained three times exactly as upstream
// chains its MaxPool2d modules (max is exact in bf16, so y2 == maxpool9(x) and y3 == maxpool13(x) bit for bit).
// 30 shared-memory reads per output instead of the 169 global loads of a direct 13x13 window.
__global__ void __launch_bounds__(256)
sppf_pool_kernel(__nv_bfloat16* __restrict__ buf, int nB, int h, int w, int c) {
  extern __shared__ uint4 sppf_smem[];
  uint4* A = sppf_smem;
  uint4* T = sppf_smem + h * w;
  const int ctot = 4 * c;
  const int v = blockIdx.x, b = blockIdx.y;
  __nv_bfloat16* base = buf + (long long)b * h * w * ctot + v * 8;
  const int n = h * w;
  for (int i = threadIdx.x; i < n; i += blockDim.x) A[i] = *reinterpret_cast<const uint4*>(base + (long long)i * ctot);
  for (int stage = 1; stage <= 3; ++stage) {
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int y = i / w, x = i - y * w;
      const int x0 = x - 2 < 0 ? 0 : x - 2, x1 = x + 2 >= w ? w - 1 : x + 2;
      uint4 m = A[y * w + x0];
      for (int xx = x0 + 1; xx <= x1; ++xx) m = bf16x8_max(m, A[y * w + xx]);
      T[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int y = i / w, x = i - y * w;
      const int y0 = y - 2 < 0 ? 0 : y - 2, y1 = y + 2 >= h ? h - 1 : y + 2;
      uint4 m = T[y0 * w + x];
      for (int yy = y0 + 1; yy <= y1; ++yy) m = bf16x8_max(m, T[yy * w + x]);
      A[i] = m;
      *reinterpret_cast<uint4*>(base + (long long)i * ctot + stage * c) = m;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LetterBox on the device: raw uint8 BGR frames -> resized (cv2.INTER_LINEAR) and 114-padded
// network input.  Replaces the host cv2.resize + copyMakeBorder of UPSTREAM data/augment.py::LetterBox for the
// down-scaling case (video frames larger than the network size), bit for bit: OpenCV's 8-bit bilinear resize is
// fixed-point — 11-bit coefficients (tables built on the host exactly as cv2 builds them), horizontal pass in int32,
// vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2.  One thread per output pixel (3 channels).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
letterbox_u8_kernel(const uint8_t* __restrict__ src, int nB, int H0, int W0, uint8_t* __restrict__ dst, int H, int W, int new_w,
                    int new_h, int top, int left, const int* __restrict__ xofs, const short* __restrict__ xa,
                    const int* __restrict__ yofs, const short* __restrict__ ya, int pad) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)H * W;
  if (idx >= per * nB) return;
  const int b = (int)(idx / per);
  const int rem = (int)(idx - (long long)b * per);
  const int y = rem / W, x = rem - y * W;
  uint8_t* o = dst + idx * 3;
  const int ry = y - top, rx = x - left;
  if (ry < 0 || ry >= new_h || rx < 0 || rx >= new_w) {
    o[0] = o[1] = o[2] = (uint8_t)pad;
    return;
  }
  const int x0 = xofs[rx], x1 = min(x0 + 1, W0 - 1), y0 = yofs[ry], y1 = min(y0 + 1, H0 - 1);
  const int a0 = xa[2 * rx], a1 = xa[2 * rx + 1], b0 = ya[2 * ry], b1 = ya[2 * ry + 1];
  const uint8_t* f = src + (long long)b * H0 * W0 * 3;
  const uint8_t* r0 = f + (long long)y0 * W0 * 3;
  const uint8_t* r1 = f + (long long)y1 * W0 * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int h0 = r0[x0 * 3 + c] * a0 + r0[x1 * 3 + c] * a1;
    const int h1 = r1[x0 * 3 + c] * a0 + r1[x1 * 3 + c] * a1;
    int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    o[c] = (uint8_t)v;
  }
}

}  // namespace ypb
