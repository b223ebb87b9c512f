// Bandwidth-bound layer kernels around the tensor-core convs: nearest 2x upsample into a concat slice, SPPF chained
// max-pools, LetterBox on the device.  (The stem lives in conv_tc.cuh: stem_tc_kernel.)
// UPSTREAM sites replaced: nn.Upsample(2,'nearest') + Concat, block.py::SPPF's three MaxPool2d(5,1,2),
// data/augment.py::LetterBox  (SURVEY.md §2.2, §8f).
#pragma once
#include "common.cuh"

namespace ypb {

// ------------------------------------------------------------------------------------------------
// Nearest 2x upsample of a channel slice into a channel slice of the (2h, 2w) concat buffer.
// One thread per (input pixel, 8-channel vector): one load, the 2x2 block of stores.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
upsample2x_kernel(const __nv_bfloat16* __restrict__ in, int in_ctot, int in_c_off, __nv_bfloat16* __restrict__ out, int out_ctot,
                  int out_c_off, int nB, int h, int w, int C) {
  // one thread per (INPUT pixel, 8-channel vector): one 16-byte load, four 16-byte stores (the 2x2 output block).
  // grid.y = (image, input row), grid.x covers the w * C/8 vectors of that row: one 32-bit division per thread
  // (the flat-index version spent most of its instructions on three 64-bit div/mod pairs).
  const int vec = C >> 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * vec) return;
  const int x = i / vec, v = i - x * vec;
  const int b = blockIdx.y / h, y = blockIdx.y - b * h;
  (void)nB;
  const uint4 val = *reinterpret_cast<const uint4*>(in + (((long long)b * h + y) * w + x) * in_ctot + in_c_off + v * 8);
  __nv_bfloat16* o = out + (((long long)b * 2 * h + 2 * y) * (2 * w) + 2 * x) * out_ctot + out_c_off + v * 8;
  const long long row = (long long)2 * w * out_ctot;
  *reinterpret_cast<uint4*>(o) = val;
  *reinterpret_cast<uint4*>(o + out_ctot) = val;
  *reinterpret_cast<uint4*>(o + row) = val;
  *reinterpret_cast<uint4*>(o + row + out_ctot) = val;
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
// network input.  Replaces the host cv2.resize + copyMakeBorder of UPSTREAM data/augment.py::LetterBox, down- and
// up-scaling, bit for bit: OpenCV's 8-bit bilinear resize is
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
  // columns: cv2 clamps the index and zeroes the fraction in the table; rows: it keeps the fraction and clamps the two
  // source rows when it fetches them (sy = -1 above the first row centre when up-scaling): restated exactly
  const int x0 = xofs[rx], x1 = min(x0 + 1, W0 - 1);
  const int sy = yofs[ry], y0 = min(max(sy, 0), H0 - 1), y1 = min(max(sy + 1, 0), H0 - 1);
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
