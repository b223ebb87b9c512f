// Fused proto-mask decode: coef x proto GEMM + letterbox un-pad + bilinear upsample + box crop + threshold
// in one kernel, writing uint8 {0,1} masks.  Logits are never materialised at frame resolution.
// UPSTREAM sites replaced: utils/ops.py::process_mask_native -> scale_masks -> crop_mask (retina,
// reference yolo_seg/app.py:49,91 and yolo_seg/yolo_with_deva.py:51) and ops.process_mask (non-retina,
// reference dev_tools/auto_speed_calc.py:62)  (SURVEY.md §8 a10, a11).
// sigmoid(interp(x)) > 0.5  <=>  interp(x) > 0, so no sigmoid is evaluated at all.
#pragma once
#include "common.cuh"

namespace ypb {

struct MaskGeom {
  int mh, mw;          // proto map size
  int nm;              // 32
  int max_det;
  int retina;          // 1: output (H0,W0) per image, crop with frame boxes; 0: output (H,W), crop in proto space
  int out_h, out_w;    // mask size (all images of a call share it)
  // source window of the bilinear resize in proto space and its scale (ATen: scale = in/out as float)
  int top, left, ch, cw;
  float scale_h, scale_w;
  // non-retina: proto-space box = letterbox box * (mw/W, mh/H)
  float ratio_w, ratio_h;
  int prefilled;       // 1: the output was zero-filled by mask_zero_kernel, only box-touching tiles are written
};

// Exclusive prefix sum of per-image detection counts -> offsets[0..B]; flags overflow of the mask capacity.
__global__ void mask_offsets_kernel(const int* __restrict__ count, int nB, int capacity, int* __restrict__ offsets,
                                    int* __restrict__ status) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < nB; ++b) { offsets[b] = acc; acc += count[b]; }
    offsets[nB] = acc;
    status[0] = acc;                       // total detections of the batch
    status[1] = acc > capacity ? 1 : 0;    // masks beyond capacity were not written
  }
}

constexpr int kMaskTile = 64;  // 64x64 output pixels per step, 256 threads x 16 px

// Streaming zero fill of the masks of the first min(total, capacity) detections: a plain grid-stride loop of 16-byte
// stores runs at the HBM write rate, which the decode kernel's per-band zero fills (L1/LSU-bound next to its
// shared-memory traffic: ncu l1tex 85 %, 2.3 TB/s) did not.  mask_decode_kernel then only writes the tiles a box touches.
__global__ void __launch_bounds__(256)
mask_zero_kernel(const int* __restrict__ offsets, int nB, int capacity, long long bytes_per_mask, uint8_t* __restrict__ out) {
  const int total = min(offsets[nB], capacity);
  const long long nbytes = (long long)total * bytes_per_mask;
  uint4* o4 = reinterpret_cast<uint4*>(out);
  const long long n16 = nbytes >> 4;
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) o4[i] = z;
  if (blockIdx.x == 0)
    for (long long i = (n16 << 4) + threadIdx.x; i < nbytes; i += blockDim.x) out[i] = 0;
}

// One CTA per (band of 64 output rows, detection).  Bands that miss the box are one contiguous, fully coalesced
// zero fill; inside a band the CTA walks the 64-pixel tiles the box touches (the others are zero-filled in one flat
// loop): it first computes the low-resolution logits of just the proto window those tiles need (<= 66 rows of dot
// products of length 32, in shared memory), then every thread produces 16 output pixels per tile with ATen's bilinear
// arithmetic, the box crop and the > 0 threshold.
// ncu (profiles/NCU_SUMMARY.md, r2 mask captures): the kernel is ISSUE-bound (72 % issue slots busy, DRAM 25 %), so the
// code below counts instructions: CTA-uniform tile windows are computed once into shared memory, prototype row / column
// come from a multiply-high instead of an integer division, the crop tests are integer bit masks, and the dot products
// are laid out so that loads are coalesced (four lanes per prototype) and reduced with three shuffles per FOUR prototypes.
template <bool RETINA>
__global__ void __launch_bounds__(256, 5)
mask_decode_kernel(const float* __restrict__ proto /*(B,mh,mw,nm) fp32*/, const float* __restrict__ coef /*(B,max_det,nm)*/,
                   const float* __restrict__ det /*(B,max_det,6) frame boxes*/, const float* __restrict__ det_lb /*(B,max_det,4)*/,
                   const int* __restrict__ offsets, int nB, int capacity, MaskGeom g, uint8_t* __restrict__ out) {
  extern __shared__ float s_band[];  // low-resolution logits of the band's proto window: rh rows x band_w columns
  // horizontally interpolated window rows of a tile; column c lives at c + 4 * (c / 32) (rows of 72 floats): the vertical
  // pass reads 16-byte pieces at columns 0 / 16 / 32 / 48 of a row and columns 0 and 32 would share their banks (ncu:
  // 2.1 wavefronts per request there before the padding)
  constexpr int kHrowPitch = kMaskTile + 8;
  __shared__ __align__(16) float s_hrow[(kMaskTile + 2) * kHrowPitch];
  __shared__ __align__(16) float s_coef[32];
  constexpr int kMaxTiles = 64;  // out_w <= 4096; wider outputs recompute the windows per thread
  __shared__ int s_tlo[kMaxTiles], s_thi[kMaxTiles];  // proto columns a tile needs; s_thi < s_tlo: the tile misses the box
  __shared__ unsigned s_tmask[2];                      // bit t: tile t touches the box
  const int slot = blockIdx.y;
  const int total = offsets[nB];
  if (slot >= total || slot >= capacity) return;
  // slot -> (image, detection): the image is the last one whose offset is <= slot.  Every warp counts those with
  // its lanes side by side (one L2 round trip; a binary search is log2(B) dependent ones before the CTA can start)
  int below = 0;
  for (int i0 = 0; i0 < nB; i0 += 32) {
    const int i = i0 + (threadIdx.x & 31);
    below += __popc(__ballot_sync(0xffffffffu, i < nB && offsets[i] <= slot));
  }
  const int b = below - 1, di = slot - offsets[b];
  const int ty0 = blockIdx.x * kMaskTile;

  float bx1, by1, bx2, by2;
  if (RETINA) {
    const float* d = det + ((long long)b * g.max_det + di) * 6;
    bx1 = d[0]; by1 = d[1]; bx2 = d[2]; by2 = d[3];
  } else {
    const float* d = det_lb + ((long long)b * g.max_det + di) * 4;
    bx1 = __fmul_rn(d[0], g.ratio_w); by1 = __fmul_rn(d[1], g.ratio_h);
    bx2 = __fmul_rn(d[2], g.ratio_w); by2 = __fmul_rn(d[3], g.ratio_h);
  }
  uint8_t* o = out + (long long)slot * g.out_h * g.out_w;
  auto src_of = [](int dst, float scale) {
    float s = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
    return s < 0.f ? 0.f : s;
  };
  const int y_last = min(ty0 + kMaskTile, g.out_h) - 1;
  const int sy_lo = (int)src_of(ty0, g.scale_h);
  const int sy_hi = min((int)src_of(y_last, g.scale_h) + 1, g.ch - 1);
  const int rh = sy_hi - sy_lo + 1;
  const bool vec_ok = ((g.out_w & 15) == 0);

  // ---- whole band outside the box: contiguous zero fill ----
  bool band_empty;
  if (RETINA) band_empty = ((float)(ty0 + kMaskTile) <= by1) || ((float)ty0 >= by2);
  else band_empty = ((float)(g.top + sy_hi) < by1) || ((float)(g.top + sy_lo) >= by2);
  if (band_empty || rh > kMaskTile + 2) {
    if (g.prefilled) return;
    const long long nbytes = (long long)(y_last - ty0 + 1) * g.out_w;
    uint8_t* dst = o + (long long)ty0 * g.out_w;
    if (((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)nbytes) & 15) == 0) {
      const int n16 = (int)(nbytes >> 4);
      for (int i = threadIdx.x; i < n16; i += 256) reinterpret_cast<uint4*>(dst)[i] = make_uint4(0, 0, 0, 0);
    } else {
      for (long long i = threadIdx.x; i < nbytes; i += 256) dst[i] = 0;
    }
    return;
  }

  if (threadIdx.x < 32) s_coef[threadIdx.x] = coef[((long long)b * g.max_det + di) * 32 + threadIdx.x];  // nm == 32 (checked by the host)
  const float* pb = proto + (long long)b * g.mh * g.mw * 32;
  const int trow = threadIdx.x >> 2, tcol = (threadIdx.x & 3) * 16;
  const int oy = ty0 + trow;
  const float sy = src_of(min(oy, g.out_h - 1), g.scale_h);
  const int y0 = (int)sy;
  const int y1 = y0 + ((y0 < g.ch - 1) ? 1 : 0);
  const float ly = __fsub_rn(sy, (float)y0), hy = __fsub_rn(1.0f, ly);
  const bool row_in = RETINA ? ((float)oy >= by1 && (float)oy < by2) : true;

  // ---- per-tile source windows, once per CTA ----
  const int n_tiles = (g.out_w + kMaskTile - 1) / kMaskTile;
  auto tile_window = [&](int tx0, int* sx_lo, int* sx_hi) {  // empty tiles come back as (1, 0)
    const int x_last = min(tx0 + kMaskTile, g.out_w) - 1;
    int lo = (int)src_of(tx0, g.scale_w);
    int hi = min((int)src_of(x_last, g.scale_w) + 1, g.cw - 1);
    bool empty = hi - lo + 1 > kMaskTile + 2;
    if (RETINA) empty = empty || ((float)(tx0 + kMaskTile) <= bx1) || ((float)tx0 >= bx2);
    else empty = empty || ((float)(g.left + hi) < bx1) || ((float)(g.left + lo) >= bx2);
    *sx_lo = empty ? 1 : lo;
    *sx_hi = empty ? 0 : hi;
  };
  // the tiles the box touches form one run [t_first, t_last]; bw_lo..bw_hi = the proto columns they need (CTA-uniform)
  int bw_lo = 1 << 30, bw_hi = -1, t_first = n_tiles, t_last = -1;
  unsigned long long tmask = 0ull;  // bit t: tile t touches the box (n_tiles <= 64 only)
  if (n_tiles <= kMaxTiles) {
    if (threadIdx.x < 64) {  // warps 0 and 1: one tile per lane, the non-empty ones as two ballot words
      int lo = 1, hi = 0;
      if (threadIdx.x < n_tiles) tile_window(threadIdx.x * kMaskTile, &lo, &hi);
      s_tlo[threadIdx.x] = lo;
      s_thi[threadIdx.x] = hi;
      const unsigned bal = __ballot_sync(0xffffffffu, hi >= lo);
      if ((threadIdx.x & 31) == 0) s_tmask[threadIdx.x >> 5] = bal;
    }
    __syncthreads();  // s_coef and the tile windows are ready
    tmask = ((unsigned long long)s_tmask[1] << 32) | s_tmask[0];
    if (tmask) {
      t_first = __ffsll((long long)tmask) - 1;
      t_last = 63 - __clzll((long long)tmask);
      bw_lo = s_tlo[t_first];  // the windows move right with the tile index
      bw_hi = s_thi[t_last];
      for (int t = t_first + 1; t < t_last; ++t)  // (a tile in the middle may be "empty" only through the width guard)
        if ((tmask >> t) & 1ull) { bw_lo = min(bw_lo, s_tlo[t]); bw_hi = max(bw_hi, s_thi[t]); }
    }
  } else {
    __syncthreads();  // s_coef
    for (int t = 0; t < n_tiles; ++t) {
      int lo, hi;
      tile_window(t * kMaskTile, &lo, &hi);
      if (hi < lo) continue;
      bw_lo = min(bw_lo, lo);
      bw_hi = max(bw_hi, hi);
      t_first = min(t_first, t);
      t_last = t;
    }
  }
  auto tile_is_empty = [&](int t) {
    if (n_tiles <= kMaxTiles) return ((tmask >> t) & 1ull) == 0ull;
    int lo, hi;
    tile_window(t * kMaskTile, &lo, &hi);
    return hi < lo;
  };
  const int band_w = bw_hi >= bw_lo ? bw_hi - bw_lo + 1 : 0;

  // ---- the band's logits, once.  Four lanes per prototype: lane l holds channels 4l..4l+3 and 16+4l..16+4l+3, so a
  // warp-wide 16-byte load covers eight 64-byte runs of full sectors (a thread-per-prototype loop touches 32 different
  // 128-byte lines with every load instruction).  Each lane group works on FOUR consecutive prototypes at a time and the
  // four partial sums are reduced with a transposing butterfly - 3 shuffles + 3 adds per lane for the four of them -
  // after which lane l holds the finished logit of one of them.  All 8 loads of a thread are in flight together. ----
  {
    const int l4 = threadIdx.x & 3, g4 = threadIdx.x >> 2;
    const bool o1 = (l4 & 1) != 0, o2 = (l4 & 2) != 0;
    const float4 c0 = *reinterpret_cast<const float4*>(&s_coef[4 * l4]);
    const float4 c1 = *reinterpret_cast<const float4*>(&s_coef[16 + 4 * l4]);
    const float4* pb4 = reinterpret_cast<const float4*>(pb) + ((g.top + sy_lo) * g.mw + g.left + bw_lo) * 8 + l4;
    const int npx = rh * band_w;  // <= 66 * 1026
    const unsigned inv_bw = band_w > 0 ? 0xFFFFFFFFu / (unsigned)band_w + 1u : 0u;  // floor(i / band_w) = umulhi(i, inv_bw), i < 2^16
    auto row_col = [&](int i, int* ry, int* rx) {
      *ry = npx < 65536 ? (int)__umulhi((unsigned)i, inv_bw) : i / band_w;
      *rx = i - *ry * band_w;
    };
    for (int i0 = 0; i0 < npx; i0 += 256) {
      const int ib = i0 + g4 * 4;
      float p[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        p[u] = 0.f;
        if (ib + u < npx) {
          int ry, rx;
          row_col(ib + u, &ry, &rx);
          const float4* a = pb4 + (ry * g.mw + rx) * 8;
          const float4 v0 = __ldg(a), v1 = __ldg(a + 4);
          float acc = __fmul_rn(c0.x, v0.x);
          acc = fmaf(c0.y, v0.y, acc); acc = fmaf(c0.z, v0.z, acc); acc = fmaf(c0.w, v0.w, acc);
          acc = fmaf(c1.x, v1.x, acc); acc = fmaf(c1.y, v1.y, acc); acc = fmaf(c1.z, v1.z, acc); acc = fmaf(c1.w, v1.w, acc);
          p[u] = acc;
        }
      }
      // lanes (l, l^1) swap halves, then (l, l^2) swap quarters: lane l ends with prototype u = 2*(l&1) + ((l>>1)&1)
      const float k0 = (o1 ? p[2] : p[0]) + __shfl_xor_sync(0xffffffffu, o1 ? p[0] : p[2], 1);
      const float k1 = (o1 ? p[3] : p[1]) + __shfl_xor_sync(0xffffffffu, o1 ? p[1] : p[3], 1);
      float f = (o2 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, o2 ? k0 : k1, 2);
      const int iu = ib + (o1 ? 2 : 0) + (o2 ? 1 : 0);
      if (iu < npx) {
        if (!RETINA) {  // ops.process_mask crops in proto space BEFORE the upsample
          int ry, rx;
          row_col(iu, &ry, &rx);
          const float fx = (float)(g.left + bw_lo + rx), fy = (float)(g.top + sy_lo + ry);
          if (!(fx >= bx1 && fx < bx2 && fy >= by1 && fy < by2)) f = 0.f;
        }
        s_band[iu] = f;
      }
    }
  }

  // ---- tiles left and right of the box: zeros, one 16-byte store per thread and tile, no per-tile bookkeeping ----
  if (!g.prefilled && oy < g.out_h) {
    const int x_a = min(max(t_first, 0) * kMaskTile, g.out_w), x_b = min((t_last + 1) * kMaskTile, g.out_w);  // box tiles = [x_a, x_b)
    uint8_t* orow = o + (long long)oy * g.out_w;
    if (t_last < 0) {  // no tile touches the box in this band (cannot happen for a non-empty band; kept for safety)
      for (int x = tcol; x < g.out_w; x += kMaskTile)
        for (int j = 0; j < 16 && x + j < g.out_w; ++j) orow[x + j] = 0;
    } else if (vec_ok) {
      for (int x = tcol; x < x_a; x += kMaskTile) *reinterpret_cast<uint4*>(orow + x) = make_uint4(0, 0, 0, 0);
      for (int x = x_b + tcol; x < g.out_w; x += kMaskTile) *reinterpret_cast<uint4*>(orow + x) = make_uint4(0, 0, 0, 0);
    } else {
      for (int x = tcol; x < x_a; x += kMaskTile)
        for (int j = 0; j < 16; ++j) orow[x + j] = 0;
      for (int x = x_b + tcol; x < g.out_w; x += kMaskTile)
        for (int j = 0; j < 16 && x + j < g.out_w; ++j) orow[x + j] = 0;
    }
  }

  // integer form of the retina crop: (float)x >= b  <=>  x >= ceil(b), (float)x < b  <=>  x < ceil(b) for integer x
  // (no int->float conversion per pixel: those run on the 16-lane XU pipe)
  const int cx_lo = RETINA ? (int)ceilf(fmaxf(bx1, -1.0f)) : 0;
  const int cx_hi = RETINA ? min((int)ceilf(fminf(bx2, 1.0e6f)), g.out_w) : g.out_w;
  for (int t = max(t_first, 0); t <= t_last; ++t) {
    const int tx0 = t * kMaskTile;
    const int ox0 = tx0 + tcol;
    if (tile_is_empty(t)) {  // CTA-uniform: a tile inside the run that is treated as empty (window too wide)
      if (oy < g.out_h && !g.prefilled)
        for (int j = 0; j < 16 && ox0 + j < g.out_w; ++j) o[(long long)oy * g.out_w + ox0 + j] = 0;
      continue;
    }
    __syncthreads();  // band logits written / previous tile's s_hrow fully consumed
    // Separable bilinear, in ATen's operation order (horizontal blend of each source row first, then the vertical
    // blend): the horizontal pass is done ONCE per (window row, output column) into shared memory instead of twice
    // per output pixel, and the vertical pass reads its two rows with 16-byte loads.
    {
      const int hx_col = threadIdx.x & (kMaskTile - 1);
      const int ox = tx0 + hx_col;
      const float sx = src_of(min(ox, g.out_w - 1), g.scale_w);
      const int x0 = (int)sx;
      const int x1 = x0 + ((x0 < g.cw - 1) ? 1 : 0);
      const float lx = __fsub_rn(sx, (float)x0), hx = __fsub_rn(1.0f, lx);
      const float* r0 = s_band + (x0 - bw_lo);
      const float* r1 = s_band + (x1 - bw_lo);
      float* dst = s_hrow + hx_col + ((hx_col >> 5) << 2);
      for (int ry = threadIdx.x >> 6; ry < rh; ry += 4)
        dst[ry * kHrowPitch] = __fadd_rn(__fmul_rn(hx, r0[ry * band_w]), __fmul_rn(lx, r1[ry * band_w]));
    }
    __syncthreads();
    if (oy >= g.out_h) continue;
    // pixels of this thread's 16 that survive the crop, as a bit mask
    const int jl = min(max(cx_lo - ox0, 0), 16), jh = min(max(cx_hi - ox0, 0), 16);
    const unsigned in16 = row_in && jh > jl ? ((1u << jh) - 1u) & ~((1u << jl) - 1u) : 0u;
    if (in16 == 0u) {  // rows above / below the box inside the band, columns left / right of it inside the tile
      if (!g.prefilled) {
        if (vec_ok && ox0 + 16 <= g.out_w) *reinterpret_cast<uint4*>(o + (long long)oy * g.out_w + ox0) = make_uint4(0, 0, 0, 0);
        else for (int j = 0; j < 16 && ox0 + j < g.out_w; ++j) o[(long long)oy * g.out_w + ox0 + j] = 0;
      }
      continue;
    }
    const int pcol = tcol + ((tcol >> 5) << 2);
    const float4* t0 = reinterpret_cast<const float4*>(s_hrow + (y0 - sy_lo) * kHrowPitch + pcol);
    const float4* t1 = reinterpret_cast<const float4*>(s_hrow + (y1 - sy_lo) * kHrowPitch + pcol);
    uint32_t packed[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 a = t0[q], c = t1[q];
      const float v0 = __fadd_rn(__fmul_rn(hy, a.x), __fmul_rn(ly, c.x)), v1 = __fadd_rn(__fmul_rn(hy, a.y), __fmul_rn(ly, c.y));
      const float v2 = __fadd_rn(__fmul_rn(hy, a.z), __fmul_rn(ly, c.z)), v3 = __fadd_rn(__fmul_rn(hy, a.w), __fmul_rn(ly, c.w));
      const uint32_t on = (v0 > 0.0f ? 0x1u : 0u) | (v1 > 0.0f ? 0x100u : 0u) | (v2 > 0.0f ? 0x10000u : 0u) | (v3 > 0.0f ? 0x1000000u : 0u);
      // bits 4q..4q+3 of the crop mask spread to one byte each: b * 0x00204081 puts bit k at position 8k (no carries)
      packed[q] = on & ((((in16 >> (4 * q)) & 0xFu) * 0x00204081u) & 0x01010101u);
    }
    if (vec_ok && ox0 + 16 <= g.out_w) {
      *reinterpret_cast<uint4*>(o + (long long)oy * g.out_w + ox0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    } else {
      for (int j = 0; j < 16 && ox0 + j < g.out_w; ++j)
        o[(long long)oy * g.out_w + ox0 + j] = (uint8_t)((packed[j >> 2] >> (8 * (j & 3))) & 1u);
    }
  }
}

// Tile-stationary mask decode (the default when a 64 x 64 output tile needs at most kMaskWin x kMaskWin prototypes, i.e.
// whenever the masks are up-sampled by >= ~3.7x: retina masks of 640 x 640 and larger frames, non-retina masks).
// mask_decode_kernel is detection-stationary: every (detection, band) CTA re-reads the prototypes under its box from L2
// (a 300 x 300 px box = 720 KB; 1.1 GB per 64-frame step) and that, not its zero fills, is what bounds it (ncu:
// 0.19 ms with 6 MB of DRAM writes once the zero fill is taken out).  Here a CTA owns one 64 x 64 OUTPUT tile of one
// image: it loads the tile's prototype window ONCE (<= 20 x 20 x 32 fp32, transposed in shared memory), then walks
// the image's detections, and for every box that touches the tile computes the window's logits (one 32-long dot
// product per prototype, operands in shared memory) and the tile's 4096 pixels.  Prototypes are read once per image;
// the zero background is written by mask_zero_kernel at the HBM write rate.  Same dot-product order, same ATen blend
// order, same crop and threshold as mask_decode_kernel: bit-identical masks.
constexpr int kMaskWin = 20;

__global__ void __launch_bounds__(256)
mask_tile_kernel(const float* __restrict__ proto, const float* __restrict__ coef, const float* __restrict__ det,
                 const float* __restrict__ det_lb, const int* __restrict__ offsets, int nB, int capacity, MaskGeom g,
                 uint8_t* __restrict__ out) {
  extern __shared__ float s_tile[];            // [32][kMaskWin * kMaskWin] prototypes (channel-major), then 2 logit windows
  float* s_p = s_tile;
  float* s_logit = s_tile + 32 * kMaskWin * kMaskWin;  // [2][kMaskWin * kMaskWin]
  const int b = blockIdx.z;
  const int slot0 = offsets[b], n_b = offsets[b + 1] - slot0;
  if (n_b <= 0) return;
  const int tx0 = blockIdx.x * kMaskTile, ty0 = blockIdx.y * kMaskTile;
  auto src_of = [](int dst, float scale) {
    float s = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
    return s < 0.f ? 0.f : s;
  };
  const int x_last = min(tx0 + kMaskTile, g.out_w) - 1, y_last = min(ty0 + kMaskTile, g.out_h) - 1;
  const int sx_lo = (int)src_of(tx0, g.scale_w), sx_hi = min((int)src_of(x_last, g.scale_w) + 1, g.cw - 1);
  const int sy_lo = (int)src_of(ty0, g.scale_h), sy_hi = min((int)src_of(y_last, g.scale_h) + 1, g.ch - 1);
  const int ww = sx_hi - sx_lo + 1, wh = sy_hi - sy_lo + 1;  // <= kMaskWin (checked by the host)
  const int npx = ww * wh;
  // ---- does any box of this image touch the tile?  (cheap scan; most tiles of most images are background) ----
  bool any = false;
  for (int di = threadIdx.x; di < n_b; di += 256) {
    if (slot0 + di >= capacity) break;
    if (g.retina) {
      const float* d = det + ((long long)b * g.max_det + di) * 6;
      any = any || !(((float)(tx0 + kMaskTile) <= d[0]) || ((float)tx0 >= d[2]) || ((float)(ty0 + kMaskTile) <= d[1]) || ((float)ty0 >= d[3]));
    } else {
      const float* d = det_lb + ((long long)b * g.max_det + di) * 4;
      const float bx1 = __fmul_rn(d[0], g.ratio_w), by1 = __fmul_rn(d[1], g.ratio_h);
      const float bx2 = __fmul_rn(d[2], g.ratio_w), by2 = __fmul_rn(d[3], g.ratio_h);
      any = any || !(((float)(g.left + sx_hi) < bx1) || ((float)(g.left + sx_lo) >= bx2) || ((float)(g.top + sy_hi) < by1) ||
                     ((float)(g.top + sy_lo) >= by2));
    }
  }
  if (!__syncthreads_or(any)) return;
  // ---- the tile's prototype window, once: s_p[k][py * ww + px] ----
  const float* pb = proto + (long long)b * g.mh * g.mw * g.nm;
  for (int i = threadIdx.x; i < npx; i += 256) {
    const int ry = i / ww, rx = i - ry * ww;
    const float4* pp = reinterpret_cast<const float4*>(pb + ((long long)(g.top + sy_lo + ry) * g.mw + (g.left + sx_lo + rx)) * g.nm);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 v = __ldg(pp + k);
      s_p[(4 * k + 0) * (kMaskWin * kMaskWin) + i] = v.x;
      s_p[(4 * k + 1) * (kMaskWin * kMaskWin) + i] = v.y;
      s_p[(4 * k + 2) * (kMaskWin * kMaskWin) + i] = v.z;
      s_p[(4 * k + 3) * (kMaskWin * kMaskWin) + i] = v.w;
    }
  }
  // ---- this thread's 16 output pixels: row oy, columns ox0 .. ox0 + 15; interpolation taps are detection-independent ----
  const int trow = threadIdx.x >> 2, tcol = (threadIdx.x & 3) * 16;
  const int oy = ty0 + trow, ox0 = tx0 + tcol;
  const float sy = src_of(min(oy, g.out_h - 1), g.scale_h);
  const int y0 = (int)sy;
  const int y1 = y0 + ((y0 < g.ch - 1) ? 1 : 0);
  const float ly = __fsub_rn(sy, (float)y0), hy = __fsub_rn(1.0f, ly);
  const int r0 = (y0 - sy_lo) * ww, r1 = (y1 - sy_lo) * ww;
  const bool vec_ok = ((g.out_w & 15) == 0) && ox0 + 16 <= g.out_w;
  // horizontal taps of the 16 pixels (detection-independent): window column of the left tap, whether the right tap is a
  // different column, and the two weights
  int xo[16];
  float lxs[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float sx = src_of(min(ox0 + j, g.out_w - 1), g.scale_w);
    const int x0 = (int)sx;
    lxs[j] = __fsub_rn(sx, (float)x0);
    xo[j] = ((x0 - sx_lo) << 1) | ((x0 < g.cw - 1) ? 1 : 0);
  }
  __syncthreads();
  int buf = 0;
  for (int di = 0; di < n_b; ++di) {
    const int slot = slot0 + di;
    if (slot >= capacity) break;
    float bx1, by1, bx2, by2;
    bool touch;
    if (g.retina) {
      const float* d = det + ((long long)b * g.max_det + di) * 6;
      bx1 = d[0]; by1 = d[1]; bx2 = d[2]; by2 = d[3];
      touch = !(((float)(tx0 + kMaskTile) <= bx1) || ((float)tx0 >= bx2) || ((float)(ty0 + kMaskTile) <= by1) || ((float)ty0 >= by2));
    } else {
      const float* d = det_lb + ((long long)b * g.max_det + di) * 4;
      bx1 = __fmul_rn(d[0], g.ratio_w); by1 = __fmul_rn(d[1], g.ratio_h);
      bx2 = __fmul_rn(d[2], g.ratio_w); by2 = __fmul_rn(d[3], g.ratio_h);
      touch = !(((float)(g.left + sx_hi) < bx1) || ((float)(g.left + sx_lo) >= bx2) || ((float)(g.top + sy_hi) < by1) ||
                ((float)(g.top + sy_lo) >= by2));
    }
    if (!touch) continue;  // CTA-uniform
    // window logits: one dot product per prototype pixel (fmaf chain k = 0..31, as mask_decode_kernel)
    const float4* cf = reinterpret_cast<const float4*>(coef + ((long long)b * g.max_det + di) * g.nm);
    float c[32];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 v = __ldg(cf + k);
      c[4 * k] = v.x; c[4 * k + 1] = v.y; c[4 * k + 2] = v.z; c[4 * k + 3] = v.w;
    }
    float* lg = s_logit + buf * (kMaskWin * kMaskWin);
    for (int i = threadIdx.x; i < npx; i += 256) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) acc = fmaf(c[k], s_p[k * (kMaskWin * kMaskWin) + i], acc);
      if (!g.retina) {  // ops.process_mask crops in proto space BEFORE the upsample
        const int ry = i / ww, rx = i - ry * ww;
        const float fx = (float)(g.left + sx_lo + rx), fy = (float)(g.top + sy_lo + ry);
        if (!(fx >= bx1 && fx < bx2 && fy >= by1 && fy < by2)) acc = 0.f;
      }
      lg[i] = acc;
    }
    __syncthreads();  // (also orders the reuse of the other logit buffer two detections later)
    buf ^= 1;
    if (oy >= g.out_h || ox0 >= g.out_w) continue;
    const bool row_in = g.retina ? ((float)oy >= by1 && (float)oy < by2) : true;
    if (g.retina && (!row_in || (float)(ox0 + 16) <= bx1 || (float)ox0 >= bx2)) continue;  // zeros already there
    uint32_t packed[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int ox = ox0 + j;
      const int c0 = xo[j] >> 1, c1 = c0 + (xo[j] & 1);
      const float lx = lxs[j], hx = __fsub_rn(1.0f, lx);
      const float top = __fadd_rn(__fmul_rn(hx, lg[r0 + c0]), __fmul_rn(lx, lg[r0 + c1]));
      const float bot = __fadd_rn(__fmul_rn(hx, lg[r1 + c0]), __fmul_rn(lx, lg[r1 + c1]));
      const float val = __fadd_rn(__fmul_rn(hy, top), __fmul_rn(ly, bot));
      bool on = val > 0.0f && row_in && ox < g.out_w;
      if (g.retina) on = on && ((float)ox >= bx1 && (float)ox < bx2);
      packed[j >> 2] |= (on ? 1u : 0u) << (8 * (j & 3));
    }
    uint8_t* o = out + (long long)slot * g.out_h * g.out_w + (long long)oy * g.out_w + ox0;
    if (vec_ok) {
      *reinterpret_cast<uint4*>(o) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    } else {
      for (int j = 0; j < 16 && ox0 + j < g.out_w; ++j) o[j] = (uint8_t)((packed[j >> 2] >> (8 * (j & 3))) & 1u);
    }
  }
}

// Generic twin for outputs SMALLER than the proto window (a frame under ~imgsz/4 with retina_masks=True: upstream's
// scale_masks simply down-samples): one thread per output pixel, four proto taps, the same dot-product order and ATen
// blend order as mask_decode_kernel.  Tiny outputs only (< proto size), so no staging is worth it.
__global__ void __launch_bounds__(256)
mask_decode_small_kernel(const float* __restrict__ proto, const float* __restrict__ coef, const float* __restrict__ det,
                         const float* __restrict__ det_lb, const int* __restrict__ offsets, int nB, int capacity, MaskGeom g,
                         uint8_t* __restrict__ out) {
  __shared__ float s_coef[32];
  const int slot = blockIdx.y;
  const int total = offsets[nB];
  if (slot >= total || slot >= capacity) return;
  int below = 0;
  for (int i0 = 0; i0 < nB; i0 += 32) {
    const int i = i0 + (threadIdx.x & 31);
    below += __popc(__ballot_sync(0xffffffffu, i < nB && offsets[i] <= slot));
  }
  const int b = below - 1, di = slot - offsets[b];
  if (threadIdx.x < g.nm) s_coef[threadIdx.x] = coef[((long long)b * g.max_det + di) * g.nm + threadIdx.x];
  __syncthreads();
  float bx1, by1, bx2, by2;
  if (g.retina) {
    const float* d = det + ((long long)b * g.max_det + di) * 6;
    bx1 = d[0]; by1 = d[1]; bx2 = d[2]; by2 = d[3];
  } else {
    const float* d = det_lb + ((long long)b * g.max_det + di) * 4;
    bx1 = __fmul_rn(d[0], g.ratio_w); by1 = __fmul_rn(d[1], g.ratio_h);
    bx2 = __fmul_rn(d[2], g.ratio_w); by2 = __fmul_rn(d[3], g.ratio_h);
  }
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.out_h * g.out_w) return;
  const int oy = idx / g.out_w, ox = idx - oy * g.out_w;
  auto src_of = [](int dst, float scale) {
    float s = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
    return s < 0.f ? 0.f : s;
  };
  const float sy = src_of(oy, g.scale_h), sx = src_of(ox, g.scale_w);
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + ((y0 < g.ch - 1) ? 1 : 0), x1 = x0 + ((x0 < g.cw - 1) ? 1 : 0);
  const float ly = __fsub_rn(sy, (float)y0), hy = __fsub_rn(1.0f, ly), lx = __fsub_rn(sx, (float)x0), hx = __fsub_rn(1.0f, lx);
  const float* pb = proto + (long long)b * g.mh * g.mw * g.nm;
  auto logit = [&](int ry, int rx) {
    const int py = g.top + ry, px = g.left + rx;
    const float4* pp = reinterpret_cast<const float4*>(pb + ((long long)py * g.mw + px) * g.nm);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 v = __ldg(pp + k);
      acc = fmaf(s_coef[4 * k + 0], v.x, acc);
      acc = fmaf(s_coef[4 * k + 1], v.y, acc);
      acc = fmaf(s_coef[4 * k + 2], v.z, acc);
      acc = fmaf(s_coef[4 * k + 3], v.w, acc);
    }
    if (!g.retina) {
      const float fx = (float)px, fy = (float)py;
      if (!(fx >= bx1 && fx < bx2 && fy >= by1 && fy < by2)) acc = 0.f;
    }
    return acc;
  };
  const float top = __fadd_rn(__fmul_rn(hx, logit(y0, x0)), __fmul_rn(lx, logit(y0, x1)));
  const float bot = __fadd_rn(__fmul_rn(hx, logit(y1, x0)), __fmul_rn(lx, logit(y1, x1)));
  const float val = __fadd_rn(__fmul_rn(hy, top), __fmul_rn(ly, bot));
  bool on = val > 0.0f;
  if (g.retina) on = on && (float)ox >= bx1 && (float)ox < bx2 && (float)oy >= by1 && (float)oy < by2;
  out[(long long)slot * g.out_h * g.out_w + idx] = on ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// Index-mask hand-off.  Replaces the per-detection Python loop of reference
// yolo_seg/yolo_with_deva.py:54-88 (`auto_segment`): for the detections of a frame in order, skip those whose mask
// area is below a threshold (`suppress_small_mask`), give the others ids 1, 2, ... and paint `output_mask[mask > 0.5] = id`
// (later detections overwrite earlier ones).  Three launches for a whole batch of frames instead of a `.sum()` host
// sync and a boolean scatter per detection.
// ------------------------------------------------------------------------------------------------
// area[i] += number of set pixels in a slice of mask i.  grid (slices, N), 256 threads, 16 bytes per thread per step.
__global__ void __launch_bounds__(256)
mask_area_kernel(const uint8_t* __restrict__ masks, long long hw, int* __restrict__ area) {
  const uint8_t* m = masks + (long long)blockIdx.y * hw;
  const long long per = (((hw + gridDim.x - 1) / gridDim.x) + 15) & ~15LL;  // slices start on 16-byte boundaries of the mask
  const long long lo = (long long)blockIdx.x * per, hi = min(lo + per, hw);
  unsigned sum = 0;
  long long i = lo + threadIdx.x * 16LL;
  if ((reinterpret_cast<uintptr_t>(m + lo) & 15) == 0) {
    for (; i + 16 <= hi; i += 256 * 16) {
      const uint4 v = *reinterpret_cast<const uint4*>(m + i);
      // bytes are 0/1: a multiply by 0x01010101 sums the four bytes of a word into its top byte
      sum += ((v.x * 0x01010101u) >> 24) + ((v.y * 0x01010101u) >> 24) + ((v.z * 0x01010101u) >> 24) + ((v.w * 0x01010101u) >> 24);
    }
    // tail of the slice (fewer than 16 bytes left for this thread's last step)
    for (long long j = i; j < hi && j < i + 16; ++j) sum += m[j];
  } else {
    for (long long j = lo + threadIdx.x; j < hi; j += 256) sum += m[j];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0 && sum) atomicAdd(area + blockIdx.y, (int)sum);
}

// ids[i] = running 1-based id of detection i inside its frame, 0 when suppressed (area < min_area; min_area < 0 keeps all)
__global__ void mask_ids_kernel(const int* __restrict__ offsets, int nB, const int* __restrict__ area, int min_area,
                                int* __restrict__ ids) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nB) return;
  int cur = 0;
  for (int i = offsets[b]; i < offsets[b + 1]; ++i) ids[i] = (min_area < 0 || area[i] >= min_area) ? ++cur : 0;
}

// The `min_side` branch of the hand-off (reference yolo_seg/yolo_with_deva.py:44-48,71-72): predict() ran on a resized
// frame, so every mask goes back to the frame size through torchvision `F.resize(mask[None], [h, w])` - bilinear WITH
// antialiasing (the default for tensors) - before the `mask.sum() < MIN_AREA` filter and the `mask > 0.5` paint.
// Restates ATen's _upsample_bilinear2d_aa: scale = in / out (float), support = max(scale, 1), taps
// [int(c - support + 0.5), int(c + support + 0.5)) around c = scale * (i + 0.5), triangle weights
// max(0, 1 - |(j + xmin - c + 0.5) / max(scale, 1)|) normalised by their sum; rows are reduced horizontally first,
// then vertically.  One thread per output pixel; writes bin = (value > 0.5) and adds the float values to area_f[mask].
__global__ void __launch_bounds__(256)
mask_resize_aa_kernel(const uint8_t* __restrict__ masks, int n, int h1, int w1, int H, int W, uint8_t* __restrict__ bins,
                      float* __restrict__ area_f) {
  const int i = blockIdx.z;
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  float v = 0.f;
  if (ox < W && oy < H) {
    const float sx = (float)w1 / (float)W, sy = (float)h1 / (float)H;
    const float supx = sx >= 1.f ? sx : 1.f, supy = sy >= 1.f ? sy : 1.f;
    const float invx = sx >= 1.f ? 1.f / sx : 1.f, invy = sy >= 1.f ? 1.f / sy : 1.f;
    const float cx = sx * ((float)ox + 0.5f), cy = sy * ((float)oy + 0.5f);
    const int xmin = max((int)(cx - supx + 0.5f), 0), xsize = min((int)(cx + supx + 0.5f), w1) - xmin;
    const int ymin = max((int)(cy - supy + 0.5f), 0), ysize = min((int)(cy + supy + 0.5f), h1) - ymin;
    float tx = 0.f, ty = 0.f;
    for (int j = 0; j < xsize; ++j) tx += fmaxf(0.f, 1.f - fabsf(((float)j + ((float)xmin - cx) + 0.5f) * invx));
    for (int j = 0; j < ysize; ++j) ty += fmaxf(0.f, 1.f - fabsf(((float)j + ((float)ymin - cy) + 0.5f) * invy));
    const uint8_t* m = masks + (long long)i * h1 * w1;
    for (int r = 0; r < ysize; ++r) {
      const uint8_t* row = m + (long long)(ymin + r) * w1 + xmin;
      float hs = 0.f;
      for (int j = 0; j < xsize; ++j) {
        float wx = fmaxf(0.f, 1.f - fabsf(((float)j + ((float)xmin - cx) + 0.5f) * invx));
        if (tx != 0.f) wx /= tx;
        hs += (float)row[j] * wx;
      }
      float wy = fmaxf(0.f, 1.f - fabsf(((float)r + ((float)ymin - cy) + 0.5f) * invy));
      if (ty != 0.f) wy /= ty;
      v += hs * wy;
    }
    bins[((long long)i * H + oy) * W + ox] = v > 0.5f ? 1 : 0;
  }
  // block sum of the float values -> one atomic per block
  __shared__ float s_part[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s_part[k];
    if (t != 0.f) atomicAdd(area_f + i, t);
  }
}

__global__ void mask_ids_f_kernel(const int* __restrict__ offsets, int nB, const float* __restrict__ area_f, float min_area,
                                  int* __restrict__ ids) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nB) return;
  int cur = 0;
  for (int i = offsets[b]; i < offsets[b + 1]; ++i) ids[i] = (min_area < 0.f || !(area_f[i] < min_area)) ? ++cur : 0;
}

// index_map[b][p] = id of the LAST kept detection of frame b whose mask covers pixel p, else 0.  8 pixels per thread.
__global__ void __launch_bounds__(256)
index_paint_kernel(const uint8_t* __restrict__ masks, const int* __restrict__ offsets, const int* __restrict__ ids, long long hw,
                   long long* __restrict__ index_map) {
  const int b = blockIdx.y;
  const long long p0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 8;
  if (p0 >= hw) return;
  const int lo = offsets[b], hi = offsets[b + 1];
  long long out[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool vec = p0 + 8 <= hw && ((hw & 7) == 0);
  for (int i = lo; i < hi; ++i) {
    const int id = ids[i];
    if (id == 0) continue;
    const uint8_t* m = masks + (long long)i * hw + p0;
    if (vec) {
      const uint2 v = *reinterpret_cast<const uint2*>(m);
      if ((v.x | v.y) == 0u) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if ((v.x >> (8 * j)) & 0xffu) out[j] = id;
        if ((v.y >> (8 * j)) & 0xffu) out[4 + j] = id;
      }
    } else {
      for (int j = 0; j < 8 && p0 + j < hw; ++j)
        if (m[j]) out[j] = id;
    }
  }
  long long* o = index_map + (long long)b * hw + p0;
  if (vec) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) *reinterpret_cast<longlong2*>(o + j) = make_longlong2(out[j], out[j + 1]);
  } else {
    for (int j = 0; j < 8 && p0 + j < hw; ++j) o[j] = out[j];
  }
}

// ---- the same hand-off when every mask is known to be zero outside a rectangle (the masks of predict() are cropped to
// their boxes): rects[i] = (x0, y0, x1, y1), x1 / y1 exclusive.  The area sum reads only the rectangle, and the paint
// works on 64 x 32 pixel tiles: a CTA first compacts (in detection order) the kept detections whose rectangle touches its
// tile, then every thread walks that short list BACKWARDS for its 8 pixels and stops as soon as all of them are assigned
// (the last kept detection covering a pixel wins, as in the reference's sequential overwrite).  Same index map, a
// fraction of the traffic: ~box area instead of n x H x W bytes, twice.  Needs W % 16 == 0 and 16-byte aligned masks.
__global__ void __launch_bounds__(256)
mask_area_rect_kernel(const uint8_t* __restrict__ masks, const int4* __restrict__ rects, int H, int W, int* __restrict__ area) {
  const int i = blockIdx.y;
  int4 r = rects[i];
  r.x = max(r.x, 0); r.y = max(r.y, 0); r.z = min(r.z, W); r.w = min(r.w, H);
  if (r.z <= r.x || r.w <= r.y) return;
  const uint8_t* m = masks + (long long)i * H * W;
  const int xa = r.x & ~15;
  const int cpr = (r.z - xa + 15) >> 4;  // 16-byte chunks per rectangle row (the last one may overhang: zeros there)
  const int total = (r.w - r.y) * cpr;
  unsigned sum = 0;
  for (int id = blockIdx.x * 256 + threadIdx.x; id < total; id += gridDim.x * 256) {
    const int row = id / cpr, c = id - row * cpr;
    const uint4 v = *reinterpret_cast<const uint4*>(m + (long long)(r.y + row) * W + xa + c * 16);
    sum += ((v.x * 0x01010101u) >> 24) + ((v.y * 0x01010101u) >> 24) + ((v.z * 0x01010101u) >> 24) + ((v.w * 0x01010101u) >> 24);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0 && sum) atomicAdd(area + i, (int)sum);
}

constexpr int kPaintTW = 64, kPaintTH = 32;  // 256 threads x 8 pixels
__global__ void __launch_bounds__(256)
index_paint_rect_kernel(const uint8_t* __restrict__ masks, const int* __restrict__ offsets, const int* __restrict__ ids,
                        const int4* __restrict__ rects, int H, int W, long long* __restrict__ index_map) {
  __shared__ int s_det[256];   // detections of this round that touch the tile (index into the batch), in order
  __shared__ int s_id[256];
  __shared__ int s_wcnt[8];
  const int b = blockIdx.z;
  const int tx0 = blockIdx.x * kPaintTW, ty0 = blockIdx.y * kPaintTH;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int y = ty0 + (threadIdx.x >> 3), x = tx0 + (threadIdx.x & 7) * 8;
  const bool inside = y < H && x < W;  // W % 16 == 0: a thread's 8 pixels are all inside or all outside
  const long long hw = (long long)H * W;
  const int lo = offsets[b], hi = offsets[b + 1];
  long long out[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned todo = inside ? 0xffu : 0u;  // pixels of this thread not yet assigned
  // rounds of 256 detections, LAST round first (later detections win)
  for (int r0 = lo + ((hi - lo - 1) / 256) * 256; r0 >= lo && hi > lo; r0 -= 256) {
    const int i = r0 + threadIdx.x;
    bool touch = false;
    int id = 0;
    if (i < hi) {
      id = ids[i];
      if (id) {
        const int4 r = rects[i];
        touch = r.x < tx0 + kPaintTW && r.z > tx0 && r.y < ty0 + kPaintTH && r.w > ty0;
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, touch);
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();
    int pos = __popc(bal & ((1u << lane) - 1u)), cnt = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (w < warp) pos += s_wcnt[w];
      cnt += s_wcnt[w];
    }
    if (touch) { s_det[pos] = i; s_id[pos] = id; }
    __syncthreads();
    for (int k = cnt - 1; k >= 0 && todo; --k) {
      const int d = s_det[k];
      const int4 r = rects[d];
      if (y < r.y || y >= r.w || x + 8 <= r.x || x >= r.z) continue;
      const uint2 v = *reinterpret_cast<const uint2*>(masks + (long long)d * hw + (long long)y * W + x);
      if ((v.x | v.y) == 0u) continue;
      const long long idv = s_id[k];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (((v.x >> (8 * j)) & 0xffu) && ((todo >> j) & 1u)) { out[j] = idv; todo &= ~(1u << j); }
        if (((v.y >> (8 * j)) & 0xffu) && ((todo >> (4 + j)) & 1u)) { out[4 + j] = idv; todo &= ~(1u << (4 + j)); }
      }
    }
    __syncthreads();  // s_det / s_id / s_wcnt are rewritten by the next round
  }
  if (inside) {
    long long* o = index_map + (long long)b * hw + (long long)y * W + x;
#pragma unroll
    for (int j = 0; j < 8; j += 2) *reinterpret_cast<longlong2*>(o + j) = make_longlong2(out[j], out[j + 1]);
  }
}

// ------------------------------------------------------------------------------------------------
// Needle length on the device.  The reference measures the needle as the long side of
// cv2.minAreaRect of the best detection's contour polygon (reference yolo_seg/app.py:97-105 ->
// utils/mask_tools.py:12-22 `get_coord_min_rect_len`), which costs a D2H copy of the full mask plus cv2.findContours
// per frame.  The minimum-area rectangle of the contour equals that of the convex hull of the mask's set pixels, and
// every hull vertex is the first or last set pixel of its row, so:
//   mask_row_extents_kernel : one warp per (mask, row) -> (xmin, xmax) of the set pixels of that row;
//   mask_min_rect_kernel    : one CTA per mask -> convex hull of the <= 2H extent points (monotone chain, integer
//                             cross products) and, for every hull edge in parallel, the bounding rectangle aligned
//                             with it; the smallest area wins (what rotating calipers enumerates).
// Output per mask: (length = long side, ratio = long / max(short, 1)), the reference's return values.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
mask_row_extents_kernel(const uint8_t* __restrict__ masks, int H, int W, int2* __restrict__ ext) {
  const int lane = threadIdx.x & 31;
  const int y = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (y >= H) return;
  const uint8_t* row = masks + ((long long)blockIdx.y * H + y) * W;
  int lo = 1 << 30, hi = -1;
  for (int x = lane; x < W; x += 32) {
    if (row[x]) {
      lo = min(lo, x);
      hi = max(hi, x);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) ext[(long long)blockIdx.y * H + y] = make_int2(lo, hi);
}

constexpr int kRectMaxPts = 2 * 2304;  // rows of the largest supported frame (2304) x 2

__global__ void __launch_bounds__(256)
mask_min_rect_kernel(const int2* __restrict__ ext, int H, float* __restrict__ out /*(n,2)*/) {
  __shared__ short2 s_pts[kRectMaxPts];
  __shared__ short2 s_hull[kRectMaxPts + 2];
  __shared__ int s_n, s_h;
  __shared__ float s_best[256], s_w[256], s_hh[256];
  const int2* e = ext + (long long)blockIdx.x * H;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int y = 0; y < H; ++y) {  // points sorted by (y, x)
      const int2 v = e[y];
      if (v.y < 0) continue;
      s_pts[n++] = make_short2((short)v.x, (short)y);
      if (v.y != v.x) s_pts[n++] = make_short2((short)v.y, (short)y);
    }
    s_n = n;
    // Andrew's monotone chain (y-major order): integer cross products, collinear points dropped
    int k = 0;
    auto cross = [](short2 o, short2 a, short2 b) { return (a.x - o.x) * (b.y - o.y) - (a.y - o.y) * (b.x - o.x); };
    for (int i = 0; i < n; ++i) {
      while (k >= 2 && cross(s_hull[k - 2], s_hull[k - 1], s_pts[i]) <= 0) --k;
      s_hull[k++] = s_pts[i];
    }
    for (int i = n - 2, t = k + 1; i >= 0; --i) {
      while (k >= t && cross(s_hull[k - 2], s_hull[k - 1], s_pts[i]) <= 0) --k;
      s_hull[k++] = s_pts[i];
    }
    s_h = n > 1 ? k - 1 : n;  // the last point repeats the first
  }
  __syncthreads();
  const int h = s_h;
  float best = 3.0e38f, bw = 0.f, bh = 0.f;
  if (h >= 3) {
    for (int i = threadIdx.x; i < h; i += 256) {
      const short2 a = s_hull[i], b = s_hull[(i + 1) % h];
      const float ex = (float)(b.x - a.x), ey = (float)(b.y - a.y);
      const float inv = rsqrtf(ex * ex + ey * ey);
      const float ux = ex * inv, uy = ey * inv;
      float umin = 3.0e38f, umax = -3.0e38f, vmin = 3.0e38f, vmax = -3.0e38f;
      for (int j = 0; j < h; ++j) {
        const float px = (float)s_hull[j].x, py = (float)s_hull[j].y;
        const float pu = px * ux + py * uy, pv = py * ux - px * uy;
        umin = fminf(umin, pu); umax = fmaxf(umax, pu);
        vmin = fminf(vmin, pv); vmax = fmaxf(vmax, pv);
      }
      const float w = umax - umin, hh = vmax - vmin;
      if (w * hh < best) { best = w * hh; bw = w; bh = hh; }
    }
  }
  s_best[threadIdx.x] = best; s_w[threadIdx.x] = bw; s_hh[threadIdx.x] = bh;
  __syncthreads();
  if (threadIdx.x == 0) {
    float length = 0.f, ratio = 0.f;
    if (h >= 3) {
      int arg = 0;
      for (int i = 1; i < 256; ++i) if (s_best[i] < s_best[arg]) arg = i;
      const float lo = fminf(s_w[arg], s_hh[arg]);
      length = fmaxf(s_w[arg], s_hh[arg]);
      ratio = length / (lo == 0.f ? 1.f : lo);
    } else if (h == 2) {  // all set pixels on one line: a degenerate rectangle of zero width
      const float dx = (float)(s_hull[1].x - s_hull[0].x), dy = (float)(s_hull[1].y - s_hull[0].y);
      length = sqrtf(dx * dx + dy * dy);
      ratio = length;
    }
    out[2 * blockIdx.x] = length;
    out[2 * blockIdx.x + 1] = ratio;
  }
}

}  // namespace ypb
