// Host-side planning of one fused conv launch: tile shape, TMA tensor maps, kernel parameters.
// (Single translation unit build: this file is included once, by ypb200.cu.)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "conv_tc.cuh"
#include "conv_halo2.cuh"

namespace ypb {

// Everything needed to describe one conv on raw device pointers (engine buffers or test tensors).
struct ConvDesc {
  // input: NHWC bf16 buffer (B, Hin, Win, in_ctot), channels [in_c_off, in_c_off + cin)
  const void* in = nullptr;
  int B = 0, Hin = 0, Win = 0, in_ctot = 0, in_c_off = 0, cin = 0;
  // weights: bf16 [k*k][cout][cin] (ConvTranspose 2x2: [1][4*cq][cin], cout = 4*cq), bias fp32 [cout]
  const void* wg = nullptr;
  const float* bias = nullptr;
  int cout = 0, k = 1, stride = 1, act = 1;
  // output
  int out_mode = OUT_BF16;
  void* out = nullptr;
  long long out_img_stride = 0;  // elements between images
  int out_pix_stride = 0;        // elements between pixels
  int out_c_off = 0;
  // optional residual (bf16, indexed like a bf16 NHWC output)
  const void* res = nullptr;
  long long res_img_stride = 0;
  int res_pix_stride = 0, res_c_off = 0;
};

struct ConvLaunch {
  CUtensorMap tmHalo;  // impl 3 experiment: halo box {64, 10, 18}
  bool halo_ok = false;
  // halo-reuse kernel for 3x3 stride-1 convs (conv3_halo_kernel)
  bool use_halo = false;
  Conv3Extra x3{};
  CUtensorMap tmHalo3, tmB3;  // halo box; weight box {64, n_tile, b_group}
  int tiles_h3 = 0, tiles_w3 = 0, total_tiles3 = 0, smem3 = 0;
  // CTA-pair kernel (conv3_halo2_kernel): 3x3 stride-1 layers whose weights do not fit in shared memory
  bool use_pair = false;
  Conv3Pair x2{};
  CUtensorMap tmHalo2, tmB2;  // halo box {64, 10, 18}; weight box {64, n_tile / 2, 3}
  int tiles_h2 = 0, tiles_w2 = 0, total_pairs = 0, smem_pair = 0;
  // CTA-pair flavour of the per-tap kernel (conv_tc2p_kernel): 1x1 and stride-2 3x3 layers
  bool use_pair_tc2 = false;
  CUtensorMap tmB2p;  // weight box {64, n_tile / 2, 1}
  int total_pairs_tc2 = 0, stages2p = 2, smem2p = 0;
  ConvParams p;       // persistent kernels (conv_tc2 geometry: tile may hold msub sub-tiles)
  ConvParams p1;      // one-tile-per-CTA geometry (impl 2 / 3, A/B experiments)
  CUtensorMap tmA1;
  dim3 grid1;
  int n_splits = 1, total_tiles = 0, stages2 = 2, smem2 = 0;  // persistent-kernel launch shape
#if YPB_DIAG
  ConvSimtGeom sg;
#endif
  CUtensorMap tmA, tmB;
  CUtensorMap tmO;  // output map of the TMA-store epilogue (1x1 convs), valid when p.tma_out != 0
  dim3 grid;
  int smem = 0;
  double flops = 0;
  int oH = 0, oW = 0;
};

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

static bool encode_map(CUtensorMap* m, CUtensorMapDataType dt, CUtensorMapSwizzle sw, const void* base, int rank,
                       const cuuint64_t* dims, const cuuint64_t* strides_bytes /*rank-1*/, const cuuint32_t* box, std::string* err);
static bool encode_bf16_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
                            const cuuint64_t* strides_bytes /*rank-1*/, const cuuint32_t* box, std::string* err) {
  return encode_map(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B, base, rank, dims, strides_bytes, box, err);
}
static bool encode_map(CUtensorMap* m, CUtensorMapDataType dt, CUtensorMapSwizzle sw, const void* base, int rank,
                       const cuuint64_t* dims, const cuuint64_t* strides_bytes /*rank-1*/, const cuuint32_t* box, std::string* err) {
  PFN_tmapEncodeTiled enc = tmap_encoder();
  if (!enc) {
    *err = "cuTensorMapEncodeTiled unavailable (no CUDA driver)";
    return false;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) rank %d dims %llu,%llu,%llu box %u,%u,%u", (int)r, rank,
             (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0), box[0],
             box[1], rank > 2 ? box[2] : 0);
    *err = buf;
    return false;
  }
  return true;
}

// Pick the CTA rectangle (TH x TW <= 128 output pixels) that covers an H x W map with the fewest tiles.
static void pick_tile(int H, int W, int* TH, int* TW) {
  long best_tiles = -1;
  int bh = 1, bw = 128, best_shape = 1 << 30;
  for (int tw = 4; tw <= 128; ++tw) {
    int th = 128 / tw;
    if (th > H) th = H;
    if (th < 1) continue;
    int twc = tw > W ? W : tw;
    const long tiles = (long)((H + th - 1) / th) * ((W + twc - 1) / twc);
    const int shape = th > twc ? th - twc : twc - th;
    if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && shape < best_shape)) {
      best_tiles = tiles; bh = th; bw = twc; best_shape = shape;
    }
  }
  *TH = bh; *TW = bw;
}

// Same for CTA tiles made of `msub` vertically stacked sub-tiles (each <= 128 rows, a multiple of 8 rows so that it
// starts on a swizzle-atom boundary of the A stage).  Returns the number of 128-row MMA slots spent per image.
static long pick_tile_msub(int H, int W, int msub, int* TH_sub, int* TW) {
  long best = -1;
  int best_shape = 1 << 30;
  for (int tw = 4; tw <= 128; ++tw) {
    const int twc = tw > W ? W : tw;
    for (int th = 128 / twc; th >= 1; --th) {
      if ((th * twc) % 8) continue;
      const long slots = (long)((H + msub * th - 1) / (msub * th)) * ((W + twc - 1) / twc) * msub;
      const int shape = th > twc ? th - twc : twc - th;
      if (best < 0 || slots < best || (slots == best && shape < best_shape)) { best = slots; best_shape = shape; *TH_sub = th; *TW = twc; }
      break;  // the tallest admissible th for this width is the only candidate worth considering
    }
  }
  return best;
}

// Geometry only (no pointers, no driver calls): usable on a box without a GPU.
static void conv3_set_taps(Conv3Extra* x, bool s2) {
  x->s2 = s2 ? 1 : 0;
  for (int t = 0; t < 9; ++t) {
    const int kh = t / 3, kw = t % 3;
    if (!s2) {
      x->tap_off[t] = ((kh * 10 + kw) * 128) >> 4;
    } else {
      const int r = kh > 0, ph = kh != 1, wq = kw > 0, half = kw != 1;
      x->tap_off[t] = ((((r * 2 + ph) * 10 + wq) * 128) + half * 64) >> 4;
    }
  }
  x->sub_off = (s2 ? 16 * 2 * 10 * 128 : 16 * 10 * 128) >> 4;
  x->a_hi = umma_desc_hi(s2 ? 2560 : 1280);
}

static bool conv_plan_geometry(const ConvDesc& d, ConvLaunch* L, std::string* err) {
  ConvParams& p = L->p;
  memset(&p, 0, sizeof p);
  if (d.cin % 16 || d.cout % 16 || d.in_ctot % 8 || d.in_c_off % 8) { *err = "conv channels must be multiples of 16"; return false; }
  if (!((d.k == 1 && d.stride == 1) || (d.k == 3 && (d.stride == 1 || d.stride == 2)))) { *err = "unsupported conv k/stride"; return false; }
  if (d.stride == 2 && ((d.Hin | d.Win) & 1)) { *err = "stride-2 conv needs even input dims"; return false; }
  const int oH = d.Hin / d.stride, oW = d.Win / d.stride;
  L->oH = oH; L->oW = oW;
  p.Cin = d.cin; p.Cout = d.cout; p.ntaps = d.k * d.k;
  p.img_HW = oH * oW; p.img_W = oW;
  if (d.k == 1) {  // flat: 128 consecutive pixels of the flattened batch
    p.tB = 1; p.tH = 1; p.tW = d.B * oH * oW;
    p.TH = 1; p.TW = 128;
    p.tiles_h = 1; p.tiles_w = (p.tW + 127) / 128;
    p.a_base[0] = d.in_c_off; p.a_cw[1] = 1;
    L->grid.x = p.tiles_w;
  } else {
    p.tB = d.B; p.tH = oH; p.tW = oW;
    pick_tile(oH, oW, &p.TH, &p.TW);
    p.tiles_h = (oH + p.TH - 1) / p.TH; p.tiles_w = (oW + p.TW - 1) / p.TW;
    p.a_base[0] = d.in_c_off;
    if (d.stride == 1) {  // view (C, W, H, B, 1)
      p.a_cw[1] = 1; p.a_ch[2] = 1; p.a_cb[3] = 1;
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) { int* t = p.tap[kh * 3 + kw]; t[1] = kw - 1; t[2] = kh - 1; }
    } else {              // view (2C, W/2, 2, H/2, B): input row 2*oh+kh-1 = 2*(oh+dh)+ph
      p.a_cw[1] = 1; p.a_ch[3] = 1; p.a_cb[4] = 1;
      const int dd[3] = {-1, 0, 0}, par[3] = {1, 0, 1};
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          int* t = p.tap[kh * 3 + kw];
          t[0] = par[kw] * d.in_ctot; t[1] = dd[kw]; t[2] = par[kh]; t[3] = dd[kh];
        }
    }
    L->grid.x = d.B * p.tiles_h * p.tiles_w;
  }
  int splits = 1;
  while (d.cout / splits > 256 || d.cout % (16 * splits)) {
    ++splits;
    if (splits > d.cout / 16) { *err = "cannot split Cout"; return false; }
  }
  // Few output tiles (small batches, the deep 20x20 / 40x40 levels): a 128-row tile x 256 channels is ONE CTA walking the
  // whole K loop at 128 cycles per K-step (3x3, 256 -> 256 at B = 1: 4 CTAs busy for ~10 us while 144 SMs idle).  Narrower
  // channel tiles put more CTAs on the layer and shorten that chain in proportion; the A tile is re-read from L2 once per
  // split, which costs less than the chain as long as a split keeps >= 64 channels.  Throughput shapes (tiles >= half
  // the SMs) are untouched.
  if (!getenv("YPB_NO_LAT_SPLIT")) {
    const long m_tiles0 = (long)L->grid.x;
    while (m_tiles0 * splits < 74 && d.cout / (2 * splits) >= 64 && d.cout % (32 * splits) == 0) splits *= 2;
  }
  p.n_tile = d.cout / splits;
  L->grid.y = splits; L->grid.z = 1;
  const int k_iters = p.ntaps * ((d.cin + 63) / 64);
  int st = (110 * 1024) / conv_stage_bytes(p.n_tile);
  if (st < 2) st = 2;
  if (st > 6) st = 6;
  if (st > k_iters) st = k_iters;
  p.stages = st;
  p.msub = 1; p.sub_rows = p.TH * p.TW;
  L->smem = conv_smem_bytes(p.n_tile, st);
  L->p1 = p;           // one-tile-per-CTA kernels keep the 128-row geometry
  L->grid1 = L->grid;
  // persistent kernel: two 128-row sub-tiles per CTA tile when TMEM has room for 2 x 2 accumulators and the layer has
  // plenty of tiles: every weight tile is fetched once per 256 rows and the per-tile hand-offs are halved
  {
    const long m_tiles = (long)L->grid.x;
    const bool room = 4 * conv2_acc_stride(p.n_tile) <= 512;
    if (room && m_tiles >= 4 * 148 && !getenv("YPB_NO_MSUB")) {
      if (d.k == 1) {
        p.msub = 2; p.sub_rows = 128; p.TW = 256;
        p.tiles_w = (p.tW + 255) / 256;
        L->grid.x = p.tiles_w;
      } else {
        int ths = 0, tw2 = 0;
        const long slots2 = pick_tile_msub(oH, oW, 2, &ths, &tw2);
        if (slots2 > 0 && slots2 * 10 <= (long)p.tiles_h * p.tiles_w * 11) {  // at most 10% more MMA slots than msub = 1
          p.msub = 2; p.sub_rows = ths * tw2;
          p.TH = 2 * ths; p.TW = tw2;
          p.tiles_h = (oH + p.TH - 1) / p.TH; p.tiles_w = (oW + p.TW - 1) / p.TW;
          L->grid.x = d.B * p.tiles_h * p.tiles_w;
        }
      }
    }
  }
  // one CTA per SM, ring as deep as ~150 KB of smem allows (>= 120 KB so CTAs never co-reside)
  L->n_splits = splits;
  L->total_tiles = (int)L->grid.x * splits;
  int ring_kb = 150;
  if (const char* ev = getenv("YPB_RING_KB")) ring_kb = atoi(ev);  // tuning knob for experiments
  const int stage2 = p.msub * kATileBytes + p.n_tile * 128;
  int st2 = (ring_kb * 1024) / stage2;
  if (st2 > 8) st2 = 8;
  if (st2 < 2) st2 = 2;
  L->stages2 = st2;
  const int epi_smem = kEpiWarps * epi_stage_bytes(d.out_mode == OUT_F32);
  const int epi_smem2 = kEpiWarps * epi_stage_bytes(d.out_mode == OUT_F32, d.k == 1) + 1024;  // conv_tc2: 1 KB-aligned tiles
  L->smem2 = 1024 + st2 * stage2 + 256 + epi_smem2 + conv_bias_smem(d.cout);
  if (L->smem2 < 120 * 1024) L->smem2 = 120 * 1024;
  p.out_mode = d.out_mode; p.act = d.act;
  p.dbg = getenv("YPB_DBG") ? atoi(getenv("YPB_DBG")) : 0;
  p.bo_prod = p.bo_mma_acc = p.bo_mma_full = p.bo_epi = 0;
  if (const char* ev = getenv("YPB_BO")) sscanf(ev, "%d,%d,%d,%d", &p.bo_prod, &p.bo_mma_acc, &p.bo_mma_full, &p.bo_epi);
  p.out_img_stride = d.out_img_stride; p.out_pix_stride = d.out_pix_stride; p.out_c_off = d.out_c_off;
  p.res_img_stride = d.res_img_stride; p.res_pix_stride = d.res_pix_stride; p.res_c_off = d.res_c_off;
  L->flops = 2.0 * d.B * oH * oW * (double)d.cout * d.cin * d.k * d.k;
  L->use_halo = false;
  if (d.k == 3 && d.stride == 1 && !getenv("YPB_NO_HALO")) {
    // Operand bytes fetched from L2 per launch are what bounds these kernels (~7 TB/s L2->SM on B200): pick the
    // cheapest of {per-tap boxes, halo tiles with msub = 1 or 2, resident or streamed weights}.
    const int kch = (d.cin + 63) / 64;
    const long avail = 227 * 1024 - 1024 - 512 - epi_smem - conv_bias_smem(d.cout);
    const long b_slot = (long)p.n_tile * 128, b_total = 9L * kch * b_slot;
    const double waste_taps = (double)L->total_tiles * p.msub * 128 - (double)d.B * oH * oW * splits;
    const double cost_taps = (double)L->total_tiles * k_iters * (p.msub * kATileBytes + b_slot) + waste_taps * 9.0 * kch * 64.0;
    double best = cost_taps;
    // Sub-tiles per CTA tile: measured (tools/conv_layers.py with YPB_HALO_MSUB=1/2, B200) a tile costs about
    // (msub + 0.7) sub-tile times - the hand-offs and the pipeline refill of a tile are worth ~0.7 of a 128-row
    // sub-tile - and a CTA runs ceil(tiles / SMs) of them, so two stacked sub-tiles win unless their padded rows cost
    // more than that (a 40-row map tiled 32 rows high wastes 37 % of its MMAs; 16-row tiles win there by 9 %).
    int msub_pick = 0;
    {
      double tbest = 1e30;
      for (int msub = 2; msub >= 1; --msub) {
        if (2 * msub * conv2_acc_stride(p.n_tile) > 512) continue;
        const long tiles3 = (long)d.B * ((oH + 16 * msub - 1) / (16 * msub)) * ((oW + 7) / 8) * splits;
        const double t = std::ceil(tiles3 / 148.0) * (msub + 0.7);
        if (t < tbest) { tbest = t; msub_pick = msub; }
      }
    }
    const int force_msub = getenv("YPB_HALO_MSUB") ? atoi(getenv("YPB_HALO_MSUB")) : 0;
    for (int msub = 1; msub <= 2; ++msub) {
      if (2 * msub * conv2_acc_stride(p.n_tile) > 512) continue;
      if (force_msub && msub != force_msub && 2 * force_msub * conv2_acc_stride(p.n_tile) <= 512) continue;
      if (!force_msub && !getenv("YPB_PLAN_BYTES") && msub_pick && msub != msub_pick) continue;
      const int halo_rows = (16 * msub + 2) * 10;
      const long a_bytes = ((long)halo_rows * 128 + 1023) & ~1023L;
      const int th3 = (oH + 16 * msub - 1) / (16 * msub), tw3 = (oW + 7) / 8;
      const long tiles3 = (long)d.B * th3 * tw3 * splits;
      for (int stat = 1; stat >= 0; --stat) {
        if (stat && (splits != 1 || b_total > 96 * 1024 || b_total + 2 * a_bytes > avail)) continue;
        long a_slots, b_slots = 0, b_bytes;
        int b_group = 3;
        if (stat) {
          b_bytes = b_total;
          a_slots = (avail - b_total) / a_bytes;
        } else {
          a_slots = 2;
          if ((avail - a_slots * a_bytes) / (3 * b_slot) < 2) b_group = 1;  // a kernel row of taps per box when two fit
          if (const char* ev = getenv("YPB_BGROUP")) b_group = atoi(ev) == 1 ? 1 : 3;  // experiment knob
          b_slots = (avail - a_slots * a_bytes) / (b_group * b_slot);
          if (b_slots > 12) b_slots = 12;
          if (b_slots < 2) continue;
          b_bytes = b_slots * b_group * b_slot;
        }
        if (a_slots > 4) a_slots = 4;
        if (a_slots < 2) continue;
        // bytes fetched from L2 + a charge for tensor work wasted on padded rows (128 B-equivalents per MMA row-step)
        const double waste = (double)tiles3 * msub * 128 - (double)d.B * oH * oW * splits;
        const double cost = (double)tiles3 * kch * (halo_rows * 128.0 + (stat ? 0.0 : 9.0 * b_slot)) +
                            (stat ? 148.0 * b_total : 0.0) + waste * 9.0 * kch * 64.0;
        if (cost < best) {
          best = cost;
          L->use_halo = true;
          L->x3.msub = msub; L->x3.a_slots = (int)a_slots; L->x3.a_bytes = (int)a_bytes; L->x3.halo_rows = halo_rows;
          L->x3.b_slots = (int)b_slots; L->x3.b_group = b_group; L->x3.b_stat = stat; L->x3.b_bytes = (int)b_bytes;
          L->tiles_h3 = th3; L->tiles_w3 = tw3; L->total_tiles3 = (int)tiles3;
          L->smem3 = (int)(1024 + a_slots * a_bytes + b_bytes + 512 + epi_smem + conv_bias_smem(d.cout));
          conv3_set_taps(&L->x3, false);
        }
      }
    }
  }
  if (d.k == 3 && d.stride == 2 && d.cin == 32 && d.in_ctot == 32 && d.in_c_off == 0 && splits == 1 && !getenv("YPB_NO_HALO") &&
      !getenv("YPB_NO_HALO_S2")) {
    // Stride 2 with 32 input channels (the layer after the stem): a pixel PAIR is one 128-byte row of the
    // (2C, W/2, 2, H/2, B) view, so the halo trick carries over with a parity-split box {64, 10, 2, 17} per 8 x 16
    // output tile and resident weights: L2 -> smem traffic drops ~3x against nine per-tap boxes plus per-tile weight
    // boxes, which is what bounded this layer (1.4 GB per launch at 640x640, B = 64).
    const long avail = 227 * 1024 - 1024 - 512 - epi_smem - conv_bias_smem(d.cout);
    const long b_total = 9L * p.n_tile * 128;
    const int halo_rows = 17 * 2 * 10;
    const long a_bytes = ((long)halo_rows * 128 + 1023) & ~1023L;
    long a_slots = (avail - b_total) / a_bytes;
    if (a_slots > 4) a_slots = 4;
    if (b_total <= 96 * 1024 && a_slots >= 2 && 2 * conv2_acc_stride(p.n_tile) <= 512) {
      const int th3 = (oH + 15) / 16, tw3 = (oW + 7) / 8;
      L->use_halo = true;
      L->x3.msub = 1; L->x3.a_slots = (int)a_slots; L->x3.a_bytes = (int)a_bytes; L->x3.halo_rows = halo_rows;
      L->x3.b_slots = 0; L->x3.b_group = 3; L->x3.b_stat = 1; L->x3.b_bytes = (int)b_total;
      L->tiles_h3 = th3; L->tiles_w3 = tw3; L->total_tiles3 = d.B * th3 * tw3;
      L->smem3 = (int)(1024 + a_slots * a_bytes + b_total + 512 + epi_smem + conv_bias_smem(d.cout));
      conv3_set_taps(&L->x3, true);
    }
  }
  L->use_pair = false;
  if (L->use_halo && !L->x3.b_stat && !L->x3.s2 && d.k == 3 && d.stride == 1 && (p.n_tile % 16) == 0 && !getenv("YPB_NO_PAIR")) {
    // Streamed weights: let the two SMs of a TPC share every weight tile (conv_halo2.cuh).  Each CTA keeps a 16 x 8
    // pixel tile; the pair's tiles are consecutive in the (image, row, column) order.
    const int th1 = (oH + 15) / 16, tw1 = (oW + 7) / 8;
    const long m_tiles = (long)d.B * th1 * tw1, pairs_m = (m_tiles + 1) / 2, total_pairs = pairs_m * splits;
    const long avail = 227 * 1024 - 1024 - 512 - kEpiWarps * epi_stage_bytes(d.out_mode == OUT_F32) - conv_bias_smem(d.cout);
    const long a_bytes = (180L * 128 + 1023) & ~1023L, grp = 3L * (p.n_tile / 2) * 128;
    long a_slots = 3;
    long b_slots = (avail - a_slots * a_bytes) / grp;
    if (b_slots < 3) { a_slots = 2; b_slots = (avail - a_slots * a_bytes) / grp; }
    if (b_slots > 12) b_slots = 12;
    const int min_pairs = getenv("YPB_PAIR_MIN") ? atoi(getenv("YPB_PAIR_MIN")) : 74;
    if (b_slots >= 2 && total_pairs >= min_pairs && 2 * conv2_acc_stride(p.n_tile) <= 512) {
      L->use_pair = true;
      L->x2.a_slots = (int)a_slots; L->x2.a_bytes = (int)a_bytes; L->x2.b_slots = (int)b_slots; L->x2.grp_bytes = (int)grp;
      L->x2.m_tiles = (int)m_tiles;
      Conv3Extra taps;
      conv3_set_taps(&taps, false);
      for (int t = 0; t < 9; ++t) L->x2.tap_off[t] = taps.tap_off[t];
      L->x2.a_hi = taps.a_hi;
      L->tiles_h2 = th1; L->tiles_w2 = tw1; L->total_pairs = (int)total_pairs;
      L->smem_pair = (int)(1024 + a_slots * a_bytes + b_slots * grp + 512 + kEpiWarps * epi_stage_bytes(d.out_mode == OUT_F32) +
                           conv_bias_smem(d.cout));
    }
  }
  L->use_pair_tc2 = false;
  if (!L->use_halo && (p.n_tile % 16) == 0 && !getenv("YPB_NO_PAIR") && !getenv("YPB_NO_PAIR_TC2")) {
    const long m_tiles = (long)L->grid.x, total_pairs = ((m_tiles + 1) / 2) * splits;
    const int min_pairs = getenv("YPB_PAIR_MIN") ? atoi(getenv("YPB_PAIR_MIN")) : 74;
    const int stage = p.msub * kATileBytes + (p.n_tile / 2) * 128;
    int ring_kb = 150;
    if (const char* ev = getenv("YPB_RING_KB")) ring_kb = atoi(ev);
    int st = (ring_kb * 1024) / stage;
    if (st > 8) st = 8;
    if (st < 2) st = 2;
    // Pairing couples the two CTAs stage by stage (cross-CTA barrier round trips): it pays when a tile is a long K loop
    // over wide weight tiles, and loses on the thin, HBM-bound 1x1 layers.  Measured on B200 (tools/conv_layers.py,
    // yolov8s/x-seg shapes): wins from k_iters x n_tile >= 12 x 256 (model.12.cv1 +4 %, model.5 +8 %, model.7 +12 %,
    // x.4.cv2 +10 %, x.3 +15 %), loses below (model.2.cv1 -34 %, proto.upsample -16 %, model.16 -3 %).
    const long pair_work = (long)k_iters * p.n_tile;
    const long pair_min_work = getenv("YPB_PAIR_WORK") ? atol(getenv("YPB_PAIR_WORK")) : 2560;
    if (total_pairs >= min_pairs && (pair_work >= pair_min_work || getenv("YPB_PAIR_MIN"))) {
      L->use_pair_tc2 = true;
      L->total_pairs_tc2 = (int)total_pairs;
      L->stages2p = st;
      L->smem2p = 1024 + st * stage + 256 + kEpiWarps * epi_stage_bytes(d.out_mode == OUT_F32, d.k == 1) + 1024 + conv_bias_smem(d.cout);
      if (L->smem2p < 120 * 1024) L->smem2p = 120 * 1024;
    }
  }
#if YPB_DIAG
  ConvSimtGeom& g = L->sg;
  memset(&g, 0, sizeof g);
  g.in_H = d.Hin; g.in_W = d.Win; g.in_ctot = d.in_ctot; g.in_c_off = d.in_c_off;
  g.k = d.k; g.stride = d.stride; g.pad = d.k / 2; g.oH = oH; g.oW = oW; g.nB = d.B;
#endif
  {  // p1 = p with the one-tile-per-CTA geometry saved above
    const ConvParams g1 = L->p1;
    L->p1 = p;
    L->p1.TH = g1.TH; L->p1.TW = g1.TW; L->p1.tiles_h = g1.tiles_h; L->p1.tiles_w = g1.tiles_w;
    L->p1.msub = 1; L->p1.sub_rows = g1.sub_rows; L->p1.stages = g1.stages;
  }
  return true;
}

// Fills pointers and encodes the TMA descriptors (needs the CUDA driver).
static bool conv_bind(const ConvDesc& d, ConvLaunch* L, std::string* err) {
  ConvParams& p = L->p;
  p.out = d.out; p.bias = d.bias; p.res = reinterpret_cast<const __nv_bfloat16*>(d.res);
  L->p1.out = d.out; L->p1.bias = d.bias; L->p1.res = p.res;
#if YPB_DIAG
  L->sg.in = reinterpret_cast<const __nv_bfloat16*>(d.in);
  L->sg.wg = reinterpret_cast<const __nv_bfloat16*>(d.wg);
#endif
  const cuuint64_t C = (cuuint64_t)d.in_ctot;
  cuuint64_t dims[5], str[4];
  cuuint32_t box[5];
  if (d.k == 1) {
    dims[0] = C; dims[1] = (cuuint64_t)d.B * d.Hin * d.Win; dims[2] = dims[3] = dims[4] = 1;
    str[0] = C * 2; str[1] = str[2] = str[3] = dims[1] * C * 2;
    box[0] = 64; box[1] = (cuuint32_t)p.TW; box[2] = box[3] = box[4] = 1;  // 128 or 256 pixels per CTA tile
  } else if (d.stride == 1) {
    dims[0] = C; dims[1] = d.Win; dims[2] = d.Hin; dims[3] = d.B; dims[4] = 1;
    str[0] = C * 2; str[1] = str[0] * d.Win; str[2] = str[1] * d.Hin; str[3] = str[2] * d.B;
    box[0] = 64; box[1] = p.TW; box[2] = p.TH; box[3] = 1; box[4] = 1;
  } else {
    dims[0] = 2 * C; dims[1] = d.Win / 2; dims[2] = 2; dims[3] = d.Hin / 2; dims[4] = d.B;
    str[0] = 2 * C * 2; str[1] = (cuuint64_t)d.Win * C * 2; str[2] = 2 * str[1]; str[3] = (cuuint64_t)d.Hin * d.Win * C * 2;
    box[0] = 64; box[1] = p.TW; box[2] = 1; box[3] = p.TH; box[4] = 1;
  }
  if (!encode_bf16_map(&L->tmA, d.in, 5, dims, str, box, err)) return false;
  {  // same view, 128-row box, for the one-tile-per-CTA kernels
    cuuint32_t box1[5] = {box[0], box[1], box[2], box[3], box[4]};
    if (d.k == 1) box1[1] = 128;
    else if (d.stride == 1) { box1[1] = L->p1.TW; box1[2] = L->p1.TH; }
    else { box1[1] = L->p1.TW; box1[3] = L->p1.TH; }
    if (!encode_bf16_map(&L->tmA1, d.in, 5, dims, str, box1, err)) return false;
  }
  if (L->use_halo) {
    cuuint32_t hb[5] = {64, 10, (cuuint32_t)(16 * L->x3.msub + 2), 1, 1};
    if (L->x3.s2) { hb[2] = 2; hb[3] = (cuuint32_t)(16 * L->x3.msub + 1); }
    if (!encode_bf16_map(&L->tmHalo3, d.in, 5, dims, str, hb, err)) return false;
  }
  if (L->use_pair) {
    cuuint32_t hb[5] = {64, 10, 18, 1, 1};
    if (!encode_bf16_map(&L->tmHalo2, d.in, 5, dims, str, hb, err)) return false;
  }
  L->halo_ok = false;
  if (d.k == 3 && d.stride == 1 && d.cout <= 128) {
    cuuint32_t hb[5] = {64, 10, 18, 1, 1};
    L->halo_ok = encode_bf16_map(&L->tmHalo, d.in, 5, dims, str, hb, err);
  }
  // TMA-store epilogue for 1x1 convs without residual: warp tiles of 32 pixels x 32 channels.  The output strides
  // come from the planned geometry (bind-time descriptors carry pointers only).
  p.tma_out = 0;
  if (d.k == 1 && d.res == nullptr && (d.out_mode == OUT_BF16 || d.out_mode == OUT_F32) && !getenv("YPB_NO_TMA_STORE")) {
    const bool f32 = d.out_mode == OUT_F32;
    const int elt = f32 ? 4 : 2;
    const long long hw = (long long)L->oH * L->oW;
    const long long pix = p.out_pix_stride, img = p.out_img_stride;
    const bool contiguous = img == hw * pix;
    const uint8_t* obase = reinterpret_cast<const uint8_t*>(d.out) + (long long)p.out_c_off * elt;
    const bool aligned = pix > 0 && img > 0 && (reinterpret_cast<uintptr_t>(obase) & 15) == 0 && (pix * elt) % 16 == 0 && (img * elt) % 16 == 0;
    if (aligned && (contiguous || hw % 32 == 0)) {
      cuuint64_t od[3] = {(cuuint64_t)d.cout, (cuuint64_t)(contiguous ? hw * d.B : hw), (cuuint64_t)d.B};
      cuuint64_t os[2] = {(cuuint64_t)(pix * elt), (cuuint64_t)(img * elt)};
      cuuint32_t ob[3] = {32, 32, 1};
      const int rank = contiguous ? 2 : 3;
      if (!encode_map(&L->tmO, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                      f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, obase, rank, od, os, ob, err))
        return false;
      p.tma_out = rank;
    }
  }
  cuuint64_t wd[3] = {(cuuint64_t)d.cin, (cuuint64_t)d.cout, (cuuint64_t)(d.k * d.k)};
  cuuint64_t ws[2] = {(cuuint64_t)d.cin * 2, (cuuint64_t)d.cin * 2 * d.cout};
  cuuint32_t wb[3] = {64, (cuuint32_t)p.n_tile, 1};
  if (!encode_bf16_map(&L->tmB, d.wg, 3, wd, ws, wb, err)) return false;
  if (L->use_halo) {
    cuuint32_t wb3[3] = {64, (cuuint32_t)p.n_tile, (cuuint32_t)L->x3.b_group};
    if (!encode_bf16_map(&L->tmB3, d.wg, 3, wd, ws, wb3, err)) return false;
  }
  if (L->use_pair) {
    cuuint32_t wb2[3] = {64, (cuuint32_t)(p.n_tile / 2), 3};
    if (!encode_bf16_map(&L->tmB2, d.wg, 3, wd, ws, wb2, err)) return false;
  }
  if (L->use_pair_tc2) {
    cuuint32_t wb2[3] = {64, (cuuint32_t)(p.n_tile / 2), 1};
    if (!encode_bf16_map(&L->tmB2p, d.wg, 3, wd, ws, wb2, err)) return false;
  }
  return true;
}

// One-line description of the launch the planner chose (diagnostics).
static void conv_describe(const ConvLaunch& L, int impl, char* out, int n) {
  if (L.use_pair && impl == 0)
    snprintf(out, n, "pair(cta_group::2) a_slots=%d b_slots=%d n_tile=%d splits=%d pairs=%d smem=%d", L.x2.a_slots, L.x2.b_slots,
             L.p.n_tile, L.n_splits, L.total_pairs, L.smem_pair);
  else if (L.use_halo && impl == 0)
    snprintf(out, n, "halo%s msub=%d a_slots=%d b_stat=%d b_slots=%d b_group=%d n_tile=%d splits=%d tiles=%d smem=%d", L.x3.s2 ? "-s2" : "", L.x3.msub,
             L.x3.a_slots, L.x3.b_stat, L.x3.b_slots, L.x3.b_group, L.p.n_tile, L.n_splits, L.total_tiles3, L.smem3);
  else if (L.use_pair_tc2 && impl == 0)
    snprintf(out, n, "tc2-pair(cta_group::2) msub=%d tile=%dx%d stages=%d n_tile=%d splits=%d pairs=%d smem=%d", L.p.msub, L.p.TH,
             L.p.TW, L.stages2p, L.p.n_tile, L.n_splits, L.total_pairs_tc2, L.smem2p);
  else
    snprintf(out, n, "tc2 msub=%d tile=%dx%d stages=%d n_tile=%d splits=%d tiles=%d smem=%d", L.p.msub, L.p.TH, L.p.TW,
             L.stages2, L.p.n_tile, L.n_splits, L.total_tiles, L.smem2);
}

// Per-device launch state.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count are properties of ONE
// device: a process that drives several GPUs (YOLO.to('cuda:1'), one host thread per GPU) needs them per ordinal.
struct DeviceState {
  bool conv_attrs = false, halo_test_attr = false, mask_tile_attr = false;
  size_t sppf_smem = 0, mask_smem = 0;
  int num_sms = 148;
};
static std::mutex g_dev_mutex;
static DeviceState g_dev_state[64];
static DeviceState& device_state() {  // of the CURRENT device
  int dev = 0;
  cudaGetDevice(&dev);
  return g_dev_state[dev & 63];
}
// Makes `device` current for the lifetime of the guard and restores the caller's device afterwards: a C-ABI entry
// point must not change the calling thread's current device behind the host application's back.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int device) {
    if (device < 0 || cudaGetDevice(&prev) != cudaSuccess) return;
    if (prev != device) switched = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

// One-time function attributes of the current device (must not happen inside a stream capture).
static cudaError_t conv_launch_init() {
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  DeviceState& ds = device_state();
  if (ds.conv_attrs) return cudaSuccess;
  cudaError_t e = cudaSuccess;
#define YPB_SET_SMEM(MODE)                                                                                        \
  e = cudaFuncSetAttribute(conv_tc2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);       \
  if (e != cudaSuccess) return e;                                                                                 \
  e = cudaFuncSetAttribute(conv3_halo_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);     \
  if (e != cudaSuccess) return e;                                                                                 \
  e = cudaFuncSetAttribute(conv3_halo2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);    \
  if (e != cudaSuccess) return e;                                                                                 \
  e = cudaFuncSetAttribute(conv_tc2p_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);      \
  if (e != cudaSuccess) return e;
  YPB_SET_SMEM(0) YPB_SET_SMEM(1) YPB_SET_SMEM(2) YPB_SET_SMEM(3) YPB_SET_SMEM(4) YPB_SET_SMEM(5) YPB_SET_SMEM(6) YPB_SET_SMEM(7)
#undef YPB_SET_SMEM
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&ds.num_sms, cudaDevAttrMultiProcessorCount, dev);
  ds.conv_attrs = true;
  return cudaSuccess;
}

// Launch of a persistent conv kernel with programmatic stream serialization (PDL): its prologue (barrier init, TMEM
// allocation, bias copy) overlaps the tail of the previous kernel in the stream / captured graph.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, int smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool no_pdl = getenv("YPB_NO_PDL") != nullptr;
  cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Same for a kernel that runs as CTA pairs (cluster of two: the two SMs of a TPC).
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl_pair(void (*kernel)(KArgs...), int grid, int block, int smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  static const bool no_pdl = getenv("YPB_NO_PDL") != nullptr;
  cfg.attrs = attr; cfg.numAttrs = no_pdl ? 1 : 2;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

static cudaError_t conv_launch(const ConvLaunch& L, cudaStream_t stream, int impl) {
#if YPB_DIAG
  if (impl == 1) {
    const long long total = (long long)L.sg.nB * L.sg.oH * L.sg.oW * (L.p.Cout / 16);
    conv_simt_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(L.sg, L.p);
    return cudaGetLastError();
  }
#else
  if (impl != 0) return cudaErrorNotSupported;  // the debugging twins live in libypb200_diag.so
#endif
  {
    cudaError_t e = conv_launch_init();
    if (e != cudaSuccess) return e;
  }
  const int num_sms = device_state().num_sms;
#if YPB_DIAG
  if (impl == 3) {  // halo-reuse experiment (3x3 stride 1, Cout <= 128): 16x8 tiles
    if (!L.halo_ok) return cudaErrorInvalidValue;
    {
      std::lock_guard<std::mutex> lock(g_dev_mutex);
      DeviceState& ds = device_state();
      if (!ds.halo_test_attr) {
        cudaFuncSetAttribute(conv_halo_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        ds.halo_test_attr = true;
      }
    }
    ConvParams p3 = L.p1;
    p3.tiles_h = (p3.tH + 15) / 16; p3.tiles_w = (p3.tW + 7) / 8;
    const int grid3 = p3.tB * p3.tiles_h * p3.tiles_w;
    const int smem3 = 1024 + kHaloBytes + 9 * p3.n_tile * 128 + 256;
    conv_halo_test_kernel<<<grid3, kConvThreads, smem3, stream>>>(L.tmHalo, L.tmB, p3);
    return cudaGetLastError();
  }
  if (impl == 2) {  // first-generation kernel: one tile per CTA (kept for A/B measurements)
    {
      std::lock_guard<std::mutex> lock(g_dev_mutex);
      DeviceState& ds = device_state();
      if (!ds.halo_test_attr) {
        cudaFuncSetAttribute(conv_halo_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        ds.halo_test_attr = true;
      }
    }
    conv_tc_kernel<<<L.grid1, kConvThreads, L.smem, stream>>>(L.tmA1, L.tmB, L.p1);
    return cudaGetLastError();
  }
#endif
  if (L.use_pair && impl == 0) {
    ConvParams p2 = L.p;
    p2.tiles_h = L.tiles_h2; p2.tiles_w = L.tiles_w2;
    conv_set_fastdiv(p2, L.n_splits);
    const int clusters = L.total_pairs < num_sms / 2 ? L.total_pairs : num_sms / 2;
    cudaError_t le = cudaSuccess;
#define YPB_PAIR_CASE(MODE) \
  case MODE: le = launch_pdl_pair(conv3_halo2_kernel<MODE>, 2 * clusters, kConv2Threads, L.smem_pair, stream, L.tmHalo2, L.tmB2, p2, L.x2, L.n_splits, L.total_pairs); break;
    switch (epi_mode_of(p2.out_mode, p2.res != nullptr, p2.act)) {
      YPB_PAIR_CASE(0) YPB_PAIR_CASE(1) YPB_PAIR_CASE(2) YPB_PAIR_CASE(3) YPB_PAIR_CASE(4) YPB_PAIR_CASE(5) YPB_PAIR_CASE(6)
      YPB_PAIR_CASE(7)
    }
#undef YPB_PAIR_CASE
    return le != cudaSuccess ? le : cudaGetLastError();
  }
  if (L.use_halo && impl == 0) {
    ConvParams p3 = L.p;
    p3.tiles_h = L.tiles_h3; p3.tiles_w = L.tiles_w3;
    conv_set_fastdiv(p3, L.n_splits);
    const int grid3 = L.total_tiles3 < num_sms ? L.total_tiles3 : num_sms;
    cudaError_t le = cudaSuccess;
#define YPB_HALO_CASE(MODE) \
  case MODE: le = launch_pdl(conv3_halo_kernel<MODE>, grid3, kConv2Threads, L.smem3, stream, L.tmHalo3, L.tmB3, p3, L.x3, L.n_splits, L.total_tiles3); break;
    switch (epi_mode_of(p3.out_mode, p3.res != nullptr, p3.act)) {
      YPB_HALO_CASE(0) YPB_HALO_CASE(1) YPB_HALO_CASE(2) YPB_HALO_CASE(3) YPB_HALO_CASE(4) YPB_HALO_CASE(5) YPB_HALO_CASE(6)
      YPB_HALO_CASE(7)
    }
#undef YPB_HALO_CASE
    return le != cudaSuccess ? le : cudaGetLastError();
  }
  if (L.use_pair_tc2 && impl == 0) {
    ConvParams pp = L.p;
    pp.stages = L.stages2p;
    conv_set_fastdiv(pp, L.n_splits);
    const int clusters = L.total_pairs_tc2 < num_sms / 2 ? L.total_pairs_tc2 : num_sms / 2;
    cudaError_t le = cudaSuccess;
#define YPB_TC2P_CASE(MODE) \
  case MODE: le = launch_pdl_pair(conv_tc2p_kernel<MODE>, 2 * clusters, kConv2Threads, L.smem2p, stream, L.tmA, L.tmB2p, pp.tma_out ? L.tmO : L.tmB2p, pp, L.n_splits, L.total_pairs_tc2); break;
    switch (epi_mode_of(pp.out_mode, pp.res != nullptr, pp.act)) {
      YPB_TC2P_CASE(0) YPB_TC2P_CASE(1) YPB_TC2P_CASE(2) YPB_TC2P_CASE(3) YPB_TC2P_CASE(4) YPB_TC2P_CASE(5) YPB_TC2P_CASE(6) YPB_TC2P_CASE(7)
    }
#undef YPB_TC2P_CASE
    return le != cudaSuccess ? le : cudaGetLastError();
  }
  ConvParams p2 = L.p;
  p2.stages = L.stages2;
  conv_set_fastdiv(p2, L.n_splits);
  const int grid = L.total_tiles < num_sms ? L.total_tiles : num_sms;
  cudaError_t le = cudaSuccess;
#define YPB_TC2_CASE(MODE) \
  case MODE: le = launch_pdl(conv_tc2_kernel<MODE>, grid, kConv2Threads, L.smem2, stream, L.tmA, L.tmB, p2.tma_out ? L.tmO : L.tmB, p2, L.n_splits, L.total_tiles); break;
  switch (epi_mode_of(p2.out_mode, p2.res != nullptr, p2.act)) {
    YPB_TC2_CASE(0) YPB_TC2_CASE(1) YPB_TC2_CASE(2) YPB_TC2_CASE(3) YPB_TC2_CASE(4) YPB_TC2_CASE(5) YPB_TC2_CASE(6) YPB_TC2_CASE(7)
  }
#undef YPB_TC2_CASE
  return le != cudaSuccess ? le : cudaGetLastError();
}

}  // namespace ypb
