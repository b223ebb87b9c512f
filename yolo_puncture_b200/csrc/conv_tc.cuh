// Fused Conv2d(+folded BN bias)(+SiLU)(+residual) as an implicit GEMM on the 5th-gen tensor cores.
//
// Replaces the cuDNN calls behind UPSTREAM ultralytics `Conv.forward_fuse` / `nn.ConvTranspose2d`
// (SURVEY.md §2.2, §8 a3-a5), i.e. everything `AutoBackend.forward` runs per frame for the
// `model.predict(...)` calls at reference yolo_seg/app.py:91 and yolo_seg/yolo_with_deva.py:51.
//
//   out[m, n] = act( sum_{tap, c} A[m @ tap, c] * Wg[tap][n][c] + bias[n] ) (+ res[m, n])
//
//   m  = one output pixel of an NHWC bf16 activation; a CTA owns a TH x TW rectangle (<=128 pixels)
//        of one image (3x3 convs) or 128 consecutive pixels of the flattened batch (1x1 convs);
//   A  = input activation, fetched per (tap, 64-channel chunk) by ONE 5-D TMA box load
//        {64 ch, TW, (1), TH, (1)} whose out-of-bounds rows are zero-filled by the TMA unit, which
//        is exactly the conv's zero padding; stride-2 convs address the input through the view
//        (2C, W/2, 2, H/2, B) so that a tap is again a dense box;
//   Wg = [tap][Cout][Cin] bf16 (K-major), one 3-D TMA box {64, n_tile, 1} per (tap, chunk);
//   D  = fp32 accumulator in TMEM (128 lanes x n_tile columns), tcgen05.mma kind::f16, M=128;
//   epilogue = tcgen05.ld -> +bias -> SiLU -> (+residual) -> bf16/fp32 store into a channel slice
//        of the destination buffer (so Concat / C2f chunk never materialise), or pixel-shuffle
//        store for ConvTranspose2d(k=2,s=2).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue
// (warp w reads TMEM lanes 32*(w%4)..+31); warp 2 also owns the TMEM allocation.
#pragma once
#include "common.cuh"

namespace ypb {

enum ConvOutMode : int { OUT_BF16 = 0, OUT_F32 = 1, OUT_SHUFFLE2_BF16 = 2 };

struct ConvParams {
  // M tiling space (rect mode: B,H,W of the OUTPUT map; flat mode: 1,1,B*H*W)
  int tB, tH, tW;
  int TH, TW;            // CTA tile rectangle, TH*TW <= 128
  int tiles_h, tiles_w;  // tiles per image
  // GEMM
  int Cin;               // reduction channels per tap (multiple of 16)
  int Cout;              // N (multiple of 16)
  int n_tile;            // N per CTA (multiple of 16, <= 256); gridDim.y = Cout / n_tile
  int ntaps;             // 1 or 9
  int stages;
  // A-operand TMA coordinates: c[d] = a_base[d] + b*a_cb[d] + h0*a_ch[d] + w0*a_cw[d] + tap[t][d]; c[0] += 64*chunk
  int a_base[5], a_cb[5], a_ch[5], a_cw[5];
  int tap[9][5];
  // epilogue
  int out_mode, act;
  int img_HW, img_W;     // real output pixels per image / width (q -> image, row, col)
  void* out;
  long long out_img_stride;
  int out_pix_stride, out_c_off;
  const float* bias;     // [Cout] fp32
  const __nv_bfloat16* res;
  long long res_img_stride;
  int res_pix_stride, res_c_off;
  int msub;      // conv_tc2: 128-row accumulator sub-tiles per CTA tile (share every weight tile); 1 or 2
  int sub_rows;  // rows of one sub-tile (<= 128, multiple of 8); sub-tile s starts at row s*sub_rows of the A stage
  int dbg;  // YPB_DBG experiments (0 in production): 1 = no bias/SiLU math, 2 = no output stores, 4 = no MMA issue
};

// Store 16 consecutive output channels [n, n+16) of output pixel q. v = raw accumulators.
__device__ __forceinline__ void conv_epilogue_store16(const ConvParams& p, int q, int n, const float (&acc)[16]) {
  float y[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float t = acc[j] + __ldg(p.bias + n + j);
    y[j] = p.act ? silu_f(t) : t;
  }
  const int b = q / p.img_HW;
  const int rem = q - b * p.img_HW;
  if (p.out_mode == OUT_F32) {
    float* o = reinterpret_cast<float*>(p.out) + b * p.out_img_stride + (long long)rem * p.out_pix_stride + p.out_c_off + n;
#pragma unroll
    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
    return;
  }
  long long off;
  if (p.out_mode == OUT_SHUFFLE2_BF16) {
    const int cq = p.Cout >> 2;
    const int g = n / cq, c = n - g * cq;
    const int h = rem / p.img_W, w = rem - h * p.img_W;
    off = b * p.out_img_stride + ((long long)(2 * h + (g >> 1)) * (2 * p.img_W) + 2 * w + (g & 1)) * p.out_pix_stride +
          p.out_c_off + c;
  } else {
    off = b * p.out_img_stride + (long long)rem * p.out_pix_stride + p.out_c_off + n;
  }
  if (p.res != nullptr) {
    // y = bf16(act(...)) first, then bf16(y + res): the same two roundings as storing the conv
    // output and adding the shortcut afterwards (Bottleneck: x + cv2(cv1(x))).
    const __nv_bfloat16* r = p.res + b * p.res_img_stride + (long long)rem * p.res_pix_stride + p.res_c_off + n;
    uint4 r0 = *reinterpret_cast<const uint4*>(r), r1 = *reinterpret_cast<const uint4*>(r + 8);
    const __nv_bfloat16* rb0 = reinterpret_cast<const __nv_bfloat16*>(&r0);
    const __nv_bfloat16* rb1 = reinterpret_cast<const __nv_bfloat16*>(&r1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      y[j] = bf16_round(y[j]) + __bfloat162float(rb0[j]);
      y[j + 8] = bf16_round(y[j + 8]) + __bfloat162float(rb1[j]);
    }
  }
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
  uint4 s0 = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
  uint4 s1 = make_uint4(pack_bf16x2(y[8], y[9]), pack_bf16x2(y[10], y[11]), pack_bf16x2(y[12], y[13]),
                        pack_bf16x2(y[14], y[15]));
  *reinterpret_cast<uint4*>(o) = s0;
  *reinterpret_cast<uint4*>(o + 8) = s1;
}

constexpr int kConvThreads = 192;
constexpr int kATileBytes = 128 * 128;  // 128 rows x 64 bf16

__host__ __device__ inline int conv_stage_bytes(int n_tile) { return kATileBytes + n_tile * 128; }
__host__ __device__ inline int conv_smem_bytes(int n_tile, int stages) {
  return 1024 /*align slack*/ + stages * conv_stage_bytes(n_tile) + 256 /*barriers*/;
}

__global__ void __launch_bounds__(kConvThreads)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = conv_stage_bytes(p.n_tile);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* accum_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile -> (b, h0, w0)
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int b = blockIdx.x / tiles_per_img;
  const int t_in = blockIdx.x - b * tiles_per_img;
  const int th = t_in / p.tiles_w;
  const int h0 = th * p.TH, w0 = (t_in - th * p.tiles_w) * p.TW;
  const int n0 = blockIdx.y * p.n_tile;

  const int kchunks = (p.Cin + 63) >> 6;
  const int k_iters = p.ntaps * kchunks;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)p.n_tile) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    mbar_init(accum_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int cbase[5];
#pragma unroll
      for (int d = 0; d < 5; ++d) cbase[d] = p.a_base[d] + b * p.a_cb[d] + h0 * p.a_ch[d] + w0 * p.a_cw[d];
      const uint32_t tx_bytes = (uint32_t)(p.TH * p.TW * 128 + p.n_tile * 128);
      int it = 0;
      for (int t = 0; t < p.ntaps; ++t) {
        for (int c = 0; c < kchunks; ++c, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(empty_bar + s, ph ^ 1, 1u);
          uint8_t* sa = smem + s * stage_bytes;
          uint8_t* sb = sa + kATileBytes;
          mbar_expect_tx(full_bar + s, tx_bytes);
          tma_load_5d(sa, &tmA, full_bar + s, cbase[0] + p.tap[t][0] + c * 64, cbase[1] + p.tap[t][1],
                      cbase[2] + p.tap[t][2], cbase[3] + p.tap[t][3], cbase[4] + p.tap[t][4]);
          tma_load_3d(sb, &tmB, full_bar + s, c * 64, n0, t);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.n_tile);
      int it = 0;
      for (int t = 0; t < p.ntaps; ++t) {
        for (int c = 0; c < kchunks; ++c, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(full_bar + s, ph, 2u);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * stage_bytes);
          const uint32_t sb = sa + kATileBytes;
          int ksteps = (p.Cin - c * 64) >> 4;
          if (ksteps > 4) ksteps = 4;
          for (int j = 0; j < ksteps; ++j) {
            umma_bf16(tmem_base, umma_desc_sw128(sa + j * 32), umma_desc_sw128(sb + j * 32), idesc,
                      (it > 0 || j > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar + s);  // frees the smem slot when these MMAs retire
        }
      }
      umma_commit(accum_bar);  // accumulator complete
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int lg = warp & 3;  // TMEM lane group this warp may access
    const int r = lg * 32 + lane;
    const int rh = r / p.TW, rw = r - rh * p.TW;
    const int h = h0 + rh, w = w0 + rw;
    const bool valid = (r < p.TH * p.TW) && (h < p.tH) && (w < p.tW);
    const int q = (b * p.tH + h) * p.tW + w;
    mbar_wait(accum_bar, 0, 4u);
    tc_fence_after();
    for (int j = 0; j < p.n_tile; j += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)j, v);
      tmem_ld_wait();
      if (valid) {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(v[i]);
        conv_epilogue_store16(p, q, n0 + j, acc);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent, warp-specialised version (the product path).  One CTA per SM loops over output tiles:
//   * the TMA producer runs ahead across tile boundaries through one shared ring of smem stages, so the
//     next tile's operands are already in flight while the current tile is multiplied and stored;
//   * the accumulator is double-buffered in TMEM (2 x n_tile columns): the MMA warp starts tile i+1 while
//     the epilogue warps drain tile i (tcgen05.ld -> bias/SiLU/residual -> global stores);
//   * tile order is M-major with the Cout splits innermost, so the CTAs that run concurrently work on
//     neighbouring rectangles of the same image and share halos / weights in L2.
// Roles: warp 0 TMA producer, warp 1 MMA issuer (also owns TMEM), warps 2.. = kEpiWarps epilogue warps
// (warp w reads TMEM lanes 32*(w%4)..+31; with 8 epilogue warps each lane group's columns are split in two).
// ------------------------------------------------------------------------------------------------
constexpr int kEpiWarps = 8;
constexpr int kProdWarps = 4;   // TMA producer warps: one thread sustains only ~1 box load per 0.35 us (tools/tma_bench.py)
constexpr int kMmaWarp = kProdWarps;
constexpr int kEpiWarp0 = kProdWarps + 1;
constexpr int kConv2Threads = 32 * (kProdWarps + 1 + kEpiWarps);
constexpr int kEpiStageBytes = 32 * (128 + 16);  // per-warp staging tile: 32 rows x (<=128 B + 16 B pad)

__host__ __device__ inline int conv2_acc_stride(int n_tile) { return (n_tile + 31) & ~31; }
__host__ __device__ inline int conv2_smem_bytes(int n_tile, int stages) {
  return 1024 /*align slack*/ + stages * conv_stage_bytes(n_tile) + 256 /*barriers*/ + kEpiWarps * kEpiStageBytes;
}

// ------------------------------------------------------------------------------------------------
// Epilogue of one 128-row accumulator (sub)tile for one warp (shared by the persistent kernels).
// The warp owns TMEM lanes [32*lg, +32) (t_addr points at them) and the 16-column chunks [c_begin, c_end).
// Phase 1: TMEM -> registers -> +bias -> SiLU -> packed bf16/fp32 into the warp's private, padded smem staging
// tile (lane = row: conflict-free thanks to the +16 B row pitch).  `release` (may be null) is arrived on as soon
// as the last tcgen05.ld of this call has landed, handing the accumulator back to the MMA warp.
// Phase 2: the staged rows are written back with lanes running along the channel dimension, so every store
// instruction covers whole 128-byte lines of NHWC rows (a thread-per-row store touches 32 lines per instruction).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void epi_drain(const ConvParams& p, uint8_t* stage, int lane, int c_begin, int c_end,
                                          uint32_t t_addr, int n0, bool valid, int q, uint64_t* release) {
  const int elt = p.out_mode == OUT_F32 ? 4 : 2;
  const int chunks_per_pass = elt == 2 ? 4 : 2;  // <= 128 B of output row per pass
  const int qb = q / p.img_HW, rem = q - qb * p.img_HW;
  // element offset of channel 0 of this row in the output (pixel-shuffle: of sub-pixel (0,0))
  long long off_row;
  if (p.out_mode == OUT_SHUFFLE2_BF16) {
    const int ph = rem / p.img_W, pw = rem - ph * p.img_W;
    off_row = qb * p.out_img_stride + ((long long)(2 * ph) * (2 * p.img_W) + 2 * pw) * p.out_pix_stride + p.out_c_off;
  } else {
    off_row = qb * p.out_img_stride + (long long)rem * p.out_pix_stride + p.out_c_off;
  }
  const long long res_row = qb * p.res_img_stride + (long long)rem * p.res_pix_stride + p.res_c_off;
  if (c_begin >= c_end) {  // n_tile == 16: the second warp of the lane group has no columns, it only hands back
    __syncwarp();
    if (lane == 0 && release != nullptr) mbar_arrive(release);
  }
  for (int cp0 = c_begin; cp0 < c_end; cp0 += chunks_per_pass) {
    const int nch = min(chunks_per_pass, c_end - cp0);
    const int row_bytes = nch * 16 * elt, pitch = row_bytes + 16;
    uint8_t* my = stage + lane * pitch;
    const int ppr = row_bytes >> 4;               // 16-byte pieces per row (<= 8)
    const int ppr_inv = (65536 + ppr - 1) / ppr;  // piece / ppr == (piece * ppr_inv) >> 16 for piece < 512
    // Residual (Bottleneck shortcut): issue this pass's coalesced 16-byte loads NOW so that their latency is
    // covered by the TMEM reads and the SiLU math of phase 1.
    uint4 rres[8];
    if (p.res != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        rres[i] = make_uint4(0, 0, 0, 0);
        if (i < ppr) {
          const int piece = lane + 32 * i;
          const int row = (piece * ppr_inv) >> 16, pc = piece - row * ppr;
          const long long r_r = __shfl_sync(0xffffffffu, res_row, row);
          const int ok = __shfl_sync(0xffffffffu, (int)valid, row);
          if (ok) rres[i] = __ldg(reinterpret_cast<const uint4*>(p.res + r_r + n0 + cp0 * 16 + pc * 8));
        }
      }
    }
    for (int ch = 0; ch < nch; ++ch) {
      uint32_t v[16];
      tmem_ld16(t_addr + (uint32_t)((cp0 + ch) * 16), v);
      tmem_ld_wait();
      const int n = n0 + (cp0 + ch) * 16;
      float y[16];
      const float4* bp = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 bv = __ldg(bp + i);
        y[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + bv.x;
        y[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + bv.y;
        y[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + bv.z;
        y[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + bv.w;
      }
      if (p.act && !(p.dbg & 1)) {
        if (elt == 2) {
#pragma unroll
          for (int i = 0; i < 16; ++i) y[i] = silu_f(y[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) y[i] = silu_precise_f(y[i]);
        }
      }
      if (elt == 2) {
        uint4* d = reinterpret_cast<uint4*>(my + ch * 32);
        d[0] = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
        d[1] = make_uint4(pack_bf16x2(y[8], y[9]), pack_bf16x2(y[10], y[11]), pack_bf16x2(y[12], y[13]),
                          pack_bf16x2(y[14], y[15]));
      } else {
        float4* d = reinterpret_cast<float4*>(my + ch * 64);
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
      }
    }
    if (cp0 + chunks_per_pass >= c_end) {  // last TMEM read of this tile: hand the accumulator back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0 && release != nullptr) mbar_arrive(release);
    } else {
      __syncwarp();
    }
    if (!(p.dbg & 2)) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i >= ppr) break;
        const int piece = lane + 32 * i;
        const int row = (piece * ppr_inv) >> 16, pc = piece - row * ppr;
        const long long o_r = __shfl_sync(0xffffffffu, off_row, row);
        const int ok = __shfl_sync(0xffffffffu, (int)valid, row);
        uint4 val = *reinterpret_cast<const uint4*>(stage + row * pitch + pc * 16);
        if (!ok) continue;
        if (elt == 4) {
          const int n = n0 + cp0 * 16 + pc * 4;
          *reinterpret_cast<uint4*>(reinterpret_cast<float*>(p.out) + o_r + n) = val;
          continue;
        }
        const int n = n0 + cp0 * 16 + pc * 8;
        if (p.res != nullptr) {
          const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&val);
          const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&rres[i]);
          uint32_t o[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float2 fa = __bfloat1622float2(a2[u]), fb = __bfloat1622float2(b2[u]);
            o[u] = pack_bf16x2(fa.x + fb.x, fa.y + fb.y);
          }
          val = make_uint4(o[0], o[1], o[2], o[3]);
        }
        long long off = o_r + n;
        if (p.out_mode == OUT_SHUFFLE2_BF16) {
          const int cq = p.Cout >> 2;
          const int g = n / cq, c = n - g * cq;
          off = o_r + ((long long)(g >> 1) * (2 * p.img_W) + (g & 1)) * p.out_pix_stride + c;
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off) = val;
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kConv2Threads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ ConvParams p, int n_splits, int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = p.msub * kATileBytes;
  const int stage_bytes = a_bytes + p.n_tile * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tfull_bar = empty_bar + p.stages;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;         // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int kchunks = (p.Cin + 63) >> 6;
  const int acc_stride = conv2_acc_stride(p.n_tile);
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * p.msub * acc_stride)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar + i, 1);
      mbar_init(tempty_bar + i, kEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kProdWarps) {
    // ===================== TMA producers: warp w owns the ring stages s with s % kProdWarps == w =====================
    if (elect_one()) {
      const uint32_t tx_bytes = (uint32_t)(p.TH * p.TW * 128 + p.n_tile * 128);  // the A box spans all sub-tiles
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / n_splits, n0 = (tile - mt * n_splits) * p.n_tile;
        const int b = mt / tiles_per_img, t_in = mt - b * tiles_per_img;
        const int th = t_in / p.tiles_w;
        const int h0 = th * p.TH, w0 = (t_in - th * p.tiles_w) * p.TW;
        int cbase[5];
#pragma unroll
        for (int d = 0; d < 5; ++d) cbase[d] = p.a_base[d] + b * p.a_cb[d] + h0 * p.a_ch[d] + w0 * p.a_cw[d];
        for (int c = 0; c < kchunks; ++c) {
          for (int t = 0; t < p.ntaps; ++t, ++it) {
            // a stage always belongs to the same producer: parity waits are only sound one phase ahead
            const int s = it % p.stages;
            if ((s % kProdWarps) != warp) continue;
            const uint32_t ph = (it / p.stages) & 1;
            mbar_wait(empty_bar + s, ph ^ 1, 1u);
            uint8_t* sa = smem + s * stage_bytes;
            mbar_expect_tx(full_bar + s, tx_bytes);
            tma_load_5d(sa, &tmA, full_bar + s, cbase[0] + p.tap[t][0] + c * 64, cbase[1] + p.tap[t][1],
                        cbase[2] + p.tap[t][2], cbase[3] + p.tap[t][3], cbase[4] + p.tap[t][4]);
            tma_load_3d(sa + a_bytes, &tmB, full_bar + s, c * 64, n0, t);
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (accumulation order: chunk-major, tap-minor, like conv3_halo_kernel) =====================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, p.n_tile);
      int it = 0, acc = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acc) {
        const int buf = acc & 1;
        mbar_wait(tempty_bar + buf, ((acc >> 1) & 1) ^ 1, 8u);  // epilogue has drained this accumulator
        tc_fence_after();
        int first = 1;
        for (int c = 0; c < kchunks; ++c) {
          int ksteps = (p.Cin - c * 64) >> 4;
          if (ksteps > 4) ksteps = 4;
          for (int t = 0; t < p.ntaps; ++t, ++it) {
            const int s = it % p.stages;
            const uint32_t ph = (it / p.stages) & 1;
            mbar_wait(full_bar + s, ph, 2u);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + s * stage_bytes);
            const uint32_t sb = sa + a_bytes;
            for (int sidx = 0; sidx < p.msub; ++sidx) {
              const uint32_t d_tmem = tmem_base + (uint32_t)((buf * p.msub + sidx) * acc_stride);
              const uint32_t sa_s = sa + (uint32_t)(sidx * p.sub_rows * 128);
              for (int j = 0; j < ((p.dbg & 4) ? 0 : ksteps); ++j)
                umma_bf16(d_tmem, umma_desc_sw128(sa_s + j * 32), umma_desc_sw128(sb + j * 32), idesc,
                          (first && j == 0) ? 0u : 1u);
            }
            first = 0;
            umma_commit(empty_bar + s);
          }
        }
        umma_commit(tfull_bar + buf);
      }
    }
  } else {
    // ===================== epilogue =====================
    // Each warp owns 32 accumulator rows (its TMEM lane group) and a contiguous range of 16-column chunks.
    // Phase 1: TMEM -> registers -> +bias -> SiLU -> packed bf16/fp32 into a private, padded smem staging
    // tile (lane = row: conflict-free thanks to the +16 B row pitch).  The accumulator is released as soon
    // as the last tcgen05.ld has landed.  Phase 2: the warp writes the staged rows back with lanes running
    // along the channel dimension, so every store instruction covers whole 128-byte lines of NHWC rows
    // (a thread-per-row store would touch 32 different lines per instruction).
    const int lg = warp & 3;
    const int part = (warp - kEpiWarp0) >> 2;
    const int nchunks = p.n_tile >> 4;
    const int half = (nchunks + 1) >> 1;
    const int c_begin = part == 0 ? 0 : half, c_end = part == 0 ? half : nchunks;
    uint8_t* stage = smem + p.stages * stage_bytes + 256 + (warp - kEpiWarp0) * kEpiStageBytes;
    const int r = lg * 32 + lane;
    int acc = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acc) {
      const int mt = tile / n_splits, n0 = (tile - mt * n_splits) * p.n_tile;
      const int b = mt / tiles_per_img, t_in = mt - b * tiles_per_img;
      const int th = t_in / p.tiles_w;
      const int buf = acc & 1;
      mbar_wait(tfull_bar + buf, (acc >> 1) & 1, 4u);
      tc_fence_after();
      for (int sidx = 0; sidx < p.msub; ++sidx) {
        const int R = sidx * p.sub_rows + r;  // row of the whole CTA tile
        const int rh = R / p.TW, rw = R - rh * p.TW;
        const int h = th * p.TH + rh, w = (t_in - th * p.tiles_w) * p.TW + rw;
        const bool valid = (r < p.sub_rows) && (h < p.tH) && (w < p.tW);
        const int q = valid ? (b * p.tH + h) * p.tW + w : 0;
        const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((buf * p.msub + sidx) * acc_stride);
        epi_drain(p, stage, lane, c_begin, c_end, t_addr, n0, valid, q, sidx == p.msub - 1 ? tempty_bar + buf : nullptr);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

__device__ __forceinline__ uint64_t umma_desc_sw128_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// ------------------------------------------------------------------------------------------------
// 3x3 stride-1 convs: halo-reuse kernel (persistent, warp-specialised like conv_tc2_kernel).
//
// Measured on B200 (tools/tma_bench.py): the L2 -> SM operand path tops out at ~7 TB/s chip-wide (~47 GB/s per
// SM), i.e. L2 gives no bandwidth amplification over HBM, and conv_tc2_kernel's per-tap boxes re-fetch the input
// tile nine times.  Here a CTA owns 8-pixel-wide column tiles (16*msub rows x 8 cols): ONE TMA box
// {64 ch, 10, 16*msub+2} brings the tile plus its halo per 64-channel chunk, and tap (kh,kw) of sub-tile s is the
// same shared-memory buffer seen through a UMMA descriptor whose start address is advanced by
// ((16*s+kh)*10 + kw) rows and whose 8-row groups are 10 rows (1280 B) apart.  That is legal because both TMA and
// tcgen05.mma apply the 128-byte swizzle on absolute shared-memory address bits (verified on hardware,
// tools/halo_experiment.py).  Input re-reads drop from 9x to ~1.4x.
// Weights: if all nine taps of all chunks fit (b_stat), they are loaded ONCE per CTA and stay resident;
// otherwise they stream through their own ring of per-tap slots and msub=2 sub-tiles share every weight tile.
// ------------------------------------------------------------------------------------------------
struct Conv3Extra {
  int msub;        // 128-row sub-tiles per CTA tile (1 or 2), stacked vertically
  int a_slots;     // halo ring depth
  int a_bytes;     // bytes per halo slot (1024 multiple)
  int halo_rows;   // (16*msub+2)*10
  int b_slots;     // weight ring depth, in groups (streaming mode)
  int b_group;     // taps per weight box / ring slot: 3 (one kernel row) or 1
  int b_stat;      // 1: all weights resident in smem
  int b_bytes;     // weight region bytes
};

__global__ void __launch_bounds__(kConv2Threads, 1)
conv3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ ConvParams p, const Conv3Extra x, int n_splits, int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + x.a_slots * x.a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + x.b_bytes);
  uint64_t* a_full = bars;            // [4]
  uint64_t* a_empty = bars + 4;       // [4]
  uint64_t* b_full = bars + 8;        // [12]
  uint64_t* b_empty = bars + 20;      // [12]
  uint64_t* tfull_bar = bars + 32;    // [2]
  uint64_t* tempty_bar = bars + 34;   // [2]
  uint64_t* ball_bar = bars + 36;     // resident weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 37);
  uint8_t* stage_base = reinterpret_cast<uint8_t*>(bars) + 512;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int kchunks = (p.Cin + 63) >> 6;
  const int acc_stride = conv2_acc_stride(p.n_tile);
  const int tap_bytes = p.n_tile * 128;
  const int grp_bytes = x.b_group * tap_bytes;
  const int ngroups = 9 / x.b_group;  // weight boxes per 64-channel chunk
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * x.msub * acc_stride)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < 4; ++i) { mbar_init(a_full + i, 1); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < 12; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar + i, 1); mbar_init(tempty_bar + i, kEpiWarps); }
    mbar_init(ball_bar, 1);
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer 0: halo tiles (and the resident weights, once) =====================
    if (elect_one()) {
      if (x.b_stat) {
        mbar_expect_tx(ball_bar, (uint32_t)(9 * kchunks * tap_bytes));
        for (int c = 0; c < kchunks; ++c)
          for (int g = 0; g < ngroups; ++g)
            tma_load_3d(sB + (c * 9 + g * x.b_group) * tap_bytes, &tmB, ball_bar, c * 64, 0, g * x.b_group);
      }
      int ia = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / n_splits;
        const int b = mt / tiles_per_img, t_in = mt - b * tiles_per_img;
        const int th = t_in / p.tiles_w;
        const int h0 = th * 16 * x.msub, w0 = (t_in - th * p.tiles_w) * 8;
        for (int c = 0; c < kchunks; ++c, ++ia) {
          const int sa = ia % x.a_slots;
          mbar_wait(a_empty + sa, ((ia / x.a_slots) & 1) ^ 1, 1u);
          mbar_expect_tx(a_full + sa, (uint32_t)(x.halo_rows * 128));
          tma_load_5d(sA + sa * x.a_bytes, &tmA, a_full + sa, p.a_base[0] + c * 64, w0 - 1, h0 - 1, b, 0);
        }
      }
    }
  } else if (warp < kProdWarps) {
    // ===================== TMA producers 1..3: streamed weight boxes, ring slot sb owned by warp 1 + sb % 3 =====================
    if (!x.b_stat && elect_one()) {
      int ib = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / n_splits, n0 = (tile - mt * n_splits) * p.n_tile;
        for (int c = 0; c < kchunks; ++c) {
          for (int g = 0; g < ngroups; ++g, ++ib) {
            const int sb = ib % x.b_slots;
            if (1 + (sb % (kProdWarps - 1)) != warp) continue;  // a slot always belongs to the same producer
            mbar_wait(b_empty + sb, ((ib / x.b_slots) & 1) ^ 1, 1u);
            mbar_expect_tx(b_full + sb, (uint32_t)grp_bytes);
            tma_load_3d(sB + sb * grp_bytes, &tmB, b_full + sb, c * 64, n0, g * x.b_group);
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, p.n_tile);
      if (x.b_stat) {
        mbar_wait(ball_bar, 0, 2u);
        tc_fence_after();
      }
      int ia = 0, ib = 0, acc = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acc) {
        const int buf = acc & 1;
        mbar_wait(tempty_bar + buf, ((acc >> 1) & 1) ^ 1, 8u);
        tc_fence_after();
        int first = 1;
        for (int c = 0; c < kchunks; ++c, ++ia) {
          const int sa = ia % x.a_slots;
          mbar_wait(a_full + sa, (ia / x.a_slots) & 1, 2u);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + sa * x.a_bytes);
          int ksteps = (p.Cin - c * 64) >> 4;
          if (ksteps > 4) ksteps = 4;
          for (int g = 0; g < ngroups; ++g) {
            uint32_t g_base;
            int sb = 0;
            if (x.b_stat) {
              g_base = smem_u32(sB + (c * 9 + g * x.b_group) * tap_bytes);
            } else {
              sb = ib % x.b_slots;
              mbar_wait(b_full + sb, (ib / x.b_slots) & 1, 2u);
              tc_fence_after();
              g_base = smem_u32(sB + sb * grp_bytes);
              ++ib;
            }
            for (int u = 0; u < x.b_group; ++u) {
              const int t = g * x.b_group + u;
              const int kh = t / 3, kw = t - kh * 3;
              const uint32_t b_base = g_base + (uint32_t)(u * tap_bytes);
              for (int sidx = 0; sidx < x.msub; ++sidx) {
                const uint32_t a0 = a_base + (uint32_t)(((sidx * 16 + kh) * 10 + kw) * 128);
                const uint32_t d_tmem = tmem_base + (uint32_t)((buf * x.msub + sidx) * acc_stride);
                for (int j = 0; j < ((p.dbg & 4) ? 0 : ksteps); ++j)
                  umma_bf16(d_tmem, umma_desc_sw128_sbo(a0 + j * 32, 1280), umma_desc_sw128(b_base + j * 32), idesc,
                            (first && j == 0) ? 0u : 1u);
              }
              first = 0;
            }
            if (!x.b_stat) umma_commit(b_empty + sb);
          }
          umma_commit(a_empty + sa);
        }
        umma_commit(tfull_bar + buf);
      }
    }
  } else {
    // ===================== epilogue =====================
    const int lg = warp & 3;
    const int part = (warp - kEpiWarp0) >> 2;
    const int nchunks = p.n_tile >> 4;
    const int half = (nchunks + 1) >> 1;
    const int c_begin = part == 0 ? 0 : half, c_end = part == 0 ? half : nchunks;
    uint8_t* stage = stage_base + (warp - kEpiWarp0) * kEpiStageBytes;
    const int r = lg * 32 + lane;
    int acc = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acc) {
      const int mt = tile / n_splits, n0 = (tile - mt * n_splits) * p.n_tile;
      const int b = mt / tiles_per_img, t_in = mt - b * tiles_per_img;
      const int th = t_in / p.tiles_w;
      const int h0 = th * 16 * x.msub, w0 = (t_in - th * p.tiles_w) * 8;
      const int buf = acc & 1;
      mbar_wait(tfull_bar + buf, (acc >> 1) & 1, 4u);
      tc_fence_after();
      for (int sidx = 0; sidx < x.msub; ++sidx) {
        const int h = h0 + sidx * 16 + (r >> 3), w = w0 + (r & 7);
        const bool valid = (h < p.tH) && (w < p.tW);
        const int q = valid ? (b * p.tH + h) * p.tW + w : 0;
        const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((buf * x.msub + sidx) * acc_stride);
        epi_drain(p, stage, lane, c_begin, c_end, t_addr, n0, valid, q, sidx == x.msub - 1 ? tempty_bar + buf : nullptr);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// Stem on the tensor pipe: uint8 BGR letterboxed frame -> Conv3x3 s2 p1 (+folded BN, /255 folded into the weights)
// -> SiLU -> bf16 NHWC.  UPSTREAM sites replaced: engine/predictor.py::preprocess (BGR->RGB, /255) + model.0.
// K = 27 is too thin for a TMA-fed pipeline, so the CTA builds the im2col tile itself: 128 output pixels (8 x 16)
// per tile; each thread converts its pixel's 27 bytes to bf16 (0..255 are exact in bf16) and writes one K-major
// SWIZZLE_128B row.  The fp32 weights are split w/255 = hi + lo into two bf16 terms, so K = 64 = [x | x] against
// [hi | lo]: four tcgen05.mma (K = 16, N = C0) per tile give fp32-grade products at bf16 tensor speed.
// A CTA loops over `tiles_per_cta` tiles; several CTAs are resident per SM and hide each other's phases.
// Weights: wq[C0][64] bf16 (K-major, k = (kh*3+kw)*3 + c_rgb in each 32-wide half), bias fp32.
// The frame is read as aligned 32-bit words straight from HBM: 3 B per pixel instead of a 12 B/pixel fp32 tensor.
// ------------------------------------------------------------------------------------------------
constexpr int kStemTH = 8, kStemTW = 16;
constexpr int kStemRowWords = 27;  // 33 pixels x 3 B = 99 B plus up to 3 B of misalignment -> 26 words, +1 spare

__global__ void __launch_bounds__(128, 6)
stem_tc_kernel(const uint8_t* __restrict__ frames, int H, int W, int nB, const __nv_bfloat16* __restrict__ wq,
               const __grid_constant__ ConvParams p, int tiles_per_img, int tiles_w, int total_tiles, int tiles_per_cta,
               int alias_stage) {
  // dynamic smem: [A tile 16 KB][weights C0*128 B, 1 KB multiple][patch 17x27 words][epilogue staging unless it aliases A]
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sA = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = sA + 128 * 128;
  uint32_t* sIn = reinterpret_cast<uint32_t*>(sB + ((p.Cout * 128 + 1023) & ~1023));
  uint8_t* sStage = alias_stage ? sA : reinterpret_cast<uint8_t*>(sIn + 17 * kStemRowWords + 4);  // A is dead once the MMAs retire
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int C0 = p.Cout;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)C0) tmem_cols <<= 1;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
  for (int i = tid; i < C0 * 8; i += 128) {  // weights -> swizzled K-major rows of 128 B
    const int n = i >> 3, j = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(wq + n * 64 + j * 8);
    *reinterpret_cast<uint4*>(sB + n * 128 + ((j ^ (n & 7)) << 4)) = v;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t idesc = umma_idesc_bf16(128, C0);
  const int oH = H >> 1, oW = W >> 1;
  const long long frame_bytes = (long long)H * W * 3, all_words = (frame_bytes * nB + 3) >> 2;
  const uint32_t* words = reinterpret_cast<const uint32_t*>(frames);  // cudaMalloc'd: at least 256-byte aligned
  uint32_t phase = 0;
  const int t_begin = blockIdx.x * tiles_per_cta;
  for (int tile = t_begin; tile < min(t_begin + tiles_per_cta, total_tiles); ++tile) {
    const int b = tile / tiles_per_img, t_in = tile - b * tiles_per_img;
    const int th = t_in / tiles_w;
    const int oh0 = th * kStemTH, ow0 = (t_in - th * tiles_w) * kStemTW;
    const int ih0 = 2 * oh0 - 1, iw0 = 2 * ow0 - 1;
    // input patch: 17 rows; row r holds the aligned words covering bytes [g0, g0+99) of the frame row, g0 = byte
    // offset of pixel (ih0+r, iw0); out-of-frame pixels are zeroed when the rows are converted below
    for (int i = tid; i < 17 * kStemRowWords; i += 128) {
      const int r = i / kStemRowWords, wd = i - r * kStemRowWords;
      const int ih = ih0 + r;
      uint32_t v = 0;
      if (ih >= 0 && ih < H) {
        const long long g0 = (long long)b * frame_bytes + ((long long)ih * W + iw0) * 3;
        const long long wi = (g0 >> 2) + wd;  // arithmetic shift: floor, also for the (only) negative case g0 = -3
        if (wi >= 0 && wi < all_words) v = __ldg(words + wi);
      }
      sIn[i] = v;
    }
    __syncthreads();
    {  // im2col row of output pixel (ty, tx) = tid: k = (kh*3+kw)*3 + c_rgb, frame bytes are BGR
      const int ty = tid >> 4, tx = tid & 15;
      __nv_bfloat16 row[32];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int r = 2 * ty + kh, ih = ih0 + r;
        const long long g0 = (long long)b * frame_bytes + ((long long)ih * W + iw0) * 3;
        const int shift = (int)(g0 & 3);  // bytes between the first loaded word and pixel (ih, iw0)
        const uint8_t* rb = reinterpret_cast<const uint8_t*>(sIn + r * kStemRowWords) + shift;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int iw = iw0 + 2 * tx + kw;
          const bool in = ih >= 0 && ih < H && iw >= 0 && iw < W;
          const uint8_t* px = rb + (2 * tx + kw) * 3;
          row[(kh * 3 + kw) * 3 + 0] = __float2bfloat16_rn(in ? (float)px[2] : 0.f);
          row[(kh * 3 + kw) * 3 + 1] = __float2bfloat16_rn(in ? (float)px[1] : 0.f);
          row[(kh * 3 + kw) * 3 + 2] = __float2bfloat16_rn(in ? (float)px[0] : 0.f);
        }
      }
#pragma unroll
      for (int k = 27; k < 32; ++k) row[k] = __float2bfloat16_rn(0.f);
      const uint4* rv = reinterpret_cast<const uint4*>(row);
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(sA + tid * 128 + ((j ^ (tid & 7)) << 4)) = rv[j & 3];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
    __syncthreads();
    if (warp == 0 && elect_one()) {
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < 4; ++j)
        umma_bf16(tmem_base, umma_desc_sw128(smem_u32(sA) + j * 32), umma_desc_sw128(smem_u32(sB) + j * 32), idesc, j ? 1u : 0u);
      umma_commit(&bar);
    }
    mbar_wait(&bar, phase, 16u);
    phase ^= 1;
    tc_fence_after();
    {
      const int r = warp * 32 + lane;
      const int oh = oh0 + (r >> 4), ow = ow0 + (r & 15);
      const bool valid = oh < oH && ow < oW;
      const int q = valid ? (b * oH + oh) * oW + ow : 0;
      epi_drain(p, sStage + warp * (alias_stage ? 4096 : kEpiStageBytes), lane, 0, C0 >> 4,
                tmem_base + ((uint32_t)(warp * 32) << 16), 0, valid, q, nullptr);
    }
    tc_fence_before();
    __syncthreads();  // TMEM, sA and sIn are reused by the next tile
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// EXPERIMENT (impl 3): 3x3 stride-1 conv whose nine taps all read ONE halo tile in shared memory.
// Tile = 16 rows x 8 cols of output; the halo box {64 ch, 10, 18} (180 rows of 128 B, SWIZZLE_128B) is loaded
// once per 64-channel chunk and tap (kh,kw) is the same buffer seen through a descriptor whose start address is
// advanced by (kh*10+kw) rows and whose 8-row groups are 10 rows (1280 B) apart.  This only works if the
// tensor core applies the 128B swizzle on absolute shared-memory address bits (as TMA does when writing).
// Unpipelined on purpose: it exists to answer that question on hardware (tests/test_gpu_conv.py).
// ------------------------------------------------------------------------------------------------
constexpr int kHaloRows = 18 * 10;
constexpr int kHaloBytes = 24 * 1024;  // 180 rows x 128 B, padded to a 1024 multiple

__global__ void __launch_bounds__(kConvThreads)
conv_halo_test_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kHaloBytes;                      // 9 taps x n_tile rows x 128 B
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + 9 * p.n_tile * 128);
  uint64_t* done_bar = full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int b = blockIdx.x / tiles_per_img, t_in = blockIdx.x - b * tiles_per_img;
  const int th = t_in / p.tiles_w;
  const int h0 = th * 16, w0 = (t_in - th * p.tiles_w) * 8;
  const int kchunks = (p.Cin + 63) >> 6;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)p.n_tile) tmem_cols <<= 1;
  if (warp == 1 && lane == 0) {
    mbar_init(full_bar, 1);
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, p.n_tile);
    for (int c = 0; c < kchunks; ++c) {
      if (c > 0) mbar_wait(done_bar, (c - 1) & 1, 1u);  // previous chunk's MMAs have finished reading smem
      mbar_expect_tx(full_bar, (uint32_t)(kHaloRows * 128 + 9 * p.n_tile * 128));
      tma_load_5d(sA, &tmA, full_bar, p.a_base[0] + c * 64, w0 - 1, h0 - 1, b, 0);
      for (int t = 0; t < 9; ++t) tma_load_3d(sB + t * p.n_tile * 128, &tmB, full_bar, c * 64, 0, t);
      mbar_wait(full_bar, c & 1, 2u);
      tc_fence_after();
      int ksteps = (p.Cin - c * 64) >> 4;
      if (ksteps > 4) ksteps = 4;
      for (int t = 0; t < 9; ++t) {
        const uint32_t a0 = smem_u32(sA) + (uint32_t)(((t / 3) * 10 + (t % 3)) * 128);
        const uint32_t b0 = smem_u32(sB) + (uint32_t)(t * p.n_tile * 128);
        for (int j = 0; j < ksteps; ++j)
          umma_bf16(tmem_base, umma_desc_sw128_sbo(a0 + j * 32, 1280), umma_desc_sw128(b0 + j * 32), idesc,
                    (c > 0 || t > 0 || j > 0) ? 1u : 0u);
      }
      umma_commit(done_bar);
    }
  }
  if (warp >= 2) {
    const int lg = warp & 3;
    const int r = lg * 32 + lane;
    const int h = h0 + (r >> 3), w = w0 + (r & 7);
    const bool valid = (h < p.tH) && (w < p.tW);
    const int q = (b * p.tH + h) * p.tW + w;
    mbar_wait(done_bar, (kchunks - 1) & 1, 4u);
    tc_fence_after();
    for (int j = 0; j < p.n_tile; j += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)j, v);
      tmem_ld_wait();
      if (valid) {
        float a[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = __uint_as_float(v[i]);
        conv_epilogue_store16(p, q, j, a);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// Bring-up / debugging twin: the same conv on CUDA cores, one thread per (pixel, 16 channels).
// Used by tests to cross-check the tensor-core kernel layer by layer on the device; the engine
// only runs it when YPB_CONV_IMPL=simt is set explicitly (never as a silent fallback).
// ------------------------------------------------------------------------------------------------
struct ConvSimtGeom {
  const __nv_bfloat16* in;
  int in_H, in_W, in_ctot, in_c_off;  // NHWC input buffer
  int k, stride, pad;
  const __nv_bfloat16* wg;            // [tap][Cout][Cin]
  int oH, oW, nB;                     // output map
};

__global__ void conv_simt_kernel(const ConvSimtGeom g, const ConvParams p) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ngrp = p.Cout / 16;
  const long long total = (long long)g.nB * g.oH * g.oW * ngrp;
  if (idx >= total) return;
  const int ng = (int)(idx % ngrp);
  const int q = (int)(idx / ngrp);
  const int b = q / (g.oH * g.oW);
  const int rem = q - b * g.oH * g.oW;
  const int oh = rem / g.oW, ow = rem - oh * g.oW;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  for (int kh = 0; kh < g.k; ++kh) {
    const int ih = oh * g.stride - g.pad + kh;
    if (ih < 0 || ih >= g.in_H) continue;
    for (int kw = 0; kw < g.k; ++kw) {
      const int iw = ow * g.stride - g.pad + kw;
      if (iw < 0 || iw >= g.in_W) continue;
      const __nv_bfloat16* a = g.in + (((long long)b * g.in_H + ih) * g.in_W + iw) * g.in_ctot + g.in_c_off;
      const __nv_bfloat16* wt = g.wg + ((long long)(kh * g.k + kw) * p.Cout + ng * 16) * p.Cin;
      for (int c = 0; c < p.Cin; ++c) {
        const float av = __bfloat162float(a[c]);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fmaf(av, __bfloat162float(wt[(long long)j * p.Cin + c]), acc[j]);
      }
    }
  }
  conv_epilogue_store16(p, q, ng * 16, acc);
}

}  // namespace ypb
