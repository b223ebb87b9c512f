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

#ifndef YPB_DIAG
#define YPB_DIAG 0  // 1: also build the debugging twins and micro-benchmarks (libypb200_diag.so; never in the product library)
#endif

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
  // exact division by launch constants as multiply-high + shift (x < 2^31): a runtime integer division is a ~150-cycle
  // dependent chain, and the per-tile index math of a warp-specialised role has no other warps to hide behind
  uint32_t fd_ns[2], fd_tpi[2], fd_tw[2], fd_hw[2], fd_iw[2];  // n_splits, tiles per image, tiles_w, img_HW, img_W
  int tma_out;  // 0: coalesced manual stores; 2 / 3: 32x32 warp tiles leave through a 2-D / 3-D TMA store map (1x1 convs)
  int bo_prod, bo_mma_acc, bo_mma_full, bo_epi;  // poll back-off (ns) of the four kinds of mbarrier waits (YPB_BO=a,b,c,d)
  int dbg;  // YPB_DBG experiments (0 in production): 1 = no bias/SiLU math, 2 = no output stores, 4 = no MMA issue
};

__host__ inline void fastdiv_make(uint32_t d, uint32_t (&f)[2]) {
  if (d < 1) d = 1;
  uint32_t sft = 0;
  while ((1ull << sft) < d) ++sft;
  f[0] = (uint32_t)(((1ull << (31 + sft)) / d) + 1);
  f[1] = 31 + sft;
}
__device__ __forceinline__ int fdiv(int x, const uint32_t (&f)[2]) {
  return (int)(((unsigned long long)(uint32_t)x * f[0]) >> f[1]);
}
__host__ inline void conv_set_fastdiv(ConvParams& p, int n_splits) {
  fastdiv_make((uint32_t)n_splits, p.fd_ns);
  fastdiv_make((uint32_t)(p.tiles_h * p.tiles_w), p.fd_tpi);
  fastdiv_make((uint32_t)p.tiles_w, p.fd_tw);
  fastdiv_make((uint32_t)p.img_HW, p.fd_hw);
  fastdiv_make((uint32_t)p.img_W, p.fd_iw);
}

#if YPB_DIAG
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
#endif  // YPB_DIAG

// YPB_DBG bit 8: per-role wait-cycle accounting (summed over CTAs; read back by ypb_conv_bench).
// [0] producer-0 wait empty  [1] MMA wait full  [2] MMA wait tempty  [3] epilogue-warp-0 wait tfull
// [4] epilogue-warp-0 drain  [5] CTA lifetime  [6] MMA wait weights (halo kernel)  [7] CTAs
__device__ unsigned long long g_conv_prof[16];  // [8..13] epilogue pass phases (all epilogue warps): ld+wait, release, math, stage, write-out, passes
#ifndef YPB_PROF
#define YPB_PROF 0  // 1: build the wait-cycle / epilogue-phase accounting in (libypb200_prof.so, tools/conv_layers.py --dbg 8)
#endif
constexpr bool kProf = YPB_PROF != 0;
#define PROF_T0() const long long _pt0 = prof ? clock64() : 0
#define PROF_ADD(var) do { if (prof) var += clock64() - _pt0; } while (0)

constexpr int kConvThreads = 192;
constexpr int kATileBytes = 128 * 128;  // 128 rows x 64 bf16

__host__ __device__ inline int conv_stage_bytes(int n_tile) { return kATileBytes + n_tile * 128; }
__host__ __device__ inline int conv_smem_bytes(int n_tile, int stages) {
  return 1024 /*align slack*/ + stages * conv_stage_bytes(n_tile) + 256 /*barriers*/;
}

#if YPB_DIAG  // first-generation kernel: one tile per CTA (libypb200_diag.so only, A/B measurements)
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
#endif  // YPB_DIAG

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
#ifndef YPB_EPI_WARPS
#define YPB_EPI_WARPS 8
#endif
constexpr int kEpiWarps = YPB_EPI_WARPS;   // 4 lane groups x kEpiParts warps; each SM sub-partition hosts kEpiParts of them
constexpr int kEpiParts = kEpiWarps / 4;
constexpr int kProdWarps = 3;   // TMA producer warps: one thread sustains only ~1 box load per 0.35 us (tools/tma_bench.py)
// Warp order = issue priority: the SM sub-partition arbiter serves the highest warp id first (B300_MICROARCH "hi-wid-first"),
// so the single-lane MMA issuer gets the LAST warp (it shares its sub-partition with two busy epilogue warps and was
// measured at ~105 cycles per 64-cycle MMA when it had the lowest id), the producers come next, the epilogue warps first.
constexpr int kEpiWarp0 = 0;
constexpr int kProdWarp0 = kEpiWarps;
constexpr int kMmaWarp = kEpiWarps + kProdWarps;
constexpr int kConv2Threads = 32 * (kProdWarps + 1 + kEpiWarps);
constexpr int kEpiStageBytes = 32 * (128 + 16);  // per-warp staging tile of an fp32 pass: 32 rows x (128 B + 16 B pad)
// per-warp staging bytes by output mode: a bf16 pass stages 32 rows x (64 B + 16 B pad)
// (flat = 1x1 conv: room for two dense, swizzled 32-row tiles, the double buffer of the TMA-store path)
__host__ __device__ constexpr int epi_stage_bytes(bool f32, bool flat = false) {
  return flat ? (f32 ? 2 * 32 * 128 : 2 * 32 * 64) : (f32 ? 32 * (128 + 16) : 32 * (64 + 16));
}
// Which 128-row sub-tile and which 16-column chunks [cb, ce) of it epilogue part `part` drains (nch = n_tile / 16).
// The kEpiParts warps of a TMEM lane group are spread over the msub sub-tiles first, then over column ranges
// (multiples of 32 columns when there are enough of them).
__host__ __device__ inline void epi_split(int msub, int nch, int part, int* sidx, int* cb, int* ce) {
  const int pps = kEpiParts / msub < 1 ? 1 : kEpiParts / msub;  // parts per sub-tile
  *sidx = part / pps;
  const int q = part - *sidx * pps;
  int per = (nch + pps - 1) / pps;
  if (nch >= 2 * pps) per = (per + 1) & ~1;
  *cb = q * per < nch ? q * per : nch;
  *ce = *cb + per < nch ? *cb + per : nch;
  if (*sidx >= msub) { *sidx = 0; *cb = *ce = 0; }  // more parts than work: idle
}
__host__ __device__ inline int conv_bias_smem(int cout) { return (cout * 4 + 15) & ~15; }

__host__ __device__ inline int conv2_acc_stride(int n_tile) { return (n_tile + 31) & ~31; }
__host__ __device__ inline int conv2_smem_bytes(int n_tile, int stages) {
  return 1024 /*align slack*/ + stages * conv_stage_bytes(n_tile) + 256 /*barriers*/ + kEpiWarps * kEpiStageBytes;  // + bias
}

// ------------------------------------------------------------------------------------------------
// Epilogue of one 128-row accumulator (sub)tile for one warp (shared by the persistent kernels and the stem).
// The warp owns TMEM lanes [32*lg, +32) (t_addr points at them) and the 16-column chunks [c_begin, c_end), which it
// drains in passes of 32 (plus a final 16) columns:
//   phase 1: ONE tcgen05.ld per pass, issued one pass ahead (the load of pass k+1 is in flight while pass k is
//            staged and written out); bias from shared memory (pre-scaled), SiLU on 32 independent values per lane,
//            packed bf16 / fp32 rows into the warp's private, padded smem staging tile (lane = row, conflict-free
//            thanks to the +16 B row pitch);
//   phase 2: the staged rows are written back with lanes running along the channel dimension, so every store
//            instruction covers whole NHWC row segments (a thread-per-row store touches 32 lines per instruction).
//            Row offsets (relative to the first valid row of the warp) are exchanged with shuffles ONCE per
//            sub-tile; all loads of a pass are issued before the first store.
// `release` (may be null) is arrived on as soon as the last tcgen05.ld of this call has landed, handing the
// accumulator back to the MMA warp.  `sbias` = shared-memory copy of the bias, indexed by absolute output channel and
// pre-multiplied by 0.5 for SiLU layers with bf16 output (h = x/2 comes straight out of one FFMA).
// MODE is a compile-time switch: one straight-line instruction stream per output layout.
// Measured (tools/conv_layers.py --dbg 8, B200): the first versions of this epilogue (chunk-at-a-time tcgen05.ld ->
// wait -> bias __ldg -> math; generic-address staging; per-piece shuffle -> load -> branch -> store chains) cost
// ~2500 cycles per 32x32 pass and bounded every layer with K <= 256, although TMEM reads (64 B/clk/SM) and the SFU
// (16 SiLU/clk/SM) allow ~16 elements per clock.
// ------------------------------------------------------------------------------------------------
// MODE = output layout (bits 0-1) | SiLU (bit 2): both are compile-time so that the pass is one straight-line stream and
// the x/2 pre-scale of the SiLU is an immediate operand (the immediate form of FFMA issues at twice the rate).
enum EpiMode : int { EPI_BF16 = 0, EPI_BF16_RES = 1, EPI_F32 = 2, EPI_SHUFFLE2 = 3, EPI_ACT = 4 };
__host__ __device__ inline int epi_mode_of(int out_mode, bool has_res, int act) {
  const int lay = out_mode == OUT_F32 ? EPI_F32 : out_mode == OUT_SHUFFLE2_BF16 ? EPI_SHUFFLE2 : has_res ? EPI_BF16_RES : EPI_BF16;
  return lay | (act ? EPI_ACT : 0);
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// read-only data (the bias copy): not volatile, no memory clobber, so the compiler may schedule it freely
__device__ __forceinline__ float4 lds128_ro(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

template <int NC>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, uint32_t (&v)[32]);
template <>
__device__ __forceinline__ void tmem_ldn<16>(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
template <>
__device__ __forceinline__ void tmem_ldn<32>(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// Row-offset exchange for a pass layout with PPR 16-byte pieces per row: piece = lane + 32*i covers row piece / PPR.
template <int PPR>
__device__ __forceinline__ void epi_row_deltas(int delta, int lane, int (&dl)[PPR]) {
#pragma unroll
  for (int i = 0; i < PPR; ++i) dl[i] = __shfl_sync(0xffffffffu, delta, (lane + 32 * i) / PPR);
}

// bias + activation on NC accumulator values, packed into the staging tile, then written out.
//   stage_sa : shared-space address of the warp's staging tile;  sb_sa : shared-space address of sbias[n]
//   dl / rl  : per-piece row deltas of the output / residual (element offsets from base / res_base; < 0 = padding row)
template <int NC, int MODE>
__device__ __forceinline__ void epi_store_pass(const ConvParams& p, const uint32_t (&v)[32], uint32_t stage_sa, uint32_t sb_sa,
                                               int lane, int n, long long base, long long res_base,
                                               const int (&dl)[(MODE & 3) == EPI_F32 ? NC / 4 : NC / 8],
                                               const int (&rl)[(MODE & 3) == EPI_F32 ? NC / 4 : NC / 8], bool prof, long long& pt,
                                               long long (&pacc)[6], const CUtensorMap* tmO = nullptr, int tma_mode = 0,
                                               int tq = 0, int tb = 0, int tbuf = 0) {
#define EPI_MARK(k) do { if (prof) { const long long _n = clock64(); pacc[(k) - 8] += _n - pt; pt = _n; } } while (0)
  constexpr int LAY = MODE & 3;
  constexpr bool F32 = LAY == EPI_F32, ACT = (MODE & EPI_ACT) != 0;
  constexpr int PPR = F32 ? NC / 4 : NC / 8;  // 16-byte pieces per staged row
  constexpr int pitch = PPR * 16 + 16;
  // residual: coalesced 16-byte loads, issued first so their latency hides behind the math
  uint4 rres[PPR];
  if (LAY == EPI_BF16_RES) {
#pragma unroll
    for (int i = 0; i < PPR; ++i) {
      const int pc = (lane + 32 * i) % PPR;
      rres[i] = make_uint4(0, 0, 0, 0);
      if (rl[i] >= 0) rres[i] = __ldg(reinterpret_cast<const uint4*>(p.res + res_base + rl[i] + n + pc * 8));
    }
  }
  float y[NC];
#pragma unroll
  for (int i = 0; i < NC / 4; ++i) {
    const float4 b4 = lds128_ro(sb_sa + 16 * i);
    if (ACT && !F32) {  // h = x/2 (the bias copy is pre-halved)
      y[4 * i + 0] = fmaf(__uint_as_float(v[4 * i + 0]), 0.5f, b4.x);
      y[4 * i + 1] = fmaf(__uint_as_float(v[4 * i + 1]), 0.5f, b4.y);
      y[4 * i + 2] = fmaf(__uint_as_float(v[4 * i + 2]), 0.5f, b4.z);
      y[4 * i + 3] = fmaf(__uint_as_float(v[4 * i + 3]), 0.5f, b4.w);
    } else {
      y[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + b4.x;
      y[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b4.y;
      y[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b4.z;
      y[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b4.w;
    }
  }
  if (ACT && !(p.dbg & 1)) {
    if (!F32) {
#pragma unroll
      for (int i = 0; i < NC; ++i) {  // y holds h = x/2: SiLU(x) = h + h*tanh(h)
#if YPB_EXACT_SILU  // A/B build (libypb200_exact.so): full-precision x * sigmoid(x), see tools/silu_ab.md
        const float x2 = 2.0f * y[i];
        y[i] = x2 / (1.0f + expf(-x2));
#else
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(y[i]));
        y[i] = fmaf(y[i], t, y[i]);
#endif
      }
    } else {
#pragma unroll
      for (int i = 0; i < NC; ++i) y[i] = silu_precise_f(y[i]);
    }
  }
  if (prof) {  // make the math retire before the timestamp
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) acc += y[i];
    if (acc == 123.456f) pt += 1;
  }
  EPI_MARK(10);
  if (NC == 32 && (LAY == EPI_BF16 || LAY == EPI_F32) && tma_mode != 0) {
    // TMA-store path (1x1 convs): the warp's 32 rows x 32 columns go to one of two dense staging tiles in the TMA
    // swizzle (64-byte rows: SWIZZLE_64B, 128-byte rows: SWIZZLE_128B - also what keeps the 16-byte row-per-lane
    // writes conflict-free), then ONE bulk tensor store writes them out; rows past the end of the tensor are clipped
    // by the TMA unit.  The other tile may still be draining: wait only for the store that last read this one.
    constexpr int ROWB = F32 ? 128 : 64;
    const uint32_t tile = stage_sa + (uint32_t)tbuf * (32 * ROWB);
    if (lane == 0) bulk_wait_read<1>();
    __syncwarp();
    const uint32_t row = tile + lane * ROWB;
    const int sw = F32 ? (lane & 7) : ((lane >> 1) & 3);
    if (!F32) {
#pragma unroll
      for (int i = 0; i < PPR; ++i)
        sts128(row + ((i ^ sw) << 4), pack_bf16x2(y[8 * i], y[8 * i + 1]), pack_bf16x2(y[8 * i + 2], y[8 * i + 3]),
               pack_bf16x2(y[8 * i + 4], y[8 * i + 5]), pack_bf16x2(y[8 * i + 6], y[8 * i + 7]));
    } else {
#pragma unroll
      for (int i = 0; i < PPR; ++i)
        sts128(row + ((i ^ sw) << 4), __float_as_uint(y[4 * i]), __float_as_uint(y[4 * i + 1]), __float_as_uint(y[4 * i + 2]),
               __float_as_uint(y[4 * i + 3]));
    }
    fence_proxy_async_smem();
    __syncwarp();
    EPI_MARK(11);
    if (lane == 0 && !(p.dbg & 2)) {
      if (tma_mode == 2) tma_store_2d(tmO, tile, n, tq);
      else tma_store_3d(tmO, tile, n, tq, tb);
      bulk_commit();
    }
    EPI_MARK(12);
    if (prof) pacc[5] += 1;
    return;
  }
  const uint32_t my = stage_sa + lane * pitch;
  if (!F32) {
#pragma unroll
    for (int i = 0; i < PPR; ++i)
      sts128(my + 16 * i, pack_bf16x2(y[8 * i], y[8 * i + 1]), pack_bf16x2(y[8 * i + 2], y[8 * i + 3]),
             pack_bf16x2(y[8 * i + 4], y[8 * i + 5]), pack_bf16x2(y[8 * i + 6], y[8 * i + 7]));
  } else {
#pragma unroll
    for (int i = 0; i < PPR; ++i)
      sts128(my + 16 * i, __float_as_uint(y[4 * i]), __float_as_uint(y[4 * i + 1]), __float_as_uint(y[4 * i + 2]),
             __float_as_uint(y[4 * i + 3]));
  }
  __syncwarp();
  EPI_MARK(11);
  if (!(p.dbg & 2)) {
    uint4 val[PPR];
#pragma unroll
    for (int i = 0; i < PPR; ++i) {
      const int piece = lane + 32 * i;
      val[i] = lds128(stage_sa + (piece / PPR) * pitch + (piece % PPR) * 16);
    }
    const int pc = lane % PPR;  // (lane + 32 i) % PPR is the same for every i: PPR divides 32
    if (F32) {
      float* out = reinterpret_cast<float*>(p.out) + base + n + pc * 4;
#pragma unroll
      for (int i = 0; i < PPR; ++i)
        if (dl[i] >= 0) *reinterpret_cast<uint4*>(out + dl[i]) = val[i];
    } else {
      const int nn = n + pc * 8;
      long long coff = nn;
      if (LAY == EPI_SHUFFLE2) {  // ConvTranspose2d(2,2): channel group g = (dy, dx) of the 2x2 output block
        const int cq = p.Cout >> 2;
        const int g = nn / cq, c = nn - g * cq;
        coff = ((long long)(g >> 1) * (2 * p.img_W) + (g & 1)) * p.out_pix_stride + c;
      }
      __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + base + coff;
#pragma unroll
      for (int i = 0; i < PPR; ++i) {
        if (LAY == EPI_BF16_RES) {
          // y = bf16(act(...)) first, then bf16(y + res): the same two roundings as storing the conv output and
          // adding the shortcut afterwards (Bottleneck: x + cv2(cv1(x))).
          const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&val[i]);
          const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&rres[i]);
          uint32_t o[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float2 fa = __bfloat1622float2(a2[u]), fb = __bfloat1622float2(b2[u]);
            o[u] = pack_bf16x2(fa.x + fb.x, fa.y + fb.y);
          }
          val[i] = make_uint4(o[0], o[1], o[2], o[3]);
        }
        if (dl[i] >= 0) *reinterpret_cast<uint4*>(out + dl[i]) = val[i];
      }
    }
  }
  __syncwarp();
  EPI_MARK(12);
  if (prof) pacc[5] += 1;
#undef EPI_MARK
}

__device__ __forceinline__ void epi_release(uint64_t* release, uint32_t release_remote) {
  if (release_remote != 0u) mbar_arrive_cluster(release_remote);
  else mbar_arrive(release);
}

template <int MODE, bool PINGPONG = true>
__device__ __forceinline__ void epi_drain(const ConvParams& p, uint32_t stage_sa, uint32_t sbias_sa, int lane, int c_begin,
                                          int c_end, uint32_t t_addr, int n0, bool valid, int qb, int rem,
                                          uint64_t* release, long long (&pacc)[6], const CUtensorMap* tmO = nullptr,
                                          int tma_mode = 0, int tq = 0, int tb = 0, int* tbuf = nullptr,
                                          uint32_t release_remote = 0) {
  // release_remote != 0: the accumulator is handed back with an arrive on a barrier of ANOTHER CTA of the cluster (its
  // shared::cluster address; the 2-CTA kernel's peer signals the leader's MMA warp) instead of the local `release`
  constexpr int LAY = MODE & 3;
  constexpr bool F32 = LAY == EPI_F32;
  constexpr int P32 = F32 ? 8 : 4, P16 = F32 ? 4 : 2;  // 16-byte pieces per staged row of a 32- / 16-column pass
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);
  if (c_begin >= c_end || vmask == 0u) {  // nothing to read (more warps than column chunks; tile rows all padding)
    __syncwarp();
    if (lane == 0 && release != nullptr) epi_release(release, release_remote);
    return;
  }
  const bool prof = kProf && (p.dbg & 8) != 0 && lane == 0;
  long long pt = prof ? clock64() : 0;
#define EPI_MARK(k) do { if (prof) { const long long _n = clock64(); pacc[(k) - 8] += _n - pt; pt = _n; } } while (0)
  // (qb, rem) = image and pixel-within-image of this lane's row; element offset of channel 0 of that row in the
  // output (pixel-shuffle: of sub-pixel (0,0))
  long long off_row;
  if (LAY == EPI_SHUFFLE2) {
    const int ph = fdiv(rem, p.fd_iw), pw = rem - ph * p.img_W;
    off_row = qb * p.out_img_stride + ((long long)(2 * ph) * (2 * p.img_W) + 2 * pw) * p.out_pix_stride + p.out_c_off;
  } else {
    off_row = qb * p.out_img_stride + (long long)rem * p.out_pix_stride + p.out_c_off;
  }
  const int src = __ffs(vmask) - 1;  // rows are addressed relative to the first valid row of the warp (offsets only grow)
  const long long base = __shfl_sync(0xffffffffu, off_row, src);
  const int delta = valid ? (int)(off_row - base) : -1;
  long long res_base = 0;
  int rdelta = -1;
  if (LAY == EPI_BF16_RES) {
    const long long res_row = qb * p.res_img_stride + (long long)rem * p.res_pix_stride + p.res_c_off;
    res_base = __shfl_sync(0xffffffffu, res_row, src);
    rdelta = valid ? (int)(res_row - res_base) : -1;
  }
  const int col0 = c_begin * 16, ncols = (c_end - c_begin) * 16;
  const int n32 = ncols >> 5;
  const bool tail16 = (ncols & 31) != 0;
  uint32_t va[32];
  if (!PINGPONG && n32 > 0) {  // register-lean variant (several CTAs per SM, e.g. the stem): one pass at a time
    int dl[P32], rl[P32];
    epi_row_deltas<P32>(delta, lane, dl);
    if (LAY == EPI_BF16_RES) epi_row_deltas<P32>(rdelta, lane, rl);
    for (int k = 0; k < n32; ++k) {
      const int col = col0 + 32 * k;
      tmem_ldn<32>(t_addr + (uint32_t)col, va);
      tmem_ld_wait();
      if (k + 1 == n32 && !tail16 && release != nullptr) {
        __syncwarp();
        if (lane == 0) epi_release(release, release_remote);
      }
      epi_store_pass<32, MODE>(p, va, stage_sa, sbias_sa + (uint32_t)(n0 + col) * 4, lane, n0 + col, base, res_base, dl, rl,
                               prof, pt, pacc);
    }
  }
  if (PINGPONG && n32 > 0) {
    uint32_t vb[32];  // ping-pong accumulator registers: the tcgen05.ld of pass k+1 is in flight during pass k
    int dl[P32], rl[P32];
    epi_row_deltas<P32>(delta, lane, dl);
    if (LAY == EPI_BF16_RES) epi_row_deltas<P32>(rdelta, lane, rl);
    tmem_ldn<32>(t_addr + (uint32_t)col0, va);
    auto pass32 = [&](uint32_t (&cur)[32], uint32_t (&nxt)[32], int k) {
      const int col = col0 + 32 * k;
      tmem_ld_wait();
      EPI_MARK(8);
      if (k + 1 < n32) {
        tmem_ldn<32>(t_addr + (uint32_t)(col + 32), nxt);
      } else if (!tail16 && release != nullptr) {
        // last TMEM read of this tile has landed (tcgen05.wait::ld): hand the accumulator back to the MMA warp
        __syncwarp();
        if (lane == 0) epi_release(release, release_remote);
      }
      EPI_MARK(9);
      int tb_idx = 0;
      if (tma_mode != 0) {
        tb_idx = *tbuf;
        *tbuf ^= 1;
      }
      epi_store_pass<32, MODE>(p, cur, stage_sa, sbias_sa + (uint32_t)(n0 + col) * 4, lane, n0 + col, base, res_base, dl, rl,
                               prof, pt, pacc, tmO, tma_mode, tq, tb, tb_idx);
    };
    for (int k = 0; k < n32; k += 2) {
      pass32(va, vb, k);
      if (k + 1 < n32) pass32(vb, va, k + 1);
    }
  }
  if (tail16) {
    const int col = col0 + 32 * n32;
    int dl[P16], rl[P16];
    epi_row_deltas<P16>(delta, lane, dl);
    if (LAY == EPI_BF16_RES) epi_row_deltas<P16>(rdelta, lane, rl);
    tmem_ldn<16>(t_addr + (uint32_t)col, va);
    tmem_ld_wait();
    if (release != nullptr) {
      __syncwarp();
      if (lane == 0) epi_release(release, release_remote);
    }
    if (tma_mode != 0) {  // the manual pass reuses the staging memory the bulk stores read from
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
    }
    epi_store_pass<16, MODE>(p, va, stage_sa, sbias_sa + (uint32_t)(n0 + col) * 4, lane, n0 + col, base, res_base, dl, rl, prof,
                             pt, pacc);
  }
#undef EPI_MARK
}

// L2 prefetch of the residual row segment this lane's accumulator row will be added to (channels [n, n + ncols)).
// Issued one tile ahead: the shortcut tensor was written several layers earlier and has left L2 by now, and a
// DRAM-latency load inside the epilogue pass cannot be covered by the pass's own math.
__device__ __forceinline__ void epi_prefetch_res(const ConvParams& p, bool valid, int qb, int rem, int n, int ncols) {
  if (!valid) return;
  const __nv_bfloat16* r = p.res + qb * p.res_img_stride + (long long)rem * p.res_pix_stride + p.res_c_off + n;
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(r) & ~uintptr_t(127), a1 = reinterpret_cast<uintptr_t>(r + ncols) - 1;
  for (uintptr_t a = a0; a <= a1; a += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
}

// Cooperative copy of the (pre-scaled) bias into shared memory; call before the setup __syncthreads().
__device__ __forceinline__ void epi_load_bias(const ConvParams& p, float* sbias) {
  const float sc = (p.act && p.out_mode != OUT_F32) ? 0.5f : 1.0f;
  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) sbias[i] = sc * __ldg(p.bias + i);
}

template <int MODE>
__global__ void __launch_bounds__(kConv2Threads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmO, const __grid_constant__ ConvParams p, int n_splits, int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = p.msub * kATileBytes;
  const int stage_bytes = a_bytes + p.n_tile * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tfull_bar = empty_bar + p.stages;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;         // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  // the shuffle makes the warp index provably warp-uniform: role branches become uniform branches and the role
  // bodies may keep their loop state on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int kchunks = (p.Cin + 63) >> 6;
  const int acc_stride = conv2_acc_stride(p.n_tile);
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * p.msub * acc_stride)) tmem_cols <<= 1;
  const bool prof = kProf && (p.dbg & 8) != 0;
  const long long prof_start = prof ? clock64() : 0;
  long long pw0 = 0, pw1 = 0;

  if (warp == kProdWarp0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_out != 0) tma_prefetch_desc(&tmO);
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
  const int kStageB = epi_stage_bytes((MODE & 3) == EPI_F32, p.ntaps == 1);
  // epilogue staging tiles start on a 1 KB boundary (TMA-store swizzle atoms)
  uint8_t* stage0 = smem + ((p.stages * stage_bytes + 256 + 1023) & ~1023);
  float* sbias = reinterpret_cast<float*>(stage0 + kEpiWarps * kStageB);
  epi_load_bias(p, sbias);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above touched only launch constants (weights' bias, tensor maps) and on-chip state: it overlaps the
  // tail of the previous kernel; activations are read / written only after the predecessor has fully completed
  pdl_trigger();
  pdl_wait();

  const int pw = warp - kProdWarp0;  // producer index
  if (pw >= 0 && pw < kProdWarps) {
    // ===================== TMA producers: producer w owns the ring stages s with s % kProdWarps == w =====================
    if (elect_one()) {
      const uint32_t tx_bytes = (uint32_t)(p.TH * p.TW * 128 + p.n_tile * 128);  // the A box spans all sub-tiles
      int s = -1;
      uint32_t ph = 1;  // ring position / phase of the current k-iteration (advanced at the top of the loop body)
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = fdiv(tile, p.fd_ns), n0 = (tile - mt * n_splits) * p.n_tile;
        const int b = fdiv(mt, p.fd_tpi), t_in = mt - b * tiles_per_img;
        const int th = fdiv(t_in, p.fd_tw);
        const int h0 = th * p.TH, w0 = (t_in - th * p.tiles_w) * p.TW;
        int cbase[5];
#pragma unroll
        for (int d = 0; d < 5; ++d) cbase[d] = p.a_base[d] + b * p.a_cb[d] + h0 * p.a_ch[d] + w0 * p.a_cw[d];
        for (int c = 0; c < kchunks; ++c) {
          for (int t = 0; t < p.ntaps; ++t) {
            if (++s == p.stages) s = 0;
            if (s == 0) ph ^= 1;
            // a stage always belongs to the same producer: parity waits are only sound one phase ahead
            if ((s % kProdWarps) != pw) continue;
            {
              PROF_T0();
              mbar_wait_bo(empty_bar + s, ph ^ 1, 1u, p.bo_prod);
              PROF_ADD(pw0);
            }
            uint8_t* sa = smem + s * stage_bytes;
            mbar_expect_tx(full_bar + s, tx_bytes);
            tma_load_5d(sa, &tmA, full_bar + s, cbase[0] + p.tap[t][0] + c * 64, cbase[1] + p.tap[t][1],
                        cbase[2] + p.tap[t][2], cbase[3] + p.tap[t][3], cbase[4] + p.tap[t][4]);
            tma_load_3d(sa + a_bytes, &tmB, full_bar + s, c * 64, n0, t);
          }
        }
      }
      if (prof && pw == 0) atomicAdd(&g_conv_prof[0], (unsigned long long)pw0);
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (accumulation order: chunk-major, tap-minor, like conv3_halo_kernel) =====================
    // The whole warp runs the loops converged (uniform-datapath address arithmetic); one elected lane - always the
    // same one for a full member mask - issues the tcgen05 instructions.  With the loops inside a single-lane region
    // ptxas computed every descriptor in vector registers and moved it over with R2UR: ~100 issue cycles per MMA,
    // more than the 64 cycles an M=128, N<=128 MMA takes.
    {
      const uint32_t idesc = umma_idesc_bf16(128, p.n_tile);
      const int msub = p.msub;
      const uint32_t sub_bytes = (uint32_t)(p.sub_rows * 128);
      int s = -1, acc = 0;
      uint32_t ph = 1;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acc) {
        const int buf = acc & 1;
        {
          PROF_T0();
          mbar_wait_warp(tempty_bar + buf, ((acc >> 1) & 1) ^ 1, 8u, p.bo_mma_acc);  // epilogue has drained this accumulator
          PROF_ADD(pw1);
        }
        tc_fence_after();
        uint32_t accf = 0;  // the first MMA of a tile overwrites the accumulators
        const uint32_t d_tmem0 = tmem_base + (uint32_t)(buf * msub * acc_stride);
        for (int c = 0; c < kchunks; ++c) {
          int ksteps = (p.Cin - c * 64) >> 4;
          if (ksteps > 4) ksteps = 4;
          if (p.dbg & 4) ksteps = 0;
          for (int t = 0; t < p.ntaps; ++t) {
            if (++s == p.stages) s = 0;
            if (s == 0) ph ^= 1;
            {
              PROF_T0();
              mbar_wait_warp(full_bar + s, ph, 2u, p.bo_mma_full);
              PROF_ADD(pw0);
            }
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + s * stage_bytes);
            const uint32_t b_lo = umma_desc_lo(sa + a_bytes);
            const uint32_t a_lo0 = umma_desc_lo(sa), a_lo1 = umma_desc_lo(sa + sub_bytes);
            constexpr uint32_t hi = umma_desc_hi(1024);
            if (elect_one()) {
              umma_bf16_ksteps_n(ksteps, d_tmem0, a_lo0, hi, b_lo, hi, idesc, accf);
              if (msub > 1) umma_bf16_ksteps_n(ksteps, d_tmem0 + (uint32_t)acc_stride, a_lo1, hi, b_lo, hi, idesc, accf);
              umma_commit(empty_bar + s);
            }
            accf = 1;
          }
        }
        if (elect_one()) umma_commit(tfull_bar + buf);
      }
      if (prof && elect_one()) {
        atomicAdd(&g_conv_prof[1], (unsigned long long)pw0);
        atomicAdd(&g_conv_prof[2], (unsigned long long)pw1);
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
    int sidx, c_begin, c_end;  // this warp's sub-tile and 16-column chunk range
    epi_split(p.msub, p.n_tile >> 4, (warp - kEpiWarp0) >> 2, &sidx, &c_begin, &c_end);
    uint8_t* stage = stage0 + (warp - kEpiWarp0) * kStageB;
    int tbuf = 0;  // which of the two TMA-store staging tiles the next pass uses
    const int r = lg * 32 + lane;
    long long pacc[6] = {0, 0, 0, 0, 0, 0};
    const bool flat = p.ntaps == 1;  // 1x1: the tile is 128 * msub consecutive pixels of the flattened batch
    const int R = sidx * p.sub_rows + r;  // this lane's row inside the CTA rectangle (tile-invariant)
    const int rh = R / p.TW, rw = R - rh * p.TW;
    int acc = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acc) {
      const int mt = fdiv(tile, p.fd_ns), n0 = (tile - mt * n_splits) * p.n_tile;
      const int b = fdiv(mt, p.fd_tpi), t_in = mt - b * tiles_per_img;
      const int th = fdiv(t_in, p.fd_tw);
      const int buf = acc & 1;
      if ((MODE & 3) == EPI_BF16_RES && tile + (int)gridDim.x < total_tiles && c_begin < c_end) {
        const int tile2 = tile + gridDim.x;
        const int mt2 = fdiv(tile2, p.fd_ns), n2 = (tile2 - mt2 * n_splits) * p.n_tile;
        const int b2 = fdiv(mt2, p.fd_tpi), t2 = mt2 - b2 * tiles_per_img;
        const int th2 = fdiv(t2, p.fd_tw);
        const int h = th2 * p.TH + rh, w = (t2 - th2 * p.tiles_w) * p.TW + rw;
        const bool valid = (r < p.sub_rows) && (h < p.tH) && (w < p.tW);
        int qb = b2, rem = h * p.tW + w;
        if (flat) {
          qb = fdiv(w, p.fd_hw);
          rem = w - qb * p.img_HW;
        }
        epi_prefetch_res(p, valid, qb, rem, n2 + c_begin * 16, (c_end - c_begin) * 16);
      }
      {
        PROF_T0();
        mbar_wait_bo(tfull_bar + buf, (acc >> 1) & 1, 4u, p.bo_epi);
        PROF_ADD(pw0);
      }
      tc_fence_after();
      PROF_T0();
      {
        const int h = th * p.TH + rh, w = (t_in - th * p.tiles_w) * p.TW + rw;
        const bool valid = (r < p.sub_rows) && (h < p.tH) && (w < p.tW);
        int qb = 0, rem = 0;
        if (valid) {
          if (flat) {
            qb = fdiv(w, p.fd_hw);
            rem = w - qb * p.img_HW;
          } else {
            qb = b;
            rem = h * p.tW + w;
          }
        }
        const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((buf * p.msub + sidx) * acc_stride);
        int tq = 0, tb = 0;
        if (p.tma_out != 0) {  // flat tiles only: pixel index of this warp's first row
          tq = t_in * p.TW + sidx * p.sub_rows + lg * 32;
          if (p.tma_out == 3) {
            tb = fdiv(tq, p.fd_hw);
            tq -= tb * p.img_HW;
          }
        }
        epi_drain<MODE>(p, smem_u32(stage), smem_u32(sbias), lane, c_begin, c_end, t_addr, n0, valid, qb, rem,
                        tempty_bar + buf, pacc, &tmO, p.tma_out, tq, tb, &tbuf);
      }
      PROF_ADD(pw1);
    }
    if (p.tma_out != 0 && lane == 0) bulk_wait_all();  // the bulk stores read this CTA's shared memory until they complete
    if (prof && warp == kEpiWarp0 && lane == 0) {
      atomicAdd(&g_conv_prof[3], (unsigned long long)pw0);
      atomicAdd(&g_conv_prof[4], (unsigned long long)pw1);
    }
    if (prof && lane == 0) {
#pragma unroll
      for (int i = 0; i < 6; ++i) atomicAdd(&g_conv_prof[8 + i], (unsigned long long)pacc[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (prof && threadIdx.x == 0) {
    atomicAdd(&g_conv_prof[5], (unsigned long long)(clock64() - prof_start));
    atomicAdd(&g_conv_prof[7], 1ull);
  }
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
  // A-operand addressing of the nine taps inside the halo buffer, in descriptor units (16 bytes):
  //   stride 1: box {64 ch, 10, 16 msub + 2}; tap (kh, kw) starts (kh*10 + kw) rows in, 8-pixel groups 1280 B apart
  //   stride 2 (Cin = 32, `s2`): the input is viewed as (2C = 64, W/2, 2, H/2, B) - a row of the view is a PAIR of
  //     input columns, [even column's 32 channels | odd column's 32 channels] - and the box is {64, 10, 2, 16 msub + 1}:
  //     tap (kh, kw) reads row-pair r = (kh > 0), row parity (kh != 1), column pair (kw > 0) and the channel half
  //     (kw != 1) as a 64-byte K offset; output rows are 2 x 10 x 128 B apart
  int s2;
  int tap_off[9];
  int sub_off;     // second sub-tile (16 output rows further down)
  uint32_t a_hi;   // high descriptor word of A: SBO = 1280 (stride 1) or 2560 (stride 2)
};

// All nine taps of one 64-channel chunk against resident weights, straight-line: 9 x KS (x 2 sub-tiles) MMAs whose
// descriptors are the chunk's base words plus launch constants.  Called by the elected lane.
template <int KS, bool TWO>
__device__ __forceinline__ void halo_issue_taps(const Conv3Extra& x, uint32_t d_tmem0, uint32_t acc_stride, uint32_t a_lo0,
                                                uint32_t a_hi, uint32_t b_lo0, uint32_t tap_units, uint32_t idesc, uint32_t accf) {
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const uint32_t a_lo = a_lo0 + (uint32_t)x.tap_off[t], b_lo = b_lo0 + (uint32_t)t * tap_units;
    umma_bf16_ksteps<KS>(d_tmem0, a_lo, a_hi, b_lo, umma_desc_hi(1024), idesc, t == 0 ? accf : 1u);
    if (TWO)
      umma_bf16_ksteps<KS>(d_tmem0 + acc_stride, a_lo + (uint32_t)x.sub_off, a_hi, b_lo, umma_desc_hi(1024), idesc,
                           t == 0 ? accf : 1u);
  }
}

// One kernel row (three taps, stride 1) against one streamed weight box, straight-line.  Called by the elected lane.
template <int KS, bool TWO>
__device__ __forceinline__ void halo_issue_row(uint32_t d_tmem0, uint32_t acc_stride, uint32_t a_lo, uint32_t sub_off, uint32_t a_hi,
                                               uint32_t b_lo, uint32_t tap_units, uint32_t idesc, uint32_t accf) {
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const uint32_t a = a_lo + (uint32_t)(u * (128 >> 4)), bq = b_lo + (uint32_t)u * tap_units;
    umma_bf16_ksteps<KS>(d_tmem0, a, a_hi, bq, umma_desc_hi(1024), idesc, u == 0 ? accf : 1u);
    if (TWO) umma_bf16_ksteps<KS>(d_tmem0 + acc_stride, a + sub_off, a_hi, bq, umma_desc_hi(1024), idesc, u == 0 ? accf : 1u);
  }
}

template <int MODE>
__global__ void __launch_bounds__(kConv2Threads, 1)
conv3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ ConvParams p, const __grid_constant__ Conv3Extra x, int n_splits, int total_tiles) {
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

  // the shuffle makes the warp index provably warp-uniform: role branches become uniform branches and the role
  // bodies may keep their loop state on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int kchunks = (p.Cin + 63) >> 6;
  const int acc_stride = conv2_acc_stride(p.n_tile);
  const int tap_bytes = p.n_tile * 128;
  const int grp_bytes = x.b_group * tap_bytes;
  const int ngroups = 9 / x.b_group;  // weight boxes per 64-channel chunk
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * x.msub * acc_stride)) tmem_cols <<= 1;
  const bool prof = kProf && (p.dbg & 8) != 0;
  const long long prof_start = prof ? clock64() : 0;
  long long pw0 = 0, pw1 = 0, pw2 = 0, pw3 = 0, pw4 = 0;

  if (warp == kProdWarp0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < 4; ++i) { mbar_init(a_full + i, 1); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < 12; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar + i, 1); mbar_init(tempty_bar + i, kEpiWarps); }
    mbar_init(ball_bar, 1);
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, tmem_cols);
  constexpr int kStageB = epi_stage_bytes((MODE & 3) == EPI_F32);
  float* sbias = reinterpret_cast<float*>(stage_base + kEpiWarps * kStageB);
  epi_load_bias(p, sbias);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();  // see conv_tc2_kernel
  pdl_wait();

  const int pw = warp - kProdWarp0;  // producer index
  if (pw == 0 || (x.b_stat && pw > 0 && pw < kProdWarps)) {
    // ===================== TMA producer 0: halo tiles (and the resident weights, once) =====================
    // With resident weights the other producer warps have nothing to stream and share the halo loads: slot sa
    // belongs to producer sa % a_prods (a thread's bulk loads complete one after the other; several threads
    // keep several boxes in flight, tools/tma_bench.py).
    const int a_prods = x.b_stat ? kProdWarps : 1;
    if (elect_one()) {
      if (x.b_stat && pw == 0) {
        mbar_expect_tx(ball_bar, (uint32_t)(9 * kchunks * tap_bytes));
        for (int c = 0; c < kchunks; ++c)
          for (int g = 0; g < ngroups; ++g)
            tma_load_3d(sB + (c * 9 + g * x.b_group) * tap_bytes, &tmB, ball_bar, c * 64, 0, g * x.b_group);
      }
      int sa = -1;
      uint32_t pa = 1;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = fdiv(tile, p.fd_ns);
        const int b = fdiv(mt, p.fd_tpi), t_in = mt - b * tiles_per_img;
        const int th = fdiv(t_in, p.fd_tw);
        const int h0 = th * 16 * x.msub, w0 = (t_in - th * p.tiles_w) * 8;
        for (int c = 0; c < kchunks; ++c) {
          if (++sa == x.a_slots) sa = 0;
          if (sa == 0) pa ^= 1;
          if (sa % a_prods != pw) continue;
          {
            PROF_T0();
            mbar_wait_bo(a_empty + sa, pa ^ 1, 1u, p.bo_prod);
            PROF_ADD(pw0);
          }
          mbar_expect_tx(a_full + sa, (uint32_t)(x.halo_rows * 128));
          if (x.s2) tma_load_5d(sA + sa * x.a_bytes, &tmA, a_full + sa, 0, w0 - 1, 0, h0 - 1, b);
          else tma_load_5d(sA + sa * x.a_bytes, &tmA, a_full + sa, p.a_base[0] + c * 64, w0 - 1, h0 - 1, b, 0);
        }
      }
      if (prof && pw == 0) atomicAdd(&g_conv_prof[0], (unsigned long long)pw0);
    }
  } else if (pw > 0 && pw < kProdWarps) {
    // ===================== TMA producers 1..3: streamed weight boxes, ring slot sb owned by warp 1 + sb % 3 =====================
    if (!x.b_stat && elect_one()) {
      int sb = -1;
      uint32_t pb = 1;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = fdiv(tile, p.fd_ns), n0 = (tile - mt * n_splits) * p.n_tile;
        for (int c = 0; c < kchunks; ++c) {
          for (int g = 0; g < ngroups; ++g) {
            if (++sb == x.b_slots) sb = 0;
            if (sb == 0) pb ^= 1;
            if (1 + (sb % (kProdWarps - 1)) != pw) continue;  // a slot always belongs to the same producer
            mbar_wait_bo(b_empty + sb, pb ^ 1, 1u, p.bo_prod);
            mbar_expect_tx(b_full + sb, (uint32_t)grp_bytes);
            tma_load_3d(sB + sb * grp_bytes, &tmB, b_full + sb, c * 64, n0, g * x.b_group);
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues: see conv_tc2_kernel) =====================
    {
      const uint32_t idesc = umma_idesc_bf16(128, p.n_tile);
      const int msub = x.msub, b_group = x.b_group;
      const bool b_stat = x.b_stat != 0;
      if (b_stat) {
        mbar_wait_warp(ball_bar, 0, 2u, 0);
        tc_fence_after();
      }
      int sa = -1, sbn = -1, acc = 0;
      uint32_t pa = 1, pb = 1;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acc) {
        const int buf = acc & 1;
        {
          PROF_T0();
          mbar_wait_warp(tempty_bar + buf, ((acc >> 1) & 1) ^ 1, 8u, p.bo_mma_acc);
          PROF_ADD(pw1);
        }
        tc_fence_after();
        uint32_t accf = 0;  // the first MMA of a tile overwrites the accumulators
        const uint32_t d_tmem0 = tmem_base + (uint32_t)(buf * msub * acc_stride);
        for (int c = 0; c < kchunks; ++c) {
          if (++sa == x.a_slots) sa = 0;
          if (sa == 0) pa ^= 1;
          {
            PROF_T0();
            mbar_wait_warp(a_full + sa, pa, 2u, p.bo_mma_full);
            PROF_ADD(pw0);
          }
          tc_fence_after();
          int ksteps = (p.Cin - c * 64) >> 4;
          if (ksteps > 4) ksteps = 4;
          if (p.dbg & 4) ksteps = 0;
          const long long _ti0 = prof ? clock64() : 0;
          // tap t of sub-tile s: the halo buffer seen from tap_off[t] (+ s * sub_off) on
          const uint32_t a_lo0 = umma_desc_lo(smem_u32(sA + sa * x.a_bytes));
          const uint32_t a_hi = x.a_hi;
          if (b_stat && (ksteps == 4 || ksteps == 2)) {
            // resident weights: one elected region issues the whole chunk (the per-tap loop below costs ~100 issue
            // cycles per tap, more than the two 64-cycle MMAs of a tap of a 32-channel layer)
            const uint32_t b_lo0 = umma_desc_lo(smem_u32(sB + c * 9 * tap_bytes)), tu = (uint32_t)(tap_bytes >> 4);
            if (elect_one()) {
              if (ksteps == 4) {
                if (msub > 1) halo_issue_taps<4, true>(x, d_tmem0, (uint32_t)acc_stride, a_lo0, a_hi, b_lo0, tu, idesc, accf);
                else halo_issue_taps<4, false>(x, d_tmem0, (uint32_t)acc_stride, a_lo0, a_hi, b_lo0, tu, idesc, accf);
              } else {
                if (msub > 1) halo_issue_taps<2, true>(x, d_tmem0, (uint32_t)acc_stride, a_lo0, a_hi, b_lo0, tu, idesc, accf);
                else halo_issue_taps<2, false>(x, d_tmem0, (uint32_t)acc_stride, a_lo0, a_hi, b_lo0, tu, idesc, accf);
              }
            }
            accf = 1;
          } else {
          // streamed weights (stride 1 only) or an odd K-step count: tap by tap, the A window advanced incrementally
          // (+1 pixel per kw, +1 halo row - 10 pixels - per kh)
          uint32_t a_lo = a_lo0;
          int kw = 0;
          for (int g = 0; g < ngroups; ++g) {
            uint32_t b_lo;
            int sb = 0;
            if (b_stat) {
              b_lo = umma_desc_lo(smem_u32(sB + (c * 9 + g * b_group) * tap_bytes));
            } else {
              if (++sbn == x.b_slots) sbn = 0;
              if (sbn == 0) pb ^= 1;
              sb = sbn;
              {
                PROF_T0();
                mbar_wait_warp(b_full + sb, pb, 2u, p.bo_mma_full);
                PROF_ADD(pw2);
              }
              tc_fence_after();
              b_lo = umma_desc_lo(smem_u32(sB + sb * grp_bytes));
            }
            if (b_group == 3 && ksteps == 4) {  // a whole kernel row per weight box: 12 / 24 MMAs back to back
              if (elect_one()) {
                if (msub > 1)
                  halo_issue_row<4, true>(d_tmem0, (uint32_t)acc_stride, a_lo, (uint32_t)x.sub_off, a_hi, b_lo,
                                          (uint32_t)(tap_bytes >> 4), idesc, accf);
                else
                  halo_issue_row<4, false>(d_tmem0, (uint32_t)acc_stride, a_lo, (uint32_t)x.sub_off, a_hi, b_lo,
                                           (uint32_t)(tap_bytes >> 4), idesc, accf);
              }
              accf = 1;
              a_lo += (10 * 128) >> 4;  // next kernel row
            } else if (b_group == 1 && ksteps == 4 && msub == 1) {  // one tap per weight box (wide N): four MMAs
              if (elect_one()) umma_bf16_ksteps<4>(d_tmem0, a_lo, a_hi, b_lo, umma_desc_hi(1024), idesc, accf);
              accf = 1;
              a_lo += 128 >> 4;
              if (++kw == 3) {
                kw = 0;
                a_lo += (7 * 128) >> 4;
              }
            } else
            for (int u = 0; u < b_group; ++u) {
              if (elect_one()) {
                umma_bf16_ksteps_n(ksteps, d_tmem0, a_lo, a_hi, b_lo, umma_desc_hi(1024), idesc, accf);
                if (msub > 1)
                  umma_bf16_ksteps_n(ksteps, d_tmem0 + (uint32_t)acc_stride, a_lo + (uint32_t)x.sub_off, a_hi, b_lo,
                                     umma_desc_hi(1024), idesc, accf);
              }
              accf = 1;
              b_lo += (uint32_t)(tap_bytes >> 4);
              a_lo += 128 >> 4;
              if (++kw == 3) {
                kw = 0;
                a_lo += (7 * 128) >> 4;
              }
            }
            if (!b_stat && elect_one()) umma_commit(b_empty + sb);
          }
          }
          const long long _ti1 = prof ? clock64() : 0;
          if (elect_one()) umma_commit(a_empty + sa);
          if (prof) { pw4 += _ti1 - _ti0; pw3 += clock64() - _ti1; }
        }
        {
          PROF_T0();
          if (elect_one()) umma_commit(tfull_bar + buf);
          PROF_ADD(pw3);
        }
      }
      if (prof && elect_one()) {
        atomicAdd(&g_conv_prof[14], (unsigned long long)pw3);
        atomicAdd(&g_conv_prof[15], (unsigned long long)pw4);
        atomicAdd(&g_conv_prof[1], (unsigned long long)pw0);
        atomicAdd(&g_conv_prof[2], (unsigned long long)pw1);
        atomicAdd(&g_conv_prof[6], (unsigned long long)pw2);
      }
    }
  } else {
    // ===================== epilogue =====================
    const int lg = warp & 3;
    int sidx, c_begin, c_end;  // this warp's sub-tile and 16-column chunk range
    epi_split(x.msub, p.n_tile >> 4, (warp - kEpiWarp0) >> 2, &sidx, &c_begin, &c_end);
    uint8_t* stage = stage_base + (warp - kEpiWarp0) * kStageB;
    const int r = lg * 32 + lane;
    long long pacc[6] = {0, 0, 0, 0, 0, 0};
    int acc = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acc) {
      const int mt = fdiv(tile, p.fd_ns), n0 = (tile - mt * n_splits) * p.n_tile;
      const int b = fdiv(mt, p.fd_tpi), t_in = mt - b * tiles_per_img;
      const int th = fdiv(t_in, p.fd_tw);
      const int h0 = th * 16 * x.msub, w0 = (t_in - th * p.tiles_w) * 8;
      const int buf = acc & 1;
      if ((MODE & 3) == EPI_BF16_RES && tile + (int)gridDim.x < total_tiles && c_begin < c_end) {
        const int tile2 = tile + gridDim.x;
        const int mt2 = fdiv(tile2, p.fd_ns), n2 = (tile2 - mt2 * n_splits) * p.n_tile;
        const int b2 = fdiv(mt2, p.fd_tpi), t2 = mt2 - b2 * tiles_per_img;
        const int th2 = fdiv(t2, p.fd_tw);
        const int h = th2 * 16 * x.msub + sidx * 16 + (r >> 3), w = (t2 - th2 * p.tiles_w) * 8 + (r & 7);
        epi_prefetch_res(p, (h < p.tH) && (w < p.tW), b2, h * p.tW + w, n2 + c_begin * 16, (c_end - c_begin) * 16);
      }
      {
        PROF_T0();
        mbar_wait_bo(tfull_bar + buf, (acc >> 1) & 1, 4u, p.bo_epi);
        PROF_ADD(pw0);
      }
      tc_fence_after();
      PROF_T0();
      {
        const int h = h0 + sidx * 16 + (r >> 3), w = w0 + (r & 7);
        const bool valid = (h < p.tH) && (w < p.tW);
        const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((buf * x.msub + sidx) * acc_stride);
        epi_drain<MODE>(p, smem_u32(stage), smem_u32(sbias), lane, c_begin, c_end, t_addr, n0, valid, b,
                        valid ? h * p.tW + w : 0, tempty_bar + buf, pacc);
      }
      PROF_ADD(pw1);
    }
    if (prof && warp == kEpiWarp0 && lane == 0) {
      atomicAdd(&g_conv_prof[3], (unsigned long long)pw0);
      atomicAdd(&g_conv_prof[4], (unsigned long long)pw1);
    }
    if (prof && lane == 0) {
#pragma unroll
      for (int i = 0; i < 6; ++i) atomicAdd(&g_conv_prof[8 + i], (unsigned long long)pacc[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (prof && threadIdx.x == 0) {
    atomicAdd(&g_conv_prof[5], (unsigned long long)(clock64() - prof_start));
    atomicAdd(&g_conv_prof[7], 1ull);
  }
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
constexpr int kStemMaxC0 = 128;
constexpr int kStemRowWords = 27;  // 33 pixels x 3 B = 99 B plus up to 3 B of misalignment -> 26 words, +1 spare

__global__ void __launch_bounds__(128, 5)
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
  __shared__ __align__(16) float sbias[kStemMaxC0];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int C0 = p.Cout;
  epi_load_bias(p, sbias);
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
  const int t_begin = blockIdx.x * tiles_per_cta, t_end = min(t_begin + tiles_per_cta, total_tiles);
  // input patch of a tile: 17 rows; row r holds the aligned words covering bytes [g0, g0+99) of the frame row, g0 =
  // byte offset of pixel (ih0+r, iw0); out-of-frame pixels are zeroed when the rows are converted below.
  // The patch of tile t+1 is fetched into registers while tile t is converted, multiplied and stored.
  constexpr int kPatchPerThread = (17 * kStemRowWords + 127) / 128;
  uint32_t pre[kPatchPerThread];
  auto fetch_patch = [&](int tile) {
    const int b = fdiv(tile, p.fd_tpi), t_in = tile - b * tiles_per_img;
    const int th = fdiv(t_in, p.fd_tw);
    const int ih0 = 2 * (th * kStemTH) - 1, iw0 = 2 * ((t_in - th * tiles_w) * kStemTW) - 1;
#pragma unroll
    for (int j = 0; j < kPatchPerThread; ++j) {
      const int i = tid + 128 * j;
      const int r = i / kStemRowWords, wd = i - r * kStemRowWords;
      const int ih = ih0 + r;
      uint32_t v = 0;
      if (i < 17 * kStemRowWords && ih >= 0 && ih < H) {
        const long long g0 = (long long)b * frame_bytes + ((long long)ih * W + iw0) * 3;
        const long long wi = (g0 >> 2) + wd;  // arithmetic shift: floor, also for the (only) negative case g0 = -3
        if (wi >= 0 && wi < all_words) v = __ldg(words + wi);
      }
      pre[j] = v;
    }
  };
  if (t_begin < t_end) fetch_patch(t_begin);
  for (int tile = t_begin; tile < t_end; ++tile) {
    const int b = fdiv(tile, p.fd_tpi), t_in = tile - b * tiles_per_img;
    const int th = fdiv(t_in, p.fd_tw);
    const int oh0 = th * kStemTH, ow0 = (t_in - th * tiles_w) * kStemTW;
    const int ih0 = 2 * oh0 - 1, iw0 = 2 * ow0 - 1;
#pragma unroll
    for (int j = 0; j < kPatchPerThread; ++j)
      if (tid + 128 * j < 17 * kStemRowWords) sIn[tid + 128 * j] = pre[j];
    __syncthreads();
    if (tile + 1 < t_end) fetch_patch(tile + 1);
    {  // im2col row of output pixel (ty, tx) = tid: k = (kh*3+kw)*3 + c_rgb, frame bytes are BGR
      const int ty = tid >> 4, tx = tid & 15;
      // uint8 -> bf16 without the conversion pipe (I2F / F2F run on the 16-lane XU pipe next to the epilogue's tanh and
      // bounded this kernel): 0x4B000000 | b is the float 2^23 + b, subtracting 2^23 leaves b exactly, and since
      // b < 256 fits bf16's 8 significant bits the upper half of that float IS the bf16 value.
      // The 9 bytes of a kernel row (3 pixels x BGR) come out of three aligned shared-memory words: two funnel
      // shifts line them up, one PRMT per byte drops it into the float pattern (explicit shared-space loads; the
      // byte-pointer version compiled to 27 generic loads with 64-bit address arithmetic each).
      uint32_t fb[32];
      // the only out-of-frame COLUMN is iw = -1 (kw = 0 of the first pixel of a left-edge tile): W is even, so the
      // right edge never overhangs; out-of-frame ROWS were zero-filled when the patch was fetched
      const bool left_pad = iw0 < 0 && tx == 0;
      const uint32_t sIn_sa = smem_u32(sIn);
      // low two bits of the byte offset of pixel (ih, iw0) in the frame buffer (what fetch_patch aligned the row on)
      const uint32_t g_lo = (uint32_t)b * (uint32_t)(frame_bytes & 3) + (uint32_t)iw0 * 3u;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int r = 2 * ty + kh;
        const uint32_t shift = (g_lo + (uint32_t)(ih0 + r) * (uint32_t)((W * 3) & 3)) & 3u;
        const uint32_t off = shift + 6u * (uint32_t)tx;
        const uint32_t wa = sIn_sa + (uint32_t)(r * kStemRowWords) * 4u + (off & ~3u);
        uint32_t w0, w1, w2;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w0) : "r"(wa));
        asm volatile("ld.shared.b32 %0, [%1+4];" : "=r"(w1) : "r"(wa));
        asm volatile("ld.shared.b32 %0, [%1+8];" : "=r"(w2) : "r"(wa));
        const uint32_t sh = (off & 3u) * 8u;
        uint32_t a0 = __funnelshift_r(w0, w1, sh), a1 = __funnelshift_r(w1, w2, sh);
        const uint32_t a2 = w2 >> sh;
        if (left_pad) a0 &= 0xFF000000u;  // pixel iw = -1: its three bytes read as zero
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {  // frame bytes are BGR, k runs over RGB
            const int i = kw * 3 + (2 - c);  // byte of the 9-byte run
            const uint32_t src = i < 4 ? a0 : (i < 8 ? a1 : a2);
            const uint32_t bits = __byte_perm(src, 0x4B000000u, 0x7440 | (i & 3));
            fb[(kh * 3 + kw) * 3 + c] = __float_as_uint(__uint_as_float(bits) - 8388608.0f);
          }
        }
      }
#pragma unroll
      for (int k = 27; k < 32; ++k) fb[k] = 0u;
      uint32_t row[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) row[k] = __byte_perm(fb[2 * k], fb[2 * k + 1], 0x7632);  // (hi16(a), hi16(b))
      const uint32_t sA_row = smem_u32(sA) + (uint32_t)tid * 128u;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts128(sA_row + (uint32_t)((j ^ (tid & 7)) << 4), row[4 * (j & 3)], row[4 * (j & 3) + 1], row[4 * (j & 3) + 2],
               row[4 * (j & 3) + 3]);
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
    mbar_wait_blocking(&bar, phase, 16u);  // several CTAs share the SM: a parked warp leaves its issue slots to them
    phase ^= 1;
    tc_fence_after();
    {
      const int r = warp * 32 + lane;
      const int oh = oh0 + (r >> 4), ow = ow0 + (r & 15);
      const bool valid = oh < oH && ow < oW;
      long long pacc[6] = {0, 0, 0, 0, 0, 0};
      epi_drain<EPI_BF16 | EPI_ACT, false>(p, smem_u32(sStage + warp * (alias_stage ? 4096 : kEpiStageBytes)), smem_u32(sbias), lane, 0, C0 >> 4,
                tmem_base + ((uint32_t)(warp * 32) << 16), 0, valid, b, valid ? oh * oW + ow : 0, nullptr, pacc);
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

#if YPB_DIAG
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
#endif  // YPB_DIAG

}  // namespace ypb
