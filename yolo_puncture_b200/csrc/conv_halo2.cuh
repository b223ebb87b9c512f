// 3x3 stride-1 convs with streamed weights on CTA PAIRS: tcgen05.mma.cta_group::2, M = 256.
//
// What bounds conv3_halo_kernel on the wide layers (Cin = Cout >= 128: every C2f bottleneck of the m / l / x models, the
// head and proto 3x3s of all of them) is not the tensor pipe but the weight stream: all nine taps of all chunks do not
// fit in shared memory, so every 128-pixel tile pulls the layer's whole weight tensor (552 KB for 160 -> 160) through
// L2 -> SM again: measured 75 GB/s per SM = 11 TB/s chip-wide on yolov8x-seg's model.4.m, 2x the time the MMAs need.
// Here two CTAs of a cluster (the two SMs of a TPC) share every weight tile: each CTA fetches HALF of its rows
// (n_tile/2 x 64 channels per tap) and the pair's single MMA issuer (the leader's elected lane) multiplies M = 256
// pixels - the leader's 128-pixel halo tile and the peer's - against the B operand assembled from both shared
// memories.  Weight bytes per SM halve, the weight ring gets twice as deep in the same shared memory, shared-memory
// reads of B per MMA halve, and each CTA keeps its own 128 x n_tile fp32 accumulator (double-buffered) in its own TMEM.
//
// Protocol (ring slots and phases advance in lock-step in both CTAs; "leader" = cluster rank 0):
//   a_full / b_full   live in the LEADER: both CTAs' TMA loads complete_tx there (cp.async.bulk.tensor.cta_group::2
//                     with the leader's barrier address); the leader's producer posts expect_tx for both halves
//   a_empty / b_empty one per CTA: tcgen05.commit.cta_group::2 ... multicast::cluster frees the slot in both CTAs
//   tfull             one per CTA, multicast commit: both epilogues start when the pair's accumulators are complete
//   tempty            in the LEADER, 2 x kEpiWarps arrivals: the peer's epilogue warps arrive through shared::cluster
// Everything else (halo-descriptor trick, epilogue, tile order) is conv3_halo_kernel's (conv_tc.cuh).
#pragma once
#include "conv_tc.cuh"

namespace ypb {

struct Conv3Pair {
  int a_slots;     // halo ring depth (per CTA)
  int a_bytes;     // bytes per halo slot (1024 multiple): box {64 ch, 10, 18}
  int b_slots;     // weight ring depth, in kernel-row groups of three taps
  int grp_bytes;   // per CTA: 3 taps x (n_tile / 2) rows x 128 B
  int tap_off[9];  // A descriptor offsets of the nine taps inside the halo box (16-byte units)
  uint32_t a_hi;   // high descriptor word of A (SBO = 1280)
  int m_tiles;     // 16 x 8 output tiles over the whole batch
};

template <int MODE>
__global__ void __launch_bounds__(kConv2Threads, 1)
conv3_halo2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ ConvParams p, const __grid_constant__ Conv3Pair x, int n_splits, int total_pairs) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + x.a_slots * x.a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + x.b_slots * x.grp_bytes);
  uint64_t* a_full = bars;            // [4]   (used in the leader)
  uint64_t* a_empty = bars + 4;       // [4]
  uint64_t* b_full = bars + 8;        // [12]  (used in the leader)
  uint64_t* b_empty = bars + 20;      // [12]
  uint64_t* tfull_bar = bars + 32;    // [2]
  uint64_t* tempty_bar = bars + 34;   // [2]   (used in the leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 37);
  uint8_t* stage_base = reinterpret_cast<uint8_t*>(bars) + 512;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int kchunks = (p.Cin + 63) >> 6;
  const int acc_stride = conv2_acc_stride(p.n_tile);
  const int half_n = p.n_tile >> 1;
  const int tap_bytes = half_n * 128;  // one tap's share of B in THIS CTA
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * acc_stride)) tmem_cols <<= 1;

  if (warp == kProdWarp0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < 4; ++i) { mbar_init(a_full + i, 1); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < 12; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar + i, 1); mbar_init(tempty_bar + i, 2 * kEpiWarps); }
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc2(tmem_slot, tmem_cols);
  constexpr int kStageB = epi_stage_bytes((MODE & 3) == EPI_F32);
  float* sbias = reinterpret_cast<float*>(stage_base + kEpiWarps * kStageB);
  epi_load_bias(p, sbias);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them from this side
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  const int pw = warp - kProdWarp0;
  if (pw == 0) {
    // ===================== producer 0: this CTA's halo tiles; bytes are counted on the leader's a_full =====================
    if (elect_one()) {
      int sa = -1;
      uint32_t pa = 1;
      for (int pt = cid; pt < total_pairs; pt += nclusters) {
        const int mp = fdiv(pt, p.fd_ns);
        const int mt = 2 * mp + (int)rank;
        const int b = fdiv(mt, p.fd_tpi), t_in = mt - b * tiles_per_img;  // mt >= m_tiles: b >= batch -> the box is all zero fill
        const int th = fdiv(t_in, p.fd_tw);
        const int h0 = th * 16, w0 = (t_in - th * p.tiles_w) * 8;
        for (int c = 0; c < kchunks; ++c) {
          if (++sa == x.a_slots) sa = 0;
          if (sa == 0) pa ^= 1;
          mbar_wait_bo(a_empty + sa, pa ^ 1, 1u, p.bo_prod);
          if (leader) mbar_expect_tx(a_full + sa, (uint32_t)(2 * 180 * 128));
          tma_load_5d_2cta(sA + sa * x.a_bytes, &tmA, mapa_u32(smem_u32(a_full + sa), 0), p.a_base[0] + c * 64, w0 - 1, h0 - 1, b, 0);
        }
      }
    }
  } else if (pw > 0 && pw < kProdWarps) {
    // ===================== producers 1..2: this CTA's half of every weight box =====================
    if (elect_one()) {
      int sb = -1;
      uint32_t pb = 1;
      for (int pt = cid; pt < total_pairs; pt += nclusters) {
        const int mp = fdiv(pt, p.fd_ns);
        const int n0 = (pt - mp * n_splits) * p.n_tile + (int)rank * half_n;
        for (int c = 0; c < kchunks; ++c) {
          for (int g = 0; g < 3; ++g) {
            if (++sb == x.b_slots) sb = 0;
            if (sb == 0) pb ^= 1;
            if (1 + (sb % (kProdWarps - 1)) != pw) continue;
            mbar_wait_bo(b_empty + sb, pb ^ 1, 1u, p.bo_prod);
            if (leader) mbar_expect_tx(b_full + sb, (uint32_t)(2 * x.grp_bytes));
            tma_load_3d_2cta(sB + sb * x.grp_bytes, &tmB, mapa_u32(smem_u32(b_full + sb), 0), c * 64, n0, g * 3);
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== the pair's MMA issuer: leader only (whole warp converged, one elected lane issues) =====================
    if (leader) {
      const uint32_t idesc = umma_idesc_bf16(256, p.n_tile);
      const uint32_t tu = (uint32_t)(tap_bytes >> 4);
      int sa = -1, sb = -1, acc = 0;
      uint32_t pa = 1, pb = 1;
      for (int pt = cid; pt < total_pairs; pt += nclusters, ++acc) {
        const int buf = acc & 1;
        mbar_wait_warp(tempty_bar + buf, ((acc >> 1) & 1) ^ 1, 8u, p.bo_mma_acc);
        tc_fence_after();
        uint32_t accf = 0;
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * acc_stride);
        for (int c = 0; c < kchunks; ++c) {
          if (++sa == x.a_slots) sa = 0;
          if (sa == 0) pa ^= 1;
          mbar_wait_warp(a_full + sa, pa, 2u, p.bo_mma_full);
          tc_fence_after();
          int ksteps = (p.Cin - c * 64) >> 4;
          if (ksteps > 4) ksteps = 4;
          const uint32_t a_lo0 = umma_desc_lo(smem_u32(sA + sa * x.a_bytes));
          for (int g = 0; g < 3; ++g) {
            if (++sb == x.b_slots) sb = 0;
            if (sb == 0) pb ^= 1;
            mbar_wait_warp(b_full + sb, pb, 2u, p.bo_mma_full);
            tc_fence_after();
            const uint32_t b_lo = umma_desc_lo(smem_u32(sB + sb * x.grp_bytes));
            if (elect_one()) {
#pragma unroll
              for (int u = 0; u < 3; ++u) {
                const uint32_t a = a_lo0 + (uint32_t)x.tap_off[g * 3 + u], bq = b_lo + (uint32_t)u * tu;
                const uint32_t af = u == 0 ? accf : 1u;
                if (ksteps == 4) umma2_bf16_ksteps<4>(d_tmem, a, x.a_hi, bq, umma_desc_hi(1024), idesc, af);
                else if (ksteps == 2) umma2_bf16_ksteps<2>(d_tmem, a, x.a_hi, bq, umma_desc_hi(1024), idesc, af);
                else {
#pragma unroll
                  for (int j = 0; j < 3; ++j)
                    if (j < ksteps) umma2_bf16_lohi(d_tmem, a + 2 * j, x.a_hi, bq + 2 * j, umma_desc_hi(1024), idesc, j == 0 ? af : 1u);
                }
              }
              umma2_commit_mc(b_empty + sb, 3);
            }
            accf = 1;
          }
          if (elect_one()) umma2_commit_mc(a_empty + sa, 3);
        }
        if (elect_one()) umma2_commit_mc(tfull_bar + buf, 3);
      }
    }
  } else {
    // ===================== epilogue: this CTA's 128 rows x n_tile columns =====================
    const int lg = warp & 3;
    int sidx, c_begin, c_end;
    epi_split(1, p.n_tile >> 4, (warp - kEpiWarp0) >> 2, &sidx, &c_begin, &c_end);
    uint8_t* stage = stage_base + (warp - kEpiWarp0) * kStageB;
    const int r = lg * 32 + lane;
    long long pacc[6] = {0, 0, 0, 0, 0, 0};
    int acc = 0;
    for (int pt = cid; pt < total_pairs; pt += nclusters, ++acc) {
      const int mp = fdiv(pt, p.fd_ns), n0 = (pt - mp * n_splits) * p.n_tile;
      const int mt = 2 * mp + (int)rank;
      const int b = fdiv(mt, p.fd_tpi), t_in = mt - b * tiles_per_img;
      const int th = fdiv(t_in, p.fd_tw);
      const int h0 = th * 16, w0 = (t_in - th * p.tiles_w) * 8;
      const int buf = acc & 1;
      if ((MODE & 3) == EPI_BF16_RES && pt + nclusters < total_pairs && c_begin < c_end) {
        const int pt2 = pt + nclusters;
        const int mp2 = fdiv(pt2, p.fd_ns), n2 = (pt2 - mp2 * n_splits) * p.n_tile;
        const int mt2 = 2 * mp2 + (int)rank;
        const int b2 = fdiv(mt2, p.fd_tpi), t2 = mt2 - b2 * tiles_per_img;
        const int th2 = fdiv(t2, p.fd_tw);
        const int h = th2 * 16 + (r >> 3), w = (t2 - th2 * p.tiles_w) * 8 + (r & 7);
        epi_prefetch_res(p, (b2 < p.tB) && (h < p.tH) && (w < p.tW), b2, h * p.tW + w, n2 + c_begin * 16, (c_end - c_begin) * 16);
      }
      mbar_wait_bo(tfull_bar + buf, (acc >> 1) & 1, 4u, p.bo_epi);
      tc_fence_after();
      const int h = h0 + (r >> 3), w = w0 + (r & 7);
      const bool valid = (b < p.tB) && (h < p.tH) && (w < p.tW);
      const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * acc_stride);
      epi_drain<MODE>(p, smem_u32(stage), smem_u32(sbias), lane, c_begin, c_end, t_addr, n0, valid, valid ? b : 0,
                      valid ? h * p.tW + w : 0, tempty_bar + buf, pacc, nullptr, 0, 0, 0, nullptr,
                      leader ? 0u : mapa_u32(smem_u32(tempty_bar + buf), 0));
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA frees its TMEM / leaves while the other may still be signalling it
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// The same pairing for the per-tap kernel (1x1 convs on 128 * msub consecutive pixels, stride-2 3x3 convs on rectangles):
// conv_tc2_kernel's ring of {A box | W box} stages, each CTA loading ITS tile's A box and half of the W rows, the
// leader issuing M = 256 MMAs per sub-tile.  Besides halving the weight bytes every SM pulls through L2 -> SM, a CTA
// now reads only half of B from shared memory per MMA: an N = 256 step (137 cycles on one SM, bound by operand reads)
// gets cheaper.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kConv2Threads, 1)
conv_tc2p_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ ConvParams p, int n_splits, int total_pairs) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = p.msub * kATileBytes;
  const int half_n = p.n_tile >> 1;
  const int stage_bytes = a_bytes + half_n * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);  // used in the leader
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tfull_bar = empty_bar + p.stages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2] used in the leader
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int kchunks = (p.Cin + 63) >> 6;
  const int acc_stride = conv2_acc_stride(p.n_tile);
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * p.msub * acc_stride)) tmem_cols <<= 1;

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
      mbar_init(tempty_bar + i, 2 * kEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc2(tmem_slot, tmem_cols);
  const int kStageB = epi_stage_bytes((MODE & 3) == EPI_F32, p.ntaps == 1);
  uint8_t* stage0 = smem + ((p.stages * stage_bytes + 256 + 1023) & ~1023);
  float* sbias = reinterpret_cast<float*>(stage0 + kEpiWarps * kStageB);
  epi_load_bias(p, sbias);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  const int pw = warp - kProdWarp0;
  if (pw >= 0 && pw < kProdWarps) {
    // ===================== TMA producers (stage s belongs to producer s % kProdWarps, in both CTAs alike) =====================
    if (elect_one()) {
      const uint32_t tx_bytes = 2u * (uint32_t)(p.TH * p.TW * 128 + half_n * 128);  // both CTAs' boxes land on the leader's barrier
      int s = -1;
      uint32_t ph = 1;
      for (int pt = cid; pt < total_pairs; pt += nclusters) {
        const int mp = fdiv(pt, p.fd_ns), n0 = (pt - mp * n_splits) * p.n_tile + (int)rank * half_n;
        const int mt = 2 * mp + (int)rank;
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
            if ((s % kProdWarps) != pw) continue;
            mbar_wait_bo(empty_bar + s, ph ^ 1, 1u, p.bo_prod);
            uint8_t* sa = smem + s * stage_bytes;
            const uint32_t fb = mapa_u32(smem_u32(full_bar + s), 0);
            if (leader) mbar_expect_tx(full_bar + s, tx_bytes);
            tma_load_5d_2cta(sa, &tmA, fb, cbase[0] + p.tap[t][0] + c * 64, cbase[1] + p.tap[t][1], cbase[2] + p.tap[t][2],
                             cbase[3] + p.tap[t][3], cbase[4] + p.tap[t][4]);
            tma_load_3d_2cta(sa + a_bytes, &tmB, fb, c * 64, n0, t);
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== the pair's MMA issuer (leader only) =====================
    if (leader) {
      const uint32_t idesc = umma_idesc_bf16(256, p.n_tile);
      const int msub = p.msub;
      const uint32_t sub_bytes = (uint32_t)(p.sub_rows * 128);
      int s = -1, acc = 0;
      uint32_t ph = 1;
      for (int pt = cid; pt < total_pairs; pt += nclusters, ++acc) {
        const int buf = acc & 1;
        mbar_wait_warp(tempty_bar + buf, ((acc >> 1) & 1) ^ 1, 8u, p.bo_mma_acc);
        tc_fence_after();
        uint32_t accf = 0;
        const uint32_t d_tmem0 = tmem_base + (uint32_t)(buf * msub * acc_stride);
        for (int c = 0; c < kchunks; ++c) {
          int ksteps = (p.Cin - c * 64) >> 4;
          if (ksteps > 4) ksteps = 4;
          for (int t = 0; t < p.ntaps; ++t) {
            if (++s == p.stages) s = 0;
            if (s == 0) ph ^= 1;
            mbar_wait_warp(full_bar + s, ph, 2u, p.bo_mma_full);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + s * stage_bytes);
            const uint32_t b_lo = umma_desc_lo(sa + a_bytes);
            const uint32_t a_lo0 = umma_desc_lo(sa), a_lo1 = umma_desc_lo(sa + sub_bytes);
            constexpr uint32_t hi = umma_desc_hi(1024);
            if (elect_one()) {
              if (ksteps == 4) {
                umma2_bf16_ksteps<4>(d_tmem0, a_lo0, hi, b_lo, hi, idesc, accf);
                if (msub > 1) umma2_bf16_ksteps<4>(d_tmem0 + (uint32_t)acc_stride, a_lo1, hi, b_lo, hi, idesc, accf);
              } else if (ksteps == 2) {
                umma2_bf16_ksteps<2>(d_tmem0, a_lo0, hi, b_lo, hi, idesc, accf);
                if (msub > 1) umma2_bf16_ksteps<2>(d_tmem0 + (uint32_t)acc_stride, a_lo1, hi, b_lo, hi, idesc, accf);
              } else {
#pragma unroll
                for (int j = 0; j < 3; ++j)
                  if (j < ksteps) {
                    umma2_bf16_lohi(d_tmem0, a_lo0 + 2 * j, hi, b_lo + 2 * j, hi, idesc, j == 0 ? accf : 1u);
                    if (msub > 1) umma2_bf16_lohi(d_tmem0 + (uint32_t)acc_stride, a_lo1 + 2 * j, hi, b_lo + 2 * j, hi, idesc, j == 0 ? accf : 1u);
                  }
              }
              umma2_commit_mc(empty_bar + s, 3);
            }
            accf = 1;
          }
        }
        if (elect_one()) umma2_commit_mc(tfull_bar + buf, 3);
      }
    }
  } else {
    // ===================== epilogue: this CTA's msub x 128 rows =====================
    const int lg = warp & 3;
    int sidx, c_begin, c_end;
    epi_split(p.msub, p.n_tile >> 4, (warp - kEpiWarp0) >> 2, &sidx, &c_begin, &c_end);
    uint8_t* stage = stage0 + (warp - kEpiWarp0) * kStageB;
    int tbuf = 0;
    const int r = lg * 32 + lane;
    long long pacc[6] = {0, 0, 0, 0, 0, 0};
    const bool flat = p.ntaps == 1;
    const int R = sidx * p.sub_rows + r;
    const int rh = R / p.TW, rw = R - rh * p.TW;
    int acc = 0;
    for (int pt = cid; pt < total_pairs; pt += nclusters, ++acc) {
      const int mp = fdiv(pt, p.fd_ns), n0 = (pt - mp * n_splits) * p.n_tile;
      const int mt = 2 * mp + (int)rank;
      const int b = fdiv(mt, p.fd_tpi), t_in = mt - b * tiles_per_img;
      const int th = fdiv(t_in, p.fd_tw);
      const int buf = acc & 1;
      if ((MODE & 3) == EPI_BF16_RES && pt + nclusters < total_pairs && c_begin < c_end) {
        const int pt2 = pt + nclusters;
        const int mp2 = fdiv(pt2, p.fd_ns), n2 = (pt2 - mp2 * n_splits) * p.n_tile;
        const int mt2 = 2 * mp2 + (int)rank;
        const int b2 = fdiv(mt2, p.fd_tpi), t2 = mt2 - b2 * tiles_per_img;
        const int th2 = fdiv(t2, p.fd_tw);
        const int h = th2 * p.TH + rh, w = (t2 - th2 * p.tiles_w) * p.TW + rw;
        const bool valid = (r < p.sub_rows) && (b2 < p.tB) && (h < p.tH) && (w < p.tW);
        int qb = b2, rem = h * p.tW + w;
        if (flat) {
          qb = fdiv(w, p.fd_hw);
          rem = w - qb * p.img_HW;
        }
        epi_prefetch_res(p, valid, qb, rem, n2 + c_begin * 16, (c_end - c_begin) * 16);
      }
      mbar_wait_bo(tfull_bar + buf, (acc >> 1) & 1, 4u, p.bo_epi);
      tc_fence_after();
      {
        const int h = th * p.TH + rh, w = (t_in - th * p.tiles_w) * p.TW + rw;
        const bool valid = (r < p.sub_rows) && (b < p.tB) && (h < p.tH) && (w < p.tW);
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
        if (p.tma_out != 0) {
          tq = t_in * p.TW + sidx * p.sub_rows + lg * 32;
          if (p.tma_out == 3) {
            tb = fdiv(tq, p.fd_hw);
            tq -= tb * p.img_HW;
          }
        }
        epi_drain<MODE>(p, smem_u32(stage), smem_u32(sbias), lane, c_begin, c_end, t_addr, n0, valid, qb, rem, tempty_bar + buf,
                        pacc, &tmO, p.tma_out, tq, tb, &tbuf, leader ? 0u : mapa_u32(smem_u32(tempty_bar + buf), 0));
      }
    }
    if (p.tma_out != 0 && lane == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, tmem_cols);
  }
}

}  // namespace ypb
