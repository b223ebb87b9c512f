// TMA load-path microbenchmark (diagnostics only; drives design decisions for conv_tc.cuh).
// One persistent CTA per SM: a producer thread streams 128-row x 128-byte SWIZZLE_128B boxes into a ring of
// shared-memory stages, a consumer thread frees each stage as soon as it lands.  No math: the number that comes
// out is the ceiling of the operand-fetch path for a given access pattern.
#pragma once
#include "common.cuh"

namespace ypb {

struct TmaBenchParams {
  int mode;        // 0: distinct rows per CTA (streaming)   1: every CTA reads the same box (weights pattern)
                   // 2: 3x3-tap pattern on a (C=64, W, H, B) map: 9 shifted 8x16 boxes per tile
  int stages;      // ring depth (16 KB per stage)
  int iters;       // box loads per CTA
  int rows_total;  // rows of the 2-D view (mode 0/1)
  int W, H, B;     // map (mode 2)
};

__global__ void __launch_bounds__(256, 1)
tma_bench_kernel(const __grid_constant__ CUtensorMap tm2d, const __grid_constant__ CUtensorMap tm5d,
                 const TmaBenchParams p) {
  // mode 0/1/2 as documented; p.W (mode 0/1) = number of producer/consumer thread pairs (1..4), each pair owns
  // the stages s with s % pairs == pair; p.H (mode 0/1) = rows per box (128 or 256; stage = rows*128 bytes).
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // mode 4: halo boxes {64 ch, 10, R} of a (C_tot, W, H, B) map, one per tile of 8 x (R-2) pixels, tiles strided
  // over the CTAs like conv3_halo_kernel does; rows_total = C_tot | R << 16 | producers << 24.
  const int halo_R = (p.rows_total >> 16) & 0xff;
  const int box_rows = p.mode == 4 ? halo_R * 10 : (p.mode == 2 || p.H == 0) ? 128 : p.H;
  const int tx_bytes = box_rows * 128;
  const int stage_bytes = (tx_bytes + 1023) & ~1023;
  const int pairs = p.mode == 4 ? ((p.rows_total >> 24) & 0xf) : (p.mode == 2 || p.W == 0) ? 1 : p.W;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    mbar_fence_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  if (!elect_one()) return;
  if (warp < pairs) {  // producers
    const int tiles_w = p.mode == 2 ? p.W / 16 : 1, tiles_h = p.mode == 2 ? p.H / 8 : 1;
    for (int it = warp; it < p.iters; it += pairs) {
      const int s = it % p.stages;
      mbar_wait(empty_bar + s, ((it / p.stages) & 1) ^ 1, 0x1000u);
      const int variant = (p.mode == 2 || p.mode == 4) ? 0 : p.B;  // 0: expect_tx then copy, 1: relaxed expect_tx, 2: copy then expect_tx
      if (variant == 0) mbar_expect_tx(full_bar + s, tx_bytes);
      if (variant == 1) mbar_expect_tx_relaxed(full_bar + s, tx_bytes);
      if (p.mode == 4) {
        const int tw = p.W / 8, th = (p.H + halo_R - 3) / (halo_R - 2);
        const int tile = (blockIdx.x + it * gridDim.x) % (tw * th * p.B);
        const int b = tile / (tw * th), r = tile % (tw * th);
        tma_load_5d(smem + s * stage_bytes, &tm5d, full_bar + s, 0, (r % tw) * 8 - 1, (r / tw) * (halo_R - 2) - 1, b, 0);
      } else if (p.mode == 2) {
        const int tile = (blockIdx.x + (it / 9) * gridDim.x) % (tiles_w * tiles_h * p.B);
        const int t = it % 9;
        const int b = tile / (tiles_w * tiles_h), r = tile % (tiles_w * tiles_h);
        tma_load_5d(smem + s * stage_bytes, &tm5d, full_bar + s, 0, (r % tiles_w) * 16 + t % 3 - 1, (r / tiles_w) * 8 + t / 3 - 1, b, 0);
      } else if (p.mode == 3) {  // rank-2 tensor map, 2d instruction
        const long long row = ((long long)(blockIdx.x + (long long)it * gridDim.x) * box_rows) % p.rows_total;
        tma_load_2d(smem + s * stage_bytes, &tm5d, full_bar + s, 0, (int)row);
      } else {
        const long long row = p.mode == 1 ? 0 : ((long long)(blockIdx.x + (long long)it * gridDim.x) * box_rows) % p.rows_total;
        tma_load_5d(smem + s * stage_bytes, &tm2d, full_bar + s, 0, (int)row, 0, 0, 0);
      }
      if (variant == 2) mbar_expect_tx(full_bar + s, tx_bytes);
    }
  } else if (warp >= 4 && warp < 4 + pairs) {  // consumers
    for (int it = warp - 4; it < p.iters; it += pairs) {
      const int s = it % p.stages;
      mbar_wait(full_bar + s, (it / p.stages) & 1, 0x2000u);
      mbar_arrive(empty_bar + s);
    }
  }
}

// MMA-side microbenchmark: one CTA per SM, operands resident in shared memory, one thread issues `iters`
// tcgen05.mma (M=128, N=n, K=16 each, SWIZZLE_128B K-major, A optionally through the halo-style shifted
// descriptor).  Optionally a second warp streams TMA boxes into spare stages at the same time (tma_iters > 0).
struct MmaBenchParams {
  int n;          // MMA N
  int iters;      // MMAs per CTA
  int shifted;    // 1: A descriptor start advanced by 11 rows, SBO 1280 (halo style)
  int tma_iters;  // concurrent 16 KB TMA box loads per CTA (0 = none)
  int rows_total;
};

__global__ void __launch_bounds__(96, 1)
mma_bench_kernel(const __grid_constant__ CUtensorMap tm2d, const MmaBenchParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;               // 32 KB (room for the shifted window)
  uint8_t* sB = smem + 32768;       // 256 rows x 128 B
  uint8_t* sT = smem + 65536;       // 4 TMA stages x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536 + 65536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 9; ++i) mbar_init(bars + i, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, p.n);
    const uint32_t a0 = smem_u32(sA) + (p.shifted ? 11 * 128 : 0), b0 = smem_u32(sB);
    for (int it = 0; it < p.iters; ++it) {
      const int j = it & 3;
      const uint64_t ad = p.shifted ? umma_desc_sw128_sbo(a0 + j * 32, 1280) : umma_desc_sw128(a0 + j * 32);
      umma_bf16(tmem_base, ad, umma_desc_sw128(b0 + j * 32), idesc, it > 0 ? 1u : 0u);
    }
    umma_commit(bars + 8);
    mbar_wait(bars + 8, 0, 0x4000u);
  } else if (warp == 1 && lane == 0) {
    for (int it = 0; it < p.tma_iters; ++it) {
      const int s = it & 3;
      if (it >= 4) mbar_wait(bars + s, ((it >> 2) - 1) & 1, 0x8000u);
      mbar_expect_tx(bars + s, 16384);
      const long long row = ((long long)(blockIdx.x + (long long)it * gridDim.x) * 128) % p.rows_total;
      tma_load_5d(sT + s * 16384, &tm2d, bars + s, 0, (int)row, 0, 0, 0);
    }
    for (int it = max(p.tma_iters - 4, 0); it < p.tma_iters; ++it) mbar_wait(bars + (it & 3), (it >> 2) & 1, 0x8000u);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}


// Latency probe (diagnostics): cycles per primitive, measured with clock64 by single threads of one CTA.
// out[0] arrive (count-1 barrier)  out[1] arrive.expect_tx(0)  out[2] test_wait on a completed phase
// out[3] tcgen05.commit -> phase visible (no MMA pending)  out[4] two-warp ping-pong round trip
// out[5] tcgen05.ld x16 + wait::ld  out[6] one 128x64x16 MMA + commit -> phase visible
__global__ void __launch_bounds__(64, 1) latency_probe_kernel(long long* out) {
  __shared__ __align__(1024) uint8_t sm[16384];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bars + i, 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 16384 / 4; i += 64) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  const int R = 256;
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 0; i < R; ++i) mbar_arrive(bars + 0);
    out[0] = (clock64() - t0) / R;
    t0 = clock64();
    for (int i = 0; i < R; ++i) mbar_expect_tx(bars + 1, 0);
    out[1] = (clock64() - t0) / R;
    t0 = clock64();
    int acc = 0;
    for (int i = 0; i < R; ++i) acc += mbar_test_wait(bars + 0, 1) ? 1 : 0;
    out[2] = (clock64() - t0) / R + (acc == -1);
    uint32_t ph = 0;
    t0 = clock64();
    for (int i = 0; i < R; ++i) {
      umma_commit(bars + 2);
      while (!mbar_test_wait(bars + 2, ph)) {}
      ph ^= 1;
    }
    out[3] = (clock64() - t0) / R;
    const uint32_t idesc = umma_idesc_bf16(128, 64);
    t0 = clock64();
    for (int i = 0; i < R; ++i) {
      umma_bf16(tm, umma_desc_sw128(smem_u32(sm)), umma_desc_sw128(smem_u32(sm)), idesc, 0u);
      umma_commit(bars + 2);
      while (!mbar_test_wait(bars + 2, ph)) {}
      ph ^= 1;
    }
    out[6] = (clock64() - t0) / R;
  }
  __syncthreads();
  // ping-pong: warp0 lane0 arrives on bars[0]->(wait by warp1), warp1 arrives on bars[3]
  if (threadIdx.x == 0) {
    // phases of bars[0] after R arrives above: parity state = R & 1 (R even -> back to 0)
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int i = 0; i < R; ++i) {
      mbar_arrive(bars + 0);
      while (!mbar_test_wait(bars + 3, ph)) {}
      ph ^= 1;
    }
    out[4] = (clock64() - t0) / R;
  } else if (threadIdx.x == 32) {
    uint32_t ph = 0;
    for (int i = 0; i < R; ++i) {
      while (!mbar_test_wait(bars + 0, ph)) {}
      ph ^= 1;
      mbar_arrive(bars + 3);
    }
  }
  __syncthreads();
  if (warp == 0) {
    long long t0 = clock64();
    uint32_t v[16];
    uint32_t sink = 0;
    for (int i = 0; i < R; ++i) {
      tmem_ld16(tm, v);
      tmem_ld_wait();
      sink += v[0];
    }
    if (lane == 0) out[5] = (clock64() - t0) / R + (sink == 0x12345678u);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 64); }
}

}  // namespace ypb
