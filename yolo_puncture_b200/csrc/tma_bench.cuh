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

__global__ void __launch_bounds__(64, 1)
tma_bench_kernel(const __grid_constant__ CUtensorMap tm2d, const __grid_constant__ CUtensorMap tm5d,
                 const TmaBenchParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * 16384);
  uint64_t* empty_bar = full_bar + p.stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int tiles_w = p.W / 16, tiles_h = p.H / 8;
    for (int it = 0; it < p.iters; ++it) {
      const int s = it % p.stages;
      mbar_wait(empty_bar + s, ((it / p.stages) & 1) ^ 1, 0x1000u);
      mbar_expect_tx(full_bar + s, 16384);
      if (p.mode == 2) {
        const int tile = (blockIdx.x + (it / 9) * gridDim.x) % (tiles_w * tiles_h * p.B);
        const int t = it % 9;
        const int b = tile / (tiles_w * tiles_h), r = tile % (tiles_w * tiles_h);
        tma_load_5d(smem + s * 16384, &tm5d, full_bar + s, 0, (r % tiles_w) * 16 + t % 3 - 1, (r / tiles_w) * 8 + t / 3 - 1, b, 0);
      } else {
        const long long row = p.mode == 1 ? 0 : ((long long)(blockIdx.x + (long long)it * gridDim.x) * 128) % p.rows_total;
        tma_load_5d(smem + s * 16384, &tm2d, full_bar + s, 0, (int)row, 0, 0, 0);
      }
    }
  } else if (threadIdx.x == 32) {
    for (int it = 0; it < p.iters; ++it) {
      const int s = it % p.stages;
      mbar_wait(full_bar + s, (it / p.stages) & 1, 0x2000u);
      mbar_arrive(empty_bar + s);
    }
  }
}

}  // namespace ypb
