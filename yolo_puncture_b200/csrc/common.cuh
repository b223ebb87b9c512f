// Common device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers.
// Everything here is hand-written inline PTX for sm_100a (compile with
// -gencode arch=compute_100a,code=sm_100a); there is no other backend.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ypb {

// Device-side error word: set when a bounded mbarrier wait gives up (pipeline bug) so that a bad
// kernel shows up as an error code on the host instead of a hung GPU box.
__device__ unsigned int g_dev_error = 0;

// One lane of a converged warp, chosen by the hardware.  Code guarded by elect_one() is known to ptxas to run on a
// single lane, so TMA / tcgen05 instructions (which take uniform-register operands) are emitted straight-line;
// behind a plain `lane == 0` test ptxas wraps each of them in an ELECT/R2UR "waterfall" loop whose back-edge waits
// on the instruction's operand scoreboard, serialising the issuing thread with the TMA unit (~750 cycles per
// cp.async.bulk.tensor on B200, measured with tools/tma_bench.py).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// while its predecessor is still draining; pdl_wait() blocks until the predecessor grid has completed and its memory is
// visible, pdl_trigger() lets the successor's CTAs be scheduled as soon as this grid's CTAs free their SMs.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_relaxed(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.relaxed.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: returns false (and flags g_dev_error) if the phase never completes.
// Polls with the NON-suspending mbarrier.test_wait: measured on B200 (tools/tma_bench.py), a thread parked in
// mbarrier.try_wait costs ~730 cycles per producer/consumer hand-off regardless of how early the phase completes,
// which capped every pipeline in this file at one stage per 0.37 us; polling reacts within tens of cycles.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, unsigned int err_code) {
#ifdef YPB_TRYWAIT
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    if (mbar_try_wait(bar, parity)) return true;
  }
#else
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    if (mbar_test_wait(bar, parity)) return true;
  }
#endif
  atomicOr(&g_dev_error, err_code);
  return false;
}

// Bounded wait that parks the thread in the hardware (mbarrier.try_wait suspends until the phase flips or a time
// limit passes).  Slower to react than polling, but it consumes no issue slots: right when other resident CTAs have
// work for the SM sub-partition (the stem kernel), wrong for a warp-specialised role that owns its sub-partition.
__device__ __forceinline__ bool mbar_wait_blocking(uint64_t* bar, uint32_t parity, unsigned int err_code) {
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    if (mbar_try_wait(bar, parity)) return true;
  }
  atomicOr(&g_dev_error, err_code);
  return false;
}

// Same, with a back-off between polls (ns = 0: tight polling).  Idle pollers share the SM's memory-instruction
// queue with the epilogue warps' LDS/STS/SHFL: see the measurements quoted at the call sites in conv_tc.cuh.
__device__ __forceinline__ bool mbar_wait_bo(uint64_t* bar, uint32_t parity, unsigned int err_code, unsigned int ns) {
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    if (mbar_test_wait(bar, parity)) return true;
    if (ns) __nanosleep(ns);
  }
  atomicOr(&g_dev_error, err_code);
  return false;
}

// Whole-warp wait: every lane polls and the loop leaves when all of them have seen the phase complete.  The exit
// condition is a vote, i.e. warp-uniform as far as ptxas can tell, which keeps the loops AROUND the wait (and the
// descriptor arithmetic in them) on the uniform datapath: the MMA warp runs its tile / tap loops converged and
// elects a lane only for the tcgen05 instructions themselves.
__device__ __forceinline__ bool mbar_wait_warp(uint64_t* bar, uint32_t parity, unsigned int err_code, unsigned int ns) {
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    if (__all_sync(0xffffffffu, mbar_test_wait(bar, parity))) return true;
    if (ns) __nanosleep(ns);
  }
  atomicOr(&g_dev_error, err_code);
  return false;
}

// ------------------------------------------------------------------------------------------------
// thread-block clusters / CTA pairs (cta_group::2)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// ------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// CTA-pair loads: the data lands in THIS CTA's shared memory, the bytes are counted on a barrier given by its
// shared::cluster address - the leader CTA's (mapa_u32(bar, 0)), where the one MMA-issuing thread of the pair waits
__device__ __forceinline__ void tma_load_5d_2cta(void* smem, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1,
                                                 int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2cta(void* smem, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1,
                                                 int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// TMA stores (shared -> global, bulk async group): the epilogue's write-out without LSU instructions
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_addr, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_addr), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t smem_addr, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_addr), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// CTA-pair flavours (cta_group::2): one warp of EACH CTA of the pair allocates / frees, the leader's single thread issues
// the M = 256 MMAs (rows 0-127 accumulate in the leader's TMEM, 128-255 in the peer's), commits arrive on the same
// barrier offset in every CTA of `cta_mask`
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <int KS>
__device__ __forceinline__ void umma2_bf16_ksteps(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                  uint32_t idesc, uint32_t accf) {
#pragma unroll
  for (int j = 0; j < KS; ++j) umma2_bf16_lohi(d_tmem, a_lo + 2 * j, a_hi, b_lo + 2 * j, b_hi, idesc, j == 0 ? accf : 1u);
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate, one CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the two shared-memory descriptors given as (low word, high word): the high words (LBO/SBO/version/layout)
// are loop constants and stepping K by 16 bf16 (32 B) is "+2" on the low word's 14-bit address field, so the issuing
// thread spends ~3 instructions per MMA instead of rebuilding two 64-bit descriptors (~15 uniform-datapath
// instructions, measured ~100 cycles per 64-cycle MMA in the halo kernel).
__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// KS consecutive K=16 steps of one 128 x N tile (descriptor start addresses advance by 32 bytes = 2 units per step).
// ksteps is dispatched to the straight-line KS = 4 / 2 forms: predicating each MMA on `j < ksteps` costs a handful of
// uniform moves per instruction (ptxas gives every predicated MMA its own operand registers).
template <int KS>
__device__ __forceinline__ void umma_bf16_ksteps(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                 uint32_t idesc, uint32_t accf) {
#pragma unroll
  for (int j = 0; j < KS; ++j) umma_bf16_lohi(d_tmem, a_lo + 2 * j, a_hi, b_lo + 2 * j, b_hi, idesc, j == 0 ? accf : 1u);
}
__device__ __forceinline__ void umma_bf16_ksteps_n(int ksteps, uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                   uint32_t b_hi, uint32_t idesc, uint32_t accf) {
  if (ksteps == 4) {
    umma_bf16_ksteps<4>(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, accf);
  } else if (ksteps == 2) {
    umma_bf16_ksteps<2>(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, accf);
  } else {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (j < ksteps) umma_bf16_lohi(d_tmem, a_lo + 2 * j, a_hi, b_lo + 2 * j, b_hi, idesc, j == 0 ? accf : 1u);
  }
}
// low / high words of a K-major SWIZZLE_128B descriptor (see umma_desc_sw128): start>>4 | LBO=1<<16 ; SBO>>4 | v1 | layout 2
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFF) >> 4) | (1u << 16); }
__host__ __device__ constexpr uint32_t umma_desc_hi(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (2u << 29); }

// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (thread i = lane base + i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 B, 8-row atoms of 1024 B):
// start>>4 [0,14) | LBO>>4 = 1 [16,30) | SBO>>4 = 64 [32,46) | version 1 [46,48) | layout 2 [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16 instruction descriptor: fp32 accum, bf16 A/B, both K-major, shape M x N.
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                      // c_format = F32
  d |= 1u << 7;                      // a_format = BF16
  d |= 1u << 10;                     // b_format = BF16
  d |= static_cast<uint32_t>(N >> 3) << 17;
  d |= static_cast<uint32_t>(M >> 4) << 24;
  return d;
}

// SiLU(x) = x*sigmoid(x) = h + h*tanh(h), h = x/2: ONE special-function op (tanh.approx.f32, |err| <= 2^-11) instead
// of ex2 + rcp.  The conv epilogues are bound by the SFU pipe on B200 (16 results/clk/SM): ~2.7 G SiLUs per
// yolov8s-seg batch of 64.  Absolute error <= |x| * 2.5e-4, below half a bf16 ulp of the result except on the
// x << 0 tail where |SiLU| < 0.05 (error <= 2e-3 there); outputs are rounded to bf16 right after.
#ifndef YPB_EXACT_SILU
#define YPB_EXACT_SILU 0  // 1: full-precision SiLU everywhere (A/B build to quantify what tanh.approx costs in parity)
#endif
__device__ __forceinline__ float silu_f(float x) {
#if YPB_EXACT_SILU
  return x / (1.0f + expf(-x));
#endif
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
// Reference-accuracy variant (ex2 + rcp), used where the result is not rounded to bf16 (fp32 proto output).
__device__ __forceinline__ float silu_precise_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

}  // namespace ypb
