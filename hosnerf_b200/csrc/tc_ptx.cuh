// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers shared by the tensor-core kernels of libhosnerf_b200.so
// (mlp_tc.cu: fused MLP kernels; gemm_tc.cu: layer-by-layer GEMMs of the training / split-precision paths).  sm_100a only.
#pragma once
#include "common.cuh"

namespace hos {

constexpr int kTileM = 128;
constexpr int kKB = 64;                    // K elements per chunk (128 bytes of fp16)
constexpr int kXChunkBytes = kTileM * 128; // 16 KB

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the waiting warp sleeps in hardware until the phase completes (or the
// hint expires) instead of spinning - a bare try_wait loop polls every few cycles and, with ~10 waiting warps per
// SM, was taking a third of all issue slots away from the epilogue / feature warps (ncu: BRA = 33 % of samples).
constexpr uint32_t kSuspendHintNs = 0x989680u;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(kSuspendHintNs) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// the same copy delivered to the same CTA-relative address (and mbarrier) of every CTA in `cta_mask` of the cluster
__device__ __forceinline__ void bulk_g2s_mcast(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- cluster-pair (cta_group::2) wrappers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Bounded waits for the pair kernel: a protocol bug traps (launch failure) instead of hanging the GPU.
// try_wait returns within a few tens of cycles whether or not a suspend-time hint is given (ncu: 57 % of all
// executed instructions were polling loops), and every poll takes an issue slot from the epilogue / feature
// warps on the same scheduler - so waits that are not on the MMA warp's critical path back off with
// nanosleep between polls.  kSleepNs == 0: pure spin (MMA issuer only).
template <int kSleepNs>
__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t polls = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (kSleepNs > 0) asm volatile("nanosleep.u32 %0;" ::"r"((uint32_t)kSleepNs));
    if (++polls > (kSleepNs > 0 ? (1u << 24) : (1u << 28))) __trap();     // seconds: far beyond any legitimate wait
  }
}
__device__ __forceinline__ void mbar_wait2_spin(uint32_t a0, uint32_t p0, uint32_t a1, uint32_t p1) {
  uint32_t polls = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%3], %4;\n\t"
        "and.pred p, p, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(a0), "r"(p0), "r"(a1), "r"(p1) : "memory");
    if (ok) return;
    if (++polls > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void mbar_wait4_spin(uint32_t a0, uint32_t p0, uint32_t a1, uint32_t p1, uint32_t a2, uint32_t p2,
                                                uint32_t a3, uint32_t p3) {
  uint32_t polls = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p, q, r, s;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%3], %4;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 r, [%5], %6;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 s, [%7], %8;\n\t"
        "and.pred p, p, q;\n\tand.pred r, r, s;\n\tand.pred p, p, r;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(a0), "r"(p0), "r"(a1), "r"(p1), "r"(a2), "r"(p2), "r"(a3), "r"(p3) : "memory");
    if (ok) return;
    if (++polls > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ bool elect_one() {      // one lane of the (converged) warp
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_commit_pair_addr(uint32_t bar_saddr) {     // arrives on the barrier in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_saddr),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_commit_mask_addr(uint32_t bar_saddr, uint16_t cta_mask) {   // arrives in every CTA of `cta_mask`
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_saddr),
               "h"(cta_mask)
               : "memory");
}
// wait until up to three phases have all completed: the polls overlap instead of paying three latencies in a row
__device__ __forceinline__ void mbar_wait3_spin(uint32_t a0, uint32_t p0, uint32_t a1, uint32_t p1, uint32_t a2, uint32_t p2) {
  uint32_t polls = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p, q, r;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%3], %4;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 r, [%5], %6;\n\t"
        "and.pred p, p, q;\n\tand.pred p, p, r;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(a0), "r"(p0), "r"(a1), "r"(p1), "r"(a2), "r"(p2) : "memory");
    if (ok) return;
    if (++polls > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {     // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {      // explicit ld.shared (a generic LD costs far more latency)
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// two fp32 (bit patterns) -> packed fp16x2 (lo = first), optionally with ReLU fused into the conversion
__device__ __forceinline__ uint32_t cvt_f16x2(uint32_t lo, uint32_t hi) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return d;
}
__device__ __forceinline__ uint32_t cvt_relu_f16x2(uint32_t lo, uint32_t hi) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return d;
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart
// (bit layout: cute/arch/mma_sm100_desc.hpp, SmemDescriptor).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);        // start address, 16 B units
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}
// Same for a K-major SWIZZLE_32B operand: [rows x 16] fp16 = 32 B per row, 8-row groups 256 B apart; the 16-byte half
// of element (r, k) is (k >> 3) ^ ((r >> 2) & 1)  (Swizzle<1,4,3>).  Used for the 4 KB constant ones operand.
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;                         // SWIZZLE_32B
  return d;
}
// Instruction descriptor, kind::f16: D=f32, A=B=f16, both K-major, M=128.
__host__ __device__ inline uint32_t umma_idesc_f16(int n, int m = kTileM) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}



// Instruction descriptor with both operands MN-major (weight-gradient GEMMs: the reduction runs over the ROWS of the
// tiled activations, so a [128 rows x 64 features] chunk is read as a [K = rows][MN = features] operand).
__host__ __device__ inline uint32_t umma_idesc_f16_mn(int n, int m) {
  return umma_idesc_f16(n, m) | (1u << 15) | (1u << 16);
}
// MN-major SWIZZLE_128B shared-memory descriptor (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>: in 16-byte
// units the canonical layout is ((8, n), (8, k)) : ((1, LBO), (8, SBO)) - 64 consecutive MN elements per K index, 8 K
// indices 128 B apart form a 1024-byte swizzle atom, atoms repeat every LBO bytes along MN and every SBO bytes along K).
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}

}  // namespace hos
