// Micro-benchmarks that size the tcgen05 MLP kernel (sm_100a):
//   1. TMEM read-out rate (tcgen05.ld 32x32b.x32 / .x64) vs number of warps, pipelined or not;
//   2. tcgen05.mma issue rate, M=128 N=256 K=16 SS, cta_group::1 and cta_group::2 (M=256 over a CTA pair),
//      alone and while 8 warps drain the other accumulator and store fp16 activations to shared memory;
//   3. a numerical check of the cta_group::2 data path (A rows / B rows split over the pair).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/bin/ubench_tc scripts/ubench_tc.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);          \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

template <int CG>
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
template <int CG>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  else
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
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ inline uint32_t umma_idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__host__ __device__ inline uint32_t tile_byte_offset(int r, int k) {
  return (uint32_t)(r * 128 + ((((k >> 3) ^ (r & 7))) << 4) + ((k & 7) << 1));
}

// ------------------------------------------------------------------------------------------------
// Test 1: TMEM read rate.  W warps, each reads `iters` x (32 lanes x 32 columns x 4 B = 4 KB).
// mode 0: ld + wait each; mode 1: four loads in flight per wait; mode 2: x16 loads, 8 in flight.
__global__ void __launch_bounds__(512, 1) k_tmem_read(int iters, int mode, long long* cycles) {
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t sink = 0;
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) {
    for (int i = 0; i < iters; ++i) {
      uint32_t v[32];
      tmem_ld32_nowait(base + (uint32_t)(((i + (warp >> 2) * 4) * 32) & 511), v);
      tmem_wait_ld();
      sink ^= v[0] ^ v[31];
    }
  } else if (mode == 1) {
    for (int i = 0; i < iters; i += 4) {
      uint32_t a[32], b[32], c[32], d[32];
      const uint32_t col = (uint32_t)(((i + (warp >> 2) * 4) * 32) & 511);
      tmem_ld32_nowait(base + ((col + 0) & 511), a);
      tmem_ld32_nowait(base + ((col + 32) & 511), b);
      tmem_ld32_nowait(base + ((col + 64) & 511), c);
      tmem_ld32_nowait(base + ((col + 96) & 511), d);
      tmem_wait_ld();
      sink ^= a[0] ^ b[31] ^ c[5] ^ d[7];
    }
  } else {
    for (int i = 0; i < iters; i += 4) {
      uint32_t a[16], b[16], c[16], d[16], e[16], f[16], g[16], h[16];
      const uint32_t col = (uint32_t)(((i + (warp >> 2) * 4) * 32) & 511);
      tmem_ld16_nowait(base + ((col + 0) & 511), a);
      tmem_ld16_nowait(base + ((col + 16) & 511), b);
      tmem_ld16_nowait(base + ((col + 32) & 511), c);
      tmem_ld16_nowait(base + ((col + 48) & 511), d);
      tmem_ld16_nowait(base + ((col + 64) & 511), e);
      tmem_ld16_nowait(base + ((col + 80) & 511), f);
      tmem_ld16_nowait(base + ((col + 96) & 511), g);
      tmem_ld16_nowait(base + ((col + 112) & 511), h);
      tmem_wait_ld();
      sink ^= a[0] ^ b[15] ^ c[5] ^ d[7] ^ e[1] ^ f[2] ^ g[3] ^ h[4];
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  if (sink == 0x12345678u) cycles[1] = sink;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512u));
  }
}

// ------------------------------------------------------------------------------------------------
// Test 2: MMA issue rate with an optional concurrent "epilogue" (8 warps: tcgen05.ld of the other
// accumulator buffer, convert to fp16, 16-byte swizzled stores into a 64 KB activation buffer).
// smem: A 4 x 16 KB | B 3 x (CG==1 ? 32 : 16) KB | H 64 KB.
template <int CG>
__global__ void __launch_bounds__(512, 1) k_mma_rate(int n_chunks, int epi_tiles, int epi_sts, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int kBStage = CG == 1 ? 32768 : 16384;
  unsigned char* sA = smem;
  unsigned char* sB = sA + 4 * 16384;
  unsigned char* sH = sB + 3 * kBStage;
  __shared__ uint64_t bar_done;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < (4 * 16384 + 3 * kBStage) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  long long t0 = clock64();
  if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_f16(CG == 1 ? 128 : 256, 256);
      if (lane == 0) {
        for (int c = 0; c < n_chunks; ++c) {
          const uint32_t a_base = smem_u32(sA + (c & 3) * 16384);
          const uint32_t b_base = smem_u32(sB + (c % 3) * kBStage);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma<CG>(tmem_base, umma_desc(a_base + k * 32), umma_desc(b_base + k * 32), idesc, (c | k) != 0);
        }
        tc_commit<CG>(&bar_done);
      }
      __syncwarp();
    }
    if (n_chunks > 0) {
      bool ok = mbar_wait(&bar_done, 0);
      tc_fence_after();
      long long t1 = clock64();
      if (lane == 0 && blockIdx.x == 0) cycles[0] = ok ? t1 - t0 : -1;
    }
  } else if (warp >= 4 && warp < 12 && epi_tiles > 0) {
    const int q = warp & 3, ch = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    const uint32_t acc = tmem_base + 256 + ((uint32_t)(q * 32) << 16);
    for (int t = 0; t < epi_tiles; ++t) {
      for (int c0 = ch * 128; c0 < ch * 128 + 128; c0 += 64) {
        uint32_t v[32], w[32];
        tmem_ld32_nowait(acc + (uint32_t)c0, v);
        tmem_ld32_nowait(acc + (uint32_t)c0 + 32, w);
        tmem_wait_ld();
        if (epi_sts) {
          unsigned char* dst = sH + (c0 >> 6) * 16384;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 pk;
            __half2 h0 = __floats2half2_rn(fmaxf(__uint_as_float(v[g * 8 + 0]) + 1.f, 0.f), fmaxf(__uint_as_float(v[g * 8 + 1]) + 1.f, 0.f));
            __half2 h1 = __floats2half2_rn(fmaxf(__uint_as_float(v[g * 8 + 2]) + 1.f, 0.f), fmaxf(__uint_as_float(v[g * 8 + 3]) + 1.f, 0.f));
            __half2 h2 = __floats2half2_rn(fmaxf(__uint_as_float(v[g * 8 + 4]) + 1.f, 0.f), fmaxf(__uint_as_float(v[g * 8 + 5]) + 1.f, 0.f));
            __half2 h3 = __floats2half2_rn(fmaxf(__uint_as_float(v[g * 8 + 6]) + 1.f, 0.f), fmaxf(__uint_as_float(v[g * 8 + 7]) + 1.f, 0.f));
            pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
            pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(dst + tile_byte_offset(r, g * 8)) = pk;
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 pk;
            __half2 h0 = __floats2half2_rn(fmaxf(__uint_as_float(w[g * 8 + 0]) + 1.f, 0.f), fmaxf(__uint_as_float(w[g * 8 + 1]) + 1.f, 0.f));
            __half2 h1 = __floats2half2_rn(fmaxf(__uint_as_float(w[g * 8 + 2]) + 1.f, 0.f), fmaxf(__uint_as_float(w[g * 8 + 3]) + 1.f, 0.f));
            __half2 h2 = __floats2half2_rn(fmaxf(__uint_as_float(w[g * 8 + 4]) + 1.f, 0.f), fmaxf(__uint_as_float(w[g * 8 + 5]) + 1.f, 0.f));
            __half2 h3 = __floats2half2_rn(fmaxf(__uint_as_float(w[g * 8 + 6]) + 1.f, 0.f), fmaxf(__uint_as_float(w[g * 8 + 7]) + 1.f, 0.f));
            pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
            pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(dst + tile_byte_offset(r, 32 + g * 8)) = pk;
          }
        } else if ((v[0] ^ w[3]) == 0x12345678u) {
          cycles[3] = 1;
        }
      }
    }
    long long t1 = clock64();
    if (threadIdx.x == 128 && blockIdx.x == 0) cycles[1] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// ------------------------------------------------------------------------------------------------
// Test 3: cta_group::2 numerics.  D[256 x 256] = A[256 x 64] * B[256 x 64]^T ; CTA r holds A rows
// [128 r, 128 r + 128) and B rows (output columns) [128 r, 128 r + 128); D rows of CTA r land in
// CTA r's TMEM lanes, all 256 columns.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k_cg2_check(const __half* A, const __half* B, float* D, int* status) {
  __shared__ __align__(1024) unsigned char sA[16384];
  __shared__ __align__(1024) unsigned char sB[16384];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int r = i >> 6, k = i & 63;
    *reinterpret_cast<__half*>(sA + tile_byte_offset(r, k)) = A[(rank * 128 + r) * 64 + k];
    *reinterpret_cast<__half*>(sB + tile_byte_offset(r, k)) = B[(rank * 128 + r) * 64 + k];
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  if (warp == 1 && rank == 0 && lane == 0) {
    const uint32_t idesc = umma_idesc_f16(256, 256);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      tc_mma<2>(tmem_base, umma_desc(smem_u32(sA) + k * 32), umma_desc(smem_u32(sB) + k * 32), idesc, k != 0);
    tc_commit<2>(&bar_done);
  }
  __syncwarp();
  bool ok = mbar_wait(&bar_done, 0);
  tc_fence_after();
  if (!ok) {
    if (threadIdx.x == 0) status[rank] = -1;
  } else {
    const int r = warp * 32 + lane;
    for (int c0 = 0; c0 < 256; c0 += 32) {
      uint32_t v[32];
      tmem_ld32_nowait(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_wait_ld();
      for (int j = 0; j < 32; ++j) D[(size_t)(rank * 128 + r) * 256 + c0 + j] = __uint_as_float(v[j]);
    }
    if (threadIdx.x == 0) status[rank] = (int)tmem_base + 1;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
  }
}


// ------------------------------------------------------------------------------------------------
// Test 4: L2 -> shared-memory streaming rate of cp.async.bulk when every SM re-reads the same weight
// image (the MLP kernel's weight ring).  `stages` copies of `chunk` bytes in flight per CTA; with
// mcast != 0 the two CTAs of a cluster each fetch half of every chunk and multicast it to both.
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__global__ void __launch_bounds__(128, 1) k_l2_stream(const unsigned char* src, int image_bytes, int chunk, int stages,
                                                       int n_chunks, int mcast, int split, int dephase, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full[16];
  const uint32_t rank = mcast ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&bar_full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (mcast) cluster_sync();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    const int per_image = image_bytes / chunk;
    for (int i = 0; i < n_chunks + stages; ++i) {
      if (i >= stages) {
        // consume chunk i - stages: wait until it has landed; with multicast both CTAs must be done with the
        // slot before either refills it - approximated by a cluster barrier every chunk (upper bound on cost)
        mbar_wait(&bar_full[(i - stages) % stages], ((i - stages) / stages) & 1);
      }
      if (i < n_chunks) {
        const int s = i % stages;
        const unsigned char* g = src + (size_t)((i + (dephase ? blockIdx.x * dephase : 0)) % per_image) * chunk;
        mbar_expect_tx(&bar_full[s], chunk);
        if (!mcast) {
          for (int q = 0; q < split; ++q)
            bulk_g2s(smem + s * chunk + q * (chunk / split), g + q * (chunk / split), chunk / split, &bar_full[s]);
        }
        else bulk_g2s_mc(smem + s * chunk + rank * (chunk / 2), g + rank * (chunk / 2), chunk / 2, &bar_full[s], 3);
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (mcast) cluster_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}
static void run_l2(const unsigned char* d_src, int image_bytes, int chunk, int stages, int mcast, long long* d_cyc, int split = 1, int dephase = 0) {
  const int n_chunks = 4096;
  const int smem = 1024 + chunk * stages;
  CK(cudaFuncSetAttribute(k_l2_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = mcast ? 2 : 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  CK(cudaMemset(d_cyc, 0, 64));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  CK(cudaLaunchKernelEx(&cfg, k_l2_stream, d_src, image_bytes, chunk, stages, n_chunks, mcast, split, dephase, d_cyc));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  long long h[1];
  CK(cudaMemcpy(h, d_cyc, 8, cudaMemcpyDeviceToHost));
  printf("l2 stream: split %d dephase %d image %4d KB chunk %2d KB stages %2d mcast %d: %9lld cyc -> %6.1f B/clk/SM landed, %.3f ms, %.2f TB/s landed aggregate\n",
         split, dephase, image_bytes >> 10, chunk >> 10, stages, mcast, h[0], (double)n_chunks * chunk / (double)h[0], ms,
         148.0 * n_chunks * chunk / (ms * 1e-3) / 1e12);
}

// Test 5: same streaming loop with (a) test_wait polling, (b) several producer warps, (c) 2-D tensor-map TMA.
__device__ __forceinline__ bool mbar_poll(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void tma_2d_g2s(void* dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// mode 0: bulk + try_wait, 1: bulk + test_wait polling, 2: tensor TMA + try_wait.  `producers` warps each run
// their own ring of `stages` slots.
__global__ void __launch_bounds__(128, 1) k_stream2(const unsigned char* src, const __grid_constant__ CUtensorMap tmap,
                                                     int image_bytes, int chunk, int stages, int n_chunks, int mode,
                                                     int producers, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full[32];
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages * producers; ++s) mbar_init(&bar_full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && warp < producers) {
    const int per_image = image_bytes / chunk;
    uint64_t* bars = bar_full + warp * stages;
    unsigned char* ring = smem + (size_t)warp * stages * chunk;
    for (int i = 0; i < n_chunks + stages; ++i) {
      if (i >= stages) {
        uint64_t* b = &bars[(i - stages) % stages];
        const uint32_t par = ((i - stages) / stages) & 1;
        if (mode == 1) mbar_poll(b, par); else mbar_wait(b, par);
      }
      if (i < n_chunks) {
        const int s = i % stages;
        const int ci = (i * producers + warp) % per_image;
        mbar_expect_tx(&bars[s], chunk);
        if (mode == 2) tma_2d_g2s(ring + s * chunk, &tmap, 0, ci * (chunk / 128), &bars[s]);
        else bulk_g2s(ring + s * chunk, src + (size_t)ci * chunk, chunk, &bars[s]);
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void run_stream2(const unsigned char* d_src, int image_bytes, int chunk, int stages, int mode, int producers, long long* d_cyc) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return; }
  }
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {128, (cuuint64_t)(image_bytes / 128)};      // bytes as uint8: rows of 128 B
  cuuint64_t gstride[1] = {128};
  cuuint32_t box[2] = {128, (cuuint32_t)(chunk / 128)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)d_src, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return; }
  const int n_chunks = 2048;
  const int smem = 1024 + chunk * stages * producers;
  CK(cudaFuncSetAttribute(k_stream2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaMemset(d_cyc, 0, 64));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  k_stream2<<<148, 128, smem>>>(d_src, tmap, image_bytes, chunk, stages, n_chunks, mode, producers, d_cyc);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  long long h[1];
  CK(cudaMemcpy(h, d_cyc, 8, cudaMemcpyDeviceToHost));
  printf("stream2: mode %d (%s) producers %d chunk %2d KB stages %2d: %9lld cyc = %7.1f cyc/chunk/producer -> %6.1f B/clk/SM, %.3f ms\n", mode,
         mode == 0 ? "bulk+try_wait" : mode == 1 ? "bulk+poll" : "tensor2d", producers, chunk >> 10, stages, h[0],
         (double)h[0] / n_chunks, (double)n_chunks * producers * chunk / (double)h[0], ms);
}

// Test 6: burst - one thread issues `n` bulk copies of `chunk` bytes back to back (distinct smem slots, one
// mbarrier), then waits once.  lanes > 1: that many lanes of warp 0 each issue n / lanes of them.
__global__ void __launch_bounds__(128, 1) k_burst(const unsigned char* src, int chunk, int n, int lanes, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64(), t_issue = 0;
  if (threadIdx.x == 0) mbar_expect_tx(&bar, chunk * n);
  __syncwarp();
  if (threadIdx.x < lanes) {
    for (int i = threadIdx.x; i < n; i += lanes) bulk_g2s(smem + i * chunk, src + (size_t)((i + blockIdx.x) % 64) * chunk, chunk, &bar);
    t_issue = clock64();
    mbar_wait(&bar, 0);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = t_issue - t0; }
}
static void run_burst(const unsigned char* d_src, int chunk, int n, int lanes, long long* d_cyc) {
  const int smem = 1024 + chunk * n;
  CK(cudaFuncSetAttribute(k_burst, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaMemset(d_cyc, 0, 64));
  k_burst<<<148, 128, smem>>>(d_src, chunk, n, lanes, d_cyc);
  CK(cudaDeviceSynchronize());
  long long h[2];
  CK(cudaMemcpy(h, d_cyc, 16, cudaMemcpyDeviceToHost));
  printf("burst: %2d x %5d B, %2d lanes: issue %6lld cyc, all landed after %6lld cyc -> %6.1f B/clk/SM\n", n, chunk, lanes, h[1], h[0],
         (double)n * chunk / (double)h[0]);
}

// Test 7: where does a ring iteration spend its time?  Stamps after wait / expect_tx / copy for 16 iterations.
// variant 0: lane 0 does everything; 1: lane 1 does the expect_tx (lane 0 waits + copies); 2: copy first, then expect_tx.
__global__ void __launch_bounds__(128, 1) k_ring_trace(const unsigned char* src, int chunk, int stages, int n_chunks, int variant,
                                                        long long* stamps) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full[16];
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&bar_full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    long long t0 = clock64();
    for (int i = 0; i < n_chunks + stages; ++i) {
      long long a = 0, b = 0, c = 0;
      if (i >= stages && lane == 0) mbar_wait(&bar_full[(i - stages) % stages], ((i - stages) / stages) & 1);
      if (variant == 1) __syncwarp();
      a = clock64();
      if (i < n_chunks) {
        const int s = i % stages;
        const unsigned char* g = src + (size_t)(i % 64) * chunk;
        if (variant == 0) {
          if (lane == 0) { mbar_expect_tx(&bar_full[s], chunk); b = clock64(); bulk_g2s(smem + s * chunk, g, chunk, &bar_full[s]); }
        } else if (variant == 1) {
          if (lane == 1) mbar_expect_tx(&bar_full[s], chunk);
          b = clock64();
          if (lane == 0) bulk_g2s(smem + s * chunk, g, chunk, &bar_full[s]);
        } else {
          if (lane == 0) { bulk_g2s(smem + s * chunk, g, chunk, &bar_full[s]); b = clock64(); mbar_expect_tx(&bar_full[s], chunk); }
        }
      }
      c = clock64();
      if (blockIdx.x == 0 && lane == 0 && i >= 200 && i < 216) {
        stamps[(i - 200) * 3 + 0] = a - t0; stamps[(i - 200) * 3 + 1] = b - t0; stamps[(i - 200) * 3 + 2] = c - t0;
      }
    }
    if (blockIdx.x == 0 && lane == 0) stamps[48] = clock64() - t0;
  }
}
static void run_trace(const unsigned char* d_src, int chunk, int stages, int variant) {
  long long* d_st;
  CK(cudaMalloc(&d_st, 64 * 8));
  CK(cudaMemset(d_st, 0, 64 * 8));
  const int smem = 1024 + chunk * stages;
  CK(cudaFuncSetAttribute(k_ring_trace, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_ring_trace<<<148, 128, smem>>>(d_src, chunk, stages, 1024, variant, d_st);
  CK(cudaDeviceSynchronize());
  long long h[49];
  CK(cudaMemcpy(h, d_st, 49 * 8, cudaMemcpyDeviceToHost));
  printf("ring trace: chunk %d stages %d variant %d: total %lld cyc = %.1f / chunk\n  (after wait, mid, after issue) deltas:", chunk, stages, variant,
         h[48], (double)h[48] / 1024);
  for (int i = 1; i < 12; ++i) printf(" [w%lld m%lld i%lld]", h[i * 3] - h[(i - 1) * 3 + 2], h[i * 3 + 1] - h[i * 3], h[i * 3 + 2] - h[i * 3 + 1]);
  printf("\n");
  CK(cudaFree(d_st));
}

// Test 8: issue cost of every tcgen05.mma / tcgen05.commit in a (4 MMA + commit) x n_chunks stream, plus the cost
// of a try_wait on an already-completed barrier placed between chunks (what the MLP kernel's MMA thread does).
template <int CG>
__global__ void __launch_bounds__(128, 1) k_issue_trace(int n_chunks, int with_wait, long long* stamps) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int kBStage = CG == 1 ? 32768 : 16384;
  unsigned char* sA = smem;
  unsigned char* sB = sA + 4 * 16384;
  __shared__ uint64_t bar_done, bar_ready, bar_stage[4];
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < (4 * 16384 + 3 * kBStage) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar_done, 1);
    mbar_init(&bar_ready, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_stage[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_ready)) : "memory");   // phase 0 complete
  }
  if (warp == 1) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  if (warp == 1 && rank == 0 && lane == 0) {
    const uint32_t idesc = umma_idesc_f16(CG == 1 ? 128 : 256, 256);
    long long t0 = clock64();
    for (int c = 0; c < n_chunks; ++c) {
      const uint32_t a_base = smem_u32(sA + (c & 3) * 16384);
      const uint32_t b_base = smem_u32(sB + (c % 3) * kBStage);
      long long* st = stamps + c * 8;
      if (with_wait) mbar_wait(&bar_ready, 0);
      const long long w = clock64();
      tc_mma<CG>(tmem_base, umma_desc(a_base), umma_desc(b_base), idesc, c != 0);
      const long long m0 = clock64();
      tc_mma<CG>(tmem_base, umma_desc(a_base + 32), umma_desc(b_base + 32), idesc, 1);
      const long long m1 = clock64();
      tc_mma<CG>(tmem_base, umma_desc(a_base + 64), umma_desc(b_base + 64), idesc, 1);
      const long long m2 = clock64();
      tc_mma<CG>(tmem_base, umma_desc(a_base + 96), umma_desc(b_base + 96), idesc, 1);
      const long long m3 = clock64();
      tc_commit<CG>(&bar_stage[c & 3]);
      const long long cm = clock64();
      if (blockIdx.x == 0 && c < 32) { st[0] = w - t0; st[1] = m0 - t0; st[2] = m1 - t0; st[3] = m2 - t0; st[4] = m3 - t0; st[5] = cm - t0; }
    }
    tc_commit<CG>(&bar_done);
    mbar_wait(&bar_done, 0);
    if (blockIdx.x == 0) stamps[32 * 8] = clock64() - t0;
  }
  if (CG == 2 && warp == 1 && rank == 1 && lane == 0) mbar_wait(&bar_done, 0);
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}
template <int CG>
static void run_issue_trace(int with_wait) {
  long long* d;
  CK(cudaMalloc(&d, 300 * 8));
  CK(cudaMemset(d, 0, 300 * 8));
  const int smem = 1024 + 4 * 16384 + 3 * 32768;
  CK(cudaFuncSetAttribute(k_issue_trace<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, k_issue_trace<CG>, 24, with_wait, d));
  CK(cudaDeviceSynchronize());
  long long h[300];
  CK(cudaMemcpy(h, d, 300 * 8, cudaMemcpyDeviceToHost));
  printf("issue trace cg=%d with_wait=%d: total %lld cyc for 24 chunks (96 MMAs)\n  per chunk [wait-or-loop, mma0, mma1, mma2, mma3, commit]:", CG, with_wait, h[256]);
  for (int c = 1; c < 12; ++c) {
    long long* st = h + c * 8;
    printf(" [%lld %lld %lld %lld %lld %lld]", st[0] - h[(c - 1) * 8 + 5], st[1] - st[0], st[2] - st[1], st[3] - st[2], st[4] - st[3], st[5] - st[4]);
  }
  printf("\n");
  CK(cudaFree(d));
}

template <int CG>
static void run_mma(int n_chunks, int epi_tiles, int epi_sts, long long* d_cyc, const char* label) {
  const int smem = 1024 + 4 * 16384 + 3 * (CG == 1 ? 32768 : 16384) + 65536;
  CK(cudaFuncSetAttribute(k_mma_rate<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaMemset(d_cyc, 0, 64));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148);
  cfg.blockDim = dim3(512);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  CK(cudaLaunchKernelEx(&cfg, k_mma_rate<CG>, n_chunks, epi_tiles, epi_sts, d_cyc));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  long long h[4];
  CK(cudaMemcpy(h, d_cyc, 32, cudaMemcpyDeviceToHost));
  printf("%-44s cg=%d chunks=%5d epi_tiles=%4d sts=%d | mma %9lld cyc", label, CG, n_chunks, epi_tiles, epi_sts, h[0]);
  if (n_chunks > 0 && h[0] > 0) printf(" = %6.1f cyc/instr", (double)h[0] / (n_chunks * 4.0));
  if (epi_tiles > 0) printf(" | epi %9lld cyc = %7.1f cyc/tile (%5.1f B/clk TMEM)", h[1], (double)h[1] / epi_tiles, 131072.0 * epi_tiles / (double)h[1]);
  printf(" | %.3f ms\n", ms);
}

int main() {
  long long* d_cyc;
  CK(cudaMalloc(&d_cyc, 64));
  if (getenv("UB_ISSUE")) {
    run_issue_trace<1>(0); run_issue_trace<1>(1); run_issue_trace<2>(0); run_issue_trace<2>(1);
    return 0;
  }
  // ---- test 4: L2 streaming
  {
    unsigned char* d_src;
    CK(cudaMalloc(&d_src, 4 << 20));
    CK(cudaMemset(d_src, 1, 4 << 20));
    for (int v = 0; v < 3; ++v) { run_trace(d_src, 16384, 3, v); run_trace(d_src, 16384, 6, v); }
    run_trace(d_src, 32768, 3, 0);
    if (getenv("UB_TRACE_ONLY")) return 0;
    for (int rep = 0; rep < 2; ++rep) {
      run_burst(d_src, 16384, 1, 1, d_cyc);
      run_burst(d_src, 16384, 2, 1, d_cyc);
      run_burst(d_src, 16384, 4, 1, d_cyc);
      run_burst(d_src, 16384, 8, 1, d_cyc);
      run_burst(d_src, 16384, 12, 1, d_cyc);
      run_burst(d_src, 16384, 12, 4, d_cyc);
      run_burst(d_src, 32768, 6 - 4 * 0, 1, d_cyc);
      run_burst(d_src, 4096, 32, 1, d_cyc);
      run_burst(d_src, 4096, 32, 32, d_cyc);
      run_burst(d_src, 1024, 128, 32, d_cyc);
    }
    for (int mode = 0; mode < 3; ++mode) {
      run_stream2(d_src, 1 << 20, 32768, 3, mode, 1, d_cyc);
      run_stream2(d_src, 1 << 20, 16384, 3, mode, 1, d_cyc);
      run_stream2(d_src, 1 << 20, 16384, 6, mode, 1, d_cyc);
      run_stream2(d_src, 1 << 20, 16384, 3, mode, 2, d_cyc);
      run_stream2(d_src, 1 << 20, 16384, 3, mode, 4, d_cyc);
      run_stream2(d_src, 1 << 20, 8192, 4, mode, 4, d_cyc);
    }
    run_l2(d_src, 1 << 20, 32768, 1, 0, d_cyc);
    run_l2(d_src, 1 << 20, 32768, 2, 0, d_cyc);
    run_l2(d_src, 1 << 20, 32768, 3, 0, d_cyc);
    run_l2(d_src, 1 << 20, 16384, 1, 0, d_cyc);
    run_l2(d_src, 1 << 20, 16384, 3, 0, d_cyc);
    run_l2(d_src, 1 << 20, 16384, 12, 0, d_cyc);
    run_l2(d_src, 1 << 20, 32768, 3, 0, d_cyc, 4, 0);
    run_l2(d_src, 1 << 20, 32768, 3, 0, d_cyc, 16, 0);
    run_l2(d_src, 1 << 20, 32768, 3, 0, d_cyc, 1, 1);
    run_l2(d_src, 1 << 20, 32768, 3, 0, d_cyc, 1, 3);
    run_l2(d_src, 1 << 20, 16384, 3, 0, d_cyc, 1, 1);
    run_l2(d_src, 1 << 20, 16384, 6, 0, d_cyc, 1, 5);
    run_l2(d_src, 1 << 20, 16384, 6, 0, d_cyc, 4, 5);
    run_l2(d_src, 4 << 20, 16384, 6, 0, d_cyc, 1, 7);
    run_l2(d_src, 1 << 20, 16384, 6, 1, d_cyc, 1, 0);
  }
  if (getenv("UB_L2_ONLY")) return 0;
  // ---- test 3 first: cg2 numerics
  {
    std::vector<__half> hA(256 * 64), hB(256 * 64);
    std::vector<float> fA(256 * 64), fB(256 * 64);
    srand(1);
    for (int i = 0; i < 256 * 64; ++i) {
      fA[i] = (float)((rand() % 9) - 4);
      fB[i] = (float)((rand() % 7) - 3);
      hA[i] = __float2half(fA[i]);
      hB[i] = __float2half(fB[i]);
    }
    __half *dA, *dB;
    float* dD;
    int* dS;
    CK(cudaMalloc(&dA, hA.size() * 2));
    CK(cudaMalloc(&dB, hB.size() * 2));
    CK(cudaMalloc(&dD, 256 * 256 * 4));
    CK(cudaMalloc(&dS, 8));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, 256 * 256 * 4));
    CK(cudaMemset(dS, 0, 8));
    k_cg2_check<<<2, 128>>>(dA, dB, dD, dS);
    CK(cudaDeviceSynchronize());
    std::vector<float> hD(256 * 256);
    int hS[2];
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hS, dS, 8, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int m = 0; m < 256; ++m)
      for (int n = 0; n < 256; ++n) {
        float ref = 0.f;
        for (int k = 0; k < 64; ++k) ref += fA[m * 64 + k] * fB[n * 64 + k];
        if (ref != hD[m * 256 + n]) {
          if (bad < 8) printf("  cg2 mismatch D[%d][%d] = %g, want %g\n", m, n, hD[m * 256 + n], ref);
          ++bad;
        }
      }
    printf("cg2 numerics: status {%d, %d}, mismatches %d / 65536 -> %s\n", hS[0], hS[1], bad, bad == 0 && hS[0] > 0 && hS[1] > 0 ? "OK" : "FAIL");
  }
  // ---- test 1: TMEM read rate
  for (int mode = 0; mode < 3; ++mode)
    for (int warps = 4; warps <= 16; warps *= 2) {
      const int iters = 4096;
      CK(cudaMemset(d_cyc, 0, 64));
      k_tmem_read<<<148, warps * 32>>>(iters, mode, d_cyc);
      CK(cudaDeviceSynchronize());
      long long h[2];
      CK(cudaMemcpy(h, d_cyc, 16, cudaMemcpyDeviceToHost));
      printf("tmem read: mode %d (%s) warps %2d: %9lld cyc, %6.1f B/clk/SM\n", mode,
             mode == 0 ? "x32, ld+wait" : mode == 1 ? "x32, 4 in flight" : "x16, 8 in flight", warps, h[0],
             (double)warps * iters * 4096.0 / (double)h[0]);
    }
  // ---- test 2: MMA rates
  run_mma<1>(2048, 0, 0, d_cyc, "mma alone");
  run_mma<2>(2048, 0, 0, d_cyc, "mma alone");
  run_mma<1>(0, 256, 0, d_cyc, "epilogue ld only");
  run_mma<1>(0, 256, 1, d_cyc, "epilogue ld+cvt+sts");
  run_mma<1>(2048, 256, 0, d_cyc, "mma + epilogue ld");
  run_mma<1>(2048, 256, 1, d_cyc, "mma + epilogue ld+cvt+sts");
  run_mma<2>(2048, 256, 0, d_cyc, "mma + epilogue ld");
  run_mma<2>(2048, 256, 1, d_cyc, "mma + epilogue ld+cvt+sts");
  run_mma<2>(2048, 512, 1, d_cyc, "mma + epilogue ld+cvt+sts (long)");
  return 0;
}
