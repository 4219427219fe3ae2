// Layer GEMMs on the 5th-gen tensor cores over ROW-MAJOR fp16 matrices, fed by TMA tensor maps (sm_100a).
// SURVEY 8a rows a9 / a17 / a19 (layer by layer) and the backward half of the same layers (VERDICT N1, N2).
//
// Three products, all with fp32 accumulation in TMEM and 2-CTA clusters (tcgen05.mma.cta_group::2, M = 256):
//
//   mode 0  "NT"  Y[rows, n]  = act([A0 | A1][rows, k0 + k1] * [W0 | W1]^T + b)      forward of nn.Linear (S1 model.py:212-259)
//   mode 1  "NN"  dA[rows, n] = (dZ[rows, k0] * W[k0, n]) .* (mask > 0)               data gradient; W read as an MN-major operand
//   wgrad   "TN"  dW[m, nq]  += P[rows, m]^T * Q[rows, nq]                            weight gradient; both operands MN-major,
//                                                                                     split over row tiles, fp32 atomics
//
// Split precision (VERDICT N2): with `*_lo` planes given every product runs as three tensor-core passes
//   A_hi W_hi + A_lo W_hi + A_hi W_lo        (x = hi + lo, hi = fp16(x), lo = fp16(x - hi): ~22 significant bits)
// into the same fp32 accumulator, and the epilogue re-splits the fp32 result into hi / lo planes for the next layer -
// the tcgen05 path that meets the "1e-4 rel fp32" parity gate.
//
// Operands are plain row-major fp16 (torch tensors), so autograd glue, tiny heads and these kernels share buffers.
// TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) lands each [128 rows x 64 columns] box in shared memory in exactly the
// canonical UMMA layout: row r at byte 128 r, 16-byte groups XOR (r & 7).  Read with a K-major descriptor the box is an
// [M or N = 128][K = 64] operand chunk; read with an MN-major descriptor the same bytes are a [K = 128 rows][MN = 64 columns]
// chunk (atoms of 8 rows = 1024 B along K, 64-column blocks one box apart along MN) - which is what dgrad (W[k][n]) and wgrad
// (reduction over rows) need, with no transposed copies anywhere.  Out-of-range rows / columns are zero-filled by TMA.
//
// Warp roles follow mlp_tc.cu's gemm_pair_kernel: warps 0, 2, 3 B-operand producers, 12-15 A-operand producers (one thread's
// copies complete serially, separate warps overlap), warp 1 MMA issuer (leader CTA) + TMEM allocation, warps 4-11 epilogue;
// mbarrier ring shared by the A and B stages, accumulator double-buffered in TMEM so the epilogue of unit u overlaps the
// MMAs of unit u + 1.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "tc_ptx.cuh"

namespace hos {

constexpr int kG2Threads = 512;
constexpr int kG2EpiWarp0 = 4, kG2EpiWarps = 8, kG2AWarp0 = 12;
constexpr int kG2BProducers = 3, kG2AProducers = 4;
constexpr int kG2MaxStages = 6;
constexpr int kG2Bars = 2 * kG2MaxStages + 4 + 4;      // + [3] mask tiles landed (+ pad)

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}

#ifdef HOS_G2_DEBUG
__device__ int* g_g2_dbg = nullptr;       // host-mapped record buffer: [0] = count, then 6 ints per record
template <int kSleepNs>
__device__ __forceinline__ void g2_wait(uint64_t* bar, uint32_t parity, int tag, int x, int y) {
  const uint32_t a = smem_u32(bar);
  uint32_t polls = 0;
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (kSleepNs > 0) asm volatile("nanosleep.u32 %0;" ::"r"((uint32_t)kSleepNs));
    if (++polls > (1u << 20)) {
      if ((threadIdx.x & 31) == 0 && g_g2_dbg) {
        const int slot = atomicAdd(g_g2_dbg, 1);
        if (slot < 100) {
          int* r = g_g2_dbg + 1 + 6 * slot;
          r[0] = tag; r[1] = blockIdx.x; r[2] = threadIdx.x >> 5; r[3] = x; r[4] = y; r[5] = (int)parity;
        }
        __threadfence_system();
      }
      __nanosleep(2000000);
      __trap();
    }
  }
}
#define G2_WAIT(ns, bar, par, tag, x, y) g2_wait<ns>(bar, par, tag, x, y)
#else
#define G2_WAIT(ns, bar, par, tag, x, y) mbar_wait_guard<ns>(bar, par)
#endif

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1),
               "r"(smem_u32(src))
               : "memory");
}

struct G2Args {
  __half* y_hi;
  __half* y_lo;
  float* y_f32;
  const float* bias;
  const float* rowbias;
  const __half* mask;
  const float* head_w;      // [hn][n] fp32
  const float* head_b;      // [hn] or null
  float* head_out;          // [rows][hn]
  float head_shift;
  int hn, head_post;
  int64_t rows;
  int rowbias_div;
  int ldy, ldy32, ld_mask;
  int ntiles;
  int n;            // valid output columns
  int n_blk;        // output columns per unit: 128 or 256
  int nblk;         // units per tile pair
  int kb0, kb1;     // 64-wide K chunks read from A0 / A1
  int relu, split, b_mn;
  int stages;
  int group;          // operand chunks per MMA issue group (non-split): the ring keeps stages - group chunks of prefetch
  int nbuf_out;     // output staging buffers (TMA stores in flight + 1)
  int bias_smem;    // bias of the current 256-column block staged in shared memory
  int dbg;
};

struct G2Maps {
  CUtensorMap a0[2], a1[2], b0[2], b1[2], y[2];     // [hi, lo]
  CUtensorMap mask;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kG2Threads, 1)
gemm_tma_kernel(const __grid_constant__ G2Args args, const __grid_constant__ G2Maps maps) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int planes = args.split ? 2 : 1;
  const int b_bytes = args.n_blk * 64;                          // this CTA's half of a [n_blk x 64] weight chunk
  const int stage_bytes = planes * (kXChunkBytes + b_bytes);   // [A_hi][A_lo][B_hi][B_lo]
  const int S = args.stages;
  unsigned char* sRing = smem;
  unsigned char* sStage = sRing + (size_t)S * stage_bytes;  // [1 or 2 planes][16 KB] output staging for the TMA stores
  const int out_buf_bytes = (args.y_lo ? 2 : 1) * kXChunkBytes;
  float* s_headx = reinterpret_cast<float*>(sStage + (size_t)args.nbuf_out * out_buf_bytes);     // [128][4]
  float* sBias = s_headx + kTileM * 4;                      // [256] bias of the current column block (bias_smem)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + (args.bias_smem ? 256 : 0));
  uint64_t* bar_full = bars;                                // leader: A + B of both CTAs landed (2 tx arrivals + 1 relay); peer: its own two
  uint64_t* bar_empty = bar_full + kG2MaxStages;            // multicast commit: stage consumed
  uint64_t* bar_tfull = bar_empty + kG2MaxStages;           // [2] multicast commit: accumulator buffer complete
  uint64_t* bar_tempty = bar_tfull + 2;                     // [2] leader only: drained by the 16 epilogue warps of the pair
  uint64_t* bar_mask = bar_tempty + 2;                      // [3] ReLU-mask tile landed in output staging buffer b (data gradient)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + kG2Bars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int n_groups = (args.ntiles + 1) >> 1;              // 2 tiles per group: tile = 2 g + rank
  const int KB = args.kb0 + args.kb1;
  const int nblk = args.nblk;

  if (args.bias_smem && threadIdx.x < 256) sBias[threadIdx.x] = (args.bias && (int)threadIdx.x < args.n) ? args.bias[threadIdx.x] : 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kG2MaxStages; ++s) { mbar_init(&bar_full[s], rank == 0 ? 3u : 2u); mbar_init(&bar_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bar_tfull[s], 1); mbar_init(&bar_tempty[s], 2 * kG2EpiWarps); }
    for (int s = 0; s < 3; ++s) mbar_init(&bar_mask[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const bool b_producer = warp == 0 || warp == 2 || warp == 3;
  const bool a_producer = warp >= kG2AWarp0;
  if (b_producer || a_producer) {
    // ===================== producers: chunk i of the flat stream -> producer i % P, ring stage i % S =====================
    if (lane == 0) {
      // A parity wait only distinguishes the current phase from the previous one: a producer may revisit a stage only when
      // that stage is at most one use behind.  Its previous chunk (P chunks back) waited for the chunk S before it, and the
      // MMAs retire in order, so that holds iff P <= S - with a 3-deep ring (split precision) only 3 producers may run.
      const uint32_t Pmax = b_producer ? kG2BProducers : kG2AProducers;
      const uint32_t P = Pmax < (uint32_t)S ? Pmax : (uint32_t)S;
      const uint32_t p = b_producer ? (warp == 0 ? 0u : (uint32_t)(warp - 1)) : (uint32_t)(warp - kG2AWarp0);
      const int half_n = args.n_blk >> 1;
      uint32_t ci = 0;
      for (int g = cluster; g < n_groups; g += n_clusters) {
        const int tile = 2 * g + (int)rank;                 // a tile past the end is all out-of-range rows: TMA zero-fills
        for (int j = 0; j < nblk; ++j) {
          for (int kb = 0; kb < KB; ++kb, ++ci) {
            if (ci % P != p) continue;
            const uint32_t st = ci % (uint32_t)S, use = ci / (uint32_t)S;
            if (args.dbg & 2) mbar_wait_guard<0>(&bar_empty[st], (use & 1) ^ 1);
            else G2_WAIT(100, &bar_empty[st], (use & 1) ^ 1, 1, (int)ci, (int)st);
            unsigned char* stage = sRing + (size_t)st * stage_bytes;
            const bool first = kb < args.kb0;
            const int kc = (first ? kb : kb - args.kb0) * kKB;
            if (!b_producer) {
              mbar_expect_tx(&bar_full[st], (uint32_t)(planes * kXChunkBytes));
              for (int pl = 0; pl < planes; ++pl)
                tma_load_2d(stage + pl * kXChunkBytes, first ? &maps.a0[pl] : &maps.a1[pl], kc, tile * kTileM, &bar_full[st]);
            } else {
              mbar_expect_tx(&bar_full[st], (uint32_t)(planes * b_bytes));
              unsigned char* sb = stage + planes * kXChunkBytes;
              const int n0 = j * args.n_blk + (int)rank * half_n;
              for (int pl = 0; pl < planes; ++pl) {
                const CUtensorMap* mp = first ? &maps.b0[pl] : &maps.b1[pl];
                if (!args.b_mn) {
                  tma_load_2d(sb + pl * b_bytes, mp, kc, n0, &bar_full[st]);                 // box [half_n rows x 64 k]
                } else {
                  for (int blk = 0; blk < half_n / 64; ++blk)                               // boxes [64 k rows x 64 n columns]
                    tma_load_2d(sb + pl * b_bytes + blk * 8192, mp, n0 + blk * 64, kc, &bar_full[st]);
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1 && rank != 0) {
    // ===================== peer CTA: completion watcher =====================
    // tells the leader "both of my operands of stage s have landed", stage by stage - the producers only issue copies and never
    // wait for one to arrive (with the relay in the producer threads each moved one chunk per L2 / HBM round trip)
    if (lane == 0) {
      const uint32_t full_leader = mapa_u32(smem_u32(&bar_full[0]), 0);
      uint32_t st = 0, par = 0;
      for (int g = cluster; g < n_groups; g += n_clusters) {
        const int nch = nblk * KB;
        for (int i = 0; i < nch; ++i) {
          mbar_wait_guard<40>(&bar_full[st], par);
          mbar_arrive_remote(full_leader + 8u * st);
          if (++st == (uint32_t)S) { st = 0; par ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: warp-uniform loop, tcgen05 instructions under elect.sync =====================
    if (rank == 0) {
      uint32_t u = 0, st = 0, par = 0;
      const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);
      const uint32_t tfull0 = smem_u32(&bar_tfull[0]);
      const uint32_t ring16 = (smem_u32(sRing) >> 4) & 0x3FFF;
      const uint32_t stage16 = (uint32_t)stage_bytes >> 4;
      const uint32_t idesc = umma_idesc_f16(args.n_blk, 2 * kTileM) | (args.b_mn ? (1u << 16) : 0u);
      const uint64_t adesc_hi = umma_desc(0);
      const uint64_t bdesc_hi = args.b_mn ? umma_desc_mn(0, 8192, 1024) : umma_desc(0);
      const uint32_t bstep = args.b_mn ? (2048u >> 4) : 2u;       // descriptor advance per K = 16
      const uint32_t a_lo16 = kXChunkBytes >> 4, b_off16 = (uint32_t)(planes * kXChunkBytes) >> 4, b_lo16 = (uint32_t)b_bytes >> 4;
      for (int g = cluster; g < n_groups; g += n_clusters) {
        for (int j = 0; j < nblk; ++j, ++u) {
          const uint32_t acc = tmem_base + (u & 1) * 256;
          if (u >= 2) mbar_wait_guard<0>(&bar_tempty[u & 1], ((u >> 1) - 1) & 1);     // buffer drained by unit u - 2
          tc_fence_after();
          int kb = 0;
          while (kb < KB) {
            // up to four chunks per poll (one parallel try_wait costs ~220 cycles; a non-split chunk is 512 cycles of tensor work)
            const int lim = args.split ? 1 : args.group;
            const int cnt = KB - kb < lim ? KB - kb : lim;
            uint32_t fb[4], fp[4];
            {
              uint32_t s_ = st, p_ = par;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                fb[i] = full0 + 8u * s_; fp[i] = p_;
                if (i + 1 < cnt) { if (++s_ == (uint32_t)S) { s_ = 0; p_ ^= 1; } }
              }
            }
#ifdef HOS_G2_DEBUG
            G2_WAIT(0, &bar_full[st], par, 3, kb, (int)st);
#endif
            mbar_wait4_spin(fb[0], fp[0], fb[1], fp[1], fb[2], fp[2], fb[3], fp[3]);
            tc_fence_after();
            const bool leader_lane = elect_one();
#pragma unroll 1
            for (int i = 0; i < cnt; ++i) {
              if (leader_lane) {
                const uint32_t s16 = ring16 + st * stage16;
                const uint64_t a_h = adesc_hi | (uint64_t)s16, a_l = adesc_hi | (uint64_t)(s16 + a_lo16);
                const uint64_t b_h = bdesc_hi | (uint64_t)(s16 + b_off16), b_l = bdesc_hi | (uint64_t)(s16 + b_off16 + b_lo16);
#pragma unroll
                for (int k = 0; k < kKB / 16; ++k)
                  tc_mma_f16_pair(acc, a_h + 2 * k, b_h + bstep * k, idesc, ((kb + i) | k) != 0 ? 1u : 0u);
                if (args.split && !(args.dbg & 1)) {
#pragma unroll
                  for (int k = 0; k < kKB / 16; ++k) tc_mma_f16_pair(acc, a_l + 2 * k, b_h + bstep * k, idesc, 1u);
#pragma unroll
                  for (int k = 0; k < kKB / 16; ++k) tc_mma_f16_pair(acc, a_h + 2 * k, b_l + bstep * k, idesc, 1u);
                }
                tc_commit_pair_addr(empty0 + 8u * st);
                if (kb + i == KB - 1) tc_commit_pair_addr(tfull0 + 8u * (u & 1));
              }
              if (++st == (uint32_t)S) { st = 0; par ^= 1; }
            }
            __syncwarp();
            kb += cnt;
          }
        }
      }
    }
  } else if (warp >= kG2EpiWarp0 && warp < kG2EpiWarp0 + kG2EpiWarps) {
    // ===================== epilogue (8 warps) =====================
    // bias / per-ray bias / ReLU / mask on the fp32 accumulator; the fp16 result (hi plane, optionally the residual lo
    // plane) is staged in shared memory in the swizzled box layout and leaves through TMA stores (one [128 x 64] box per
    // plane and 64 output columns: full-line writes instead of 32 scattered 16-byte stores per warp instruction);
    // optional fp32 copy and an fp32 head (<= 4 outputs, e.g. the density layer) straight from the registers.
    const int q = warp & 3;
    const int ch = (warp - kG2EpiWarp0) >> 2;
    const int r = q * 32 + lane;
    const bool epi_leader = threadIdx.x == kG2EpiWarp0 * 32;
    const uint32_t tempty0 = mapa_u32(smem_u32(&bar_tempty[0]), 0);
    const uint32_t st_row = smem_u32(sStage) + 128u * (uint32_t)r;
    uint32_t st_off[4];
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) st_off[gq] = st_row + ((uint32_t)((4 * ch + gq) ^ (r & 7)) << 4);
    const bool has_head = args.hn > 0;
    uint32_t u = 0;
    bool stage_busy = false;
    int ob = 0;
    // Data gradient: the ReLU mask (the layer's saved fp16 input) arrives by TMA, one [128 x 64] tile per output chunk, straight
    // into the staging buffer the chunk's result will be written to (every thread reads exactly the 16-byte slots it then
    // overwrites); the tile of chunk i + 1 is requested while chunk i is processed.  (Per-thread global loads of the mask -
    // 32 different lines per warp instruction - kept the L1 tag stage busy for twice the chunk's tensor time.)
    const int nchunk = args.n_blk / 64;
    uint32_t mpar = 0;                                      // per staging buffer: parity of the next mask phase
    auto next_chunk = [&](int& g2, int& j2, int& c2) {      // advance (g2, j2, c2) to the next chunk that stages output
      ++c2;
      if (c2 >= nchunk || j2 * args.n_blk + c2 * 64 >= args.n) {
        c2 = 0;
        if (++j2 >= nblk) { j2 = 0; g2 += n_clusters; }
      }
      return g2 < n_groups;
    };
    auto request_mask = [&](int g2, int j2, int c2, int buf) {
      mbar_expect_tx(&bar_mask[buf], (uint32_t)kXChunkBytes);
      tma_load_2d(sStage + (size_t)buf * out_buf_bytes, &maps.mask, j2 * args.n_blk + c2 * 64, (2 * g2 + (int)rank) * kTileM, &bar_mask[buf]);
    };
    if (args.mask && epi_leader && cluster < n_groups) request_mask(cluster, 0, 0, 0);
    for (int g = cluster; g < n_groups; g += n_clusters) {
      const int tile = 2 * g + (int)rank;
      const int64_t row = (int64_t)tile * kTileM + r;
      const bool row_ok = row < args.rows, tile_ok = tile < args.ntiles;
      float hacc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < nblk; ++j, ++u) {
        const uint32_t acc = tmem_base + (u & 1) * 256 + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32);
        if (args.bias_smem && nblk > 1) {       // several column blocks: restage this block's bias (block 0 was staged at start-up)
          asm volatile("bar.sync 1, %0;" ::"n"(kG2EpiWarps * 32) : "memory");
          const int e = threadIdx.x - kG2EpiWarp0 * 32, col = j * args.n_blk + e;
          sBias[e] = (args.bias && col < args.n) ? __ldg(args.bias + col) : 0.f;
          asm volatile("bar.sync 1, %0;" ::"n"(kG2EpiWarps * 32) : "memory");
        }
        G2_WAIT(20, &bar_tfull[u & 1], (u >> 1) & 1, 4, (int)u, 0);
        tc_fence_after();
        for (int c = 0; c < args.n_blk / 64; ++c) {
          const int nc0 = j * args.n_blk + c * 64;                    // first column of this 64-column chunk
          const int n0 = nc0 + ch * 32;                               // first of this thread's 32 output columns
          uint32_t v[32];
          tmem_ld32_nowait(acc + (uint32_t)(c * 64), v);
          tmem_wait_ld();
          if (nc0 >= args.n) continue;                                // uniform over the CTA: whole chunk beyond the matrix
          const bool col_ok = n0 < args.n;
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          if (args.bias_smem) {
            // (global bias loads in this loop cost a quarter of the kernel: 8 dependent L1 round trips per chunk per warp)
            const uint32_t sb = smem_u32(sBias) + 4u * (uint32_t)(c * 64 + ch * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 bb = lds128(sb + 16u * i);
              f[4 * i + 0] += bb.x; f[4 * i + 1] += bb.y; f[4 * i + 2] += bb.z; f[4 * i + 3] += bb.w;
            }
          } else if (args.bias && col_ok) {
            const float4* b4 = reinterpret_cast<const float4*>(args.bias + n0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (n0 + i * 4 >= args.n) break;
              const float4 bb = __ldg(b4 + i);
              f[4 * i + 0] += bb.x; f[4 * i + 1] += bb.y; f[4 * i + 2] += bb.z; f[4 * i + 3] += bb.w;
            }
          }
          if (args.rowbias && row_ok && col_ok) {           // per-ray term (view-direction encoding), fp32
            const float4* rb = reinterpret_cast<const float4*>(args.rowbias + (row / args.rowbias_div) * args.n + n0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (n0 + i * 4 >= args.n) break;
              const float4 bb = __ldg(rb + i);
              f[4 * i + 0] += bb.x; f[4 * i + 1] += bb.y; f[4 * i + 2] += bb.z; f[4 * i + 3] += bb.w;
            }
          }
          if (args.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
          }
          const uint32_t ob_off = (uint32_t)ob * (uint32_t)out_buf_bytes;
          if (args.mask) {                                  // dgrad through a ReLU: keep where the saved activation is positive
            if (epi_leader) {
              int g2 = g, j2 = j, c2 = c;
              if (next_chunk(g2, j2, c2)) {
                if (stage_busy) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // buffer ob + 1: store of chunk i - 2 has read it
                request_mask(g2, j2, c2, ob + 1 == args.nbuf_out ? 0 : ob + 1);
              }
            }
            mbar_wait_guard<20>(&bar_mask[ob], (mpar >> ob) & 1u);
            mpar ^= 1u << ob;
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              const float4 mv4 = lds128(st_off[gq] + ob_off);
              const uint32_t w[4] = {__float_as_uint(mv4.x), __float_as_uint(mv4.y), __float_as_uint(mv4.z), __float_as_uint(mv4.w)};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const __half2 h2 = *reinterpret_cast<const __half2*>(&w[i]);
                if (!(__low2float(h2) > 0.f)) f[gq * 8 + 2 * i] = 0.f;
                if (!(__high2float(h2) > 0.f)) f[gq * 8 + 2 * i + 1] = 0.f;
              }
            }
          }
          if (args.y_hi) {
            if (!args.mask) {
              // staging buffer `ob` is free once the TMA stores issued nbuf_out chunks ago have read it
              if (stage_busy && epi_leader) {
                if (args.nbuf_out >= 3) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
                else if (args.nbuf_out == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              }
              asm volatile("bar.sync 1, %0;" ::"n"(kG2EpiWarps * 32) : "memory");
            }
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              uint32_t ph[4], pl[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float x0 = f[gq * 8 + 2 * i], x1 = f[gq * 8 + 2 * i + 1];
                const __half2 h2 = __floats2half2_rn(x0, x1);
                ph[i] = *reinterpret_cast<const uint32_t*>(&h2);
                const __half2 l2 = __floats2half2_rn(x0 - __low2float(h2), x1 - __high2float(h2));
                pl[i] = *reinterpret_cast<const uint32_t*>(&l2);
              }
              sts128(st_off[gq] + ob_off, ph[0], ph[1], ph[2], ph[3]);
              if (args.y_lo) sts128(st_off[gq] + ob_off + kXChunkBytes, pl[0], pl[1], pl[2], pl[3]);
            }
            fence_proxy_async();                            // generic-proxy stores -> visible to the TMA engine
            asm volatile("bar.sync 1, %0;" ::"n"(kG2EpiWarps * 32) : "memory");
            if (epi_leader && tile_ok) {
              tma_store_2d(&maps.y[0], sStage + ob_off, nc0, tile * kTileM);
              if (args.y_lo) tma_store_2d(&maps.y[1], sStage + ob_off + kXChunkBytes, nc0, tile * kTileM);
            }
            if (epi_leader) asm volatile("cp.async.bulk.commit_group;" ::: "memory");     // one group per chunk, even when empty
            stage_busy = true;
            if (++ob == args.nbuf_out) ob = 0;
          }
          if (args.y_f32 && row_ok && col_ok) {
            float4* d32 = reinterpret_cast<float4*>(args.y_f32 + row * args.ldy32 + n0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (n0 + i * 4 >= args.n) break;
              d32[i] = make_float4(f[4 * i + 0], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
            }
          }
          if (has_head && col_ok) {
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              if (h < args.hn) {
                const float4* w4 = reinterpret_cast<const float4*>(args.head_w + (size_t)h * args.n + n0);
                float a = hacc[h];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  if (n0 + i * 4 >= args.n) break;
                  const float4 w = __ldg(w4 + i);
                  a = fmaf(f[4 * i + 0], w.x, a); a = fmaf(f[4 * i + 1], w.y, a);
                  a = fmaf(f[4 * i + 2], w.z, a); a = fmaf(f[4 * i + 3], w.w, a);
                }
                hacc[h] = a;
              }
            }
          }
        }
        tc_fence_before();          // TMEM loads ordered before the arrive
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(tempty0 + 8u * (u & 1));
      }
      if (has_head) {                           // combine the two column halves of this tile's rows, then post-process
        if (ch == 1) sts128(smem_u32(s_headx) + 16u * r, __float_as_uint(hacc[0]), __float_as_uint(hacc[1]), __float_as_uint(hacc[2]),
                            __float_as_uint(hacc[3]));
        asm volatile("bar.sync 1, %0;" ::"n"(kG2EpiWarps * 32) : "memory");
        if (ch == 0 && row_ok) {
          const float4 o4 = lds128(smem_u32(s_headx) + 16u * r);
          hacc[0] += o4.x; hacc[1] += o4.y; hacc[2] += o4.z; hacc[3] += o4.w;
          float* o = args.head_out + row * args.hn;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            if (h >= args.hn) break;
            float x = hacc[h] + (args.head_b ? __ldg(args.head_b + h) : 0.f);
            if (args.head_post == 1) {
              const float z = x + args.head_shift;
              x = z > 20.f ? z : log1pf(expf(z));
            } else if (args.head_post == 2) {
              x = (1.f / (1.f + expf(-x))) * (1.f + 2.f * args.head_shift) - args.head_shift;
            } else if (args.head_post == 4) {
              x = (h < 3) ? 1.f / (1.f + expf(-x)) : fmaxf(x, 0.f);
            }
            o[h] = x;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kG2EpiWarps * 32) : "memory");
      }
    }
    if (stage_busy && epi_leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// ============================================================================ weight gradient
// out[i][j] (+)= sum_rows P[row, i] Q[row, j],  i < 256 (P columns, M of the pair MMA), j < nq <= 512 (Q columns).
// Cluster c reduces row tiles c, c + C, ...: per 128-row tile both CTAs read the SAME rows - CTA r the P columns
// [128 r, 128 r + 128) (its half of M) and the Q columns [nq/2 r, ...) of every 256-column accumulator block (its half of N) -
// and issue 8 K = 16 steps; the fp32 accumulators stay in TMEM until the cluster has seen all of its tiles, then the
// epilogue adds them to the global gradient with vector atomics.  Columns beyond the matrices are zero-filled by TMA.
constexpr int kWgMaxStages = 3;           // ring depth: 3 stages of 64 KB (one accumulator block) or 2 of 96 KB (two)

struct WgArgs {
  float* out;
  float* colsum;      // optional [m]: += column sums of P (the bias gradient when P is dL/dZ), fp32 atomics
  int64_t rows;
  int ld_out;
  int ntiles;
  int m, nq;          // valid P / Q columns of the whole product; blockIdx.y walks its 256 x 512 output blocks
  int nqblocks;       // 512-column blocks of Q (blockIdx.y = mblock * nqblocks + qblock)
  int transpose_out;  // 0: out[i * ld + j], 1: out[j * ld + i]
};
struct WgMaps {
  CUtensorMap p, q;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kG2Threads, 1)
wgrad_tma_kernel(const __grid_constant__ WgArgs args, const __grid_constant__ WgMaps maps) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // this cluster's output block: P columns [m0, m0 + 256) x Q columns [q0, q0 + 512)
  const int m0 = ((int)blockIdx.y / args.nqblocks) * 256, q0 = ((int)blockIdx.y % args.nqblocks) * 512;
  const int m_here = args.m - m0 < 256 ? args.m - m0 : 256, nq_here = args.nq - q0 < 512 ? args.nq - q0 : 512;
  const int nacc = nq_here > 256 ? 2 : 1;                     // 256-column accumulator blocks
  float* const out = args.out + (args.transpose_out ? (size_t)q0 * args.ld_out + m0 : (size_t)m0 * args.ld_out + q0);
  float* const colsum = (args.colsum && q0 == 0) ? args.colsum + m0 : nullptr;
  const int p_bytes = 2 * kXChunkBytes;                       // this CTA's 128 P columns = two [128 x 64] boxes
  const int q_bytes = nacc * 2 * kXChunkBytes;                // 128 Q columns per accumulator block
  const int stage_bytes = p_bytes + q_bytes;
  unsigned char* sRing = smem;
  const uint32_t kWgStages = nacc == 1 ? 3u : 2u;             // 3 x 64 KB or 2 x 96 KB: the same 192 KB either way
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + (size_t)kWgStages * stage_bytes);
  uint64_t* bar_full = bars;                                  // [kWgMaxStages]
  uint64_t* bar_empty = bar_full + kWgMaxStages;              // [kWgMaxStages]
  uint64_t* bar_done = bar_empty + kWgMaxStages;              // [1] all MMAs of this cluster retired
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    // a stage is free when its MMAs have retired (multicast commit) and, with column sums requested, when the 8 epilogue
    // warps of this CTA have read the P tile out of it
    for (int s = 0; s < kWgMaxStages; ++s) {
      mbar_init(&bar_full[s], rank == 0 ? 3u : 2u);         // leader: own P + own Q + ONE relay from the peer's watcher warp
      mbar_init(&bar_empty[s], colsum ? 1u + kG2EpiWarps : 1u);
    }
    mbar_init(&bar_done[0], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  int my_tiles = 0;
  for (int t = cluster; t < args.ntiles; t += n_clusters) ++my_tiles;

  if (warp == 0 || warp == 2) {
    // ===================== producers: warp 0 loads P, warp 2 loads Q =====================
    if (lane == 0) {
      uint32_t ci = 0;
      for (int t = cluster; t < args.ntiles; t += n_clusters, ++ci) {
        const uint32_t st = ci % kWgStages, use = ci / kWgStages;
        mbar_wait_guard<100>(&bar_empty[st], (use & 1) ^ 1);
        unsigned char* stage = sRing + (size_t)st * stage_bytes;
        if (warp == 0) {
          mbar_expect_tx(&bar_full[st], (uint32_t)p_bytes);
          for (int b = 0; b < 2; ++b)
            tma_load_2d(stage + b * kXChunkBytes, &maps.p, m0 + (int)rank * 128 + b * 64, t * kTileM, &bar_full[st]);
        } else {
          mbar_expect_tx(&bar_full[st], (uint32_t)q_bytes);
          for (int a = 0; a < nacc; ++a)
            for (int b = 0; b < 2; ++b)
              tma_load_2d(stage + p_bytes + (a * 2 + b) * kXChunkBytes, &maps.q, q0 + a * 256 + (int)rank * 128 + b * 64, t * kTileM,
                          &bar_full[st]);
        }
      }
    }
  } else if (warp == 3) {
    // ===================== peer CTA: completion watcher =====================
    // tells the leader "both of my operands of stage s have landed".  A separate warp, so that the producers above never wait
    // for a copy to ARRIVE before issuing the next one (with the relay in the producer thread only one stage per CTA was in
    // flight and the kernel ran at the latency of one HBM round trip per tile: 110-140 us per 0.54 GB).
    if (rank != 0 && lane == 0) {
      uint32_t st = 0, par = 0;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait_guard<40>(&bar_full[st], par);
        mbar_arrive_remote(mapa_u32(smem_u32(&bar_full[st]), 0));
        if (++st == kWgStages) { st = 0; par ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && my_tiles > 0) {
      uint32_t st = 0, par = 0;
      const uint32_t empty0 = smem_u32(&bar_empty[0]);
      const uint32_t ring16 = (smem_u32(sRing) >> 4) & 0x3FFF;
      const uint32_t stage16 = (uint32_t)stage_bytes >> 4;
      const uint32_t idesc = umma_idesc_f16_mn(256, 2 * kTileM);
      const uint64_t desc_hi = umma_desc_mn(0, kXChunkBytes, 1024);        // 64-column blocks one box apart, 8-row atoms 1 KB apart
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait_guard<0>(&bar_full[st], par);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t s16 = ring16 + st * stage16;
          const uint64_t pd = desc_hi | (uint64_t)s16;
#pragma unroll 1
          for (int a = 0; a < nacc; ++a) {
            const uint64_t qd = desc_hi | (uint64_t)(s16 + (uint32_t)((p_bytes + a * 2 * kXChunkBytes) >> 4));
#pragma unroll
            for (int k = 0; k < kTileM / 16; ++k)          // 16 rows per step = two 8-row atoms = 2048 B
              tc_mma_f16_pair(tmem_base + a * 256, pd + 128u * k, qd + 128u * k, idesc, (it | k) != 0 ? 1u : 0u);
          }
          tc_commit_pair_addr(empty0 + 8u * st);
          if (it == my_tiles - 1) tc_commit_pair_addr(smem_u32(&bar_done[0]));
        }
        __syncwarp();
        if (++st == kWgStages) { st = 0; par ^= 1; }
      }
    }
  } else if (warp >= kG2EpiWarp0 && warp < kG2EpiWarp0 + kG2EpiWarps && my_tiles > 0) {
    // ===================== epilogue: this CTA's 128 accumulator rows (P columns) -> atomics =====================
    const int q = warp & 3;
    const int ch = (warp - kG2EpiWarp0) >> 2;
    const int i = (int)rank * 128 + q * 32 + lane;            // P column = output row
    if (colsum) {
      // While the tensor core works, these warps add up the columns of every P tile straight from the staged (swizzled)
      // boxes: thread -> column e & 127 of this CTA's 128, rows [64 (e >> 7), +64).  The bias gradient costs no extra pass.
      const int e = threadIdx.x - kG2EpiWarp0 * 32;
      const int col = e & 127, r0 = (e >> 7) * 64;
      const uint32_t cbase = (uint32_t)(col >> 6) * kXChunkBytes + (uint32_t)(col & 7) * 2u;
      const uint32_t cg = (uint32_t)((col & 63) >> 3);
      float acc = 0.f;
      uint32_t st = 0, par = 0;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait_guard<40>(&bar_full[st], par);
        const uint32_t sbase = smem_u32(sRing + (size_t)st * stage_bytes) + cbase;
#pragma unroll 8
        for (int rr = 0; rr < 64; ++rr) {
          const uint32_t r = (uint32_t)(r0 + rr);
          unsigned short hv;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hv) : "r"(sbase + r * 128u + ((cg ^ (r & 7u)) << 4)));
          acc += __half2float(__ushort_as_half(hv));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[st]);
        if (++st == kWgStages) { st = 0; par ^= 1; }
      }
      const int pc = (int)rank * 128 + col;
      if (pc < m_here) atomicAdd(colsum + pc, acc);
    }
    mbar_wait_guard<200>(&bar_done[0], 0);
    tc_fence_after();
    // Every MMA has retired, so the operand ring is free: each warp transposes its 32 x 32 accumulator blocks through a
    // private 32 x 33 float patch of it and adds them to the global gradient with COALESCED atomics (a warp instruction
    // covers 32 consecutive floats of one output row = one 128-byte line; thread-per-row atomics touched 32 lines each
    // and made this epilogue - not the reduction over the rows - the cost of a small-batch launch).
    asm volatile("bar.sync 1, %0;" ::"n"(kG2EpiWarps * 32) : "memory");     // every warp is done reading P tiles (column sums)
    float* patch = reinterpret_cast<float*>(sRing) + (size_t)(warp - kG2EpiWarp0) * (32 * 33);
    const int i_base = (int)rank * 128 + q * 32;              // first output row of this warp
    for (int a = 0; a < nacc; ++a) {
      for (int c = 0; c < 4; ++c) {
        const int j0 = a * 256 + c * 64 + ch * 32;
        uint32_t v[32];
        tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 256 + c * 64 + ch * 32), v);
        tmem_wait_ld();
        if (i_base >= m_here || j0 >= nq_here) continue;       // warp-uniform
        if (!args.transpose_out) {
#pragma unroll
          for (int x = 0; x < 32; ++x) patch[lane * 33 + x] = __uint_as_float(v[x]);
          __syncwarp();
          const bool col_ok = j0 + lane < nq_here;
          for (int rr = 0; rr < 32; ++rr) {
            if (i_base + rr >= m_here) break;
            if (col_ok) atomicAdd(out + (size_t)(i_base + rr) * args.ld_out + j0 + lane, patch[rr * 33 + lane]);
          }
          __syncwarp();
        } else if (i < m_here) {                               // out[j][i]: lanes are consecutive i already
#pragma unroll
          for (int x = 0; x < 32; ++x)
            if (j0 + x < nq_here) atomicAdd(out + (size_t)(j0 + x) * args.ld_out + i, __uint_as_float(v[x]));
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// ============================================================================ small SIMT companions
// out[c] += sum_rows g[row] * X[row, c]   (g == null: plain column sums = bias gradients); hn weight vectors at once:
// out[h * ld_out + c] += sum_rows g[row * hn + h] * X[row, c].  X fp16 row-major.  One CTA per slab of rows, thread = column pair.
__global__ void __launch_bounds__(256)
colsum_f16_kernel(const __half* __restrict__ X, int64_t rows, int n, int ld, const float* __restrict__ g, int hn,
                  float* __restrict__ out, int ld_out, int rows_per_cta) {
  // thread -> (column pair, row phase): the 256 threads cover n / 2 column pairs x `phases` interleaved row sets, four
  // independent rows in flight per thread (the one-row-at-a-time version ran at 1 TB/s)
  const int pairs = n >> 1;
  const int phases = pairs >= 256 ? 1 : 256 / pairs;
  const int cp = threadIdx.x % pairs, ph = threadIdx.x / pairs;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  if (ph >= phases) return;
  for (int c = 2 * cp; c < n; c += 2 * 256) {
    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t r = r0 + ph; r < r1; r += 4 * phases) {
      __half2 h2[4];
      float gv[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t rr = r + (int64_t)u * phases;
        const bool ok = rr < r1;
        h2[u] = ok ? *reinterpret_cast<const __half2*>(X + rr * ld + c) : __floats2half2_rn(0.f, 0.f);
#pragma unroll
        for (int h = 0; h < 4; ++h) gv[u][h] = !g ? (h == 0 ? 1.f : 0.f) : ((ok && h < hn) ? __ldg(g + rr * hn + h) : 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float x0 = __low2float(h2[u]), x1 = __high2float(h2[u]);
#pragma unroll
        for (int h = 0; h < 4; ++h) { a0[h] = fmaf(gv[u][h], x0, a0[h]); a1[h] = fmaf(gv[u][h], x1, a1[h]); }
      }
    }
    for (int h = 0; h < (g ? hn : 1); ++h) {
      atomicAdd(out + (size_t)h * ld_out + c, a0[h]);
      if (c + 1 < n) atomicAdd(out + (size_t)h * ld_out + c + 1, a1[h]);
    }
  }
}

// The same reduction with 16-byte loads: thread -> (8-column group, row phase), a warp reads 512 contiguous bytes of a row,
// two rows in flight per thread; the CTA's partial sums meet in shared memory (atomics between the row phases only) and leave
// as one global atomic per output element.  Needs n % 8 == 0, n <= 2048, 16-byte aligned rows.  (The 4-byte-per-thread kernel
// above ran at 1.0 - 1.4 TB/s: 190 us per 524 288 x 256 head gradient.)
__global__ void __launch_bounds__(256)
colsum_f16_v8_kernel(const __half* __restrict__ X, int64_t rows, int n, int ld, const float* __restrict__ g, int hn,
                     float* __restrict__ out, int ld_out, int rows_per_cta) {
  extern __shared__ float s_sum[];                       // [hn][n]
  const int groups = n >> 3;
  const int phases = 256 / groups;
  const int cgp = threadIdx.x % groups, ph = threadIdx.x / groups;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  for (int i = threadIdx.x; i < hn * n; i += 256) s_sum[i] = 0.f;
  __syncthreads();
  if (ph < phases) {
    float acc[4][8];
#pragma unroll
    for (int h = 0; h < 4; ++h)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[h][c] = 0.f;
    const __half* col = X + 8 * cgp;
    for (int64_t r = r0 + ph; r < r1; r += 2 * phases) {
      const int64_t rb = r + phases;
      const bool okb = rb < r1;
      const uint4 va = __ldg(reinterpret_cast<const uint4*>(col + r * ld));
      const uint4 vb = okb ? __ldg(reinterpret_cast<const uint4*>(col + rb * ld)) : make_uint4(0u, 0u, 0u, 0u);
      float ga[4], gb[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        ga[h] = !g ? (h == 0 ? 1.f : 0.f) : (h < hn ? __ldg(g + r * hn + h) : 0.f);
        gb[h] = !g ? (h == 0 ? 1.f : 0.f) : ((okb && h < hn) ? __ldg(g + rb * hn + h) : 0.f);
      }
      const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __half2 a2 = *reinterpret_cast<const __half2*>(&wa[k]), b2 = *reinterpret_cast<const __half2*>(&wb[k]);
        const float a0 = __low2float(a2), a1 = __high2float(a2), b0 = __low2float(b2), b1 = __high2float(b2);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          acc[h][2 * k] = fmaf(ga[h], a0, acc[h][2 * k]);
          acc[h][2 * k + 1] = fmaf(ga[h], a1, acc[h][2 * k + 1]);
          acc[h][2 * k] = fmaf(gb[h], b0, acc[h][2 * k]);
          acc[h][2 * k + 1] = fmaf(gb[h], b1, acc[h][2 * k + 1]);
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 4; ++h)
      if (h < hn) {
#pragma unroll
        for (int c = 0; c < 8; ++c) atomicAdd(&s_sum[h * n + 8 * cgp + c], acc[h][c]);
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < hn * n; i += 256) {
    const int h = i / n, c = i - h * n;
    atomicAdd(out + (size_t)h * ld_out + c, s_sum[i]);
  }
}

// Y[row, c] = (sum_h g[row, h] W[h, c] (+ Y_add[row, c])) .* (mask[row, c] > 0)   as fp16: the data gradient of a head with
// <= 4 outputs (density / rgb / raw4 / offset heads), fused with the ReLU mask of the layer it reads.
__global__ void __launch_bounds__(256)
head_dgrad_kernel(const float* __restrict__ g, int hn, const float* __restrict__ W, int ldw, const __half* __restrict__ add,
                  int ld_add, const __half* __restrict__ mask, int ld_mask, int64_t rows, int n, __half* __restrict__ Y, int ldy) {
  const int vpr = n >> 3;                                    // 16-byte vectors (8 columns) per row
  const int64_t total = rows * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / vpr;
    const int c = 8 * (int)(i - row * vpr);
    float y[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 4; ++h)
      if (h < hn) {
        const float gv = __ldg(g + row * hn + h);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(W + (size_t)h * ldw + c));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(W + (size_t)h * ldw + c) + 1);
        y[0] = fmaf(gv, w0.x, y[0]); y[1] = fmaf(gv, w0.y, y[1]); y[2] = fmaf(gv, w0.z, y[2]); y[3] = fmaf(gv, w0.w, y[3]);
        y[4] = fmaf(gv, w1.x, y[4]); y[5] = fmaf(gv, w1.y, y[5]); y[6] = fmaf(gv, w1.z, y[6]); y[7] = fmaf(gv, w1.w, y[7]);
      }
    if (add) {
      const uint4 av = __ldg(reinterpret_cast<const uint4*>(add + row * ld_add + c));
      const uint32_t aw[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __half2 a2 = *reinterpret_cast<const __half2*>(&aw[k]);
        y[2 * k] += __low2float(a2); y[2 * k + 1] += __high2float(a2);
      }
    }
    if (mask) {
      const uint4 mv = __ldg(reinterpret_cast<const uint4*>(mask + row * ld_mask + c));
      const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __half2 m2 = *reinterpret_cast<const __half2*>(&mw[k]);
        if (!(__low2float(m2) > 0.f)) y[2 * k] = 0.f;
        if (!(__high2float(m2) > 0.f)) y[2 * k + 1] = 0.f;
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __half2 h2 = __floats2half2_rn(y[2 * k], y[2 * k + 1]);
      o[k] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    *reinterpret_cast<uint4*>(Y + row * ldy + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp16 row-major [rows, cols] with leading dimension ld (elements); box [box_rows x 64 columns], SWIZZLE_128B.
static int make_map(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return HOS_ERR_CUDA; }
  if (!base || (reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16 != 0 || rows < 1 || cols < 1) {
    set_error("tensor map: base must be 16-byte aligned and the row pitch a multiple of 16 bytes (base %p, ld %lld, rows %lld, cols %lld)",
              base, (long long)ld, (long long)rows, (long long)cols);
    return HOS_ERR_ARG;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return HOS_ERR_CUDA; }
  return HOS_OK;
}

static int max_clusters_for(const void* kernel, size_t smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kNumSMs);
  cfg.blockDim = dim3(kG2Threads);
  cfg.dynamicSmemBytes = smem;
  int nc = 0;
  if (cudaOccupancyMaxActiveClusters(&nc, kernel, &cfg) != cudaSuccess || nc < 1) {
    cudaGetLastError();
    nc = kNumSMs / 2;
  }
  return nc < kNumSMs / 2 ? nc : kNumSMs / 2;
}

}  // namespace hos

using namespace hos;

extern "C" {
#ifdef HOS_G2_DEBUG
int hos_g2_debug_buffer(void* p) {
  return cudaMemcpyToSymbol(g_g2_dbg, &p, sizeof(p)) == cudaSuccess ? 0 : -3;
}
#endif

int hos_gemm_tma(const hos_gemm_tma_desc* d, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(d && d->rows >= 0 && d->a0_hi && d->w0_hi && d->k0 >= 1 && d->n >= 8 && (d->n % 8) == 0,
              "hos_gemm_tma: need A0, W0, k0 >= 1 and n a positive multiple of 8");
  HOS_REQUIRE(d->mode == 0 || d->mode == 1, "hos_gemm_tma: mode 0 (NT forward) or 1 (NN dgrad)");
  HOS_REQUIRE(d->k1 == 0 || (d->a1_hi && d->w1_hi), "hos_gemm_tma: second input needs A1 and W1");
  const bool split = d->a0_lo != nullptr;
  HOS_REQUIRE(!split || (d->w0_lo && (d->k1 == 0 || (d->a1_lo && d->w1_lo))), "hos_gemm_tma: split precision needs every lo plane");
  HOS_REQUIRE(d->y_hi || d->y_f32 || d->hn > 0, "hos_gemm_tma: nothing to write");
  HOS_REQUIRE(!d->y_lo || d->y_hi, "hos_gemm_tma: y_lo needs y_hi");
  HOS_REQUIRE(!d->y_hi || ((d->ldy % 8) == 0 && (reinterpret_cast<uintptr_t>(d->y_hi) & 15) == 0), "hos_gemm_tma: y pitch / alignment");
  HOS_REQUIRE(!d->y_f32 || ((d->ldy32 % 4) == 0 && (reinterpret_cast<uintptr_t>(d->y_f32) & 15) == 0), "hos_gemm_tma: y_f32 pitch / alignment");
  HOS_REQUIRE(!d->mask || ((d->ld_mask % 8) == 0 && (reinterpret_cast<uintptr_t>(d->mask) & 15) == 0), "hos_gemm_tma: mask pitch / alignment");
  HOS_REQUIRE(!d->rowbias || ((d->n % 4) == 0 && (reinterpret_cast<uintptr_t>(d->rowbias) & 15) == 0), "hos_gemm_tma: rowbias alignment");
  HOS_REQUIRE(!d->bias || (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0, "hos_gemm_tma: bias must be 16-byte aligned");
  HOS_REQUIRE(d->hn >= 0 && d->hn <= 4 && (d->hn == 0 || (d->head_w && d->head_out && (d->n % 4) == 0 &&
                                                        (reinterpret_cast<uintptr_t>(d->head_w) & 15) == 0)),
              "hos_gemm_tma: head needs hn <= 4, head_w [hn, n] (16-byte aligned) and head_out");
  if (d->rows == 0) return HOS_OK;
  G2Args a;
  memset(&a, 0, sizeof(a));
  G2Maps maps;
  memset(&maps, 0, sizeof(maps));
  a.y_hi = (__half*)d->y_hi; a.y_lo = (__half*)d->y_lo; a.y_f32 = d->y_f32;
  a.bias = d->bias; a.mask = (const __half*)d->mask;
  a.rowbias = d->rowbias; a.rowbias_div = d->rowbias_div < 1 ? 1 : d->rowbias_div;
  a.head_w = d->head_w; a.head_b = d->head_b; a.head_out = d->head_out; a.head_shift = d->head_shift;
  a.hn = d->hn; a.head_post = d->head_post;
  a.rows = d->rows; a.ldy = d->ldy; a.ldy32 = d->ldy32; a.ld_mask = d->ld_mask;
  a.ntiles = (int)((d->rows + kTileM - 1) / kTileM);
  a.n = d->n;
  a.n_blk = d->n <= 128 ? 128 : 256;
  a.nblk = (d->n + a.n_blk - 1) / a.n_blk;
  a.kb0 = (d->k0 + kKB - 1) / kKB;
  a.kb1 = (d->k1 + kKB - 1) / kKB;
  a.relu = d->relu; a.split = split; a.b_mn = d->mode == 1;
  const int planes = split ? 2 : 1;
  const int stage_bytes = planes * (kXChunkBytes + a.n_blk * 64);
  // split precision: the 64 KB stages leave room for one staging buffer (its epilogue is far shorter than its MMAs);
  // single plane: three staging buffers so that two chunks' TMA stores stay in flight behind the epilogue
  a.nbuf_out = !d->y_hi ? 0 : (split ? 1 : 3);
  a.bias_smem = !split && d->bias != nullptr;
  const size_t fixed = 1024 + (size_t)a.nbuf_out * (d->y_lo ? 2 : 1) * kXChunkBytes + kTileM * 4 * sizeof(float) +
                       (a.bias_smem ? 1024 : 0) + kG2Bars * 8 + 64;
  int stages = (int)((227 * 1024 - fixed) / stage_bytes);
  if (stages > kG2MaxStages) stages = kG2MaxStages;
  HOS_REQUIRE(stages >= 2, "hos_gemm_tma: shared memory budget leaves %d pipeline stages", stages);
  { const char* e = getenv("HOS_G2_STAGES"); if (e) stages = atoi(e); }
  a.stages = stages;
  { const char* e = getenv("HOS_G2_GROUP"); a.group = e ? atoi(e) : 2; }
  if (a.group < 1) a.group = 1;
  if (a.group > 4) a.group = 4;
  if (a.group > stages - 1) a.group = stages - 1;
  { const char* e = getenv("HOS_G2_DBG"); a.dbg = e ? atoi(e) : 0; }
  const size_t smem = fixed + (size_t)stages * stage_bytes;
  const int half_n = a.n_blk / 2;
  int rc;
  for (int pl = 0; pl < planes; ++pl) {
    const void* a0 = pl ? d->a0_lo : d->a0_hi;
    const void* w0 = pl ? d->w0_lo : d->w0_hi;
    if ((rc = make_map(&maps.a0[pl], a0, d->rows, d->k0, d->lda0, kTileM)) != HOS_OK) return rc;
    // mode 0: W0 [n, k0] box [half_n rows x 64 k];  mode 1: W0 [k0, n] box [64 k rows x 64 n columns]
    if (d->mode == 0) rc = make_map(&maps.b0[pl], w0, d->n, d->k0, d->ldw0, half_n);
    else rc = make_map(&maps.b0[pl], w0, d->k0, d->n, d->ldw0, 64);
    if (rc != HOS_OK) return rc;
    if (d->k1 > 0) {
      const void* a1 = pl ? d->a1_lo : d->a1_hi;
      const void* w1 = pl ? d->w1_lo : d->w1_hi;
      if ((rc = make_map(&maps.a1[pl], a1, d->rows, d->k1, d->lda1, kTileM)) != HOS_OK) return rc;
      if (d->mode == 0) rc = make_map(&maps.b1[pl], w1, d->n, d->k1, d->ldw1, half_n);
      else rc = make_map(&maps.b1[pl], w1, d->k1, d->n, d->ldw1, 64);
      if (rc != HOS_OK) return rc;
    }
  }
  if (d->mask) {
    HOS_REQUIRE(d->y_hi && !split && a.nbuf_out >= 3, "hos_gemm_tma: a mask needs the single-plane fp16 output path");
    if ((rc = make_map(&maps.mask, d->mask, d->rows, d->n, d->ld_mask, kTileM)) != HOS_OK) return rc;
  }
  if (d->y_hi) {
    if ((rc = make_map(&maps.y[0], d->y_hi, d->rows, d->n, d->ldy, kTileM)) != HOS_OK) return rc;
    if (d->y_lo && (rc = make_map(&maps.y[1], d->y_lo, d->rows, d->n, d->ldy, kTileM)) != HOS_OK) return rc;
  }
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_dev != dev) {
    HOS_CUDA(cudaFuncSetAttribute(gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    HOS_CUDA(cudaFuncSetAttribute(wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    attr_dev = dev;
  }
  static thread_local int max_clusters = 0;
  if (!max_clusters) max_clusters = max_clusters_for((const void*)gemm_tma_kernel, 227 * 1024 - 1024);
  const int n_groups = (a.ntiles + 1) / 2;
  const int clusters = n_groups < max_clusters ? n_groups : max_clusters;
  gemm_tma_kernel<<<2 * clusters, kG2Threads, smem, (cudaStream_t)stream>>>(a, maps);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_wgrad_tma(const void* p, int m, int ldp, const void* q, int nq, int ldq, int64_t rows, float* out, int ld_out,
                  int transpose_out, float* colsum_p, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(p && q && out && m >= 1 && nq >= 1 && rows >= 0 && ld_out >= 1, "hos_wgrad_tma: need P [rows, m], Q [rows, nq] and out");
  if (rows == 0) return HOS_OK;
  WgArgs a;
  memset(&a, 0, sizeof(a));
  WgMaps maps;
  memset(&maps, 0, sizeof(maps));
  a.out = out; a.rows = rows; a.ld_out = ld_out; a.colsum = colsum_p;
  a.ntiles = (int)((rows + kTileM - 1) / kTileM);
  a.m = m; a.nq = nq; a.nqblocks = (nq + 511) / 512; a.transpose_out = transpose_out;
  const int nblocks = ((m + 255) / 256) * a.nqblocks;        // 256 x 512 output blocks, one grid row each
  HOS_REQUIRE(nblocks <= 65535, "hos_wgrad_tma: product too large (%d output blocks)", nblocks);
  int rc;
  if ((rc = make_map(&maps.p, p, rows, m, ldp, kTileM)) != HOS_OK) return rc;
  if ((rc = make_map(&maps.q, q, rows, nq, ldq, kTileM)) != HOS_OK) return rc;
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_dev != dev) {
    HOS_CUDA(cudaFuncSetAttribute(wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    attr_dev = dev;
  }
  const size_t smem = 1024 + (size_t)12 * kXChunkBytes + (2 * kWgMaxStages + 2) * 8 + 64;       // 3 x 64 KB = 2 x 96 KB of ring
  static thread_local int max_clusters = 0;
  if (!max_clusters) max_clusters = max_clusters_for((const void*)wgrad_tma_kernel, 227 * 1024 - 1024);
  // clusters per output block: >= 16 row tiles each before a cluster pays the atomic epilogue, and about one wave of the
  // machine over all the blocks (a 1024 x 1024 layer is 8 blocks: 9 clusters each instead of 8 launches of 32)
  int clusters = (a.ntiles + 15) / 16;
  const int cap = max_clusters / nblocks > 1 ? max_clusters / nblocks : 1;
  if (clusters > cap) clusters = cap;
  if (clusters < 1) clusters = 1;
  wgrad_tma_kernel<<<dim3(2 * clusters, nblocks, 1), kG2Threads, smem, (cudaStream_t)stream>>>(a, maps);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_colsum_f16(const void* x, int64_t rows, int n, int ld, const float* g, int hn, float* out, int ld_out, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(x && out && rows >= 0 && n >= 2 && (n % 2) == 0 && (ld % 2) == 0 && hn >= 0 && hn <= 4 && (g || hn <= 1),
              "hos_colsum_f16: X [rows, n even], hn <= 4");
  if (rows == 0) return HOS_OK;
  const int hn_eff = g ? hn : 1;
  if ((n % 8) == 0 && n <= 2048 && (ld % 8) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    // about four CTAs per SM in one wave, at least 256 rows each
    int64_t per = (rows + 148 * 4 - 1) / (148 * 4);
    if (per < 256) per = 256;
    const unsigned grid = (unsigned)((rows + per - 1) / per);
    colsum_f16_v8_kernel<<<grid, 256, (size_t)hn_eff * n * sizeof(float), (cudaStream_t)stream>>>((const __half*)x, rows, n, ld, g, hn_eff,
                                                                                               out, ld_out, (int)per);
    HOS_LAUNCH_CHECK();
    return HOS_OK;
  }
  const int rows_per_cta = 512;
  const unsigned grid = (unsigned)((rows + rows_per_cta - 1) / rows_per_cta);
  const int threads = 256;
  colsum_f16_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>((const __half*)x, rows, n, ld, g, hn_eff, out, ld_out, rows_per_cta);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_head_dgrad(const float* g, int hn, const float* W, int ldw, const void* add, int ld_add, const void* mask, int ld_mask,
                   int64_t rows, int n, void* y, int ldy, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(g && W && y && hn >= 1 && hn <= 4 && n >= 8 && (n % 8) == 0 && (ldy % 8) == 0 && (ldw % 4) == 0 &&
                  (!mask || (ld_mask % 8) == 0) && (!add || (ld_add % 8) == 0) &&
                  ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(add) |
                    reinterpret_cast<uintptr_t>(mask)) & 15) == 0,
              "hos_head_dgrad: hn <= 4, n and the fp16 pitches multiples of 8, ldw of 4, 16-byte aligned pointers");
  if (rows == 0) return HOS_OK;
  const int64_t total = rows * (n / 8);
  const unsigned grid = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  head_dgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, hn, W, ldw, (const __half*)add, ld_add, (const __half*)mask, ld_mask,
                                                          rows, n, (__half*)y, ldy);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

}  // extern "C"
