// Fused multi-layer MLP on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
// SURVEY 8a rows a9 (Prop/NeRF MLP), a17 (non-rigid MLP), a19 (canonical MLP).
//
// Two kernels share the program format, the packed weights and the fused IPE prologue:
//   mlp_pair_kernel  (default)  2-CTA clusters, tcgen05.mma.cta_group::2, deep weight ring, TMEM double-buffered
//                               accumulator, chunk-pipelined epilogue - see the comment above that kernel;
//   mlp_tc_kernel    (fallback) the single-CTA kernel described next.
//
// mlp_tc_kernel: one persistent CTA per SM walks 128-row tiles of the flattened (rays x samples) batch.
// For every tile the whole layer stack runs without touching HBM in between:
//
//   warp 0  (producer)  1-D bulk copies (cp.async.bulk -> UBLKCP) of pre-tiled, pre-swizzled
//                       weight chunks [N x 64] fp16 and, for layers that read the encoded
//                       input (first layer, skip layer), of the input-feature chunks
//                       [128 x 64] fp16 into a shared-memory ring, signalled by mbarriers.
//   warp 1  (MMA)       one lane issues tcgen05.mma (M=128, N=layer width, K=16, fp16 x fp16 ->
//                       fp32) with A = activations in shared memory (K-major, 128B swizzle),
//                       B = weight chunk, D = TMEM accumulator; tcgen05.commit releases ring
//                       stages and signals the epilogue.
//   warps 2-5 (epilogue) tcgen05.ld the accumulator (lane = row), + bias (+ per-ray bias),
//                       ReLU, pack to fp16 and store straight back into shared memory in the
//                       UMMA canonical layout as the next layer's A operand; small output
//                       heads (density / rgb / raw4 / xyz offset) are evaluated in fp32 from
//                       the registers and are the only thing written to HBM.
//
// HBM layout ("tiled fp16"): activations entering the kernel and all weights are stored as
// consecutive 16 KB (resp. N*128 B) images of exactly what the tensor core reads from shared
// memory - element (r, k) of a [rows x 64] K-block lives at  r*128 + (((k>>3) ^ (r&7))<<4) + (k&7)*2 -
// so a chunk is one contiguous bulk copy, no tensor map needed.
#include <string.h>

#include <vector>

#include "tc_ptx.cuh"

namespace hos {

constexpr int kMaxLayers = 12;
constexpr int kMaxHeads = 4;
constexpr int kStages = 3;       // weight ring depth
constexpr int kStagesX = 3;      // input-feature ring depth
constexpr int kMlpThreads = 512; // warp 0 producer, 1 MMA, (2-3 idle), 4-11 epilogue, 12-15 feature generators
constexpr int kEpiWarp0 = 4, kEpiWarps = 8, kFeatWarp0 = 12;
constexpr int kTmemCols = 512;   // two accumulator buffers: consecutive tiles alternate, so the first layer of
                                 // tile t+1 (input = features, no dependence on tile t) overlaps tile t's last read-out
constexpr int kIpeB = 21;        // geodesic basis directions (icosahedron, 2 subdivisions)
constexpr int kIpeDeg = 12;      // octaves 2^0 .. 2^11

struct LayerDev {
  uint32_t w_off;      // byte offset of the first weight chunk in the packed fp16 buffer
  uint32_t bias_off;   // float offset of bias[n] in the fp32 parameter block
  uint16_t n;          // output width (multiple of 16, <= 256)
  uint8_t kb_h;        // K-blocks read from the previous layer's activations
  uint8_t kb_x;        // K-blocks read from the streamed input features
  uint8_t relu;
  uint8_t rowbias;
  int8_t head;         // head evaluated on this layer's output, or -1
  uint8_t pad_;
};
struct HeadDev {
  uint32_t w_off;      // float offset of W[hn][n]
  uint32_t b_off;
  float shift;
  uint8_t hn;
  uint8_t post;        // 0 identity, 1 softplus(v+shift), 2 sigmoid*(1+2 shift)-shift, 3 v+add, 4 rgb sigmoid / sigma relu
  uint8_t slot;        // which output pointer
  uint8_t pad_;
};
struct MlpProgram {
  int n_layers;
  int n_heads;
  int kbx;             // K-blocks per row tile in the input feature buffer
  int kbh;             // K-blocks of the hidden activation buffer (width / 64)
  int n_max;           // widest layer
  int param_floats;
  int head_base;       // params[head_base ..) = head weights / biases (the part the cluster-pair kernel keeps on chip)
  LayerDev layers[kMaxLayers];
  HeadDev heads[kMaxHeads];
};

// Inputs of the fused integrated-positional-encoding prologue (S1 helper.py:242-302, 26-78):
// the feature warps turn ray intervals straight into fp16 A-operand chunks in shared memory.
struct IpeArgs {
  const float* tdist;    // [N, S+1]
  const float* rays_o;   // [N, 3]
  const float* rays_d;   // [N, 3]
  const float* radii;    // [N]
  int S;
  float basis[3 * kIpeB];
};

struct MlpArgs {
  const unsigned char* x_tiled;   // [ntiles][kbx][16 KB], or null when the IPE prologue is fused
  const unsigned char* w_packed;  // fp16 chunks
  const float* params;            // biases + head weights
  const float* rowbias;           // [rows / rowbias_div][n] or null
  const float* add;               // [rows][hn] for post 3, or null
  float* out[2];
  int64_t rows;
  int ntiles;
  int rowbias_div;
  int fused_ipe;                  // 0: features from x_tiled; 1: IPE prologue (IpeArgs); 2: Fourier prologue (pair kernel only):
                                  // the feature warps encode fx [rows,3] straight into the A-operand ring, one 64-column chunk
  const float* fx;                // mode 2: points [rows, 3]
  int f_nfreq, f_ident;           // mode 2: octaves 2^0 .. 2^(f_nfreq-1); identity columns first (fourier.py) or not (hannw_fourier.py)
  float f_hann[16];               // mode 2: per-octave window weights (all 1 for the plain embedder)
  long long* timeline;            // optional debug: per (tile iteration, layer) 4 clock64() stamps of CTA 0
  int duo;                        // pair kernel: two tile pairs in flight per cluster, units interleaved layer by layer (narrow programs)
};

// ----------------------------------------------------------------------------- fused IPE prologue
// One feature thread owns one row (sample) of the 128-row tile and produces its 504 features
// octave by octave.  Kernel column order is  col = (l*21 + j)*2 + {0: sin, 1: cos}  (the weight
// columns are permuted to match at upload time), so octave l lands in columns [42 l, 42 l + 42) and
// the K-blocks come out in order while the recurrences run along l:
//   sin/cos(2^l m): angle doubling, re-seeded from MUFU sin/cos after Cody-Waite reduction every 4
//                   octaves (error <= ~1e-5, below the fp16 resolution of the operand);
//   exp(-.5 4^l v): ex2.approx every 4th octave, e_{l+1} = e_l^4 in between.
//   (the sine of odd octaves comes out negated - three instructions per doubling instead of four - and the weight
//   columns carry the sign.)
// Per row and pass: ~210 MUFU + ~2.5 k FP instructions instead of 756 libm calls, and the features
// never exist outside shared memory.
// Geometry of one sample, shared by the three j-group passes of a feature pass.
struct IpeRowGeom {
  float x[3], d[3];        // pre-contraction mean, ray direction
  float a, bq, m, xd;      // contraction: z = a x, J = a I + bq x x^T ; m = |x|^2 ; xd = x.d
  float t_var, r_var, inv_dsq;
};

__device__ __forceinline__ void ipe_row_setup(const IpeArgs& A, int64_t row, bool row_ok, IpeRowGeom& G) {
  float o[3] = {0.f, 0.f, 0.f}, t0 = 1.f, t1 = 2.f, radius = 0.f;
  G.d[0] = 0.f; G.d[1] = 0.f; G.d[2] = 1.f;
  if (row_ok) {
    const int64_t ray = row / A.S;
    const int smp = (int)(row % A.S);
#pragma unroll
    for (int i = 0; i < 3; ++i) { o[i] = A.rays_o[ray * 3 + i]; G.d[i] = A.rays_d[ray * 3 + i]; }
    t0 = A.tdist[ray * (A.S + 1) + smp];
    t1 = A.tdist[ray * (A.S + 1) + smp + 1];
    radius = A.radii[ray];
  }
  // conical frustum -> Gaussian (helper.py:257-267)
  const float mu = 0.5f * (t0 + t1), hw = 0.5f * (t1 - t0);
  const float mu2 = mu * mu, hw2 = hw * hw, hw4 = hw2 * hw2;
  const float denom = fmaxf(3.f * mu2 + hw2, 1.1920929e-07f);
  const float inv_den = 1.f / denom;
  const float t_mean = mu + 2.f * mu * hw2 * inv_den;
  G.t_var = hw2 * (1.f / 3.f) - (4.f / 15.f) * hw4 * (12.f * mu2 - hw2) * inv_den * inv_den;
  G.r_var = (mu2 * 0.25f + (5.f / 12.f) * hw2 - (4.f / 15.f) * hw4 * inv_den) * radius * radius;
  const float dsq = fmaxf(G.d[0] * G.d[0] + G.d[1] * G.d[1] + G.d[2] * G.d[2], 1e-10f);
  G.inv_dsq = 1.f / dsq;
#pragma unroll
  for (int i = 0; i < 3; ++i) G.x[i] = fmaf(G.d[i], t_mean, o[i]);
  // contraction (helper.py:26-60): z = a x, J = a I + bq x x^T (identity inside the unit ball)
  G.m = fmaxf(G.x[0] * G.x[0] + G.x[1] * G.x[1] + G.x[2] * G.x[2], 1e-32f);
  G.a = 1.f;
  G.bq = 0.f;
  if (G.m > 1.f) {
    const float r = sqrtf(G.m), inv_m = 1.f / G.m;
    G.a = (2.f * r - 1.f) * inv_m;
    G.bq = 2.f * (1.f - r) * inv_m * inv_m;
  }
  G.xd = G.x[0] * G.d[0] + G.x[1] * G.d[1] + G.x[2] * G.d[2];
}

// One group of GS basis directions (JBASE .. JBASE + GS - 1), all 12 octaves, two directions per packed fp32x2
// instruction (sm_100 FMUL2 / FFMA2): per direction pair and octave 6 packed instructions instead of 16
// scalar ones.  emit(p, e sin, e cos) receives the pair index p = PBASE + l * GS + jj (columns 2p, 2p + 1).
template <int GS, int JBASE, int PBASE, class Emit>
__device__ __forceinline__ void ipe_group(const IpeArgs& A, const IpeRowGeom& G, Emit&& emit) {
  constexpr float kInv2Pi = 0.15915494309189535f;
  constexpr float k2PiHi = 6.2831854820251465f;           // fl32(2 pi)
  constexpr float k2PiLo = -1.7484555e-07f;               // 2 pi - fl32(2 pi)
  constexpr float kHalfLog2e = 0.72134752044448170f;      // 0.5 * log2(e)
  constexpr int NP = (GS + 1) / 2;                        // packed pairs (an odd group pads with a copy of its last direction)
  float2 S[NP], C[NP], e[NP], lv[NP], lm[NP];             // S = e * (+-sin), C = e * cos: the emitted values themselves
  // lifted mean  m_j = z.b_j  and variance  b_j^T J cov J^T b_j  with cov = t_var d d^T + r_var (I - d d^T/|d|^2):
  //   v = J b_j = a b_j + bq (x.b_j) x ;  var = t_var (d.v)^2 + r_var (|v|^2 - (d.v)^2 / |d|^2)
  // two directions per packed instruction, the row's scalars broadcast to both halves
  {
    const float2 x0 = make_float2(G.x[0], G.x[0]), x1 = make_float2(G.x[1], G.x[1]), x2 = make_float2(G.x[2], G.x[2]);
    const float2 d0 = make_float2(G.d[0], G.d[0]), d1 = make_float2(G.d[1], G.d[1]), d2 = make_float2(G.d[2], G.d[2]);
    const float2 pa = make_float2(G.a, G.a), pbq = make_float2(G.bq, G.bq), pxd = make_float2(G.xd, G.xd);
    const float2 pm = make_float2(G.m, G.m), p2a = make_float2(2.f * G.a, 2.f * G.a), paa = make_float2(G.a * G.a, G.a * G.a);
    const float2 ptv = make_float2(G.t_var, G.t_var), prv = make_float2(G.r_var, G.r_var);
    const float2 pnid = make_float2(-G.inv_dsq, -G.inv_dsq);
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      const int j0 = JBASE + 2 * q, j1 = JBASE + (2 * q + 1 < GS ? 2 * q + 1 : GS - 1);
      const float2 b0 = make_float2(A.basis[j0], A.basis[j1]);
      const float2 b1 = make_float2(A.basis[kIpeB + j0], A.basis[kIpeB + j1]);
      const float2 b2 = make_float2(A.basis[2 * kIpeB + j0], A.basis[2 * kIpeB + j1]);
      const float2 xb = __ffma2_rn(x2, b2, __ffma2_rn(x1, b1, __fmul2_rn(x0, b0)));
      const float2 db = __ffma2_rn(d2, b2, __ffma2_rn(d1, b1, __fmul2_rn(d0, b0)));
      const float2 k = __fmul2_rn(pbq, xb);
      const float2 dv = __ffma2_rn(k, pxd, __fmul2_rn(pa, db));
      const float2 vv = __ffma2_rn(k, __ffma2_rn(k, pm, __fmul2_rn(p2a, xb)), paa);       // |b_j| = 1
      const float2 dv2 = __fmul2_rn(dv, dv);
      const float2 var = __ffma2_rn(ptv, dv2, __fmul2_rn(prv, __ffma2_rn(dv2, pnid, vv)));
      lv[q] = make_float2(fmaxf(var.x, 0.f), fmaxf(var.y, 0.f));
      lm[q] = __fmul2_rn(pa, xb);
    }
  }
  const float2 kNeg2 = make_float2(-2.f, -2.f);
#pragma unroll
  for (int l = 0; l < kIpeDeg; ++l) {
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      if ((l & 3) == 0) {
        // ---- re-seed every 4 octaves: sin/cos of 2^l m_j from MUFU after Cody-Waite reduction, exp(-0.5 4^l var_j)
        // from ex2.approx
        const float ax = lm[q].x * (float)(1 << l), ay = lm[q].y * (float)(1 << l);
        const float kx = rintf(ax * kInv2Pi), ky = rintf(ay * kInv2Pi);
        const float rx = fmaf(-kx, k2PiLo, fmaf(-kx, k2PiHi, ax)), ry = fmaf(-ky, k2PiLo, fmaf(-ky, k2PiHi, ay));
        const float sc4 = -kHalfLog2e * (float)(1 << (2 * l));
        e[q] = make_float2(exp2f(sc4 * lv[q].x), exp2f(sc4 * lv[q].y));
        S[q] = __fmul2_rn(e[q], make_float2(__sinf(rx), __sinf(ry)));
        C[q] = __fmul2_rn(e[q], make_float2(__cosf(rx), __cosf(ry)));
      } else {
        // ---- in between, angle doubling and e <- e^4 on the products themselves (six packed instructions):
        //   S' = e^4 (-2 s c) = (-2 e^2 S) C ,  C' = e^4 (1 - 2 s^2) = e^4 + (-2 e^2 S) S
        // S carries (-1)^(l & 3) sin: the sign that -2 s c leaves behind is folded into the weights of the
        // odd-octave sine columns at upload time (pack_weight_kernel).  Error <= ~1e-5 (sin/cos) and <= 64 x the
        // ex2.approx error (exp) at the last octave before a re-seed, below the fp16 resolution of the operand.
        const float2 e2 = __fmul2_rn(e[q], e[q]);
        e[q] = __fmul2_rn(e2, e2);
        const float2 gS = __fmul2_rn(__fmul2_rn(e2, kNeg2), S[q]);
        const float2 nC = __ffma2_rn(gS, S[q], e[q]);
        S[q] = __fmul2_rn(gS, C[q]);
        C[q] = nC;
      }
      emit(PBASE + l * GS + 2 * q, S[q].x, C[q].x);
      if (2 * q + 1 < GS) emit(PBASE + l * GS + 2 * q + 1, S[q].y, C[q].y);
    }
  }
}

// Produce the 8 feature chunks of one pass for row r into the X ring.  Kernel column order (the weight columns are
// permuted to match at upload time, pack_weight_kernel): three groups of 8, 8 and 5 basis directions,
//   col = 2 * (pbase_g + l * gs_g + jj) + {0: sin, 1: cos},  j = jbase_g + jj,  (pbase, jbase, gs) = (0,0,8) (96,8,8) (192,16,5)
// so that a thread only carries one group's recurrence state at a time and the K-chunks come out in order.
//   sin/cos(2^l m): angle doubling, re-seeded from MUFU sin/cos after Cody-Waite reduction every 4
//                   octaves (error <= ~1e-5, below the fp16 resolution of the operand);
//   exp(-.5 4^l v): ex2.approx every 4th octave, e_{l+1} = e_l^4 in between.
//   (the sine of odd octaves comes out negated - three instructions per doubling instead of four - and the weight
//   columns carry the sign.)
// kStg: ring depth.  kWarpArrive == false: every thread arrives on bar_xfull (count 128).
// kWarpArrive == true (cluster-pair kernel): one arrival per warp, on the local bar_xfull (xfull_remote == 0)
// or on the leader CTA's barrier at cluster address xfull_remote + 8 * stage.
template <int kStg, bool kWarpArrive>
__device__ __forceinline__ void ipe_generate_pass(const IpeArgs& A, const IpeRowGeom& G, int r, unsigned char* sRingX,
                                                  uint64_t* bar_xfull, uint64_t* bar_xempty, uint32_t& xi,
                                                  uint32_t xfull_remote = 0, long long* tstamp = nullptr) {
  const uint32_t xempty_u32 = smem_u32(bar_xempty);
  uint32_t pk[4];
  const uint32_t row_u32 = smem_u32(sRingX) + 128u * (uint32_t)r, r7 = (uint32_t)(r & 7);
  uint32_t slot = 0;                                       // shared-space address of row r in the current ring slot
  auto acquire = [&]() {                                   // wait until ring slot xi % kStg has been consumed
    const int xs = xi % kStg;
    if (kWarpArrive) mbar_wait_guard<40>(&bar_xempty[xs], ((xi / kStg) & 1) ^ 1);
    else mbar_wait(&bar_xempty[xs], ((xi / kStg) & 1) ^ 1);
    if (tstamp) *tstamp++ = clock64();
    slot = row_u32 + (uint32_t)xs * kXChunkBytes;
  };
  // Publish chunk xi.  with_next: also poll the NEXT slot's "consumed" barrier - the phase check is issued before the
  // proxy fence so that its ~150-cycle latency hides behind the fence instead of following it.
  auto chunk_done = [&](bool with_next) {
    if (tstamp) *tstamp++ = clock64();
    uint32_t next_free = 0;
    if (with_next) {
      const uint32_t nx = (xi + 1) % kStg, npar = (((xi + 1) / kStg) & 1) ^ 1;
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "fence.proxy.async.shared::cta;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(next_free) : "r"(xempty_u32 + 8u * nx), "r"(npar) : "memory");
    } else {
      fence_proxy_async();
    }
    if (!kWarpArrive) {
      mbar_arrive(&bar_xfull[xi % kStg]);
    } else {
      __syncwarp();
      if ((threadIdx.x & 31) == 0) {
        if (xfull_remote) mbar_arrive_remote(xfull_remote + 8u * (xi % kStg));
        else mbar_arrive(&bar_xfull[xi % kStg]);
      }
    }
    ++xi;
    if (with_next) {
      if (next_free) {
        if (tstamp) *tstamp++ = clock64();
        slot = row_u32 + (uint32_t)(xi % kStg) * kXChunkBytes;
      } else {
        acquire();
      }
    }
  };
  auto emit = [&](int p, float sv, float cv) {             // p is a compile-time constant after unrolling
    if (p == 0) acquire();                                 // first chunk of the pass; later slots come from chunk_done(true)
    pk[p & 3] = cvt_f16x2(__float_as_uint(sv), __float_as_uint(cv));
    if ((p & 3) == 3) sts128(slot + ((((uint32_t)(p & 31) >> 2) ^ r7) << 4), pk[0], pk[1], pk[2], pk[3]);
    if ((p & 31) == 31) chunk_done(true);                  // chunk complete, more follow in this pass
  };
  ipe_group<8, 0, 0>(A, G, emit);
  ipe_group<8, 8, 96>(A, G, emit);
  ipe_group<5, 16, 192>(A, G, emit);
  // 252 pairs = 7 chunks + 28 pairs: groups 0..6 of the last chunk are written, zero the 8th
  sts128(slot + ((7u ^ r7) << 4), 0u, 0u, 0u, 0u);
  chunk_done(false);
}

__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_tc_kernel(const __grid_constant__ MlpProgram prog, const __grid_constant__ MlpArgs args,
              const __grid_constant__ IpeArgs ipe) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve-up (all chunk bases 1024-aligned, required by SWIZZLE_128B)
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int w_stage_bytes = prog.n_max * 128;
  unsigned char* sH = smem;                                            // kbh * 16 KB
  unsigned char* sRingW = sH + prog.kbh * kXChunkBytes;                // kStages * w_stage_bytes
  unsigned char* sRingX = sRingW + kStages * w_stage_bytes;            // kStagesX * 16 KB
  float* sParams = reinterpret_cast<float*>(sRingX + kStagesX * kXChunkBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sParams + ((prog.param_floats + 3) & ~3));
  uint64_t* bar_wfull = bars;                           // [kStages]
  uint64_t* bar_wempty = bar_wfull + kStages;           // [kStages]
  uint64_t* bar_xfull = bar_wempty + kStages;           // [kStagesX]
  uint64_t* bar_xempty = bar_xfull + kStagesX;          // [kStagesX]
  uint64_t* bar_tmem_full = bar_xempty + kStagesX;
  uint64_t* bar_act = bar_tmem_full + 2;                // [2] each: layer iterations alternate between the two,
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_act + 2);   // so a waiter is never two phases behind
  float* s_headx = reinterpret_cast<float*>(s_tmem + 4);     // [128][4] partial head sums of the upper column half

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool fused = args.fused_ipe != 0;

  for (int i = threadIdx.x; i < prog.param_floats; i += kMlpThreads) sParams[i] = args.params[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&bar_wfull[s], 1); mbar_init(&bar_wempty[s], 1); }
    for (int s = 0; s < kStagesX; ++s) { mbar_init(&bar_xfull[s], fused ? 128 : 1); mbar_init(&bar_xempty[s], 1); }
    mbar_init(&bar_tmem_full[0], 1);
    mbar_init(&bar_tmem_full[1], 1);
    mbar_init(&bar_act[0], kEpiWarps * 32);
    mbar_init(&bar_act[1], kEpiWarps * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===================== producer: bulk copies of weight (and, un-fused, input) chunks =====================
    if (lane == 0) {
      uint32_t wi = 0, xi = 0;
      for (int tile = blockIdx.x; tile < args.ntiles; tile += gridDim.x) {
        for (int l = 0; l < prog.n_layers; ++l) {
          const LayerDev L = prog.layers[l];
          const uint32_t wbytes = (uint32_t)L.n * 128u;
          const int nkb = L.kb_h + L.kb_x;
          for (int kb = 0; kb < nkb; ++kb, ++wi) {
            const int ws = wi % kStages;
            mbar_wait(&bar_wempty[ws], ((wi / kStages) & 1) ^ 1);
            mbar_expect_tx(&bar_wfull[ws], wbytes);
            bulk_g2s(sRingW + ws * w_stage_bytes, args.w_packed + L.w_off + (size_t)kb * wbytes, wbytes, &bar_wfull[ws]);
            if (kb >= L.kb_h && !fused) {
              const int xs = xi % kStagesX;
              mbar_wait(&bar_xempty[xs], ((xi / kStagesX) & 1) ^ 1);
              mbar_expect_tx(&bar_xfull[xs], kXChunkBytes);
              bulk_g2s(sRingX + xs * kXChunkBytes,
                       args.x_tiled + ((size_t)tile * prog.kbx + (kb - L.kb_h)) * kXChunkBytes, kXChunkBytes,
                       &bar_xfull[xs]);
              ++xi;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    uint32_t wi = 0, xi = 0, li = 0, ti = 0;
    for (int tile = blockIdx.x; tile < args.ntiles; tile += gridDim.x, ++ti) {
      const uint32_t acc = tmem_base + (ti & 1) * 256;
      for (int l = 0; l < prog.n_layers; ++l, ++li) {
        const LayerDev L = prog.layers[l];
        const uint32_t idesc = umma_idesc_f16(L.n);
        // Layer l > 0 needs the previous layer's epilogue (activations written, accumulator drained).
        // Layer 0 reads only the input ring and writes the OTHER accumulator buffer, so it may start
        // while the previous tile's last read-out is still running; that buffer was drained two
        // tiles ago, which the wait of this tile's layer 1 (two phases later on the same barrier
        // pair) transitively guarantees for every tile with >= 2 layers.
        if (l > 0) {
          const uint32_t pl = li - 1;         // phase index; barrier (pl & 1) completes every other layer
          mbar_wait(&bar_act[pl & 1], (pl >> 1) & 1);
          tc_fence_after();
        }
        if (args.timeline && blockIdx.x == 0 && lane == 0 && li < 64) args.timeline[li * 4 + 0] = clock64();
        const int nkb = L.kb_h + L.kb_x;
        for (int kb = 0; kb < nkb; ++kb, ++wi) {
          const int ws = wi % kStages;
          const bool from_x = kb >= L.kb_h;
          const int xs = xi % kStagesX;
          mbar_wait(&bar_wfull[ws], (wi / kStages) & 1);
          if (from_x) mbar_wait(&bar_xfull[xs], (xi / kStagesX) & 1);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_base = smem_u32(from_x ? sRingX + xs * kXChunkBytes : sH + kb * kXChunkBytes);
            const uint32_t b_base = smem_u32(sRingW + ws * w_stage_bytes);
#pragma unroll
            for (int k = 0; k < kKB / 16; ++k)
              tc_mma_f16(acc, umma_desc(a_base + k * 32), umma_desc(b_base + k * 32), idesc,
                         (kb | k) != 0 ? 1u : 0u);
            tc_commit(&bar_wempty[ws]);         // ring slots are released when these MMAs retire
            if (from_x) tc_commit(&bar_xempty[xs]);
            if (kb == nkb - 1) tc_commit(&bar_tmem_full[li & 1]);
          }
          __syncwarp();
          if (from_x) ++xi;
        }
        if (args.timeline && blockIdx.x == 0 && lane == 0 && li < 64) args.timeline[li * 4 + 1] = clock64();
      }
    }
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + kEpiWarps) {
    // ===================== epilogue (8 warps): 2 warps per TMEM lane quarter, half the columns each ============
    const int q = warp & 3;                     // TMEM lane quarter this warp may access (warp id % 4)
    const int ch = (warp - kEpiWarp0) >> 2;     // which half of the layer's columns
    const int r = q * 32 + lane;                // row within the tile
    uint32_t li = 0, ti = 0;
    for (int tile = blockIdx.x; tile < args.ntiles; tile += gridDim.x, ++ti) {
      const int64_t row = (int64_t)tile * kTileM + r;
      const bool row_ok = row < args.rows;
      const uint32_t acc = tmem_base + (ti & 1) * 256 + ((uint32_t)(q * 32) << 16);
      for (int l = 0; l < prog.n_layers; ++l, ++li) {
        const LayerDev L = prog.layers[l];
        const bool last = (l == prog.n_layers - 1);
        const bool has_head = L.head >= 0;
        const HeadDev Hd = prog.heads[has_head ? L.head : 0];
        const float* rb = (L.rowbias && args.rowbias && row_ok) ? args.rowbias + (row / args.rowbias_div) * L.n : nullptr;
        const int nh = L.n >> 1;
        const int cbeg = ch * nh, cend = cbeg + nh;
        float hacc[4] = {0.f, 0.f, 0.f, 0.f};
        mbar_wait(&bar_tmem_full[li & 1], (li >> 1) & 1);
        tc_fence_after();
        if (args.timeline && blockIdx.x == 0 && threadIdx.x == kEpiWarp0 * 32 && li < 64) args.timeline[li * 4 + 2] = clock64();
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(acc + (uint32_t)c0, v);
          float f[32];
          const float4* b4 = reinterpret_cast<const float4*>(sParams + L.bias_off + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = b4[j];
            f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + bb.x;
            f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bb.y;
            f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bb.z;
            f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bb.w;
          }
          if (rb) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(rb + c0) + j);
              f[4 * j + 0] += bb.x; f[4 * j + 1] += bb.y; f[4 * j + 2] += bb.z; f[4 * j + 3] += bb.w;
            }
          }
          if (L.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (has_head) {
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              if (n < Hd.hn) {
                const float4* w4 = reinterpret_cast<const float4*>(sParams + Hd.w_off + n * L.n + c0);
                float a = hacc[n];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 w = w4[j];
                  a = fmaf(f[4 * j + 0], w.x, a); a = fmaf(f[4 * j + 1], w.y, a);
                  a = fmaf(f[4 * j + 2], w.z, a); a = fmaf(f[4 * j + 3], w.w, a);
                }
                hacc[n] = a;
              }
            }
          }
          if (!last) {
            unsigned char* dst = sH + (c0 >> 6) * kXChunkBytes;
            const int kk = c0 & 63;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              __half2 h0 = __floats2half2_rn(f[g * 8 + 0], f[g * 8 + 1]);
              __half2 h1 = __floats2half2_rn(f[g * 8 + 2], f[g * 8 + 3]);
              __half2 h2 = __floats2half2_rn(f[g * 8 + 4], f[g * 8 + 5]);
              __half2 h3 = __floats2half2_rn(f[g * 8 + 6], f[g * 8 + 7]);
              uint4 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&h0);
              pk.y = *reinterpret_cast<uint32_t*>(&h1);
              pk.z = *reinterpret_cast<uint32_t*>(&h2);
              pk.w = *reinterpret_cast<uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(dst + tile_byte_offset(r, kk + g * 8)) = pk;
            }
          }
        }
        if (has_head) {                         // combine the two column halves, then post-process
          if (ch == 1) *reinterpret_cast<float4*>(s_headx + r * 4) = make_float4(hacc[0], hacc[1], hacc[2], hacc[3]);
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          if (ch == 0 && row_ok) {
            const float4 o4 = *reinterpret_cast<const float4*>(s_headx + r * 4);
            hacc[0] += o4.x; hacc[1] += o4.y; hacc[2] += o4.z; hacc[3] += o4.w;
            float* o = args.out[Hd.slot] + row * Hd.hn;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              if (n >= Hd.hn) break;
              float x = hacc[n] + sParams[Hd.b_off + n];
              if (Hd.post == 1) {
                float z = x + Hd.shift;
                x = z > 20.f ? z : log1pf(expf(z));
              } else if (Hd.post == 2) {
                x = (1.f / (1.f + expf(-x))) * (1.f + 2.f * Hd.shift) - Hd.shift;
              } else if (Hd.post == 3) {
                x = args.add[row * Hd.hn + n] + x;
              } else if (Hd.post == 4) {
                x = (n < 3) ? 1.f / (1.f + expf(-x)) : fmaxf(x, 0.f);
              }
              o[n] = x;
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");   // s_headx may be reused by the next head
        }
        fence_proxy_async();        // H stores (generic proxy) -> visible to the tensor core (async proxy)
        tc_fence_before();          // TMEM loads ordered before the arrive
        mbar_arrive(&bar_act[li & 1]);
        if (args.timeline && blockIdx.x == 0 && threadIdx.x == kEpiWarp0 * 32 && li < 64) args.timeline[li * 4 + 3] = clock64();
      }
    }
  } else if (fused && warp >= kFeatWarp0) {
    // ===================== feature generators (warps 12..15): fused IPE prologue =====================
    const int r = (warp - kFeatWarp0) * 32 + lane;
    uint32_t xi = 0;
    for (int tile = blockIdx.x; tile < args.ntiles; tile += gridDim.x) {
      const int64_t row = (int64_t)tile * kTileM + r;
      IpeRowGeom G;
      ipe_row_setup(ipe, row, row < args.rows, G);
      for (int l = 0; l < prog.n_layers; ++l) {
        if (prog.layers[l].kb_x == 0) continue;
        ipe_generate_pass<kStagesX, false>(ipe, G, r, sRingX, bar_xfull, bar_xempty, xi);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols));
  }
}


// ============================================================================ cluster-pair kernel
// Two CTAs (one cluster, two SMs) run every layer as ONE tcgen05.mma.cta_group::2 of M = 256 rows:
// CTA r owns rows [128 r, 128 r + 128) of the A operand and of the accumulator (its own TMEM) and stages
// only output columns [N/2 r, N/2 r + N/2) of every weight chunk.  Each SM therefore pulls half the
// weight bytes through L2 and a ring stage is 16 KB instead of 32 KB, which buys a 7-deep weight ring:
// the refill latency of a stage (MMA retire -> producer wake -> L2 -> peer relay) is ~2.5-3.5 k cycles,
// several times the 512 cycles a chunk lasts on the tensor core, so a 3-deep ring left the MMA warp
// waiting ~400 cycles per chunk (timeline in profiles/).
//
// Layers of one tile are dependent, so MMA and epilogue overlap at CHUNK granularity instead: the
// accumulator is double-buffered in TMEM (layer u -> columns 256 (u & 1)), the epilogue drains it 64
// columns at a time and signals each finished K-chunk of the next A operand separately; the MMA warp
// starts layer u + 1 on chunk 0 while the epilogue is still converting chunks 1..3 of layer u.  The
// first layer of a tile reads only generated features, so it also overlaps the previous tile's last
// epilogue.
//
// Measured building blocks (scripts/ubench_tc.cu, B200): tcgen05.ld sustains ~900 B/clk/SM; M=256 N=256
// K=16 cta_group::2 issues every 128 cycles (issue blocks while the pipe is busy: anything the MMA warp
// does between MMAs - waits, commits - idles the tensor core); ONE thread's bulk copies complete one
// mbarrier phase per ~800 cycles regardless of size while separate warps overlap - hence three producer
// warps.
//
// Warp roles (both CTAs unless noted): 0, 2, 3 weight producers (chunk i -> producer i % 3, stage i % S;
// in the peer CTA they also relay "my half has landed" to the leader), 1 MMA issuer (leader CTA only)
// + TMEM alloc, 4-11 epilogue, 12-15 feature generators.  Barriers that collect arrivals from both CTAs
// live in the leader; tcgen05.commit multicasts "stage free" / "accumulator ready" to both.
constexpr int kPairMaxStagesW = 8;
constexpr int kPairStagesX = 3;
constexpr int kOnesBytes = kTileM * 32;               // constant ones operand: [128 x 16] fp16, SWIZZLE_32B
constexpr int kPairProducers = 3;
constexpr int kPairMaxKbh = 4;
constexpr int kRowBiasVecs = 4;
constexpr int kPairBars = 2 * kPairMaxStagesW + 2 * kPairStagesX + 4 + 2 * kPairMaxKbh + 4;
static_assert(kPairBars % 2 == 0, "s_headx behind the barriers is read as float4");

// One row of the Fourier / Hann-windowed embedding as 64 packed fp16 values (column order of the reference embedders; columns
// beyond the embedding width are zero): one accurate sincos per coordinate, higher octaves by angle doubling.
template <bool IDENT>
__device__ __forceinline__ void fourier_row(const float (&p)[3], int nfreq, const float* hann, bool ok, uint32_t (&pk)[32]) {
  float vals[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) vals[j] = 0.f;
  float sn[3], cs[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) sincosf(p[a], &sn[a], &cs[a]);
  constexpr int base = IDENT ? 3 : 0;
  if (IDENT) { vals[0] = p[0]; vals[1] = p[1]; vals[2] = p[2]; }
#pragma unroll
  for (int k = 0; k < 10; ++k) {                               // 3 + 6 * 10 = 63 columns at most
    if (k < nfreq) {
      const float w = hann[k];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        vals[base + 6 * k + a] = w * sn[a];
        vals[base + 6 * k + 3 + a] = w * cs[a];
        const float s2 = 2.f * sn[a] * cs[a], c2 = 1.f - 2.f * sn[a] * sn[a];
        sn[a] = s2; cs[a] = c2;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) pk[j] = cvt_f16x2(__float_as_uint(ok ? vals[2 * j] : 0.f), __float_as_uint(ok ? vals[2 * j + 1] : 0.f));
}

// DUO is a template parameter (not a runtime flag): the one-tile-pair schedule of the 256-wide networks - the headline kernel -
// must compile exactly as it did before the duo schedule existed (as a runtime flag the extra loop level and indexing cost
// it 4 %: 1.140 -> 1.188 ms per C2 step).
template <bool DUO>
__device__ __forceinline__ void mlp_pair_body(const MlpProgram& prog, const MlpArgs& args, const IpeArgs& ipe, const int stages_w) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int w_stage_bytes = prog.n_max * 64;               // this CTA's half of an [n_max x 64] fp16 chunk
  // "duo" mode (narrow programs: n_max <= 128, so four 128-column accumulators fit in TMEM): a cluster keeps TWO tile pairs in
  // flight and walks the units (layer l, tile pair t) as l0t0 l0t1 l1t0 l1t1 ...: while one tile pair is in the layer-to-layer
  // hand-off (commit -> epilogue -> first operand chunk: 1.5 - 2.8 k cycles against 0.5 k cycles of tensor work for a 128-wide
  // layer) the other one's MMAs run.  Each tile pair has its own activation buffer, accumulator pair and hready barriers.
  constexpr int T = DUO ? 2 : 1;
  const uint32_t h_tile_bytes = (uint32_t)prog.kbh * kXChunkBytes;
  unsigned char* sH = smem;                                // [T][kbh][16 KB] activations (A operand of the next layer)
  unsigned char* sOnes = sH + T * h_tile_bytes;            // [4 KB] constant A operand of the bias MMA: [128 x 16], columns 0, 1 = 1.0
  unsigned char* sRingW = sOnes + kOnesBytes;              // [stages_w][w_stage_bytes]
  unsigned char* sRingX = sRingW + stages_w * w_stage_bytes;       // [kPairStagesX][16 KB]
  float* sParams = reinterpret_cast<float*>(sRingX + kPairStagesX * kXChunkBytes);   // head weights / biases only: the layer
  const int head_floats = prog.param_floats - prog.head_base;                      // biases ride in the GEMM (bias chunk)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sParams + ((head_floats + 3) & ~3));
  uint64_t* bar_wfull = bars;                              // [S] leader: both halves landed (own tx bytes + the peer's relay
                                                           //     arrive); peer: its own half landed
  uint64_t* bar_wempty = bar_wfull + kPairMaxStagesW;      // [S] multicast commit: stage consumed
  uint64_t* bar_xfull = bar_wempty + kPairMaxStagesW;      // [2] leader: feature chunk complete in BOTH CTAs (fused: one arrive
                                                           //     per feature warp of the pair; else own tx bytes + relay)
  uint64_t* bar_xempty = bar_xfull + kPairStagesX;         // [2] multicast commit
  uint64_t* bar_tfull = bar_xempty + kPairStagesX;         // [2 accumulator buffers; duo: 4 = (tile pair, layer parity)] multicast commit
  uint64_t* bar_hready = bar_tfull + 4;                    // [T][kPairMaxKbh] leader only: K-chunk c of the next A operand written
                                                           //       by the 16 epilogue warps of the pair
  uint64_t* bar_tfree = bar_hready + 2 * kPairMaxKbh;      // [1] leader only: a head layer's epilogue has finished its deferred
                                                           //     second pass over the accumulator (16 warp arrivals)
  uint64_t* bar_tlast = bar_tfree + 2;                     // [2] leader only, duo: the LAST layer's epilogue of tile pair t has read
                                                           //     its accumulator (nothing else orders the next iteration's layer 0 after it)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + kPairBars);       // kPairBars is even: 16-byte aligned
  float* s_headx = reinterpret_cast<float*>(s_tmem + 4);   // [128][4], read as float4
  float* s_rowbias = s_headx + kTileM * 4;                 // [kRowBiasVecs][128] per-ray bias vectors of the current tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool fused = args.fused_ipe != 0;                  // features written by the feature warps (IPE or Fourier prologue)
  const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int n_groups = (args.ntiles + 1) >> 1;             // 2 tiles per group: tile = 2 g + rank
  const int n_layers = prog.n_layers;
  const uint32_t S = (uint32_t)stages_w;

  for (int i = threadIdx.x; i < head_floats; i += kMlpThreads) sParams[i] = args.params[prog.head_base + i];
  // Bias: every layer's weight stream ends with one extra chunk whose K-columns 0 / 1 hold fp16 hi / lo parts of
  // the bias; multiplied by this constant operand it lands in the accumulator, so the epilogue adds nothing.
  for (int i = threadIdx.x; i < kOnesBytes / 16; i += kMlpThreads) reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  if (threadIdx.x < kTileM)      // (1.0h, 1.0h) in columns 0, 1 of row r: 16-byte half (r >> 2) & 1 of its 32-byte row
    *reinterpret_cast<uint32_t*>(sOnes + threadIdx.x * 32 + (((threadIdx.x >> 2) & 1) << 4)) = 0x3c003c00u;
  fence_proxy_async();
  if (threadIdx.x == 0) {
    const uint32_t both = rank == 0 ? 2u : 1u;             // the leader's barriers also count the peer
    for (int s = 0; s < kPairMaxStagesW; ++s) { mbar_init(&bar_wfull[s], both); mbar_init(&bar_wempty[s], 1); }
    for (int s = 0; s < kPairStagesX; ++s) {
      mbar_init(&bar_xfull[s], fused ? 8 : both);
      mbar_init(&bar_xempty[s], 1);
    }
    for (int s = 0; s < 4; ++s) mbar_init(&bar_tfull[s], 1);
    const uint32_t epi_arrivals = DUO ? (uint32_t)kEpiWarps : 2u * kEpiWarps;       // duo: four epilogue warps per CTA and tile pair
    for (int s = 0; s < 2 * kPairMaxKbh; ++s) mbar_init(&bar_hready[s], epi_arrivals);
    mbar_init(&bar_tfree[0], 2 * kEpiWarps);
    for (int s = 0; s < 2; ++s) mbar_init(&bar_tlast[s], epi_arrivals);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                   // both CTAs: barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0 || warp == 2 || (warp == 3 && !DUO)) {
    // ===================== weight producers: chunk i -> producer i % P, ring stage i % S =====================
    // (duo: two producers - warp 3 of the leader is the second MMA issuer; with the relay in the watcher warp a producer only
    // issues copies and never waits for one to land)
    if (lane == 0) {
      const uint32_t n_prod = DUO ? 2u : (uint32_t)kPairProducers;
      const uint32_t p = warp == 0 ? 0u : (uint32_t)(warp - 1);
      uint32_t wi = 0, xi = 0;
      for (int g0 = cluster; g0 < n_groups; g0 += T * n_clusters) {
        for (int l = 0; l < n_layers; ++l) {
         for (int t = 0; t < T; ++t) {
          int tile = 2 * (g0 + t * n_clusters) + (int)rank;
          if (tile >= args.ntiles) tile = args.ntiles - 1;          // padding tile: any valid rows, outputs masked
          const LayerDev L = prog.layers[l];
          const uint32_t half_bytes = (uint32_t)L.n * 64u;
          const int nkb = L.kb_h + L.kb_x;
          for (int kk = 0; kk <= nkb; ++kk, ++wi) {            // consumption order: the bias chunk (stored last) first
            const int kb = kk == 0 ? nkb : kk - 1;
            const bool from_x = kb >= L.kb_h && kb < nkb;
            const uint32_t xs = xi % kPairStagesX, xuse = xi / kPairStagesX;
            if (from_x) ++xi;
            // un-fused input: the 2-deep X ring is filled by ONE thread in order (a barrier waiter must never
            // be two phases ahead, which independent producers on a 2-stage ring could be)
            const bool load_x = from_x && !fused && p == 0;
            if (load_x) {
              mbar_wait_guard<100>(&bar_xempty[xs], (xuse & 1) ^ 1);
              mbar_expect_tx(&bar_xfull[xs], kXChunkBytes);
              bulk_g2s(sRingX + xs * kXChunkBytes, args.x_tiled + ((size_t)tile * prog.kbx + (kb - L.kb_h)) * kXChunkBytes,
                       kXChunkBytes, &bar_xfull[xs]);
            }
            const bool load_w = wi % n_prod == p;
            const uint32_t ws = wi % S, use = wi / S;
            if (load_w) {
              mbar_wait_guard<100>(&bar_wempty[ws], (use & 1) ^ 1);
              mbar_expect_tx(&bar_wfull[ws], half_bytes);
              bulk_g2s(sRingW + ws * w_stage_bytes, args.w_packed + L.w_off + (size_t)kb * (2u * half_bytes) + rank * half_bytes,
                       half_bytes, &bar_wfull[ws]);
            }
            if (rank != 0 && load_x) {          // relay to the leader once this CTA's bytes are in shared memory
              mbar_wait_guard<100>(&bar_xfull[xs], xuse & 1);
              mbar_arrive_remote(mapa_u32(smem_u32(&bar_xfull[xs]), 0));
            }
            // (the weight stages are relayed by the peer's watcher warp below: a producer that waited for its own copy to land
            // before issuing the next one moved one chunk per L2 round trip - three producers barely fed one 512-cycle chunk
            // of tensor work each)
          }
         }
        }
      }
    }
  } else if (warp == 1 && rank != 0) {
    // ===================== peer CTA: weight-stage watcher =====================
    // tells the leader "my half of weight stage s has landed", stage by stage in ring order
    if (lane == 0) {
      const uint32_t wfull_leader = mapa_u32(smem_u32(&bar_wfull[0]), 0);
      uint32_t ws = 0, wpar = 0;
      for (int g0 = cluster; g0 < n_groups; g0 += T * n_clusters) {
        for (int l = 0; l < n_layers; ++l) {
          const int nchunks_w = (prog.layers[l].kb_h + prog.layers[l].kb_x + 1) * T;
          for (int i = 0; i < nchunks_w; ++i) {
            mbar_wait_guard<40>(&bar_wfull[ws], wpar);
            mbar_arrive_remote(wfull_leader + 8u * ws);
            if (++ws == S) { ws = 0; wpar ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== MMA issuer (one thread of the leader CTA; duo: two - warp 1 issues the units of tile pair 0, warp 3
    // those of tile pair 1: a 128-wide layer is 9 MMAs of 64 cycles, and the polls + commits around them cost one thread
    // ~1.5 k cycles per unit) =====================
    // A single thread runs ~5 cycles per dependent instruction and every try_wait / commit costs 60-180 cycles
    // (scripts/ubench_tc.cu, test 8), while a K = 64 chunk lasts only 512 cycles on the tensor core: the loop
    // therefore works on GROUPS of up to three chunks - all their barriers polled in one parallel try_wait,
    // MMAs and commits issued back to back, ring state advanced by increments (no division).  Grouping of a
    // layer with kb_h activation chunks: [h0] [h1 h2] [h3 ...] (the first alone so that the layer starts as soon
    // as the previous epilogue has produced one K-chunk), feature chunks alone (their ring is 2 deep), and the
    // bias chunk rides with the last group.
    // The loop runs on ALL lanes of the warp with warp-uniform values and only the tcgen05 instructions sit under
    // elect.sync: tcgen05.mma takes its descriptors from uniform registers, and when the issue code ran on one lane of
    // a divergent branch the compiler had to move 7 values per MMA through R2UR in a wait-loop (~130 cycles per MMA).
    if (rank == 0) {
      uint32_t u = 0;                                // unit counter
      uint32_t hpar_bits = 0u;                       // bit t: parity of the hready phase tile pair t's next H-reading unit waits for
      uint32_t ws = 0, wpar = 0, xs = 0, xpar = 0;   // weight / feature ring cursors + phase parities
      bool prev_had_h = true;
      const uint64_t desc_hi = umma_desc(0);
      const uint32_t wfull0 = smem_u32(&bar_wfull[0]), wempty0 = smem_u32(&bar_wempty[0]);
      const uint32_t xfull0 = smem_u32(&bar_xfull[0]), xempty0 = smem_u32(&bar_xempty[0]);
      const uint32_t hready_a = smem_u32(&bar_hready[0]), tfull0 = smem_u32(&bar_tfull[0]);
      const uint32_t tfree_a = smem_u32(&bar_tfree[0]), tlast_a = smem_u32(&bar_tlast[0]);
      uint32_t tfpar = 0;
      bool head_hist[2] = {false, false};
      const uint32_t h16 = (smem_u32(sH) >> 4) & 0x3FFF, x16 = (smem_u32(sRingX) >> 4) & 0x3FFF;
      const uint32_t w16 = (smem_u32(sRingW) >> 4) & 0x3FFF;
      const uint64_t ones_desc = umma_desc_sw32(smem_u32(sOnes));
      const uint32_t wstage16 = (uint32_t)w_stage_bytes >> 4;
      uint32_t iter = 0;
      const int my_t = warp == 3 ? 1 : 0;
      for (int g0 = cluster; g0 < n_groups; g0 += T * n_clusters, ++iter) {
        for (int l = 0; l < n_layers; ++l) {
         for (int t = 0; t < T; ++t, ++u) {
          const LayerDev L = prog.layers[l];
          const uint32_t idesc = umma_idesc_f16(L.n, 2 * kTileM);
          const int kb_h = L.kb_h, nkb = L.kb_h + L.kb_x;
          if (DUO && t != my_t) {               // the other issuer's unit: step the ring cursors over its stages
            uint32_t adv = (uint32_t)nkb + 1u + ws;
            while (adv >= S) { adv -= S; wpar ^= 1u; }
            ws = adv;
            uint32_t advx = (uint32_t)L.kb_x + xs;
            while (advx >= (uint32_t)kPairStagesX) { advx -= kPairStagesX; xpar ^= 1u; }
            xs = advx;
            continue;
          }

          const uint32_t abuf = DUO ? (uint32_t)(2 * t + (l & 1)) : (u & 1);       // accumulator buffer / tfull barrier
          const uint32_t acc = tmem_base + (DUO ? abuf * 128u : abuf * 256u);
          const uint32_t hready_t = hready_a + 8u * (uint32_t)(t * kPairMaxKbh);
          const uint32_t h16_t = h16 + (uint32_t)t * (h_tile_bytes >> 4);
          const uint32_t hpar = (hpar_bits >> t) & 1u;
          const bool tl = args.timeline && blockIdx.x == 0 && u < 64 && lane == 0;
          long long wait_sum = 0;
          if (tl) args.timeline[u * 12 + 0] = clock64();
          // H chunk kb is overwritten by the epilogue of unit u only after all MMAs of unit u retired (tfull).
          {
            // Bias first: ones (columns 0, 1 of the constant operand) x (hi, lo) -> one K = 16 MMA that INITIALISES the
            // accumulator.  It needs only its weight stage, so it is issued (and runs) while the previous layer's
            // epilogue is still producing this layer's first operand chunk.
            // Accumulator buffer (u & 1) was last read by the epilogue of unit u - 2.  If unit u - 1 read activation
            // chunks it waited for that epilogue's last chunk signal, so the buffer is free now; if unit u - 1 read only
            // features (first layer of a tile) nothing has ordered us after epilogue(u - 2) yet: then wait for the first
            // operand chunk of this unit as well (signalled by epilogue(u - 1), which runs after epilogue(u - 2)).
            // A hidden layer with an fp32 head re-reads its accumulator after it has signalled all operand chunks
            // (deferred head pass): its buffer is free only once bar_tfree completes.
            if (DUO) {
              // Buffer (t, l & 1) was last read by the epilogue of unit (t, l - 2), which ran before the epilogue of (t, l - 1)
              // whose chunks this warp waited for when it issued that unit - except across iterations: nothing waits for the
              // previous iteration's LAST layer's epilogue, so the first unit that reuses ITS buffer (layer 0 or 1, by parity)
              // waits for bar_tlast.  (With an even layer count that is layer 1: the slow head epilogue gets a whole unit of slack.)
              const int l_reuse = (n_layers - 1) & 1;
              if (l == l_reuse && iter > 0 && l_reuse < n_layers) mbar_wait2_spin(wfull0 + 8u * ws, wpar, tlast_a + 8u * (uint32_t)t, (iter - 1) & 1u);
              else mbar_wait2_spin(wfull0 + 8u * ws, wpar, wfull0 + 8u * ws, wpar);
            } else {
            if (head_hist[u & 1]) {
              mbar_wait2_spin(tfree_a, tfpar, tfree_a, tfpar);
              tfpar ^= 1u;
            }
            head_hist[u & 1] = L.head >= 0 && l != n_layers - 1 && !(L.rowbias && args.rowbias);
            if (u < 2 || prev_had_h) mbar_wait2_spin(wfull0 + 8u * ws, wpar, wfull0 + 8u * ws, wpar);
            else mbar_wait2_spin(wfull0 + 8u * ws, wpar, kb_h > 0 ? hready_a : xfull0 + 8u * xs, kb_h > 0 ? hpar : xpar);
            prev_had_h = kb_h > 0;
            }
            tc_fence_after();
            if (elect_one()) {
              tc_mma_f16_pair(acc, ones_desc, desc_hi | (uint64_t)(w16 + ws * wstage16), idesc, 0u);
              tc_commit_pair_addr(wempty0 + 8u * ws);
            }
            if (++ws == S) { ws = 0; wpar ^= 1; }
            __syncwarp();
          }
          int kb = 0, gi3 = 0;
          while (kb < nkb) {
            const bool is_h = kb < kb_h;
            const int cnt = (is_h && kb > 0 && kb + 1 < kb_h) ? 2 : 1;
            const bool last_group = kb + cnt == nkb;
            const int nst = cnt;                                   // ring stages consumed by this group (<= 2)
            // ---- one parallel poll (several try_waits in flight together cost ~220 cycles, separate phase checks
            // ~150 each): the weight stages + the newest operand chunk (earlier ones are implied)
            uint32_t wb[2], wp[2];
            {
              uint32_t s_ = ws, p_ = wpar;
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                wb[i] = wfull0 + 8u * s_; wp[i] = p_;
                if (i + 1 < nst) { if (++s_ == S) { s_ = 0; p_ ^= 1; } }
              }
            }
            const uint32_t ob = is_h ? hready_t + 8u * (uint32_t)(kb + cnt - 1) : xfull0 + 8u * xs;
            const uint32_t op = is_h ? hpar : xpar;
            const long long c0 = tl ? clock64() : 0;
            mbar_wait3_spin(wb[0], wp[0], wb[1], wp[1], ob, op);
            const bool tl3 = tl && (u == 11 || u == 12) && gi3 < 5;
            if (tl) {
              const long long c1 = clock64();
              wait_sum += c1 - c0;
              if (tl3) { args.timeline[840 + (u - 11) * 16 + gi3 * 3 + 0] = c0; args.timeline[840 + (u - 11) * 16 + gi3 * 3 + 1] = c1; }
            }
            tc_fence_after();
            // ---- issue
            const bool leader_lane = elect_one();
#pragma unroll 1
            for (int i = 0; i < cnt; ++i) {
              const uint32_t a16 = is_h ? h16_t + (uint32_t)(kb + i) * (kXChunkBytes >> 4) : x16 + xs * (kXChunkBytes >> 4);
              const uint64_t adesc = desc_hi | (uint64_t)a16;
              const uint64_t bdesc = desc_hi | (uint64_t)(w16 + ws * wstage16);
              if (leader_lane) {
#pragma unroll
                for (int k = 0; k < kKB / 16; ++k)
                  tc_mma_f16_pair(acc, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                tc_commit_pair_addr(wempty0 + 8u * ws);
                if (!is_h) tc_commit_pair_addr(xempty0 + 8u * xs);
                if (last_group && i == cnt - 1) tc_commit_pair_addr(tfull0 + 8u * abuf);       // accumulator complete
              }
              if (++ws == S) { ws = 0; wpar ^= 1; }
              if (!is_h) { if (++xs == kPairStagesX) { xs = 0; xpar ^= 1; } }
            }
            __syncwarp();
            if (tl3) args.timeline[840 + (u - 11) * 16 + gi3 * 3 + 2] = clock64();
            ++gi3;
            kb += cnt;
          }
          if (kb_h) hpar_bits ^= 1u << t;
          if (tl) { args.timeline[u * 12 + 1] = clock64(); args.timeline[u * 12 + 8] = wait_sum; }
         }
        }
      }
    }
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + kEpiWarps) {
    // ===================== epilogue (8 warps): 2 warps per TMEM lane quarter; K-chunk c of the output is columns
    // [64 c, 64 c + 64): warp half `ch` converts its 32 of them, then signals the chunk ============
    const int q = warp & 3;
    const int ch = (warp - kEpiWarp0) >> 2;
    const int r = q * 32 + lane;
    const uint32_t hready0 = mapa_u32(smem_u32(&bar_hready[0]), 0);
    const uint32_t tfree0 = mapa_u32(smem_u32(&bar_tfree[0]), 0);
    const uint32_t tlast0 = mapa_u32(smem_u32(&bar_tlast[0]), 0);
    const uint32_t sparams_u32 = smem_u32(sParams) - 4u * (uint32_t)prog.head_base;   // indexed with whole-block offsets
    const uint32_t headx_u32 = smem_u32(s_headx) + 16u * r;
    // this thread's 16-byte slots in a K-chunk of H: row r, 16-byte groups 4 ch .. 4 ch + 3 (128B swizzle)
    const uint32_t h_row = smem_u32(sH) + 128u * r;
    uint32_t h_off0[4];
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) h_off0[gq] = h_row + ((uint32_t)((4 * ch + gq) ^ (r & 7)) << 4);
    uint32_t u = 0;
    uint32_t tf_use = 0;                                      // bit b: parity of the next phase of bar_tfull[b] to wait for
    if (DUO) {
      // ---------- duo schedule: warps 4-7 drain the units of tile pair 0, warps 8-11 those of tile pair 1, concurrently (an
      // epilogue of a 128-wide unit is a ~1.2 k-cycle latency chain - TMEM load, convert, store, fence, arrive - not a
      // throughput problem).  Each warp takes all 64 columns of a chunk for its 32 rows; the head's dot products stay in one thread.
      const int t = ch;
      uint32_t hq[2][4];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int gq = 0; gq < 4; ++gq)
          hq[hh][gq] = h_row + (uint32_t)t * h_tile_bytes + ((uint32_t)((4 * hh + gq) ^ (r & 7)) << 4);
      const uint32_t hready_t = hready0 + 8u * (uint32_t)(t * kPairMaxKbh);
      for (int g0 = cluster; g0 < n_groups; g0 += 2 * n_clusters) {
        const int g = g0 + t * n_clusters;
        const int64_t row = (int64_t)(2 * g + (int)rank) * kTileM + r;
        const bool row_ok = g < n_groups && row < args.rows;
        for (int l = 0; l < n_layers; ++l) {
          const LayerDev L = prog.layers[l];
          const bool last = (l == n_layers - 1);
          const bool relu = L.relu != 0;
          const int nchunks = L.n >> 6;
          const uint32_t abuf = (uint32_t)(2 * t + (l & 1));
          const uint32_t acc = tmem_base + abuf * 128u + ((uint32_t)(q * 32) << 16);
          mbar_wait_guard<20>(&bar_tfull[abuf], (tf_use >> abuf) & 1u);
          tf_use ^= 1u << abuf;
          tc_fence_after();
          const HeadDev Hd = prog.heads[L.head >= 0 ? L.head : 0];
          float hacc[4] = {0.f, 0.f, 0.f, 0.f};
          for (int c = 0; c < nchunks; ++c) {
            uint32_t va[32], vb[32];
            tmem_ld32_nowait(acc + (uint32_t)(c * 64), va);
            tmem_ld32_nowait(acc + (uint32_t)(c * 64 + 32), vb);
            tmem_wait_ld();
            if (!last) {
              const uint32_t cb = (uint32_t)c * kXChunkBytes;
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const uint32_t (&v)[32] = hh ? vb : va;
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                  uint32_t p0, p1, p2, p3;
                  if (relu) {
                    p0 = cvt_relu_f16x2(v[gq * 8 + 0], v[gq * 8 + 1]); p1 = cvt_relu_f16x2(v[gq * 8 + 2], v[gq * 8 + 3]);
                    p2 = cvt_relu_f16x2(v[gq * 8 + 4], v[gq * 8 + 5]); p3 = cvt_relu_f16x2(v[gq * 8 + 6], v[gq * 8 + 7]);
                  } else {
                    p0 = cvt_f16x2(v[gq * 8 + 0], v[gq * 8 + 1]); p1 = cvt_f16x2(v[gq * 8 + 2], v[gq * 8 + 3]);
                    p2 = cvt_f16x2(v[gq * 8 + 4], v[gq * 8 + 5]); p3 = cvt_f16x2(v[gq * 8 + 6], v[gq * 8 + 7]);
                  }
                  sts128(hq[hh][gq] + cb, p0, p1, p2, p3);
                }
              }
              fence_proxy_async();        // H stores (generic proxy) -> visible to the tensor core (async proxy)
              tc_fence_before();          // TMEM loads of this chunk ordered before the arrive
              __syncwarp();
              if (lane == 0) mbar_arrive_remote(hready_t + 8u * c);
            } else if (L.head >= 0) {
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const uint32_t (&v)[32] = hh ? vb : va;
                const int c0 = c * 64 + hh * 32;
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                  if (n < Hd.hn) {
                    const uint32_t w4 = sparams_u32 + 4u * (uint32_t)(Hd.w_off + n * L.n + c0);
                    float a = hacc[n];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      const float4 w = lds128(w4 + 16u * j);
                      const float f0 = __uint_as_float(v[4 * j + 0]), f1 = __uint_as_float(v[4 * j + 1]);
                      const float f2 = __uint_as_float(v[4 * j + 2]), f3 = __uint_as_float(v[4 * j + 3]);
                      a = fmaf(relu ? fmaxf(f0, 0.f) : f0, w.x, a); a = fmaf(relu ? fmaxf(f1, 0.f) : f1, w.y, a);
                      a = fmaf(relu ? fmaxf(f2, 0.f) : f2, w.z, a); a = fmaf(relu ? fmaxf(f3, 0.f) : f3, w.w, a);
                    }
                    hacc[n] = a;
                  }
                }
              }
            }
          }
          if (last) {                       // accumulator read: the next iteration's layer 0 of this tile pair may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(tlast0 + 8u * (uint32_t)t);
            if (L.head >= 0 && row_ok) {
              const float4 hb = lds128(sparams_u32 + 4u * Hd.b_off);
              const float hbias[4] = {hb.x, hb.y, hb.z, hb.w};
              float* o = args.out[Hd.slot] + row * Hd.hn;
#pragma unroll
              for (int n = 0; n < 4; ++n) {
                if (n >= Hd.hn) break;
                float x = hacc[n] + hbias[n];
                if (Hd.post == 1) {
                  float z = x + Hd.shift;
                  x = z > 20.f ? z : log1pf(expf(z));
                } else if (Hd.post == 2) {
                  x = (1.f / (1.f + expf(-x))) * (1.f + 2.f * Hd.shift) - Hd.shift;
                } else if (Hd.post == 3) {
                  x = args.add[row * Hd.hn + n] + x;
                } else if (Hd.post == 4) {
                  x = (n < 3) ? 1.f / (1.f + expf(-x)) : fmaxf(x, 0.f);
                }
                o[n] = x;
              }
            }
          }
        }
      }
    } else
    for (int g0 = cluster; g0 < n_groups; g0 += T * n_clusters) {
      for (int l = 0; l < n_layers; ++l) {
       for (int t = 0; t < T; ++t, ++u) {
        const int g = g0 + t * n_clusters;                     // >= n_groups: padding tile pair of a duo iteration (nothing stored)
        const int tile = 2 * g + (int)rank;
        const int64_t row = (int64_t)tile * kTileM + r;
        const bool row_ok = g < n_groups && row < args.rows;
        uint32_t h_off[4];
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) h_off[gq] = h_off0[gq] + (uint32_t)t * h_tile_bytes;
        const uint32_t hready_t = hready0 + 8u * (uint32_t)(t * kPairMaxKbh);
        const LayerDev L = prog.layers[l];
        const bool last = (l == n_layers - 1);
        const bool has_head = L.head >= 0;
        const bool relu = L.relu != 0;
        const int nchunks = L.n >> 6;
        const uint32_t abuf = DUO ? (uint32_t)(2 * t + (l & 1)) : (u & 1);
        const uint32_t acc = tmem_base + (DUO ? abuf * 128u : abuf * 256u) + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32);
        const bool tl = args.timeline && blockIdx.x == 0 && threadIdx.x == kEpiWarp0 * 32 && u < 64;
        // Per-ray bias (view-direction term): when a tile spans a whole number of rays (or one ray spans whole tiles) its
        // <= 4 bias vectors are fetched into shared memory BEFORE the wait for the accumulator, so the L2 latency is
        // hidden and the 8 loads per 32 columns become broadcast ld.shared; otherwise each row reads global memory.
        uint32_t rb_s = 0;
        const float* rb_g = nullptr;
        if (L.rowbias && args.rowbias) {
          const int div = args.rowbias_div;
          const int nr = (div % kTileM == 0) ? 1 : ((kTileM % div == 0 && kTileM / div <= kRowBiasVecs && L.n <= 128) ? kTileM / div : 0);
          if (nr > 0 && nr * L.n <= kRowBiasVecs * 128) {
            const int te = threadIdx.x - kEpiWarp0 * 32;
            for (int i = te; i < nr * L.n; i += kEpiWarps * 32) {
              const int64_t row0 = (int64_t)tile * kTileM + (int64_t)(i / L.n) * div;
              s_rowbias[i] = row0 < args.rows ? __ldg(args.rowbias + (row0 / div) * L.n + (i % L.n)) : 0.f;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
            rb_s = smem_u32(s_rowbias) + 4u * (uint32_t)((nr == 1 ? 0 : r / div) * L.n);
          } else if (row_ok) {
            rb_g = args.rowbias + (row / div) * L.n;
          }
        }
        if (tl) args.timeline[u * 12 + 3] = clock64();
        mbar_wait_guard<20>(&bar_tfull[abuf], (tf_use >> abuf) & 1u);
        tf_use ^= 1u << abuf;
        tc_fence_after();
        if (tl) args.timeline[u * 12 + 4] = clock64();

        // chunk c done: its 32 columns of this warp are in shared memory -> tell the MMA warp (leader CTA)
        auto chunk_ready = [&](int c) {
          fence_proxy_async();        // H stores (generic proxy) -> visible to the tensor core (async proxy)
          tc_fence_before();          // TMEM loads of this chunk ordered before the arrive
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(hready_t + 8u * c);
        };

        const HeadDev Hd = prog.heads[has_head ? L.head : 0];
        float hacc[4] = {0.f, 0.f, 0.f, 0.f};
        if (!last && !rb_s && !rb_g) {
          // ---------- hidden layer (bias already in the accumulator): convert (+ ReLU), store, signal; a layer that also
          // feeds an fp32 head (density) accumulates it AFTER the signal, so the next layer's MMAs are not held up.
          // Two TMEM loads in flight: the next chunk's 32 columns arrive while these are converted.
          auto head32 = [&](const uint32_t (&v)[32], int c) {
            const int c0 = c * 64 + ch * 32;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              if (n < Hd.hn) {
                const uint32_t w4 = sparams_u32 + 4u * (uint32_t)(Hd.w_off + n * L.n + c0);
                float a = hacc[n];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 w = lds128(w4 + 16u * j);
                  const float f0 = __uint_as_float(v[4 * j + 0]), f1 = __uint_as_float(v[4 * j + 1]);
                  const float f2 = __uint_as_float(v[4 * j + 2]), f3 = __uint_as_float(v[4 * j + 3]);
                  a = fmaf(relu ? fmaxf(f0, 0.f) : f0, w.x, a); a = fmaf(relu ? fmaxf(f1, 0.f) : f1, w.y, a);
                  a = fmaf(relu ? fmaxf(f2, 0.f) : f2, w.z, a); a = fmaf(relu ? fmaxf(f3, 0.f) : f3, w.w, a);
                }
                hacc[n] = a;
              }
            }
          };
          auto store32 = [&](const uint32_t (&v)[32], int c) {
            const uint32_t cb = (uint32_t)c * kXChunkBytes;
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              uint32_t p0, p1, p2, p3;
              if (relu) {
                p0 = cvt_relu_f16x2(v[gq * 8 + 0], v[gq * 8 + 1]); p1 = cvt_relu_f16x2(v[gq * 8 + 2], v[gq * 8 + 3]);
                p2 = cvt_relu_f16x2(v[gq * 8 + 4], v[gq * 8 + 5]); p3 = cvt_relu_f16x2(v[gq * 8 + 6], v[gq * 8 + 7]);
              } else {
                p0 = cvt_f16x2(v[gq * 8 + 0], v[gq * 8 + 1]); p1 = cvt_f16x2(v[gq * 8 + 2], v[gq * 8 + 3]);
                p2 = cvt_f16x2(v[gq * 8 + 4], v[gq * 8 + 5]); p3 = cvt_f16x2(v[gq * 8 + 6], v[gq * 8 + 7]);
              }
              sts128(h_off[gq] + cb, p0, p1, p2, p3);
            }
          };
          uint32_t va[32], vb[32];
          tmem_ld32_nowait(acc, va);
          for (int c = 0; c < nchunks; c += 2) {
            tmem_wait_ld();
            const bool more1 = c + 1 < nchunks;
            if (more1) tmem_ld32_nowait(acc + (uint32_t)((c + 1) * 64), vb);
            store32(va, c);
            chunk_ready(c);
            if (more1) {
              tmem_wait_ld();
              if (c + 2 < nchunks) tmem_ld32_nowait(acc + (uint32_t)((c + 2) * 64), va);
              store32(vb, c + 1);
              chunk_ready(c + 1);
            }
          }
          if (has_head) {
            // Deferred head pass: every operand chunk of the next layer is signalled; now read the (still intact) fp32
            // accumulator once more for the head dot products, then hand the buffer back (bar_tfree).
            tmem_ld32_nowait(acc, va);
            for (int c = 0; c < nchunks; c += 2) {
              tmem_wait_ld();
              const bool more1 = c + 1 < nchunks;
              if (more1) tmem_ld32_nowait(acc + (uint32_t)((c + 1) * 64), vb);
              head32(va, c);
              if (more1) {
                tmem_wait_ld();
                if (c + 2 < nchunks) tmem_ld32_nowait(acc + (uint32_t)((c + 2) * 64), va);
                head32(vb, c + 1);
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(tfree0);
          }
        } else {
          // ---------- general layer: per-row bias, fp32 output head, last layer (no activations stored)
          auto process_general = [&](uint32_t (&v)[32], int c) {
            const int c0 = c * 64 + ch * 32;
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (rb_s) {                         // per-ray bias staged in shared memory (see above)
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 bb = lds128(rb_s + 4u * (uint32_t)c0 + 16u * j);
                f[4 * j + 0] += bb.x; f[4 * j + 1] += bb.y; f[4 * j + 2] += bb.z; f[4 * j + 3] += bb.w;
              }
            } else if (rb_g) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(rb_g + c0) + j);
                f[4 * j + 0] += bb.x; f[4 * j + 1] += bb.y; f[4 * j + 2] += bb.z; f[4 * j + 3] += bb.w;
              }
            }
            if (!last) {
              // store + signal FIRST: the next layer's MMAs start while this warp is still busy with the head
              const uint32_t cb = (uint32_t)c * kXChunkBytes;
#pragma unroll
              for (int gq = 0; gq < 4; ++gq) {
                uint32_t p0, p1, p2, p3;
                if (relu) {
                  p0 = cvt_relu_f16x2(__float_as_uint(f[gq * 8 + 0]), __float_as_uint(f[gq * 8 + 1]));
                  p1 = cvt_relu_f16x2(__float_as_uint(f[gq * 8 + 2]), __float_as_uint(f[gq * 8 + 3]));
                  p2 = cvt_relu_f16x2(__float_as_uint(f[gq * 8 + 4]), __float_as_uint(f[gq * 8 + 5]));
                  p3 = cvt_relu_f16x2(__float_as_uint(f[gq * 8 + 6]), __float_as_uint(f[gq * 8 + 7]));
                } else {
                  p0 = cvt_f16x2(__float_as_uint(f[gq * 8 + 0]), __float_as_uint(f[gq * 8 + 1]));
                  p1 = cvt_f16x2(__float_as_uint(f[gq * 8 + 2]), __float_as_uint(f[gq * 8 + 3]));
                  p2 = cvt_f16x2(__float_as_uint(f[gq * 8 + 4]), __float_as_uint(f[gq * 8 + 5]));
                  p3 = cvt_f16x2(__float_as_uint(f[gq * 8 + 6]), __float_as_uint(f[gq * 8 + 7]));
                }
                sts128(h_off[gq] + cb, p0, p1, p2, p3);
              }
              chunk_ready(c);
            }
            if (has_head) {
              if (relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
              }
#pragma unroll
              for (int n = 0; n < 4; ++n) {
                if (n < Hd.hn) {
                  const uint32_t w4 = sparams_u32 + 4u * (uint32_t)(Hd.w_off + n * L.n + c0);
                  float a = hacc[n];
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float4 w = lds128(w4 + 16u * j);
                    a = fmaf(f[4 * j + 0], w.x, a); a = fmaf(f[4 * j + 1], w.y, a);
                    a = fmaf(f[4 * j + 2], w.z, a); a = fmaf(f[4 * j + 3], w.w, a);
                  }
                  hacc[n] = a;
                }
              }
            }
          };
          for (int c = 0; c < nchunks; ++c) {       // single-buffered: this path also carries the head accumulators
            uint32_t v[32];
            tmem_ld32_nowait(acc + (uint32_t)(c * 64), v);
            tmem_wait_ld();
            process_general(v, c);
          }
        }
        {
          if (last) tc_fence_before();
          if (last && DUO) {                 // accumulator read: the next iteration's layer 0 of this tile pair may overwrite it
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(tlast0 + 8u * (uint32_t)t);
          }
          if (has_head) {                         // combine the two column halves, then post-process
            if (ch == 1) sts128(headx_u32, __float_as_uint(hacc[0]), __float_as_uint(hacc[1]), __float_as_uint(hacc[2]), __float_as_uint(hacc[3]));
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
            if (ch == 0 && row_ok) {
              const float4 o4 = lds128(headx_u32);
              hacc[0] += o4.x; hacc[1] += o4.y; hacc[2] += o4.z; hacc[3] += o4.w;
              const float4 hb = lds128(sparams_u32 + 4u * Hd.b_off);
              const float hbias[4] = {hb.x, hb.y, hb.z, hb.w};
              float* o = args.out[Hd.slot] + row * Hd.hn;
#pragma unroll
              for (int n = 0; n < 4; ++n) {
                if (n >= Hd.hn) break;
                float x = hacc[n] + hbias[n];
                if (Hd.post == 1) {
                  float z = x + Hd.shift;
                  x = z > 20.f ? z : log1pf(expf(z));
                } else if (Hd.post == 2) {
                  x = (1.f / (1.f + expf(-x))) * (1.f + 2.f * Hd.shift) - Hd.shift;
                } else if (Hd.post == 3) {
                  x = args.add[row * Hd.hn + n] + x;
                } else if (Hd.post == 4) {
                  x = (n < 3) ? 1.f / (1.f + expf(-x)) : fmaxf(x, 0.f);
                }
                o[n] = x;
              }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          }
        }
        if (tl) args.timeline[u * 12 + 5] = clock64();
       }
      }
    }
  } else if (args.fused_ipe == 2 && warp >= kFeatWarp0) {
    // ===================== feature generators, Fourier prologue (human branch) =====================
    // identity | per octave k: w_k sin(2^k p) (3), w_k cos(2^k p) (3)  (fourier.py:13-57, hannw_fourier.py:15-71): one sincos per
    // coordinate, higher octaves by angle doubling (error 2^k x 1e-7, below the fp16 resolution of the operand); one 64-column
    // chunk per reading layer, so the encoded points never exist outside shared memory.
    const int r = (warp - kFeatWarp0) * 32 + lane;
    const uint32_t xremote = rank != 0 ? mapa_u32(smem_u32(&bar_xfull[0]), 0) : 0u;
    const uint32_t row_u32 = smem_u32(sRingX) + 128u * (uint32_t)r, r7 = (uint32_t)(r & 7);
    uint32_t xi = 0;
    for (int g0 = cluster; g0 < n_groups; g0 += T * n_clusters) {
      float pa[3] = {0.f, 0.f, 0.f}, pb[3] = {0.f, 0.f, 0.f};
      const int64_t row_a = (int64_t)(2 * g0 + (int)rank) * kTileM + r;
      const int64_t row_b = (int64_t)(2 * (g0 + n_clusters) + (int)rank) * kTileM + r;
      const bool ok_a = row_a < args.rows, ok_b = T == 2 && g0 + n_clusters < n_groups && row_b < args.rows;
      if (ok_a) { pa[0] = args.fx[row_a * 3 + 0]; pa[1] = args.fx[row_a * 3 + 1]; pa[2] = args.fx[row_a * 3 + 2]; }
      if (ok_b) { pb[0] = args.fx[row_b * 3 + 0]; pb[1] = args.fx[row_b * 3 + 1]; pb[2] = args.fx[row_b * 3 + 2]; }
      for (int l = 0; l < n_layers; ++l) {
       if (prog.layers[l].kb_x == 0) continue;
       for (int t = 0; t < T; ++t) {
        // the row's 64 fp16 features: re-encoded per reading layer (3 sincos + doublings), so that two tile pairs in flight do not
        // hold two encoded rows in registers
        const float p[3] = {t ? pb[0] : pa[0], t ? pb[1] : pa[1], t ? pb[2] : pa[2]};
        const bool ok = t ? ok_b : ok_a;
        uint32_t pk[32];
        if (args.f_ident) fourier_row<true>(p, args.f_nfreq, args.f_hann, ok, pk);
        else fourier_row<false>(p, args.f_nfreq, args.f_hann, ok, pk);
        const int xs = xi % kPairStagesX;
        mbar_wait_guard<40>(&bar_xempty[xs], ((xi / kPairStagesX) & 1) ^ 1);
        const uint32_t slot = row_u32 + (uint32_t)xs * kXChunkBytes;
#pragma unroll
        for (int gq = 0; gq < 8; ++gq) sts128(slot + (((uint32_t)gq ^ r7) << 4), pk[4 * gq], pk[4 * gq + 1], pk[4 * gq + 2], pk[4 * gq + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (xremote) mbar_arrive_remote(xremote + 8u * (uint32_t)xs);
          else mbar_arrive(&bar_xfull[xs]);
        }
        ++xi;
       }
      }
    }
  } else if (args.fused_ipe == 1 && warp >= kFeatWarp0) {
    // ===================== feature generators (warps 12..15): fused IPE prologue =====================
    const int r = (warp - kFeatWarp0) * 32 + lane;
    const uint32_t xremote = rank != 0 ? mapa_u32(smem_u32(&bar_xfull[0]), 0) : 0u;
    uint32_t xi = 0, u = 0;
    for (int g = cluster; g < n_groups; g += n_clusters) {
      const int64_t row = (int64_t)(2 * g + (int)rank) * kTileM + r;
      for (int l = 0; l < n_layers; ++l, ++u) {
        if (prog.layers[l].kb_x == 0) continue;
        const bool tl = args.timeline && blockIdx.x == 0 && threadIdx.x == kFeatWarp0 * 32 && u < 64;
        if (tl) args.timeline[u * 12 + 6] = clock64();
        IpeRowGeom G;
        ipe_row_setup(ipe, row, row < args.rows, G);
        ipe_generate_pass<kPairStagesX, true>(ipe, G, r, sRingX, bar_xfull, bar_xempty, xi, xremote,
                                              (tl && (u == 9 || u == 14)) ? args.timeline + 900 + (u == 14 ? 20 : 0) : nullptr);
        if (tl) args.timeline[u * 12 + 7] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                   // no CTA leaves (or frees TMEM) while its partner may still signal it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols));
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1)
mlp_pair_kernel(const __grid_constant__ MlpProgram prog, const __grid_constant__ MlpArgs args,
                const __grid_constant__ IpeArgs ipe, const int stages_w) {
  mlp_pair_body<false>(prog, args, ipe, stages_w);
}
// two tile pairs in flight per cluster (narrow networks), see mlp_pair_body
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1)
mlp_duo_kernel(const __grid_constant__ MlpProgram prog, const __grid_constant__ MlpArgs args,
               const __grid_constant__ IpeArgs ipe, const int stages_w) {
  mlp_pair_body<true>(prog, args, ipe, stages_w);
}

// ----------------------------------------------------------------------------- packing
// W fp32 [N, in_h + in_x] (nn.Linear layout, columns ordered [x|h] if x_first else [h|x])
// -> fp16 chunks in kernel K order: kb_h chunks of h columns, then kb_x chunks of x columns.
// ipe_perm bit 0 (x columns) / bit 1 (h columns): that input holds generated IPE features, re-ordered from the reference layout
// f = half*252 + l*21 + j  to the kernel's generation order
// col = 2 * (pbase_g + l * gs_g + jj) + half,  j = jbase_g + jj  (groups of 8, 8, 5 directions; see ipe_generate_pass).
__global__ void pack_weight_kernel(const float* __restrict__ W, int N, int in_h, int in_x, int x_first, int kb_h,
                                   int kb_x, int ipe_perm, unsigned char* __restrict__ dst) {
  const int nkb = kb_h + kb_x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)nkb * N * kKB) return;
  int kk = (int)(i % kKB);
  int n = (int)((i / kKB) % N);
  int kb = (int)(i / ((int64_t)kKB * N));
  float v = 0.f;
  const int ktot = in_h + in_x;
  auto ipe_col = [](int c) {          // kernel generation order -> reference feature index
    const int half = c & 1, pr = c >> 1;
    const int gs = pr < 192 ? 8 : 5, pb = pr < 96 ? 0 : (pr < 192 ? 96 : 192), jb = pr < 96 ? 0 : (pr < 192 ? 8 : 16);
    const int l = (pr - pb) / gs, jj = (pr - pb) % gs;
    return half * (kIpeDeg * kIpeB) + l * kIpeB + jb + jj;
  };
  // the generator emits (-1)^(l & 3) sin(2^l m) (ipe_group): odd octaves of the sine half take the sign here
  auto ipe_sign = [](int c) {
    const int pr = c >> 1;
    const int gs = pr < 192 ? 8 : 5, pb = pr < 96 ? 0 : (pr < 192 ? 96 : 192);
    const int l = (pr - pb) / gs;
    return ((c & 1) == 0 && (l & 1)) ? -1.f : 1.f;
  };
  if (kb < kb_h) {
    int c = kb * kKB + kk;
    if (c < in_h) {
      float sg = 1.f;
      if (ipe_perm & 2) { sg = ipe_sign(c); c = ipe_col(c); }
      v = sg * W[(int64_t)n * ktot + (x_first ? in_x + c : c)];
    }
  } else {
    int c = (kb - kb_h) * kKB + kk;
    if (c < in_x) {
      float sg = 1.f;
      if (ipe_perm & 1) { sg = ipe_sign(c); c = ipe_col(c); }
      v = sg * W[(int64_t)n * ktot + (x_first ? c : in_h + c)];
    }
  }
  *reinterpret_cast<__half*>(dst + (size_t)kb * N * 128 + tile_byte_offset(n, kk)) = __float2half_rn(v);
}

// Bias chunk of a layer (cluster-pair kernel): an [N x 64] fp16 chunk whose K-columns 0 / 1 hold the fp16 hi / lo
// parts of b[n] (hi + lo reproduces the fp32 bias to ~2^-22); b == nullptr: all zero.
__global__ void pack_bias_chunk_kernel(const float* __restrict__ b, int N, unsigned char* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * kKB) return;
  const int n = i / kKB, kk = i % kKB;
  __half v = __float2half_rn(0.f);
  if (b && kk < 2) {
    const float x = b[n];
    const __half hi = __float2half_rn(x);
    v = kk == 0 ? hi : __float2half_rn(x - __half2float(hi));
  }
  *reinterpret_cast<__half*>(dst + tile_byte_offset(n, kk)) = v;
}

// X fp32 [rows, ld] (first K columns) -> tiled fp16 [ntiles][kbx][16 KB], zero padded.
__global__ void pack_rows_kernel(const float* __restrict__ X, int64_t rows, int ld, int K, int kbx,
                                 unsigned char* __restrict__ dst, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int k = (int)(i % (kbx * kKB));
  int64_t row = i / (kbx * kKB);
  float v = (row < rows && k < K) ? X[row * ld + k] : 0.f;
  int64_t tile = row / kTileM;
  int r = (int)(row % kTileM);
  *reinterpret_cast<__half*>(dst + ((size_t)tile * kbx + (k >> 6)) * kXChunkBytes + tile_byte_offset(r, k & 63)) =
      __float2half_rn(v);
}

// ============================================================================ wide-layer GEMM (cluster pair)
// Layers wider than 256 (the reference's default NeRFMLP is 1024 wide) cannot keep a tile's activations on chip
// (128 rows x 1024 fp16 = 256 KB), so they run layer by layer through L2/HBM:
//     Y[rows, N] = act([A0 | A1][rows, K0 + K1] * W^T + b)         A0, A1, Y in the tiled fp16 layout
// with the same building blocks as mlp_pair_kernel: tcgen05.mma.cta_group::2 (M = 256 rows per cluster, N = 256 per
// unit), half of every weight chunk per CTA, mbarrier rings fed by several bulk-copy producer warps per operand
// (one thread's copies serialise), accumulator double-buffered in TMEM.  Units (tile pair, 256-column block) are
// independent, so the epilogue of unit u (bias, ReLU, optional fp32 head, 16-byte stores of the tiled output)
// simply overlaps the MMAs of unit u + 1 - no per-layer dead time here.
constexpr int kGemmStages = 6;          // per operand ring (A: 16 KB stages, W: 16 KB = this CTA's 128 of 256 rows)
constexpr int kGemmNB = 256;            // output columns per unit
constexpr int kGemmBars = 2 * kGemmStages + 4;
constexpr int kGemmWProducers = 3;      // warps 0, 2, 3
constexpr int kGemmAProducers = 4;      // warps 12..15

struct GemmArgs {
  const unsigned char* a0;     // tiled fp16 [ntiles][kb0][16 KB]
  const unsigned char* a1;     // optional second input (skip connection), [ntiles][kb1][16 KB]
  const unsigned char* w;      // packed [N / 256][kb0 + kb1][256 x 128 B]
  const float* params;         // [N] bias, then [hn][N] head weights, then [4] head bias
  unsigned char* y;            // tiled fp16 [ntiles][N / 64][16 KB], or null
  float* head_out;             // [rows][hn] or null
  int64_t rows;
  int ntiles, kb0, kb1, n, relu;
  int hn, head_post;
  float head_shift;
  int group;                   // operand chunks per MMA issue group (1 .. 4)
};

// CS = 4 (gemm_quad_kernel, opt-in): two CTA pairs per cluster work on different row tiles and the SAME weight block.  At
// 256 x 256 tiles a pair pulls 64 B per clock and SM of operands out of L2; in the quad kernel every CTA fetches only HALF of
// its weight stage and multicasts it to its sibling in the other pair (which fetches the other half): 48 B per clock and SM.
// A stage is refilled once BOTH pairs have consumed it (their commits are multicast to all four CTAs).  Bit-identical
// results; measured slower than the pair kernel (see hos_gemm_forward), so it is kept as an A/B switch only.
template <int CS>
__device__ __forceinline__ void gemm_cluster_body(const GemmArgs& args) {
  constexpr bool QUAD = CS == 4;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int kWStage = kGemmNB * 64;                     // 16 KB: this CTA's half of a [256 x 64] weight chunk
  unsigned char* sRingA = smem;                             // [kGemmStages][16 KB]
  unsigned char* sRingW = sRingA + kGemmStages * kXChunkBytes;
  float* sParams = reinterpret_cast<float*>(sRingW + kGemmStages * kWStage);
  const int param_floats = args.n * (1 + args.hn) + 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sParams + ((param_floats + 3) & ~3));
  // The A and W rings move in lock step, so stage s of both shares one pair of barriers:
  uint64_t* bar_full = bars;                                // leader: A + W of both CTAs landed (2 tx arrivals + 1 relay); peer: its own two
  uint64_t* bar_empty = bar_full + kGemmStages;             // multicast commit: stage s of both rings consumed
  uint64_t* bar_tfull = bar_empty + kGemmStages;            // [2] multicast commit: accumulator buffer complete
  uint64_t* bar_tempty = bar_tfull + 2;                     // [2] leader only: drained by the 16 epilogue warps of the pair
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + kGemmBars);
  float* s_headx = reinterpret_cast<float*>(s_tmem + 4);    // [128][4]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();                 // 0 .. CS - 1
  const uint32_t rank = crank & 1u, pair = crank >> 1;      // rank inside the CTA pair (0 = leader: issues the MMAs); pair index
  const uint32_t leader = crank & ~1u;                      // cluster rank of this pair's leader
  const int cluster = blockIdx.x / CS, n_clusters = gridDim.x / CS;
  const int n_groups = (args.ntiles + CS - 1) / CS;         // CS tiles per group: tile = CS g + crank
  const int KB = args.kb0 + args.kb1;
  const int nblk = args.n / kGemmNB;

  for (int i = threadIdx.x; i < param_floats; i += kMlpThreads) sParams[i] = args.params[i];
  if (threadIdx.x == 0) {
    // leader: own A + own W + ONE relay from the peer's watcher warp; peer: its own two
    for (int s = 0; s < kGemmStages; ++s) { mbar_init(&bar_full[s], rank == 0 ? 3u : 2u); mbar_init(&bar_empty[s], CS / 2); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bar_tfull[s], 1); mbar_init(&bar_tempty[s], 2 * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const bool w_producer = warp == 0 || warp == 2 || warp == 3;
  const bool a_producer = warp >= kFeatWarp0;
  if (w_producer || a_producer) {
    // ===================== producers: chunk i of the flat stream -> producer i % P, ring stage i % kGemmStages ========
    if (lane == 0) {
      const uint32_t P = w_producer ? kGemmWProducers : kGemmAProducers;
      const uint32_t p = w_producer ? (warp == 0 ? 0u : (uint32_t)(warp - 1)) : (uint32_t)(warp - kFeatWarp0);
      uint64_t* full = bar_full;
      uint64_t* empty = bar_empty;
      unsigned char* ring = w_producer ? sRingW : sRingA;
      uint32_t ci = 0;
      for (int g = cluster; g < n_groups; g += n_clusters) {
        int tile = CS * g + (int)crank;
        if (tile >= args.ntiles) tile = args.ntiles - 1;           // padding tile: any valid rows, nothing is stored
        for (int j = 0; j < nblk; ++j) {
          for (int kb = 0; kb < KB; ++kb, ++ci) {
            if (ci % P != p) continue;
            const uint32_t st = ci % kGemmStages, use = ci / kGemmStages;
            mbar_wait_guard<100>(&empty[st], (use & 1) ^ 1);
            mbar_expect_tx(&full[st], kXChunkBytes);
            if (w_producer) {
              const unsigned char* src = args.w + ((size_t)j * KB + kb) * (2u * kWStage) + rank * kWStage;
              if (QUAD) {       // this CTA's half of the stage: rows [64 pair, +64) of its 128, delivered to both pairs
                const uint32_t off = pair * (uint32_t)(kWStage / 2);
                bulk_g2s_mcast(ring + st * kXChunkBytes + off, src + off, kWStage / 2, &full[st],
                               (uint16_t)((1u << rank) | (1u << (rank + 2))));
              } else {
                bulk_g2s(ring + st * kXChunkBytes, src, kXChunkBytes, &full[st]);
              }
            } else {
              const unsigned char* src = kb < args.kb0 ? args.a0 + ((size_t)tile * args.kb0 + kb) * kXChunkBytes
                                                       : args.a1 + ((size_t)tile * args.kb1 + (kb - args.kb0)) * kXChunkBytes;
              bulk_g2s(ring + st * kXChunkBytes, src, kXChunkBytes, &full[st]);
            }
          }
        }
      }
    }
  } else if (warp == 1 && rank != 0) {
    // ===================== peer CTA: completion watcher =====================
    // tells the leader "both of my operands of stage s have landed", stage by stage.  A separate warp, so that a producer never
    // waits for a copy to ARRIVE before issuing its next one (with the relay in the producer threads each of them moved one chunk
    // per L2 round trip).
    if (lane == 0) {
      const uint32_t full_leader = mapa_u32(smem_u32(&bar_full[0]), leader);
      uint32_t st = 0, par = 0;
      for (int g = cluster; g < n_groups; g += n_clusters) {
        const int nch = nblk * KB;
        for (int i = 0; i < nch; ++i) {
          mbar_wait_guard<40>(&bar_full[st], par);
          mbar_arrive_remote(full_leader + 8u * st);
          if (++st == kGemmStages) { st = 0; par ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: warp-uniform loop, tcgen05 instructions under elect.sync =====================
    if (rank == 0) {
      uint32_t u = 0, st = 0, par = 0;                 // unit counter; ring cursor + phase parity (A and W rings move in lock step)
      const uint64_t desc_hi = umma_desc(0);
      const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);
      const uint32_t tfull0 = smem_u32(&bar_tfull[0]);
      const uint32_t a16 = (smem_u32(sRingA) >> 4) & 0x3FFF, w16 = (smem_u32(sRingW) >> 4) & 0x3FFF;
      const uint32_t idesc = umma_idesc_f16(kGemmNB, 2 * kTileM);
      for (int g = cluster; g < n_groups; g += n_clusters) {
        for (int j = 0; j < nblk; ++j, ++u) {
          const uint32_t acc = tmem_base + (u & 1) * 256;
          if (u >= 2) mbar_wait_guard<0>(&bar_tempty[u & 1], ((u >> 1) - 1) & 1);     // buffer drained by unit u - 2
          int kb = 0;
          while (kb < KB) {
            // up to four chunks per group: one parallel poll (~220 cycles) and the issue overhead amortised over 2048
            // cycles of tensor work
            const int cnt = KB - kb < args.group ? KB - kb : args.group;
            uint32_t fb[4], fp[4];
            {
              uint32_t s_ = st, p_ = par;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                fb[i] = full0 + 8u * s_; fp[i] = p_;
                if (i + 1 < cnt) { if (++s_ == kGemmStages) { s_ = 0; p_ ^= 1; } }
              }
            }
            mbar_wait4_spin(fb[0], fp[0], fb[1], fp[1], fb[2], fp[2], fb[3], fp[3]);
            tc_fence_after();
            const bool leader_lane = elect_one();
#pragma unroll 1
            for (int i = 0; i < cnt; ++i) {
              if (leader_lane) {
                const uint64_t adesc = desc_hi | (uint64_t)(a16 + st * (kXChunkBytes >> 4));
                const uint64_t bdesc = desc_hi | (uint64_t)(w16 + st * (kWStage >> 4));
#pragma unroll
                for (int k = 0; k < kKB / 16; ++k)
                  tc_mma_f16_pair(acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb + i) != 0 || k != 0 ? 1u : 0u);
                tc_commit_mask_addr(empty0 + 8u * st, QUAD ? (uint16_t)0xF : (uint16_t)0x3);
                if (kb + i == KB - 1) tc_commit_mask_addr(tfull0 + 8u * (u & 1), (uint16_t)(3u << (2 * pair)));
              }
              if (++st == kGemmStages) { st = 0; par ^= 1; }
            }
            __syncwarp();
            kb += cnt;
          }
        }
      }
    }
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + kEpiWarps) {
    // ===================== epilogue (8 warps): bias, ReLU, optional fp32 head, tiled fp16 store =====================
    const int q = warp & 3;
    const int ch = (warp - kEpiWarp0) >> 2;
    const int r = q * 32 + lane;
    const uint32_t tempty0 = mapa_u32(smem_u32(&bar_tempty[0]), leader);
    const uint32_t sparams_u32 = smem_u32(sParams);
    const uint32_t headx_u32 = smem_u32(s_headx) + 16u * r;
    const bool has_head = args.hn > 0 && args.head_out != nullptr;
    uint32_t u = 0;
    for (int g = cluster; g < n_groups; g += n_clusters) {
      const int tile = CS * g + (int)crank;
      const int64_t row = (int64_t)tile * kTileM + r;
      const bool row_ok = row < args.rows, tile_ok = tile < args.ntiles;
      float hacc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < nblk; ++j, ++u) {
        const uint32_t acc = tmem_base + (u & 1) * 256 + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32);
        mbar_wait_guard<20>(&bar_tfull[u & 1], (u >> 1) & 1);
        tc_fence_after();
        for (int c = 0; c < kGemmNB / 64; ++c) {
          const int n0 = j * kGemmNB + c * 64 + ch * 32;           // first of this thread's 32 output columns
          uint32_t v[32];
          tmem_ld32_nowait(acc + (uint32_t)(c * 64), v);
          tmem_wait_ld();
          float f[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = lds128(sparams_u32 + 4u * (uint32_t)n0 + 16u * i);
            f[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + bb.x; f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + bb.y;
            f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + bb.z; f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + bb.w;
          }
          if (args.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
          }
          if (args.y && tile_ok) {
            unsigned char* dst = args.y + ((size_t)tile * (args.n >> 6) + (size_t)(j * (kGemmNB / 64) + c)) * kXChunkBytes;
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              uint4 pk;
              pk.x = cvt_f16x2(__float_as_uint(f[gq * 8 + 0]), __float_as_uint(f[gq * 8 + 1]));
              pk.y = cvt_f16x2(__float_as_uint(f[gq * 8 + 2]), __float_as_uint(f[gq * 8 + 3]));
              pk.z = cvt_f16x2(__float_as_uint(f[gq * 8 + 4]), __float_as_uint(f[gq * 8 + 5]));
              pk.w = cvt_f16x2(__float_as_uint(f[gq * 8 + 6]), __float_as_uint(f[gq * 8 + 7]));
              *reinterpret_cast<uint4*>(dst + tile_byte_offset(r, ch * 32 + gq * 8)) = pk;
            }
          }
          if (has_head) {
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              if (h < args.hn) {
                const uint32_t w4 = sparams_u32 + 4u * (uint32_t)(args.n * (1 + h) + n0);
                float a = hacc[h];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 w = lds128(w4 + 16u * i);
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
        if (ch == 1) sts128(headx_u32, __float_as_uint(hacc[0]), __float_as_uint(hacc[1]), __float_as_uint(hacc[2]), __float_as_uint(hacc[3]));
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (ch == 0 && row_ok) {
          const float4 o4 = lds128(headx_u32);
          hacc[0] += o4.x; hacc[1] += o4.y; hacc[2] += o4.z; hacc[3] += o4.w;
          const float4 hb = lds128(sparams_u32 + 4u * (uint32_t)(args.n * (1 + args.hn)));
          const float hbias[4] = {hb.x, hb.y, hb.z, hb.w};
          float* o = args.head_out + row * args.hn;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            if (h >= args.hn) break;
            float x = hacc[h] + hbias[h];
            if (args.head_post == 1) {
              float z = x + args.head_shift;
              x = z > 20.f ? z : log1pf(expf(z));
            } else if (args.head_post == 2) {
              x = (1.f / (1.f + expf(-x))) * (1.f + 2.f * args.head_shift) - args.head_shift;
            } else if (args.head_post == 4) {
              x = (h < 3) ? 1.f / (1.f + expf(-x)) : fmaxf(x, 0.f);
            }
            o[h] = x;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols));
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1)
gemm_pair_kernel(const __grid_constant__ GemmArgs args) { gemm_cluster_body<2>(args); }
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(kMlpThreads, 1)
gemm_quad_kernel(const __grid_constant__ GemmArgs args) { gemm_cluster_body<4>(args); }


// Generated IPE features to HBM in the tiled fp16 layout (kernel column order, see ipe_generate_pass): the wide-layer
// path reads them as a GEMM operand.  One thread per sample, the same packed recurrences as the fused prologue, one
// 16-byte store per four (sin, cos) pairs; ~20x faster than the accurate fp32 parity kernel (hos_ipe_features).
__global__ void __launch_bounds__(kTileM)
ipe_features_fast_kernel(const __grid_constant__ IpeArgs ipe, int64_t rows, unsigned char* __restrict__ out) {
  const int r = threadIdx.x;
  const int64_t tile = blockIdx.x;
  const int64_t row = tile * kTileM + r;
  IpeRowGeom G;
  ipe_row_setup(ipe, row, row < rows, G);
  unsigned char* base = out + (size_t)tile * 8 * kXChunkBytes + (size_t)r * 128;
  const uint32_t r7 = (uint32_t)(r & 7);
  uint32_t pk[4];
  auto emit = [&](int p, float sv, float cv) {             // p is a compile-time constant after unrolling
    pk[p & 3] = cvt_f16x2(__float_as_uint(sv), __float_as_uint(cv));
    if ((p & 3) == 3)
      *reinterpret_cast<uint4*>(base + (size_t)(p >> 5) * kXChunkBytes + ((((uint32_t)(p & 31) >> 2) ^ r7) << 4)) =
          make_uint4(pk[0], pk[1], pk[2], pk[3]);
  };
  ipe_group<8, 0, 0>(ipe, G, emit);
  ipe_group<8, 8, 96>(ipe, G, emit);
  ipe_group<5, 16, 192>(ipe, G, emit);
  *reinterpret_cast<uint4*>(base + (size_t)7 * kXChunkBytes + ((7u ^ r7) << 4)) = make_uint4(0u, 0u, 0u, 0u);   // columns 504..511
}
}  // namespace hos

using namespace hos;

struct hos_mlp {
  MlpProgram prog;
  int in_dim;
  std::vector<hos_mlp_layer> layers;
  std::vector<hos_mlp_head> heads;
  unsigned char* d_w = nullptr;      // packed fp16 weights
  float* d_params = nullptr;         // biases + head weights (fp32)
  size_t w_bytes = 0;
  size_t smem_bytes = 0;
  size_t smem_pair = 0;    // shared memory of the cluster-pair kernel; 0: this program only runs on mlp_tc_kernel
  int pair_stages = 0;     // depth of its weight ring
  int max_clusters = 0;    // co-resident 2-CTA clusters (persistent grid of the pair kernel)
  int ipe_perm = 0;        // weights packed for the fused-IPE column order
  int variant = 0;         // kernel selection of THIS handle: 0 automatic, 1 single-CTA kernel, 2 cluster-pair kernel,
                           // 3 cluster-pair kernel with one tile pair in flight (A/B against the duo schedule)
  int duo_ok = 0;          // the pair kernel can keep two tile pairs in flight for this program (narrow layers)
  long long* timeline = nullptr;   // optional debug buffer of THIS handle (hos_mlp_debug_timeline)
};


extern "C" {

hos_mlp_t* hos_mlp_create(int in_dim, int n_layers, const hos_mlp_layer* layers, int n_heads,
                          const hos_mlp_head* heads) {
  if (hos::check_arch() != HOS_OK) return nullptr;
  if (!layers || n_layers < 2 || n_layers > kMaxLayers || n_heads < 0 || n_heads > kMaxHeads || in_dim < 1) {
    hos::set_error("hos_mlp_create: bad layer/head count");
    return nullptr;
  }
  hos_mlp* m = new hos_mlp();
  m->in_dim = in_dim;
  m->layers.assign(layers, layers + n_layers);
  if (n_heads) m->heads.assign(heads, heads + n_heads);
  MlpProgram& P = m->prog;
  memset(&P, 0, sizeof(P));
  P.n_layers = n_layers;
  P.n_heads = n_heads;
  P.kbx = (in_dim + kKB - 1) / kKB;
  uint32_t woff = 0, poff = 0;
  int width = 0, nmax = 0;
  for (int l = 0; l < n_layers; ++l) {
    const hos_mlp_layer& L = layers[l];
    bool ok = L.out_dim >= 16 && L.out_dim <= 256 && (L.out_dim % 16) == 0 && L.in_h >= 0 && L.in_x >= 0 &&
              (L.in_h + L.in_x) > 0 && L.in_x <= in_dim && (L.in_h % kKB) == 0 &&
              (l == 0 ? L.in_h == 0 : L.in_h == layers[l - 1].out_dim) && L.head < n_heads;
    if (!ok) {
      hos::set_error("hos_mlp_create: unsupported layer %d (out=%d in_h=%d in_x=%d)", l, L.out_dim, L.in_h, L.in_x);
      delete m;
      return nullptr;
    }
    LayerDev& D = P.layers[l];
    D.n = (uint16_t)L.out_dim;
    D.kb_h = (uint8_t)(L.in_h / kKB);
    D.kb_x = (uint8_t)((L.in_x + kKB - 1) / kKB);
    D.relu = (uint8_t)(L.relu != 0);
    D.rowbias = (uint8_t)(L.rowbias != 0);
    D.head = (int8_t)L.head;
    D.w_off = woff;
    D.bias_off = poff;                      // multiples of 16 floats: the epilogue reads float4
    woff += (uint32_t)(D.kb_h + D.kb_x + 1) * D.n * 128u;   // + the bias chunk read by the cluster-pair kernel
    poff += D.n;
    if (l < n_layers - 1 && L.out_dim > width) width = L.out_dim;
    if (L.out_dim > nmax) nmax = L.out_dim;
    if (L.head >= 0 && heads[L.head].out_dim > 4) {
      hos::set_error("hos_mlp_create: head width > 4");
      delete m;
      return nullptr;
    }
  }
  P.head_base = (int)poff;
  for (int h = 0; h < n_heads; ++h) {
    int owner = -1;
    for (int l = 0; l < n_layers; ++l) if (layers[l].head == h) owner = l;
    if (owner < 0) { hos::set_error("hos_mlp_create: head %d unused", h); delete m; return nullptr; }
    HeadDev& H = P.heads[h];
    H.hn = (uint8_t)heads[h].out_dim;
    H.post = (uint8_t)heads[h].post;
    H.shift = heads[h].shift;
    H.slot = (uint8_t)heads[h].out_slot;
    H.w_off = poff;                         // [hn][n], n % 16 == 0 keeps every row float4-aligned
    poff += (uint32_t)H.hn * layers[owner].out_dim;
    H.b_off = poff;
    poff += 4;
  }
  if (width == 0) width = kKB;
  P.kbh = (width + kKB - 1) / kKB;
  P.n_max = nmax;
  P.param_floats = (int)poff;
  m->w_bytes = woff;
  m->smem_bytes = 1024 + (size_t)P.kbh * kXChunkBytes + (size_t)kStages * nmax * 128 +
                  (size_t)kStagesX * kXChunkBytes + (((size_t)poff + 3) & ~(size_t)3) * 4 +
                  (2 * kStages + 2 * kStagesX + 4) * 8 + 16 + kTileM * 4 * sizeof(float);
  if (m->smem_bytes > 227 * 1024) {
    hos::set_error("hos_mlp_create: needs %zu B shared memory (> 227 KB)", m->smem_bytes);
    delete m;
    return nullptr;
  }
  // cluster-pair kernel: every layer width a multiple of 64 (two CTAs x two epilogue column halves x 32-column
  // TMEM loads), two activation slots + half-width weight stages must fit
  bool pair_ok = true;
  for (int l = 0; l < n_layers; ++l) pair_ok = pair_ok && (layers[l].out_dim % 64) == 0;
  for (int l = 0; l + 1 < n_layers; ++l) pair_ok = pair_ok && layers[l].out_dim == P.kbh * kKB;   // one hready phase count
  pair_ok = pair_ok && P.kbh <= kPairMaxKbh;
  // duo schedule (two tile pairs in flight): four 128-column accumulators, so every layer <= 128 wide; no per-ray bias and no
  // fp32 head on a hidden layer (their epilogues hold the accumulator beyond the last operand chunk)
  bool duo_ok = pair_ok && nmax <= 128;
  for (int l = 0; l < n_layers; ++l) duo_ok = duo_ok && !layers[l].rowbias && (layers[l].head < 0 || l == n_layers - 1);
  const size_t pair_fixed = 1024 + (size_t)(duo_ok ? 2 : 1) * P.kbh * kXChunkBytes + kOnesBytes + (size_t)kPairStagesX * kXChunkBytes +
                            (((size_t)(poff - (uint32_t)P.head_base) + 3) & ~(size_t)3) * 4 + (kPairBars + 1) * 8 + 16 +
                            kTileM * 4 * sizeof(float) +
                            kRowBiasVecs * 128 * sizeof(float);
  int pair_stages = pair_fixed < 227 * 1024 ? (int)((227 * 1024 - pair_fixed) / ((size_t)nmax * 64)) : 0;
  if (pair_stages > kPairMaxStagesW) pair_stages = kPairMaxStagesW;
  pair_ok = pair_ok && pair_stages >= 3;
  const size_t smem_pair = pair_fixed + (size_t)pair_stages * nmax * 64;
  m->pair_stages = pair_stages;
  m->duo_ok = duo_ok && pair_ok && pair_stages >= 3;
  if (pair_ok && smem_pair <= 227 * 1024) {
    m->smem_pair = smem_pair;
    if (cudaFuncSetAttribute(mlp_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)) != cudaSuccess ||
        cudaFuncSetAttribute(mlp_duo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)) != cudaSuccess) {
      cudaGetLastError();
      m->smem_pair = 0;
    } else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(kNumSMs);
      cfg.blockDim = dim3(kMlpThreads);
      cfg.dynamicSmemBytes = m->smem_pair;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, mlp_pair_kernel, &cfg) != cudaSuccess || nc < 1) {
        cudaGetLastError();
        nc = kNumSMs / 2;
      }
      m->max_clusters = nc < kNumSMs / 2 ? nc : kNumSMs / 2;
    }
  }
  if (cudaMalloc(&m->d_w, m->w_bytes) != cudaSuccess || cudaMalloc(&m->d_params, (size_t)poff * 4) != cudaSuccess ||
      cudaMemset(m->d_params, 0, (size_t)poff * 4) != cudaSuccess ||
      cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)) != cudaSuccess) {
    hos::set_error("hos_mlp_create: CUDA allocation/attribute failed: %s", cudaGetErrorString(cudaGetLastError()));
    hos_mlp_destroy(m);
    return nullptr;
  }
  return m;
}

void hos_mlp_destroy(hos_mlp_t* m) {
  if (!m) return;
  if (m->d_w) cudaFree(m->d_w);
  if (m->d_params) cudaFree(m->d_params);
  delete m;
}

int hos_mlp_set_layer(hos_mlp_t* m, int layer, const float* W, const float* b, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(m && W && layer >= 0 && layer < m->prog.n_layers, "hos_mlp_set_layer: bad handle/layer");
  const hos_mlp_layer& L = m->layers[layer];
  const LayerDev& D = m->prog.layers[layer];
  cudaStream_t st = (cudaStream_t)stream;
  int64_t tot = (int64_t)(D.kb_h + D.kb_x) * D.n * kKB;
  pack_weight_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(W, D.n, L.in_h, L.in_x, L.x_first, D.kb_h, D.kb_x,
                                                                   m->ipe_perm, m->d_w + D.w_off);
  HOS_LAUNCH_CHECK();
  pack_bias_chunk_kernel<<<(D.n * kKB + 255) / 256, 256, 0, st>>>(b, D.n, m->d_w + D.w_off + (size_t)(D.kb_h + D.kb_x) * D.n * 128);
  HOS_LAUNCH_CHECK();
  if (b) HOS_CUDA(cudaMemcpyAsync(m->d_params + D.bias_off, b, (size_t)D.n * 4, cudaMemcpyDeviceToDevice, st));
  else HOS_CUDA(cudaMemsetAsync(m->d_params + D.bias_off, 0, (size_t)D.n * 4, st));
  return HOS_OK;
}

int hos_mlp_set_bias(hos_mlp_t* m, int layer, const float* b, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(m && b && layer >= 0 && layer < m->prog.n_layers, "hos_mlp_set_bias: bad handle/layer");
  const LayerDev& D = m->prog.layers[layer];
  pack_bias_chunk_kernel<<<(D.n * kKB + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      b, D.n, m->d_w + D.w_off + (size_t)(D.kb_h + D.kb_x) * D.n * 128);
  HOS_LAUNCH_CHECK();
  HOS_CUDA(cudaMemcpyAsync(m->d_params + D.bias_off, b, (size_t)D.n * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return HOS_OK;
}

int hos_mlp_set_head(hos_mlp_t* m, int head, const float* W, const float* b, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(m && W && head >= 0 && head < m->prog.n_heads, "hos_mlp_set_head: bad handle/head");
  const HeadDev& H = m->prog.heads[head];
  cudaStream_t st = (cudaStream_t)stream;
  HOS_CUDA(cudaMemcpyAsync(m->d_params + H.w_off, W, (size_t)(H.b_off - H.w_off) * 4, cudaMemcpyDeviceToDevice, st));
  if (b) HOS_CUDA(cudaMemcpyAsync(m->d_params + H.b_off, b, (size_t)H.hn * 4, cudaMemcpyDeviceToDevice, st));
  return HOS_OK;
}

int hos_mlp_in_kblocks(const hos_mlp_t* m) { return m ? m->prog.kbx : 0; }

struct FourierArgs {
  const float* x;
  int nfreq, ident;
  float hann[16];
};

static int mlp_launch(hos_mlp_t* m, const void* x_tiled, const IpeArgs* ipe, int64_t rows, const float* rowbias,
                      int rowbias_div, const float* add, float* out0, float* out1, void* stream, const FourierArgs* fe = nullptr) {
  for (int l = 0; l < m->prog.n_layers; ++l)
    HOS_REQUIRE(!m->prog.layers[l].rowbias || (rowbias && rowbias_div >= 1), "hos_mlp_forward: layer %d needs rowbias", l);
  for (int h = 0; h < m->prog.n_heads; ++h) {
    HOS_REQUIRE(m->prog.heads[h].slot < 2 && (m->prog.heads[h].slot == 0 ? out0 : out1), "hos_mlp_forward: missing output for head %d", h);
    HOS_REQUIRE(m->prog.heads[h].post != 3 || add, "hos_mlp_forward: head %d needs `add`", h);
  }
  if (rows == 0) return HOS_OK;
  MlpArgs a;
  a.x_tiled = (const unsigned char*)x_tiled;
  a.w_packed = m->d_w;
  a.params = m->d_params;
  a.rowbias = rowbias;
  a.add = add;
  a.out[0] = out0;
  a.out[1] = out1;
  a.rows = rows;
  a.ntiles = (int)((rows + kTileM - 1) / kTileM);
  a.rowbias_div = rowbias_div < 1 ? 1 : rowbias_div;
  a.fused_ipe = ipe != nullptr;
  a.fx = nullptr; a.f_nfreq = 0; a.f_ident = 0;
  for (int i = 0; i < 16; ++i) a.f_hann[i] = 1.f;
  if (fe) {
    a.fused_ipe = 2;
    a.fx = fe->x; a.f_nfreq = fe->nfreq; a.f_ident = fe->ident;
    for (int i = 0; i < 16; ++i) a.f_hann[i] = fe->hann[i];
  }
  a.timeline = m->timeline;
  // two tile pairs in flight when the program allows it and every cluster still gets at least two tile pairs
  a.duo = m->duo_ok && m->variant != 3 && a.fused_ipe != 1 && (a.ntiles + 1) / 2 >= 2 * m->max_clusters;
  static const IpeArgs kNoIpe = {};
  HOS_REQUIRE(m->variant < 2 || m->smem_pair, "hos_mlp_forward: the cluster-pair kernel does not support this program");
  // the pair kernel walks groups of 4 tiles; tiny batches keep more SMs busy on the single-CTA kernel
  const bool pair = m->smem_pair && m->variant != 1 && (m->variant >= 2 || a.ntiles >= 2);
  HOS_REQUIRE(!fe || pair, "hos_mlp_forward_fourier: needs the cluster-pair kernel (uniform hidden width, >= 2 row tiles)");
  if (pair) {
    const int n_groups = (a.ntiles + 1) / 2;
    const int clusters = n_groups < m->max_clusters ? n_groups : m->max_clusters;
    if (a.duo)
      mlp_duo_kernel<<<2 * clusters, kMlpThreads, m->smem_pair, (cudaStream_t)stream>>>(m->prog, a, ipe ? *ipe : kNoIpe, m->pair_stages);
    else
      mlp_pair_kernel<<<2 * clusters, kMlpThreads, m->smem_pair, (cudaStream_t)stream>>>(m->prog, a, ipe ? *ipe : kNoIpe,
                                                                                        m->pair_stages);
  } else {
    int grid = a.ntiles < kNumSMs ? a.ntiles : kNumSMs;
    mlp_tc_kernel<<<grid, kMlpThreads, m->smem_bytes, (cudaStream_t)stream>>>(m->prog, a, ipe ? *ipe : kNoIpe);
  }
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_mlp_forward(hos_mlp_t* m, const void* x_tiled, int64_t rows, const float* rowbias, int rowbias_div,
                    const float* add, float* out0, float* out1, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(m && x_tiled && rows >= 0, "hos_mlp_forward: bad handle/input");
  HOS_REQUIRE(!m->ipe_perm, "hos_mlp_forward: this MLP was created for the fused IPE prologue (use hos_mlp_forward_ipe)");
  return mlp_launch(m, x_tiled, nullptr, rows, rowbias, rowbias_div, add, out0, out1, stream);
}

int hos_mlp_fourier_supported(const hos_mlp_t* m, int64_t rows) {
  return m && m->smem_pair && m->variant != 1 && !m->ipe_perm && m->prog.kbx == 1 && (m->variant >= 2 || (rows + kTileM - 1) / kTileM >= 2);
}

int hos_mlp_forward_fourier(hos_mlp_t* m, const float* x, int64_t rows, int n_freqs, int include_input, const float* hann_host,
                            const float* add, float* out0, float* out1, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(m && x && rows >= 0 && n_freqs >= 1 && n_freqs <= 10, "hos_mlp_forward_fourier: bad arguments (1 <= n_freqs <= 10)");
  HOS_REQUIRE(m->in_dim == (include_input ? 3 : 0) + 6 * n_freqs && m->prog.kbx == 1 && !m->ipe_perm,
              "hos_mlp_forward_fourier: the MLP reads %d encoded columns, the prologue produces %d", m->in_dim, (include_input ? 3 : 0) + 6 * n_freqs);
  FourierArgs fe;
  fe.x = x; fe.nfreq = n_freqs; fe.ident = include_input != 0;
  for (int i = 0; i < 16; ++i) fe.hann[i] = (hann_host && i < n_freqs) ? hann_host[i] : 1.f;
  return mlp_launch(m, nullptr, nullptr, rows, nullptr, 1, add, out0, out1, stream, &fe);
}

int hos_mlp_set_variant(hos_mlp_t* m, int variant) {
  HOS_REQUIRE(m && variant >= 0 && variant <= 3,
              "hos_mlp_set_variant: handle + 0 = auto, 1 = single-CTA kernel, 2 = cluster-pair kernel, 3 = pair kernel, one tile pair in flight");
  m->variant = variant;
  return HOS_OK;
}

int hos_mlp_debug_timeline(hos_mlp_t* m, long long* device_buf_1024) {
  HOS_REQUIRE(m, "hos_mlp_debug_timeline: null handle");
  m->timeline = device_buf_1024;
  return HOS_OK;
}

int hos_mlp_set_ipe_input(hos_mlp_t* m, int enable) {
  HOS_REQUIRE(m, "hos_mlp_set_ipe_input: null handle");
  HOS_REQUIRE(!enable || m->in_dim == 2 * kIpeDeg * kIpeB, "hos_mlp_set_ipe_input: the fused prologue produces %d features, MLP reads %d",
              2 * kIpeDeg * kIpeB, m->in_dim);
  m->ipe_perm = enable ? 1 : 0;
  return HOS_OK;
}

int hos_mlp_forward_ipe(hos_mlp_t* m, const float* tdist, const float* rays_o, const float* rays_d,
                        const float* radii, const float* basis_host, int N, int S, const float* rowbias,
                        int rowbias_div, float* out0, float* out1, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(m && tdist && rays_o && rays_d && radii && basis_host && N >= 0 && S >= 1, "hos_mlp_forward_ipe: bad arguments");
  HOS_REQUIRE(m->ipe_perm, "hos_mlp_forward_ipe: call hos_mlp_set_ipe_input(mlp, 1) before uploading the weights");
  IpeArgs ipe;
  ipe.tdist = tdist;
  ipe.rays_o = rays_o;
  ipe.rays_d = rays_d;
  ipe.radii = radii;
  ipe.S = S;
  for (int i = 0; i < 3 * kIpeB; ++i) ipe.basis[i] = basis_host[i];
  return mlp_launch(m, nullptr, &ipe, (int64_t)N * S, rowbias, rowbias_div, nullptr, out0, out1, stream);
}

int hos_ipe_features_fast(const float* tdist, const float* rays_o, const float* rays_d, const float* radii,
                          const float* basis_host, int N, int S, void* feat_tiled, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(tdist && rays_o && rays_d && radii && basis_host && feat_tiled && N >= 0 && S >= 1, "hos_ipe_features_fast: bad arguments");
  const int64_t rows = (int64_t)N * S;
  if (rows == 0) return HOS_OK;
  IpeArgs ipe;
  ipe.tdist = tdist;
  ipe.rays_o = rays_o;
  ipe.rays_d = rays_d;
  ipe.radii = radii;
  ipe.S = S;
  for (int i = 0; i < 3 * kIpeB; ++i) ipe.basis[i] = basis_host[i];
  const int64_t ntiles = (rows + kTileM - 1) / kTileM;
  ipe_features_fast_kernel<<<(unsigned)ntiles, kTileM, 0, (cudaStream_t)stream>>>(ipe, rows, (unsigned char*)feat_tiled);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_pack_rows_f16(const float* X, int64_t rows, int ld, int K, void* dst_tiled, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(X && dst_tiled && rows >= 0 && K >= 1 && ld >= K, "hos_pack_rows_f16: bad arguments");
  if (rows == 0) return HOS_OK;
  int kbx = (K + kKB - 1) / kKB;
  int64_t ntiles = (rows + kTileM - 1) / kTileM;
  int64_t total = ntiles * kTileM * (int64_t)kbx * kKB;
  pack_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      X, rows, ld, K, kbx, (unsigned char*)dst_tiled, total);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

// ----------------------------------------------------------------------------- wide-layer GEMM handle
struct hos_gemm {
  int n = 0, k0 = 0, k1 = 0, kb0 = 0, kb1 = 0, x_first = 0, hn = 0, ipe_mask = 0;
  unsigned char* d_w = nullptr;     // packed fp16 [n / 256][kb0 + kb1][256 x 128 B]
  float* d_params = nullptr;        // [n] bias | [4][n] head weights | [4] head bias
  size_t smem_bytes = 0;
  int max_clusters = 0;        // co-resident CTA pairs (gemm_pair_kernel)
  int max_quads = 0;           // co-resident 4-CTA clusters (gemm_quad_kernel); 0: not available
  int cluster = 0;             // kernel selection of THIS handle: 0 automatic, 2 pair kernel, 4 quad kernel
};

hos_gemm_t* hos_gemm_create(int n_out, int k0, int k1, int x_first, int ipe_inputs) {
  if (hos::check_arch() != HOS_OK) return nullptr;
  if (((ipe_inputs & 1) && k0 != 2 * kIpeDeg * kIpeB) || ((ipe_inputs & 2) && k1 != 2 * kIpeDeg * kIpeB)) {
    hos::set_error("hos_gemm_create: an IPE input must have %d features", 2 * kIpeDeg * kIpeB);
    return nullptr;
  }
  if (n_out < kGemmNB || (n_out % kGemmNB) != 0 || n_out > 4096 || k0 < 1 || k1 < 0) {
    hos::set_error("hos_gemm_create: n_out must be a multiple of %d (got %d), k0 >= 1, k1 >= 0", kGemmNB, n_out);
    return nullptr;
  }
  hos_gemm* m = new hos_gemm();
  m->n = n_out; m->k0 = k0; m->k1 = k1; m->x_first = x_first;
  // bit 0: A0 / bit 1: A1 hold generated IPE features (hos_ipe_features_fast) -> pack_weight_kernel's h / x flags
  m->ipe_mask = ((ipe_inputs & 1) ? 2 : 0) | ((ipe_inputs & 2) ? 1 : 0);
  m->kb0 = (k0 + kKB - 1) / kKB;
  m->kb1 = (k1 + kKB - 1) / kKB;
  const size_t wbytes = (size_t)(n_out / kGemmNB) * (m->kb0 + m->kb1) * kGemmNB * 128;
  const size_t pfloats = (size_t)n_out * 5 + 4;
  m->smem_bytes = 1024 + (size_t)kGemmStages * (kXChunkBytes + kGemmNB * 64) + ((pfloats + 3) & ~(size_t)3) * 4 +
                  kGemmBars * 8 + 16 + kTileM * 4 * sizeof(float);
  if (m->smem_bytes > 227 * 1024) {
    hos::set_error("hos_gemm_create: needs %zu B shared memory (> 227 KB)", m->smem_bytes);
    delete m;
    return nullptr;
  }
  if (cudaMalloc(&m->d_w, wbytes) != cudaSuccess || cudaMalloc(&m->d_params, pfloats * 4) != cudaSuccess ||
      cudaMemset(m->d_params, 0, pfloats * 4) != cudaSuccess ||
      cudaFuncSetAttribute(gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)) != cudaSuccess ||
      cudaFuncSetAttribute(gemm_quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)) != cudaSuccess) {
    hos::set_error("hos_gemm_create: CUDA allocation/attribute failed: %s", cudaGetErrorString(cudaGetLastError()));
    hos_gemm_destroy(m);
    return nullptr;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kNumSMs);
  cfg.blockDim = dim3(kMlpThreads);
  cfg.dynamicSmemBytes = m->smem_bytes;
  int nc = 0;
  if (cudaOccupancyMaxActiveClusters(&nc, gemm_pair_kernel, &cfg) != cudaSuccess || nc < 1) {
    cudaGetLastError();
    nc = kNumSMs / 2;
  }
  m->max_clusters = nc < kNumSMs / 2 ? nc : kNumSMs / 2;
  int nq = 0;
  if (cudaOccupancyMaxActiveClusters(&nq, gemm_quad_kernel, &cfg) != cudaSuccess || nq < 1) {
    cudaGetLastError();
    nq = 0;
  }
  m->max_quads = nq < kNumSMs / 4 ? nq : kNumSMs / 4;
  return m;
}

void hos_gemm_destroy(hos_gemm_t* m) {
  if (!m) return;
  if (m->d_w) cudaFree(m->d_w);
  if (m->d_params) cudaFree(m->d_params);
  delete m;
}

int hos_gemm_set_cluster(hos_gemm_t* m, int cluster_size) {
  HOS_REQUIRE(m && (cluster_size == 0 || cluster_size == 2 || cluster_size == 4), "hos_gemm_set_cluster: 0 (automatic), 2 or 4");
  HOS_REQUIRE(cluster_size != 4 || m->max_quads > 0, "hos_gemm_set_cluster: no 4-CTA cluster of this kernel fits on the device");
  m->cluster = cluster_size;
  return HOS_OK;
}

int hos_gemm_set_weight(hos_gemm_t* m, const float* W, const float* b, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(m && W, "hos_gemm_set_weight: bad handle/weights");
  cudaStream_t st = (cudaStream_t)stream;
  const int KB = m->kb0 + m->kb1;
  for (int j = 0; j < m->n / kGemmNB; ++j) {
    const int64_t tot = (int64_t)KB * kGemmNB * kKB;
    // rows [256 j, 256 j + 256) of W [n, k0 + k1]; the first input's columns become chunks 0 .. kb0 - 1
    pack_weight_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(W + (size_t)j * kGemmNB * (m->k0 + m->k1), kGemmNB, m->k0, m->k1,
                                                                     m->x_first, m->kb0, m->kb1, m->ipe_mask,
                                                                     m->d_w + (size_t)j * KB * kGemmNB * 128);
    HOS_LAUNCH_CHECK();
  }
  if (b) HOS_CUDA(cudaMemcpyAsync(m->d_params, b, (size_t)m->n * 4, cudaMemcpyDeviceToDevice, st));
  else HOS_CUDA(cudaMemsetAsync(m->d_params, 0, (size_t)m->n * 4, st));
  return HOS_OK;
}

int hos_gemm_set_head(hos_gemm_t* m, int hn, const float* W, const float* b, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(m && W && hn >= 1 && hn <= 4, "hos_gemm_set_head: 1 <= hn <= 4");
  cudaStream_t st = (cudaStream_t)stream;
  m->hn = hn;
  HOS_CUDA(cudaMemcpyAsync(m->d_params + m->n, W, (size_t)hn * m->n * 4, cudaMemcpyDeviceToDevice, st));
  if (b) HOS_CUDA(cudaMemcpyAsync(m->d_params + (size_t)m->n * (1 + hn), b, (size_t)hn * 4, cudaMemcpyDeviceToDevice, st));
  else HOS_CUDA(cudaMemsetAsync(m->d_params + (size_t)m->n * (1 + hn), 0, 16, st));
  return HOS_OK;
}

int hos_gemm_forward(hos_gemm_t* m, const void* a0_tiled, const void* a1_tiled, int64_t rows, int relu, void* y_tiled,
                     float* head_out, int head_post, float head_shift, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(m && a0_tiled && rows >= 0 && (m->kb1 == 0 || a1_tiled), "hos_gemm_forward: bad handle/inputs");
  HOS_REQUIRE(y_tiled || (head_out && m->hn > 0), "hos_gemm_forward: nothing to write");
  HOS_REQUIRE(!head_out || m->hn > 0, "hos_gemm_forward: call hos_gemm_set_head first");
  if (rows == 0) return HOS_OK;
  GemmArgs a;
  a.a0 = (const unsigned char*)a0_tiled;
  a.a1 = (const unsigned char*)a1_tiled;
  a.w = m->d_w;
  a.params = m->d_params;
  a.y = (unsigned char*)y_tiled;
  a.head_out = head_out;
  a.rows = rows;
  a.ntiles = (int)((rows + kTileM - 1) / kTileM);
  a.kb0 = m->kb0;
  a.kb1 = m->kb1;
  a.n = m->n;
  a.relu = relu;
  a.hn = head_out ? m->hn : 0;
  a.head_post = head_post;
  a.head_shift = head_shift;
  static const int group_env = getenv("HOS_GEMM_GROUP") ? atoi(getenv("HOS_GEMM_GROUP")) : 0;
  // two chunks per issue group: with the 6-stage ring that leaves four stages of prefetch (groups of four left two, and the
  // tensor pipe waited for the refill after every group: 1144 -> 1209 TFLOP/s on a 1024 x 1024 layer at 4.2 M rows)
  a.group = group_env >= 1 && group_env <= 4 ? group_env : 2;
  const int force = m->cluster;                             // A/B switch (hos_gemm_set_cluster): 2 or 4
  const int quads_needed = (a.ntiles + 3) / 4;
  // the quad kernel is opt-in (hos_gemm_set_cluster): measured on a 1024 x 1024 layer at 4.2 M rows it runs at 1029 TFLOP/s
  // against the pair kernel's 1209 - the halved weight traffic buys nothing (the pair kernel is not L2-bound at the clocks the
  // power cap allows) and fewer 4-CTA clusters than CTA pairs are co-resident
  const bool quad = force == 4;
  if (quad && m->max_quads > 0) {
    const int clusters = quads_needed < m->max_quads ? quads_needed : m->max_quads;
    gemm_quad_kernel<<<4 * clusters, kMlpThreads, m->smem_bytes, (cudaStream_t)stream>>>(a);
  } else {
    const int n_groups = (a.ntiles + 1) / 2;
    const int clusters = n_groups < m->max_clusters ? n_groups : m->max_clusters;
    // parameter block in shared memory: bias + the heads actually used
    gemm_pair_kernel<<<2 * clusters, kMlpThreads, m->smem_bytes, (cudaStream_t)stream>>>(a);
  }
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

}  // extern "C"
