// fp32 SIMT MLP building blocks (the "fp32" precision mode: full-precision FFMA
// accumulation, used for the 1e-4 parity gate; the fp16 tcgen05 path lives in mlp_tc.cu).
//
// hos_linear_f32: Y = act([X1 | X2] W^T + b), 128x128x8 tiles, 8x8 register micro-tiles,
// both operands K-contiguous (activations row-major, nn.Linear weights [out,in]).
// The two-segment input avoids ever materialising the skip-connection concat
// (S1 model.py:215-216, mlp_rgb_sigma.py:53, mlp_offset.py:61).
#include "common.cuh"

namespace hos {

constexpr int BM = 128, BN = 128, BK = 8, TM = 8, TN = 8;
constexpr int kLinThreads = 256;

struct Seg {
  const float* p;
  int ld;
  int K;
  int row_div;   // row of this segment = m / row_div (per-ray inputs broadcast over samples)
};

__global__ void __launch_bounds__(kLinThreads)
linear_f32_kernel(Seg s1, Seg s2, const float* __restrict__ W, const float* __restrict__ bias, int64_t M,
                  int N, int act, float* __restrict__ Y, int ldy) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int Ktot = s1.K + s2.K;
  const int tx = tid % 16, ty = tid / 16;      // 16 x 16 threads, each 8x8
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // loader mapping: 128 rows x 8 k = 1024 elements, 4 per thread: row = tid/2, k = (tid%2)*4 .. +3
  const int lrow = tid >> 1;
  const int lk = (tid & 1) * 4;

  for (int seg = 0; seg < 2; ++seg) {
    const Seg s = seg == 0 ? s1 : s2;
    if (s.K == 0) continue;
    const int wofs = seg == 0 ? 0 : s1.K;
    const int64_t am = m0 + lrow;
    const float* arow = (am < M) ? s.p + (am / s.row_div) * (int64_t)s.ld : nullptr;
    const int bn = n0 + lrow;
    const float* brow = (bn < N) ? W + (int64_t)bn * Ktot + wofs : nullptr;
    const bool avec = ((s.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(s.p) & 15) == 0);
    const bool bvec = ((Ktot & 3) == 0) && ((wofs & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
    for (int k0 = 0; k0 < s.K; k0 += BK) {
      float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
      const int k = k0 + lk;
      if (arow) {
        if (avec && k + 3 < s.K) {
          float4 v = *reinterpret_cast<const float4*>(arow + k);
          a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) if (k + q < s.K) a[q] = arow[k + q];
        }
      }
      if (brow) {
        if (bvec && k + 3 < s.K) {
          float4 v = __ldg(reinterpret_cast<const float4*>(brow + k));
          b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) if (k + q < s.K) b[q] = __ldg(brow + k + q);
        }
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) { As[lk + q][lrow] = a[q]; Bs[lk + q][lrow] = b[q]; }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float ra[TM], rb[TN];
        const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * TM + 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN + 4]);
        ra[0] = a0.x; ra[1] = a0.y; ra[2] = a0.z; ra[3] = a0.w; ra[4] = a1.x; ra[5] = a1.y; ra[6] = a1.z; ra[7] = a1.w;
        rb[0] = b0.x; rb[1] = b0.y; rb[2] = b0.z; rb[3] = b0.w; rb[4] = b1.x; rb[5] = b1.y; rb[6] = b1.z; rb[7] = b1.w;
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(ra[i], rb[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (act == 1) v = fmaxf(v, 0.f);
      Y[m * ldy + n] = v;
    }
  }
}

// Skinny products (K1 + K2 <= 32, e.g. the per-ray view-direction term [rays, 27] x [27, 128]): one thread per output element,
// the weight matrix transposed in shared memory (conflict-free: consecutive threads read consecutive columns), inputs of the
// rows of a pass staged next to it.  The 128 x 128-tile kernel above gives such a product 32 CTAs and 15 us; this one a few
// microseconds.  Same arithmetic: fmaf over k in ascending order from zero, then + bias.
constexpr int kSmallRows = 4;           // rows per thread and pass
__global__ void __launch_bounds__(256)
linear_small_f32_kernel(Seg s1, Seg s2, const float* __restrict__ W, const float* __restrict__ bias, int64_t M, int N, int act,
                        float* __restrict__ Y, int ldy) {
  extern __shared__ float s_lin[];
  const int Ktot = s1.K + s2.K;
  float* sW = s_lin;                    // [Ktot][N]
  float* sX = s_lin + Ktot * N;         // [rows per pass][Ktot]
  for (int i = threadIdx.x; i < N * Ktot; i += blockDim.x) {       // consecutive threads: consecutive n of one k (no bank conflicts)
    const int k = i / N, n = i - k * N;
    sW[i] = __ldg(W + (size_t)n * Ktot + k);
  }
  const int rpp = (blockDim.x / N) * kSmallRows;                   // rows per pass (N divides the block size)
  const int n = threadIdx.x % N, rl = (threadIdx.x / N) * kSmallRows;
  const float bn = bias ? bias[n] : 0.f;
  for (int64_t r0 = (int64_t)blockIdx.x * rpp; r0 < M; r0 += (int64_t)gridDim.x * rpp) {
    __syncthreads();
    for (int i = threadIdx.x; i < rpp * Ktot; i += blockDim.x) {
      const int rr = i / Ktot, k = i - rr * Ktot;
      const int64_t m = r0 + rr;
      float v = 0.f;
      if (m < M) v = k < s1.K ? s1.p[(m / s1.row_div) * (int64_t)s1.ld + k] : s2.p[(m / s2.row_div) * (int64_t)s2.ld + (k - s1.K)];
      sX[i] = v;
    }
    __syncthreads();
    float acc[kSmallRows];
#pragma unroll
    for (int q = 0; q < kSmallRows; ++q) acc[q] = 0.f;
    for (int k = 0; k < Ktot; ++k) {
      const float w = sW[k * N + n];
#pragma unroll
      for (int q = 0; q < kSmallRows; ++q) acc[q] = fmaf(sX[(rl + q) * Ktot + k], w, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < kSmallRows; ++q) {
      const int64_t m = r0 + rl + q;
      if (m < M) {
        float v = acc[q] + bn;
        if (act == 1) v = fmaxf(v, 0.f);
        Y[m * ldy + n] = v;
      }
    }
  }
}

// Small-N heads: one warp per row.
__global__ void __launch_bounds__(256)
head_f32_kernel(const float* __restrict__ X, int ldx, int K, const float* __restrict__ W,
                const float* __restrict__ b, int64_t M, int N, int post, float shift,
                const float* __restrict__ add, float* __restrict__ Y, int ldy) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* x = X + row * ldx;
  for (int n = 0; n < N; ++n) {
    const float* w = W + (int64_t)n * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(x[k], __ldg(w + k), acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc + (b ? b[n] : 0.f);
      if (post == 1) {                 // softplus(v + shift), torch threshold 20
        float z = v + shift;
        v = z > 20.f ? z : log1pf(expf(z));
      } else if (post == 2) {          // sigmoid(v) * (1 + 2 pad) - pad
        v = (1.f / (1.f + expf(-v))) * (1.f + 2.f * shift) - shift;
      } else if (post == 3) {
        v = add[row * N + n] + v;
      } else if (post == 4) {          // human branch: sigmoid(rgb), relu(sigma)  (S3 network.py:539-540)
        v = (n < 3) ? 1.f / (1.f + expf(-v)) : fmaxf(v, 0.f);
      }
      Y[row * ldy + n] = v;
    }
  }
}

}  // namespace hos

using namespace hos;

extern "C" {

int hos_linear_f32_ex(const float* X1, int ld1, int K1, const float* X2, int ld2, int K2, int x2_row_div,
                      const float* W, const float* b, int64_t M, int N, int act, float* Y, int ldy,
                      void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(X1 && W && Y, "hos_linear_f32: null pointer");
  HOS_REQUIRE(K1 >= 1 && K2 >= 0 && (K2 == 0 || X2), "hos_linear_f32: bad K1/K2/X2");
  HOS_REQUIRE(ld1 >= K1 && (K2 == 0 || ld2 >= K2) && ldy >= N && N >= 1 && M >= 0, "hos_linear_f32: bad strides");
  HOS_REQUIRE(x2_row_div >= 1, "hos_linear_f32: x2_row_div must be >= 1");
  if (M == 0) return HOS_OK;
  Seg s1{X1, ld1, K1, 1}, s2{X2, ld2, K2, x2_row_div};
  if (K1 + K2 <= 32 && N <= 256 && (256 % N) == 0) {
    const int rpp = (256 / N) * kSmallRows;
    int64_t blocks = (M + rpp - 1) / rpp;
    if (blocks > kNumSMs * 2) blocks = kNumSMs * 2;          // the weight matrix is staged once per CTA: few CTAs, many passes
    const size_t smem = (size_t)(K1 + K2) * (N + rpp) * sizeof(float);
    linear_small_f32_kernel<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(s1, s2, W, b, M, N, act, Y, ldy);
    HOS_LAUNCH_CHECK();
    return HOS_OK;
  }
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
  linear_f32_kernel<<<grid, kLinThreads, 0, (cudaStream_t)stream>>>(s1, s2, W, b, M, N, act, Y, ldy);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_linear_f32(const float* X1, int ld1, int K1, const float* X2, int ld2, int K2, const float* W,
                   const float* b, int64_t M, int N, int act, float* Y, int ldy, void* stream) {
  return hos_linear_f32_ex(X1, ld1, K1, X2, ld2, K2, 1, W, b, M, N, act, Y, ldy, stream);
}

int hos_head_f32(const float* X, int ldx, int K, const float* W, const float* b, int64_t M, int N,
                 int post, float shift, const float* add, float* Y, int ldy, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(X && W && Y, "hos_head_f32: null pointer");
  HOS_REQUIRE(K >= 1 && ldx >= K && N >= 1 && N <= 16 && ldy >= N && M >= 0, "hos_head_f32: bad shape");
  HOS_REQUIRE(post >= 0 && post <= 4 && (post != 3 || add), "hos_head_f32: bad post-op");
  if (M == 0) return HOS_OK;
  const int wpb = 8;
  head_f32_kernel<<<(unsigned)((M + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      X, ldx, K, W, b, M, N, post, shift, add, Y, ldy);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

}  // extern "C"
