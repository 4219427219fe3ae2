// Sampler kernels.  Compiled with -fmad=false: every
// elementwise expression below mirrors the reference's torch op order so that, given
// the same CDF, interval indices and sample positions are bit-identical to CPU torch.
//
// One CTA per ray.  A ray's whole working set (<= 1024 knots) lives in shared memory;
// HBM traffic is the algorithmic minimum (read t/w once, write sdist/tdist once).
#include <math_constants.h>

#include <atomic>

#include "common.cuh"

namespace hos {

constexpr int kSamplerThreads = 256;
constexpr int kMaxKnots = 1024;
constexpr float kEps = 1.1920929e-07f;   // S1 helper.py:18

struct SamplerSmem {
  float knots[kMaxKnots];      // sorted / input knot positions
  float wts[kMaxKnots];        // per-interval weights or softmax probabilities
  float cw[kMaxKnots + 1];     // CDF
  float centers[kMaxKnots];
  float pdf[kMaxKnots / 3 + 1];
  float t0[kMaxKnots / 3 + 1];
  float t1[kMaxKnots / 3 + 1];
  float red[32];
  double tot[32];              // per-chunk totals / carries of the CDF scan
  double carry;
};

__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (wid == 0) {
    r = warp_sum(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  r = red[0];
  __syncthreads();
  return r;
}

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -CUDART_INF_F;
  if (wid == 0) {
    r = warp_max(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  r = red[0];
  __syncthreads();
  return r;
}

// ascending bitonic sort of a[0..n2), n2 a power of two, by the whole CTA
__device__ void bitonic_sort(float* a, int n2) {
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        int p = i ^ j;
        if (p > i) {
          float x = a[i], y = a[p];
          bool up = ((i & k) == 0);
          if ((x > y) == up) { a[i] = y; a[p] = x; }
        }
      }
      __syncthreads();
    }
  }
}

// max_dilate_weights on one ray.  in: t[M+1], w[M] (global).  out (smem): sm.knots[0..3M]
// sorted+clipped, sm.wts[0..3M) renormalised.  S1 helper.py:130-143,152-164.
__device__ void dilate_ray(SamplerSmem& sm, const float* __restrict__ t, const float* __restrict__ w,
                           int M, float dilation, float lo, float hi, bool sorted_input) {
  const int K = 3 * M + 1;
  int n2 = 1;
  while (n2 < K) n2 <<= 1;
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    float a = t[j], b = t[j + 1];
    sm.pdf[j] = w[j] / fmaxf(b - a, kEps);          // weight_to_pdf
    float l = a - dilation, r = b + dilation;
    sm.t0[j] = l;
    sm.t1[j] = r;
    sm.knots[j] = a;
    sm.knots[M + 1 + j] = l;
    sm.knots[2 * M + 1 + j] = r;
  }
  if (threadIdx.x == 0) sm.knots[M] = t[M];
  if (sorted_input) {
    // t ascending => the three lists A = t[0..M], B = t[:-1] - d, C = t[1:] + d are each ascending: the sorted
    // union is a 3-way merge, done by rank: position of an element = its own index + how many elements of the other
    // two lists precede it (ties broken by list order A < B < C, so every position is hit exactly once).  The
    // result is the same ascending array of VALUES that torch.sort produces - one pass instead of 45 bitonic steps.
    __syncthreads();
    for (int e = threadIdx.x; e < K; e += blockDim.x) {
      float x;
      int pos;
      if (e <= M) {                      // from A: count B < x and C < x
        x = sm.knots[e];
        int a = 0, b = M;
        while (a < b) { int m = (a + b) >> 1; if (sm.t0[m] < x) a = m + 1; else b = m; }
        pos = e + a;
        a = 0; b = M;
        while (a < b) { int m = (a + b) >> 1; if (sm.t1[m] < x) a = m + 1; else b = m; }
        pos += a;
      } else if (e <= 2 * M) {           // from B: count A <= x and C < x
        const int j = e - (M + 1);
        x = sm.t0[j];
        int a = 0, b = M + 1;
        while (a < b) { int m = (a + b) >> 1; if (sm.knots[m] <= x) a = m + 1; else b = m; }
        pos = j + a;
        a = 0; b = M;
        while (a < b) { int m = (a + b) >> 1; if (sm.t1[m] < x) a = m + 1; else b = m; }
        pos += a;
      } else {                           // from C: count A <= x and B <= x
        const int j = e - (2 * M + 1);
        x = sm.t1[j];
        int a = 0, b = M + 1;
        while (a < b) { int m = (a + b) >> 1; if (sm.knots[m] <= x) a = m + 1; else b = m; }
        pos = j + a;
        a = 0; b = M;
        while (a < b) { int m = (a + b) >> 1; if (sm.t0[m] <= x) a = m + 1; else b = m; }
        pos += a;
      }
      sm.cw[pos] = fminf(fmaxf(x, lo), hi);          // cw is free until cdf_ray: staging for the merged knots
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x) sm.knots[i] = sm.cw[i];
    __syncthreads();
  } else {
    for (int i = K + threadIdx.x; i < n2; i += blockDim.x) sm.knots[i] = CUDART_INF_F;
    __syncthreads();
    bitonic_sort(sm.knots, n2);
    for (int i = threadIdx.x; i < K; i += blockDim.x) sm.knots[i] = fminf(fmaxf(sm.knots[i], lo), hi);
    __syncthreads();
  }
  float part = 0.f;
  for (int k = threadIdx.x; k < K - 1; k += blockDim.x) {
    float x = sm.knots[k];
    float pm = 0.f;                                  // where(mask, p, 0).max()
    if (sorted_input) {
      // t ascending => t0, t1 ascending => {j : t0[j] <= x < t1[j]} is the contiguous range
      // [#{t1 <= x}, #{t0 <= x}): two binary searches and a walk over the handful of overlapping
      // intervals instead of a scan over all M (same set, same max - still bit-exact).
      int a = 0, b = M;
      while (a < b) { int m = (a + b) >> 1; if (sm.t1[m] <= x) a = m + 1; else b = m; }
      const int lo_j = a;
      a = 0; b = M;
      while (a < b) { int m = (a + b) >> 1; if (sm.t0[m] <= x) a = m + 1; else b = m; }
      for (int j = lo_j; j < a; ++j) pm = fmaxf(pm, sm.pdf[j]);
    } else {
      for (int j = 0; j < M; ++j) {
        bool in = (sm.t0[j] <= x) && (sm.t1[j] > x);
        pm = fmaxf(pm, in ? sm.pdf[j] : 0.f);
      }
    }
    float wd = pm * (sm.knots[k + 1] - x);           // pdf_to_weight
    sm.wts[k] = wd;
    part += wd;
  }
  float tot = block_reduce_sum(part, sm.red);
  float den = fmaxf(tot, kEps);
  for (int k = threadIdx.x; k < K - 1; k += blockDim.x) sm.wts[k] = sm.wts[k] / den;
  __syncthreads();
}

// softmax(logits) -> CDF in sm.cw[0..M].  logits live in sm.wts[0..M).
// S1 helper.py:166-173,193-194.  cumsum accumulates in double like ATen's CPU kernel.
__device__ void cdf_ray(SamplerSmem& sm, int M) {
  float mx = -CUDART_INF_F;
  for (int j = threadIdx.x; j < M; j += blockDim.x) mx = fmaxf(mx, sm.wts[j]);
  mx = block_reduce_max(mx, sm.red);
  float part = 0.f;
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    float e = expf(sm.wts[j] - mx);
    sm.wts[j] = e;
    part += e;
  }
  float tot = block_reduce_sum(part, sm.red);
  for (int j = threadIdx.x; j < M; j += blockDim.x) sm.wts[j] = sm.wts[j] / tot;
  __syncthreads();
  // Inclusive sum in double, 32 elements per warp scan.  Every warp scans its own chunks (all chunks of a ray are in flight at
  // once instead of one warp walking them one after the other: the double-precision scan was the longest serial stretch of
  // the kernel); the running carry is then added in chunk order - carry(c+1) = scan_c[31] + carry(c), inc = scan + carry -
  // which is operation for operation what the single-warp loop computed.
  {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int nchunk = (M - 1 + 31) >> 5;                    // <= 32 (M <= kMaxKnots)
    double* tot = sm.tot;                                    // [nchunk] chunk totals, then carries
    double loc[4];                                           // this warp's chunks c = wid, wid + nw, ... (<= 4 with 8 warps)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = wid + i * nw;
      loc[i] = 0.0;
      if (c < nchunk) {
        const int j = c * 32 + lane;
        const double v = (j < M - 1) ? (double)sm.wts[j] : 0.0;
        loc[i] = warp_incl_sum_d(v, lane);
        if (lane == 31) tot[c] = loc[i];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {                                  // carries in chunk order (exclusive): tot[c] <- carry(c)
      double carry = 0.0;
      for (int c = 0; c < nchunk; ++c) {
        const double t = tot[c];
        tot[c] = carry;
        carry = t + carry;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = wid + i * nw;
      if (c < nchunk) {
        const int j = c * 32 + lane;
        const double inc = loc[i] + tot[c];
        if (j < M - 1) sm.cw[j + 1] = fminf((float)inc, 1.0f);
      }
    }
    if (threadIdx.x == 0) { sm.cw[0] = 0.f; sm.cw[M] = 1.f; }
  }
  __syncthreads();
}

// invert the CDF at S quantiles; knots t in sm.knots[off .. off+M], CDF in sm.cw[0..M].
// S1 helper.py:175-196, 336-359.
__device__ void invert_ray(SamplerSmem& sm, int off, int M, const float* __restrict__ u_base,
                           const float* __restrict__ jitter, int jitter_cols, float max_jitter, int S,
                           float lo, float hi, float* __restrict__ out, float* __restrict__ centers_out,
                           int32_t* __restrict__ idx_out, float* __restrict__ tdist_out = nullptr, float s_near = 0.f,
                           float s_far = 0.f) {
  const float* t = sm.knots + off;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    float u = u_base[s];
    if (jitter) u = u + jitter[jitter_cols == 1 ? 0 : s] * max_jitter;
    // ub = #{j in [0,M] : cw[j] <= u}
    int a = 0, b = M + 1;
    while (a < b) {
      int mid = (a + b) >> 1;
      if (sm.cw[mid] <= u) a = mid + 1; else b = mid;
    }
    int ilo = a - 1, ihi = a;
    float f0 = t[ilo < 0 ? 0 : ilo], x0 = sm.cw[ilo < 0 ? 0 : ilo];
    float f1 = t[ihi > M ? M : ihi], x1 = sm.cw[ihi > M ? M : ihi];
    float r = (u - x0) / (x1 - x0);
    if (isnan(r)) r = 0.f;                           // nan_to_num(., 0)
    r = fminf(fmaxf(r, 0.f), 1.f);
    float c = f0 + r * (f1 - f0);
    sm.centers[s] = c;
    if (centers_out) centers_out[s] = c;
    if (idx_out) idx_out[s] = ilo;
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= S; i += blockDim.x) {
    float v;
    if (i == 0) {
      float mid0 = (sm.centers[1] + sm.centers[0]) / 2.f;
      v = fmaxf(2.f * sm.centers[0] - mid0, lo);
    } else if (i == S) {
      float midl = (sm.centers[S - 1] + sm.centers[S - 2]) / 2.f;
      v = fminf(2.f * sm.centers[S - 1] - midl, hi);
    } else {
      v = (sm.centers[i] + sm.centers[i - 1]) / 2.f;
    }
    out[i] = v;
    if (tdist_out) tdist_out[i] = 1.f / (v * s_far + (1.f - v) * s_near);    // s_to_t (helper.py:146-150)
  }
}

__global__ void __launch_bounds__(kSamplerThreads)
max_dilate_kernel(const float* __restrict__ t, const float* __restrict__ w, int M, float dilation,
                  float lo, float hi, float* __restrict__ t_out, float* __restrict__ w_out) {
  __shared__ SamplerSmem sm;
  const int ray = blockIdx.x;
  dilate_ray(sm, t + (size_t)ray * (M + 1), w + (size_t)ray * M, M, dilation, lo, hi, false);   // arbitrary t
  const int K = 3 * M + 1;
  for (int i = threadIdx.x; i < K; i += blockDim.x) t_out[(size_t)ray * K + i] = sm.knots[i];
  for (int i = threadIdx.x; i < K - 1; i += blockDim.x) w_out[(size_t)ray * (K - 1) + i] = sm.wts[i];
}

__global__ void __launch_bounds__(kSamplerThreads)
sample_intervals_kernel(const float* __restrict__ t, const float* __restrict__ logits,
                        const float* __restrict__ u_base, const float* __restrict__ jitter,
                        int jitter_cols, float max_jitter, int M, int S, float lo, float hi,
                        float* __restrict__ t_out, float* __restrict__ centers_out,
                        int32_t* __restrict__ idx_out) {
  __shared__ SamplerSmem sm;
  const int ray = blockIdx.x;
  for (int j = threadIdx.x; j <= M; j += blockDim.x) sm.knots[j] = t[(size_t)ray * (M + 1) + j];
  for (int j = threadIdx.x; j < M; j += blockDim.x) sm.wts[j] = logits[(size_t)ray * M + j];
  __syncthreads();
  cdf_ray(sm, M);
  invert_ray(sm, 0, M, u_base, jitter ? jitter + (size_t)ray * jitter_cols : nullptr, jitter_cols,
             max_jitter, S, lo, hi, t_out + (size_t)ray * (S + 1),
             centers_out ? centers_out + (size_t)ray * S : nullptr,
             idx_out ? idx_out + (size_t)ray * S : nullptr);
}

// invert_cdf / sorted_interp alone (S1 helper.py:175-196) on a CDF given by the caller: the integer part of the sampler
// (interval index per sample) separated from the floating-point softmax that produces the CDF.
__global__ void __launch_bounds__(kSamplerThreads)
invert_cdf_kernel(const float* __restrict__ t, const float* __restrict__ cw, const float* __restrict__ u_base,
                  const float* __restrict__ jitter, int jitter_cols, float max_jitter, int M, int S, float lo, float hi,
                  float* __restrict__ t_out, float* __restrict__ centers_out, int32_t* __restrict__ idx_out) {
  __shared__ SamplerSmem sm;
  const int ray = blockIdx.x;
  for (int j = threadIdx.x; j <= M; j += blockDim.x) {
    sm.knots[j] = t[(size_t)ray * (M + 1) + j];
    sm.cw[j] = cw[(size_t)ray * (M + 1) + j];
  }
  __syncthreads();
  invert_ray(sm, 0, M, u_base, jitter ? jitter + (size_t)ray * jitter_cols : nullptr, jitter_cols, max_jitter, S, lo, hi,
             t_out + (size_t)ray * (S + 1), centers_out ? centers_out + (size_t)ray * S : nullptr,
             idx_out ? idx_out + (size_t)ray * S : nullptr);
}

// One resampling step of MipNeRF360.forward (S1 model.py:362-408), fused.
__global__ void __launch_bounds__(kSamplerThreads, 8)
resample_level_kernel(const float* __restrict__ sdist, const float* __restrict__ weights, int M_in,
                      int dilate, float dilation, float anneal, float pad,
                      const float* __restrict__ u_base, const float* __restrict__ jitter,
                      int jitter_cols, float max_jitter, int S, float lo, float hi, float s_near,
                      float s_far, float* __restrict__ sdist_out, float* __restrict__ tdist_out) {
  __shared__ SamplerSmem sm;
  const int ray = blockIdx.x;
  const float* t = sdist + (size_t)ray * (M_in + 1);
  const float* w = weights + (size_t)ray * M_in;
  int off, M;
  if (dilate) {
    // sdist comes from the previous level's sample_intervals: ascending by construction; a ray that
    // is not (NaNs, foreign input) falls back to the exhaustive scan
    bool asc = true;
    for (int j = threadIdx.x; j < M_in; j += blockDim.x) asc = asc && (t[j] <= t[j + 1]);
    asc = __syncthreads_and(asc);
    dilate_ray(sm, t, w, M_in, dilation, lo, hi, asc);
    off = 1;                 // sdist[..., 1:-1], weights[..., 1:-1]   (model.py:381-382)
    M = 3 * M_in - 2;
    // logits overwrite the trimmed weights in place: wts[j] <- f(wts[j+1])
    float tmp[(kMaxKnots + kSamplerThreads - 1) / kSamplerThreads];
    int c = 0;
    for (int j = threadIdx.x; j < M; j += blockDim.x) tmp[c++] = sm.wts[j + 1];
    __syncthreads();
    c = 0;
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
      float wj = tmp[c++];
      bool pos = sm.knots[off + j + 1] > sm.knots[off + j];
      sm.wts[j] = pos ? anneal * logf(wj + pad) : -CUDART_INF_F;
    }
  } else {
    off = 0;
    M = M_in;
    for (int j = threadIdx.x; j <= M; j += blockDim.x) sm.knots[j] = t[j];
    __syncthreads();
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
      bool pos = sm.knots[j + 1] > sm.knots[j];
      sm.wts[j] = pos ? anneal * logf(w[j] + pad) : -CUDART_INF_F;
    }
  }
  __syncthreads();
  if (M == 1 && sm.knots[off + 1] > sm.knots[off]) {
    // first level: one interval.  softmax of a single finite logit is exp(0) / exp(0) = 1 whatever the logit, so the CDF is
    // {0, 1}: what cdf_ray would store, without its block-wide max / sum reductions
    if (threadIdx.x == 0) { sm.cw[0] = 0.f; sm.cw[1] = 1.f; }
    __syncthreads();
  } else {
    cdf_ray(sm, M);
  }
  float* so = sdist_out + (size_t)ray * (S + 1);
  invert_ray(sm, off, M, u_base, jitter ? jitter + (size_t)ray * jitter_cols : nullptr, jitter_cols,
             max_jitter, S, lo, hi, so, nullptr, nullptr, tdist_out + (size_t)ray * (S + 1), s_near, s_far);
}

// Human branch samples: z = near*(1-t) + far*t (+ stratified jitter), pts = o + d*z.
// S3 network.py:401-424, 451.
__global__ void human_samples_kernel(const float* __restrict__ ro, const float* __restrict__ rd,
                                     const float* __restrict__ near, const float* __restrict__ far,
                                     const float* __restrict__ t_lin, const float* __restrict__ rand,
                                     int n, int S, float* __restrict__ z_out, float* __restrict__ pts_out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * S) return;
  int r = (int)(i / S), s = (int)(i % S);
  float nr = near[r], fr = far[r];
  auto zf = [&](int k) { float t = t_lin[k]; return nr * (1.f - t) + fr * t; };
  float z = zf(s);
  if (rand) {
    float up = (s == S - 1) ? z : 0.5f * (zf(s + 1) + z);
    float lw = (s == 0) ? z : 0.5f * (z + zf(s - 1));
    z = lw + (up - lw) * rand[i];
  }
  z_out[i] = z;
#pragma unroll
  for (int a = 0; a < 3; ++a) pts_out[i * 3 + a] = ro[r * 3 + a] + rd[r * 3 + a] * z;
}

}  // namespace hos

using namespace hos;

extern "C" {

int hos_max_dilate(const float* t, const float* w, int N, int S, float dilation, float dom_lo,
                   float dom_hi, float* t_out, float* w_out, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(t && w && t_out && w_out, "hos_max_dilate: null pointer");
  HOS_REQUIRE(N >= 0 && S >= 1 && 3 * S + 1 <= kMaxKnots, "hos_max_dilate: need 1 <= S and 3S+1 <= %d (S=%d)", kMaxKnots, S);
  if (N == 0) return HOS_OK;
  max_dilate_kernel<<<N, kSamplerThreads, 0, (cudaStream_t)stream>>>(t, w, S, dilation, dom_lo, dom_hi, t_out, w_out);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_sample_intervals(const float* t, const float* logits, const float* u_base, const float* jitter,
                         int jitter_cols, float max_jitter, int N, int M, int S, float dom_lo,
                         float dom_hi, float* t_out, float* centers_out, int32_t* idx_out, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(t && logits && u_base && t_out, "hos_sample_intervals: null pointer");
  HOS_REQUIRE(N >= 0 && M >= 1 && M + 1 <= kMaxKnots, "hos_sample_intervals: need 1 <= M < %d (M=%d)", kMaxKnots, M);
  HOS_REQUIRE(S >= 2 && S <= kMaxKnots, "hos_sample_intervals: need 2 <= S <= %d (S=%d)", kMaxKnots, S);
  HOS_REQUIRE(!jitter || jitter_cols == 1 || jitter_cols == S, "hos_sample_intervals: jitter_cols must be 1 or S");
  if (N == 0) return HOS_OK;
  sample_intervals_kernel<<<N, kSamplerThreads, 0, (cudaStream_t)stream>>>(
      t, logits, u_base, jitter, jitter_cols, max_jitter, M, S, dom_lo, dom_hi, t_out, centers_out, idx_out);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_invert_cdf(const float* t, const float* cw, const float* u_base, const float* jitter, int jitter_cols, float max_jitter,
                   int N, int M, int S, float dom_lo, float dom_hi, float* t_out, float* centers_out, int32_t* idx_out, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(t && cw && u_base && t_out, "hos_invert_cdf: null pointer");
  HOS_REQUIRE(N >= 0 && M >= 1 && M + 1 <= kMaxKnots && S >= 2 && S <= kMaxKnots, "hos_invert_cdf: bad M / S (M=%d S=%d)", M, S);
  HOS_REQUIRE(!jitter || jitter_cols == 1 || jitter_cols == S, "hos_invert_cdf: jitter_cols must be 1 or S");
  if (N == 0) return HOS_OK;
  invert_cdf_kernel<<<N, kSamplerThreads, 0, (cudaStream_t)stream>>>(t, cw, u_base, jitter, jitter_cols, max_jitter, M, S, dom_lo,
                                                                      dom_hi, t_out, centers_out, idx_out);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_resample_level(const float* sdist, const float* weights, int N, int M_in, int dilate,
                       float dilation, float anneal, float resample_padding, const float* u_base,
                       const float* jitter, int jitter_cols, float max_jitter, int S, float dom_lo,
                       float dom_hi, float s_near, float s_far, float* sdist_out, float* tdist_out,
                       void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(sdist && weights && u_base && sdist_out && tdist_out, "hos_resample_level: null pointer");
  HOS_REQUIRE(N >= 0 && M_in >= 1, "hos_resample_level: bad N/M_in");
  HOS_REQUIRE(dilate ? (3 * M_in + 1 <= kMaxKnots) : (M_in + 1 <= kMaxKnots),
              "hos_resample_level: too many knots (M_in=%d, dilate=%d, max %d)", M_in, dilate, kMaxKnots);
  HOS_REQUIRE(S >= 2 && S <= kMaxKnots, "hos_resample_level: need 2 <= S <= %d", kMaxKnots);
  HOS_REQUIRE(!jitter || jitter_cols == 1 || jitter_cols == S, "hos_resample_level: jitter_cols must be 1 or S");
  if (N == 0) return HOS_OK;
  // 8 CTAs of 20.6 KB static shared memory per SM need the large shared-memory carve-out (a preference, set once)
  // (function attributes are per device: one bit per device ordinal, set with an atomic so concurrent host threads agree)
  static std::atomic<unsigned long long> carveout_set{0ull};
  int dev_ord = 0;
  HOS_CUDA(cudaGetDevice(&dev_ord));
  const unsigned long long bit = 1ull << (dev_ord & 63);
  if (!(carveout_set.load(std::memory_order_relaxed) & bit)) {
    HOS_CUDA(cudaFuncSetAttribute(resample_level_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)cudaSharedmemCarveoutMaxShared));
    carveout_set.fetch_or(bit, std::memory_order_relaxed);
  }
  resample_level_kernel<<<N, kSamplerThreads, 0, (cudaStream_t)stream>>>(
      sdist, weights, M_in, dilate, dilation, anneal, resample_padding, u_base, jitter, jitter_cols,
      max_jitter, S, dom_lo, dom_hi, s_near, s_far, sdist_out, tdist_out);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_human_samples(const float* rays_o, const float* rays_d, const float* near, const float* far,
                      const float* t_lin, const float* rand, int n, int S, float* z_out,
                      float* pts_out, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(rays_o && rays_d && near && far && t_lin && z_out && pts_out, "hos_human_samples: null pointer");
  HOS_REQUIRE(n >= 0 && S >= 1, "hos_human_samples: bad n/S");
  if (n == 0) return HOS_OK;
  size_t tot = (size_t)n * S;
  human_samples_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, near, far, t_lin, rand, n, S, z_out, pts_out);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

}  // extern "C"
