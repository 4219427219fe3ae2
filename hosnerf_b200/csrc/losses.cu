// Stage-1 regularisers evaluated on the ray history: the proposal ("interlevel")
// loss and the distortion loss of mip-NeRF 360, plus the two small reductions the training-step
// loss value needs.  Forward values only.
//
// One warp per ray, the ray's edges / weights staged in shared memory.  HBM traffic = the
// algorithmic minimum: distortion reads (2S+1) floats and writes 4 B per ray; the interlevel term
// reads (2S+1) + (2Se+1) floats and writes S floats (+4 B).  The O(S^2) pair sum of the distortion
// term stays in registers / shared memory (S <= 1024).
#include "common.cuh"

namespace hos {

constexpr int kLossWarps = 4;   // rays per CTA
constexpr float kLossEps = 1.1920929e-07f;   // S1 helper.py:18

// lossfun_distortion, S1 helper.py:122-128:
//   u = interval mid-points; loss = sum_i w_i sum_j w_j |u_i - u_j| + sum_i w_i^2 (t_{i+1} - t_i) / 3
__global__ void __launch_bounds__(kLossWarps * 32)
loss_distortion_kernel(const float* __restrict__ t, const float* __restrict__ w, int N, int S,
                       float* __restrict__ per_ray) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ray = blockIdx.x * kLossWarps + wid;
  if (ray >= N) return;
  float* su = smem + (size_t)wid * 2 * S;
  float* sw = su + S;
  const float* tr = t + (size_t)ray * (S + 1);
  const float* wr = w + (size_t)ray * S;
  float intra = 0.f;
  for (int i = lane; i < S; i += 32) {
    float a = tr[i], b = tr[i + 1], wi = wr[i];
    su[i] = (b + a) / 2.f;
    sw[i] = wi;
    intra += wi * wi * (b - a);
  }
  __syncwarp();
  float inter = 0.f;
  for (int i = lane; i < S; i += 32) {
    const float ui = su[i];
    float a0 = 0.f, a1 = 0.f;     // two chains: the shared-memory reads are broadcasts, the adds are the limit
    int j = 0;
    for (; j + 1 < S; j += 2) {
      a0 += sw[j] * fabsf(ui - su[j]);
      a1 += sw[j + 1] * fabsf(ui - su[j + 1]);
    }
    if (j < S) a0 += sw[j] * fabsf(ui - su[j]);
    inter += sw[i] * (a0 + a1);
  }
  inter = warp_sum(inter);
  intra = warp_sum(intra);
  if (lane == 0) per_ray[ray] = inter + intra / 3.f;
}

// lossfun_outer, S1 helper.py:92-120: for every interval [t_i, t_{i+1}] of the fine histogram the
// envelope weight that could overlap it is cy[hi(t_{i+1})] - cy[lo(t_i)], with cy the exclusive
// prefix sum of w_env, lo(v) = last envelope edge <= v (edge 0 if none) and hi(v) = first envelope
// edge > v (last edge if none); loss_i = max(w_i - w_outer_i, 0)^2 / (w_i + eps).
__global__ void __launch_bounds__(kLossWarps * 32)
loss_outer_kernel(const float* __restrict__ t, const float* __restrict__ w, const float* __restrict__ t_env,
                  const float* __restrict__ w_env, int N, int S, int Se, float* __restrict__ loss,
                  float* __restrict__ per_ray) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ray = blockIdx.x * kLossWarps + wid;
  if (ray >= N) return;
  // per warp: envelope edges [Se+1] | prefix sums [Se+1] | cy at lo [S+1] | cy at hi [S+1]
  float* ste = smem + (size_t)wid * (2 * (Se + 1) + 2 * (S + 1));
  float* scy = ste + (Se + 1);
  float* slo = scy + (Se + 1);
  float* shi = slo + (S + 1);
  const float* te = t_env + (size_t)ray * (Se + 1);
  const float* we = w_env + (size_t)ray * Se;
  double carry = 0.0;
  for (int base = 0; base < Se + 1; base += 32) {
    int k = base + lane;
    if (k <= Se) ste[k] = te[k];
    // cy[0] = 0, cy[k] = w_env[0] + ... + w_env[k-1]   (ATen's CPU cumsum accumulates float in double)
    double v = (k >= 1 && k <= Se) ? (double)we[k - 1] : 0.0;
    double inc = warp_incl_sum_d(v, lane) + carry;
    carry = __shfl_sync(0xffffffffu, inc, 31);
    if (k <= Se) scy[k] = (float)inc;
  }
  __syncwarp();
  const float* tr = t + (size_t)ray * (S + 1);
  for (int i = lane; i <= S; i += 32) {
    const float v = tr[i];
    int lo = 0, hi = Se + 1;          // cnt = number of envelope edges <= v (edges ascend)
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (ste[mid] <= v) lo = mid + 1; else hi = mid;
    }
    const int cnt = lo;
    slo[i] = scy[cnt > 0 ? cnt - 1 : 0];
    shi[i] = scy[cnt <= Se ? cnt : Se];
  }
  __syncwarp();
  const float* wr = w + (size_t)ray * S;
  float acc = 0.f;
  for (int i = lane; i < S; i += 32) {
    float wi = wr[i];
    float wo = shi[i + 1] - slo[i];
    float d = fmaxf(wi - wo, 0.f);
    float l = d * d / (wi + kLossEps);
    if (loss) loss[(size_t)ray * S + i] = l;
    acc += l;
  }
  if (per_ray) {
    acc = warp_sum(acc);
    if (lane == 0) per_ray[ray] = acc;
  }
}

// Backward of lossfun_distortion w.r.t. the weights (the edges are detached in the reference, S1 model.py:405-406):
//   dL/dw_k = g_ray * (2 sum_j w_j |u_k - u_j| + 2 w_k (t_{k+1} - t_k) / 3),   g_ray = upstream gradient of the ray's term
__global__ void __launch_bounds__(kLossWarps * 32)
loss_distortion_backward_kernel(const float* __restrict__ t, const float* __restrict__ w, const float* __restrict__ g_ray,
                                float g_scalar, int N, int S, float* __restrict__ g_w) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ray = blockIdx.x * kLossWarps + wid;
  if (ray >= N) return;
  float* su = smem + (size_t)wid * 2 * S;
  float* sw = su + S;
  const float* tr = t + (size_t)ray * (S + 1);
  const float* wr = w + (size_t)ray * S;
  for (int i = lane; i < S; i += 32) {
    su[i] = (tr[i + 1] + tr[i]) / 2.f;
    sw[i] = wr[i];
  }
  __syncwarp();
  const float g = g_ray ? g_ray[ray] * g_scalar : g_scalar;
  for (int i = lane; i < S; i += 32) {
    const float ui = su[i];
    float a0 = 0.f, a1 = 0.f;
    int j = 0;
    for (; j + 1 < S; j += 2) {
      a0 += sw[j] * fabsf(ui - su[j]);
      a1 += sw[j + 1] * fabsf(ui - su[j + 1]);
    }
    if (j < S) a0 += sw[j] * fabsf(ui - su[j]);
    g_w[(size_t)ray * S + i] = g * (2.f * (a0 + a1) + 2.f * sw[i] * (tr[i + 1] - tr[i]) / 3.f);
  }
}

// Backward of lossfun_outer w.r.t. the ENVELOPE weights (interlevel_loss detaches the fine histogram, S1 model.py:613-614):
// with q_i = 2 clip(w_i - w_outer_i, 0) / (w_i + eps), w_outer_i = sum_{lo_i <= j < hi_{i+1}} w_env[j]:
//   dL/dw_env[j] = -g * sum_{i : lo_i <= j < hi_{i+1}} q_i.
// lo_i and hi_{i+1} ascend with i, so the fine intervals that cover envelope bin j are the range [b_j, a_j) with
// a_j = #{i : lo_i <= j}, b_j = #{i : hi_{i+1} <= j}: two binary searches and a difference of prefix sums of q -
// deterministic, no atomics.
__global__ void __launch_bounds__(kLossWarps * 32)
loss_outer_backward_kernel(const float* __restrict__ t, const float* __restrict__ w, const float* __restrict__ t_env,
                           const float* __restrict__ w_env, float g_scalar, int N, int S, int Se,
                           float* __restrict__ g_w_env) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ray = blockIdx.x * kLossWarps + wid;
  if (ray >= N) return;
  // per warp: envelope edges [Se+1] | prefix sums cy [Se+1] | lo index [S+1] | hi index [S+1] | prefix sums of q [S+1]
  float* ste = smem + (size_t)wid * (2 * (Se + 1) + 3 * (S + 1));
  float* scy = ste + (Se + 1);
  int* slo = reinterpret_cast<int*>(scy + (Se + 1));
  int* shi = slo + (S + 1);
  float* sq = reinterpret_cast<float*>(shi + (S + 1));
  const float* te = t_env + (size_t)ray * (Se + 1);
  const float* we = w_env + (size_t)ray * Se;
  double carry = 0.0;
  for (int base = 0; base < Se + 1; base += 32) {
    const int k = base + lane;
    if (k <= Se) ste[k] = te[k];
    const double v = (k >= 1 && k <= Se) ? (double)we[k - 1] : 0.0;
    const double inc = warp_incl_sum_d(v, lane) + carry;
    carry = __shfl_sync(0xffffffffu, inc, 31);
    if (k <= Se) scy[k] = (float)inc;
  }
  __syncwarp();
  const float* tr = t + (size_t)ray * (S + 1);
  for (int i = lane; i <= S; i += 32) {
    const float v = tr[i];
    int lo = 0, hi = Se + 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (ste[mid] <= v) lo = mid + 1; else hi = mid;
    }
    slo[i] = lo > 0 ? lo - 1 : 0;
    shi[i] = lo <= Se ? lo : Se;
  }
  __syncwarp();
  const float* wr = w + (size_t)ray * S;
  carry = 0.0;                                       // sq[i] = q_0 + ... + q_{i-1}
  for (int base = 0; base < S + 1; base += 32) {
    const int k = base + lane;
    double q = 0.0;
    if (k >= 1 && k <= S) {
      const int i = k - 1;
      const float wi = wr[i];
      const float wo = scy[shi[i + 1]] - scy[slo[i]];
      q = (double)(2.f * fmaxf(wi - wo, 0.f) / (wi + kLossEps));
    }
    const double inc = warp_incl_sum_d(q, lane) + carry;
    carry = __shfl_sync(0xffffffffu, inc, 31);
    if (k <= S) sq[k] = (float)inc;
  }
  __syncwarp();
  for (int j = lane; j < Se; j += 32) {
    int a = 0, b = S;                                // a_j = #{i < S : lo_i <= j}
    while (a < b) { const int m = (a + b) >> 1; if (slo[m] <= j) a = m + 1; else b = m; }
    const int aj = a;
    a = 0; b = S;                                    // b_j = #{i < S : hi_{i+1} <= j}
    while (a < b) { const int m = (a + b) >> 1; if (shi[m + 1] <= j) a = m + 1; else b = m; }
    const int bj = a;
    g_w_env[(size_t)ray * Se + j] = aj > bj ? -g_scalar * (sq[aj] - sq[bj]) : 0.f;
  }
}

// out[0] = scale * sum(x[0..n)), or scale * sum((x - y)^2) when y is given.  One CTA, fixed order
// (deterministic), double accumulation; the inputs here are per-ray terms (n = rays of one step).
__global__ void __launch_bounds__(1024)
reduce_scaled_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t n, double scale,
                     float* __restrict__ out) {
  __shared__ double part[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    float v = x[i];
    if (y) { v -= y[i]; v *= v; }
    acc += (double)v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) out[0] = (float)(v * scale);
  }
}

}  // namespace hos

using namespace hos;

extern "C" {

int hos_lossfun_distortion(const float* t, const float* w, int N, int S, float* per_ray, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(t && w && per_ray, "hos_lossfun_distortion: null pointer");
  HOS_REQUIRE(N >= 0 && S >= 1 && S <= 1024, "hos_lossfun_distortion: bad shape (1 <= S <= 1024)");
  if (N == 0) return HOS_OK;
  size_t smem = (size_t)kLossWarps * 2 * S * sizeof(float);
  loss_distortion_kernel<<<(N + kLossWarps - 1) / kLossWarps, kLossWarps * 32, smem, (cudaStream_t)stream>>>(
      t, w, N, S, per_ray);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_lossfun_outer(const float* t, const float* w, const float* t_env, const float* w_env, int N, int S,
                      int S_env, float* loss, float* per_ray, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(t && w && t_env && w_env, "hos_lossfun_outer: null pointer");
  HOS_REQUIRE(loss || per_ray, "hos_lossfun_outer: no output requested");
  HOS_REQUIRE(N >= 0 && S >= 1 && S_env >= 1 && S <= 1024 && S_env <= 1024,
              "hos_lossfun_outer: bad shape (1 <= S, S_env <= 1024)");
  if (N == 0) return HOS_OK;
  size_t smem = (size_t)kLossWarps * (2 * (S_env + 1) + 2 * (S + 1)) * sizeof(float);
  if (smem > 48 * 1024)
    HOS_CUDA(cudaFuncSetAttribute(loss_outer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  loss_outer_kernel<<<(N + kLossWarps - 1) / kLossWarps, kLossWarps * 32, smem, (cudaStream_t)stream>>>(
      t, w, t_env, w_env, N, S, S_env, loss, per_ray);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_lossfun_distortion_backward(const float* t, const float* w, const float* g_ray, float g_scalar, int N, int S,
                                    float* g_w, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(t && w && g_w, "hos_lossfun_distortion_backward: null pointer");
  HOS_REQUIRE(N >= 0 && S >= 1 && S <= 1024, "hos_lossfun_distortion_backward: bad shape (1 <= S <= 1024)");
  if (N == 0) return HOS_OK;
  size_t smem = (size_t)kLossWarps * 2 * S * sizeof(float);
  loss_distortion_backward_kernel<<<(N + kLossWarps - 1) / kLossWarps, kLossWarps * 32, smem, (cudaStream_t)stream>>>(
      t, w, g_ray, g_scalar, N, S, g_w);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_lossfun_outer_backward(const float* t, const float* w, const float* t_env, const float* w_env, float g_scalar,
                               int N, int S, int S_env, float* g_w_env, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(t && w && t_env && w_env && g_w_env, "hos_lossfun_outer_backward: null pointer");
  HOS_REQUIRE(N >= 0 && S >= 1 && S_env >= 1 && S <= 1024 && S_env <= 1024,
              "hos_lossfun_outer_backward: bad shape (1 <= S, S_env <= 1024)");
  if (N == 0) return HOS_OK;
  size_t smem = (size_t)kLossWarps * (2 * (S_env + 1) + 3 * (S + 1)) * sizeof(float);
  if (smem > 48 * 1024)
    HOS_CUDA(cudaFuncSetAttribute(loss_outer_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  loss_outer_backward_kernel<<<(N + kLossWarps - 1) / kLossWarps, kLossWarps * 32, smem, (cudaStream_t)stream>>>(
      t, w, t_env, w_env, g_scalar, N, S, S_env, g_w_env);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_reduce_scaled(const float* x, const float* y, int64_t n, double scale, float* out, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(x && out, "hos_reduce_scaled: null pointer");
  HOS_REQUIRE(n >= 0, "hos_reduce_scaled: bad size");
  reduce_scaled_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, y, n, scale, out);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

}  // extern "C"
