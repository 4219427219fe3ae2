// Alpha-composite kernels.
//
// One warp per ray; samples are strided across lanes in chunks of 32 and the running
// transmittance is carried between chunks with a warp scan.  The scans accumulate in
// double because ATen's CPU cumsum/cumprod do (acc_type<float> = double) and the oracle
// is CPU torch; everything else is fp32.  HBM traffic = the algorithmic minimum
// (20 B/sample in, 4 B/sample weights out, 12 B/ray rgb out).
#include <math_constants.h>

#include "common.cuh"

namespace hos {

constexpr int kCompWarps = 4;   // rays per CTA

// compute_alpha_weights + volumetric_rendering, S1 helper.py:198-238
__global__ void __launch_bounds__(kCompWarps * 32)
composite_mip360_kernel(const float* __restrict__ density, const float* __restrict__ tdist,
                        const float* __restrict__ dirs, const float* __restrict__ rgb, int N, int S,
                        int opaque, float bg, float* __restrict__ weights, float* __restrict__ rgb_out) {
  const int lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kCompWarps + (threadIdx.x >> 5);
  if (ray >= N) return;
  const float* t = tdist + (size_t)ray * (S + 1);
  const float* sg = density + (size_t)ray * S;
  float dx = dirs[ray * 3], dy = dirs[ray * 3 + 1], dz = dirs[ray * 3 + 2];
  float dn = sqrtf(dx * dx + dy * dy + dz * dz);
  double carry = 0.0;
  float cr = 0.f, cg = 0.f, cb = 0.f, acc = 0.f;
  for (int base = 0; base < S; base += 32) {
    int i = base + lane;
    float dd = 0.f;
    if (i < S) {
      dd = sg[i] * ((t[i + 1] - t[i]) * dn);
      if (opaque && i == S - 1) dd = 1e10f;
    }
    double inc = warp_incl_sum_d((double)dd, lane) + carry;
    float excl = (float)(inc - (double)dd);
    // torch: cumsum(dd[..., :-1]) in double, cast to float, 0 prepended.  (inc - dd) is the
    // same double sum of the preceding elements.
    carry = __shfl_sync(0xffffffffu, inc, 31);
    if (i < S) {
      float alpha = 1.f - expf(-dd);
      float trans = expf(-excl);
      float w = alpha * trans;
      weights[(size_t)ray * S + i] = w;
      acc += w;
      if (rgb) {
        const float* c = rgb + ((size_t)ray * S + i) * 3;
        cr += w * c[0]; cg += w * c[1]; cb += w * c[2];
      }
    }
  }
  if (rgb && rgb_out) {
    acc = warp_sum(acc); cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
    if (lane == 0) {
      float bw = fmaxf(1.f - acc, 0.f) * bg;
      rgb_out[ray * 3 + 0] = cr + bw;
      rgb_out[ray * 3 + 1] = cg + bw;
      rgb_out[ray * 3 + 2] = cb + bw;
    }
  }
}

// Backward of compute_alpha_weights + volumetric_rendering (S1 helper.py:198-238): the first kernel of the training
// path.  Upstream: g_w [N,S] = dL/dweights from the loss terms (optional) and g_out [N,3] = dL/d(composited rgb)
// (optional, final level).  With dd_i = sigma_i * delta_i, T_i = exp(-sum_{j<i} dd_j), w_i = (1 - e^{-dd_i}) T_i:
//   dw_i/ddd_k = -w_i (k < i),  T_k e^{-dd_k} (k = i),  0 (k > i)
//   gw_i     = g_w[i] + g_out . c_i - [1 - sum w >= 0] * bg * sum(g_out)        (the clip of the background weight)
//   g_sigma_k = (gw_k T_k e^{-dd_k} - sum_{i>k} gw_i w_i) * delta_k             (0 for the opaque last sample)
//   g_c_i     = w_i g_out
// One warp per ray; w, T e^{-dd}, delta and gw live in shared memory between the forward and the reverse sweep; the
// scans run in double like the forward kernel.  HBM: 20 B/sample in (+4 B g_w), 4 (+12) B/sample out.
__global__ void __launch_bounds__(kCompWarps * 32)
composite_mip360_backward_kernel(const float* __restrict__ density, const float* __restrict__ tdist,
                                 const float* __restrict__ dirs, const float* __restrict__ rgb,
                                 const float* __restrict__ g_w, const float* __restrict__ g_out, int N, int S, int opaque,
                                 float bg, float* __restrict__ g_density, float* __restrict__ g_rgb) {
  extern __shared__ float bsm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ray = blockIdx.x * kCompWarps + wid;
  if (ray >= N) return;
  float* sw = bsm + (size_t)wid * 4 * S;      // w_i
  float* sa = sw + S;                         // T_i e^{-dd_i}
  float* sd = sa + S;                         // delta_i (0 where sigma gets no gradient)
  float* sg = sd + S;                         // gw_i
  const float* t = tdist + (size_t)ray * (S + 1);
  const float* sgm = density + (size_t)ray * S;
  const float dx = dirs[ray * 3], dy = dirs[ray * 3 + 1], dz = dirs[ray * 3 + 2];
  const float dn = sqrtf(dx * dx + dy * dy + dz * dz);
  float G0 = 0.f, G1 = 0.f, G2 = 0.f;
  if (g_out) { G0 = g_out[ray * 3]; G1 = g_out[ray * 3 + 1]; G2 = g_out[ray * 3 + 2]; }
  double carry = 0.0;
  float acc = 0.f;
  for (int base = 0; base < S; base += 32) {          // forward sweep, same arithmetic as composite_mip360_kernel
    const int i = base + lane;
    float dd = 0.f, delta = 0.f;
    if (i < S) {
      delta = (t[i + 1] - t[i]) * dn;
      dd = sgm[i] * delta;
      if (opaque && i == S - 1) { dd = 1e10f; delta = 0.f; }
    }
    const double inc = warp_incl_sum_d((double)dd, lane) + carry;
    const float excl = (float)(inc - (double)dd);
    carry = __shfl_sync(0xffffffffu, inc, 31);
    if (i < S) {
      const float e = expf(-dd), trans = expf(-excl);
      const float w = (1.f - e) * trans;
      sw[i] = w; sa[i] = trans * e; sd[i] = delta;
      acc += w;
    }
  }
  acc = warp_sum(acc);
  const float base_term = (g_out && (1.f - acc) >= 0.f) ? bg * (G0 + G1 + G2) : 0.f;     // torch.clip passes the gradient at the bound
  __syncwarp();
  for (int i = lane; i < S; i += 32) {
    float gw = g_w ? g_w[(size_t)ray * S + i] : 0.f;
    if (g_out && rgb) {
      const float* c = rgb + ((size_t)ray * S + i) * 3;
      gw += G0 * c[0] + G1 * c[1] + G2 * c[2];
      if (g_rgb) {
        float* o = g_rgb + ((size_t)ray * S + i) * 3;
        const float w = sw[i];
        o[0] = w * G0; o[1] = w * G1; o[2] = w * G2;
      }
    }
    sg[i] = gw - base_term;
  }
  __syncwarp();
  // reverse sweep: suffix_k = sum_{i>k} gw_i w_i, chunks of 32 from the far end
  double tail = 0.0;
  for (int base = ((S - 1) / 32) * 32; base >= 0; base -= 32) {
    const int i = base + (31 - lane);                  // lane 0 holds the farthest sample of the chunk
    const double p = (i < S) ? (double)sg[i] * (double)sw[i] : 0.0;
    const double inc = warp_incl_sum_d(p, lane) + tail;      // inclusive over samples >= i within the sweep
    tail = __shfl_sync(0xffffffffu, inc, 31);
    if (i < S) g_density[(size_t)ray * S + i] = (float)((double)sg[i] * (double)sa[i] - (inc - p)) * sd[i];
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// Network._raw2outputs (S2 network.py:273-299, activate=1) / module-level _raw2outputs
// (S3 model.py:73-99, activate=0)
struct Bg3 { float c[3]; int has; };

__global__ void __launch_bounds__(kCompWarps * 32)
composite_nerf_kernel(const float* __restrict__ raw, const float* __restrict__ mask,
                      const float* __restrict__ z, const float* __restrict__ dirs, Bg3 bgc, int n, int S,
                      int activate, float* __restrict__ rgb_out, float* __restrict__ acc_out,
                      float* __restrict__ weights, float* __restrict__ depth_out) {
  const int lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kCompWarps + (threadIdx.x >> 5);
  if (ray >= n) return;
  const float* zz = z + (size_t)ray * S;
  float dx = dirs[ray * 3], dy = dirs[ray * 3 + 1], dz = dirs[ray * 3 + 2];
  float dn = sqrtf(dx * dx + dy * dy + dz * dz);
  double carry = 1.0;
  float cr = 0.f, cg = 0.f, cb = 0.f, acc = 0.f, dep = 0.f;
  for (int base = 0; base < S; base += 32) {
    int i = base + lane;
    float alpha = 0.f, r = 0.f, g = 0.f, b = 0.f, zi = 0.f;
    if (i < S) {
      const float* q = raw + ((size_t)ray * S + i) * 4;
      zi = zz[i];
      float dist = ((i == S - 1) ? 1e10f : (zz[i + 1] - zi)) * dn;
      float sg = q[3];
      r = q[0]; g = q[1]; b = q[2];
      if (activate) { sg = fmaxf(sg, 0.f); r = sigmoidf_(r); g = sigmoidf_(g); b = sigmoidf_(b); }
      alpha = 1.f - expf(-sg * dist);
      if (mask) alpha = alpha * mask[(size_t)ray * S + i];
    }
    float e = (1.f - alpha) + 1e-10f;
    double inc = warp_incl_prod_d((i < S) ? (double)e : 1.0, lane) * carry;
    double prev = shfl_up_d(inc, 1);
    float T = (float)(lane == 0 ? carry : prev);         // exclusive product
    carry = __shfl_sync(0xffffffffu, inc, 31);
    if (i < S) {
      float w = alpha * T;
      if (weights) weights[(size_t)ray * S + i] = w;
      acc += w; dep += w * zi;
      cr += w * r; cg += w * g; cb += w * b;
    }
  }
  acc = warp_sum(acc); dep = warp_sum(dep);
  cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
  if (lane == 0) {
    if (bgc.has) {
      float rem = 1.f - acc;
      cr += rem * (bgc.c[0] / 255.f); cg += rem * (bgc.c[1] / 255.f); cb += rem * (bgc.c[2] / 255.f);
    }
    if (rgb_out) { rgb_out[ray * 3] = cr; rgb_out[ray * 3 + 1] = cg; rgb_out[ray * 3 + 2] = cb; }
    if (acc_out) acc_out[ray] = acc;
    if (depth_out) depth_out[ray] = dep;
  }
}

// ---------------------------------------------------------------------------
// Stage-3 composite, S3 model.py:1524-1596
// ---------------------------------------------------------------------------
struct Mat4 { float m[16]; };

__global__ void any_small_dir_kernel(const float* __restrict__ d, int n3, int* __restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n3 && fabsf(d[i]) < 1e-5f) *flag = 1;      // model.py:1526
}

constexpr int kS3MaxTot = 1024;

__global__ void __launch_bounds__(kCompWarps * 32)
composite_s3_kernel(const float* __restrict__ bkg_rgb, const float* __restrict__ bkg_density,
                    const float* __restrict__ bkg_tdist, const float* __restrict__ human_rgb,
                    const float* __restrict__ human_density, const float* __restrict__ pts_mask,
                    const float* __restrict__ pts, Mat4 M, const float* __restrict__ rays_o,
                    const float* __restrict__ rays_d, const int* __restrict__ degenerate_flag, int n,
                    int Sb, int Sh, int n2, float thre_fg, float* __restrict__ rgb_out,
                    uint8_t* __restrict__ is_fg, float* __restrict__ human_w) {
  extern __shared__ unsigned char s3_smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* sz = reinterpret_cast<float*>(s3_smem) + (size_t)wid * n2;
  uint16_t* sidx = reinterpret_cast<uint16_t*>(s3_smem + (size_t)kCompWarps * n2 * sizeof(float)) + (size_t)wid * n2;
  const int ray = blockIdx.x * kCompWarps + wid;
  if (ray >= n) return;
  const float* pm = pts_mask + (size_t)ray * Sh;
  float ms = 0.f;
  for (int s = lane; s < Sh; s += 32) ms += pm[s];
  ms = warp_sum(ms);
  const bool fg = ms > thre_fg;                      // model.py:1547-1551
  if (is_fg && lane == 0) is_fg[ray] = fg ? 1 : 0;
  float o[3], d[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) { o[a] = rays_o[ray * 3 + a]; d[a] = rays_d[ray * 3 + a]; }
  const float dn = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  const int tot = fg ? (Sb + Sh) : Sb;
  for (int j = lane; j < Sb; j += 32) { sz[j] = bkg_tdist[(size_t)ray * (Sb + 1) + j]; sidx[j] = (uint16_t)j; }
  if (fg) {
    const int degenerate = *degenerate_flag;
    int axis = -1;
    if (degenerate) {                                // model.py:1527-1538: first axis with |d| > 1e-5
      axis = fabsf(d[0]) > 1e-5f ? 0 : (fabsf(d[1]) > 1e-5f ? 1 : (fabsf(d[2]) > 1e-5f ? 2 : -1));
    }
    for (int s = lane; s < Sh; s += 32) {
      const float* p = pts + ((size_t)ray * Sh + s) * 3;
      float zq[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        float wa = M.m[a * 4 + 0] * p[0] + M.m[a * 4 + 1] * p[1] + M.m[a * 4 + 2] * p[2] + M.m[a * 4 + 3];
        zq[a] = (wa - o[a]) / (d[a] + 1e-10f);
      }
      float zh = (axis >= 0) ? zq[axis] : (zq[0] + zq[1] + zq[2]) / 3.f;
      sz[Sb + s] = zh;
      sidx[Sb + s] = (uint16_t)(Sb + s);
    }
    for (int j = tot + lane; j < n2; j += 32) { sz[j] = CUDART_INF_F; sidx[j] = 0xFFFF; }
    __syncwarp();
    // bitonic sort on (z, idx), ascending
    for (int k = 2; k <= n2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < n2; i += 32) {
          int p2 = i ^ j;
          if (p2 > i) {
            float x = sz[i], y = sz[p2];
            uint16_t xi = sidx[i], yi = sidx[p2];
            bool gt = (x > y) || (x == y && xi > yi) || (isnan(x) && !isnan(y));
            bool up = ((i & k) == 0);
            if (gt == up) { sz[i] = y; sz[p2] = x; sidx[i] = yi; sidx[p2] = xi; }
          }
        }
        __syncwarp();
      }
  }
  __syncwarp();
  if (human_w) for (int s = lane; s < Sh; s += 32) human_w[(size_t)ray * Sh + s] = 0.f;
  __syncwarp();
  double carry = 1.0;
  float cr = 0.f, cg = 0.f, cb = 0.f;
  int hrank = 0;                                     // human samples met so far along the merged (depth-sorted) ray
  for (int base = 0; base < tot; base += 32) {
    int i = base + lane;
    float alpha = 0.f, r = 0.f, g = 0.f, b = 0.f;
    int src = 0;
    if (i < tot) {
      src = sidx[i];
      float zi = sz[i];
      float dist = ((i == tot - 1) ? 1e10f : (sz[i + 1] - zi)) * dn;
      float sg, mk = 1.f;
      if (src < Sb) {
        const float* c = bkg_rgb + ((size_t)ray * Sb + src) * 3;
        r = c[0]; g = c[1]; b = c[2];
        sg = bkg_density[(size_t)ray * Sb + src];
      } else {
        int s = src - Sb;
        const float* c = human_rgb + ((size_t)ray * Sh + s) * 3;
        r = c[0]; g = c[1]; b = c[2];
        sg = human_density[(size_t)ray * Sh + s];
        mk = pm[s];
      }
      alpha = (1.f - expf(-sg * dist)) * mk;
    }
    float e = (1.f - alpha) + 1e-10f;
    double inc = warp_incl_prod_d((i < tot) ? (double)e : 1.0, lane) * carry;
    double prev = shfl_up_d(inc, 1);
    float T = (float)(lane == 0 ? carry : prev);
    carry = __shfl_sync(0xffffffffu, inc, 31);
    // model.py:1577,1588: human_weights = weights[total_order >= Sb].reshape(z_vals_human.shape) - the k-th human sample
    // met in DEPTH order lands in column k (equal to the sample index only while the projected depths are monotone).
    const unsigned hm = __ballot_sync(0xffffffffu, i < tot && src >= Sb);
    if (i < tot) {
      float w = alpha * T;
      cr += w * r; cg += w * g; cb += w * b;
      if (human_w && src >= Sb) human_w[(size_t)ray * Sh + hrank + __popc(hm & ((1u << lane) - 1u))] = w;
    }
    hrank += __popc(hm);
  }
  cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
  if (lane == 0) { rgb_out[ray * 3] = cr; rgb_out[ray * 3 + 1] = cg; rgb_out[ray * 3 + 2] = cb; }
}

}  // namespace hos

using namespace hos;

extern "C" {

int hos_composite_mip360(const float* density, const float* tdist, const float* dirs, const float* rgb,
                         int N, int S, int opaque_background, float bg, float* weights, float* rgb_out,
                         void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(density && tdist && dirs && weights, "hos_composite_mip360: null pointer");
  HOS_REQUIRE(!rgb || rgb_out, "hos_composite_mip360: rgb given but rgb_out is NULL");
  HOS_REQUIRE(N >= 0 && S >= 1, "hos_composite_mip360: bad shape");
  if (N == 0) return HOS_OK;
  composite_mip360_kernel<<<(N + kCompWarps - 1) / kCompWarps, kCompWarps * 32, 0, (cudaStream_t)stream>>>(
      density, tdist, dirs, rgb, N, S, opaque_background, bg, weights, rgb_out);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_composite_mip360_backward(const float* density, const float* tdist, const float* dirs, const float* rgb,
                                  const float* g_weights, const float* g_rgb_out, int N, int S, int opaque_background,
                                  float bg, float* g_density, float* g_rgb, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(density && tdist && dirs && g_density, "hos_composite_mip360_backward: null pointer");
  HOS_REQUIRE(g_weights || g_rgb_out, "hos_composite_mip360_backward: no upstream gradient");
  HOS_REQUIRE(!g_rgb_out || rgb, "hos_composite_mip360_backward: g_rgb_out needs the per-sample rgb");
  HOS_REQUIRE(N >= 0 && S >= 1 && S <= 1024, "hos_composite_mip360_backward: bad shape (1 <= S <= 1024)");
  if (N == 0) return HOS_OK;
  const size_t smem = (size_t)kCompWarps * 4 * S * sizeof(float);
  if (smem > 48 * 1024)
    HOS_CUDA(cudaFuncSetAttribute(composite_mip360_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  composite_mip360_backward_kernel<<<(N + kCompWarps - 1) / kCompWarps, kCompWarps * 32, smem, (cudaStream_t)stream>>>(
      density, tdist, dirs, rgb, g_weights, g_rgb_out, N, S, opaque_background, bg, g_density, g_rgb);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_composite_nerf(const float* raw, const float* mask, const float* z, const float* dirs,
                       const float* bgcolor_host, int n, int S, int activate, float* rgb_out, float* acc,
                       float* weights, float* depth, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(raw && z && dirs, "hos_composite_nerf: null pointer");
  HOS_REQUIRE(n >= 0 && S >= 1, "hos_composite_nerf: bad shape");
  if (n == 0) return HOS_OK;
  Bg3 bgc;
  bgc.has = bgcolor_host != nullptr;
  for (int i = 0; i < 3; ++i) bgc.c[i] = bgcolor_host ? bgcolor_host[i] : 0.f;
  composite_nerf_kernel<<<(n + kCompWarps - 1) / kCompWarps, kCompWarps * 32, 0, (cudaStream_t)stream>>>(
      raw, mask, z, dirs, bgc, n, S, activate, rgb_out, acc, weights, depth);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_composite_s3(const float* bkg_rgb, const float* bkg_density, const float* bkg_tdist,
                     const float* human_rgb, const float* human_density, const float* pts_mask,
                     const float* newsmpl_pts, const float* M_host, const float* rays_o,
                     const float* rays_d, int n, int Sb, int Sh, float thre_fg, float* rgb_out,
                     uint8_t* is_fg, float* human_w, int* flag_ws, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(bkg_rgb && bkg_density && bkg_tdist && human_rgb && human_density && pts_mask && newsmpl_pts &&
              M_host && rays_o && rays_d && rgb_out, "hos_composite_s3: null pointer");
  HOS_REQUIRE(n >= 0 && Sb >= 1 && Sh >= 1 && Sb + Sh <= kS3MaxTot, "hos_composite_s3: need Sb+Sh <= %d", kS3MaxTot);
  if (n == 0) return HOS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  HOS_REQUIRE(flag_ws, "hos_composite_s3: flag_ws (one int of caller-owned device memory) is NULL");
  int* d_flag = flag_ws;                             // "any |rays_d| < 1e-5 in the batch" (model.py:1526), caller-owned:
  HOS_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));   // no allocation in the library, per device, capture-safe
  any_small_dir_kernel<<<(n * 3 + 255) / 256, 256, 0, st>>>(rays_d, n * 3, d_flag);
  HOS_LAUNCH_CHECK();
  int n2 = 1;
  while (n2 < Sb + Sh) n2 <<= 1;
  Mat4 M;
  for (int i = 0; i < 16; ++i) M.m[i] = M_host[i];
  size_t smem = (size_t)kCompWarps * n2 * (sizeof(float) + sizeof(uint16_t));
  composite_s3_kernel<<<(n + kCompWarps - 1) / kCompWarps, kCompWarps * 32, smem, st>>>(
      bkg_rgb, bkg_density, bkg_tdist, human_rgb, human_density, pts_mask, newsmpl_pts, M, rays_o, rays_d,
      d_flag, n, Sb, Sh, n2, thre_fg, rgb_out, is_fg, human_w);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

}  // extern "C"
