// Backward of the LBS motion field (S3/core/nets/human_nerf/network.py:304-354 and the forward warp :357-398) - what the
// reference gets from autograd through 26 x {matmul, F.grid_sample}: gradients with respect to the motion-weight volume
// (scatter of the eight trilinear corner weights, fp32 atomics into the L2-resident [bones, G, G, G] gradient volume), the
// bone rotations / translations (warp-shuffle reduction, then one shared-memory and one global atomic per value) and, for the
// forward warp, the canonical sample position.
//
//   warp:     q_i = R_i p + T_i,  w_i = trilinear(vol_i, u(q_i)),  W = sum w_i,  x = sum w_i q_i / max(W, 1e-4),  mask = W
//   forward:  w_i = trilinear(vol_i, u(c)),  q_i = Rf_i c + Tf_i,  x = sum w_i q_i / max(W, 1e-4)
//   u(q) = (q - bbox_min) * bbox_scale - 1,  voxel index = (u + 1) / 2 * (G - 1)   (align_corners=True, zeros padding)
#include "common.cuh"

namespace hos {

constexpr int kBwdMaxBones = 32;
struct LbsBwdParams {
  float bbox_min[3];
  float bbox_scale[3];
};

// Trilinear sample + its gradient with respect to the voxel index (ix, iy, iz); optionally scatters g * corner weight into gvol.
// Returns the sample value; d[3] = dw/d(ix, iy, iz).  Corners outside the grid contribute nothing (zeros padding).
__device__ __forceinline__ float trilinear_grad(const float* __restrict__ v, float* __restrict__ gvol, int G, float ix, float iy,
                                                float iz, float d[3], float gscatter, bool scatter) {
  d[0] = d[1] = d[2] = 0.f;
  if (!(ix > -1.f && ix < (float)G && iy > -1.f && iy < (float)G && iz > -1.f && iz < (float)G)) return 0.f;
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
  const float wx1 = ix - fx, wy1 = iy - fy, wz1 = iz - fz;
  const float wx0 = 1.f - wx1, wy0 = 1.f - wy1, wz0 = 1.f - wz1;
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int dx = c & 1, dy = (c >> 1) & 1, dz = c >> 2;
    const int x = x0 + dx, y = y0 + dy, z = z0 + dz;
    if (x < 0 || x >= G || y < 0 || y >= G || z < 0 || z >= G) continue;
    const float wx = dx ? wx1 : wx0, wy = dy ? wy1 : wy0, wz = dz ? wz1 : wz0;
    const size_t off = ((size_t)z * G + y) * G + x;
    const float val = __ldg(v + off);
    acc += val * (wx * wy * wz);
    d[0] += val * ((dx ? 1.f : -1.f) * wy * wz);
    d[1] += val * ((dy ? 1.f : -1.f) * wx * wz);
    d[2] += val * ((dz ? 1.f : -1.f) * wx * wy);
    if (scatter) atomicAdd(gvol + off, gscatter * (wx * wy * wz));
  }
  return acc;
}

// FORWARD == false: backward of lbs_warp_kernel (gradient to vol, R, T).
// FORWARD == true:  backward of lbs_forward_kernel (gradient to vol, Rf, Tf and the canonical points).
template <bool FORWARD>
__global__ void __launch_bounds__(256)
lbs_backward_kernel(const float* __restrict__ pts, const float* __restrict__ R, const float* __restrict__ T,
                    const float* __restrict__ vol, LbsBwdParams prm, int64_t P, int bones, int G,
                    const float* __restrict__ g_x, const float* __restrict__ g_mask, float* __restrict__ g_vol,
                    float* __restrict__ g_R, float* __restrict__ g_T, float* __restrict__ g_pts) {
  __shared__ float sR[kBwdMaxBones * 9];
  __shared__ float sT[kBwdMaxBones * 3];
  __shared__ float sG[kBwdMaxBones * 12];             // per bone: dL/dR (9) then dL/dT (3), block partial sums
  for (int i = threadIdx.x; i < bones * 9; i += blockDim.x) sR[i] = R[i];
  for (int i = threadIdx.x; i < bones * 3; i += blockDim.x) sT[i] = T[i];
  for (int i = threadIdx.x; i < bones * 12; i += blockDim.x) sG[i] = 0.f;
  __syncthreads();
  const size_t vstride = (size_t)G * G * G;
  const float sidx = 0.5f * (float)(G - 1);            // d(voxel index) / d(u)
  const int lane = threadIdx.x & 31;
  const int64_t span = (int64_t)gridDim.x * blockDim.x;
  const int64_t iters = (P + span - 1) / span;          // every thread runs the same number of iterations (warp collectives)
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t i = it * span + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = i < P;
    float p[3] = {0.f, 0.f, 0.f}, gx[3] = {0.f, 0.f, 0.f}, gm = 0.f;
    if (ok) {
#pragma unroll
      for (int a = 0; a < 3; ++a) { p[a] = pts[i * 3 + a]; gx[a] = g_x[i * 3 + a]; }
      if (g_mask) gm = g_mask[i];
    }
    // fixed sample position of the forward warp
    float ic[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) ic[a] = (((p[a] - prm.bbox_min[a]) * prm.bbox_scale[a] - 1.f) + 1.f) * sidx;
    // ---- pass 1: W and N = sum w_i q_i
    float W = 0.f, N[3] = {0.f, 0.f, 0.f};
    for (int b = 0; b < bones; ++b) {
      const float* r = sR + b * 9;
      float q[3], d[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) q[a] = fmaf(r[a * 3 + 2], p[2], fmaf(r[a * 3 + 1], p[1], r[a * 3] * p[0])) + sT[b * 3 + a];
      float w;
      if (FORWARD) w = trilinear_grad(vol + b * vstride, nullptr, G, ic[0], ic[1], ic[2], d, 0.f, false);
      else w = trilinear_grad(vol + b * vstride, nullptr, G, ((q[0] - prm.bbox_min[0]) * prm.bbox_scale[0]) * sidx,
                              ((q[1] - prm.bbox_min[1]) * prm.bbox_scale[1]) * sidx, ((q[2] - prm.bbox_min[2]) * prm.bbox_scale[2]) * sidx,
                              d, 0.f, false);
      W += w;
#pragma unroll
      for (int a = 0; a < 3; ++a) N[a] += w * q[a];
    }
    const float D = fmaxf(W, 0.0001f);
    float gN[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) gN[a] = gx[a] / D;
    const float gD = -(gx[0] * N[0] + gx[1] * N[1] + gx[2] * N[2]) / (D * D);
    const float gW = (W > 0.0001f ? gD : 0.f) + gm;
    float gp[3] = {0.f, 0.f, 0.f};
    // ---- pass 2: per bone gradients
    for (int b = 0; b < bones; ++b) {
      const float* r = sR + b * 9;
      float q[3], d[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) q[a] = fmaf(r[a * 3 + 2], p[2], fmaf(r[a * 3 + 1], p[1], r[a * 3] * p[0])) + sT[b * 3 + a];
      const float gw = gN[0] * q[0] + gN[1] * q[1] + gN[2] * q[2] + gW;          // dL/dw_b
      float w;
      if (FORWARD) w = trilinear_grad(vol + b * vstride, g_vol + b * vstride, G, ic[0], ic[1], ic[2], d, gw, ok && g_vol != nullptr);
      else w = trilinear_grad(vol + b * vstride, g_vol + b * vstride, G, ((q[0] - prm.bbox_min[0]) * prm.bbox_scale[0]) * sidx,
                              ((q[1] - prm.bbox_min[1]) * prm.bbox_scale[1]) * sidx, ((q[2] - prm.bbox_min[2]) * prm.bbox_scale[2]) * sidx,
                              d, gw, ok && g_vol != nullptr);
      float gq[3];                                      // dL/dq_b
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        gq[a] = w * gN[a];
        const float gpos = gw * d[a] * (prm.bbox_scale[a] * sidx);              // through the sample position
        if (FORWARD) gp[a] += gpos;
        else gq[a] += gpos;
      }
      if (FORWARD) {                                    // q_b = Rf_b c + Tf_b also depends on the point
#pragma unroll
        for (int c = 0; c < 3; ++c) gp[c] += r[0 * 3 + c] * gq[0] + r[1 * 3 + c] * gq[1] + r[2 * 3 + c] * gq[2];
      }
      // dL/dR_b[a][c] = gq[a] p[c], dL/dT_b[a] = gq[a]: reduce over the warp, one shared atomic per value
      float vals[12];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int c = 0; c < 3; ++c) vals[a * 3 + c] = ok ? gq[a] * p[c] : 0.f;
        vals[9 + a] = ok ? gq[a] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const float s = warp_sum(vals[k]);
        if (lane == 0 && s != 0.f) atomicAdd(&sG[b * 12 + k], s);
      }
    }
    if (FORWARD && ok && g_pts) {
#pragma unroll
      for (int a = 0; a < 3; ++a) g_pts[i * 3 + a] = gp[a];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bones * 12; i += blockDim.x) {
    const int b = i / 12, k = i % 12;
    const float s = sG[i];
    if (s == 0.f) continue;
    if (k < 9) atomicAdd(g_R + b * 9 + k, s);
    else atomicAdd(g_T + b * 3 + (k - 9), s);
  }
}

}  // namespace hos

using namespace hos;

static int lbs_backward_launch(bool forward, const float* pts, const float* R, const float* T, const float* vol,
                               const float* bbox_min_host, const float* bbox_scale_host, int64_t P, int bones, int G,
                               const float* g_x, const float* g_mask, float* g_vol, float* g_R, float* g_T, float* g_pts, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(pts && R && T && vol && bbox_min_host && bbox_scale_host && g_x && g_R && g_T, "hos_lbs_*_backward: null pointer");
  HOS_REQUIRE(P >= 0 && bones >= 1 && bones <= kBwdMaxBones && G >= 2, "hos_lbs_*_backward: bad shape (bones=%d G=%d)", bones, G);
  if (P == 0) return HOS_OK;
  LbsBwdParams prm;
  for (int i = 0; i < 3; ++i) { prm.bbox_min[i] = bbox_min_host[i]; prm.bbox_scale[i] = bbox_scale_host[i]; }
  int64_t blocks = (P + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  if (forward)
    lbs_backward_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pts, R, T, vol, prm, P, bones, G, g_x, g_mask, g_vol, g_R, g_T, g_pts);
  else
    lbs_backward_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pts, R, T, vol, prm, P, bones, G, g_x, g_mask, g_vol, g_R, g_T, g_pts);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

extern "C" int hos_lbs_warp_backward(const float* pts, const float* R, const float* T, const float* vol,
                                     const float* bbox_min_host, const float* bbox_scale_host, int64_t P, int bones, int G,
                                     const float* g_x_skel, const float* g_mask, float* g_vol, float* g_R, float* g_T, void* stream) {
  return lbs_backward_launch(false, pts, R, T, vol, bbox_min_host, bbox_scale_host, P, bones, G, g_x_skel, g_mask, g_vol, g_R, g_T,
                             nullptr, stream);
}

extern "C" int hos_lbs_forward_backward(const float* cnl_pts, const float* R_fwd, const float* T_fwd, const float* vol,
                                        const float* bbox_min_host, const float* bbox_scale_host, int64_t P, int bones, int G,
                                        const float* g_x_deform, float* g_vol, float* g_R, float* g_T, float* g_pts, void* stream) {
  return lbs_backward_launch(true, cnl_pts, R_fwd, T_fwd, vol, bbox_min_host, bbox_scale_host, P, bones, G, g_x_deform, nullptr, g_vol,
                             g_R, g_T, g_pts, stream);
}
