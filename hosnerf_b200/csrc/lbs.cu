// LBS motion field: S3/core/nets/human_nerf/network.py:304-354.
//
// The reference loops 26x over {3x3 matmul, F.grid_sample(32^3 volume)} in Python and
// materialises 26 [P,1] weight tensors and 26 [P,3] positions.  Here one thread owns one
// sample point: bone transforms sit in shared memory, the 3.4 MiB weight volume is read
// through the read-only path and stays L2-resident (gathers from neighbouring samples of
// a ray hit the same cache lines), and only pts (12 B) in / x_skel+mask (16 B) out touch HBM.
// Compiled with -fmad=false: the canonical MLP downstream sees sin(2^9 x), so the warp is kept
// rounding-compatible with the reference's un-fused torch ops (see tests/test_gpu_models.py).
#include "common.cuh"

namespace hos {

constexpr int kMaxBones = 32;
struct LbsParams {
  float bbox_min[3];
  float bbox_scale[3];
};

// Trilinear sample with zeros padding, align_corners=True; term order follows ATen's grid_sampler_3d CPU kernel (tnw, tne,
// tsw, tse, bnw, bne, bsw, bse).  Split in two so the forward warp - all bones sampled at ONE position - sets the cell up once:
// tri_setup() finds the cell and the eight corner weights, tri_sample() gathers one channel.  32-bit index arithmetic, one base
// offset plus constant corner strides, and an interior fast path without per-corner predicates (the common case); the
// arithmetic of each term (three un-fused multiplies and one add, -fmad=false) is unchanged.
struct TriCell {
  int base;          // (z0 * G + y0) * G + x0 - may point outside when a corner is outside; such corners are never read
  unsigned valid;    // bit c: corner c inside the volume; 0xff = interior cell; 0 = sample completely outside
  float w[8];
};

__device__ __forceinline__ void tri_setup(int G, float gx, float gy, float gz, TriCell& c) {
  float s = (float)(G - 1);
  float ix = ((gx + 1.f) / 2.f) * s;
  float iy = ((gy + 1.f) / 2.f) * s;
  float iz = ((gz + 1.f) / 2.f) * s;
  c.valid = 0u;
  // completely outside (or NaN): every corner is out of range
  if (!(ix > -1.f && ix < (float)G && iy > -1.f && iy < (float)G && iz > -1.f && iz < (float)G)) return;
  float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
  int x1 = x0 + 1, y1 = y0 + 1, z1 = z0 + 1;
  float wx1 = ix - fx, wy1 = iy - fy, wz1 = iz - fz;     // weight of the +1 corner
  float wx0 = (float)x1 - ix, wy0 = (float)y1 - iy, wz0 = (float)z1 - iz;
  c.base = (z0 * G + y0) * G + x0;
  c.w[0] = wx0 * wy0 * wz0; c.w[1] = wx1 * wy0 * wz0; c.w[2] = wx0 * wy1 * wz0; c.w[3] = wx1 * wy1 * wz0;
  c.w[4] = wx0 * wy0 * wz1; c.w[5] = wx1 * wy0 * wz1; c.w[6] = wx0 * wy1 * wz1; c.w[7] = wx1 * wy1 * wz1;
  // interior cell (the common case): both corners of every axis inside, i.e. 0 <= x0 <= G - 2 - one unsigned compare per axis
  const unsigned g1 = (unsigned)(G - 1);
  if ((unsigned)x0 < g1 && (unsigned)y0 < g1 && (unsigned)z0 < g1) { c.valid = 0xffu; return; }
  const unsigned bx0 = x0 >= 0 && x0 < G, bx1 = x1 >= 0 && x1 < G;
  const unsigned by0 = y0 >= 0 && y0 < G, by1 = y1 >= 0 && y1 < G;
  const unsigned bz0 = z0 >= 0 && z0 < G, bz1 = z1 >= 0 && z1 < G;
  c.valid = ((bz0 & by0 & bx0) << 0) | ((bz0 & by0 & bx1) << 1) | ((bz0 & by1 & bx0) << 2) | ((bz0 & by1 & bx1) << 3) |
            ((bz1 & by0 & bx0) << 4) | ((bz1 & by0 & bx1) << 5) | ((bz1 & by1 & bx0) << 6) | ((bz1 & by1 & bx1) << 7);
}

__device__ __forceinline__ float tri_sample(const float* __restrict__ v, int G, const TriCell& c) {
  if (c.valid == 0u) return 0.f;
  const float* p = v + c.base;
  const int sy = G, sz = G * G;
  float acc = 0.f;
  if (c.valid == 0xffu) {
    acc += __ldg(p) * c.w[0];
    acc += __ldg(p + 1) * c.w[1];
    acc += __ldg(p + sy) * c.w[2];
    acc += __ldg(p + sy + 1) * c.w[3];
    acc += __ldg(p + sz) * c.w[4];
    acc += __ldg(p + sz + 1) * c.w[5];
    acc += __ldg(p + sz + sy) * c.w[6];
    acc += __ldg(p + sz + sy + 1) * c.w[7];
    return acc;
  }
  if (c.valid & 1u) acc += __ldg(p) * c.w[0];
  if (c.valid & 2u) acc += __ldg(p + 1) * c.w[1];
  if (c.valid & 4u) acc += __ldg(p + sy) * c.w[2];
  if (c.valid & 8u) acc += __ldg(p + sy + 1) * c.w[3];
  if (c.valid & 16u) acc += __ldg(p + sz) * c.w[4];
  if (c.valid & 32u) acc += __ldg(p + sz + 1) * c.w[5];
  if (c.valid & 64u) acc += __ldg(p + sz + sy) * c.w[6];
  if (c.valid & 128u) acc += __ldg(p + sz + sy + 1) * c.w[7];
  return acc;
}

__device__ __forceinline__ float trilinear_zeros(const float* __restrict__ v, int G, float gx, float gy, float gz) {
  TriCell c;
  tri_setup(G, gx, gy, gz, c);
  return tri_sample(v, G, c);
}

__global__ void __launch_bounds__(256)
lbs_warp_kernel(const float* __restrict__ pts, const float* __restrict__ R, const float* __restrict__ T,
                const float* __restrict__ vol, LbsParams prm, int64_t P, int bones, int G,
                float* __restrict__ x_skel, float* __restrict__ mask) {
  __shared__ float4 sRT[kMaxBones * 3];                  // per bone: (R row 0, Tx), (R row 1, Ty), (R row 2, Tz) - three 16-byte loads
  for (int i = threadIdx.x; i < bones * 3; i += blockDim.x)
    sRT[i] = make_float4(R[i * 3 + 0], R[i * 3 + 1], R[i * 3 + 2], T[i]);
  __syncthreads();
  const size_t vstride = (size_t)G * G * G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    float px = pts[i * 3 + 0], py = pts[i * 3 + 1], pz = pts[i * 3 + 2];
    float ax = 0.f, ay = 0.f, az = 0.f, wsum = 0.f;
#pragma unroll 2
    for (int b = 0; b < bones; ++b) {
      const float4 r0 = sRT[b * 3 + 0], r1 = sRT[b * 3 + 1], r2 = sRT[b * 3 + 2];
      // matmul(R_i, pts.T): K=3 GEMM micro-kernel order (fma chain), then the separate "+ T_i"
      float qx = fmaf(r0.z, pz, fmaf(r0.y, py, r0.x * px)) + r0.w;
      float qy = fmaf(r1.z, pz, fmaf(r1.y, py, r1.x * px)) + r1.w;
      float qz = fmaf(r2.z, pz, fmaf(r2.y, py, r2.x * px)) + r2.w;
      float gx = (qx - prm.bbox_min[0]) * prm.bbox_scale[0] - 1.0f;
      float gy = (qy - prm.bbox_min[1]) * prm.bbox_scale[1] - 1.0f;
      float gz = (qz - prm.bbox_min[2]) * prm.bbox_scale[2] - 1.0f;
      float w = trilinear_zeros(vol + b * vstride, G, gx, gy, gz);
      wsum += w;
      ax += w * qx;
      ay += w * qy;
      az += w * qz;
    }
    float den = fmaxf(wsum, 0.0001f);
    x_skel[i * 3 + 0] = ax / den;
    x_skel[i * 3 + 1] = ay / den;
    x_skel[i * 3 + 2] = az / den;
    mask[i] = wsum;
  }
}

// Forward warp of the cycle / flow side paths, S3 network.py:357-398: all bone channels are
// sampled at ONE position (the canonical point), x_deform = sum_i w_i (Rf_i p + Tf_i) / max(sum w, 1e-4).
__global__ void __launch_bounds__(256)
lbs_forward_kernel(const float* __restrict__ pts, const float* __restrict__ R, const float* __restrict__ T,
                   const float* __restrict__ vol, LbsParams prm, int64_t P, int bones, int G,
                   float* __restrict__ x_deform, float* __restrict__ mask) {
  __shared__ float sR[kMaxBones * 9];
  __shared__ float sT[kMaxBones * 3];
  for (int i = threadIdx.x; i < bones * 9; i += blockDim.x) sR[i] = R[i];
  for (int i = threadIdx.x; i < bones * 3; i += blockDim.x) sT[i] = T[i];
  __syncthreads();
  const size_t vstride = (size_t)G * G * G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    float px = pts[i * 3 + 0], py = pts[i * 3 + 1], pz = pts[i * 3 + 2];
    float gx = (px - prm.bbox_min[0]) * prm.bbox_scale[0] - 1.0f;
    float gy = (py - prm.bbox_min[1]) * prm.bbox_scale[1] - 1.0f;
    float gz = (pz - prm.bbox_min[2]) * prm.bbox_scale[2] - 1.0f;
    float ax = 0.f, ay = 0.f, az = 0.f, wsum = 0.f;
    TriCell cell;
    tri_setup(G, gx, gy, gz, cell);                       // every bone channel is sampled at the same position
    for (int b = 0; b < bones; ++b) {
      const float* r = sR + b * 9;
      float w = tri_sample(vol + b * vstride, G, cell);
      float qx = fmaf(r[2], pz, fmaf(r[1], py, r[0] * px)) + sT[b * 3 + 0];
      float qy = fmaf(r[5], pz, fmaf(r[4], py, r[3] * px)) + sT[b * 3 + 1];
      float qz = fmaf(r[8], pz, fmaf(r[7], py, r[6] * px)) + sT[b * 3 + 2];
      wsum += w;
      ax += w * qx;
      ay += w * qy;
      az += w * qz;
    }
    float den = fmaxf(wsum, 0.0001f);
    x_deform[i * 3 + 0] = ax / den;
    x_deform[i * 3 + 1] = ay / den;
    x_deform[i * 3 + 2] = az / den;
    if (mask) mask[i] = wsum;
  }
}

}  // namespace hos

using namespace hos;

extern "C" int hos_lbs_forward(const float* cnl_pts, const float* R_fwd, const float* T_fwd, const float* vol,
                               const float* bbox_min_host, const float* bbox_scale_host, int64_t P,
                               int bones, int G, float* x_deform, float* mask, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(cnl_pts && R_fwd && T_fwd && vol && bbox_min_host && bbox_scale_host && x_deform, "hos_lbs_forward: null pointer");
  HOS_REQUIRE(P >= 0 && bones >= 1 && bones <= kMaxBones && G >= 2, "hos_lbs_forward: bad shape (bones=%d G=%d)", bones, G);
  if (P == 0) return HOS_OK;
  LbsParams prm;
  for (int i = 0; i < 3; ++i) { prm.bbox_min[i] = bbox_min_host[i]; prm.bbox_scale[i] = bbox_scale_host[i]; }
  int64_t blocks = (P + 255) / 256;
  int64_t cap = (int64_t)kNumSMs * 16;
  if (blocks > cap) blocks = cap;
  lbs_forward_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(cnl_pts, R_fwd, T_fwd, vol, prm, P, bones, G, x_deform, mask);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

extern "C" int hos_lbs_warp(const float* pts, const float* R, const float* T, const float* vol,
                            const float* bbox_min_host, const float* bbox_scale_host, int64_t P,
                            int bones, int G, float* x_skel, float* mask, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(pts && R && T && vol && bbox_min_host && bbox_scale_host && x_skel && mask, "hos_lbs_warp: null pointer");
  HOS_REQUIRE(P >= 0 && bones >= 1 && bones <= kMaxBones && G >= 2, "hos_lbs_warp: bad shape (bones=%d G=%d)", bones, G);
  if (P == 0) return HOS_OK;
  LbsParams prm;
  for (int i = 0; i < 3; ++i) { prm.bbox_min[i] = bbox_min_host[i]; prm.bbox_scale[i] = bbox_scale_host[i]; }
  int64_t blocks = (P + 255) / 256;
  int64_t cap = (int64_t)kNumSMs * 16;
  if (blocks > cap) blocks = cap;
  lbs_warp_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pts, R, T, vol, prm, P, bones, G, x_skel, mask);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}
