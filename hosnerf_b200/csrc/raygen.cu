// On-GPU ray generation and bounding-box clipping: S3 core/utils/camera_util.py:154-265.
// The reference builds 2 M rays per 1080p frame with numpy on the host; here one thread owns one pixel / ray.
// Arithmetic is float64 like numpy's (K, R, T are float64 there), outputs are written as float32 or float64.
// HBM-bound: 0 B in, 40 B/ray out (o, d, viewdir, radius in fp32); the box test reads 24 B and writes 9 B per ray.
#include "common.cuh"

namespace hos {

struct CamParams {
  double kinv[9];      // K^-1, row major
  double r[9];         // R, row major
  double t[3];
  double origin[3];    // -R^T T
};

template <typename TOut>
__global__ void rays_from_krt_kernel(CamParams cam, int H, int W, int want_bkg, TOut* __restrict__ rays_o,
                                     TOut* __restrict__ rays_d, TOut* __restrict__ viewdirs, TOut* __restrict__ radii) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)H * W) return;
  const int j = (int)(idx / W), i = (int)(idx % W);
  auto dir = [&](int row, double (&d)[3]) {
    // pixel_camera = [i, row, 1] . inv(K)^T ; pixel_world = (pixel_camera - T) . R ; d = pixel_world - origin
    const double px = (double)i, py = (double)row;
    double pc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) pc[c] = px * cam.kinv[c * 3 + 0] + py * cam.kinv[c * 3 + 1] + cam.kinv[c * 3 + 2] - cam.t[c];
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c] = pc[0] * cam.r[0 * 3 + c] + pc[1] * cam.r[1 * 3 + c] + pc[2] * cam.r[2 * 3 + c] - cam.origin[c];
  };
  double d[3];
  dir(j, d);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    rays_o[idx * 3 + c] = (TOut)cam.origin[c];
    rays_d[idx * 3 + c] = (TOut)d[c];
  }
  if (!want_bkg) return;
  const double inv = 1.0 / sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
#pragma unroll
  for (int c = 0; c < 3; ++c) viewdirs[idx * 3 + c] = (TOut)(d[c] * inv);
  // radii: distance to the ray of the next image row; the last row repeats dx[-2:-1] = row H - 3 (camera_util.py:211-212)
  const int ja = (j < H - 1) ? j : (H >= 3 ? H - 3 : 0);
  double da[3], db[3];
  dir(ja, da);
  dir(ja + 1, db);
  const double dx = sqrt((da[0] - db[0]) * (da[0] - db[0]) + (da[1] - db[1]) * (da[1] - db[1]) + (da[2] - db[2]) * (da[2] - db[2]));
  radii[idx] = (TOut)(dx * 2.0 / sqrt(12.0));
}

// rays_intersect_3d_bbox: near / far of the rays that cross the (0.01-padded) box exactly twice.
template <typename TIn>
__global__ void rays_bbox_kernel(const TIn* __restrict__ ray_o, TIn* __restrict__ ray_d, int64_t N, double lo0, double lo1,
                                 double lo2, double hi0, double hi1, double hi2, float* __restrict__ near,
                                 float* __restrict__ far, unsigned char* __restrict__ mask) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const double lo[3] = {lo0, lo1, lo2}, hi[3] = {hi0, hi1, hi2};
  TIn o[3], d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o[c] = ray_o[n * 3 + c];
    d[c] = ray_d[n * 3 + c];
    if (fabs((double)d[c]) < 1e-5) {              // ray_d[np.abs(ray_d) < 1e-5] = 1e-5, written back like the reference
      d[c] = (TIn)1e-5;
      ray_d[n * 3 + c] = d[c];
    }
  }
  const double eps = 1e-6;
  int hits = 0;
  double dist[2] = {0.0, 0.0};
#pragma unroll
  for (int k = 0; k < 6; ++k) {                  // plane order of the reference: min x, y, z, max x, y, z
    const int c = k % 3;
    const double bound = k < 3 ? lo[c] : hi[c];
    const double t = (bound - (double)o[c]) / (double)d[c];
    double p[3];
    bool in = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      p[a] = t * (double)d[a] + (double)o[a];
      in = in && p[a] >= lo[a] - eps && p[a] <= hi[a] + eps;
    }
    if (in) {
      if (hits < 2) {
        const double ex = p[0] - (double)o[0], ey = p[1] - (double)o[1], ez = p[2] - (double)o[2];
        dist[hits] = sqrt(ex * ex + ey * ey + ez * ez);
      }
      ++hits;
    }
  }
  const bool ok = hits == 2;
  mask[n] = ok ? 1 : 0;
  float nr = 0.f, fr = 0.f;
  if (ok) {
    const TIn nd = (TIn)sqrt((double)(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]));      // np.linalg.norm in the rays' dtype
    const double d0 = dist[0] / (double)nd, d1 = dist[1] / (double)nd;
    nr = (float)fmin(d0, d1);
    fr = (float)fmax(d0, d1);
  }
  near[n] = nr;
  far[n] = fr;
}

}  // namespace hos

using namespace hos;

extern "C" {

int hos_rays_from_krt(int H, int W, const double* Kinv, const double* R, const double* T, int want_bkg, int out_f64,
                      void* rays_o, void* rays_d, void* viewdirs, void* radii, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(H >= 1 && W >= 1 && Kinv && R && T && rays_o && rays_d, "hos_rays_from_krt: bad arguments");
  HOS_REQUIRE(!want_bkg || (viewdirs && radii && H >= 2), "hos_rays_from_krt: viewdirs / radii outputs need H >= 2");
  CamParams cam;
  for (int i = 0; i < 9; ++i) { cam.kinv[i] = Kinv[i]; cam.r[i] = R[i]; }
  for (int c = 0; c < 3; ++c) {
    cam.t[c] = T[c];
    cam.origin[c] = -(R[0 * 3 + c] * T[0] + R[1 * 3 + c] * T[1] + R[2 * 3 + c] * T[2]);
  }
  const int64_t tot = (int64_t)H * W;
  const unsigned grid = (unsigned)((tot + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_f64)
    rays_from_krt_kernel<double><<<grid, 256, 0, st>>>(cam, H, W, want_bkg, (double*)rays_o, (double*)rays_d, (double*)viewdirs, (double*)radii);
  else
    rays_from_krt_kernel<float><<<grid, 256, 0, st>>>(cam, H, W, want_bkg, (float*)rays_o, (float*)rays_d, (float*)viewdirs, (float*)radii);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

int hos_rays_intersect_bbox(const double* bounds_min, const double* bounds_max, const void* ray_o, void* ray_d, int64_t N,
                            int in_f64, float* near, float* far, unsigned char* mask, void* stream) {
  HOS_ARCH_GUARD();
  HOS_REQUIRE(bounds_min && bounds_max && N >= 0, "hos_rays_intersect_bbox: bad arguments");
  if (N == 0) return HOS_OK;
  HOS_REQUIRE(ray_o && ray_d && near && far && mask, "hos_rays_intersect_bbox: null pointer");
  const unsigned grid = (unsigned)((N + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  const double l0 = bounds_min[0] - 0.01, l1 = bounds_min[1] - 0.01, l2 = bounds_min[2] - 0.01;
  const double h0 = bounds_max[0] + 0.01, h1 = bounds_max[1] + 0.01, h2 = bounds_max[2] + 0.01;
  if (in_f64) rays_bbox_kernel<double><<<grid, 256, 0, st>>>((const double*)ray_o, (double*)ray_d, N, l0, l1, l2, h0, h1, h2, near, far, mask);
  else rays_bbox_kernel<float><<<grid, 256, 0, st>>>((const float*)ray_o, (float*)ray_d, N, l0, l1, l2, h0, h1, h2, near, far, mask);
  HOS_LAUNCH_CHECK();
  return HOS_OK;
}

}  // extern "C"
