// Shared helpers for libhosnerf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hosnerf_b200.h"

namespace hos {

void set_error(const char* fmt, ...);
int check_arch();   // HOS_OK iff current device is sm_100

#define HOS_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      hos::set_error(__VA_ARGS__);        \
      return HOS_ERR_ARG;                 \
    }                                     \
  } while (0)

#define HOS_CUDA(expr)                                                         \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) {                                                   \
      hos::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                     __FILE__, __LINE__);                                      \
      return HOS_ERR_CUDA;                                                     \
    }                                                                          \
  } while (0)

#define HOS_ARCH_GUARD()                  \
  do {                                    \
    int _a = hos::check_arch();           \
    if (_a != HOS_OK) return _a;          \
  } while (0)

#define HOS_LAUNCH_CHECK() HOS_CUDA(cudaGetLastError())

constexpr int kNumSMs = 148;   // B200

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double shfl_up_d(double v, int d) {
  return __shfl_up_sync(0xffffffffu, v, d);
}
// inclusive warp scans in double (torch CPU cumsum/cumprod accumulate float in double)
__device__ __forceinline__ double warp_incl_sum_d(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double n = shfl_up_d(v, o);
    if (lane >= o) v += n;
  }
  return v;
}
__device__ __forceinline__ double warp_incl_prod_d(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double n = shfl_up_d(v, o);
    if (lane >= o) v *= n;
  }
  return v;
}


// "Tiled fp16" activation layout shared by the encoders and the tcgen05 MLP (see
// include/hosnerf_b200.h): [tile of 128 rows][K-block of 64][128 rows x 128 B, 128B-swizzled].
constexpr int kTileRows = 128;
constexpr int kTileK = 64;
constexpr int kTileChunkBytes = kTileRows * 128;
__host__ __device__ __forceinline__ uint32_t tile_byte_offset(int r, int k) {   // inside one chunk
  return (uint32_t)(r * 128 + ((((k >> 3) ^ (r & 7))) << 4) + ((k & 7) << 1));
}
__device__ __forceinline__ void store_tiled_f16(void* base, int64_t row, int k, int kblocks, float v) {
  int64_t tile = row / kTileRows;
  int r = (int)(row % kTileRows);
  unsigned char* p = reinterpret_cast<unsigned char*>(base) + ((size_t)tile * kblocks + (k >> 6)) * kTileChunkBytes +
                     tile_byte_offset(r, k & 63);
  *reinterpret_cast<__half*>(p) = __float2half_rn(v);
}

}  // namespace hos
