// fused_skeleton.cuh -- the CTA skeleton shared by the memory-bound NHWC kernels
// (gan_fused.cu, recolor.cu): 256 threads = 8 channel lanes (one float4 = 4 channels each
// -> 32 channels) x 32 pixel lanes; per-thread partial sums, a shared-memory reduction
// over the pixel lanes, one atomicAdd per channel.
#pragma once
#include "hg_common.cuh"

namespace hg {

constexpr int kFusedThreads = 256;
constexpr int kPixLanes = 32;

// reduce NQ float4 partials per thread over the 32 pixel lanes; returns the sums in
// the threads with pixel lane 0 (valid for those threads only)
template <int NQ>
__device__ __forceinline__ void reduce_pixel_lanes(float4 (&acc)[NQ], float4* smem /*[NQ][32][8]*/,
                                                   int cl, int pl) {
#pragma unroll
  for (int q = 0; q < NQ; ++q) smem[(q * kPixLanes + pl) * 8 + cl] = acc[q];
  __syncthreads();
  if (pl == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4      // (fully unrolled, ptxas hoisted all 32 x NQ loads and spilled ~1.4 KB)
      for (int k = 0; k < kPixLanes; ++k) {
        const float4 v = smem[(q * kPixLanes + k) * 8 + cl];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      acc[q] = s;
    }
  }
}

__device__ __forceinline__ void atomic_add4(float* p, const float4& v) {
  atomicAdd(p + 0, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w);
}

static inline int pick_pix_per_cta(int HW, int B, int cblocks) {
  // aim for ~4 waves of CTAs over 148 SMs, at least 64 pixels (2 per lane) per CTA
  const long long target = 4LL * 148 * 4;
  long long chunks = (target + (long long)B * cblocks - 1) / ((long long)B * cblocks);
  if (chunks < 1) chunks = 1;
  int per = (int)((HW + chunks - 1) / chunks);
  if (per < 64) per = 64;
  per = (per + kPixLanes - 1) / kPixLanes * kPixLanes;
  return per;
}


}  // namespace hg
