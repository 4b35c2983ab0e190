// hg_common.cuh -- shared host/device helpers for libhistogan_b200.so (sm_100a).
#pragma once

#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/histogan_b200.h"

namespace hg {

// ---------------------------------------------------------------- errors ----
extern thread_local char g_err[512];

inline int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define HG_CUDA_OK(expr)                                                        \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess)                                                      \
      return ::hg::set_error((int)_e, "%s failed: %s (%s:%d)", #expr,           \
                             cudaGetErrorString(_e), __FILE__, __LINE__);       \
  } while (0)

extern unsigned long long g_launches;   // kernels launched by this library (bench evidence)

#define HG_LAUNCH_OK(name)                                                      \
  do {                                                                          \
    __atomic_fetch_add(&::hg::g_launches, 1ULL, __ATOMIC_RELAXED);              \
    cudaError_t _e = cudaGetLastError();                                        \
    if (_e != cudaSuccess)                                                      \
      return ::hg::set_error((int)_e, "launch of %s failed: %s", name,          \
                             cudaGetErrorString(_e));                           \
  } while (0)

struct DeviceInfo {
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  bool ok = false;
};
const DeviceInfo& device_info();   // for the current device (cached)

// "has this one-time, per-DEVICE setup (cudaFuncSetAttribute: function attributes belong to the
// current device's context) been done?" -- one atomic bit per device ordinal.  A process that
// drives several GPUs gets the attribute set on each of them.
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
  bool need() {                      // true: the caller must do the setup now (idempotent if raced)
    int dev = 0;
    cudaGetDevice(&dev);
    bit = 1ull << (dev & 63);
    return !(done.load(std::memory_order_acquire) & bit);
  }
  void mark() { done.fetch_or(bit, std::memory_order_release); }
  static thread_local unsigned long long bit;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------- histogram geometry -
constexpr int kMaxBins = 128;      // generic path limit on h
constexpr float kEps = 1e-6f;      // RGBuvHistBlock.py:25

enum ResizeMode { kNone = 0, kBilinear = 1, kSampling = 2 };

// Everything a kernel needs to know about one RGBuvHistBlock call (by value).
struct HistGeom {
  int B, C, H, W;
  long long sb, sc, sh, sw;
  int h, nc;
  int mode;                 // ResizeMode
  int OH, OW, N;            // pixel grid entering the histogram
  float scale_h, scale_w;   // bilinear: in/out (area_pixel_compute_scale)
  int method, intensity, green_only;
  int projection;           // HG_PROJ_*
  float inv_sigma2;         // float(1/sigma^2)
  double sigma2;            // sigma**2 as the reference computes it (float64)
  double thr_half;          // thresholding: eps/2 (RGBuvHistBlock.py:70-71,126)
};

// Bin centres (np.linspace in float64, RGBuvHistBlock.py:117-118) as float64 and
// as a float32 hi/lo pair, plus the 'sampling' row/col index tables
// (RGBuvHistBlock.py:82-87).  Passed by value in kernel parameter space.
struct HistTables {
  double c[kMaxBins];
  float c_hi[kMaxBins];
  float c_lo[kMaxBins];
  int rows[kMaxBins];
  int cols[kMaxBins];
};

int make_hist_geom(const hg_hist_params* p, HistGeom* g, HistTables* t);

#ifdef __CUDACC__

// float32 natural log, correctly rounded (float64 log rounded once).  The
// reference takes torch.log in float32 (RGBuvHistBlock.py:112-115); a 1-ulp
// difference in a log moves individual bins by ~1e-5 relative (DESIGN.md
// "precision"), so we use the best float32 value available.
__device__ __forceinline__ float log_f32(float v) {
  return __double2float_rn(log((double)v));
}

__device__ __forceinline__ float clamp01(float v) {
  // torch.clamp(x, 0, 1): NaN propagates
  return v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
}

// source index of F.interpolate(mode='bilinear', align_corners=False).  The FMA
// placement reproduces torch's CPU kernel bit-for-bit (verified against
// F.interpolate for 5 geometries, DESIGN.md "precision"): a resized pixel that is
// off by one ulp moves its log by ~1e-7 and the bins it dominates by ~1e-5.
__device__ __forceinline__ void bilinear_taps(float scale, int dst, int in_size,
                                              int& i0, int& i1, float& l0, float& l1) {
  float src = fmaf(scale, __fadd_rn((float)dst, 0.5f), -0.5f);
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = __fadd_rn(src, -(float)i0);
  l0 = __fadd_rn(1.f, -l1);
}

// One pre-processed pixel (clamp -> resize -> first 3 channels),
// RGBuvHistBlock.py:76-99.  p indexes the OH x OW grid row-major.
__device__ __forceinline__ void load_pixel(const float* __restrict__ x, const HistGeom& g,
                                           const HistTables& t, int b, int p,
                                           float& r, float& gg, float& bb) {
  const float* xb = x + (long long)b * g.sb;
  float v[3];
  if (g.mode == kBilinear) {
    const int oy = p / g.OW, ox = p - oy * g.OW;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    bilinear_taps(g.scale_h, oy, g.H, y0, y1, ly0, ly1);
    bilinear_taps(g.scale_w, ox, g.W, x0, x1, lx0, lx1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* xc = xb + c * g.sc;
      const float v00 = clamp01(__ldg(xc + y0 * g.sh + x0 * g.sw));
      const float v01 = clamp01(__ldg(xc + y0 * g.sh + x1 * g.sw));
      const float v10 = clamp01(__ldg(xc + y1 * g.sh + x0 * g.sw));
      const float v11 = clamp01(__ldg(xc + y1 * g.sh + x1 * g.sw));
      const float top = fmaf(v00, lx0, __fmul_rn(v01, lx1));
      const float bot = fmaf(v10, lx0, __fmul_rn(v11, lx1));
      v[c] = fmaf(top, ly0, __fmul_rn(bot, ly1));
    }
  } else {
    int y, xx;
    if (g.mode == kSampling) {
      const int oy = p / g.OW, ox = p - oy * g.OW;
      y = t.rows[oy];
      xx = t.cols[ox];
    } else {
      y = p / g.W;
      xx = p - y * g.W;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      v[c] = clamp01(__ldg(xb + c * g.sc + y * g.sh + xx * g.sw));
  }
  r = v[0]; gg = v[1]; bb = v[2];
}

// Per-pixel projection (RGBuvHistBlock.py:104-115): intensity and the three
// distinct log-chroma coordinates  u_RG = L_R-L_G, u_RB = L_R-L_B, u_GB = L_G-L_B.
// (The six coordinates of the reference are +/- these)
struct PixelProj {
  float iy;            // sqrt(R^2+G^2+B^2+eps)  (1 when !intensity)
  float u_rg, u_rb, u_gb;
};

__device__ __forceinline__ PixelProj project_pixel(float r, float g, float b, bool intensity) {
  PixelProj q;
  const float lr = log_f32(__fadd_rn(r, kEps));
  const float lg = log_f32(__fadd_rn(g, kEps));
  const float lb = log_f32(__fadd_rn(b, kEps));
  q.u_rg = __fadd_rn(lr, -lg);
  q.u_rb = __fadd_rn(lr, -lb);
  q.u_gb = __fadd_rn(lg, -lb);
  if (intensity) {
    const float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r, r), __fmul_rn(g, g)),
                                        __fmul_rn(b, b)), kEps);
    q.iy = __fsqrt_rn(s);
  } else {
    q.iy = 1.f;
  }
  return q;
}

__device__ __forceinline__ float fast_rcp(float v);

// float32 kernel value for the fast path.  d = u - c is formed with the centre
// split into hi+lo float32 parts (the reference subtracts a float64 centre,
// RGBuvHistBlock.py:116-123); everything after that is plain float32.
template <int METHOD>
__device__ __forceinline__ float kernel_f32(float u, float c_hi, float c_lo, float inv_s2) {
  const float d = __fadd_rn(__fadd_rn(u, -c_hi), -c_lo);
  const float q = __fmul_rn(d, d);
  if (METHOD == HG_METHOD_INVERSE_QUADRATIC) {
    return fast_rcp(fmaf(q, inv_s2, 1.f));
  } else {
    return expf(-__fmul_rn(q, inv_s2));
  }
}

// d/du of kernel_f32 given k = kernel value: -2 d / sigma^2 * k^2 (inverse
// quadratic) or -2 d / sigma^2 * k (RBF).
template <int METHOD>
__device__ __forceinline__ float kernel_grad_f32(float u, float c_hi, float c_lo, float inv_s2,
                                                 float k) {
  const float d = __fadd_rn(__fadd_rn(u, -c_hi), -c_lo);
  const float w = -2.f * d * inv_s2 * k;
  return METHOD == HG_METHOD_INVERSE_QUADRATIC ? w * k : w;
}

// float64 kernel exactly as the reference evaluates it (generic path).
__device__ __forceinline__ float kernel_f64(float u, double c, int method, double sigma2,
                                            double thr_half) {
  const double d = fabs((double)u - c);
  if (method == HG_METHOD_THRESHOLDING) return d <= thr_half ? 1.f : 0.f;
  const double q = d * d / sigma2;
  if (method == HG_METHOD_RBF) return (float)exp(-q);
  return (float)(1.0 / (1.0 + q));
}

// 1/x by MUFU.RCP (<= 1 ulp); the extra Newton step of __frcp_rn buys nothing at
// the 1e-5 parity budget (measured: 4e-7 max element-wise on the histogram).
__device__ __forceinline__ float fast_rcp(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#endif  // __CUDACC__

}  // namespace hg
