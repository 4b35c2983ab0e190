// hist_host.cu -- host-side geometry for the RGB-uv histogram block + misc ABI.
#include "hg_common.cuh"
#include "sm100_ptx.cuh"

#include <math.h>
#include <mutex>

namespace hg {

thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
thread_local unsigned long long PerDeviceOnce::bit = 1;

const DeviceInfo& device_info() {
  static DeviceInfo infos[64];
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  std::lock_guard<std::mutex> lk(mu);
  DeviceInfo& d = infos[dev];
  if (!d.ok) {
    cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
    d.ok = d.sm_count > 0;
  }
  return d;
}

// np.linspace(lo, hi, num) as numpy computes it: step = (hi-lo)/(num-1),
// y[i] = lo + i*step, y[num-1] = hi.
static void linspace(double lo, double hi, int num, double* out) {
  if (num == 1) { out[0] = lo; return; }
  const double step = (hi - lo) / (double)(num - 1);
  for (int i = 0; i < num; ++i) out[i] = lo + (double)i * step;
  out[num - 1] = hi;
}

int make_hist_geom(const hg_hist_params* p, HistGeom* g, HistTables* t) {
  if (!p) return set_error(HG_EINVAL, "null hg_hist_params");
  if (p->B < 0 || p->C < 3 || p->H <= 0 || p->W <= 0)
    return set_error(HG_EINVAL, "bad input shape (B=%d,C=%d,H=%d,W=%d); need C>=3", p->B, p->C,
                     p->H, p->W);
  if (p->h < 1 || p->h > kMaxBins)
    return set_error(HG_ENOSUP, "h=%d outside the supported range [1,%d]", p->h, kMaxBins);
  if (p->method < HG_METHOD_THRESHOLDING || p->method > HG_METHOD_INVERSE_QUADRATIC)
    return set_error(HG_EINVAL, "Wrong kernel method id %d", p->method);
  if (p->resizing != HG_RESIZE_INTERPOLATION && p->resizing != HG_RESIZE_SAMPLING)
    return set_error(HG_EINVAL, "Wrong resizing method id %d", p->resizing);
  if (!(p->lo <= p->hi)) return set_error(HG_EINVAL, "hist_boundary must be sorted");
  if (p->method != HG_METHOD_THRESHOLDING && !(p->sigma > 0.0))
    return set_error(HG_EINVAL, "sigma must be > 0");

  g->B = p->B; g->C = p->C; g->H = p->H; g->W = p->W;
  g->sb = p->sb; g->sc = p->sc; g->sh = p->sh; g->sw = p->sw;
  g->h = p->h;
  if (p->projection < HG_PROJ_RGB_UV || p->projection > HG_PROJ_LAB)
    return set_error(HG_EINVAL, "unknown projection id %d", p->projection);
  g->projection = p->projection;
  g->green_only = (p->green_only && p->projection == HG_PROJ_RGB_UV) ? 1 : 0;
  g->nc = (g->green_only || p->projection != HG_PROJ_RGB_UV) ? 1 : 3;
  g->method = p->method;
  g->intensity = p->intensity_scale ? 1 : 0;
  g->sigma2 = p->sigma * p->sigma;
  g->inv_sigma2 = (float)(1.0 / g->sigma2);
  // RGBuvHistBlock.py:70-71: eps = (|lo|+|hi|)/h ; bins hit when |d| <= eps/2
  g->thr_half = (fabs(p->lo) + fabs(p->hi)) / (double)p->h / 2.0;
  g->scale_h = g->scale_w = 1.f;
  if (p->H > p->insz || p->W > p->insz) {          // RGBuvHistBlock.py:77
    if (p->resizing == HG_RESIZE_INTERPOLATION) {
      if (p->insz < 1) return set_error(HG_EINVAL, "insz must be >= 1");
      g->mode = kBilinear;
      g->OH = g->OW = p->insz;
      g->scale_h = (float)p->H / (float)p->insz;
      g->scale_w = (float)p->W / (float)p->insz;
    } else {
      g->mode = kSampling;
      g->OH = g->OW = p->h;                        // uses self.h, not insz (:82-87)
    }
  } else {
    g->mode = kNone;
    g->OH = p->H; g->OW = p->W;
  }
  const long long n = (long long)g->OH * g->OW;
  if (n > 0x7fffffffLL / 4) return set_error(HG_ENOSUP, "image too large");
  g->N = (int)n;

  if (t) {
    linspace(p->lo, p->hi, p->h, t->c);
    for (int i = 0; i < p->h; ++i) {
      t->c_hi[i] = (float)t->c[i];
      t->c_lo[i] = (float)(t->c[i] - (double)t->c_hi[i]);
    }
    for (int i = p->h; i < kMaxBins; ++i) { t->c[i] = 0; t->c_hi[i] = t->c_lo[i] = 0; }
    // np.linspace(0, H, h, endpoint=False) -> step = H/h ; LongTensor truncates
    const double sh = (double)p->H / (double)p->h, sw = (double)p->W / (double)p->h;
    for (int i = 0; i < kMaxBins; ++i) {
      t->rows[i] = i < p->h ? (int)((double)i * sh) : 0;
      t->cols[i] = i < p->h ? (int)((double)i * sw) : 0;
    }
  }
  return 0;
}

}  // namespace hg

extern "C" {

int hg_abi_version(void) { return HG_ABI_VERSION; }

uint64_t hg_launch_count(void) { return __atomic_load_n(&hg::g_launches, __ATOMIC_RELAXED); }

const char* hg_last_error(void) { return hg::g_err; }

int hg_device_check(int dev) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return hg::set_error(HG_EARCH, "no CUDA device visible (%s)", cudaGetErrorString(e));
  if (dev < 0 || dev >= n) return hg::set_error(HG_EINVAL, "device %d out of range", dev);
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10)
    return hg::set_error(HG_EARCH, "device %d has compute capability %d.x; this library is "
                         "built for sm_100a only", dev, major);
  return 0;
}

int64_t hg_hist_num_pixels(const hg_hist_params* p) {
  hg::HistGeom g;
  int rc = hg::make_hist_geom(p, &g, nullptr);
  if (rc) return rc;
  return g.N;
}

}  // extern "C"

// ----------------------------------------------------- small elementwise ----
namespace hg {

// out[b,h,w,c] = tf32_round(x[b,h,w,c] * (mod ? mod[b,c] : 1)), NHWC
__global__ void __launch_bounds__(256)
modulate_round_kernel(const float4* __restrict__ x, const float* __restrict__ mod,
                      float4* __restrict__ out, long long n4, int C, long long per_image4,
                      int do_round) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = x[i];
  if (mod) {
    const int b = (int)(i / per_image4);
    const int c = (int)((i * 4) % C);
    const float4 m = *reinterpret_cast<const float4*>(mod + (long long)b * C + c);
    v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
  }
  if (do_round) { v.x = tf32_round(v.x); v.y = tf32_round(v.y); v.z = tf32_round(v.z); v.w = tf32_round(v.w); }
  out[i] = v;
}

// out[b,c] += sum_{h,w} a[b,h,w,c] * g[b,h,w,c]   (NHWC); grid (C/32, B, splits)
__global__ void __launch_bounds__(256)
channel_dot_kernel(const float* __restrict__ a, const float* __restrict__ g,
                   float* __restrict__ out, int HW, int C) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int r = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int per = (HW + gridDim.z - 1) / gridDim.z;
  const int p0 = blockIdx.z * per, p1 = min(HW, p0 + per);
  float s = 0.f;
  if (c < C) {
    const float* ab = a + (long long)b * HW * C + c;
    const float* gb = g + (long long)b * HW * C + c;
    for (int p = p0 + r; p < p1; p += 8) s = fmaf(ab[(long long)p * C], gb[(long long)p * C], s);
  }
  red[r][threadIdx.x & 31] = s;
  __syncthreads();
  if (r == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x & 31];
    atomicAdd(out + (long long)b * C + c, t);
  }
}

}  // namespace hg

extern "C" int hg_modulate_round(const float* x, const float* mod, float* out, int32_t B, int32_t HW,
                                 int32_t C, int32_t do_round, hg_stream_t stream_) {
  if (!x || !out) return hg::set_error(HG_EINVAL, "null tensor pointer");
  if (C % 4 != 0) return hg::set_error(HG_ENOSUP, "modulate_round: C=%d must be a multiple of 4", C);
  const long long n4 = (long long)B * HW * C / 4;
  if (n4 <= 0) return 0;
  hg::modulate_round_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      reinterpret_cast<const float4*>(x), mod, reinterpret_cast<float4*>(out), n4, C,
      (long long)HW * C / 4, do_round);
  HG_LAUNCH_OK("modulate_round_kernel");
  return 0;
}

extern "C" int hg_channel_dot(const float* a, const float* g, float* out, int32_t B, int32_t HW,
                              int32_t C, hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a || !g || !out) return hg::set_error(HG_EINVAL, "null tensor pointer");
  if (B <= 0 || HW <= 0 || C <= 0) return 0;
  HG_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * C, stream));
  int splits = (HW + 2047) / 2048;
  if (splits > 64) splits = 64;
  dim3 grid((C + 31) / 32, B, splits);
  hg::channel_dot_kernel<<<grid, 256, 0, stream>>>(a, g, out, HW, C);
  HG_LAUNCH_OK("channel_dot_kernel");
  return 0;
}
