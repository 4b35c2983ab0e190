// recolor.cu -- the memory-bound pieces the ReHistoGAN recolouring step adds to the
// HistoGAN kernels (ReHistoGAN/rehistoGAN.py):
//
//   instnorm_lrelu fwd/bwd   nn.InstanceNorm2d(affine=False, eps=1e-5) + LeakyReLU(0.2) of
//                            EncoderBlock (:489-495), on NHWC activations: one statistics
//                            pass (per (sample, channel) sum / sum of squares, fp64
//                            accumulators) and one apply pass; backward likewise
//   laplacian_l1 fwd/bwd     reconstruction_loss('2nd gradient') (:293-299,321-324):
//                            mean | lap(sum_c a_c) - lap(sum_c b_c) |, deterministic
//                            two-stage reduction; backward wrt b from the stored signs
//   depthwise_conv           gaussian_op (:228-232) = the same KxK filter on every plane,
//                            no padding; with pad = K-1 and the flipped filter it is its
//                            own adjoint (used for the backward)
#include "hg_common.cuh"
#include "sm100_ptx.cuh"
#include "fused_skeleton.cuh"

namespace hg {

// --------------------------------------------------------------- instance norm ----
// sums[b][c] = {sum_p x, sum_p x^2}; grid (C/32, pixel chunks, B)
__global__ void __launch_bounds__(kFusedThreads)
instnorm_stats_kernel(const float* __restrict__ x, double* __restrict__ sums, int HW, int C,
                      int pix_per_cta) {
  __shared__ float4 red[2 * kPixLanes * 8];
  const int cl = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = blockIdx.x * 32 + cl * 4;
  const int b = blockIdx.z;
  const int p0 = blockIdx.y * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  const bool cvalid = c < C;
  float4 acc[2];
  acc[0] = acc[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cvalid) {
    for (int p = p0 + pl; p < p1; p += kPixLanes) {
      const float4 v = *reinterpret_cast<const float4*>(x + ((long long)b * HW + p) * C + c);
      acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
      acc[1].x = fmaf(v.x, v.x, acc[1].x); acc[1].y = fmaf(v.y, v.y, acc[1].y);
      acc[1].z = fmaf(v.z, v.z, acc[1].z); acc[1].w = fmaf(v.w, v.w, acc[1].w);
    }
  }
  reduce_pixel_lanes<2>(acc, red, cl, pl);
  if (pl == 0 && cvalid) {
    double* s = sums + ((long long)b * C + c) * 2;
    atomicAdd(s + 0, (double)acc[0].x); atomicAdd(s + 1, (double)acc[1].x);
    atomicAdd(s + 2, (double)acc[0].y); atomicAdd(s + 3, (double)acc[1].y);
    atomicAdd(s + 4, (double)acc[0].z); atomicAdd(s + 5, (double)acc[1].z);
    atomicAdd(s + 6, (double)acc[0].w); atomicAdd(s + 7, (double)acc[1].w);
  }
}

__device__ __forceinline__ void mean_rstd(const double* s, int HW, float eps, float& mean, float& rstd) {
  const double m = s[0] / HW;
  double var = s[1] / HW - m * m;             // biased variance, as F.instance_norm
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(var + (double)eps));
}

// y = lrelu((x - mean) * rstd)   [optionally TF32-rounded for the next convolution]
__global__ void __launch_bounds__(kFusedThreads)
instnorm_lrelu_apply_kernel(const float* __restrict__ x, const double* __restrict__ sums,
                            float* __restrict__ y, int HW, int C, float eps, float slope,
                            int round, int pix_per_cta) {
  const int cl = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = blockIdx.x * 32 + cl * 4;
  const int b = blockIdx.z;
  const int p0 = blockIdx.y * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  if (c >= C) return;
  float mu[4], rs[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) mean_rstd(sums + ((long long)b * C + c + k) * 2, HW, eps, mu[k], rs[k]);
  for (int p = p0 + pl; p < p1; p += kPixLanes) {
    const long long off = ((long long)b * HW + p) * C + c;
    const float4 v = *reinterpret_cast<const float4*>(x + off);
    float4 o;
#define HG_IN(F, K)                                               \
    {                                                             \
      const float xh = (v.F - mu[K]) * rs[K];                     \
      const float a = xh > 0.f ? xh : xh * slope;                 \
      o.F = round ? tf32_round(a) : a;                            \
    }
    HG_IN(x, 0) HG_IN(y, 1) HG_IN(z, 2) HG_IN(w, 3)
#undef HG_IN
    *reinterpret_cast<float4*>(y + off) = o;
  }
}

// backward statistics: bs[b][c] = {sum_p g, sum_p g * xhat},  g = dy * lrelu'(xhat)
__global__ void __launch_bounds__(kFusedThreads)
instnorm_lrelu_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                const double* __restrict__ sums, double* __restrict__ bs, int HW,
                                int C, float eps, float slope, int pix_per_cta) {
  __shared__ float4 red[2 * kPixLanes * 8];
  const int cl = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = blockIdx.x * 32 + cl * 4;
  const int b = blockIdx.z;
  const int p0 = blockIdx.y * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  const bool cvalid = c < C;
  float4 acc[2];
  acc[0] = acc[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cvalid) {
    float mu[4], rs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) mean_rstd(sums + ((long long)b * C + c + k) * 2, HW, eps, mu[k], rs[k]);
    for (int p = p0 + pl; p < p1; p += kPixLanes) {
      const long long off = ((long long)b * HW + p) * C + c;
      const float4 v = *reinterpret_cast<const float4*>(x + off);
      const float4 g = *reinterpret_cast<const float4*>(dy + off);
#define HG_INB(F, K)                                              \
      {                                                           \
        const float xh = (v.F - mu[K]) * rs[K];                   \
        const float gg = xh > 0.f ? g.F : g.F * slope;            \
        acc[0].F += gg;                                           \
        acc[1].F = fmaf(gg, xh, acc[1].F);                        \
      }
      HG_INB(x, 0) HG_INB(y, 1) HG_INB(z, 2) HG_INB(w, 3)
#undef HG_INB
    }
  }
  reduce_pixel_lanes<2>(acc, red, cl, pl);
  if (pl == 0 && cvalid) {
    double* s = bs + ((long long)b * C + c) * 2;
    atomicAdd(s + 0, (double)acc[0].x); atomicAdd(s + 1, (double)acc[1].x);
    atomicAdd(s + 2, (double)acc[0].y); atomicAdd(s + 3, (double)acc[1].y);
    atomicAdd(s + 4, (double)acc[0].z); atomicAdd(s + 5, (double)acc[1].z);
    atomicAdd(s + 6, (double)acc[0].w); atomicAdd(s + 7, (double)acc[1].w);
  }
}

// dx = rstd * (g - mean_p(g) - xhat * mean_p(g * xhat))
__global__ void __launch_bounds__(kFusedThreads)
instnorm_lrelu_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                const double* __restrict__ sums, const double* __restrict__ bs,
                                float* __restrict__ dx, int HW, int C, float eps, float slope,
                                int round, int pix_per_cta) {
  const int cl = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = blockIdx.x * 32 + cl * 4;
  const int b = blockIdx.z;
  const int p0 = blockIdx.y * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  if (c >= C) return;
  float mu[4], rs[4], m1[4], m2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    mean_rstd(sums + ((long long)b * C + c + k) * 2, HW, eps, mu[k], rs[k]);
    const double* s = bs + ((long long)b * C + c + k) * 2;
    m1[k] = (float)(s[0] / HW);
    m2[k] = (float)(s[1] / HW);
  }
  for (int p = p0 + pl; p < p1; p += kPixLanes) {
    const long long off = ((long long)b * HW + p) * C + c;
    const float4 v = *reinterpret_cast<const float4*>(x + off);
    const float4 g = *reinterpret_cast<const float4*>(dy + off);
    float4 o;
#define HG_INA(F, K)                                              \
    {                                                             \
      const float xh = (v.F - mu[K]) * rs[K];                     \
      const float gg = xh > 0.f ? g.F : g.F * slope;              \
      const float d = rs[K] * (gg - m1[K] - xh * m2[K]);          \
      o.F = round ? tf32_round(d) : d;                            \
    }
    HG_INA(x, 0) HG_INA(y, 1) HG_INA(z, 2) HG_INA(w, 3)
#undef HG_INA
    *reinterpret_cast<float4*>(dx + off) = o;
  }
}

// ------------------------------------------------------------ laplacian L1 loss ----
constexpr int kLapThreads = 256;
constexpr int kLapBlocks = 592;            // 4 per SM; fixed so the reduction is reproducible

__device__ __forceinline__ float lap_sum3(const float* __restrict__ img, int H, int W, int y, int x) {
  // sum over the 3 colour planes of the 5-point laplacian with zero padding (F.conv2d padding=1)
  float s = 0.f;
  const long long plane = (long long)H * W;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* q = img + c * plane;
    float v = -4.f * q[(long long)y * W + x];
    if (y > 0) v += q[(long long)(y - 1) * W + x];
    if (y + 1 < H) v += q[(long long)(y + 1) * W + x];
    if (x > 0) v += q[(long long)y * W + x - 1];
    if (x + 1 < W) v += q[(long long)y * W + x + 1];
    s += v;
  }
  return s;
}

__global__ void __launch_bounds__(kLapThreads)
laplacian_l1_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                            signed char* __restrict__ sign, double* __restrict__ partial, int B,
                            int H, int W) {
  __shared__ double red[kLapThreads];
  const long long n = (long long)B * H * W;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * kLapThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kLapThreads) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const long long img = (i / ((long long)H * W)) * 3 * H * W;
    const float d = lap_sum3(a + img, H, W, y, x) - lap_sum3(b + img, H, W, y, x);
    sign[i] = d > 0.f ? 1 : (d < 0.f ? -1 : 0);
    acc += (double)fabsf(d);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kLapThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void laplacian_l1_final_kernel(const double* __restrict__ partial, int nblocks,
                                          float* __restrict__ loss, double inv_n) {
  __shared__ double red[kLapThreads];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += kLapThreads) acc += partial[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kLapThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = (float)(red[0] * inv_n);
}

// d loss / d b[bi,c,y,x] = -(gout / N) * lap(sign)[bi,y,x]   (the stencil is symmetric)
__global__ void __launch_bounds__(256)
laplacian_l1_bwd_kernel(const signed char* __restrict__ sign, const float* __restrict__ gout,
                        float* __restrict__ db, int B, int H, int W, float inv_n) {
  const long long n = (long long)B * H * W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % W);
  const int y = (int)((i / W) % H);
  const long long bi = i / ((long long)H * W);
  const signed char* s = sign + bi * H * W;
  int v = -4 * s[(long long)y * W + x];
  if (y > 0) v += s[(long long)(y - 1) * W + x];
  if (y + 1 < H) v += s[(long long)(y + 1) * W + x];
  if (x > 0) v += s[(long long)y * W + x - 1];
  if (x + 1 < W) v += s[(long long)y * W + x + 1];
  const float g = -(*gout) * inv_n * (float)v;
  float* o = db + bi * 3 * H * W + (long long)y * W + x;
  o[0] = g; o[(long long)H * W] = g; o[2LL * H * W] = g;
}

// ------------------------------------------------------- depthwise KxK filter ----
constexpr int kDwMaxK = 15;
constexpr int kDwTW = 32, kDwTH = 16;

// y[pl, oy, ox] = sum_{i,j} k[i][j] * x[pl, oy + i - pad, ox + j - pad]  (zero outside x);
// flip: use k[K-1-i][K-1-j]
__global__ void __launch_bounds__(kDwTW * kDwTH)
depthwise_conv_kernel(const float* __restrict__ x, const float* __restrict__ kern,
                      float* __restrict__ y, int H, int W, int OH, int OW, int K, int pad, int flip) {
  __shared__ float tile[(kDwTH + kDwMaxK - 1) * (kDwTW + kDwMaxK - 1)];
  __shared__ float kw[kDwMaxK * kDwMaxK];
  const int tid = threadIdx.y * kDwTW + threadIdx.x;
  const int plane = blockIdx.z;
  const int ox0 = blockIdx.x * kDwTW, oy0 = blockIdx.y * kDwTH;
  const int tw = kDwTW + K - 1, th = kDwTH + K - 1;
  for (int i = tid; i < K * K; i += kDwTW * kDwTH) kw[i] = kern[flip ? K * K - 1 - i : i];
  const float* xp = x + (long long)plane * H * W;
  for (int i = tid; i < tw * th; i += kDwTW * kDwTH) {
    const int ty = i / tw, tx = i - ty * tw;
    const int iy = oy0 + ty - pad, ix = ox0 + tx - pad;
    tile[i] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? xp[(long long)iy * W + ix] : 0.f;
  }
  __syncthreads();
  const int ox = ox0 + threadIdx.x, oy = oy0 + threadIdx.y;
  if (ox >= OW || oy >= OH) return;
  float acc = 0.f;
  for (int i = 0; i < K; ++i) {
    const float* row = tile + (threadIdx.y + i) * tw + threadIdx.x;
    for (int j = 0; j < K; ++j) acc = fmaf(kw[i * K + j], row[j], acc);
  }
  y[((long long)plane * OH + oy) * OW + ox] = acc;
}

}  // namespace hg

using namespace hg;

static int in_args_ok(const void* a, const void* b, const void* c, int C) {
  if (!a || !b || !c) return set_error(HG_EINVAL, "null tensor pointer");
  if (C % 4) return set_error(HG_ENOSUP, "C=%d must be a multiple of 4", C);
  return 0;
}

extern "C" int hg_instnorm_lrelu_fwd(const float* x, float* y, double* stats, int32_t B, int32_t HW,
                                     int32_t C, float eps, float slope, int32_t round_tf32,
                                     hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = in_args_ok(x, y, stats, C)) return rc;
  if (B <= 0 || HW <= 0) return 0;
  HG_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * (size_t)B * C, stream));
  const int cblocks = (C + 31) / 32;
  const int per = pick_pix_per_cta(HW, B, cblocks);
  dim3 grid(cblocks, (HW + per - 1) / per, B);
  instnorm_stats_kernel<<<grid, kFusedThreads, 0, stream>>>(x, stats, HW, C, per);
  HG_LAUNCH_OK("instnorm_stats_kernel");
  instnorm_lrelu_apply_kernel<<<grid, kFusedThreads, 0, stream>>>(x, stats, y, HW, C, eps, slope,
                                                                  round_tf32, per);
  HG_LAUNCH_OK("instnorm_lrelu_apply_kernel");
  return 0;
}

extern "C" int hg_instnorm_lrelu_bwd(const float* dy, const float* x, const double* stats, float* dx,
                                     double* ws, int32_t B, int32_t HW, int32_t C, float eps,
                                     float slope, int32_t round_tf32, hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = in_args_ok(dy, x, stats, C)) return rc;
  if (!dx || !ws) return set_error(HG_EINVAL, "null tensor pointer");
  if (B <= 0 || HW <= 0) return 0;
  HG_CUDA_OK(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * (size_t)B * C, stream));
  const int cblocks = (C + 31) / 32;
  const int per = pick_pix_per_cta(HW, B, cblocks);
  dim3 grid(cblocks, (HW + per - 1) / per, B);
  instnorm_lrelu_bwd_stats_kernel<<<grid, kFusedThreads, 0, stream>>>(dy, x, stats, ws, HW, C, eps,
                                                                      slope, per);
  HG_LAUNCH_OK("instnorm_lrelu_bwd_stats_kernel");
  instnorm_lrelu_bwd_apply_kernel<<<grid, kFusedThreads, 0, stream>>>(dy, x, stats, ws, dx, HW, C, eps,
                                                                      slope, round_tf32, per);
  HG_LAUNCH_OK("instnorm_lrelu_bwd_apply_kernel");
  return 0;
}

extern "C" size_t hg_laplacian_l1_workspace_bytes(void) { return sizeof(double) * kLapBlocks; }

extern "C" int hg_laplacian_l1_fwd(const float* a, const float* b, float* loss, int8_t* sign, void* ws,
                                   size_t ws_bytes, int32_t B, int32_t H, int32_t W,
                                   hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a || !b || !loss || !sign || !ws) return set_error(HG_EINVAL, "null tensor pointer");
  if (ws_bytes < hg_laplacian_l1_workspace_bytes()) return set_error(HG_EINVAL, "workspace too small");
  if (B <= 0 || H <= 0 || W <= 0) return set_error(HG_EINVAL, "empty image batch");
  laplacian_l1_partial_kernel<<<kLapBlocks, kLapThreads, 0, stream>>>(
      a, b, reinterpret_cast<signed char*>(sign), reinterpret_cast<double*>(ws), B, H, W);
  HG_LAUNCH_OK("laplacian_l1_partial_kernel");
  laplacian_l1_final_kernel<<<1, kLapThreads, 0, stream>>>(reinterpret_cast<const double*>(ws), kLapBlocks,
                                                           loss, 1.0 / ((double)B * H * W));
  HG_LAUNCH_OK("laplacian_l1_final_kernel");
  return 0;
}

extern "C" int hg_laplacian_l1_bwd(const int8_t* sign, const float* gout, float* db, int32_t B, int32_t H,
                                   int32_t W, hg_stream_t stream_) {
  if (!sign || !gout || !db) return set_error(HG_EINVAL, "null tensor pointer");
  const long long n = (long long)B * H * W;
  if (n <= 0) return 0;
  laplacian_l1_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      reinterpret_cast<const signed char*>(sign), gout, db, B, H, W, (float)(1.0 / (double)n));
  HG_LAUNCH_OK("laplacian_l1_bwd_kernel");
  return 0;
}

extern "C" int hg_depthwise_conv(const float* x, const float* kernel, float* y, int32_t planes, int32_t H,
                                 int32_t W, int32_t K, int32_t pad, int32_t flip, hg_stream_t stream_) {
  if (!x || !kernel || !y) return set_error(HG_EINVAL, "null tensor pointer");
  if (K < 1 || K > kDwMaxK) return set_error(HG_ENOSUP, "filter size %d not in 1..%d", K, kDwMaxK);
  const int OH = H + 2 * pad - K + 1, OW = W + 2 * pad - K + 1;
  if (OH <= 0 || OW <= 0) return set_error(HG_EINVAL, "filter larger than the padded image");
  if (planes <= 0) return 0;
  if (planes > 65535) return set_error(HG_ENOSUP, "too many planes (%d)", planes);
  dim3 grid((OW + kDwTW - 1) / kDwTW, (OH + kDwTH - 1) / kDwTH, planes);
  depthwise_conv_kernel<<<grid, dim3(kDwTW, kDwTH), 0, (cudaStream_t)stream_>>>(x, kernel, y, H, W, OH, OW,
                                                                              K, pad, flip);
  HG_LAUNCH_OK("depthwise_conv_kernel");
  return 0;
}
