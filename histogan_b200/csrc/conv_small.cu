// conv_small.cu -- the discriminator's FIRST layers on the CUDA cores: convolutions whose input is the
// 3- (or 4-) channel image itself (DiscriminatorBlock 0: conv_res 1x1 and net[0] 3x3,
// histoGAN/histoGAN.py:507-511,520-523, called on the real and the generated batch :613-617).
//
// With K = 3*k*k <= 36 multiply-adds per output these layers are pure HBM traffic.  The tensor-core path
// needs 32-channel boxes, i.e. the image zero-padded to 32 channels (a 268 MB tensor at 32 x 256^2 that
// carries 25 MB of information) and read once per conv; here the planar image is read as it is (any
// strides), the arithmetic is exact fp32, and the only large tensor touched is the NHWC output / upstream
// gradient.  Three primitives with the same meaning as their tcgen05 counterparts (ops._raw_conv /
// _raw_grad_input / _raw_grad_weight), so autograd of any order -- the gradient penalty differentiates
// d D(x)/d x once more -- composes them exactly as it composes the tensor-core ones:
//   conv_small_fwd     y[b,oh,ow,co] = act(sum x[b,ci,oh+kh-p,ow+kw-p] w[co,ci,kh,kw] + bias)   (NHWC out)
//   conv_small_dgrad   dx[b,ci,ih,iw] = sum dy[b,ih-kh+p,iw-kw+p,co] w[co,ci,kh,kw]             (planar out)
//   conv_small_wgrad   dw[co,ci,kh,kw] = sum_pix dy[pix,co] x[pix + tap, ci]   (deterministic two-stage sum)
// k in {1, 3}, stride 1, pad k/2, Cin <= 4, Cout % 4 == 0 and <= 64.
#include "hg_common.cuh"
#include "sm100_ptx.cuh"

namespace hg {

constexpr int kSmallMaxCin = 4;
constexpr int kSmallMaxTaps = 9;
constexpr int kSmallMaxCout = 64;
constexpr int kSmallWgradCtas = 148 * 6;     // two waves of 3 resident CTAs per SM

struct SmallConvArgs {
  int B, Cin, H, W, Cout, Cp, k;           // Cp = channel count of the NHWC tensor (>= Cout, % 4 == 0)
  long long sb, sc, sh, sw;                // element strides of the planar tensor (x or dx)
  int flags;
  float slope;
};

// thread = one pixel x 4 output channels, consecutive threads = consecutive channel quads, then
// consecutive pixels: a warp's store (and its residual load) is 512 contiguous bytes of the NHWC tensor,
// the k*k*Cin image values of a pixel are warp-level broadcasts / 32-byte segments that hit L1, the
// weights come from shared memory ([ci][tap][Cp], one 16-byte read per multiply group).  Channels
// beyond Cout are written as zeros.
__global__ void __launch_bounds__(256)
conv_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      const float* __restrict__ residual, float* __restrict__ y, const SmallConvArgs a,
                      int n_items) {
  __shared__ __align__(16) float sw[kSmallMaxCin * kSmallMaxTaps * kSmallMaxCout];   // [ci][tap][Cp]
  __shared__ __align__(16) float sb[kSmallMaxCout];
  const int taps = a.k * a.k;
  for (int i = threadIdx.x; i < a.Cin * taps * a.Cp; i += 256) {
    const int co = i % a.Cp, t = (i / a.Cp) % taps, ci = i / (a.Cp * taps);
    sw[i] = co < a.Cout ? w[(co * a.Cin + ci) * taps + t] : 0.f;
  }
  if (threadIdx.x < a.Cp) sb[threadIdx.x] = (bias && threadIdx.x < a.Cout) ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int item = blockIdx.x * 256 + threadIdx.x;
  if (item >= n_items) return;
  const int nq = a.Cp >> 2;
  const int p = item / nq, c0 = (item - p * nq) * 4;
  const int ow = p % a.W, r = p / a.W, oh = r % a.H, b = r / a.H;
  float4 acc = *reinterpret_cast<const float4*>(sb + c0);
  const int pad = a.k / 2;
  const float* xb = x + (long long)b * a.sb;
  const float* wq = sw + c0;
  for (int ci = 0; ci < a.Cin; ++ci) {
    for (int kh = 0; kh < a.k; ++kh) {
      const int ih = oh + kh - pad;
      if (ih < 0 || ih >= a.H) continue;
      const float* xr = xb + ci * a.sc + (long long)ih * a.sh;
      for (int kw = 0; kw < a.k; ++kw) {
        const int iw = ow + kw - pad;
        if (iw < 0 || iw >= a.W) continue;
        const float xv = __ldg(xr + (long long)iw * a.sw);
        const float4 wv = *reinterpret_cast<const float4*>(wq + (ci * taps + kh * a.k + kw) * a.Cp);
        acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y);
        acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
      }
    }
  }
  float v[4] = {acc.x, acc.y, acc.z, acc.w};
  float rv[4] = {0.f, 0.f, 0.f, 0.f};
  if (residual) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(residual) + item);
    rv[0] = t.x; rv[1] = t.y; rv[2] = t.z; rv[3] = t.w;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float t = v[e];
    if (c0 + e >= a.Cout) t = 0.f;
    else {
      if (a.flags & HG_CONV_LRELU) t = t > 0.f ? t : t * a.slope;
      t += rv[e];                                                 // added after the activation (:523)
      if (a.flags & HG_CONV_ROUND_TF32) t = tf32_round(t);
    }
    v[e] = t;
  }
  reinterpret_cast<float4*>(y)[item] = make_float4(v[0], v[1], v[2], v[3]);
}

// thread = one input pixel x 4 upstream channels (consecutive threads = consecutive quads, then pixels:
// every dy load of a warp is 512 contiguous bytes); the Cout/4 partial sums of a pixel are added in a
// fixed order through shared memory by the threads (pixel, ci), which write the planar (strided) dx.
__global__ void __launch_bounds__(256)
conv_small_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                        const SmallConvArgs a, int n_pix) {
  __shared__ __align__(16) float sw[kSmallMaxCout * kSmallMaxCin * kSmallMaxTaps];   // [tap][ci][co]
  __shared__ float red[kSmallMaxCin * 256];
  const int taps = a.k * a.k;
  for (int i = threadIdx.x; i < a.Cout * a.Cin * taps; i += 256) {
    const int t = i % taps, ci = (i / taps) % a.Cin, co = i / (taps * a.Cin);
    sw[(t * a.Cin + ci) * a.Cout + co] = w[i];
  }
  __syncthreads();
  const int q = a.Cout >> 2;                       // quads that carry information (<= 16)
  const int pix_cta = 256 / q;                     // pixels per CTA
  const int pl = threadIdx.x / q, cq = threadIdx.x - pl * q;
  const int p = blockIdx.x * pix_cta + pl;
  const bool live = pl < pix_cta && p < n_pix;
  float acc[kSmallMaxCin] = {0.f, 0.f, 0.f, 0.f};
  const int pad = a.k / 2;
  if (live) {
    const int iw = p % a.W, r = p / a.W, ih = r % a.H;
    for (int kh = 0; kh < a.k; ++kh) {
      const int oh = ih - kh + pad;
      if (oh < 0 || oh >= a.H) continue;
      for (int kw = 0; kw < a.k; ++kw) {
        const int ow = iw - kw + pad;
        if (ow < 0 || ow >= a.W) continue;
        const int po = p + (oh - ih) * a.W + (ow - iw);
        const float4 gv = __ldg(reinterpret_cast<const float4*>(dy + (long long)po * a.Cp + cq * 4));
        const float* wt = sw + (kh * a.k + kw) * a.Cin * a.Cout + cq * 4;
#pragma unroll
        for (int ci = 0; ci < kSmallMaxCin; ++ci) {
          if (ci < a.Cin) {
            const float4 wv = *reinterpret_cast<const float4*>(wt + ci * a.Cout);
            acc[ci] = fmaf(gv.x, wv.x, fmaf(gv.y, wv.y, fmaf(gv.z, wv.z, fmaf(gv.w, wv.w, acc[ci]))));
          }
        }
      }
    }
  }
#pragma unroll
  for (int ci = 0; ci < kSmallMaxCin; ++ci)
    if (ci < a.Cin) red[ci * 256 + threadIdx.x] = acc[ci];
  __syncthreads();
  // output thread = (ci, local pixel): consecutive threads -> consecutive pixels of one plane
  for (int o = threadIdx.x; o < a.Cin * pix_cta; o += 256) {
    const int ci = o / pix_cta, l = o - ci * pix_cta;
    const int po = blockIdx.x * pix_cta + l;
    if (po >= n_pix) continue;
    float s = 0.f;
    for (int j = 0; j < q; ++j) s += red[ci * 256 + l * q + j];
    const int iw = po % a.W, r = po / a.W, ih = r % a.H, b = r / a.H;
    dx[(long long)b * a.sb + ci * a.sc + (long long)ih * a.sh + (long long)iw * a.sw] = s;
  }
}

// dw[co][ci][tap]: thread = (ci, pixel lane, co quad) -- quads fastest, so a warp's dy load is
// contiguous and its image loads are consecutive pixels of one plane -- keeps 4 x taps accumulators in
// registers while it walks the CTA's contiguous pixel range; per CTA the lanes are summed through shared
// memory in a fixed order, each CTA writes ONE partial vector, conv_small_wgrad_finish adds the CTAs in a
// fixed order.
template <int K>
__global__ void __launch_bounds__(256, 3)
conv_small_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ partial,
                        const SmallConvArgs a, int n_pix) {
  constexpr int T = K * K;
  __shared__ float red[256 * 4];
  const int q = a.Cout / 4;                       // co quads (<= 16)
  const int lanes = 256 / (q * a.Cin);            // pixel lanes per CTA
  const int per_ci = lanes * q;
  const int ci = threadIdx.x / per_ci, rem = threadIdx.x - ci * per_ci;
  const int pl = rem / q, cq = rem - pl * q;
  const bool live = ci < a.Cin;
  float acc[4][T];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int t = 0; t < T; ++t) acc[e][t] = 0.f;
  constexpr int pad = K / 2;
  const int chunk = (n_pix + gridDim.x - 1) / gridDim.x;
  const int p_begin = blockIdx.x * chunk, p_end = min(n_pix, p_begin + chunk);
  if (live) {
    for (int p = p_begin + pl; p < p_end; p += lanes) {
      const int ow = p % a.W, r = p / a.W, oh = r % a.H, b = r / a.H;
      const float4 g = __ldg(reinterpret_cast<const float4*>(dy + (long long)p * a.Cp + cq * 4));
      const float* xc = x + (long long)b * a.sb + ci * a.sc;
      float xv[T];
#pragma unroll
      for (int kh = 0; kh < K; ++kh) {
        const int ih = oh + kh - pad;
#pragma unroll
        for (int kw = 0; kw < K; ++kw) {
          const int iw = ow + kw - pad;
          xv[kh * K + kw] = (ih >= 0 && ih < a.H && iw >= 0 && iw < a.W)
                                ? __ldg(xc + (long long)ih * a.sh + (long long)iw * a.sw) : 0.f;
        }
      }
#pragma unroll
      for (int t = 0; t < T; ++t) {
        acc[0][t] = fmaf(g.x, xv[t], acc[0][t]);
        acc[1][t] = fmaf(g.y, xv[t], acc[1][t]);
        acc[2][t] = fmaf(g.z, xv[t], acc[2][t]);
        acc[3][t] = fmaf(g.w, xv[t], acc[3][t]);
      }
    }
  }
  // reduce over the pixel lanes, one tap at a time: red[thread][e]; output thread = (co, ci)
  float* out = partial + (long long)blockIdx.x * a.Cout * a.Cin * T;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) red[threadIdx.x * 4 + e] = live ? acc[e][t] : 0.f;
    __syncthreads();
    if (threadIdx.x < a.Cout * a.Cin) {
      const int co = threadIdx.x % a.Cout, c2 = threadIdx.x / a.Cout;
      const int cq2 = co / 4, e = co % 4;
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += red[(c2 * per_ci + l * q + cq2) * 4 + e];
      out[(co * a.Cin + c2) * T + t] = s;
    }
  }
}

// one warp per weight: lanes stride over the CTA partials, then a fixed xor tree
__global__ void conv_small_wgrad_finish_kernel(const float* __restrict__ partial, float* __restrict__ dw, int n,
                                               int ctas) {
  const int i = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  float s = 0.f;
  for (int c = lane; c < ctas; c += 32) s += partial[(long long)c * n + i];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) dw[i] = s;
}

static int small_args(SmallConvArgs& a, int B, int Cin, int H, int W, int Cout, int Cp, int k, int64_t sb,
                      int64_t sc, int64_t sh, int64_t sw) {
  if (Cin < 1 || Cin > kSmallMaxCin) return set_error(HG_ENOSUP, "conv_small: Cin=%d not in [1, 4]", Cin);
  if (Cout % 4 || Cout < 4 || Cout > kSmallMaxCout) return set_error(HG_ENOSUP, "conv_small: Cout=%d (multiple of 4, <= 64)", Cout);
  if (k != 1 && k != 3) return set_error(HG_ENOSUP, "conv_small: k=%d (1 or 3)", k);
  if (Cp % 4 || Cp < Cout || Cp > kSmallMaxCout)
    return set_error(HG_EINVAL, "conv_small: Cp=%d must be a multiple of 4 in [Cout, 64]", Cp);
  a.B = B; a.Cin = Cin; a.H = H; a.W = W; a.Cout = Cout; a.Cp = Cp; a.k = k;
  a.sb = sb; a.sc = sc; a.sh = sh; a.sw = sw; a.flags = 0; a.slope = 0.2f;
  return 0;
}

}  // namespace hg

using namespace hg;

extern "C" int hg_conv_small_fwd(const float* x, const float* w, const float* bias, const float* residual,
                                 float* y, int32_t B,
                                 int32_t Cin, int32_t H, int32_t W, int32_t Cout, int32_t Cp, int32_t k,
                                 int64_t sb, int64_t sc, int64_t sh, int64_t sw, int32_t flags, float slope,
                                 hg_stream_t stream_) {
  if (!x || !w || !y) return set_error(HG_EINVAL, "null tensor pointer");
  SmallConvArgs a;
  int rc = small_args(a, B, Cin, H, W, Cout, Cp, k, sb, sc, sh, sw);
  if (rc) return rc;
  a.flags = flags; a.slope = slope;
  const long long n_pix = (long long)B * H * W;
  if (n_pix <= 0) return 0;
  if (n_pix >= (1LL << 31)) return set_error(HG_ENOSUP, "conv_small: too many pixels");
  const long long n_items = n_pix * (Cp / 4);
  if (n_items >= (1LL << 31)) return set_error(HG_ENOSUP, "conv_small: too many outputs");
  conv_small_fwd_kernel<<<(unsigned)((n_items + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(x, w, bias, residual, y,
                                                                                            a, (int)n_items);
  HG_LAUNCH_OK("conv_small_fwd_kernel");
  return 0;
}

extern "C" int hg_conv_small_dgrad(const float* dy, const float* w, float* dx, int32_t B, int32_t Cin,
                                   int32_t H, int32_t W, int32_t Cout, int32_t Cp, int32_t k, int64_t sb,
                                   int64_t sc, int64_t sh, int64_t sw, hg_stream_t stream_) {
  if (!dy || !w || !dx) return set_error(HG_EINVAL, "null tensor pointer");
  SmallConvArgs a;
  int rc = small_args(a, B, Cin, H, W, Cout, Cp, k, sb, sc, sh, sw);
  if (rc) return rc;
  const long long total = (long long)B * H * W;
  if (total <= 0) return 0;
  if (total >= (1LL << 31)) return set_error(HG_ENOSUP, "conv_small: too many pixels");
  const int pix_cta = 256 / (Cout / 4);
  conv_small_dgrad_kernel<<<(unsigned)((total + pix_cta - 1) / pix_cta), 256, 0, (cudaStream_t)stream_>>>(
      dy, w, dx, a, (int)total);
  HG_LAUNCH_OK("conv_small_dgrad_kernel");
  return 0;
}

extern "C" size_t hg_conv_small_wgrad_workspace_bytes(int32_t Cin, int32_t Cout, int32_t k) {
  return sizeof(float) * (size_t)kSmallWgradCtas * Cout * Cin * k * k;
}

extern "C" int hg_conv_small_wgrad(const float* dy, const float* x, float* dw, void* ws, size_t ws_bytes,
                                   int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t Cout, int32_t Cp,
                                   int32_t k, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
                                   hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!dy || !x || !dw || !ws) return set_error(HG_EINVAL, "null tensor pointer");
  SmallConvArgs a;
  int rc = small_args(a, B, Cin, H, W, Cout, Cp, k, sb, sc, sh, sw);
  if (rc) return rc;
  const int n = Cout * Cin * k * k;
  const long long n_pix = (long long)B * H * W;
  if (n_pix <= 0) { HG_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * n, stream)); return 0; }
  const int ctas = kSmallWgradCtas;
  if (n_pix >= (1LL << 31)) return set_error(HG_ENOSUP, "conv_small: too many pixels");
  if (ws_bytes < hg_conv_small_wgrad_workspace_bytes(Cin, Cout, k))
    return set_error(HG_EWS, "conv_small_wgrad: workspace too small");
  if (k == 3) conv_small_wgrad_kernel<3><<<ctas, 256, 0, stream>>>(dy, x, (float*)ws, a, (int)n_pix);
  else conv_small_wgrad_kernel<1><<<ctas, 256, 0, stream>>>(dy, x, (float*)ws, a, (int)n_pix);
  HG_LAUNCH_OK("conv_small_wgrad_kernel");
  conv_small_wgrad_finish_kernel<<<(n + 7) / 8, 256, 0, stream>>>((const float*)ws, dw, n, ctas);
  HG_LAUNCH_OK("conv_small_wgrad_finish_kernel");
  return 0;
}
