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

struct SmallConvArgs {
  int B, Cin, H, W, Cout, Cp, k;           // Cp = channel count of the NHWC tensor (>= Cout, % 4 == 0)
  long long sb, sc, sh, sw;                // element strides of the planar tensor (x or dx)
  int flags;
  float slope;
};

// thread = (pixel, channel quad of the NHWC output); the quads beyond Cout write zeros (padding channels)
__global__ void __launch_bounds__(256)
conv_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      const float* __restrict__ residual, float* __restrict__ y, const SmallConvArgs a,
                      long long total) {
  __shared__ float sw[kSmallMaxCout * kSmallMaxCin * kSmallMaxTaps];   // [co][ci][tap]
  __shared__ float sb[kSmallMaxCout];
  const int taps = a.k * a.k, wn = a.Cout * a.Cin * taps;
  for (int i = threadIdx.x; i < wn; i += 256) sw[i] = w[i];
  for (int i = threadIdx.x; i < a.Cout; i += 256) sb[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int q = a.Cp / 4;
  const int c0 = (int)(i % q) * 4;
  const long long p = i / q;
  const int ow = (int)(p % a.W), oh = (int)((p / a.W) % a.H);
  const long long b = p / ((long long)a.W * a.H);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (c0 < a.Cout) {
    const int pad = a.k / 2;
    const float* xb = x + b * a.sb;
#pragma unroll 1
    for (int ci = 0; ci < a.Cin; ++ci) {
      for (int kh = 0; kh < a.k; ++kh) {
        const int ih = oh + kh - pad;
        if (ih < 0 || ih >= a.H) continue;
        for (int kw = 0; kw < a.k; ++kw) {
          const int iw = ow + kw - pad;
          if (iw < 0 || iw >= a.W) continue;
          const float xv = __ldg(xb + ci * a.sc + (long long)ih * a.sh + (long long)iw * a.sw);
          const int t = kh * a.k + kw;
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[e] = fmaf(xv, sw[((c0 + e) * a.Cin + ci) * taps + t], acc[e]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v = acc[e] + sb[c0 + e];
      if (a.flags & HG_CONV_LRELU) v = v > 0.f ? v : v * a.slope;
      if (residual) v += residual[p * a.Cp + c0 + e];          // added after the activation (:523)
      if (a.flags & HG_CONV_ROUND_TF32) v = tf32_round(v);
      acc[e] = v;
    }
  }
  *reinterpret_cast<float4*>(y + p * a.Cp + c0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// thread = one input pixel: all Cin channels of dx (planar, strided)
__global__ void __launch_bounds__(256)
conv_small_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                        const SmallConvArgs a, long long total) {
  __shared__ float sw[kSmallMaxCout * kSmallMaxCin * kSmallMaxTaps];   // [tap][ci][co]  (co contiguous)
  const int taps = a.k * a.k;
  for (int i = threadIdx.x; i < a.Cout * a.Cin * taps; i += 256) {
    const int t = i % taps, ci = (i / taps) % a.Cin, co = i / (taps * a.Cin);
    sw[(t * a.Cin + ci) * a.Cout + co] = w[i];
  }
  __syncthreads();
  const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= total) return;
  const int iw = (int)(p % a.W), ih = (int)((p / a.W) % a.H);
  const long long b = p / ((long long)a.W * a.H);
  const int pad = a.k / 2;
  float acc[kSmallMaxCin] = {0.f, 0.f, 0.f, 0.f};
  for (int kh = 0; kh < a.k; ++kh) {
    const int oh = ih - kh + pad;
    if (oh < 0 || oh >= a.H) continue;
    for (int kw = 0; kw < a.k; ++kw) {
      const int ow = iw - kw + pad;
      if (ow < 0 || ow >= a.W) continue;
      const float* g = dy + ((b * a.H + oh) * a.W + ow) * a.Cp;
      const float* wt = sw + (kh * a.k + kw) * a.Cin * a.Cout;
      for (int co = 0; co < a.Cout; co += 4) {
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g + co));
#pragma unroll
        for (int ci = 0; ci < kSmallMaxCin; ++ci) {
          if (ci < a.Cin) {
            const float* wc = wt + ci * a.Cout + co;
            acc[ci] = fmaf(gv.x, wc[0], fmaf(gv.y, wc[1], fmaf(gv.z, wc[2], fmaf(gv.w, wc[3], acc[ci]))));
          }
        }
      }
    }
  }
  float* o = dx + b * a.sb + (long long)ih * a.sh + (long long)iw * a.sw;
#pragma unroll
  for (int ci = 0; ci < kSmallMaxCin; ++ci)
    if (ci < a.Cin) o[ci * a.sc] = acc[ci];
}

// dw[co][ci][tap]: thread = (pixel lane, co quad) keeps 4 x Cin x taps accumulators in registers while it
// walks its pixels; per CTA the lanes are summed through shared memory (fixed order), each CTA writes ONE
// partial vector, conv_small_wgrad_finish adds the CTAs in index order.
template <int K>
__global__ void __launch_bounds__(256)
conv_small_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ partial,
                        const SmallConvArgs a, long long n_pix) {
  constexpr int T = K * K;
  __shared__ float red[256 * 4];
  const int q = a.Cout / 4;                       // co quads (<= 16)
  const int lanes = 256 / q;                      // pixel lanes per CTA
  const int cq = threadIdx.x % q, pl = threadIdx.x / q;
  float acc[4][kSmallMaxCin][T];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int ci = 0; ci < kSmallMaxCin; ++ci)
#pragma unroll
      for (int t = 0; t < T; ++t) acc[e][ci][t] = 0.f;
  const int pad = K / 2;
  if (pl < lanes) {
    for (long long p = (long long)blockIdx.x * lanes + pl; p < n_pix; p += (long long)gridDim.x * lanes) {
      const int ow = (int)(p % a.W), oh = (int)((p / a.W) % a.H);
      const long long b = p / ((long long)a.W * a.H);
      const float4 g = __ldg(reinterpret_cast<const float4*>(dy + p * a.Cp + cq * 4));
      const float gv[4] = {g.x, g.y, g.z, g.w};
      const float* xb = x + b * a.sb;
#pragma unroll
      for (int ci = 0; ci < kSmallMaxCin; ++ci) {
        if (ci >= a.Cin) break;
#pragma unroll
        for (int kh = 0; kh < K; ++kh) {
          const int ih = oh + kh - pad;
#pragma unroll
          for (int kw = 0; kw < K; ++kw) {
            const int iw = ow + kw - pad;
            float xv = 0.f;
            if (ih >= 0 && ih < a.H && iw >= 0 && iw < a.W)
              xv = __ldg(xb + ci * a.sc + (long long)ih * a.sh + (long long)iw * a.sw);
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e][ci][kh * K + kw] = fmaf(gv[e], xv, acc[e][ci][kh * K + kw]);
          }
        }
      }
    }
  }
  // reduce over the pixel lanes, one (ci, tap) at a time: red[pl][cq][e]
  float* out = partial + (long long)blockIdx.x * a.Cout * a.Cin * T;
#pragma unroll 1
  for (int ci = 0; ci < a.Cin; ++ci) {
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = 0.f;
#pragma unroll
        for (int c2 = 0; c2 < kSmallMaxCin; ++c2)
#pragma unroll
          for (int t2 = 0; t2 < T; ++t2)
            if (c2 == ci && t2 == t) v = acc[e][c2][t2];      // constant indices keep acc in registers
        red[threadIdx.x * 4 + e] = (pl < lanes) ? v : 0.f;
      }
      __syncthreads();
      if (threadIdx.x < a.Cout) {                   // thread = co
        const int co = threadIdx.x, cqq = co / 4, e = co % 4;
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += red[(l * q + cqq) * 4 + e];
        out[(co * a.Cin + ci) * T + t] = s;
      }
    }
  }
}

__global__ void conv_small_wgrad_finish_kernel(const float* __restrict__ partial, float* __restrict__ dw, int n,
                                               int ctas) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int c = 0; c < ctas; ++c) s += partial[(long long)c * n + i];
  dw[i] = s;
}

static int small_args(SmallConvArgs& a, int B, int Cin, int H, int W, int Cout, int Cp, int k, int64_t sb,
                      int64_t sc, int64_t sh, int64_t sw) {
  if (Cin < 1 || Cin > kSmallMaxCin) return set_error(HG_ENOSUP, "conv_small: Cin=%d not in [1, 4]", Cin);
  if (Cout % 4 || Cout < 4 || Cout > kSmallMaxCout) return set_error(HG_ENOSUP, "conv_small: Cout=%d (multiple of 4, <= 64)", Cout);
  if (k != 1 && k != 3) return set_error(HG_ENOSUP, "conv_small: k=%d (1 or 3)", k);
  if (Cp % 4 || Cp < Cout) return set_error(HG_EINVAL, "conv_small: Cp=%d must be a multiple of 4 and >= Cout", Cp);
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
  const long long total = (long long)B * H * W * (Cp / 4);
  if (total <= 0) return 0;
  conv_small_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(x, w, bias, residual, y, a,
                                                                                         total);
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
  conv_small_dgrad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(dy, w, dx, a, total);
  HG_LAUNCH_OK("conv_small_dgrad_kernel");
  return 0;
}

extern "C" size_t hg_conv_small_wgrad_workspace_bytes(int32_t Cin, int32_t Cout, int32_t k) {
  return sizeof(float) * (size_t)(2 * 148) * Cout * Cin * k * k;
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
  const int ctas = 2 * 148;
  if (ws_bytes < hg_conv_small_wgrad_workspace_bytes(Cin, Cout, k))
    return set_error(HG_EWS, "conv_small_wgrad: workspace too small");
  if (k == 3) conv_small_wgrad_kernel<3><<<ctas, 256, 0, stream>>>(dy, x, (float*)ws, a, n_pix);
  else conv_small_wgrad_kernel<1><<<ctas, 256, 0, stream>>>(dy, x, (float*)ws, a, n_pix);
  HG_LAUNCH_OK("conv_small_wgrad_kernel");
  conv_small_wgrad_finish_kernel<<<(n + 255) / 256, 256, 0, stream>>>((const float*)ws, dw, n, ctas);
  HG_LAUNCH_OK("conv_small_wgrad_finish_kernel");
  return 0;
}
