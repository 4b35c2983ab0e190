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

// thread = one pixel x 16 output channels (blockIdx.y selects the group of 16): the k*k*Cin image values
// are loaded once per pixel, the weights come from shared memory as warp-wide broadcasts ([ci][tap][co]),
// 32-bit index arithmetic throughout (B*H*W < 2^31).  Channels beyond Cout are written as zeros.
__global__ void __launch_bounds__(256)
conv_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      const float* __restrict__ residual, float* __restrict__ y, const SmallConvArgs a,
                      int n_pix) {
  __shared__ __align__(16) float sw[kSmallMaxCin * kSmallMaxTaps * 16];   // [ci][tap][16 co of this group]
  __shared__ float sb[16];
  const int taps = a.k * a.k;
  const int cg = blockIdx.y * 16;                    // first output channel of this thread group
  for (int i = threadIdx.x; i < a.Cin * taps * 16; i += 256) {
    const int co = i & 15, t = (i >> 4) % taps, ci = (i >> 4) / taps;
    sw[i] = cg + co < a.Cout ? w[((cg + co) * a.Cin + ci) * taps + t] : 0.f;
  }
  if (threadIdx.x < 16) sb[threadIdx.x] = (bias && cg + threadIdx.x < a.Cout) ? bias[cg + threadIdx.x] : 0.f;
  __syncthreads();
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= n_pix) return;
  const int ow = p % a.W, r = p / a.W, oh = r % a.H, b = r / a.H;
  float acc[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = sb[e];
  const int pad = a.k / 2;
  const float* xb = x + (long long)b * a.sb;
  for (int ci = 0; ci < a.Cin; ++ci) {
    for (int kh = 0; kh < a.k; ++kh) {
      const int ih = oh + kh - pad;
      if (ih < 0 || ih >= a.H) continue;
      for (int kw = 0; kw < a.k; ++kw) {
        const int iw = ow + kw - pad;
        if (iw < 0 || iw >= a.W) continue;
        const float xv = __ldg(xb + ci * a.sc + (long long)ih * a.sh + (long long)iw * a.sw);
        const float4* wp = reinterpret_cast<const float4*>(sw + (ci * taps + kh * a.k + kw) * 16);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          const float4 wv = wp[qd];
          acc[qd * 4 + 0] = fmaf(xv, wv.x, acc[qd * 4 + 0]); acc[qd * 4 + 1] = fmaf(xv, wv.y, acc[qd * 4 + 1]);
          acc[qd * 4 + 2] = fmaf(xv, wv.z, acc[qd * 4 + 2]); acc[qd * 4 + 3] = fmaf(xv, wv.w, acc[qd * 4 + 3]);
        }
      }
    }
  }
  float* yo = y + (long long)p * a.Cp + cg;
  const float* ro = residual ? residual + (long long)p * a.Cp + cg : nullptr;
#pragma unroll
  for (int qd = 0; qd < 4; ++qd) {
    if (cg + qd * 4 >= a.Cp) break;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float t = acc[qd * 4 + e];
      if (cg + qd * 4 + e >= a.Cout) t = 0.f;
      else {
        if (a.flags & HG_CONV_LRELU) t = t > 0.f ? t : t * a.slope;
        if (ro) t += ro[qd * 4 + e];                              // added after the activation (:523)
        if (a.flags & HG_CONV_ROUND_TF32) t = tf32_round(t);
      }
      v[e] = t;
    }
    *reinterpret_cast<float4*>(yo + qd * 4) = make_float4(v[0], v[1], v[2], v[3]);
  }
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
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= (int)total) return;
  const int iw = p % a.W, r = p / a.W, ih = r % a.H;
  const long long b = r / a.H;
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

// dw[co][ci][tap]: thread = (pixel lane, co quad, ci) keeps 4 x taps accumulators in registers while it
// walks its pixels (32-bit index arithmetic); per CTA the lanes are summed through shared memory in a
// fixed order, each CTA writes ONE partial vector, conv_small_wgrad_finish adds the CTAs in index order.
template <int K>
__global__ void __launch_bounds__(256)
conv_small_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ partial,
                        const SmallConvArgs a, long long n_pix_) {
  constexpr int T = K * K;
  __shared__ float red[256 * 4];
  const int n_pix = (int)n_pix_;
  const int q = a.Cout / 4;                       // co quads (<= 16)
  const int roles = q * a.Cin;                    // (co quad, ci) roles per pixel lane
  const int lanes = 256 / roles;                  // pixel lanes per CTA
  const int role = threadIdx.x % roles, pl = threadIdx.x / roles;
  const int cq = role % q, ci = role / q;
  float acc[4][T];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int t = 0; t < T; ++t) acc[e][t] = 0.f;
  constexpr int pad = K / 2;
  if (pl < lanes) {
    for (int p = blockIdx.x * lanes + pl; p < n_pix; p += gridDim.x * lanes) {
      const int ow = p % a.W, r = p / a.W, oh = r % a.H, b = r / a.H;
      const float4 g = __ldg(reinterpret_cast<const float4*>(dy + (long long)p * a.Cp + cq * 4));
      const float* xc = x + (long long)b * a.sb + ci * a.sc;
#pragma unroll
      for (int kh = 0; kh < K; ++kh) {
        const int ih = oh + kh - pad;
#pragma unroll
        for (int kw = 0; kw < K; ++kw) {
          const int iw = ow + kw - pad;
          float xv = 0.f;
          if (ih >= 0 && ih < a.H && iw >= 0 && iw < a.W) xv = __ldg(xc + (long long)ih * a.sh + (long long)iw * a.sw);
          acc[0][kh * K + kw] = fmaf(g.x, xv, acc[0][kh * K + kw]);
          acc[1][kh * K + kw] = fmaf(g.y, xv, acc[1][kh * K + kw]);
          acc[2][kh * K + kw] = fmaf(g.z, xv, acc[2][kh * K + kw]);
          acc[3][kh * K + kw] = fmaf(g.w, xv, acc[3][kh * K + kw]);
        }
      }
    }
  }
  // reduce over the pixel lanes, one tap at a time: red[thread][e]; output thread = (co, ci)
  float* out = partial + (long long)blockIdx.x * a.Cout * a.Cin * T;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) red[threadIdx.x * 4 + e] = pl < lanes ? acc[e][t] : 0.f;
    __syncthreads();
    if (threadIdx.x < a.Cout * a.Cin) {
      const int co = threadIdx.x % a.Cout, c2 = threadIdx.x / a.Cout;
      const int rl = c2 * q + co / 4, e = co % 4;
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += red[(l * roles + rl) * 4 + e];
      out[(co * a.Cin + c2) * T + t] = s;
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
  const long long n_pix = (long long)B * H * W;
  if (n_pix <= 0) return 0;
  if (n_pix >= (1LL << 31)) return set_error(HG_ENOSUP, "conv_small: too many pixels");
  dim3 grid((unsigned)((n_pix + 255) / 256), (Cp + 15) / 16);
  conv_small_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(x, w, bias, residual, y, a, (int)n_pix);
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
