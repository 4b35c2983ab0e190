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
#include <mutex>

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

// thread = one pixel x 16 output channels (blockIdx.y selects the group of 16): the k*k*Cin image values
// are loaded once per pixel, the weights come from shared memory as warp-wide broadcasts ([ci][tap][co]),
// 32-bit index arithmetic throughout (B*H*W < 2^31).  Channels beyond Cout are written as zeros.
__global__ void __launch_bounds__(256)
conv_small_fwd_generic_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
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
conv_small_dgrad_generic_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
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

// ---------------------------------------------------------------------------------------------
// The layer that matters (DiscriminatorBlock 0 at network_capacity 16: 3 -> 16 channels, dense 16-channel
// NHWC output) has its own kernels.  Measured on the generic ones (ncu, 32 x 256^2): issue slots 84 % busy,
// DRAM at 3 % -- the weights came from shared memory, one LDS.128 per four multiply-adds, and the
// address arithmetic was per tap.  Here the 432 weights sit in CONSTANT memory (a device-to-device copy
// node in front of the kernel, capturable) and, with every loop unrolled, each multiply-add names its
// weight as an immediate constant-bank operand: no load instruction at all.
constexpr int kFastCin = 3, kFastCout = 16;
__constant__ __align__(16) float c_small_w[kFastCin * kSmallMaxTaps * kFastCout + kFastCout];   // [ci][tap][co], then bias
#define C_SMALL_B (c_small_w + kFastCin * kSmallMaxTaps * kFastCout)

// OIHW filter (+ bias or zeros) -> the constant-memory layout, into a per-device scratch
__global__ void small_pack_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                  float* __restrict__ out, int T) {
  const int i = threadIdx.x;
  if (i < kFastCin * T * kFastCout) {
    const int co = i % kFastCout, t = (i / kFastCout) % T, ci = i / (kFastCout * T);
    out[(ci * kSmallMaxTaps + t) * kFastCout + co] = w[(co * kFastCin + ci) * T + t];
  }
  if (i < kFastCout) out[kFastCin * kSmallMaxTaps * kFastCout + i] = bias ? bias[i] : 0.f;
}

// thread = one pixel x 16 channels
template <int K>
__global__ void __launch_bounds__(256)
conv_small_fwd16_kernel(const float* __restrict__ x, const float* __restrict__ residual, float* __restrict__ y,
                        const SmallConvArgs a, int n_pix) {
  constexpr int pad = K / 2;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= n_pix) return;
  const int ow = p % a.W, r = p / a.W, oh = r % a.H, b = r / a.H;
  float acc[kFastCout];
#pragma unroll
  for (int e = 0; e < kFastCout; ++e) acc[e] = C_SMALL_B[e];
  // taps outside the image read a clamped (valid) address and are zeroed by a select: no divergence, and
  // the nine tap pointers are formed once and advanced by one plane stride per input channel (the first
  // version spent 850 of its 1280 instructions per pixel on per-tap bounds checks and 64-bit address math)
  const float* tap[K][K];
  bool ok[K][K];
  {
    const float* xb = x + (long long)b * a.sb;
    const float* row[K];
    bool okh[K];
#pragma unroll
    for (int kh = 0; kh < K; ++kh) {
      const int ih = oh + kh - pad;
      okh[kh] = ih >= 0 && ih < a.H;
      row[kh] = xb + (long long)min(max(ih, 0), a.H - 1) * a.sh;
    }
#pragma unroll
    for (int kw = 0; kw < K; ++kw) {
      const int iw = ow + kw - pad;
      const bool okw = iw >= 0 && iw < a.W;
      const long long off = (long long)min(max(iw, 0), a.W - 1) * a.sw;
#pragma unroll
      for (int kh = 0; kh < K; ++kh) {
        tap[kh][kw] = row[kh] + off;
        ok[kh][kw] = okh[kh] && okw;
      }
    }
  }
#pragma unroll
  for (int ci = 0; ci < kFastCin; ++ci) {
#pragma unroll
    for (int kh = 0; kh < K; ++kh) {
#pragma unroll
      for (int kw = 0; kw < K; ++kw) {
        const float v = __ldg(tap[kh][kw]);
        const float xv = ok[kh][kw] ? v : 0.f;
        tap[kh][kw] += a.sc;
#pragma unroll
        for (int co = 0; co < kFastCout; ++co)
          acc[co] = fmaf(xv, c_small_w[(ci * kSmallMaxTaps + kh * K + kw) * kFastCout + co], acc[co]);
      }
    }
  }
  float4* yo = reinterpret_cast<float4*>(y + (long long)p * kFastCout);
  const float4* ro = residual ? reinterpret_cast<const float4*>(residual + (long long)p * kFastCout) : nullptr;
#pragma unroll
  for (int qd = 0; qd < 4; ++qd) {
    float v[4] = {acc[qd * 4], acc[qd * 4 + 1], acc[qd * 4 + 2], acc[qd * 4 + 3]};
    float rv[4] = {0.f, 0.f, 0.f, 0.f};
    if (ro) { const float4 t = __ldg(ro + qd); rv[0] = t.x; rv[1] = t.y; rv[2] = t.z; rv[3] = t.w; }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float t = v[e];
      if (a.flags & HG_CONV_LRELU) t = t > 0.f ? t : t * a.slope;
      t += rv[e];                                                  // added after the activation (:523)
      if (a.flags & HG_CONV_ROUND_TF32) t = tf32_round(t);
      v[e] = t;
    }
    yo[qd] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// CTA = a 32 x 8 tile of input pixels of one image.  The (32 + 2 pad) x (8 + 2 pad) x 16-channel patch of
// dy it needs is staged in shared memory with fully coalesced loads (pixel stride 20 floats: the 16-byte
// reads of 8 consecutive lanes then fall into 8 different bank groups); thread = one pixel, 3 sums.
constexpr int kDgTW = 32, kDgTH = 8, kDgStride = 20;
template <int K>
__global__ void __launch_bounds__(256)
conv_small_dgrad16_kernel(const float* __restrict__ dy, float* __restrict__ dx, const SmallConvArgs a) {
  constexpr int pad = K / 2, PW = kDgTW + 2 * pad, PH = kDgTH + 2 * pad;
  __shared__ __align__(16) float tile[PH * PW * kDgStride];
  const int b = blockIdx.z, w0 = blockIdx.x * kDgTW, h0 = blockIdx.y * kDgTH;
  const float* dyb = dy + (long long)b * a.H * a.W * kFastCout;
  for (int i = threadIdx.x; i < PH * PW * 4; i += 256) {
    const int q = i & 3, px = (i >> 2) % PW, py = (i >> 2) / PW;
    const int gh = h0 + py - pad, gw = w0 + px - pad;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gh >= 0 && gh < a.H && gw >= 0 && gw < a.W)
      v = __ldg(reinterpret_cast<const float4*>(dyb + ((long long)gh * a.W + gw) * kFastCout) + q);
    *reinterpret_cast<float4*>(tile + (py * PW + px) * kDgStride + q * 4) = v;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int iw = w0 + tx, ih = h0 + ty;
  float acc[kFastCin] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int kh = 0; kh < K; ++kh) {
#pragma unroll
    for (int kw = 0; kw < K; ++kw) {
      // dx[ih, iw] takes dy[ih - kh + pad, iw - kw + pad] = patch (ty + 2 pad - kh, tx + 2 pad - kw)
      const float* g = tile + ((ty + 2 * pad - kh) * PW + (tx + 2 * pad - kw)) * kDgStride;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 gv = *reinterpret_cast<const float4*>(g + q * 4);
#pragma unroll
        for (int ci = 0; ci < kFastCin; ++ci) {
          acc[ci] = fmaf(gv.x, c_small_w[(ci * kSmallMaxTaps + kh * K + kw) * kFastCout + q * 4 + 0], acc[ci]);
          acc[ci] = fmaf(gv.y, c_small_w[(ci * kSmallMaxTaps + kh * K + kw) * kFastCout + q * 4 + 1], acc[ci]);
          acc[ci] = fmaf(gv.z, c_small_w[(ci * kSmallMaxTaps + kh * K + kw) * kFastCout + q * 4 + 2], acc[ci]);
          acc[ci] = fmaf(gv.w, c_small_w[(ci * kSmallMaxTaps + kh * K + kw) * kFastCout + q * 4 + 3], acc[ci]);
        }
      }
    }
  }
  if (iw < a.W && ih < a.H) {
    float* o = dx + (long long)b * a.sb + (long long)ih * a.sh + (long long)iw * a.sw;
#pragma unroll
    for (int ci = 0; ci < kFastCin; ++ci) o[ci * a.sc] = acc[ci];
  }
}

static bool small_fast(const SmallConvArgs& a) {
  static const bool off = [] { const char* e = getenv("HG_SMALL_GENERIC"); return e && e[0] == '1'; }();
  return !off && a.Cin == kFastCin && a.Cout == kFastCout && a.Cp == kFastCout;
}

// the filter (and bias) into constant memory, in stream order: a pack kernel into a per-device scratch
// and a device-to-device copy into the symbol (both capturable).  Returns 1 when the fast path is not
// available (first use inside a capture: the scratch cannot be allocated), < 0 on error.
static int small_fast_upload(const float* w, const float* bias, int k, cudaStream_t stream) {
  static float* scratch[64] = {};
  static std::mutex mu;
  constexpr size_t bytes = sizeof(float) * (kFastCin * kSmallMaxTaps * kFastCout + kFastCout);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 1;
  float* buf;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!scratch[dev]) {
      cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return 1;
      float* p = nullptr;
      if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return 1; }
      scratch[dev] = p;
    }
    buf = scratch[dev];
  }
  small_pack_kernel<<<1, 512, 0, stream>>>(w, bias, buf, k * k);
  HG_LAUNCH_OK("small_pack_kernel");
  HG_CUDA_OK(cudaMemcpyToSymbolAsync(c_small_w, buf, bytes, 0, cudaMemcpyDeviceToDevice, stream));
  return 0;
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

// 3 -> 16, 3x3: thread = (row-segment lane, input channel, output quad) walks 32 consecutive pixels of one
// image row with the 3x3 window of its input plane sliding in registers: per pixel ONE 16-byte dy load and
// THREE new image values feed 36 multiply-adds (the generic kernel: 10 loads with their own bounds checks).
// 21 segment lanes x 12 roles per CTA; per-CTA partial vectors, fixed-order finish as above.
constexpr int kWgSeg = 32;
__global__ void __launch_bounds__(256, 2)
conv_small_wgrad16_k3_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ partial,
                             const SmallConvArgs a, int n_units, int segs) {
  constexpr int T = 9, ROLES = 12, LANES = 256 / ROLES;          // 21 lanes, 4 idle threads
  __shared__ float red[256 * 4];
  const int sl = threadIdx.x / ROLES, role = threadIdx.x - sl * ROLES;
  const int cq = role & 3, ci = role >> 2;
  const bool live = sl < LANES;
  float acc[4][T];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int t = 0; t < T; ++t) acc[e][t] = 0.f;
  if (live) {
    for (int u = blockIdx.x * LANES + sl; u < n_units; u += gridDim.x * LANES) {
      const int s = u % segs, r = u / segs, h = r % a.H, b = r / a.H;
      const int w0 = s * kWgSeg, w1 = min(a.W, w0 + kWgSeg);
      const float* xc = x + (long long)b * a.sb + ci * a.sc;
      const float* xr[3];
      bool ok[3];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int ih = h + kh - 1;
        ok[kh] = ih >= 0 && ih < a.H;
        xr[kh] = xc + (long long)ih * a.sh;
      }
      float c0[3], c1[3], c2[3];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        c0[kh] = (ok[kh] && w0 > 0) ? __ldg(xr[kh] + (long long)(w0 - 1) * a.sw) : 0.f;
        c1[kh] = ok[kh] ? __ldg(xr[kh] + (long long)w0 * a.sw) : 0.f;
      }
      const float4* g4 = reinterpret_cast<const float4*>(dy + ((long long)r * a.W + w0) * kFastCout) + cq;
#pragma unroll 4
      for (int w = w0; w < w1; ++w) {
        const float4 g = __ldg(g4 + (w - w0) * 4);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
          c2[kh] = (ok[kh] && w + 1 < a.W) ? __ldg(xr[kh] + (long long)(w + 1) * a.sw) : 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const float xv[3] = {c0[kh], c1[kh], c2[kh]};
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            acc[0][kh * 3 + kw] = fmaf(g.x, xv[kw], acc[0][kh * 3 + kw]);
            acc[1][kh * 3 + kw] = fmaf(g.y, xv[kw], acc[1][kh * 3 + kw]);
            acc[2][kh * 3 + kw] = fmaf(g.z, xv[kw], acc[2][kh * 3 + kw]);
            acc[3][kh * 3 + kw] = fmaf(g.w, xv[kw], acc[3][kh * 3 + kw]);
          }
          c0[kh] = c1[kh]; c1[kh] = c2[kh];
        }
      }
    }
  }
  float* out = partial + (long long)blockIdx.x * kFastCout * kFastCin * T;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) red[threadIdx.x * 4 + e] = live ? acc[e][t] : 0.f;
    __syncthreads();
    if (threadIdx.x < kFastCout * kFastCin) {
      const int co = threadIdx.x % kFastCout, c2i = threadIdx.x / kFastCout;
      const int rl = c2i * 4 + co / 4, e = co % 4;                 // role (ci, cq) of this output
      float sum = 0.f;
      for (int l = 0; l < LANES; ++l) sum += red[(l * ROLES + rl) * 4 + e];
      out[(co * kFastCin + c2i) * T + t] = sum;
    }
  }
}

// 3 -> 16, 1x1: thread = (pixel lane, output quad), consecutive threads = consecutive quads then pixels: the dy
// loads of a warp are 512 contiguous bytes, the three image values of a pixel feed 12 multiply-adds.
__global__ void __launch_bounds__(256)
conv_small_wgrad16_k1_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ partial,
                             const SmallConvArgs a, int n_pix) {
  __shared__ float red[256 * 12];
  const int pl = threadIdx.x >> 2, cq = threadIdx.x & 3;
  float acc[4][3];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[e][c] = 0.f;
  const int chunk = (n_pix + gridDim.x - 1) / gridDim.x;
  const int p_begin = blockIdx.x * chunk, p_end = min(n_pix, p_begin + chunk);
#pragma unroll 2
  for (int p = p_begin + pl; p < p_end; p += 64) {
    const int ow = p % a.W, r = p / a.W, oh = r % a.H, b = r / a.H;
    const float4 g = __ldg(reinterpret_cast<const float4*>(dy + (long long)p * kFastCout) + cq);
    const float* xp = x + (long long)b * a.sb + (long long)oh * a.sh + (long long)ow * a.sw;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float xv = __ldg(xp + c * a.sc);
      acc[0][c] = fmaf(g.x, xv, acc[0][c]); acc[1][c] = fmaf(g.y, xv, acc[1][c]);
      acc[2][c] = fmaf(g.z, xv, acc[2][c]); acc[3][c] = fmaf(g.w, xv, acc[3][c]);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int c = 0; c < 3; ++c) red[(e * 3 + c) * 256 + threadIdx.x] = acc[e][c];
  __syncthreads();
  if (threadIdx.x < kFastCout * kFastCin) {
    const int co = threadIdx.x % kFastCout, c = threadIdx.x / kFastCout;
    const float* col = red + ((co & 3) * 3 + c) * 256 + (co >> 2);
    float sum = 0.f;
    for (int l = 0; l < 64; ++l) sum += col[l * 4];
    partial[(long long)blockIdx.x * kFastCout * kFastCin + co * kFastCin + c] = sum;
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
  cudaStream_t stream = (cudaStream_t)stream_;
  if (small_fast(a) && (rc = small_fast_upload(w, bias, k, stream)) <= 0) {
    if (rc) return rc;
    const unsigned grid = (unsigned)((n_pix + 255) / 256);
    if (k == 3) conv_small_fwd16_kernel<3><<<grid, 256, 0, stream>>>(x, residual, y, a, (int)n_pix);
    else conv_small_fwd16_kernel<1><<<grid, 256, 0, stream>>>(x, residual, y, a, (int)n_pix);
    HG_LAUNCH_OK("conv_small_fwd16_kernel");
    return 0;
  }
  dim3 grid((unsigned)((n_pix + 255) / 256), (Cp + 15) / 16);
  conv_small_fwd_generic_kernel<<<grid, 256, 0, stream>>>(x, w, bias, residual, y, a, (int)n_pix);
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
  cudaStream_t stream = (cudaStream_t)stream_;
  if (small_fast(a) && B <= 65535 && (rc = small_fast_upload(w, nullptr, k, stream)) <= 0) {
    if (rc) return rc;
    dim3 grid((W + kDgTW - 1) / kDgTW, (H + kDgTH - 1) / kDgTH, B);
    if (k == 3) conv_small_dgrad16_kernel<3><<<grid, 256, 0, stream>>>(dy, dx, a);
    else conv_small_dgrad16_kernel<1><<<grid, 256, 0, stream>>>(dy, dx, a);
    HG_LAUNCH_OK("conv_small_dgrad16_kernel");
    return 0;
  }
  conv_small_dgrad_generic_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(dy, w, dx, a, total);
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
  if (small_fast(a) && k == 3) {
    const int segs = (W + kWgSeg - 1) / kWgSeg;
    const long long units = (long long)B * H * segs;
    if (units >= (1LL << 31)) return set_error(HG_ENOSUP, "conv_small: too many pixels");
    conv_small_wgrad16_k3_kernel<<<ctas, 256, 0, stream>>>(dy, x, (float*)ws, a, (int)units, segs);
  } else if (small_fast(a) && k == 1) {
    conv_small_wgrad16_k1_kernel<<<ctas, 256, 0, stream>>>(dy, x, (float*)ws, a, (int)n_pix);
  } else if (k == 3) {
    conv_small_wgrad_kernel<3><<<ctas, 256, 0, stream>>>(dy, x, (float*)ws, a, (int)n_pix);
  } else {
    conv_small_wgrad_kernel<1><<<ctas, 256, 0, stream>>>(dy, x, (float*)ws, a, (int)n_pix);
  }
  HG_LAUNCH_OK("conv_small_wgrad_kernel");
  conv_small_wgrad_finish_kernel<<<(n + 7) / 8, 256, 0, stream>>>((const float*)ws, dw, n, ctas);
  HG_LAUNCH_OK("conv_small_wgrad_finish_kernel");
  return 0;
}
