// style.cu -- the "style path" of the generator: every small dense op between the latent
// vectors and the modulated convolutions, as grouped skinny GEMMs with the batch (<= 32 rows)
// mapped onto the lanes of a warp.  One launch handles ALL layers of a generator pass.
//
//   grouped_linear_fwd   y_g = act(x_g W_g^T + b_g) (+1)         GeneratorBlock.to_style1/2,
//                                                                 RGBBlock.to_style (histoGAN.py:
//                                                                 372,451,455,461-462,381) -> mod = style + 1
//                        d_g = rsqrt((x_g^2) Wsq_g^T + eps)       Conv2DMod demodulation (:427-429)
//   grouped_linear_bwd   gW_g = gy_g^T x_g, gb_g = sum_b gy_g, gx_g = gy_g W_g   (their adjoints)
//   weight_sqsum         Wsq[co][ci] = sum_taps w[co][tap][ci]^2               (:428, shared by the batch)
//   demod_bwd            adjoint of d = rsqrt(mod^2 Wsq^T + eps) given gd: adds the mod part to gmod and
//                        the weight part 2 w gWsq straight into the weight gradient
//
// The reference evaluates these with per-sample weight tensors (B x Cout x Cin x k x k, :423-429);
// here they are O(B (Cin + Cout) + Cout Cin) per layer.  All kernels are bound by reading the weight
// matrices once (25 MB of to_style weights, 38 MB of Wsq per generator pass).
#include <mutex>
#include "hg_common.cuh"
#include "sm100_ptx.cuh"

namespace hg {

constexpr int kMaxGroups = 24;
constexpr int kKC = 512;                   // K chunk staged in shared memory
constexpr int kXS = kKC + 4;               // row stride: conflict-free 128-bit reads for lane = batch row
constexpr int kRows = 4;                   // output rows per warp step
constexpr int kStyleThreads = 256;

struct LinearGroups {
  const float* x[kMaxGroups];              // [B][K]
  const float* w[kMaxGroups];              // [J][K]
  const float* bias[kMaxGroups];           // [J] or null
  float* y[kMaxGroups];                    // [B][J]
  int J[kMaxGroups];
  int K[kMaxGroups];
  int first_block[kMaxGroups + 1];         // prefix sum of ceil(J / rows_per_cta)
  int count;
  int B;
  int flags;                               // HG_LIN_*
  float slope, eps;
  int pairs_per_warp;                      // 1..4 row pairs per warp: rows per CTA = 16 * pairs_per_warp
  int ksplit;                              // > 1 (one group, long K): gridDim.z CTAs share the K range;
  float* partial;                          //   their raw sums land in partial[z][B][J], ordered finish
};

constexpr int kRowsPerCta = 64;            // 8 warps x 4 row pairs

// lane <- sum over the 32 lanes of acc[lane] (a 31-shuffle reduce-scatter instead of 32 x 5)
__device__ __forceinline__ float warp_reduce_scatter(float (&acc)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = lane & s;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? acc[i] : acc[i + s];
      const float keep = up ? acc[i + s] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return acc[0];
}

// CTA = 64 output rows x (up to) 32 batch rows.  The x chunk (32 x 512) sits in shared memory; a
// warp takes a PAIR of W rows at a time with the lanes across K (coalesced 512 B per load, all the
// loads of the pair issued before the math), accumulates the 32 batch rows in registers and
// finishes each row with a reduce-scatter so that lane b ends up holding y[b][row].
__global__ void __launch_bounds__(kStyleThreads, 2)
grouped_linear_fwd_kernel(const LinearGroups t) {
  extern __shared__ __align__(16) float xs[];          // [32][kXS] then partial sums [64][32]
  float* part = xs + 32 * kXS;
  int gi = 0;
  while (gi + 1 < t.count && (int)blockIdx.x >= t.first_block[gi + 1]) ++gi;
  const int J = t.J[gi], K = t.K[gi];
  const int ppw = t.pairs_per_warp;
  const int r0 = (blockIdx.x - t.first_block[gi]) * 16 * ppw;
  const int b0 = blockIdx.y * 32;
  const int nb = min(32, t.B - b0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* __restrict__ x = t.x[gi] + (long long)b0 * K;
  const float* __restrict__ w = t.w[gi];
  const bool sq = t.flags & HG_LIN_SQUARE_INPUT;
  // this CTA's K range: whole chunks of kKC
  const int nchunks = (K + kKC - 1) / kKC;
  const int cper = (nchunks + t.ksplit - 1) / t.ksplit;
  const int kbeg = min(K, (int)blockIdx.z * cper * kKC), kend = min(K, kbeg + cper * kKC);
  const bool multi = kend - kbeg > kKC;                // several K chunks: partial sums go through smem

  for (int k0 = kbeg; k0 < kend; k0 += kKC) {
    const int kc = min(kKC, kend - k0);                // multiple of 4
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * (kKC / 4); e += kStyleThreads) {
      const int b = e / (kKC / 4), q = e - b * (kKC / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b < nb && q * 4 < kc) v = *reinterpret_cast<const float4*>(x + (long long)b * K + k0 + q * 4);
      if (sq) { v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w; }
      *reinterpret_cast<float4*>(xs + b * kXS + q * 4) = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int pr = 0; pr < ppw; ++pr) {
      const int j0 = r0 + (warp * ppw + pr) * 2;       // rows j0, j0 + 1
      if (j0 >= J) break;
      const float* w0 = w + (long long)j0 * K + k0;
      const float* w1 = w + (long long)min(j0 + 1, J - 1) * K + k0;
      float4 wa[4], wb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = i * 128 + lane * 4;
        const bool ok = k < kc;
        wa[i] = ok ? __ldg(reinterpret_cast<const float4*>(w0 + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        wb[i] = ok ? __ldg(reinterpret_cast<const float4*>(w1 + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float a0[32], a1[32];
#pragma unroll
      for (int b = 0; b < 32; ++b) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 xv = *reinterpret_cast<const float4*>(xs + b * kXS + i * 128 + lane * 4);
          s0 = fmaf(xv.x, wa[i].x, s0); s0 = fmaf(xv.y, wa[i].y, s0);
          s0 = fmaf(xv.z, wa[i].z, s0); s0 = fmaf(xv.w, wa[i].w, s0);
          s1 = fmaf(xv.x, wb[i].x, s1); s1 = fmaf(xv.y, wb[i].y, s1);
          s1 = fmaf(xv.z, wb[i].z, s1); s1 = fmaf(xv.w, wb[i].w, s1);
        }
        a0[b] = s0; a1[b] = s1;
      }
      float v0 = warp_reduce_scatter(a0, lane);        // lane b: row j0
      float v1 = warp_reduce_scatter(a1, lane);        //         row j0 + 1
      float* pp = part + ((warp * ppw + pr) * 2) * 32 + lane;
      if (multi) {
        if (k0 > kbeg) { v0 += pp[0]; v1 += pp[32]; }
        if (k0 + kKC < kend) { pp[0] = v0; pp[32] = v1; continue; }
      }
      if (lane >= nb) continue;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int j = j0 + r;
        if (j >= J) break;
        float v = r ? v1 : v0;
        if (t.ksplit > 1) {                            // raw partial sum; epilogue in the finish kernel
          t.partial[((long long)blockIdx.z * t.B + b0 + lane) * J + j] = v;
          continue;
        }
        if (t.bias[gi]) v += t.bias[gi][j];
        if (t.flags & HG_LIN_RSQRT_EPS) v = rsqrtf(v + t.eps);
        if (t.flags & HG_LIN_LRELU) v = v > 0.f ? v : v * t.slope;
        if (t.flags & HG_LIN_ADD_ONE) v += 1.f;
        t.y[gi][(long long)(b0 + lane) * J + j] = v;
      }
    }
  }
}

// y[b][j] = f(sum over the K splits, in index order, + bias)
__global__ void __launch_bounds__(256)
grouped_linear_fwd_finish_kernel(const LinearGroups t) {
  const long long n = (long long)t.B * t.J[0];
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float v = t.partial[i];
  for (int z = 1; z < t.ksplit; ++z) v += t.partial[z * n + i];
  if (t.bias[0]) v += t.bias[0][i % t.J[0]];
  if (t.flags & HG_LIN_RSQRT_EPS) v = rsqrtf(v + t.eps);
  if (t.flags & HG_LIN_LRELU) v = v > 0.f ? v : v * t.slope;
  if (t.flags & HG_LIN_ADD_ONE) v += 1.f;
  t.y[0][i] = v;
}

// ------------------------------------------------------------------ backward --
struct LinearBwdGroups {
  const float* x[kMaxGroups];              // [B][K]
  const float* w[kMaxGroups];              // [J][K]
  const float* gy[kMaxGroups];             // [B][J]
  float* gw[kMaxGroups];                   // [J][K]   (or null)
  float* gb[kMaxGroups];                   // [J]      (or null)
  float* gx[kMaxGroups];                   // [B][K]   (or null); accumulate: see flags
  int J[kMaxGroups];
  int K[kMaxGroups];
  int first_block[kMaxGroups + 1];
  int count;
  int B;                                   // <= 32
  int flags;
  int jsplit;                              // dgrad: the J rows are divided over gridDim.y CTAs ...
  float* partial;                          // ... whose sums land in partial[(split * nslab + slab) * 1024 + b*32 + k]
};

// gW[j][k..k+3] = sum_b gy[b][j] x[b][k..k+3];  gb[j] = sum_b gy[b][j].
// CTA = 8 rows j x one 512-wide K chunk; thread = (row, float4 of k).  Write-bound.
__global__ void __launch_bounds__(kStyleThreads)
grouped_linear_wgrad_kernel(const LinearBwdGroups t) {
  __shared__ float gys[8][32];
  int gi = 0;
  while (gi + 1 < t.count && (int)blockIdx.x >= t.first_block[gi + 1]) ++gi;
  const int J = t.J[gi], K = t.K[gi];
  const int kchunks = (K + kKC - 1) / kKC;
  const int blk = blockIdx.x - t.first_block[gi];
  const int j0 = (blk / kchunks) * 8, k0 = (blk % kchunks) * kKC;
  const int B = t.B;
  {
    const int r = threadIdx.x >> 5, b = threadIdx.x & 31;
    gys[r][b] = (b < B && j0 + r < J) ? t.gy[gi][(long long)b * J + j0 + r] : 0.f;
  }
  __syncthreads();
  const float* __restrict__ x = t.x[gi];
  const bool sq = t.flags & HG_LIN_SQUARE_INPUT;
  const int q = threadIdx.x & 127, rh = threadIdx.x >> 7;          // 128 float4 columns x 2 row halves
  const int k = k0 + q * 4;
  if (k < K) {
    float4 acc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
      float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long long)b * K + k));
      if (sq) { xv.x *= xv.x; xv.y *= xv.y; xv.z *= xv.z; xv.w *= xv.w; }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float g = gys[rh * 4 + r][b];
        acc[r].x = fmaf(g, xv.x, acc[r].x); acc[r].y = fmaf(g, xv.y, acc[r].y);
        acc[r].z = fmaf(g, xv.z, acc[r].z); acc[r].w = fmaf(g, xv.w, acc[r].w);
      }
    }
    if (t.gw[gi]) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int j = j0 + rh * 4 + r;
        if (j < J) *reinterpret_cast<float4*>(t.gw[gi] + (long long)j * K + k) = acc[r];
      }
    }
  }
  if (t.gb[gi] && k0 == 0 && threadIdx.x < 8 && j0 + threadIdx.x < J) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += gys[threadIdx.x][b];
    t.gb[gi][j0 + threadIdx.x] = s;
  }
}

// gx[b][k] (+)= post[b][k] * sum_j gy[b][j] W[j][k]:  CTA = one 32-wide K slab of one group (lanes = k)
// x one share of the J rows (gridDim.y shares); its 8 warps split those rows and are reduced in
// shared memory in a fixed order.  With several shares the CTA sums go to a scratch buffer and
// grouped_linear_dgrad_finish adds them in share order -- deterministic, no atomics.
// post (HG_LIN_POST_2X): multiply by 2 x[b][k] (the demodulation's  d(mod^2)/d mod).
__device__ __forceinline__ void dgrad_store(const LinearBwdGroups& t, int gi, int b, int k, int K, float s) {
  const long long o = (long long)b * K + k;
  if (t.flags & HG_LIN_POST_2X) s *= 2.f * t.x[gi][o];
  if (t.flags & HG_LIN_ACCUMULATE) s += t.gx[gi][o];
  t.gx[gi][o] = s;
}

__global__ void __launch_bounds__(kStyleThreads)
grouped_linear_dgrad_kernel(const LinearBwdGroups t) {
  extern __shared__ __align__(16) float sm[];          // gy tile [32 b][kJT + 1] then reduction [8][32 b][32 k]
  constexpr int kJT = 256;
  float* gys = sm;
  float* red = sm + 32 * (kJT + 1);
  int gi = 0;
  while (gi + 1 < t.count && (int)blockIdx.x >= t.first_block[gi + 1]) ++gi;
  const int J = t.J[gi], K = t.K[gi], B = t.B;
  const int k = (blockIdx.x - t.first_block[gi]) * 32 + (threadIdx.x & 31);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* __restrict__ w = t.w[gi];
  // this CTA's share of the rows (multiples of 8 so that the warps stay balanced)
  const int per = ((J + t.jsplit - 1) / t.jsplit + 7) / 8 * 8;
  const int jlo = min(J, (int)blockIdx.y * per), jhi = min(J, jlo + per);
  float acc[32];
#pragma unroll
  for (int b = 0; b < 32; ++b) acc[b] = 0.f;
  for (int jt = jlo; jt < jhi; jt += kJT) {
    const int jn = min(kJT, jhi - jt);
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * jn; e += kStyleThreads) {
      const int b = e / jn, j = e - b * jn;
      gys[b * (kJT + 1) + j] = b < B ? t.gy[gi][(long long)b * J + jt + j] : 0.f;
    }
    __syncthreads();
    if (k < K) {
      // 8 rows per step, their loads issued together (one 128 B line per row and warp)
      for (int j = warp; j < jn; j += 64) {
        float wv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          wv[u] = j + 8 * u < jn ? __ldg(w + (long long)(jt + j + 8 * u) * K + k) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int jj = min(j + 8 * u, jn - 1);
#pragma unroll
          for (int b = 0; b < 32; ++b) acc[b] = fmaf(gys[b * (kJT + 1) + jj], wv[u], acc[b]);
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int b = 0; b < 32; ++b) red[(warp * 32 + b) * 32 + lane] = acc[b];
  __syncthreads();
  // thread (warp, lane = k): batch rows b = warp, warp + 8, ...
  if (!t.gx[gi]) return;
  for (int b = warp; b < 32; b += 8) {
    float s = 0.f;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) s += red[(wq * 32 + b) * 32 + lane];
    if (t.jsplit > 1)
      t.partial[((long long)blockIdx.y * gridDim.x + blockIdx.x) * 1024 + b * 32 + lane] = s;
    else if (k < K && b < B)
      dgrad_store(t, gi, b, k, K, s);
  }
}

// several J shares: gx = ordered sum of the shares' partial sums.  grid = slabs (as above), 1024 threads
__global__ void __launch_bounds__(1024)
grouped_linear_dgrad_finish_kernel(const LinearBwdGroups t, int nslab) {
  int gi = 0;
  while (gi + 1 < t.count && (int)blockIdx.x >= t.first_block[gi + 1]) ++gi;
  const int K = t.K[gi];
  const int b = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = (blockIdx.x - t.first_block[gi]) * 32 + lane;
  if (!t.gx[gi] || k >= K || b >= t.B) return;
  float s = 0.f;
  for (int sp = 0; sp < t.jsplit; ++sp) s += t.partial[((long long)sp * nslab + blockIdx.x) * 1024 + b * 32 + lane];
  dgrad_store(t, gi, b, k, K, s);
}

// ------------------------------------------------------------ weight helpers --
// Wsq[co][ci] = sum_tap w[co][tap][ci]^2  (w channels_last: [Cout][T][Cin]).  One pass over the weight.
__global__ void __launch_bounds__(256)
weight_sqsum_kernel(const float* __restrict__ w, float* __restrict__ wsq, int T, int Cin, long long n4) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;       // (co, ci4)
  if (i >= n4) return;
  const int q = Cin / 4;
  const long long co = i / q;
  const int ci = (int)(i - co * q) * 4;
  const float* p = w + (co * T) * Cin + ci;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int tp = 0; tp < T; ++tp) {
    const float4 v = *reinterpret_cast<const float4*>(p + (long long)tp * Cin);
    s.x = fmaf(v.x, v.x, s.x); s.y = fmaf(v.y, v.y, s.y); s.z = fmaf(v.z, v.z, s.z); s.w = fmaf(v.w, v.w, s.w);
  }
  *reinterpret_cast<float4*>(wsq + co * Cin + ci) = s;
}

// dW[co][tap][ci] += 2 w[co][tap][ci] * sum_b t[b][co] mod[b][ci]^2    (t = -1/2 gd d^3)
// thread = (co, ci4): the batch sum once, then the T taps.  CTA = 64 ci4 x 4 co.
__global__ void __launch_bounds__(256)
demod_weight_grad_kernel(float* __restrict__ dw, const float* __restrict__ w, const float* __restrict__ gd,
                         const float* __restrict__ d, const float* __restrict__ mod, int B, int Cout,
                         int T, int Cin) {
  __shared__ float ts[4][64];
  const int q = threadIdx.x & 63, r = threadIdx.x >> 6;
  const int co = blockIdx.y * 4 + r;
  const int ci = (blockIdx.x * 64 + q) * 4;
  for (int e = threadIdx.x; e < 4 * B; e += 256) {
    const int rr = e / B, b = e - rr * B;
    const int c = blockIdx.y * 4 + rr;
    float v = 0.f;
    if (c < Cout) {
      const float dv = d[(long long)b * Cout + c];
      v = -0.5f * gd[(long long)b * Cout + c] * dv * dv * dv;
    }
    ts[rr][b] = v;
  }
  __syncthreads();
  if (co >= Cout || ci >= Cin) return;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < B; ++b) {
    const float4 m = __ldg(reinterpret_cast<const float4*>(mod + (long long)b * Cin + ci));
    const float tv = ts[r][b];
    s.x = fmaf(tv, m.x * m.x, s.x); s.y = fmaf(tv, m.y * m.y, s.y);
    s.z = fmaf(tv, m.z * m.z, s.z); s.w = fmaf(tv, m.w * m.w, s.w);
  }
  s.x *= 2.f; s.y *= 2.f; s.z *= 2.f; s.w *= 2.f;
  const long long base = ((long long)co * T) * Cin + ci;
  for (int tp = 0; tp < T; ++tp) {
    const long long o = base + (long long)tp * Cin;
    const float4 wv = *reinterpret_cast<const float4*>(w + o);
    float4 g = *reinterpret_cast<const float4*>(dw + o);
    g.x = fmaf(wv.x, s.x, g.x); g.y = fmaf(wv.y, s.y, g.y); g.z = fmaf(wv.z, s.z, g.z); g.w = fmaf(wv.w, s.w, g.w);
    *reinterpret_cast<float4*>(dw + o) = g;
  }
}

// t[b][co] = -1/2 gd d^3  (so that the demod adjoint w.r.t. mod is a grouped_linear_dgrad with W = Wsq)
__global__ void demod_t_kernel(const float* __restrict__ gd, const float* __restrict__ d, float* __restrict__ t,
                               int n) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) {
    const float dv = d[i];
    t[i] = -0.5f * gd[i] * dv * dv * dv;
  }
}

// per-device scratch for the dgrad J-split partial sums (4 KB per CTA; a few MB at most): allocated on
// first use outside a stream capture, never moved.  Launches of one device are issued on one stream.
constexpr size_t kStyleWsBytes = 32u << 20;

static float* style_workspace(size_t bytes, cudaStream_t stream) {
  static float* ws[64] = {};
  static std::mutex mu;
  if (bytes > kStyleWsBytes) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  if (ws[dev]) return ws[dev];
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return nullptr;
  float* p = nullptr;
  if (cudaMalloc(&p, kStyleWsBytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  ws[dev] = p;
  return p;
}

static int check_groups(int count, int B, const int32_t* J, const int32_t* K) {
  if (count < 0 || count > kMaxGroups) return set_error(HG_EINVAL, "group count %d not in [0, %d]", count, kMaxGroups);
  if (B < 0) return set_error(HG_EINVAL, "negative batch");
  for (int i = 0; i < count; ++i)
    if (J[i] <= 0 || K[i] <= 0 || K[i] % 4) return set_error(HG_ENOSUP, "group %d: J=%d, K=%d (K %% 4 == 0 required)", i, J[i], K[i]);
  return 0;
}

}  // namespace hg

using namespace hg;

extern "C" int hg_grouped_linear_fwd(int32_t count, const float* const* x, const float* const* w,
                                     const float* const* bias, float* const* y, const int32_t* J,
                                     const int32_t* K, int32_t B, int32_t flags, float slope, float eps,
                                     hg_stream_t stream_) {
  int rc = check_groups(count, B, J, K);
  if (rc) return rc;
  if (count == 0 || B == 0) return 0;
  LinearGroups t{};
  // rows per CTA: 64 when that still gives >= 2 CTAs per SM, else 32 / 16 (a 512-row layer of the
  // S / H MLPs would otherwise run on 8 CTAs; the 1024 x 12288 first layer of H on 16)
  long long rows = 0;
  for (int i = 0; i < count; ++i) rows += J[i];
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  int ppw = 4;
  while (ppw > 1 && rows / (16 * ppw) < 2 * sms) ppw >>= 1;
  t.pairs_per_warp = ppw;
  const int rpc = 16 * ppw;
  int blocks = 0;
  for (int i = 0; i < count; ++i) {
    if (!x[i] || !w[i] || !y[i]) return set_error(HG_EINVAL, "null pointer in group %d", i);
    if (((uintptr_t)x[i] | (uintptr_t)w[i]) & 15) return set_error(HG_EINVAL, "group %d: x / W must be 16-byte aligned", i);
    t.x[i] = x[i]; t.w[i] = w[i]; t.bias[i] = bias ? bias[i] : nullptr; t.y[i] = y[i];
    t.J[i] = J[i]; t.K[i] = K[i];
    t.first_block[i] = blocks;
    blocks += (J[i] + rpc - 1) / rpc;
  }
  t.first_block[count] = blocks;
  t.count = count; t.B = B; t.flags = flags; t.slope = slope; t.eps = eps;
  // one layer with a long K (the 12288-wide first layer of the histogram MLP: 50 MB of weights on
  // J/16 = 64 CTAs ran at 0.36 TB/s): the K chunks are divided over gridDim.z, ordered finish
  t.ksplit = 1; t.partial = nullptr;
  if (count == 1 && K[0] >= 8 * kKC) {
    int ks = 1;
    const int nchunks = (K[0] + kKC - 1) / kKC;
    while (ks * 2 <= 16 && ks * 2 <= nchunks / 2 && (long long)blocks * ((B + 31) / 32) * ks < 4LL * sms) ks *= 2;
    ks = (nchunks + (nchunks + ks - 1) / ks - 1) / ((nchunks + ks - 1) / ks);      // no empty split
    if (ks > 1) {
      float* ws = style_workspace(sizeof(float) * (size_t)ks * B * J[0], (cudaStream_t)stream_);
      if (ws) { t.ksplit = ks; t.partial = ws; }
    }
  }
  const size_t smem = sizeof(float) * (32 * kXS + kRowsPerCta * 32);
  static PerDeviceOnce once;
  if (once.need()) {
    HG_CUDA_OK(cudaFuncSetAttribute(grouped_linear_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    once.mark();
  }
  grouped_linear_fwd_kernel<<<dim3(blocks, (B + 31) / 32, t.ksplit), kStyleThreads, smem, (cudaStream_t)stream_>>>(t);
  HG_LAUNCH_OK("grouped_linear_fwd_kernel");
  if (t.ksplit > 1) {
    const long long n = (long long)B * J[0];
    grouped_linear_fwd_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(t);
    HG_LAUNCH_OK("grouped_linear_fwd_finish_kernel");
  }
  return 0;
}

extern "C" int hg_grouped_linear_bwd(int32_t count, const float* const* x, const float* const* w,
                                     const float* const* gy, float* const* gw, float* const* gb,
                                     float* const* gx, const int32_t* J, const int32_t* K, int32_t B,
                                     int32_t flags, hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_groups(count, B, J, K);
  if (rc) return rc;
  if (count == 0 || B == 0) return 0;
  if (B > 32) return set_error(HG_ENOSUP, "grouped_linear_bwd: batch %d > 32", B);
  LinearBwdGroups t{};
  bool any_w = false, any_x = false;
  for (int i = 0; i < count; ++i) {
    if (!x[i] || !w[i] || !gy[i]) return set_error(HG_EINVAL, "null pointer in group %d", i);
    t.x[i] = x[i]; t.w[i] = w[i]; t.gy[i] = gy[i];
    t.gw[i] = gw ? gw[i] : nullptr; t.gb[i] = gb ? gb[i] : nullptr; t.gx[i] = gx ? gx[i] : nullptr;
    t.J[i] = J[i]; t.K[i] = K[i];
    any_w |= t.gw[i] || t.gb[i];
    any_x |= t.gx[i] != nullptr;
  }
  t.count = count; t.B = B; t.flags = flags;
  if (any_w) {
    int blocks = 0;
    for (int i = 0; i < count; ++i) {
      t.first_block[i] = blocks;
      blocks += ((J[i] + 7) / 8) * ((K[i] + kKC - 1) / kKC);
    }
    t.first_block[count] = blocks;
    grouped_linear_wgrad_kernel<<<blocks, kStyleThreads, 0, stream>>>(t);
    HG_LAUNCH_OK("grouped_linear_wgrad_kernel");
  }
  if (any_x) {
    int blocks = 0, maxJ = 0;
    for (int i = 0; i < count; ++i) {
      t.first_block[i] = blocks;
      blocks += (K[i] + 31) / 32;
      if (J[i] > maxJ) maxJ = J[i];
    }
    t.first_block[count] = blocks;
    // few K slabs (a single layer's demodulation adjoint: <= 64) -> divide the J rows over several
    // CTAs per slab, sums through a per-device scratch (fixed address: capturable), ordered finish
    const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
    int jsplit = 1;
    while (jsplit < 16 && blocks * jsplit < 2 * sms && maxJ / (jsplit * 2) >= 64) jsplit *= 2;
    t.jsplit = 1; t.partial = nullptr;
    if (jsplit > 1) {
      const size_t need = sizeof(float) * 1024 * (size_t)blocks * jsplit;
      float* ws = style_workspace(need, stream);
      if (ws) { t.jsplit = jsplit; t.partial = ws; }
    }
    const size_t smem = sizeof(float) * (32 * (256 + 1) + 8 * 32 * 32);
    static PerDeviceOnce once;
    if (once.need()) {
      HG_CUDA_OK(cudaFuncSetAttribute(grouped_linear_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      once.mark();
    }
    grouped_linear_dgrad_kernel<<<dim3(blocks, t.jsplit), kStyleThreads, smem, stream>>>(t);
    HG_LAUNCH_OK("grouped_linear_dgrad_kernel");
    if (t.jsplit > 1) {
      grouped_linear_dgrad_finish_kernel<<<blocks, 1024, 0, stream>>>(t, blocks);
      HG_LAUNCH_OK("grouped_linear_dgrad_finish_kernel");
    }
  }
  return 0;
}

extern "C" int hg_weight_sqsum(const float* w, float* wsq, int32_t Cout, int32_t T, int32_t Cin,
                               hg_stream_t stream_) {
  if (!w || !wsq) return set_error(HG_EINVAL, "null tensor pointer");
  if (Cin % 4) return set_error(HG_ENOSUP, "Cin=%d must be a multiple of 4", Cin);
  const long long n4 = (long long)Cout * (Cin / 4);
  if (n4 <= 0) return 0;
  weight_sqsum_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(w, wsq, T, Cin, n4);
  HG_LAUNCH_OK("weight_sqsum_kernel");
  return 0;
}

extern "C" int hg_demod_bwd(const float* gd, const float* d, const float* mod, const float* wsq,
                            const float* w, float* gmod_accum, float* dw_accum, float* t_ws, int32_t B,
                            int32_t Cout, int32_t T, int32_t Cin, hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!gd || !d || !mod || !wsq || !t_ws) return set_error(HG_EINVAL, "null tensor pointer");
  if (Cin % 4) return set_error(HG_ENOSUP, "Cin=%d must be a multiple of 4", Cin);
  if (B <= 0) return 0;
  if (B > 32) return set_error(HG_ENOSUP, "demod_bwd: batch %d > 32", B);
  if (gmod_accum) {       // gmod[b][ci] += 2 mod[b][ci] sum_co t[b][co] Wsq[co][ci]
    demod_t_kernel<<<(B * Cout + 255) / 256, 256, 0, stream>>>(gd, d, t_ws, B * Cout);
    HG_LAUNCH_OK("demod_t_kernel");
    const float* xs[1] = {mod};  const float* ws[1] = {wsq};  const float* gys[1] = {t_ws};
    float* gxs[1] = {gmod_accum};
    const int32_t Js[1] = {Cout}, Ks[1] = {Cin};
    int rc = hg_grouped_linear_bwd(1, xs, ws, gys, nullptr, nullptr, gxs, Js, Ks, B,
                                   HG_LIN_POST_2X | HG_LIN_ACCUMULATE, stream_);
    if (rc) return rc;
  }
  if (dw_accum) {
    if (!w) return set_error(HG_EINVAL, "dw_accum without w");
    dim3 grid((Cin / 4 + 63) / 64, (Cout + 3) / 4);
    demod_weight_grad_kernel<<<grid, 256, 0, stream>>>(dw_accum, w, gd, d, mod, B, Cout, T, Cin);
    HG_LAUNCH_OK("demod_weight_grad_kernel");
  }
  return 0;
}
