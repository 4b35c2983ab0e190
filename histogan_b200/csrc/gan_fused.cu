// gan_fused.cu -- memory-bound companions of the tensor-core convolutions: the
// element-wise / reduction halves of GeneratorBlock (histoGAN/histoGAN.py:461-479)
// and RGBBlock (:380-390), each fused into ONE pass over the NHWC activation.
//
//   modconv_epilogue_bwd   adjoint of the conv epilogue  y = lrelu(d*z + noise)
//                          -> dz (TF32-rounded, ready for dgrad/wgrad), d/d(d),
//                             d/d(noise weight), d/d(noise bias)
//   modulate_bwd           adjoint of xm = x * mod[b,c]: dx (in place) and d/d(mod)
//   torgb_fwd / torgb_bwd  1x1 modulated conv to 3 channels (+ previous rgb), planar
//                          NCHW output; backward gives dx and d/d(per-sample weights)
//
// Common skeleton: a CTA = 256 threads = 8 channel lanes (one float4 = 4 channels
// each -> 32 channels) x 32 pixel lanes; it walks a chunk of pixels of one image,
// keeps per-thread partial sums for its 4 channels, reduces them over the pixel
// lanes through shared memory and finishes with one atomicAdd per channel.
#include "hg_common.cuh"
#include "sm100_ptx.cuh"
#include "fused_skeleton.cuh"

namespace hg {

// ---------------------------------------------------------------------------
// y = lrelu(d[b,c]*z + nz[b,p]*nw[c] + nb[c])   (conv epilogue, conv_tc.cu)
// given dy, y:  dpre = dy * (y > 0 ? 1 : slope)
//   dz  = tf32_round(dpre * d)          gd[b,c] += sum_p dpre * z ,  z = (pre - noise)/d
//   gnw[c] += sum_{b,p} dpre * nz       gnb[c] += sum_{b,p} dpre
// grid (C/32, pixel chunks, B)
__global__ void __launch_bounds__(kFusedThreads)
modconv_epilogue_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                            const float* __restrict__ d, const float* __restrict__ noise,
                            const float* __restrict__ nw, const float* __restrict__ nb,
                            float* __restrict__ dz, float* __restrict__ gd, float* __restrict__ gnw,
                            float* __restrict__ gnb, int H, int W, int C, int noise_size,
                            float slope, int pix_per_cta) {
  __shared__ float4 red[3 * kPixLanes * 8];
  const int cl = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = blockIdx.x * 32 + cl * 4;
  const int b = blockIdx.z;
  const int HW = H * W;
  const int p0 = blockIdx.y * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  const bool cvalid = c < C;
  float4 dv = make_float4(1.f, 1.f, 1.f, 1.f), nwv = make_float4(0.f, 0.f, 0.f, 0.f), nbv = nwv;
  if (cvalid) {
    if (d) dv = *reinterpret_cast<const float4*>(d + (long long)b * C + c);
    if (noise) {
      nwv = *reinterpret_cast<const float4*>(nw + c);
      nbv = *reinterpret_cast<const float4*>(nb + c);
    }
  }
  const float inv_slope = 1.f / slope;
  float4 acc[3];
  acc[0] = acc[1] = acc[2] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cvalid) {
    for (int p = p0 + pl; p < p1; p += kPixLanes) {
      const long long off = ((long long)b * HW + p) * C + c;
      const float4 g = *reinterpret_cast<const float4*>(dy + off);
      const float4 yv = *reinterpret_cast<const float4*>(y + off);
      float nz = 0.f;
      if (noise) {
        const int oh = p / W, ow = p - oh * W;      // transposed noise image (histoGAN.py:466-467)
        nz = noise[((long long)b * noise_size + ow) * noise_size + oh];
      }
      float4 dp, o;
#define HG_EBWD(F)                                                                   \
      {                                                                              \
        const bool pos = yv.F > 0.f;                                                 \
        dp.F = pos ? g.F : g.F * slope;                                              \
        const float pre = pos ? yv.F : yv.F * inv_slope;                             \
        const float z_times_d = pre - fmaf(nz, nwv.F, nbv.F);                        \
        acc[0].F = fmaf(dp.F, z_times_d, acc[0].F);                                  \
        acc[1].F = fmaf(dp.F, nz, acc[1].F);                                         \
        acc[2].F += dp.F;                                                            \
        o.F = tf32_round(dp.F * dv.F);                                               \
      }
      HG_EBWD(x) HG_EBWD(y) HG_EBWD(z) HG_EBWD(w)
#undef HG_EBWD
      *reinterpret_cast<float4*>(dz + off) = o;
    }
  }
  reduce_pixel_lanes<3>(acc, red, cl, pl);
  if (pl == 0 && cvalid) {
    if (gd) {
      float4 v = acc[0];
      v.x /= dv.x; v.y /= dv.y; v.z /= dv.z; v.w /= dv.w;
      atomic_add4(gd + (long long)b * C + c, v);
    }
    if (gnw) { atomic_add4(gnw + c, acc[1]); atomic_add4(gnb + c, acc[2]); }
  }
}

// ---------------------------------------------------------------------------
// xm = x * mod[b,c]  ->  dx = dxm * mod (written over dxm), gmod[b,c] += sum_p dxm * x
__global__ void __launch_bounds__(kFusedThreads)
modulate_bwd_kernel(float* __restrict__ dxm, const float* __restrict__ x,
                    const float* __restrict__ mod, float* __restrict__ gmod, int HW, int C,
                    int pix_per_cta) {
  __shared__ float4 red[kPixLanes * 8];
  const int cl = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = blockIdx.x * 32 + cl * 4;
  const int b = blockIdx.z;
  const int p0 = blockIdx.y * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  const bool cvalid = c < C;
  float4 acc[1];
  acc[0] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cvalid) {
    const float4 m = *reinterpret_cast<const float4*>(mod + (long long)b * C + c);
    for (int p = p0 + pl; p < p1; p += kPixLanes) {
      const long long off = ((long long)b * HW + p) * C + c;
      float4 g = *reinterpret_cast<const float4*>(dxm + off);
      const float4 xv = *reinterpret_cast<const float4*>(x + off);
      acc[0].x = fmaf(g.x, xv.x, acc[0].x); acc[0].y = fmaf(g.y, xv.y, acc[0].y);
      acc[0].z = fmaf(g.z, xv.z, acc[0].z); acc[0].w = fmaf(g.w, xv.w, acc[0].w);
      g.x *= m.x; g.y *= m.y; g.z *= m.z; g.w *= m.w;
      *reinterpret_cast<float4*>(dxm + off) = g;
    }
  }
  reduce_pixel_lanes<1>(acc, red, cl, pl);
  if (pl == 0 && cvalid) atomic_add4(gmod + (long long)b * C + c, acc[0]);
}

// ---------------------------------------------------------------------------
// rgb[b,o,p] = sum_c x[b,p,c] * wmod[b,o,c] (+ prev[b,o,p]);  x NHWC, rgb/prev planar NCHW.
// G consecutive threads (a power of two >= 4, 4 channels each) share one pixel; a tile = U x (256 / G)
// consecutive pixels = ONE contiguous piece of x (16 KB at U = 4).  A CTA walks gridDim.x-strided tiles of
// one image; the tiles arrive through a 3-stage ring of 1-D bulk copies (cp.async.bulk + mbarrier, issued
// by one thread): the bytes in flight no longer depend on registers or on where the CTA is in its
// load -> shuffle -> store chain.  History (ncu --set full, profiles/r02_torgb_fwd_ncu.md): 7 dependent
// rounds per CTA with register loads ran at 1.9 TB/s; one tile per CTA, everything issued up front, at
// 2.2 TB/s (filter load, sync, pixel loads, store: three serial latencies per CTA lifetime, loads in
// flight for a third of it; 43 % issue slots, long-scoreboard stalls).
// G > 32 (C >= 256, the low-resolution levels): the warps of a pixel meet in shared memory.
constexpr int kRgbStages = 3;
template <int U>
__global__ void __launch_bounds__(kFusedThreads)
torgb_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wmod,
                 const float* __restrict__ prev, float* __restrict__ rgb, int HW, int C, int G, int tiles) {
  extern __shared__ __align__(128) float smem[];    // ring [stages][tile_pix * C], then sw [3][C]
  __shared__ float red[8][U][3];
  __shared__ __align__(8) uint64_t full[kRgbStages];
  const int b = blockIdx.y;
  const int sub = threadIdx.x % G, pl = threadIdx.x / G, npl = kFusedThreads / G;
  const int tile_pix = U * npl;
  const int tile_floats = tile_pix * C;
  float* sw = smem + kRgbStages * tile_floats;
  const float* xb = x + (long long)b * HW * C;
  auto fetch = [&](int it) {                        // thread 0: tile of iteration `it` into its stage
    const int tile = blockIdx.x + it * gridDim.x;
    if (tile >= tiles) return;
    const int p0 = tile * tile_pix;
    const uint32_t bytes = (uint32_t)(min(HW - p0, tile_pix) * C) * 4u;
    const int st = it % kRgbStages;
    ptx::mbar_expect_tx(&full[st], bytes);
    ptx::bulk_load_1d(smem + st * tile_floats, xb + (long long)p0 * C, bytes, &full[st]);
  };
  if (threadIdx.x == 0) {
    for (int st = 0; st < kRgbStages; ++st) ptx::mbar_init(&full[st], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int it = 0; it < kRgbStages; ++it) fetch(it);
  for (int i = threadIdx.x; i < 3 * C; i += kFusedThreads) sw[i] = wmod[(long long)b * 3 * C + i];
  __syncthreads();
  int it = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
    const int st = it % kRgbStages;
    const int p = tile * tile_pix + pl;
    float pv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pu = p + u * npl;
      pv[u] = (prev && sub < 3 && pu < HW) ? __ldg(prev + ((long long)b * 3 + sub) * HW + pu) : 0.f;
    }
    ptx::mbar_wait(&full[st], (uint32_t)((it / kRgbStages) & 1));
    const float* xt = smem + st * tile_floats;
    float a0[U], a1[U], a2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a0[u] = a1[u] = a2[u] = 0.f;
    for (int c = sub * 4; c < C; c += G * 4) {
      const float4 w0 = *reinterpret_cast<const float4*>(sw + c);
      const float4 w1 = *reinterpret_cast<const float4*>(sw + C + c);
      const float4 w2 = *reinterpret_cast<const float4*>(sw + 2 * C + c);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pu = p + u * npl;
        if (pu < HW) {                              // pixels beyond the image were not copied
          const float4 v = *reinterpret_cast<const float4*>(xt + (pl + u * npl) * C + c);
          a0[u] += v.x * w0.x + v.y * w0.y + v.z * w0.z + v.w * w0.w;
          a1[u] += v.x * w1.x + v.y * w1.y + v.z * w1.z + v.w * w1.w;
          a2[u] += v.x * w2.x + v.y * w2.y + v.z * w2.z + v.w * w2.w;
        }
      }
    }
    const int gw = G < 32 ? G : 32;                 // lanes of this pixel inside the warp
#pragma unroll
    for (int u = 0; u < U; ++u)
      for (int o = gw >> 1; o > 0; o >>= 1) {
        a0[u] += __shfl_xor_sync(0xffffffffu, a0[u], o);
        a1[u] += __shfl_xor_sync(0xffffffffu, a1[u], o);
        a2[u] += __shfl_xor_sync(0xffffffffu, a2[u], o);
      }
    if (G > 32) {                                   // block-uniform: G / 32 warps per pixel, fixed order
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      if (lane == 0)
#pragma unroll
        for (int u = 0; u < U; ++u) { red[warp][u][0] = a0[u]; red[warp][u][1] = a1[u]; red[warp][u][2] = a2[u]; }
    }
    __syncthreads();                                // every thread is done with this stage (and `red` is written)
    if (threadIdx.x == 0) fetch(it + kRgbStages);   // refill the stage just consumed
    if (G > 32) {
      const int wpp = G >> 5;
      if (sub < 3) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float s = 0.f;
          for (int w = 0; w < wpp; ++w) s += red[pl * wpp + w][u][sub];
          a0[u] = s;                                // lane `sub` now holds output channel `sub`
        }
      }
      __syncthreads();                              // `red` may be rewritten by the next tile
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) a0[u] = sub == 0 ? a0[u] : (sub == 1 ? a1[u] : a2[u]);
    }
    if (sub < 3) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pu = p + u * npl;
        if (pu < HW) rgb[((long long)b * 3 + sub) * HW + pu] = a0[u] + pv[u];
      }
    }
  }
}

// dx[b,p,c] = sum_o drgb[b,o,p] * wmod[b,o,c] ; gw[b,o,c] += sum_p drgb[b,o,p] * x[b,p,c]
// grid (C/32, pixel chunks, B)
__global__ void __launch_bounds__(kFusedThreads)
torgb_bwd_kernel(const float* __restrict__ drgb, const float* __restrict__ x,
                 const float* __restrict__ wmod, float* __restrict__ dx, float* __restrict__ gw,
                 int HW, int C, int pix_per_cta, int accumulate_dx) {
  __shared__ float4 red[3 * kPixLanes * 8];
  const int cl = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = blockIdx.x * 32 + cl * 4;
  const int b = blockIdx.z;
  const int p0 = blockIdx.y * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  const bool cvalid = c < C;
  float4 acc[3];
  acc[0] = acc[1] = acc[2] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cvalid) {
    const float* wb = wmod + (long long)b * 3 * C + c;
    const float4 w0 = *reinterpret_cast<const float4*>(wb);
    const float4 w1 = *reinterpret_cast<const float4*>(wb + C);
    const float4 w2 = *reinterpret_cast<const float4*>(wb + 2 * C);
    const float* gb = drgb + (long long)b * 3 * HW;
    for (int p = p0 + pl; p < p1; p += kPixLanes) {
      const float g0 = gb[p], g1 = gb[HW + p], g2 = gb[2 * HW + p];
      const long long off = ((long long)b * HW + p) * C + c;
      const float4 xv = *reinterpret_cast<const float4*>(x + off);
      float4 o;
      o.x = g0 * w0.x + g1 * w1.x + g2 * w2.x; o.y = g0 * w0.y + g1 * w1.y + g2 * w2.y;
      o.z = g0 * w0.z + g1 * w1.z + g2 * w2.z; o.w = g0 * w0.w + g1 * w1.w + g2 * w2.w;
      if (accumulate_dx) {
        const float4 old = *reinterpret_cast<const float4*>(dx + off);
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      *reinterpret_cast<float4*>(dx + off) = o;
      acc[0].x = fmaf(g0, xv.x, acc[0].x); acc[0].y = fmaf(g0, xv.y, acc[0].y);
      acc[0].z = fmaf(g0, xv.z, acc[0].z); acc[0].w = fmaf(g0, xv.w, acc[0].w);
      acc[1].x = fmaf(g1, xv.x, acc[1].x); acc[1].y = fmaf(g1, xv.y, acc[1].y);
      acc[1].z = fmaf(g1, xv.z, acc[1].z); acc[1].w = fmaf(g1, xv.w, acc[1].w);
      acc[2].x = fmaf(g2, xv.x, acc[2].x); acc[2].y = fmaf(g2, xv.y, acc[2].y);
      acc[2].z = fmaf(g2, xv.z, acc[2].z); acc[2].w = fmaf(g2, xv.w, acc[2].w);
    }
  }
  reduce_pixel_lanes<3>(acc, red, cl, pl);
  if (pl == 0 && cvalid) {
    float* gwb = gw + (long long)b * 3 * C + c;
    atomic_add4(gwb, acc[0]); atomic_add4(gwb + C, acc[1]); atomic_add4(gwb + 2 * C, acc[2]);
  }
}

// ---------------------------------------------------------------------------
// adjoint of y = act(conv + bias): dpre = tf32_round(dy * (y > 0 ? 1 : slope)) (y may be
// null: no activation) and gb[c] += sum_{b,p} dpre (before rounding).  grid (C/32, chunks, B)
__global__ void __launch_bounds__(kFusedThreads)
bias_act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                    float* __restrict__ dpre, float* __restrict__ gb, int HW, int C, float slope,
                    int pix_per_cta) {
  __shared__ float4 red[kPixLanes * 8];
  const int cl = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = blockIdx.x * 32 + cl * 4;
  const int b = blockIdx.z;
  const int p0 = blockIdx.y * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  const bool cvalid = c < C;
  float4 acc[1];
  acc[0] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cvalid) {
    for (int p = p0 + pl; p < p1; p += kPixLanes) {
      const long long off = ((long long)b * HW + p) * C + c;
      float4 g = *reinterpret_cast<const float4*>(dy + off);
      if (y) {
        const float4 yv = *reinterpret_cast<const float4*>(y + off);
        g.x = yv.x > 0.f ? g.x : g.x * slope; g.y = yv.y > 0.f ? g.y : g.y * slope;
        g.z = yv.z > 0.f ? g.z : g.z * slope; g.w = yv.w > 0.f ? g.w : g.w * slope;
      }
      acc[0].x += g.x; acc[0].y += g.y; acc[0].z += g.z; acc[0].w += g.w;
      g.x = tf32_round(g.x); g.y = tf32_round(g.y); g.z = tf32_round(g.z); g.w = tf32_round(g.w);
      *reinterpret_cast<float4*>(dpre + off) = g;
    }
  }
  if (gb) {
    reduce_pixel_lanes<1>(acc, red, cl, pl);
    if (pl == 0 && cvalid) atomic_add4(gb + c, acc[0]);
  }
}

// ---------------------------------------------------------------------------
// nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) (histoGAN.py:446-447,
// 462-463) fused with the style modulation of the conv that consumes it:
//   xm[b, oh, ow, c] = tf32_round( bilerp2x(x)[b, oh, ow, c] * mod[b, c] )
// taps/weights as torch: src = (o + 0.5)/2 - 0.5 clamped at 0, i1 = min(i0 + 1, n - 1).
__device__ __forceinline__ void up2_taps(int o, int n, int& i0, int& i1, float& l0, float& l1) {
  float src = fmaf(0.5f, (float)o + 0.5f, -0.5f);
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  i1 = i0 + (i0 < n - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

__device__ __forceinline__ float4 f4_axpy(float a, const float4& x, const float4& y) {
  return make_float4(fmaf(a, x.x, y.x), fmaf(a, x.y, y.y), fmaf(a, x.z, y.z), fmaf(a, x.w, y.w));
}

__device__ __forceinline__ float4 up2_sample(const float* __restrict__ xb, int H, int W, int C, int c,
                                             int oh, int ow) {
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  up2_taps(oh, H, y0, y1, ly0, ly1);
  up2_taps(ow, W, x0, x1, lx0, lx1);
  const float4 v00 = *reinterpret_cast<const float4*>(xb + ((long long)y0 * W + x0) * C + c);
  const float4 v01 = *reinterpret_cast<const float4*>(xb + ((long long)y0 * W + x1) * C + c);
  const float4 v10 = *reinterpret_cast<const float4*>(xb + ((long long)y1 * W + x0) * C + c);
  const float4 v11 = *reinterpret_cast<const float4*>(xb + ((long long)y1 * W + x1) * C + c);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 r = f4_axpy(ly0 * lx0, v00, z);
  r = f4_axpy(ly0 * lx1, v01, r);
  r = f4_axpy(ly1 * lx0, v10, r);
  r = f4_axpy(ly1 * lx1, v11, r);
  return r;
}

__device__ __forceinline__ float4 bilerp4(const float4& v00, const float4& v01, const float4& v10,
                                          const float4& v11, float ly0, float ly1, float lx0, float lx1) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 r = f4_axpy(ly0 * lx0, v00, z);
  r = f4_axpy(ly0 * lx1, v01, r);
  r = f4_axpy(ly1 * lx0, v10, r);
  r = f4_axpy(ly1 * lx1, v11, r);
  return r;
}

// 3x3 low-resolution neighbourhood of pixel (h, w) (rows / cols h-1, h, h+1 clamped): all the
// taps of its 2x2 output quad.  Output row 2h+dy uses rows (dy, dy+1) of it with exactly the
// weights up2_taps returns (at the borders the clamped duplicate row carries the zero weight).
struct Nbhd { float4 v[3][3]; };

__device__ __forceinline__ void load_nbhd(Nbhd& n, const float* __restrict__ xb, int H, int W, int C,
                                          int c, int h, int w) {
  const int ys[3] = {max(h - 1, 0), h, min(h + 1, H - 1)};
  const int xs[3] = {max(w - 1, 0), w, min(w + 1, W - 1)};
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q)
      n.v[r][q] = *reinterpret_cast<const float4*>(xb + ((long long)ys[r] * W + xs[q]) * C + c);
}

// one thread = one LOW-resolution pixel x 4 channels -> its 2x2 output quad (9 loads, 4 stores)
// grid (ceil(H*W*C/4 / 256), B)
__global__ void __launch_bounds__(256)
upsample_modulate_round_kernel(const float* __restrict__ x, const float* __restrict__ mod,
                               float* __restrict__ xm, int H, int W, int C) {
  const unsigned q4 = (unsigned)C / 4;
  const unsigned i = blockIdx.x * 256u + threadIdx.x;
  if (i >= (unsigned)H * W * q4) return;
  const int c = (int)(i % q4) * 4;
  const unsigned p = i / q4;
  const int w = (int)(p % (unsigned)W), h = (int)(p / (unsigned)W);
  const int b = blockIdx.y;
  Nbhd n;
  load_nbhd(n, x + (long long)b * H * W * C, H, W, C, c, h, w);
  const float4 m = *reinterpret_cast<const float4*>(mod + (long long)b * C + c);
  float ly[2][2], lx[2][2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    int i0, i1;
    up2_taps(2 * h + d, H, i0, i1, ly[d][0], ly[d][1]);
    up2_taps(2 * w + d, W, i0, i1, lx[d][0], lx[d][1]);
  }
  float* ob = xm + ((long long)b * 2 * H * 2 * W) * C + c;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      float4 o = bilerp4(n.v[dy][dx], n.v[dy][dx + 1], n.v[dy + 1][dx], n.v[dy + 1][dx + 1],
                         ly[dy][0], ly[dy][1], lx[dx][0], lx[dx][1]);
      o.x = tf32_round(o.x * m.x); o.y = tf32_round(o.y * m.y);
      o.z = tf32_round(o.z * m.z); o.w = tf32_round(o.w * m.w);
      *reinterpret_cast<float4*>(ob + ((long long)(2 * h + dy) * 2 * W + (2 * w + dx)) * C) = o;
    }
}

// adjoint: dx[b,h,w,c] = sum over the <=4x4 output pixels that tap (h,w) of weight * dxm * mod,
//          gmod[b,c]  += sum_{h,w} (dx / mod)[b,h,w,c] * x[b,h,w,c]
// A CTA owns a 4x8 tile of low-resolution pixels x 32 channels and stages the (2*4+2) x (2*8+2)
// high-resolution gradients it needs in shared memory once.  grid (C/32, tiles, B).
constexpr int kUpTH = 4, kUpTW = 8;
constexpr int kUpRows = 2 * kUpTH + 2, kUpCols = 2 * kUpTW + 2;

// weight with which output index o (of 2n) taps input index i
__device__ __forceinline__ float up2_weight(int o, int n, int i) {
  if (o < 0 || o >= 2 * n) return 0.f;
  int i0, i1; float l0, l1;
  up2_taps(o, n, i0, i1, l0, l1);
  return (i0 == i ? l0 : 0.f) + (i1 == i ? l1 : 0.f);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;                       // src-size 0 -> the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kFusedThreads, 4)
upsample_modulate_bwd_kernel(const float* __restrict__ dxm, const float* __restrict__ x,
                             const float* __restrict__ mod, float* __restrict__ dx,
                             float* __restrict__ gmod, int H, int W, int C, int tiles_w,
                             int n_tiles, int tiles_per_cta) {
  __shared__ float4 tiles[2][kUpRows * kUpCols * 8];      // 2 x 22.5 KB (static limit: 48 KB)
  float4* red = tiles[0];                                 // reused for the final reduction
  const int cl = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = blockIdx.x * 32 + cl * 4;
  const int b = blockIdx.z;
  const bool cvalid = c < C;
  const int HW = H * W;
  const float* gb = dxm + (long long)b * 4 * HW * C;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 acc[1];
  acc[0] = zero;
  const float4 m = cvalid ? *reinterpret_cast<const float4*>(mod + (long long)b * C + c) : zero;
  const float* xb = x + (long long)b * HW * C;

  // stage the (2*TH+2) x (2*TW+2) high-resolution gradients of tile t (zero outside the image)
  auto prefetch = [&](int t, int buf) {
    const int h0 = (t / tiles_w) * kUpTH, w0 = (t % tiles_w) * kUpTW;
    for (int e = pl; e < kUpRows * kUpCols; e += kPixLanes) {
      const int r = e / kUpCols, q = e - r * kUpCols;
      const int oh = 2 * h0 - 1 + r, ow = 2 * w0 - 1 + q;
      const bool ok = cvalid && oh >= 0 && oh < 2 * H && ow >= 0 && ow < 2 * W;
      cp_async16(&tiles[buf][e * 8 + cl], ok ? gb + ((long long)oh * 2 * W + ow) * C + c : gb, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // a CTA walks several tiles (double-buffered) and issues ONE atomic per channel at the end
  const int t_begin = blockIdx.y * tiles_per_cta;
  const int t_end = min(n_tiles, t_begin + tiles_per_cta);
  if (t_begin < t_end) prefetch(t_begin, 0);
  for (int t = t_begin; t < t_end; ++t) {
    const int buf = (t - t_begin) & 1;
    const int h0 = (t / tiles_w) * kUpTH, w0 = (t % tiles_w) * kUpTW;
    const int ty = pl / kUpTW, tx = pl - ty * kUpTW;
    const int h = h0 + ty, w = w0 + tx;
    float4 xc = zero;                       // issued before the wait so its latency overlaps
    if (cvalid && h < H && w < W)
      xc = *reinterpret_cast<const float4*>(xb + ((long long)h * W + w) * C + c);
    if (t + 1 < t_end) {
      prefetch(t + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (cvalid && h < H && w < W) {
      float wy[4], wx[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        wy[r] = up2_weight(2 * h - 1 + r, H, h);
        wx[r] = up2_weight(2 * w - 1 + r, W, w);
      }
      const float4* t0 = tiles[buf] + ((2 * ty) * kUpCols + 2 * tx) * 8 + cl;
      float4 g = zero;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) g = f4_axpy(wy[r] * wx[q], t0[(r * kUpCols + q) * 8], g);
      // d/d mod = <dxm, up(x)> = <up^T(dxm), x>: the un-modulated dx times x, no second gather
      acc[0].x = fmaf(g.x, xc.x, acc[0].x); acc[0].y = fmaf(g.y, xc.y, acc[0].y);
      acc[0].z = fmaf(g.z, xc.z, acc[0].z); acc[0].w = fmaf(g.w, xc.w, acc[0].w);
      g.x *= m.x; g.y *= m.y; g.z *= m.z; g.w *= m.w;
      *reinterpret_cast<float4*>(dx + ((long long)b * HW + (long long)h * W + w) * C + c) = g;
    }
    __syncthreads();          // everyone is done with tiles[buf] before it is refilled
  }
  reduce_pixel_lanes<1>(acc, red, cl, pl);
  if (pl == 0 && cvalid) atomic_add4(gmod + (long long)b * C + c, acc[0]);
}

// ---------------------------------------------------------------------------
// nn.Upsample(scale_factor=2, bilinear, align_corners=False) of the PLANAR rgb skip tensor
// (RGBBlock, histoGAN/histoGAN.py:377-378,388-389): `planes` = B*3 images of H x W -> 2H x 2W.
// Same taps / evaluation order as torch's upsample_bilinear2d.  HBM-bound (<= 33 MB per pass).
__global__ void __launch_bounds__(256)
upsample2x_planar_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W,
                             long long total) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int OW = 2 * W, OH = 2 * H;
  const int ow = (int)(i % OW);
  const int oh = (int)((i / OW) % OH);
  const long long pl = i / ((long long)OW * OH);
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  up2_taps(oh, H, y0, y1, ly0, ly1);
  up2_taps(ow, W, x0, x1, lx0, lx1);
  const float* xp = x + pl * H * W;
  y[i] = ly0 * (lx0 * xp[y0 * W + x0] + lx1 * xp[y0 * W + x1]) +
         ly1 * (lx0 * xp[y1 * W + x0] + lx1 * xp[y1 * W + x1]);
}

// adjoint as a gather (deterministic): low-res pixel (h, w) collects from output rows 2h-1..2h+2
__global__ void __launch_bounds__(256)
upsample2x_planar_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int H, int W,
                             long long total) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int w = (int)(i % W);
  const int h = (int)((i / W) % H);
  const long long pl = i / ((long long)W * H);
  const int OW = 2 * W, OH = 2 * H;
  const float* gp = dy + pl * OH * OW;
  float acc = 0.f;
#pragma unroll
  for (int a = -1; a <= 2; ++a) {
    const int oh = 2 * h + a;
    if (oh < 0 || oh >= OH) continue;
    int y0, y1; float ly0, ly1;
    up2_taps(oh, H, y0, y1, ly0, ly1);
    const float wy = (y0 == h ? ly0 : 0.f) + (y1 == h ? ly1 : 0.f);
    if (wy == 0.f) continue;
#pragma unroll
    for (int c = -1; c <= 2; ++c) {
      const int ow = 2 * w + c;
      if (ow < 0 || ow >= OW) continue;
      int x0, x1; float lx0, lx1;
      up2_taps(ow, W, x0, x1, lx0, lx1);
      const float wx = (x0 == w ? lx0 : 0.f) + (x1 == w ? lx1 : 0.f);
      acc = fmaf(wy * wx, gp[(long long)oh * OW + ow], acc);
    }
  }
  dx[i] = acc;
}

// ---------------------------------------------------------------------------
// Discriminator input (histoGAN/histoGAN.py:613-617): planar image (B,C,H,W), arbitrary element
// strides, C = 3 or 4 -> NHWC with the channels zero-padded to Cp (a multiple of 4; 32 for the
// tensor-core kernels' boxes), TF32-rounded.  One pass: reads 12 B and writes 4*Cp B per pixel
// (torch: zeros + layout copy + slice copy + rounding pass = 4 passes over the padded tensor).
__global__ void __launch_bounds__(256)
pad_round_nhwc_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int H, int W, int Cp,
                      long long sb, long long sc, long long sh, long long sw, long long total4) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;       // (pixel, channel quad)
  if (i >= total4) return;
  const int q = Cp / 4;
  const int cq = (int)(i % q) * 4;
  const long long p = i / q;
  const int w = (int)(p % W);
  const int h = (int)((p / W) % H);
  const long long b = p / ((long long)W * H);
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (cq < C) {
    const float* src = x + b * sb + (long long)h * sh + (long long)w * sw;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (cq + e < C) v[e] = tf32_round(src[(cq + e) * sc]);
  }
  *reinterpret_cast<float4*>(out + p * Cp + cq) = make_float4(v[0], v[1], v[2], v[3]);
}

}  // namespace hg

using namespace hg;

extern "C" int hg_modconv_epilogue_bwd(const float* dy, const float* y, const float* d,
                                       const float* noise, const float* noise_w,
                                       const float* noise_b, float* dz, float* gd, float* gnw,
                                       float* gnb, int32_t B, int32_t H, int32_t W, int32_t C,
                                       int32_t noise_size, float slope, hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!dy || !y || !dz) return set_error(HG_EINVAL, "null tensor pointer");
  if (C % 4) return set_error(HG_ENOSUP, "C=%d must be a multiple of 4", C);
  if (d && !gd) return set_error(HG_EINVAL, "gd required with d");
  if (noise && (!noise_w || !noise_b || !gnw || !gnb)) return set_error(HG_EINVAL, "noise grads required");
  if (B <= 0) return 0;
  if (gd) HG_CUDA_OK(cudaMemsetAsync(gd, 0, sizeof(float) * (size_t)B * C, stream));
  if (gnw) {
    HG_CUDA_OK(cudaMemsetAsync(gnw, 0, sizeof(float) * (size_t)C, stream));
    HG_CUDA_OK(cudaMemsetAsync(gnb, 0, sizeof(float) * (size_t)C, stream));
  }
  const int cblocks = (C + 31) / 32;
  const int per = pick_pix_per_cta(H * W, B, cblocks);
  dim3 grid(cblocks, (H * W + per - 1) / per, B);
  modconv_epilogue_bwd_kernel<<<grid, kFusedThreads, 0, stream>>>(
      dy, y, d, noise, noise_w, noise_b, dz, gd, gnw, gnb, H, W, C, noise_size, slope, per);
  HG_LAUNCH_OK("modconv_epilogue_bwd_kernel");
  return 0;
}

extern "C" int hg_modulate_bwd(float* dxm_inout, const float* x, const float* mod, float* gmod,
                               int32_t B, int32_t HW, int32_t C, hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!dxm_inout || !x || !mod || !gmod) return set_error(HG_EINVAL, "null tensor pointer");
  if (C % 4) return set_error(HG_ENOSUP, "C=%d must be a multiple of 4", C);
  if (B <= 0) return 0;
  HG_CUDA_OK(cudaMemsetAsync(gmod, 0, sizeof(float) * (size_t)B * C, stream));
  const int cblocks = (C + 31) / 32;
  const int per = pick_pix_per_cta(HW, B, cblocks);
  dim3 grid(cblocks, (HW + per - 1) / per, B);
  modulate_bwd_kernel<<<grid, kFusedThreads, 0, stream>>>(dxm_inout, x, mod, gmod, HW, C, per);
  HG_LAUNCH_OK("modulate_bwd_kernel");
  return 0;
}

extern "C" int hg_torgb_fwd(const float* x, const float* wmod, const float* prev, float* rgb,
                            int32_t B, int32_t HW, int32_t C, hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!x || !wmod || !rgb) return set_error(HG_EINVAL, "null tensor pointer");
  if (C % 4) return set_error(HG_ENOSUP, "C=%d must be a multiple of 4", C);
  if (B <= 0) return 0;
  if (C > 4096) return set_error(HG_ENOSUP, "torgb_fwd: C=%d too wide", C);
  if ((uintptr_t)x & 15) return set_error(HG_EINVAL, "torgb_fwd: x must be 16-byte aligned");
  int G = 4;                                        // >= 3: lanes 0..2 of a pixel write the three outputs
  while (G * 2 <= C / 4 && G < kFusedThreads) G <<= 1;
  const int npl = kFusedThreads / G;
  // 4 pixels per thread where that still leaves >= 2 CTAs per SM, else 1 (the low-resolution levels);
  // about 4 CTAs per SM in all, each walking its share of the image's tiles through a 3-stage ring
  const bool wide = (long long)((HW + 4 * npl - 1) / (4 * npl)) * B >= 2 * 148;
  const int U = wide ? 4 : 1;
  const int tiles = (HW + U * npl - 1) / (U * npl);
  int gx = (148 * 4 + B - 1) / B;
  if (gx > tiles) gx = tiles;
  const size_t smem = sizeof(float) * ((size_t)kRgbStages * U * npl * C + 3 * (size_t)C);
  static PerDeviceOnce once;
  if (once.need()) {
    HG_CUDA_OK(cudaFuncSetAttribute(torgb_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    HG_CUDA_OK(cudaFuncSetAttribute(torgb_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    once.mark();
  }
  if (smem > 160 * 1024) return set_error(HG_ENOSUP, "torgb_fwd: C=%d too wide", C);
  if (wide) torgb_fwd_kernel<4><<<dim3(gx, B), kFusedThreads, smem, stream>>>(x, wmod, prev, rgb, HW, C, G, tiles);
  else torgb_fwd_kernel<1><<<dim3(gx, B), kFusedThreads, smem, stream>>>(x, wmod, prev, rgb, HW, C, G, tiles);
  HG_LAUNCH_OK("torgb_fwd_kernel");
  return 0;
}

extern "C" int hg_torgb_bwd(const float* drgb, const float* x, const float* wmod, float* dx,
                            float* gw, int32_t B, int32_t HW, int32_t C, int32_t accumulate_dx,
                            hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!drgb || !x || !wmod || !dx || !gw) return set_error(HG_EINVAL, "null tensor pointer");
  if (C % 4) return set_error(HG_ENOSUP, "C=%d must be a multiple of 4", C);
  if (B <= 0) return 0;
  HG_CUDA_OK(cudaMemsetAsync(gw, 0, sizeof(float) * (size_t)B * 3 * C, stream));
  const int cblocks = (C + 31) / 32;
  const int per = pick_pix_per_cta(HW, B, cblocks);
  dim3 grid(cblocks, (HW + per - 1) / per, B);
  torgb_bwd_kernel<<<grid, kFusedThreads, 0, stream>>>(drgb, x, wmod, dx, gw, HW, C, per, accumulate_dx);
  HG_LAUNCH_OK("torgb_bwd_kernel");
  return 0;
}

extern "C" int hg_bias_act_bwd(const float* dy, const float* y, float* dpre, float* gb, int32_t B,
                               int32_t HW, int32_t C, float slope, hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!dy || !dpre) return set_error(HG_EINVAL, "null tensor pointer");
  if (C % 4) return set_error(HG_ENOSUP, "C=%d must be a multiple of 4", C);
  if (B <= 0) return 0;
  if (gb) HG_CUDA_OK(cudaMemsetAsync(gb, 0, sizeof(float) * (size_t)C, stream));
  const int cblocks = (C + 31) / 32;
  const int per = pick_pix_per_cta(HW, B, cblocks);
  dim3 grid(cblocks, (HW + per - 1) / per, B);
  bias_act_bwd_kernel<<<grid, kFusedThreads, 0, stream>>>(dy, y, dpre, gb, HW, C, slope, per);
  HG_LAUNCH_OK("bias_act_bwd_kernel");
  return 0;
}

extern "C" int hg_upsample_modulate_round(const float* x, const float* mod, float* xm, int32_t B,
                                          int32_t H, int32_t W, int32_t C, hg_stream_t stream_) {
  if (!x || !mod || !xm) return set_error(HG_EINVAL, "null tensor pointer");
  if (C % 4) return set_error(HG_ENOSUP, "C=%d must be a multiple of 4", C);
  const long long n4 = (long long)H * W * (C / 4);
  if (n4 <= 0 || B <= 0) return 0;
  if (n4 >= (1LL << 31) || B > 65535)
    return set_error(HG_ENOSUP, "upsample_modulate_round: tensor too large (%d x %d x %d, B=%d)", H, W, C, B);
  upsample_modulate_round_kernel<<<dim3((unsigned)((n4 + 255) / 256), B), 256, 0, (cudaStream_t)stream_>>>(
      x, mod, xm, H, W, C);
  HG_LAUNCH_OK("upsample_modulate_round_kernel");
  return 0;
}

extern "C" int hg_upsample_modulate_bwd(const float* dxm, const float* x, const float* mod, float* dx,
                                        float* gmod, int32_t B, int32_t H, int32_t W, int32_t C,
                                        hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!dxm || !x || !mod || !dx || !gmod) return set_error(HG_EINVAL, "null tensor pointer");
  if (C % 4) return set_error(HG_ENOSUP, "C=%d must be a multiple of 4", C);
  if (B <= 0) return 0;
  HG_CUDA_OK(cudaMemsetAsync(gmod, 0, sizeof(float) * (size_t)B * C, stream));
  const int cblocks = (C + 31) / 32;
  const int tiles_w = (W + kUpTW - 1) / kUpTW, tiles_h = (H + kUpTH - 1) / kUpTH;
  if (B > 65535) return set_error(HG_ENOSUP, "upsample_modulate_bwd: batch too large (%d)", B);
  const int n_tiles = tiles_w * tiles_h;
  long long per = ((long long)n_tiles * cblocks * B + 148 * 16 - 1) / (148 * 16);      // ~16 CTAs per SM in all
  if (per < 1) per = 1;
  if (per > n_tiles) per = n_tiles;
  dim3 grid(cblocks, (unsigned)((n_tiles + per - 1) / per), B);
  upsample_modulate_bwd_kernel<<<grid, kFusedThreads, 0, stream>>>(dxm, x, mod, dx, gmod, H, W, C,
                                                                   tiles_w, n_tiles, (int)per);
  HG_LAUNCH_OK("upsample_modulate_bwd_kernel");
  return 0;
}

extern "C" int hg_upsample2x_planar(const float* x, float* y, int32_t planes, int32_t H, int32_t W,
                                    int32_t backward, hg_stream_t stream_) {
  if (!x || !y) return set_error(HG_EINVAL, "null tensor pointer");
  if (planes <= 0 || H <= 0 || W <= 0) return 0;
  // forward: x (planes,H,W) -> y (planes,2H,2W), one thread per output; backward: x = dy
  // (planes,2H,2W) -> y = dx (planes,H,W), one thread per low-resolution pixel
  const long long total = (long long)planes * H * W * (backward ? 1 : 4);
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (backward)
    upsample2x_planar_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(x, y, H, W, total);
  else
    upsample2x_planar_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(x, y, H, W, total);
  HG_LAUNCH_OK("upsample2x_planar_kernel");
  return 0;
}

extern "C" int hg_pad_round_nhwc(const float* x, float* out, int32_t B, int32_t C, int32_t H, int32_t W,
                                 int32_t Cp, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
                                 hg_stream_t stream_) {
  if (!x || !out) return set_error(HG_EINVAL, "null tensor pointer");
  if (Cp % 4 || Cp < C) return set_error(HG_EINVAL, "pad_round: Cp=%d must be a multiple of 4 and >= C=%d", Cp, C);
  const long long total4 = (long long)B * H * W * (Cp / 4);
  if (total4 <= 0) return 0;
  pad_round_nhwc_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      x, out, C, H, W, Cp, sb, sc, sh, sw, total4);
  HG_LAUNCH_OK("pad_round_nhwc_kernel");
  return 0;
}
