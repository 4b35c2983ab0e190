// hist_fwd.cu -- RGB-uv histogram forward (replaces RGBuvHistBlock.py:75-228).
//
// Data flow (all float32, one pass over the image):
//   x (B,C,H,W, strided) --load_pixel: clamp/resize--> registers
//     --project_pixel--> Iy, u_RG, u_RB, u_GB          (3 logs + 1 sqrt / pixel)
//     --soft binning--> 4 shared-memory operand rows of 64 values per pixel
//     --rank-1 updates--> 3x64x64 accumulators in registers (per CTA)
//     --> partial[b][chunk][3][64][64] --reduce--> raw[b] --normalise--> hist[b]
//
// Fast path (h == 64, lo == -hi, 3 channels, RBF / inverse-quadratic): only
// three distinct kernel matrices exist per pixel:
//   E_RG = k(L_R-L_G-c), E_RB = k(L_R-L_B-c), E_GB = k(L_G-L_B-c)
//   hist0 = (Iy E_RG)^T E_RB,  hist1 = flipud((Iy E_RG)^T E_GB),
//   hist2 = flipud(fliplr((Iy E_RB)^T E_GB))
// because the reference's u/v of channels 1 and 2 are exact negations and the
// bin centres are symmetric.  Everything else goes through the generic kernel,
// which evaluates the soft-binning kernel in float64 exactly as the reference.
#include "hg_common.cuh"

namespace hg {

// ============================================================ fast path =====
constexpr int kFP = 32;            // pixels per tile
constexpr int kFThreads = 128;
constexpr int kFBlocksPerSM = 3;

template <int METHOD, bool INTENSITY>
__global__ void __launch_bounds__(kFThreads, kFBlocksPerSM)
hist_fwd_fast_kernel(const float* __restrict__ x, const HistGeom g, const HistTables t,
                     float* __restrict__ partial, const int chunks) {
  __shared__ __align__(16) float sA_rg[kFP][64];
  __shared__ __align__(16) float sA_rb[kFP][64];
  __shared__ __align__(16) float sE_rb[kFP][64];
  __shared__ __align__(16) float sE_gb[kFP][64];
  __shared__ float sU[3][kFP];
  __shared__ float sW[kFP];

  const int tid = threadIdx.x;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tiles_total = (g.N + kFP - 1) / kFP;
  const int tile0 = (int)((long long)chunk * tiles_total / chunks);
  const int tile1 = (int)((long long)(chunk + 1) * tiles_total / chunks);

  // phase-2 role: 8 rows x 4 cols of each of the three 64x64 products
  const int ti = tid >> 4, tj = tid & 15;
  float acc0[8][4], acc1[8][4], acc2[8][4];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc0[r][c] = acc1[r][c] = acc2[r][c] = 0.f;

  // phase-1 role: bins 4*iq .. 4*iq+3 of pixel rows grp, grp+8, ...
  const int iq = tid & 15, grp = tid >> 4;
  float chi[4], clo[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { chi[k] = t.c_hi[iq * 4 + k]; clo[k] = t.c_lo[iq * 4 + k]; }
  const float inv_s2 = g.inv_sigma2;

  for (int tile = tile0; tile < tile1; ++tile) {
    const int p0 = tile * kFP;
    if (tid < kFP) {
      const int p = p0 + tid;
      float w = 0.f, u0 = 0.f, u1 = 0.f, u2 = 0.f;
      if (p < g.N) {
        float r, gg, bb;
        load_pixel(x, g, t, b, p, r, gg, bb);
        const PixelProj q = project_pixel(r, gg, bb, INTENSITY);
        w = q.iy; u0 = q.u_rg; u1 = q.u_rb; u2 = q.u_gb;
      }
      sW[tid] = w; sU[0][tid] = u0; sU[1][tid] = u1; sU[2][tid] = u2;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kFP / 8; ++it) {
      const int p = grp + it * 8;
      const float w = sW[p];
      float4 k;
      // E_RG -> A_RG = Iy * E_RG                       (RGBuvHistBlock.py:147)
      float u = sU[0][p];
      k.x = kernel_f32<METHOD>(u, chi[0], clo[0], inv_s2);
      k.y = kernel_f32<METHOD>(u, chi[1], clo[1], inv_s2);
      k.z = kernel_f32<METHOD>(u, chi[2], clo[2], inv_s2);
      k.w = kernel_f32<METHOD>(u, chi[3], clo[3], inv_s2);
      *reinterpret_cast<float4*>(&sA_rg[p][iq * 4]) =
          make_float4(__fmul_rn(w, k.x), __fmul_rn(w, k.y), __fmul_rn(w, k.z), __fmul_rn(w, k.w));
      // E_RB (plain, and Iy-weighted for hist2)
      u = sU[1][p];
      k.x = kernel_f32<METHOD>(u, chi[0], clo[0], inv_s2);
      k.y = kernel_f32<METHOD>(u, chi[1], clo[1], inv_s2);
      k.z = kernel_f32<METHOD>(u, chi[2], clo[2], inv_s2);
      k.w = kernel_f32<METHOD>(u, chi[3], clo[3], inv_s2);
      *reinterpret_cast<float4*>(&sE_rb[p][iq * 4]) = k;
      *reinterpret_cast<float4*>(&sA_rb[p][iq * 4]) =
          make_float4(__fmul_rn(w, k.x), __fmul_rn(w, k.y), __fmul_rn(w, k.z), __fmul_rn(w, k.w));
      // E_GB
      u = sU[2][p];
      k.x = kernel_f32<METHOD>(u, chi[0], clo[0], inv_s2);
      k.y = kernel_f32<METHOD>(u, chi[1], clo[1], inv_s2);
      k.z = kernel_f32<METHOD>(u, chi[2], clo[2], inv_s2);
      k.w = kernel_f32<METHOD>(u, chi[3], clo[3], inv_s2);
      *reinterpret_cast<float4*>(&sE_gb[p][iq * 4]) = k;
    }
    __syncthreads();
#pragma unroll 2
    for (int p = 0; p < kFP; ++p) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sA_rg[p][ti * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sA_rg[p][32 + ti * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&sA_rb[p][ti * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sA_rb[p][32 + ti * 4]);
      const float4 e4 = *reinterpret_cast<const float4*>(&sE_rb[p][tj * 4]);
      const float4 f4 = *reinterpret_cast<const float4*>(&sE_gb[p][tj * 4]);
      const float ar[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float br[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const float ec[4] = {e4.x, e4.y, e4.z, e4.w};
      const float fc[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc0[r][c] = fmaf(ar[r], ec[c], acc0[r][c]);
          acc1[r][c] = fmaf(ar[r], fc[c], acc1[r][c]);
          acc2[r][c] = fmaf(br[r], fc[c], acc2[r][c]);
        }
    }
    __syncthreads();
  }

  // epilogue: un-flip and store this CTA's partial histogram
  float* out = partial + ((long long)b * chunks + chunk) * (3 * 64 * 64);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = (r < 4) ? (ti * 4 + r) : (32 + ti * 4 + (r - 4));
    const int j0 = tj * 4;
    *reinterpret_cast<float4*>(out + i * 64 + j0) =
        make_float4(acc0[r][0], acc0[r][1], acc0[r][2], acc0[r][3]);
    *reinterpret_cast<float4*>(out + 4096 + (63 - i) * 64 + j0) =
        make_float4(acc1[r][0], acc1[r][1], acc1[r][2], acc1[r][3]);
    *reinterpret_cast<float4*>(out + 8192 + (63 - i) * 64 + (60 - j0)) =
        make_float4(acc2[r][3], acc2[r][2], acc2[r][1], acc2[r][0]);
  }
}

// ========================================================= generic path =====
// One CTA = one (image, channel, pixel chunk).  256 threads as a 16x16 grid;
// thread (ti,tj) owns bins i = ti+16a, j = tj+16b, a,b < NA.
constexpr int kGP = 16;
constexpr int kGThreads = 256;

// (u,v) of output channel c in terms of the three logs (RGBuvHistBlock.py:112-115,
// 150-153, 190-193)
__device__ __forceinline__ void channel_uv(int c, float lr, float lg, float lb, float& u,
                                           float& v) {
  if (c == 0) { u = __fadd_rn(lr, -lg); v = __fadd_rn(lr, -lb); }
  else if (c == 1) { u = __fadd_rn(lg, -lr); v = __fadd_rn(lg, -lb); }
  else { u = __fadd_rn(lb, -lr); v = __fadd_rn(lb, -lg); }
}

template <int NA>
__global__ void __launch_bounds__(kGThreads)
hist_fwd_generic_kernel(const float* __restrict__ x, const HistGeom g, const HistTables t,
                        float* __restrict__ partial, const int chunks) {
  __shared__ float sA[kGP][kMaxBins];
  __shared__ float sK[kGP][kMaxBins];
  __shared__ double sC[kMaxBins];
  __shared__ float sU[kGP], sV[kGP], sW[kGP];

  const int tid = threadIdx.x;
  const int b = blockIdx.y / g.nc, cc = blockIdx.y % g.nc;
  const int ch = g.green_only ? 1 : cc;
  const int chunk = blockIdx.x;
  const int h = g.h;
  const int tiles_total = (g.N + kGP - 1) / kGP;
  const int tile0 = (int)((long long)chunk * tiles_total / chunks);
  const int tile1 = (int)((long long)(chunk + 1) * tiles_total / chunks);
  const int ti = tid >> 4, tj = tid & 15;

  for (int i = tid; i < kMaxBins; i += kGThreads) sC[i] = t.c[i];
  for (int i = tid; i < kGP * kMaxBins; i += kGThreads) {
    (&sA[0][0])[i] = 0.f;
    (&sK[0][0])[i] = 0.f;
  }
  float acc[NA][NA];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int c = 0; c < NA; ++c) acc[a][c] = 0.f;
  __syncthreads();

  for (int tile = tile0; tile < tile1; ++tile) {
    const int p0 = tile * kGP;
    if (tid < kGP) {
      const int p = p0 + tid;
      float w = 0.f, u = 0.f, v = 0.f;
      if (p < g.N) {
        float r, gg, bb;
        load_pixel(x, g, t, b, p, r, gg, bb);
        if (g.projection == HG_PROJ_RG_CHROMA) {   // rgChromaHistBlock.py:104-112
          const float ssum = __fadd_rn(__fadd_rn(__fadd_rn(r, gg), bb), kEps);
          u = __fdiv_rn(r, ssum);
          v = __fdiv_rn(gg, ssum);
          w = g.intensity ? __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r, r), __fmul_rn(gg, gg)),
                                                           __fmul_rn(bb, bb)), kEps)) : 1.f;
        } else if (g.projection == HG_PROJ_LAB) {    // LabHistBlock.py:104-111 (input is Lab in [0,1])
          u = gg; v = bb; w = g.intensity ? r : 1.f;
        } else {
          const PixelProj q = project_pixel(r, gg, bb, g.intensity != 0);
          const float lr = log_f32(__fadd_rn(r, kEps)), lg = log_f32(__fadd_rn(gg, kEps)),
                      lb = log_f32(__fadd_rn(bb, kEps));
          channel_uv(ch, lr, lg, lb, u, v);
          w = q.iy;
        }
      }
      sU[tid] = u; sV[tid] = v; sW[tid] = w;
    }
    __syncthreads();
    for (int idx = tid; idx < kGP * h; idx += kGThreads) {
      const int p = idx / h, i = idx - p * h;
      const float ku = kernel_f64(sU[p], sC[i], g.method, g.sigma2, g.thr_half);
      const float kv = kernel_f64(sV[p], sC[i], g.method, g.sigma2, g.thr_half);
      sA[p][i] = __fmul_rn(sW[p], ku);
      sK[p][i] = kv;
    }
    __syncthreads();
    for (int p = 0; p < kGP; ++p) {
      float av[NA], kv[NA];
#pragma unroll
      for (int a = 0; a < NA; ++a) { av[a] = sA[p][ti + 16 * a]; kv[a] = sK[p][tj + 16 * a]; }
#pragma unroll
      for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int c = 0; c < NA; ++c) acc[a][c] = fmaf(av[a], kv[c], acc[a][c]);
    }
    __syncthreads();
  }

  float* out = partial + (((long long)b * chunks + chunk) * g.nc + cc) * (long long)(h * h);
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int c = 0; c < NA; ++c) {
      const int i = ti + 16 * a, j = tj + 16 * c;
      if (i < h && j < h) out[i * h + j] = acc[a][c];
    }
}

// ===================================================== reduce + normalise ===
constexpr int kRElems = 1024;   // elements per reduce CTA (256 threads x 4)

__global__ void __launch_bounds__(256)
hist_reduce_kernel(const float* __restrict__ partial, const int chunks, const int E,
                   float* __restrict__ raw, float* __restrict__ blocksums) {
  __shared__ float red[8];
  const int b = blockIdx.y;
  const float* pb = partial + (long long)b * chunks * E;
  float local = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = blockIdx.x * kRElems + k * 256 + threadIdx.x;
    if (e < E) {
      float s = 0.f;
      for (int c = 0; c < chunks; ++c) s += pb[(long long)c * E + e];   // fixed order
      raw[(long long)b * E + e] = s;
      local += s;
    }
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
    blocksums[b * gridDim.x + blockIdx.x] = s;
  }
}

// hist = raw / (sum + EPS)                                (RGBuvHistBlock.py:225-226)
__global__ void __launch_bounds__(256)
hist_normalize_kernel(const float* __restrict__ raw, const float* __restrict__ blocksums,
                      const int nblk, const int E, float* __restrict__ hist,
                      float* __restrict__ hist_sum) {
  const int b = blockIdx.y;
  float S = 0.f;
  for (int k = 0; k < nblk; ++k) S += blocksums[b * nblk + k];
  if (blockIdx.x == 0 && threadIdx.x == 0) hist_sum[b] = S;
  const float den = __fadd_rn(S, kEps);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = blockIdx.x * kRElems + k * 256 + threadIdx.x;
    if (e < E) hist[(long long)b * E + e] = __fdiv_rn(raw[(long long)b * E + e], den);
  }
}

// ============================================================ debug hooks ===
__global__ void hist_preprocess_kernel(const float* __restrict__ x, const HistGeom g,
                                       const HistTables t, float* __restrict__ pixels) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.N) return;
  float r, gg, bb;
  load_pixel(x, g, t, b, p, r, gg, bb);
  float* o = pixels + (long long)b * 3 * g.N;
  o[p] = r; o[g.N + p] = gg; o[2 * g.N + p] = bb;
}

__global__ void debug_logf_kernel(const float* __restrict__ in, float* __restrict__ out,
                                  long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = log_f32(in[i]);
}

// ============================================================ host side =====
bool hist_fast_path(const HistGeom& g, const hg_hist_params* p) {
  return g.h == 64 && !g.green_only && g.projection == HG_PROJ_RGB_UV && p->lo == -p->hi &&
         (g.method == HG_METHOD_RBF || g.method == HG_METHOD_INVERSE_QUADRATIC);
}

struct FwdPlan {
  bool fast;
  int chunks, E, nblk;
  size_t off_partial, off_raw, off_sums, total;
};

static FwdPlan make_fwd_plan(const HistGeom& g, const hg_hist_params* p) {
  FwdPlan pl;
  pl.fast = hist_fast_path(g, p);
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int P = pl.fast ? kFP : kGP;
  const int tiles = (g.N + P - 1) / P;
  const int slots = sms * (pl.fast ? kFBlocksPerSM : 4);
  const int units = g.B * (pl.fast ? 1 : g.nc);
  int want = units > 0 ? (2 * slots) / units : 1;      // ~2 waves of CTAs
  if (want < 1) want = 1;
  int max_chunks = tiles / 4;
  if (max_chunks < 1) max_chunks = 1;
  pl.chunks = want < max_chunks ? want : max_chunks;
  pl.E = g.nc * g.h * g.h;
  pl.nblk = (pl.E + kRElems - 1) / kRElems;
  size_t off = 0;
  pl.off_partial = off; off = align_up(off + sizeof(float) * (size_t)g.B * pl.chunks * pl.E, 256);
  pl.off_raw = off;     off = align_up(off + sizeof(float) * (size_t)g.B * pl.E, 256);
  pl.off_sums = off;    off = align_up(off + sizeof(float) * (size_t)g.B * pl.nblk, 256);
  pl.total = off;
  return pl;
}

}  // namespace hg

using namespace hg;

extern "C" size_t hg_hist_fwd_workspace_bytes(const hg_hist_params* p) {
  HistGeom g;
  if (make_hist_geom(p, &g, nullptr)) return 0;
  return make_fwd_plan(g, p).total;
}

extern "C" int hg_hist_fwd(const float* x, const hg_hist_params* p, float* hist, float* hist_sum,
                           void* ws, size_t ws_bytes, hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  HistGeom g;
  HistTables t;
  int rc = make_hist_geom(p, &g, &t);
  if (rc) return rc;
  if (g.B == 0) return 0;
  if (!x || !hist || !hist_sum) return set_error(HG_EINVAL, "null tensor pointer");
  const FwdPlan pl = make_fwd_plan(g, p);
  if (!ws || ws_bytes < pl.total)
    return set_error(HG_EWS, "workspace too small: %zu < %zu", ws_bytes, pl.total);
  char* w = (char*)ws;
  float* partial = (float*)(w + pl.off_partial);
  float* raw = (float*)(w + pl.off_raw);
  float* sums = (float*)(w + pl.off_sums);

  if (pl.fast) {
    dim3 grid(pl.chunks, g.B);
    const bool iq = g.method == HG_METHOD_INVERSE_QUADRATIC;
    if (iq && g.intensity)
      hist_fwd_fast_kernel<HG_METHOD_INVERSE_QUADRATIC, true><<<grid, kFThreads, 0, stream>>>(x, g, t, partial, pl.chunks);
    else if (iq)
      hist_fwd_fast_kernel<HG_METHOD_INVERSE_QUADRATIC, false><<<grid, kFThreads, 0, stream>>>(x, g, t, partial, pl.chunks);
    else if (g.intensity)
      hist_fwd_fast_kernel<HG_METHOD_RBF, true><<<grid, kFThreads, 0, stream>>>(x, g, t, partial, pl.chunks);
    else
      hist_fwd_fast_kernel<HG_METHOD_RBF, false><<<grid, kFThreads, 0, stream>>>(x, g, t, partial, pl.chunks);
    HG_LAUNCH_OK("hist_fwd_fast_kernel");
  } else {
    dim3 grid(pl.chunks, g.B * g.nc);
    if (g.h <= 64)
      hist_fwd_generic_kernel<4><<<grid, kGThreads, 0, stream>>>(x, g, t, partial, pl.chunks);
    else
      hist_fwd_generic_kernel<8><<<grid, kGThreads, 0, stream>>>(x, g, t, partial, pl.chunks);
    HG_LAUNCH_OK("hist_fwd_generic_kernel");
  }
  dim3 rgrid(pl.nblk, g.B);
  hist_reduce_kernel<<<rgrid, 256, 0, stream>>>(partial, pl.chunks, pl.E, raw, sums);
  HG_LAUNCH_OK("hist_reduce_kernel");
  hist_normalize_kernel<<<rgrid, 256, 0, stream>>>(raw, sums, pl.nblk, pl.E, hist, hist_sum);
  HG_LAUNCH_OK("hist_normalize_kernel");
  return 0;
}

extern "C" int hg_hist_preprocess(const float* x, const hg_hist_params* p, float* pixels,
                                  hg_stream_t stream_) {
  HistGeom g;
  HistTables t;
  int rc = make_hist_geom(p, &g, &t);
  if (rc) return rc;
  if (g.B == 0) return 0;
  dim3 grid((g.N + 255) / 256, g.B);
  hist_preprocess_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(x, g, t, pixels);
  HG_LAUNCH_OK("hist_preprocess_kernel");
  return 0;
}

extern "C" int hg_debug_logf(const float* in, float* out, int64_t n, hg_stream_t stream_) {
  if (n <= 0) return 0;
  debug_logf_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(in, out, n);
  HG_LAUNCH_OK("debug_logf_kernel");
  return 0;
}
