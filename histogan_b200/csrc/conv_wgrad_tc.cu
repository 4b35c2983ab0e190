// conv_wgrad_tc.cu -- weight gradient of the NHWC convolution on tcgen05 (TF32).
//
//   dW[co][kh][kw][ci] = sum_{b,oh,ow} dy[b,oh,ow,co] * x[b, oh*s+kh-p, ow*s+kw-p, ci]
//
// (what autograd computes for Conv2DMod's F.conv2d, histoGAN/histoGAN.py:436, and
// for the nn.Conv2d layers of DiscriminatorBlock, :505-526).  Per filter tap this is
// a GEMM whose reduction dimension is the PIXEL index: M = Cout, N = Cin,
// K = B*OH*OW.  In NHWC both operands are "MN-major" (channels contiguous, pixels
// strided), which the UMMA shared-memory descriptors support for TF32, so the
// same TMA boxes as the forward kernel feed the tensor cores without a transpose:
//   A chunk: {32 co, PW, PH, PB} of dy        -> 64 pixel rows x 128 B (swizzle 128B,
//            32B atoms: the layout tcgen05 requires for MN-major 32-bit operands)
//   B chunk: {32 ci, PW*s, PH*s, PB} of x at the tap's shifted origin (zero-filled
//            outside the image == padding), element stride s for stride-2 convs.
// One CTA owns one (tap, 128-co tile, BN-ci tile) accumulator in TMEM and a slice
// of the pixel range (split-K); partial results are combined with fp32 atomics.
#include <cstdlib>
#include <mutex>
#include "hg_common.cuh"
#include "sm100_ptx.cuh"

namespace hg {

constexpr int kWgM = 128;               // co per CTA
constexpr int kWgPix = 64;              // pixels per k-block
constexpr int kWgChunkBytes = kWgPix * 128;   // 8 KB: 64 rows x 32 fp32
constexpr int kWgThreads = 192;

struct WgradArgs {
  int B, H, W, Cin, Cout, KH, KW, stride, pad, OH, OW;
  int PB, PH, PW;
  int tiles_w, tiles_h, tiles_b;
  int co_tiles, ci_tiles, splits;
  int atomic;                           // splits > 1 and no scratch: fp32 atomics into a zeroed dw (fallback)
  float* dw;                            // packed [Cout][KH*KW][Kp], Kp = round_up(Cin, 32)
  int Kp;
  float* partial;                       // splits > 1: split s writes its partial dW to partial + s*slice;
  long long slice;                      //   wgrad_finish_kernel adds the slices in index order (deterministic)
};

template <int BN, int STAGES, int TAPS = 1>
struct WgradSmem {
  static constexpr int kAChunks = kWgM / 32;
  static constexpr int kBChunks = (BN / 32) * TAPS;
  static constexpr int kStageBytes = (kAChunks + kBChunks) * kWgChunkBytes;
  static constexpr int kTotal = STAGES * kStageBytes + 1024 + 256;
};

// TAPS == 1: one CTA per (filter tap, co tile, ci tile).
// TAPS == 9 (3x3 filters with <= 32 input channels -- the 256^2 / 128^2 layers, where the
//   pixel reduction is longest): one CTA keeps all nine tap accumulators (9 x 32 TMEM columns),
//   so a dy chunk is fetched once per pixel block instead of once per tap and each pixel block
//   feeds 72 MMAs instead of 8.
template <int BN, int STAGES, int TAPS = 1>
__global__ void __launch_bounds__(kWgThreads)
conv_wgrad_tf32_kernel(const __grid_constant__ CUtensorMap tmdy,
                       const __grid_constant__ CUtensorMap tmx, const WgradArgs a) {
  using SM = WgradSmem<BN, STAGES, TAPS>;
  static_assert(TAPS == 1 || (TAPS == 9 && BN == 32), "all-taps variant: 3x3, BN = 32");
  constexpr uint32_t kTmemCols = TAPS == 9 ? 512 : (BN < 32 ? 32 : BN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + STAGES * SM::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmdy);
    ptx::prefetch_tmap(&tmx);
  }
  if (warp == 1) {
    if (ptx::elect_one()) {
      for (int s = 0; s < STAGES; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      ptx::mbar_init(tmem_full_bar, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_ptr, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  int id = blockIdx.x;
  const int ci_t = id % a.ci_tiles; id /= a.ci_tiles;
  const int co_t = id % a.co_tiles; id /= a.co_tiles;
  const int tap = id;                        // 0 when TAPS == 9
  const int kh = tap / a.KW, kw = tap - kh * a.KW;
  const int co0 = co_t * kWgM, ci0 = ci_t * BN;
  const int kb_total = a.tiles_w * a.tiles_h * a.tiles_b;
  const int kb0 = (int)((long long)blockIdx.y * kb_total / a.splits);
  const int kb1 = (int)((long long)(blockIdx.y + 1) * kb_total / a.splits);

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int a_chunks = min(SM::kAChunks, (a.Cout - co0 + 31) / 32);
      for (int kb = kb0; kb < kb1; ++kb) {
        const int tw_i = kb % a.tiles_w;
        const int th_i = (kb / a.tiles_w) % a.tiles_h;
        const int tb_i = kb / (a.tiles_w * a.tiles_h);
        const int ow0 = tw_i * a.PW, oh0 = th_i * a.PH, b0 = tb_i * a.PB;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sA = base + stage * SM::kStageBytes;
        uint8_t* sB = sA + SM::kAChunks * kWgChunkBytes;
        // only the 32-channel chunks of dy that exist are fetched; the MMA still runs M = 128 and
        // the accumulator rows of the missing channels hold garbage that the epilogue never reads
        ptx::mbar_expect_tx(&full_bar[stage], (a_chunks + SM::kBChunks) * kWgChunkBytes);
#pragma unroll
        for (int c = 0; c < SM::kAChunks; ++c)
          if (c < a_chunks)
            ptx::tma_load_4d(sA + c * kWgChunkBytes, &tmdy, &full_bar[stage], co0 + 32 * c, ow0, oh0, b0);
        if (TAPS == 1) {
#pragma unroll
          for (int c = 0; c < SM::kBChunks; ++c)
            ptx::tma_load_4d(sB + c * kWgChunkBytes, &tmx, &full_bar[stage], ci0 + 32 * c,
                             ow0 * a.stride + kw - a.pad, oh0 * a.stride + kh - a.pad, b0);
        } else {
#pragma unroll
          for (int t = 0; t < TAPS; ++t)       // chunk t = the x tile shifted by tap t
            ptx::tma_load_4d(sB + t * kWgChunkBytes, &tmx, &full_bar[stage], ci0,
                             ow0 * a.stride + (t % 3) - a.pad, oh0 * a.stride + (t / 3) - a.pad, b0);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc(2 /*tf32*/, kWgM, BN, 1 /*A MN-major*/, 1 /*B MN-major*/);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sA = ptx::smem_u32(base + stage * SM::kStageBytes);
        const uint32_t sB = sA + SM::kAChunks * kWgChunkBytes;
        // MN-major tf32: 128B swizzle with 32B atoms (4 pixel rows x 128 B per atom):
        // LBO = stride between 32-channel chunks, SBO = stride between 4-row groups
        const uint64_t a_desc = ptx::make_smem_desc(sA, kWgChunkBytes, 512, ptx::kLayoutSW128Base32);
#pragma unroll
        for (int t = 0; t < TAPS; ++t) {
          const uint64_t b_desc = ptx::make_smem_desc(sB + (TAPS == 1 ? 0 : t * kWgChunkBytes),
                                                      kWgChunkBytes, 512, ptx::kLayoutSW128Base32);
#pragma unroll
          for (int k = 0; k < kWgPix / 8; ++k) {
            // next 8 pixels = 8 rows x 128 B = two swizzle atoms: +64 in (addr >> 4) units
            ptx::mma_tf32_ss(tmem_base + (uint32_t)(t * BN), a_desc + (uint64_t)(k * 64),
                             b_desc + (uint64_t)(k * 64), idesc, (uint32_t)((kb > kb0) | (k != 0)));
          }
        }
        ptx::tc_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      ptx::tc_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    const int taps = a.KH * a.KW;
    if (kb1 > kb0) {
      ptx::mbar_wait(tmem_full_bar, 0);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < BN * TAPS; cc += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
        ptx::tmem_ld_wait();
        const int c0 = TAPS == 1 ? cc : 0;
        const int tap_o = TAPS == 1 ? tap : cc / 32;
        if (co < a.Cout) {
          float* o = (a.partial ? a.partial + (long long)blockIdx.y * a.slice : a.dw) +
                     ((long long)co * taps + tap_o) * a.Kp + ci0 + c0;
          if (a.atomic) {
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(o + j, __uint_as_float(v[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(o + j) =
                  make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                              __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------
// Column-halo variant for 3x3 / stride 1 / pad 1 layers with FEW output channels (the 256^2 ..
// 64^2 layers, where the pixel reduction is longest and M = Cout would leave the 128-row MMA
// mostly empty).  Roles are swapped:  dW^T[(kh,ci)][co] = sum_pix x[pix + tap][ci] * dy[pix][co]
//   A (M side) = the x tile, fetched ONCE per filter column kw as a {32 ci, 16, 4+2, 1} box with a
//       one-row halo above and below (96 pixel rows x 128 B = 12 KB).  Tap kh of that column is
//       the same box 16 rows = 2048 B further on, so the three vertical taps are three "32-channel
//       chunks" of one MN-major operand with LBO = 2048: one M = 128 MMA (chunk 3 reads past the
//       box; its accumulator rows are never stored) covers kh = 0..2.
//   B (N side) = the dy tile {32 co, 16, 4, 1} x NC chunks, fetched once.
// Per 64 pixels: 3 x 8 MMAs instead of 9 x 8, 3 x-boxes instead of 9, one dy fetch instead of 9.
// CTA = (co tile of NC*32, ci tile of 32, pixel split); accumulators: 3 x (NC*32) TMEM columns.
constexpr int kColBoxBytes = (4 + 2) * 16 * 128;          // 12 KB
constexpr int kColSlack = 4096;                            // chunk 3 of the last box reads past it

template <int NC, int STAGES>
struct WgradColSmem {
  static constexpr int kStageBytes = NC * kWgChunkBytes + 3 * kColBoxBytes;
  static constexpr int kTotal = STAGES * kStageBytes + kColSlack + 1024 + 256;
};

template <int NC, int STAGES>
__global__ void __launch_bounds__(kWgThreads)
conv_wgrad_col_kernel(const __grid_constant__ CUtensorMap tmdy, const __grid_constant__ CUtensorMap tmx,
                      const WgradArgs a) {
  using SM = WgradColSmem<NC, STAGES>;
  constexpr int N = NC * 32;
  constexpr uint32_t kTmemCols = NC == 1 ? 128 : (NC == 2 ? 256 : 512);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + STAGES * SM::kStageBytes + kColSlack);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmdy);
    ptx::prefetch_tmap(&tmx);
  }
  if (warp == 1) {
    if (ptx::elect_one()) {
      for (int s = 0; s < STAGES; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      ptx::mbar_init(tmem_full_bar, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_ptr, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int ci_t = blockIdx.x % a.ci_tiles, co_t = blockIdx.x / a.ci_tiles;
  const int co0 = co_t * N, ci0 = ci_t * 32;
  const int kb_total = a.tiles_w * a.tiles_h * a.tiles_b;
  const int kb0 = (int)((long long)blockIdx.y * kb_total / a.splits);
  const int kb1 = (int)((long long)(blockIdx.y + 1) * kb_total / a.splits);

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int n_chunks = min(NC, (a.Cout - co0 + 31) / 32);
      for (int kb = kb0; kb < kb1; ++kb) {
        const int tw_i = kb % a.tiles_w;
        const int th_i = (kb / a.tiles_w) % a.tiles_h;
        const int b0 = kb / (a.tiles_w * a.tiles_h);
        const int ow0 = tw_i * 16, oh0 = th_i * 4;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sDy = base + stage * SM::kStageBytes;
        uint8_t* sX = sDy + NC * kWgChunkBytes;
        ptx::mbar_expect_tx(&full_bar[stage], n_chunks * kWgChunkBytes + 3 * kColBoxBytes);
#pragma unroll
        for (int c = 0; c < NC; ++c)
          if (c < n_chunks)
            ptx::tma_load_4d(sDy + c * kWgChunkBytes, &tmdy, &full_bar[stage], co0 + 32 * c, ow0, oh0, b0);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
          ptx::tma_load_4d(sX + kw * kColBoxBytes, &tmx, &full_bar[stage], ci0, ow0 + kw - 1, oh0 - 1, b0);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc(2 /*tf32*/, 128, N, 1 /*A MN-major*/, 1 /*B MN-major*/);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sDy = ptx::smem_u32(base + stage * SM::kStageBytes);
        const uint32_t sX = sDy + NC * kWgChunkBytes;
        const uint64_t b_desc = ptx::make_smem_desc(sDy, kWgChunkBytes, 512, ptx::kLayoutSW128Base32);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          // M chunks = vertical taps: 16 pixel rows (2048 B) apart inside the halo box
          const uint64_t a_desc = ptx::make_smem_desc(sX + kw * kColBoxBytes, 2048, 512,
                                                      ptx::kLayoutSW128Base32);
#pragma unroll
          for (int k = 0; k < kWgPix / 8; ++k)
            ptx::mma_tf32_ss(tmem_base + (uint32_t)(kw * N), a_desc + (uint64_t)(k * 64),
                             b_desc + (uint64_t)(k * 64), idesc, (uint32_t)((kb > kb0) | (k != 0)));
        }
        ptx::tc_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      ptx::tc_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quadrant == vertical tap kh (3 = unused rows)
    const int ci = ci0 + lane;
    if (kb1 > kb0) {
      ptx::mbar_wait(tmem_full_bar, 0);
      ptx::tc_fence_after();
      if (q < 3) {
#pragma unroll 1
        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll 1
          for (int cc = 0; cc < N; cc += 32) {
            uint32_t v[32];
            ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kw * N + cc), v);
            ptx::tmem_ld_wait();
            float* o = (a.partial ? a.partial + (long long)blockIdx.y * a.slice : a.dw) +
                       ((long long)(co0 + cc) * 9 + q * 3 + kw) * a.Kp + ci;
            const int ncols = min(32, a.Cout - (co0 + cc));
            if (a.partial) {             // lanes = consecutive ci: coalesced 128 B rows
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncols) o[(long long)j * 9 * a.Kp] = __uint_as_float(v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncols) atomicAdd(o + (long long)j * 9 * a.Kp, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

// dW = sum over the split-K slices, in index order: bit-reproducible (fp32 atomics are not, and the
// randomly initialised networks amplify a 1e-7 reordering to percents in later gradients)
__global__ void __launch_bounds__(256)
wgrad_finish_kernel(const float* __restrict__ partial, float* __restrict__ dw, long long n4, long long slice,
                    int splits, int G) {
  // G (a power of two <= 32, fixed by the shape) consecutive lanes share one float4 of dW: lane `sub`
  // adds the slices sub, sub + G, ... and a fixed xor tree adds the G sums.  Small filters with many
  // slices (one CTA walking 100+ slices took 46 us) become wide; large ones keep G = 1.
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long i = t / G;
  const int sub = (int)(t - i * G);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n4) {
    for (int k = sub; k < splits; k += G) {
      const float4 v = *reinterpret_cast<const float4*>(partial + k * slice + i * 4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  for (int o = G >> 1; o; o >>= 1) {
    s.x += __shfl_xor_sync(0xffffffffu, s.x, o); s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
    s.z += __shfl_xor_sync(0xffffffffu, s.z, o); s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
  }
  if (i < n4 && sub == 0) *reinterpret_cast<float4*>(dw + i * 4) = s;
}

// Per-device scratch for the split-K slices: one fixed allocation made on first use (never inside a
// stream capture, never moved: captured graphs keep its address).  CONTRACT as for the forward
// split-K scratch: weight-gradient launches of one device are issued on one stream.
constexpr size_t kWgradWsBytes = 96u << 20;

static float* wgrad_workspace(size_t bytes, cudaStream_t stream) {
  static float* ws[64] = {};
  static std::mutex mu;
  if (bytes > kWgradWsBytes) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  if (ws[dev]) return ws[dev];
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return nullptr;
  float* p = nullptr;
  if (cudaMalloc(&p, kWgradWsBytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  ws[dev] = p;
  return p;
}

// splits > 1: route the partial sums through the scratch (deterministic); without scratch (request too
// large, or first use inside a capture) fall back to atomics into a zeroed dw.  Returns 0 / error.
static int setup_split(WgradArgs& a, float* dw_packed, size_t out_bytes, cudaStream_t stream) {
  a.partial = nullptr; a.slice = 0; a.atomic = 0; a.dw = dw_packed;
  if (a.splits <= 1) return 0;
  static const bool deterministic = [] {
    const char* e = getenv("HG_WGRAD_ATOMIC");
    return !(e && e[0] == '1');
  }();
  float* ws = deterministic ? wgrad_workspace(out_bytes * (size_t)a.splits, stream) : nullptr;
  if (ws) {
    a.partial = ws; a.slice = (long long)(out_bytes / sizeof(float));
  } else {
    a.atomic = 1;
    HG_CUDA_OK(cudaMemsetAsync(dw_packed, 0, out_bytes, stream));
  }
  return 0;
}

static int finish_split(const WgradArgs& a, size_t out_bytes, cudaStream_t stream) {
  if (!a.partial) return 0;
  const long long n4 = (long long)(out_bytes / 16);
  int G = 1;
  while (G < 32 && G * 2 <= a.splits && n4 * G < 148LL * 256 * 4) G <<= 1;
  wgrad_finish_kernel<<<(unsigned)((n4 * G + 255) / 256), 256, 0, stream>>>(a.partial, a.dw, n4, a.slice, a.splits, G);
  HG_LAUNCH_OK("wgrad_finish_kernel");
  return 0;
}

// packed [Cout][KH][KW][Cin] gradient -> OIHW parameter gradient (optionally +=)
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Cout,
                                    int Cin, int KH, int KW, int accumulate) {
  const long long total = (long long)Cout * Cin * KH * KW;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int kw = (int)(e % KW);
  long long r = e / KW;
  const int kh = (int)(r % KH); r /= KH;
  const int ci = (int)(r % Cin);
  const int co = (int)(r / Cin);
  const int Kp = (Cin + 31) / 32 * 32;
  const float v = dwp[(((long long)co * KH + kh) * KW + kw) * Kp + ci];
  dw[e] = accumulate ? dw[e] + v : v;
}

typedef CUresult (*PFN_encodeTiledW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                     const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_nhwc_map(CUtensorMap* m, const void* ptr, int C, int W, int H, int B, int bw, int bh,
                           int bb, int s) {
  static PFN_encodeTiledW fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_encodeTiledW>(p);
  }();
  if (!fn) return set_error(HG_EARCH, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)(bw * s), (cuuint32_t)(bh * s), (cuuint32_t)bb};
  cuuint32_t estr[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(HG_EINVAL, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return 0;
}

template <int BN, int STAGES, int TAPS = 1>
static int launch_wgrad(const CUtensorMap& tmdy, const CUtensorMap& tmx, const WgradArgs& a,
                        cudaStream_t stream) {
  using SM = WgradSmem<BN, STAGES, TAPS>;
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    HG_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_tf32_kernel<BN, STAGES, TAPS>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal));
    attr_once.mark();
  }
  dim3 grid((a.KH * a.KW / TAPS) * a.co_tiles * a.ci_tiles, a.splits);
  conv_wgrad_tf32_kernel<BN, STAGES, TAPS><<<grid, kWgThreads, SM::kTotal, stream>>>(tmdy, tmx, a);
  HG_LAUNCH_OK("conv_wgrad_tf32_kernel");
  return 0;
}

template <int NC, int STAGES>
static int launch_wgrad_col(const CUtensorMap& tmdy, const CUtensorMap& tmx, const WgradArgs& a,
                            int co_tiles, cudaStream_t stream) {
  using SM = WgradColSmem<NC, STAGES>;
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    HG_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_col_kernel<NC, STAGES>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal));
    attr_once.mark();
  }
  dim3 grid(co_tiles * a.ci_tiles, a.splits);
  conv_wgrad_col_kernel<NC, STAGES><<<grid, kWgThreads, SM::kTotal, stream>>>(tmdy, tmx, a);
  HG_LAUNCH_OK("conv_wgrad_col_kernel");
  return 0;
}

}  // namespace hg

using namespace hg;

extern "C" int hg_conv2d_wgrad(const float* dy, const float* x, float* dw_packed,
                               const hg_conv_params* p, hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!dy || !x || !dw_packed || !p) return set_error(HG_EINVAL, "null pointer");
  if (p->Cin % 4 != 0 || p->Cout % 4 != 0)
    return set_error(HG_ENOSUP, "wgrad: Cin=%d and Cout=%d must be multiples of 4", p->Cin, p->Cout);
  if (p->stride < 1 || p->stride > 2) return set_error(HG_ENOSUP, "wgrad: stride must be 1 or 2");
  const int OH = (p->H + 2 * p->pad - p->KH) / p->stride + 1;
  const int OW = (p->W + 2 * p->pad - p->KW) / p->stride + 1;
  if (OH != p->OH || OW != p->OW) return set_error(HG_EINVAL, "wgrad: inconsistent OH/OW");
  const int Kp = (p->Cin + 31) / 32 * 32;
  const size_t out_bytes = sizeof(float) * (size_t)p->Cout * p->KH * p->KW * Kp;
  if (p->B <= 0) { HG_CUDA_OK(cudaMemsetAsync(dw_packed, 0, out_bytes, stream)); return 0; }

  WgradArgs a{};
  a.B = p->B; a.H = p->H; a.W = p->W; a.Cin = p->Cin; a.Cout = p->Cout; a.KH = p->KH; a.KW = p->KW;
  a.stride = p->stride; a.pad = p->pad; a.OH = OH; a.OW = OW;
  const int Np = (p->Cout + 31) / 32 * 32;
  if (p->KH == 3 && p->KW == 3 && p->stride == 1 && p->pad == 1 && OW >= 16 && OH >= 4 && Np <= 128 && Np != 96 &&
      (long long)p->B * OH * OW >= 64 * 1024) {
    // few output channels, many pixels: column-halo variant (see conv_wgrad_col_kernel)
    a.PW = 16; a.PH = 4; a.PB = 1;
    a.tiles_w = (OW + 15) / 16; a.tiles_h = (OH + 3) / 4; a.tiles_b = p->B;
    a.ci_tiles = Kp / 32; a.Kp = Kp;
    const int NC = Np / 32;                              // 1, 2 or 4
    const int kb_total = a.tiles_w * a.tiles_h * a.tiles_b;
    const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
    int splits = (2 * sms + a.ci_tiles - 1) / a.ci_tiles;
    if (splits > kb_total / 2) splits = kb_total / 2;
    if (splits < 1) splits = 1;
    a.splits = splits;
    int rc = setup_split(a, dw_packed, out_bytes, stream);
    if (rc) return rc;
    if (splits == 1) { a.atomic = 1; HG_CUDA_OK(cudaMemsetAsync(dw_packed, 0, out_bytes, stream)); }
    alignas(64) CUtensorMap tmdy, tmx;
    rc = encode_nhwc_map(&tmdy, dy, p->Cout, OW, OH, p->B, 16, 4, 1, 1);
    if (rc) return rc;
    rc = encode_nhwc_map(&tmx, x, p->Cin, p->W, p->H, p->B, 16, 6, 1, 1);
    if (rc) return rc;
    if (NC == 1) rc = launch_wgrad_col<1, 4>(tmdy, tmx, a, 1, stream);
    else if (NC == 2) rc = launch_wgrad_col<2, 4>(tmdy, tmx, a, 1, stream);
    else rc = launch_wgrad_col<4, 3>(tmdy, tmx, a, 1, stream);
    return rc ? rc : finish_split(a, out_bytes, stream);
  }
  int PW = 1; while (PW < 16 && PW < OW) PW <<= 1;
  int PH = 1; while (PW * PH < kWgPix && PH < OH) PH <<= 1;
  const int PB = kWgPix / (PW * PH);
  a.PW = PW; a.PH = PH; a.PB = PB;
  a.tiles_w = (OW + PW - 1) / PW; a.tiles_h = (OH + PH - 1) / PH; a.tiles_b = (p->B + PB - 1) / PB;
  const int BN = (Kp % 128 == 0) ? 128 : (Kp % 64 == 0 ? 64 : 32);
  a.co_tiles = (p->Cout + kWgM - 1) / kWgM;
  a.ci_tiles = Kp / BN;
  a.Kp = Kp;
  const int kb_total = a.tiles_w * a.tiles_h * a.tiles_b;
  const bool all_taps = p->KH == 3 && p->KW == 3 && Kp == 32 && kb_total >= 64;
  const int base_ctas = (all_taps ? 1 : p->KH * p->KW) * a.co_tiles * a.ci_tiles;
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  int splits = (2 * sms + base_ctas - 1) / base_ctas;
  if (splits > kb_total / 2) splits = kb_total / 2;
  if (splits < 1) splits = 1;
  a.splits = splits;
  int rc = setup_split(a, dw_packed, out_bytes, stream);
  if (rc) return rc;

  alignas(64) CUtensorMap tmdy, tmx;
  rc = encode_nhwc_map(&tmdy, dy, p->Cout, OW, OH, p->B, PW, PH, PB, 1);
  if (rc) return rc;
  rc = encode_nhwc_map(&tmx, x, p->Cin, p->W, p->H, p->B, PW, PH, PB, p->stride);
  if (rc) return rc;
  if (all_taps) rc = launch_wgrad<32, 2, 9>(tmdy, tmx, a, stream);
  else if (BN == 128) rc = launch_wgrad<128, 3>(tmdy, tmx, a, stream);
  else if (BN == 64) rc = launch_wgrad<64, 4>(tmdy, tmx, a, stream);
  else rc = launch_wgrad<32, 4>(tmdy, tmx, a, stream);
  return rc ? rc : finish_split(a, out_bytes, stream);
}

extern "C" int hg_unpack_conv_wgrad(const float* dw_packed, float* dw_oihw, int32_t Cout, int32_t Cin,
                                    int32_t KH, int32_t KW, int32_t accumulate, hg_stream_t stream_) {
  const long long total = (long long)Cout * Cin * KH * KW;
  if (total <= 0) return 0;
  unpack_wgrad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      dw_packed, dw_oihw, Cout, Cin, KH, KW, accumulate);
  HG_LAUNCH_OK("unpack_wgrad_kernel");
  return 0;
}
