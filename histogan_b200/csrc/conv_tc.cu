// conv_tc.cu -- NHWC implicit-GEMM convolution on the 5th-gen tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM, operands staged by TMA).
//
// This one kernel is the dense contraction behind every conv of the StyleGAN2
// style generator / discriminator of the reference:
//   Conv2DMod.forward          histoGAN/histoGAN.py:420-440   (3x3 and 1x1, pad "same")
//   DiscriminatorBlock convs   histoGAN/histoGAN.py:505-526   (3x3 s1, 3x3 s2, 1x1, bias)
// and, with flipped/transposed weights, their input gradients.  The reference
// materialises per-sample weights and runs a grouped conv (:423-437); here the
// style modulation is applied to the ACTIVATIONS by the producer of `x`, the
// weights are shared, and the demodulation is a per-(sample, out-channel) scale
// in the epilogue -- algebraically identical (4e-7).
//
// GEMM view:  M = B*OH*OW (pixels), N = Cout, K = KH*KW*Cin.
//   A tile: 128 pixels x 32 channels of ONE filter tap, fetched by a 4-D TMA box
//           {32ch, TW, TH, TB} (TW*TH*TB = 128) at spatially shifted coordinates;
//           out-of-bounds rows are zero-filled by TMA = the conv's zero padding.
//   B tile: BLOCK_N out-channels x 32 channels of the same tap from the packed
//           weight [Cout][KH*KW*Cin] (K-major).
//   Both land in 128B-swizzled rows == the canonical K-major UMMA layout, so no
//   thread ever touches operand data.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA
// issuer (one elected lane), warps 2-5 and 6-9 = two epilogue sets (TMEM ->
// registers -> fused scale / bias / noise / LeakyReLU / residual -> NHWC global);
// set s drains accumulator stage s, i.e. every other tile, so two tiles are in
// their (latency-bound: TMEM load, parameter loads, 128-bit stores) epilogue at a
// time.  The kernel is persistent (one CTA per SM) with two TMEM accumulator stages.
#include <cstdlib>
#include <mutex>
#include "hg_common.cuh"
#include "sm100_ptx.cuh"

namespace hg {

constexpr int kBlockM = 128;
constexpr int kBlockK = 32;                       // fp32 channels per k-block = 128 B rows
constexpr int kABytes = kBlockM * kBlockK * 4;    // 16 KB
constexpr int kEpilogueSets = 2;                 // two sets of 4 epilogue warps alternate over the tiles
constexpr int kConvThreads = 64 + 128 * kEpilogueSets;

struct ConvArgs {
  int B, H, W, Cin, Cout, KH, KW, stride, pad, OH, OW;
  int TB, TH, TW;                 // pixel tile (TB*TH*TW == 128)
  int tiles_w, tiles_h;           // tiles per image row / column
  int m_tiles;                    // tiles_w * tiles_h * ceil(B / TB)
  int kc_per_tap;                 // Kp / 32
  int Kp;                         // Cin rounded up to 32: K extent of one tap in the packed weight
  int n_tiles;                    // round_up(Cout, 32) / BLOCK_N
  int flags;
  float slope;
  float* y;
  const float* scale;             // [B][Cout]  demodulation          (or null)
  const float* bias;              // [Cout]                            (or null)
  const float* noise;             // [B][NS][NS] image noise           (or null)
  const float* noise_w;           // [Cout]  to_noise Linear(1->Cout) weight
  const float* noise_b;           // [Cout]  to_noise bias
  const float* residual;          // NHWC like y, added after the activation (or null)
  int noise_size;
  long long y_img, y_row, y_pix;  // output strides in floats (dense NHWC unless the caller says otherwise)
  int tma_store;                  // TMAST kernels: 1 = stage + TMA-store the output tiles (0: direct stores)
  int desc_base_offset;           // HALO == 2: put (start >> 7) & 7 into the descriptor's base-offset field
  int ksplit;                     // > 1: the K loop (taps x 32-channel chunks) is split over ksplit work
                                  // items per output tile; each writes its raw partial sums to its own
                                  // slice of `partial`, conv_finish_kernel adds the slices in a fixed
                                  // order (deterministic) and applies the epilogue
  float* partial;                 // [ksplit][B*OH*OW][Cout]
};

// HALO variant (3x3, stride 1, 16x8-pixel tiles inside one image): one A box with a one-row
// halo above and below -- {32 ch, 16, 8+2, 1} = 160 pixel rows = 20 KB -- serves the three
// vertical taps of a filter column: tap kh starts kh*16 rows = kh*2048 B into the box, a
// multiple of the 1024 B swizzle atom, so the same canonical UMMA descriptor applies.  The
// activation is then fetched 3x instead of 9x from L2, which is what bounds the 256^2/128^2
// layers.
constexpr int kHaloABytes = (8 + 2) * 16 * kBlockK * 4;   // 20 KB

// RESW (halo variant, one n-tile, whole filter <= 72 KB -- the 32-channel 256^2 / 128^2 layers):
// the packed filter is loaded into shared memory ONCE per CTA and stays resident while the CTA
// walks its tiles; the ring then carries activation boxes only.  These layers are bound by the
// bytes TMA moves into shared memory (~30 B/clk/SM measured), of which the per-tile filter
// re-fetch was 36 of 96 KB.
constexpr int kResWBytes = 72 * 1024;

// HALO == 2 (resident filter only; 8-wide x 16-tall pixel tiles): ONE activation box with a one-pixel
// halo on every side -- {32 ch, 8+2, 16+2, 1} = 180 pixel rows = 22.5 KB -- serves all NINE taps.  The
// 128 MMA rows are 16 groups of 8 consecutive pixels (one tile row each); inside the box consecutive
// tile rows are 10 pixels = 1280 B apart, which is the descriptor's stride between 8-row groups (SBO),
// and tap (kh, kw) is the same matrix started (kh*10 + kw) rows = (kh*10 + kw)*128 B further on.  TMA
// and the tensor core both derive the 128B-swizzle phase from the shared-memory ADDRESS, so a start
// that is not a multiple of the 1024 B swizzle atom addresses exactly the rows TMA wrote.  The
// activation is fetched ONCE per tile (22.5 KB per 128 pixels instead of 3 x 20 KB): the 32/64-channel
// 256^2 / 128^2 layers were bound by the bytes TMA moves into shared memory.
constexpr int kHalo2ABytes = (16 + 2) * (8 + 2) * kBlockK * 4;     // 23040
constexpr int kHalo2Stage = (kHalo2ABytes + 1023) / 1024 * 1024;   // ring slots stay 1024 B aligned

// TMAST: the epilogue stages each 128-pixel x 32-channel chunk of the output tile in shared memory
// (128 B rows, 128B-swizzled: conflict-free although every lane writes its own row) and ONE thread
// stores it with cp.async.bulk.tensor (TMA): full 128 B lines leave the SM in one bulk operation.
// The direct form (each lane 16 B of its own pixel row: 32 different lines per store instruction)
// kept the LSU busy for ~256 cycles per warp and tile on the 32-channel layers.
constexpr int kStoreStage = kBlockM * 128;         // 16 KB per epilogue set

template <int BLOCK_N, int STAGES, int HALO = 0, bool RESW = false, bool TMAST = false>
struct ConvSmem {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 4;
  static constexpr int kAStage = HALO == 2 ? kHalo2Stage : (HALO ? kHaloABytes : kABytes);
  static constexpr int kStageBytes = kAStage + (RESW ? 0 : (HALO ? 3 : 1) * kBBytes);
  static constexpr int kRing = STAGES * kStageBytes;
  static constexpr int kStore = TMAST ? kEpilogueSets * kStoreStage : 0;
  static constexpr int kTotal = kRing + (RESW ? kResWBytes : 0) + kStore + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Persistent kernel: one CTA per SM walks output tiles (n fastest, so the CTAs running
// together share A tiles through L2).  Two TMEM accumulator stages let the epilogue of
// tile t overlap the main loop of tile t+1; the smem ring (STAGES deep) runs straight
// through tile boundaries.
template <int BLOCK_N, int STAGES, int HALO = 0, bool RESW = false, bool TMAST = false>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tf32_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmw,
                 const __grid_constant__ CUtensorMap tmy, const ConvArgs a) {
  static_assert(!RESW || HALO, "resident weights: halo variants only");
  static_assert(HALO != 2 || RESW, "single-box halo: resident weights only");
  using SM = ConvSmem<BLOCK_N, STAGES, HALO, RESW, TMAST>;
  constexpr uint32_t kAccCols = BLOCK_N < 32 ? 32 : BLOCK_N;
  constexpr uint32_t kTmemCols = 2 * kAccCols;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* wres = base + SM::kRing;                  // RESW: [tap][kc] filter tiles, kBBytes each
  uint8_t* store_stage = base + SM::kRing + (RESW ? kResWBytes : 0);   // TMAST: [set][128 rows x 128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + SM::kRing + (RESW ? kResWBytes : 0) + SM::kStore);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint64_t* wres_bar = tmem_empty_bar + 2;           // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(wres_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmx);
    ptx::prefetch_tmap(&tmw);
    if (TMAST) ptx::prefetch_tmap(&tmy);
  }
  if (warp == 1) {
    if (ptx::elect_one()) {
      for (int s = 0; s < STAGES; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        ptx::mbar_init(&tmem_full_bar[s], 1);
        ptx::mbar_init(&tmem_empty_bar[s], 4);        // one arrival per epilogue warp
      }
      ptx::mbar_init(wres_bar, 1);
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

  const int n_tiles = a.n_tiles;
  const int ksplit = HALO ? 1 : a.ksplit;
  const int total_tiles = a.m_tiles * n_tiles * ksplit;      // work items; the split index runs fastest
  const int taps = a.KH * a.KW;
  const int total_kb = taps * a.kc_per_tap;

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      if (RESW) {                                    // the whole filter, once
        const int n_w = 9 * a.kc_per_tap;
        ptx::mbar_expect_tx(wres_bar, n_w * SM::kBBytes);
        for (int i = 0; i < n_w; ++i) {
          const int tap = i / a.kc_per_tap, kc = i - tap * a.kc_per_tap;
          ptx::tma_load_2d(wres + i * SM::kBBytes, &tmw, wres_bar, tap * a.Kp + kc * kBlockK, 0);
        }
      }
      for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
        const int tile = item / ksplit, ks = item - tile * ksplit;
        const int mt = tile / n_tiles, n0 = (tile - mt * n_tiles) * BLOCK_N;
        const int tw_i = mt % a.tiles_w;
        const int th_i = (mt / a.tiles_w) % a.tiles_h;
        const int tb_i = mt / (a.tiles_w * a.tiles_h);
        const int iw0 = tw_i * a.TW * a.stride - a.pad, ih0 = th_i * a.TH * a.stride - a.pad;
        const int b0 = tb_i * a.TB;
        if (HALO == 2) {
          for (int kc = 0; kc < a.kc_per_tap; ++kc) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sA = base + stage * SM::kStageBytes;
            ptx::mbar_expect_tx(&full_bar[stage], kHalo2ABytes);
            ptx::tma_load_4d(sA, &tmx, &full_bar[stage], kc * kBlockK, iw0, ih0, b0);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        } else if (HALO) {
          for (int kw = 0; kw < 3; ++kw) {
            for (int kc = 0; kc < a.kc_per_tap; ++kc) {
              ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
              uint8_t* sA = base + stage * SM::kStageBytes;
              uint8_t* sB = sA + SM::kAStage;
              ptx::mbar_expect_tx(&full_bar[stage], SM::kStageBytes);
              ptx::tma_load_4d(sA, &tmx, &full_bar[stage], kc * kBlockK, iw0 + kw, ih0, b0);
              if (!RESW) {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
                  ptx::tma_load_2d(sB + kh * SM::kBBytes, &tmw, &full_bar[stage],
                                   (kh * 3 + kw) * a.Kp + kc * kBlockK, n0);
              }
              if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
          }
        } else {
          const int it0 = (int)((long long)ks * total_kb / ksplit);
          const int it1 = (int)((long long)(ks + 1) * total_kb / ksplit);
          for (int it = it0; it < it1; ++it) {
            const int tap = it / a.kc_per_tap, kc = it - tap * a.kc_per_tap;
            const int kh = tap / a.KW, kw = tap - kh * a.KW;
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sA = base + stage * SM::kStageBytes;
            uint8_t* sB = sA + kABytes;
            ptx::mbar_expect_tx(&full_bar[stage], kABytes + SM::kBBytes);
            ptx::tma_load_4d(sA, &tmx, &full_bar[stage], kc * kBlockK, iw0 + kw, ih0 + kh, b0);
            ptx::tma_load_2d(sB, &tmw, &full_bar[stage], tap * a.Kp + kc * kBlockK, n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc(2 /*tf32*/, kBlockM, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int t = 0;
      if (RESW) ptx::mbar_wait(wres_bar, 0);           // filter resident
      for (int item = blockIdx.x; item < total_tiles; item += gridDim.x, ++t) {
        const int acc = t & 1;
        const uint32_t acc_phase = (uint32_t)((t >> 1) & 1);
        ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);   // epilogue drained this stage
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccCols;
        const int ks = item % ksplit;
        const int iters = HALO == 2 ? a.kc_per_tap : HALO ? 3 * a.kc_per_tap
                               : (int)((long long)(ks + 1) * total_kb / ksplit) -
                                     (int)((long long)ks * total_kb / ksplit);
        for (int kb = 0; kb < iters; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sA = ptx::smem_u32(base + stage * SM::kStageBytes);
          if (HALO == 2) {
            // kb = kc; all nine taps read the one box: start (kh*10 + kw) pixel rows in, 8-row
            // groups 10 rows (1280 B) apart
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t a_addr = sA + (uint32_t)(((tap / 3) * 10 + tap % 3) * 128);
              const uint64_t a_desc = ptx::make_smem_desc(a_addr, 16, 1280, ptx::kLayoutSW128) |
                                      ((uint64_t)(a.desc_base_offset ? ((a_addr >> 7) & 7u) : 0u) << 49);
              const uint64_t b_desc = ptx::make_smem_desc(
                  ptx::smem_u32(wres) + (uint32_t)((tap * a.kc_per_tap + kb) * SM::kBBytes), 16, 1024,
                  ptx::kLayoutSW128);
#pragma unroll
              for (int k = 0; k < kBlockK / 8; ++k)
                ptx::mma_tf32_ss(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                                 (uint32_t)((kb | tap | k) != 0));
            }
            ptx::tc_commit(&empty_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            continue;
          }
#pragma unroll
          for (int kh = 0; kh < (HALO ? 3 : 1); ++kh) {
            // HALO: vertical tap kh = the same box 16 pixel rows (2 swizzle atoms) further down
            const uint64_t a_desc = ptx::make_smem_desc(sA + (uint32_t)(kh * 2048), 16, 1024,
                                                        ptx::kLayoutSW128);
            // HALO iteration kb = (kw, kc): kw = kb / kc_per_tap
            const uint32_t sBt =
                RESW ? ptx::smem_u32(wres) + (uint32_t)(((kh * 3 + kb / a.kc_per_tap) * a.kc_per_tap +
                                                        kb % a.kc_per_tap) * SM::kBBytes)
                     : sA + SM::kAStage + (uint32_t)(kh * SM::kBBytes);
            const uint64_t b_desc = ptx::make_smem_desc(sBt, 16, 1024, ptx::kLayoutSW128);
#pragma unroll
            for (int k = 0; k < kBlockK / 8; ++k) {
              // advance 8 tf32 = 32 B along K inside the 128 B swizzle row: +2 in (addr >> 4) units
              ptx::mma_tf32_ss(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                               (uint32_t)((kb | kh | k) != 0));
            }
          }
          ptx::tc_commit(&empty_bar[stage]);          // frees the smem slot when the MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        ptx::tc_commit(&tmem_full_bar[acc]);          // accumulator of this tile complete
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue --
    const int q = warp & 3;                          // TMEM lane quadrant this warp may read
    const int set = (warp - 2) >> 2;                 // this warp's epilogue set == its accumulator stage
    const int row = q * 32 + lane;
    const int tw = row % a.TW, th = (row / a.TW) % a.TH, tb = row / (a.TW * a.TH);
    int t = 0;
    for (int item = blockIdx.x; item < total_tiles; item += gridDim.x, ++t) {
      if ((t & (kEpilogueSets - 1)) != set) continue;
      const int acc = t & 1;
      const uint32_t acc_phase = (uint32_t)((t >> 1) & 1);
      const int tile = item / ksplit;
      const int mt = tile / n_tiles, n0 = (tile - mt * n_tiles) * BLOCK_N;
      const int tw_i = mt % a.tiles_w;
      const int th_i = (mt / a.tiles_w) % a.tiles_h;
      const int tb_i = mt / (a.tiles_w * a.tiles_h);
      const int b = tb_i * a.TB + tb, oh = th_i * a.TH + th, ow = tw_i * a.TW + tw;
      const bool valid = b < a.B && oh < a.OH && ow < a.OW;
      const long long pix = ((long long)b * a.OH + oh) * a.OW + ow;
      float nz = 0.f;
      if (a.noise && valid)                          // spatially transposed (histoGAN.py:466-467)
        nz = a.noise[((long long)b * a.noise_size + ow) * a.noise_size + oh];
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * kAccCols +
                                    (uint32_t)c0, v);
        ptx::tmem_ld_wait();
        if (c0 + 32 >= BLOCK_N) {                    // last read of this accumulator stage:
          ptx::tc_fence_before();                    // hand it back to the MMA warp
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
        }
        if (TMAST && a.tma_store) {
          // every thread of the set takes part (named barriers); rows outside the image / channels
          // beyond Cout are clipped by the TMA store, their parameter loads are guarded
          const int n = n0 + c0;
          const bool live = n < a.Cout;                  // warp-uniform
          const int ncols = live ? min(32, a.Cout - n) : 0;
          const float* ro = (a.residual && valid) ? a.residual + pix * a.Cout + n : nullptr;
          const float* sc = (a.scale && valid) ? a.scale + (long long)b * a.Cout + n : nullptr;
          uint8_t* stg = store_stage + set * kStoreStage;
          const bool issuer = q == 2 && lane == 0;       // first warp of the set (warp & 3 == 2)
          if (live) {
            if (issuer) ptx::bulk_wait_group_read0();    // the previous store has read the staging tile
            ptx::named_bar_sync(1 + set, 128);
#define HG_EPILOGUE_ST(J)                                                                     \
            {                                                                                 \
              float4 o = make_float4(__uint_as_float(v[(J)]), __uint_as_float(v[(J) + 1]),    \
                                     __uint_as_float(v[(J) + 2]), __uint_as_float(v[(J) + 3])); \
              if ((J) < ncols) {                                                              \
                if (sc) {                                                                     \
                  const float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + (J)));         \
                  o.x *= s4.x; o.y *= s4.y; o.z *= s4.z; o.w *= s4.w;                         \
                }                                                                             \
                if (a.bias) {                                                                 \
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + n + (J))); \
                  o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;                         \
                }                                                                             \
                if (a.noise) {                                                                \
                  const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.noise_w + n + (J))); \
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.noise_b + n + (J))); \
                  o.x = fmaf(nz, w4.x, o.x + b4.x); o.y = fmaf(nz, w4.y, o.y + b4.y);         \
                  o.z = fmaf(nz, w4.z, o.z + b4.z); o.w = fmaf(nz, w4.w, o.w + b4.w);         \
                }                                                                             \
                if (a.flags & HG_CONV_LRELU) {                                                \
                  o.x = o.x > 0.f ? o.x : o.x * a.slope; o.y = o.y > 0.f ? o.y : o.y * a.slope; \
                  o.z = o.z > 0.f ? o.z : o.z * a.slope; o.w = o.w > 0.f ? o.w : o.w * a.slope; \
                }                                                                             \
                if (ro) {                                                                     \
                  const float4 r4 = __ldg(reinterpret_cast<const float4*>(ro + (J)));         \
                  o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;                         \
                }                                                                             \
                if (a.flags & HG_CONV_ROUND_TF32) {                                           \
                  o.x = tf32_round(o.x); o.y = tf32_round(o.y);                               \
                  o.z = tf32_round(o.z); o.w = tf32_round(o.w);                               \
                }                                                                             \
              }                                                                               \
              /* 128B swizzle: 16-byte chunk j of row r sits at chunk j ^ (r & 7) */           \
              *reinterpret_cast<float4*>(stg + row * 128 + ((((J) >> 2) ^ (row & 7)) << 4)) = o; \
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4) HG_EPILOGUE_ST(j)
#undef HG_EPILOGUE_ST
            ptx::fence_proxy_async();                    // generic-proxy writes -> visible to the TMA engine
            ptx::named_bar_sync(1 + set, 128);
            if (issuer) {
              ptx::tma_store_4d(&tmy, stg, n, tw_i * a.TW, th_i * a.TH, tb_i * a.TB);
              ptx::bulk_commit_group();
            }
          }
          continue;
        }
        if (valid && n0 + c0 < a.Cout) {
          const int n = n0 + c0;
          const int ncols = min(32, a.Cout - n);       // Cout % 4 == 0
          float* yo = a.y + (long long)b * a.y_img + (long long)oh * a.y_row + (long long)ow * a.y_pix + n;
          const float* ro = a.residual ? a.residual + pix * a.Cout + n : nullptr;
          const float* sc = a.scale ? a.scale + (long long)b * a.Cout + n : nullptr;
          if (ksplit > 1) {             // raw partial sums; conv_finish_kernel does the rest
            float* po = a.partial + ((long long)(item - tile * ksplit) * a.B * a.OH * a.OW + pix) * a.Cout + n;
#pragma unroll
            for (int j = 0; j < 32; j += 4)      // constant indices keep v[] in registers
              if (j < ncols)
                *reinterpret_cast<float4*>(po + j) =
                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            continue;
          }
// One group of 4 channels: vector loads of the per-channel operands, branches only on
          // warp-uniform kernel arguments.  (The scalar per-element form of this chain cost ~970
          // instructions per warp and tile and bounded the 32-channel layers: ncu showed 63 M warp
          // instructions for a 178 us launch with the tensor pipe 21 % active.)
#define HG_EPILOGUE_4(J)                                                                      \
          {                                                                                   \
            float4 o = make_float4(__uint_as_float(v[(J)]), __uint_as_float(v[(J) + 1]),      \
                                   __uint_as_float(v[(J) + 2]), __uint_as_float(v[(J) + 3])); \
            if (sc) {                                                                         \
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + (J)));             \
              o.x *= s4.x; o.y *= s4.y; o.z *= s4.z; o.w *= s4.w;                             \
            }                                                                                 \
            if (a.bias) {                                                                     \
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + n + (J)));     \
              o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;                             \
            }                                                                                 \
            if (a.noise) {                                                                    \
              const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.noise_w + n + (J)));  \
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.noise_b + n + (J)));  \
              o.x = fmaf(nz, w4.x, o.x + b4.x); o.y = fmaf(nz, w4.y, o.y + b4.y);             \
              o.z = fmaf(nz, w4.z, o.z + b4.z); o.w = fmaf(nz, w4.w, o.w + b4.w);             \
            }                                                                                 \
            if (a.flags & HG_CONV_LRELU) {                                                    \
              o.x = o.x > 0.f ? o.x : o.x * a.slope; o.y = o.y > 0.f ? o.y : o.y * a.slope;   \
              o.z = o.z > 0.f ? o.z : o.z * a.slope; o.w = o.w > 0.f ? o.w : o.w * a.slope;   \
            }                                                                                 \
            if (ro) {                                                                         \
              const float4 r4 = __ldg(reinterpret_cast<const float4*>(ro + (J)));             \
              o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;                             \
            }                                                                                 \
            if (a.flags & HG_CONV_ROUND_TF32) {                                               \
              o.x = tf32_round(o.x); o.y = tf32_round(o.y);                                   \
              o.z = tf32_round(o.z); o.w = tf32_round(o.w);                                   \
            }                                                                                 \
            *reinterpret_cast<float4*>(yo + (J)) = o;                                         \
          }
          if (ncols == 32) {            // common case: fully unrolled, loads batched by the compiler
#pragma unroll
            for (int j = 0; j < 32; j += 4) HG_EPILOGUE_4(j)
          } else {                      // channel tail (Cout not a multiple of 32)
#pragma unroll 1
            for (int j = 0; j < ncols; j += 4) {
              switch (j) {              // keep v[] in registers: constant indices only
                case 0: HG_EPILOGUE_4(0) break;
                case 4: HG_EPILOGUE_4(4) break;
                case 8: HG_EPILOGUE_4(8) break;
                case 12: HG_EPILOGUE_4(12) break;
                case 16: HG_EPILOGUE_4(16) break;
                case 20: HG_EPILOGUE_4(20) break;
                case 24: HG_EPILOGUE_4(24) break;
                default: HG_EPILOGUE_4(28) break;
              }
            }
          }
#undef HG_EPILOGUE_4
        }
      }
    }
  }

  if (TMAST && warp >= 2 && (warp & 3) == 2 && lane == 0) ptx::bulk_wait_group0();   // stores landed
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

constexpr bool kResidentWDefault = true;   // HG_CONV_RESIDENT_W=0 disables it
constexpr bool kSplitKDefault = true;      // HG_CONV_SPLITK=0 disables it

// epilogue of a split-K convolution: adds the ksplit partial slices in index order and applies
// the same element-wise chain as HG_EPILOGUE_4; writes the dense NHWC y
__global__ void __launch_bounds__(256)
conv_finish_kernel(const ConvArgs a) {
  const long long n4 = (long long)a.B * a.OH * a.OW * (a.Cout / 4);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int n = (int)(i % (a.Cout / 4)) * 4;
  const long long pix = i / (a.Cout / 4);
  const int ow = (int)(pix % a.OW);
  const int oh = (int)((pix / a.OW) % a.OH);
  const int b = (int)(pix / ((long long)a.OW * a.OH));
  float* yo = a.y + pix * a.Cout + n;
  const long long slice = (long long)a.B * a.OH * a.OW * a.Cout;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ks = 0; ks < a.ksplit; ++ks) {
    const float4 p4 = *reinterpret_cast<const float4*>(a.partial + ks * slice + pix * a.Cout + n);
    v[0] += p4.x; v[1] += p4.y; v[2] += p4.z; v[3] += p4.w;
  }
  float nz = 0.f;
  if (a.noise) nz = a.noise[((long long)b * a.noise_size + ow) * a.noise_size + oh];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float tv = v[e];
    if (a.scale) tv *= a.scale[(long long)b * a.Cout + n + e];
    if (a.bias) tv += a.bias[n + e];
    if (a.noise) tv = fmaf(nz, a.noise_w[n + e], tv + a.noise_b[n + e]);
    if (a.flags & HG_CONV_LRELU) tv = tv > 0.f ? tv : tv * a.slope;
    if (a.residual) tv += a.residual[pix * a.Cout + n + e];
    if (a.flags & HG_CONV_ROUND_TF32) tv = tf32_round(tv);
    v[e] = tv;
  }
  *reinterpret_cast<float4*>(yo) = make_float4(v[0], v[1], v[2], v[3]);
}

// ------------------------------------------------- weight packing kernel ----
// OIHW parameter (state_dict layout of Conv2DMod.weight / nn.Conv2d.weight)
//   -> [Np][KH][KW][Kp] K-major, TF32-rounded, Np = round_up(N, 32), Kp = round_up(K, 32),
//      zero padded (the kernels always move 32-channel boxes; tensors with fewer channels are
//      completed by TMA's out-of-bounds zero fill, the weights by these explicit zeros).
// mode 0 (forward):  N = Cout, K = Cin : out[co][kh][kw][ci] = w[co][ci][kh][kw]
// mode 1 (dgrad):    N = Cin, K = Cout : out[ci][kh][kw][co] = w[co][ci][KH-1-kh][KW-1-kw]
// ohwi: the parameter is stored channels_last (memory [Cout][KH][KW][Cin], what the modules of
// this package use): mode 0 is then a rounding copy, mode 1 a per-tap transpose (kernel below).
__global__ void pack_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout,
                                   int Cin, int KH, int KW, int mode, int Np, int Kp, int ohwi) {
  const long long total = (long long)Np * KH * KW * Kp;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int N = mode ? Cin : Cout, K = mode ? Cout : Cin;
  const int k = (int)(e % Kp);
  long long r = e / Kp;
  const int kw = (int)(r % KW); r /= KW;
  const int kh = (int)(r % KH); r /= KH;
  const int n = (int)r;
  float v = 0.f;                                   // zero padding of both the N and the K extent
  if (n < N && k < K) {
    if (ohwi) {
      if (mode == 0) v = w[(((long long)n * KH + kh) * KW + kw) * Cin + k];
      else v = w[(((long long)k * KH + (KH - 1 - kh)) * KW + (KW - 1 - kw)) * Cin + n];
    } else {
      if (mode == 0) v = w[(((long long)n * Cin + k) * KH + kh) * KW + kw];
      else v = w[(((long long)k * Cin + n) * KH + (KH - 1 - kh)) * KW + (KW - 1 - kw)];
    }
  }
  out[e] = tf32_round(v);
}

// mode 1 from a channels_last parameter: out[ci][tap][co] = w[co][T-1-tap][ci]: a 64x64 shared-memory
// transpose per tap with 128-bit accesses on both sides (reads along ci, writes along co).
// grid (Np/64, Kp/64, KH*KW), block 256.  HBM-bound: one read + one write of the weight tensor.
__global__ void __launch_bounds__(256)
pack_weight_transpose_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin,
                             int T, int Np, int Kp) {
  __shared__ float tile[64][65];
  const int ci0 = blockIdx.x * 64, co0 = blockIdx.y * 64, tap = blockIdx.z;
  const int q = threadIdx.x & 15, r = threadIdx.x >> 4;            // 16 float4 columns x 16 rows
  const bool vec_in = (Cin & 3) == 0;
#pragma unroll
  for (int rr = 0; rr < 64; rr += 16) {
    const int co = co0 + rr + r, ci = ci0 + q * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (co < Cout) {
      const float* src = w + ((long long)co * T + (T - 1 - tap)) * Cin + ci;
      if (vec_in && ci + 3 < Cin) v = *reinterpret_cast<const float4*>(src);
      else {
        if (ci < Cin) v.x = src[0];
        if (ci + 1 < Cin) v.y = src[1];
        if (ci + 2 < Cin) v.z = src[2];
        if (ci + 3 < Cin) v.w = src[3];
      }
    }
    tile[rr + r][q * 4 + 0] = v.x; tile[rr + r][q * 4 + 1] = v.y;
    tile[rr + r][q * 4 + 2] = v.z; tile[rr + r][q * 4 + 3] = v.w;
  }
  __syncthreads();
#pragma unroll
  for (int rr = 0; rr < 64; rr += 16) {
    const int ci = ci0 + rr + r, co = co0 + q * 4;                 // Np, Kp are multiples of 32: guard
    if (ci < Np && co < Kp) {
      const float4 o = make_float4(tf32_round(tile[q * 4 + 0][rr + r]), tf32_round(tile[q * 4 + 1][rr + r]),
                                   tf32_round(tile[q * 4 + 2][rr + r]), tf32_round(tile[q * 4 + 3][rr + r]));
      *reinterpret_cast<float4*>(out + ((long long)ci * T + tap) * Kp + co) = o;
    }
  }
}

// ------------------------------------------------------------ host side -----
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_encodeTiled>(p);
  }();
  return fn;
}

static int encode_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims,
                      const cuuint64_t* strides_bytes, const cuuint32_t* box,
                      const cuuint32_t* estr, CUtensorMapL2promotion promo) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return set_error(HG_EARCH, "cuTensorMapEncodeTiled not available from the driver");
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), dims,
                  strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(HG_EINVAL, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return 0;
}

// Per-device scratch for the split-K partial sums: ONE fixed 64 MB allocation made on first use
// (never during stream capture, never freed or moved -- captured CUDA graphs keep its address).
// A convolution that would need more, or whose first use falls inside a capture, runs unsplit.
// CONTRACT: one scratch per device => split-K convolutions of one device must be issued on ONE
// stream (or on streams ordered against each other); the Trainer does exactly that.  Creation is
// guarded by a mutex (several host threads / devices).
constexpr size_t kSplitKWsBytes = 64u << 20;

static float* splitk_workspace(size_t bytes, cudaStream_t stream) {
  static float* ws[64] = {};
  static std::mutex mu;
  if (bytes > kSplitKWsBytes) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  if (ws[dev]) return ws[dev];
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return nullptr;
  float* p = nullptr;
  if (cudaMalloc(&p, kSplitKWsBytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  ws[dev] = p;
  return p;
}

template <int BLOCK_N, int STAGES, int HALO = 0, bool RESW = false, bool TMAST = false>
static int launch_conv(const CUtensorMap& tmx, const CUtensorMap& tmw, const CUtensorMap& tmy,
                       const ConvArgs& a, int m_tiles, cudaStream_t stream) {
  using SM = ConvSmem<BLOCK_N, STAGES, HALO, RESW, TMAST>;
  static_assert(SM::kTotal <= 227 * 1024, "shared memory budget");
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    HG_CUDA_OK(cudaFuncSetAttribute(conv_tf32_kernel<BLOCK_N, STAGES, HALO, RESW, TMAST>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal));
    attr_once.mark();
  }
  const int total = m_tiles * a.n_tiles * (HALO ? 1 : a.ksplit);
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int grid = total < sms ? total : sms;
  const bool split = !HALO && a.ksplit > 1;
  conv_tf32_kernel<BLOCK_N, STAGES, HALO, RESW, TMAST><<<grid, kConvThreads, SM::kTotal, stream>>>(tmx, tmw, tmy, a);
  HG_LAUNCH_OK("conv_tf32_kernel");
  if (split) {
    const long long n4 = (long long)a.B * a.OH * a.OW * (a.Cout / 4);
    conv_finish_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(a);
    HG_LAUNCH_OK("conv_finish_kernel");
  }
  return 0;
}

}  // namespace hg

using namespace hg;

extern "C" int hg_pack_conv_weight(const float* w, float* w_packed, int32_t Cout, int32_t Cin,
                                   int32_t KH, int32_t KW, int32_t mode_, hg_stream_t stream_) {
  if (!w || !w_packed) return set_error(HG_EINVAL, "null tensor pointer");
  if ((long long)Cout * Cin * KH * KW <= 0) return set_error(HG_EINVAL, "empty weight");
  if (mode_ < 0 || mode_ > 3) return set_error(HG_EINVAL, "unknown pack mode %d", mode_);
  const int mode = mode_ & 1, ohwi = (mode_ & HG_PACK_FROM_OHWI) ? 1 : 0;
  const int N = mode ? Cin : Cout, K = mode ? Cout : Cin;
  const int Np = (N + 31) / 32 * 32, Kp = (K + 31) / 32 * 32;
  if (ohwi && mode == 1 && KH * KW <= 65535) {
    dim3 grid((Np + 63) / 64, (Kp + 63) / 64, KH * KW);
    pack_weight_transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(w, w_packed, Cout, Cin, KH * KW, Np,
                                                                          Kp);
    HG_LAUNCH_OK("pack_weight_transpose_kernel");
    return 0;
  }
  const long long total = (long long)Np * KH * KW * Kp;
  pack_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      w, w_packed, Cout, Cin, KH, KW, mode, Np, Kp, ohwi);
  HG_LAUNCH_OK("pack_weight_kernel");
  return 0;
}

extern "C" int hg_conv2d_fwd(const float* x, const float* w_packed, float* y,
                             const hg_conv_params* p, const hg_conv_epilogue* ep,
                             hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!x || !w_packed || !y || !p) return set_error(HG_EINVAL, "null pointer");
  if (p->B <= 0) return 0;
  if (p->Cin % 4 != 0 || p->Cout % 4 != 0)
    return set_error(HG_ENOSUP, "conv: Cin=%d and Cout=%d must be multiples of 4 (16-byte TMA rows)",
                     p->Cin, p->Cout);
  if (p->stride < 1 || p->stride > 2) return set_error(HG_ENOSUP, "conv: stride must be 1 or 2");
  const int OHn = (p->H + 2 * p->pad - p->KH) / p->stride + 1;
  const int OWn = (p->W + 2 * p->pad - p->KW) / p->stride + 1;
  const int OH = p->OH, OW = p->OW;          // >= the natural extent: the excess reads zeros
  if (OH < OHn || OW < OWn || OH < 1 || OW < 1 || OH > p->H + 2 * p->pad || OW > p->W + 2 * p->pad)
    return set_error(HG_EINVAL, "conv: OH/OW (%d,%d) inconsistent with geometry (%d,%d)", p->OH,
                     p->OW, OHn, OWn);
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_packed) |
       reinterpret_cast<uintptr_t>(y)) & 15)
    return set_error(HG_EINVAL, "conv: pointers must be 16-byte aligned");

  ConvArgs a{};
  a.B = p->B; a.H = p->H; a.W = p->W; a.Cin = p->Cin; a.Cout = p->Cout;
  a.KH = p->KH; a.KW = p->KW; a.stride = p->stride; a.pad = p->pad; a.OH = OH; a.OW = OW;
  // pixel tile: as wide as the row allows (<= 16), then rows, then images
  int TW = 1; while (TW < 16 && TW < OW) TW <<= 1;
  int TH = 1; while (TW * TH < kBlockM && TH < OH) TH <<= 1;
  int TB = kBlockM / (TW * TH);
  // single-box halo (HALO == 2): 3x3 / stride 1, one n-tile, the whole filter resident (<= 72 KB, one
  // 32-channel chunk per tap or two), enough tiles per CTA -- the 32/64-channel 256^2 / 128^2 layers
  static const int halo2_mode = [] {
    const char* e = getenv("HG_CONV_HALO2");        // 0 = off, 1 = on, 2 = on with the base-offset field
    return e ? atoi(e) : 1;
  }();
  const int Kp_ = (p->Cin + 31) / 32 * 32, Np_ = (p->Cout + 31) / 32 * 32;
  const int BN_ = (Np_ % 128 == 0) ? 128 : (Np_ % 64 == 0 ? 64 : 32);
  const bool halo2 = halo2_mode > 0 && p->KH == 3 && p->KW == 3 && p->stride == 1 && p->pad == 1 &&
                     OW >= 8 && OH >= 16 && Np_ == BN_ && BN_ <= 64 &&
                     9 * (Kp_ / kBlockK) * BN_ * kBlockK * 4 <= kResWBytes &&
                     (long long)((OW + 7) / 8) * ((OH + 15) / 16) * p->B >= 4 * 148;
  if (halo2) { TW = 8; TH = 16; TB = 1; }
  a.desc_base_offset = halo2_mode == 2;
  a.TW = TW; a.TH = TH; a.TB = TB;
  a.tiles_w = (OW + TW - 1) / TW;
  a.tiles_h = (OH + TH - 1) / TH;
  const int tiles_b = (p->B + TB - 1) / TB;
  const int m_tiles = a.tiles_w * a.tiles_h * tiles_b;
  a.m_tiles = m_tiles;
  const int Kp = (p->Cin + 31) / 32 * 32, Np = (p->Cout + 31) / 32 * 32;
  a.Kp = Kp;
  a.kc_per_tap = Kp / kBlockK;
  a.flags = ep ? ep->flags : 0;
  a.slope = ep ? ep->lrelu_slope : 0.2f;
  a.y = y;
  a.scale = ep ? ep->scale : nullptr;
  a.bias = ep ? ep->bias : nullptr;
  a.noise = ep ? ep->noise : nullptr;
  a.noise_w = ep ? ep->noise_w : nullptr;
  a.noise_b = ep ? ep->noise_b : nullptr;
  a.noise_size = ep ? ep->noise_size : 0;
  a.residual = ep ? ep->residual : nullptr;
  a.y_img = (long long)OH * OW * p->Cout; a.y_row = (long long)OW * p->Cout; a.y_pix = p->Cout;
  if (ep && (ep->out_img_stride || ep->out_row_stride || ep->out_pix_stride)) {
    if (a.residual) return set_error(HG_ENOSUP, "conv: strided output together with a residual");
    if (ep->out_pix_stride < p->Cout || (ep->out_pix_stride | ep->out_row_stride | ep->out_img_stride) & 3)
      return set_error(HG_EINVAL, "conv: output strides must be multiples of 4 floats, pixel stride >= Cout");
    a.y_img = ep->out_img_stride; a.y_row = ep->out_row_stride; a.y_pix = ep->out_pix_stride;
  }
  if (a.noise && (!a.noise_w || !a.noise_b || a.noise_size < OH || a.noise_size < OW))
    return set_error(HG_EINVAL, "conv: noise needs noise_w/noise_b and noise_size >= OH,OW");

  // 3x3 / stride 1 / 16x8 tiles within one image: fetch each activation box once per filter
  // COLUMN (with a one-row halo) instead of once per tap
  const bool halo = !halo2 && p->KH == 3 && p->KW == 3 && p->stride == 1 && p->pad == 1 && TW == 16 &&
                    TH == 8 && TB == 1;
  // x: NHWC as a 4-D tensor {C, W, H, B}; strided boxes implement stride-2 convs
  alignas(64) CUtensorMap tmx, tmw;
  {
    cuuint64_t dims[4] = {(cuuint64_t)p->Cin, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->B};
    cuuint64_t strides[3] = {(cuuint64_t)p->Cin * 4, (cuuint64_t)p->W * p->Cin * 4,
                             (cuuint64_t)p->H * p->W * p->Cin * 4};
    cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)(halo2 ? TW + 2 : TW * p->stride),
                         (cuuint32_t)(halo || halo2 ? TH + 2 : TH * p->stride), (cuuint32_t)TB};
    cuuint32_t estr[4] = {1, (cuuint32_t)p->stride, (cuuint32_t)p->stride, 1};
    int rc = encode_map(&tmx, x, 4, dims, strides, box, estr, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc) return rc;
  }
  const int Ktot = p->KH * p->KW * Kp;
  const int BN = (Np % 128 == 0) ? 128 : (Np % 64 == 0 ? 64 : 32);
  a.n_tiles = Np / BN;
  // few output tiles but a long K loop (the 4x4 / 2x2 layers stream up to 151 MB of weights
  // through a handful of SMs): split K over several CTAs per tile
  a.ksplit = 1;
  {
    const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
    const int base = m_tiles * a.n_tiles, iters = p->KH * p->KW * a.kc_per_tap;
    const bool dense_out = a.y_pix == p->Cout && a.y_row == (long long)OW * p->Cout &&
                           a.y_img == (long long)OH * OW * p->Cout;
    static const bool enabled = [] {
      const char* e = getenv("HG_CONV_SPLITK");
      return e ? e[0] != '0' : kSplitKDefault;
    }();
    if (enabled && !halo && !halo2 && dense_out && 2 * base <= sms && iters >= 32) {
      int ks = sms / base;
      if (ks > iters / 8) ks = iters / 8;
      if (ks >= 2) {
        a.partial = splitk_workspace(sizeof(float) * (size_t)ks * p->B * OH * OW * p->Cout, stream);
        if (a.partial) a.ksplit = ks;
      }
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Np};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 4};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    int rc = encode_map(&tmw, w_packed, 2, dims, strides, box, estr, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
  }
  // TMA-store epilogue for the narrow layers (BLOCK_N <= 64: little math per output byte, the epilogue
  // dominates) whenever the output is expressible as a tensor map and the K loop is not split
  static const bool tmast_enabled = [] {
    const char* e = getenv("HG_CONV_TMA_STORE");
    return e ? e[0] != '0' : true;
  }();
  alignas(64) CUtensorMap tmy = tmx;
  a.tma_store = 0;
  if (tmast_enabled && BN <= 64 && a.ksplit == 1) {
    cuuint64_t dims[4] = {(cuuint64_t)p->Cout, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)p->B};
    cuuint64_t strides[3] = {(cuuint64_t)a.y_pix * 4, (cuuint64_t)a.y_row * 4, (cuuint64_t)a.y_img * 4};
    cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TB};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (encode_map(&tmy, y, 4, dims, strides, box, estr, CU_TENSOR_MAP_L2_PROMOTION_NONE) == 0)
      a.tma_store = 1;
  }
  if (halo2) {
    a.ksplit = 1;
    if (BN == 64) return launch_conv<64, 4, 2, true, true>(tmx, tmw, tmy, a, m_tiles, stream);
    return launch_conv<32, 4, 2, true, true>(tmx, tmw, tmy, a, m_tiles, stream);
  }
  if (halo) {
    static const bool resw_enabled = [] {
      const char* e = getenv("HG_CONV_RESIDENT_W");
      return e ? e[0] != '0' : kResidentWDefault;
    }();
    // whole filter resident in shared memory (one n-tile, <= 72 KB, enough tiles per CTA to pay
    // for the up-front load)
    // (measured: 197 -> 180 us on 32->32 @256^2; with two k-chunks per tap, 64->32, it is a
    // loss -- 242 -> 255 us -- so one chunk per tap only)
    if (resw_enabled && a.n_tiles == 1 && a.kc_per_tap == 1 &&
        9 * a.kc_per_tap * BN * kBlockK * 4 <= kResWBytes && m_tiles >= 4 * 148) {
      if (BN == 64) return launch_conv<64, 5, 1, true, true>(tmx, tmw, tmy, a, m_tiles, stream);
      if (BN == 32) return launch_conv<32, 5, 1, true, true>(tmx, tmw, tmy, a, m_tiles, stream);
    }
    if (BN == 128) return launch_conv<128, 3, 1>(tmx, tmw, tmy, a, m_tiles, stream);
    if (BN == 64) return launch_conv<64, 4, 1, false, true>(tmx, tmw, tmy, a, m_tiles, stream);
    return launch_conv<32, 5, 1, false, true>(tmx, tmw, tmy, a, m_tiles, stream);
  }
  if (BN == 128) return launch_conv<128, 6>(tmx, tmw, tmy, a, m_tiles, stream);
  if (BN == 64) return launch_conv<64, 7, 0, false, true>(tmx, tmw, tmy, a, m_tiles, stream);
  return launch_conv<32, 8, 0, false, true>(tmx, tmw, tmy, a, m_tiles, stream);
}
