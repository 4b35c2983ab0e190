// optim.cu -- fused multi-tensor DiffGrad step.
//
// The reference steps torch_optimizer.DiffGrad (histoGAN/histoGAN.py:670-671,932,989),
// a per-parameter Python loop of ~10 element-wise launches.  One pass here:
//   m = b1 m + (1-b1) g ;  v = b2 v + (1-b2) g^2 ;  xi = sigmoid(|g_prev - g|)
//   p -= step_size * (m * xi) / (sqrt(v) + eps) ;   g_prev = g
// for up to kMaxTensors tensors per launch (pointers in kernel-parameter space).  For conv weights
// whose tensor-core operand is a plain TF32-rounded copy of the parameter the kernel also writes
// that copy (`packed`), so no separate packing pass reads the weights again.
#include "hg_common.cuh"
#include "sm100_ptx.cuh"

namespace hg {

constexpr int kMaxTensors = 48;
constexpr int kChunk = 1 << 16;            // elements per CTA

struct DiffGradBatch {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];
  float* m[kMaxTensors];
  float* v[kMaxTensors];
  float* prev[kMaxTensors];
  float* packed[kMaxTensors];              // optional TF32-rounded copy of the NEW parameter (or null)
  long long n[kMaxTensors];
  int first_block[kMaxTensors + 1];        // prefix sum of ceil(n / kChunk)
  int count;
};

__device__ __forceinline__ float diffgrad_one(float pi, float gi, float& mi, float& vi, float& pvi,
                                              float beta1, float beta2, float eps, float step_size,
                                              float weight_decay) {
  if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
  mi = fmaf(beta1, mi, (1.f - beta1) * gi);
  vi = fmaf(beta2, vi, (1.f - beta2) * gi * gi);
  const float diff = fabsf(pvi - gi);
  const float xi = 1.f / (1.f + expf(-diff));
  pvi = gi;
  return pi - step_size * (mi * xi) / (sqrtf(vi) + eps);
}

// HBM-bound: 5 reads + 4 writes (+1 for the packed copy) of 4 B per element; 128-bit accesses
// (VEC: every pointer of the batch is 16-byte aligned -- torch allocations always are).
template <bool VEC>
__global__ void __launch_bounds__(256)
diffgrad_kernel(const DiffGradBatch t, float beta1, float beta2, float eps, float step_size,
                float weight_decay) {
  int ti = 0;
  while (ti + 1 < t.count && (int)blockIdx.x >= t.first_block[ti + 1]) ++ti;
  const long long base = (long long)(blockIdx.x - t.first_block[ti]) * kChunk;
  const long long n = t.n[ti];
  float* __restrict__ p = t.p[ti];
  const float* __restrict__ g = t.g[ti];
  float* __restrict__ m = t.m[ti];
  float* __restrict__ v = t.v[ti];
  float* __restrict__ pv = t.prev[ti];
  float* __restrict__ pk = t.packed[ti];
  const long long end = min(n, base + kChunk);
  long long i = base + (VEC ? threadIdx.x * 4 : threadIdx.x);
  if (VEC) {
    for (; i + 3 < end; i += 256 * 4) {
      const float4 g4 = *reinterpret_cast<const float4*>(g + i);
      float4 p4 = *reinterpret_cast<const float4*>(p + i);
      float4 m4 = *reinterpret_cast<const float4*>(m + i);
      float4 v4 = *reinterpret_cast<const float4*>(v + i);
      float4 q4 = *reinterpret_cast<const float4*>(pv + i);
      p4.x = diffgrad_one(p4.x, g4.x, m4.x, v4.x, q4.x, beta1, beta2, eps, step_size, weight_decay);
      p4.y = diffgrad_one(p4.y, g4.y, m4.y, v4.y, q4.y, beta1, beta2, eps, step_size, weight_decay);
      p4.z = diffgrad_one(p4.z, g4.z, m4.z, v4.z, q4.z, beta1, beta2, eps, step_size, weight_decay);
      p4.w = diffgrad_one(p4.w, g4.w, m4.w, v4.w, q4.w, beta1, beta2, eps, step_size, weight_decay);
      *reinterpret_cast<float4*>(m + i) = m4;
      *reinterpret_cast<float4*>(v + i) = v4;
      *reinterpret_cast<float4*>(pv + i) = q4;
      *reinterpret_cast<float4*>(p + i) = p4;
      if (pk)
        *reinterpret_cast<float4*>(pk + i) =
            make_float4(tf32_round(p4.x), tf32_round(p4.y), tf32_round(p4.z), tf32_round(p4.w));
    }
    // tail of the chunk (n % 4 elements of the last chunk): the thread whose group is cut finishes it
    if (i < end) {
      for (long long j = i; j < end; ++j) {
        float mi = m[j], vi = v[j], qi = pv[j];
        const float pn = diffgrad_one(p[j], g[j], mi, vi, qi, beta1, beta2, eps, step_size, weight_decay);
        m[j] = mi; v[j] = vi; pv[j] = qi; p[j] = pn;
        if (pk) pk[j] = tf32_round(pn);
      }
    }
  } else {
    for (; i < end; i += 256) {
      float mi = m[i], vi = v[i], qi = pv[i];
      const float pn = diffgrad_one(p[i], g[i], mi, vi, qi, beta1, beta2, eps, step_size, weight_decay);
      m[i] = mi; v[i] = vi; pv[i] = qi; p[i] = pn;
      if (pk) pk[i] = tf32_round(pn);
    }
  }
}

// ---------------------------------------------------------------------------
// Exponential moving average of the generator-side weights (HistoGAN.EMA, histoGAN/histoGAN.py
// :698-707, every 10th step after step 20 000):  ma = beta * ma + (1 - beta) * cur  for up to
// kMaxTensors tensors per launch.  The reference loops over ~150 parameters with 3 element-wise
// launches each; this is one pass (2 reads + 1 write per element).
struct EmaBatch {
  float* ma[kMaxTensors];
  const float* cur[kMaxTensors];
  long long n[kMaxTensors];
  int first_block[kMaxTensors + 1];
  int count;
};

__global__ void __launch_bounds__(256)
ema_kernel(const EmaBatch t, float beta) {
  int ti = 0;
  while (ti + 1 < t.count && (int)blockIdx.x >= t.first_block[ti + 1]) ++ti;
  const long long base = (long long)(blockIdx.x - t.first_block[ti]) * kChunk;
  const long long end = min(t.n[ti], base + kChunk);
  float* __restrict__ ma = t.ma[ti];
  const float* __restrict__ cur = t.cur[ti];
  const float omb = 1.f - beta;
  const bool vec = (((uintptr_t)ma | (uintptr_t)cur) & 15) == 0;
  long long i = base + (vec ? threadIdx.x * 4 : threadIdx.x);
  if (vec) {
    for (; i + 3 < end; i += 256 * 4) {
      float4 m = *reinterpret_cast<const float4*>(ma + i);
      const float4 c = *reinterpret_cast<const float4*>(cur + i);
      // same expression as EMA.update_average (:69): old * beta + (1 - beta) * new
      m.x = m.x * beta + omb * c.x; m.y = m.y * beta + omb * c.y;
      m.z = m.z * beta + omb * c.z; m.w = m.w * beta + omb * c.w;
      *reinterpret_cast<float4*>(ma + i) = m;
    }
    if (i < end)
      for (long long j = i; j < end; ++j) ma[j] = ma[j] * beta + omb * cur[j];
  } else {
    for (; i < end; i += 256) ma[i] = ma[i] * beta + omb * cur[i];
  }
}

}  // namespace hg

using namespace hg;

extern "C" int hg_diffgrad_step(int32_t count, float* const* p, const float* const* g,
                                float* const* m, float* const* v, float* const* prev,
                                float* const* packed, const int64_t* numel, float beta1, float beta2,
                                float eps, float step_size, float weight_decay, hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (count < 0 || (count > 0 && (!p || !g || !m || !v || !prev || !numel)))
    return set_error(HG_EINVAL, "null pointer table");
  int i = 0;
  while (i < count) {
    DiffGradBatch b;
    b.count = 0;
    int blocks = 0;
    uintptr_t align = 0;
    while (i < count && b.count < kMaxTensors) {
      if (numel[i] > 0) {
        const int k = b.count++;
        b.p[k] = p[i]; b.g[k] = g[i]; b.m[k] = m[i]; b.v[k] = v[i]; b.prev[k] = prev[i];
        b.packed[k] = packed ? packed[i] : nullptr;
        align |= (uintptr_t)p[i] | (uintptr_t)g[i] | (uintptr_t)m[i] | (uintptr_t)v[i] |
                 (uintptr_t)prev[i] | (uintptr_t)b.packed[k];
        b.n[k] = numel[i];
        b.first_block[k] = blocks;
        blocks += (int)((numel[i] + kChunk - 1) / kChunk);
      }
      ++i;
    }
    b.first_block[b.count] = blocks;
    if (b.count == 0) continue;
    if (align & 15)
      diffgrad_kernel<false><<<blocks, 256, 0, stream>>>(b, beta1, beta2, eps, step_size, weight_decay);
    else
      diffgrad_kernel<true><<<blocks, 256, 0, stream>>>(b, beta1, beta2, eps, step_size, weight_decay);
    HG_LAUNCH_OK("diffgrad_kernel");
  }
  return 0;
}

extern "C" int hg_ema_update(int32_t count, float* const* ma, const float* const* cur,
                             const int64_t* numel, float beta, hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (count < 0 || (count > 0 && (!ma || !cur || !numel))) return set_error(HG_EINVAL, "null pointer table");
  int i = 0;
  while (i < count) {
    EmaBatch b;
    b.count = 0;
    int blocks = 0;
    while (i < count && b.count < kMaxTensors) {
      if (numel[i] > 0) {
        const int k = b.count++;
        b.ma[k] = ma[i]; b.cur[k] = cur[i]; b.n[k] = numel[i];
        b.first_block[k] = blocks;
        blocks += (int)((numel[i] + kChunk - 1) / kChunk);
      }
      ++i;
    }
    b.first_block[b.count] = blocks;
    if (b.count == 0) continue;
    ema_kernel<<<blocks, 256, 0, stream>>>(b, beta);
    HG_LAUNCH_OK("ema_kernel");
  }
  return 0;
}
