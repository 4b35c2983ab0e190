// optim.cu -- fused multi-tensor DiffGrad step.
//
// The reference steps torch_optimizer.DiffGrad (histoGAN/histoGAN.py:670-671,932,989),
// a per-parameter Python loop of ~10 element-wise launches.  One pass here:
//   m = b1 m + (1-b1) g ;  v = b2 v + (1-b2) g^2 ;  xi = sigmoid(|g_prev - g|)
//   p -= step_size * (m * xi) / (sqrt(v) + eps) ;   g_prev = g
// for up to kMaxTensors tensors per launch (pointers in kernel-parameter space).
#include "hg_common.cuh"

namespace hg {

constexpr int kMaxTensors = 48;
constexpr int kChunk = 1 << 16;            // elements per CTA

struct DiffGradBatch {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];
  float* m[kMaxTensors];
  float* v[kMaxTensors];
  float* prev[kMaxTensors];
  long long n[kMaxTensors];
  int first_block[kMaxTensors + 1];        // prefix sum of ceil(n / kChunk)
  int count;
};

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
  const long long end = min(n, base + kChunk);
  for (long long i = base + threadIdx.x; i < end; i += 256) {
    float gi = g[i];
    const float pi = p[i];
    if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
    const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);
    const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
    const float diff = fabsf(pv[i] - gi);
    const float xi = 1.f / (1.f + expf(-diff));
    m[i] = mi; v[i] = vi; pv[i] = gi;
    p[i] = pi - step_size * (mi * xi) / (sqrtf(vi) + eps);
  }
}

}  // namespace hg

using namespace hg;

extern "C" int hg_diffgrad_step(int32_t count, float* const* p, const float* const* g,
                                float* const* m, float* const* v, float* const* prev,
                                const int64_t* numel, float beta1, float beta2, float eps,
                                float step_size, float weight_decay, hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (count < 0 || (count > 0 && (!p || !g || !m || !v || !prev || !numel)))
    return set_error(HG_EINVAL, "null pointer table");
  int i = 0;
  while (i < count) {
    DiffGradBatch b;
    b.count = 0;
    int blocks = 0;
    while (i < count && b.count < kMaxTensors) {
      if (numel[i] > 0) {
        const int k = b.count++;
        b.p[k] = p[i]; b.g[k] = g[i]; b.m[k] = m[i]; b.v[k] = v[i]; b.prev[k] = prev[i];
        b.n[k] = numel[i];
        b.first_block[k] = blocks;
        blocks += (int)((numel[i] + kChunk - 1) / kChunk);
      }
      ++i;
    }
    b.first_block[b.count] = blocks;
    if (b.count == 0) continue;
    diffgrad_kernel<<<blocks, 256, 0, stream>>>(b, beta1, beta2, eps, step_size, weight_decay);
    HG_LAUNCH_OK("diffgrad_kernel");
  }
  return 0;
}
