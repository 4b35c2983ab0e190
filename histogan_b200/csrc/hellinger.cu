// hellinger.cu -- Hellinger histogram loss (histoGAN/histoGAN.py:957-960,
// ReHistoGAN/rehistoGAN.py:1011-1014):
//   loss = alpha * (1/sqrt 2) * sqrt( sum_{b,c,u,v} (sqrt T - sqrt H)^2 ) / B
// One global sqrt over the whole micro-batch.  Single-CTA
// kernels: the tensors are B*3*64*64 floats (1.5 MB at B=32) -- latency-, not
// bandwidth-bound; partial sums are combined in a fixed order, so the result is
// deterministic.
#include "hg_common.cuh"

namespace hg {

constexpr float kScale = 0.70710678118654752440f;   // SCALE = 1/np.sqrt(2.0)

constexpr int kHellThreads = 256;
constexpr int kHellMaxBlocks = 296;

// stage 1: per-CTA partial sums; the last CTA to finish adds them in a fixed order
// (deterministic) and writes q and loss.
__global__ void __launch_bounds__(kHellThreads)
hellinger_fwd_kernel(const float* __restrict__ target, const float* __restrict__ hist,
                     const long long n, const float coef, float* __restrict__ loss,
                     float* __restrict__ q, float* __restrict__ partials,
                     unsigned int* __restrict__ counter) {
  __shared__ float red[kHellThreads / 32];
  __shared__ bool is_last;
  float local = 0.f;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    const float d = __fadd_rn(__fsqrt_rn(target[e]), -__fsqrt_rn(hist[e]));
    local = fmaf(d, d, local);
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kHellThreads / 32; ++w) v += red[w];
    partials[blockIdx.x] = v;
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    __threadfence();
    float v = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) v += partials[i];
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      *q = v;
      *loss = coef * __fsqrt_rn(v);
      *counter = 0u;
    }
  }
}

// d loss / d H = -coef * (sqrtT - sqrtH) / (2 sqrtQ sqrtH) ; d loss / d T mirrors it.
__global__ void __launch_bounds__(256)
hellinger_bwd_kernel(const float* __restrict__ target, const float* __restrict__ hist,
                     const long long n, const float coef, const float* __restrict__ q,
                     const float* __restrict__ grad_loss, float* __restrict__ grad_hist,
                     float* __restrict__ grad_target) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float gl = grad_loss ? grad_loss[0] : 1.f;
  const float k = gl * coef / (2.f * __fsqrt_rn(q[0]));
  const float st = __fsqrt_rn(target[e]), sh = __fsqrt_rn(hist[e]);
  const float d = st - sh;
  if (grad_hist) grad_hist[e] = -k * d / sh;
  if (grad_target) grad_target[e] = k * d / st;
}

}  // namespace hg

using namespace hg;

extern "C" size_t hg_hellinger_workspace_bytes(void) {
  return sizeof(float) * kHellMaxBlocks + 256;
}

extern "C" int hg_hellinger_fwd(const float* target, const float* hist, int64_t numel, int32_t B,
                                float alpha, float* loss, float* q, void* ws, size_t ws_bytes,
                                hg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!target || !hist || !loss || !q) return set_error(HG_EINVAL, "null tensor pointer");
  if (B <= 0 || numel <= 0) return set_error(HG_EINVAL, "empty batch");
  if (!ws || ws_bytes < hg_hellinger_workspace_bytes())
    return set_error(HG_EWS, "workspace too small: %zu < %zu", ws_bytes, hg_hellinger_workspace_bytes());
  const float coef = alpha * kScale / (float)B;
  float* partials = (float*)ws;
  unsigned int* counter = (unsigned int*)((char*)ws + sizeof(float) * kHellMaxBlocks);
  HG_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
  long long blocks = (numel + 4 * kHellThreads - 1) / (4 * kHellThreads);
  if (blocks > kHellMaxBlocks) blocks = kHellMaxBlocks;
  hellinger_fwd_kernel<<<(unsigned)blocks, kHellThreads, 0, stream>>>(target, hist, numel, coef, loss,
                                                                    q, partials, counter);
  HG_LAUNCH_OK("hellinger_fwd_kernel");
  return 0;
}

extern "C" int hg_hellinger_bwd(const float* target, const float* hist, int64_t numel, int32_t B,
                                float alpha, const float* q, const float* grad_loss,
                                float* grad_hist, float* grad_target, hg_stream_t stream_) {
  if (!target || !hist || !q) return set_error(HG_EINVAL, "null tensor pointer");
  if (B <= 0 || numel <= 0) return set_error(HG_EINVAL, "empty batch");
  const float coef = alpha * kScale / (float)B;
  hellinger_bwd_kernel<<<(unsigned)((numel + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      target, hist, numel, coef, q, grad_loss, grad_hist, grad_target);
  HG_LAUNCH_OK("hellinger_bwd_kernel");
  return 0;
}
