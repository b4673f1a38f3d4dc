// Classification losses on the logits that leave the head GEMM: nn.CrossEntropyLoss (train.py:855,
// benchmark.py:380), LabelSmoothingCrossEntropy and SoftTargetCrossEntropy (timm/loss/cross_entropy.py:11-36), forward
// and backward, 'mean' reduction.  The data is tiny ([256, 1000]) -- torch spends ~10 launches on cast + log_softmax +
// gather / mean + their backwards; here it is one row-parallel pass forward (+ a one-block mean) and one pass backward
// that recomputes softmax from the saved log-sum-exp.  One 256-thread block per row, fp32 arithmetic.
#include "common.h"
#include "ptx.cuh"

namespace b200 {

constexpr int kLossThreads = 256;

template <bool F32>
__device__ __forceinline__ float ldx(const void* p, long long i) {
  return F32 ? __ldg(reinterpret_cast<const float*>(p) + i)
             : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, u) : v + u;
  }
  __syncthreads();                              // scratch may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = scratch[0];
#pragma unroll
  for (int i = 1; i < kLossThreads / 32; ++i) r = is_max ? fmaxf(r, scratch[i]) : r + scratch[i];
  return r;
}

// hard targets: loss = (1 - s) * (lse - x[t]) + s * (lse - mean(x));  soft targets: loss = lse * sum(t) - sum(t * x)
template <bool XF32, bool TF32>
__global__ void __launch_bounds__(kLossThreads) ce_fwd_kernel(const void* __restrict__ logits, const long long* __restrict__ target,
                                                              const void* __restrict__ soft, float smoothing,
                                                              long long ignore_index, float* __restrict__ loss_rows,
                                                              float* __restrict__ lse_out, float* __restrict__ tsum_out, int C) {
  __shared__ float scratch[kLossThreads / 32];
  const long long row = blockIdx.x;
  const long long base = row * C;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += kLossThreads) mx = fmaxf(mx, ldx<XF32>(logits, base + c));
  mx = block_reduce(mx, true, scratch);
  float se = 0.f, sx = 0.f, st = 0.f, stx = 0.f;
  for (int c = threadIdx.x; c < C; c += kLossThreads) {
    const float x = ldx<XF32>(logits, base + c);
    se += expf(x - mx);
    sx += x;
    if (soft != nullptr) {
      const float t = ldx<TF32>(soft, base + c);
      st += t;
      stx = fmaf(t, x, stx);
    }
  }
  se = block_reduce(se, false, scratch);
  const float lse = mx + logf(se);
  float loss;
  if (soft != nullptr) {
    st = block_reduce(st, false, scratch);
    stx = block_reduce(stx, false, scratch);
    loss = lse * st - stx;
  } else {
    sx = block_reduce(sx, false, scratch);
    const long long t = target[row];
    st = 1.f;
    if (t == ignore_index || t < 0 || t >= C) {
      loss = 0.f;                                // ignore_index (or an out-of-range class index): the row contributes nothing
      st = 0.f;
    } else {
      const float xt = ldx<XF32>(logits, base + t);
      loss = (1.f - smoothing) * (lse - xt) + smoothing * (lse - sx / static_cast<float>(C));
    }
  }
  if (threadIdx.x == 0) {
    loss_rows[row] = loss;
    lse_out[row] = lse;
    tsum_out[row] = st;
  }
}

// out[0] = mean loss, out[1] = the denominator of the mean: the number of non-ignored rows for class-index targets
// (torch's 'mean' reduction of nll_loss divides by the count of rows whose target != ignore_index; all rows ignored ->
// 0 / 0 = NaN, as in torch), B for soft targets.
__global__ void __launch_bounds__(kLossThreads) ce_mean_kernel(const float* __restrict__ loss_rows, const float* __restrict__ tsum,
                                                               int hard, float* __restrict__ out, int B) {
  __shared__ float scratch[kLossThreads / 32];
  float s = 0.f, n = 0.f;
  for (int i = threadIdx.x; i < B; i += kLossThreads) {
    s += loss_rows[i];
    n += tsum[i];
  }
  s = block_reduce(s, false, scratch);
  n = hard ? block_reduce(n, false, scratch) : static_cast<float>(B);
  if (threadIdx.x == 0) {
    out[0] = s / n;
    out[1] = n;
  }
}

// d logits[b, c] = (softmax * sum(t) - target distribution) * grad_out / denom
template <bool XF32, bool TF32>
__global__ void __launch_bounds__(kLossThreads) ce_bwd_kernel(const void* __restrict__ logits, const long long* __restrict__ target,
                                                              const void* __restrict__ soft, float smoothing,
                                                              const float* __restrict__ lse, const float* __restrict__ tsum,
                                                              const float* __restrict__ grad_out, const float* __restrict__ denom,
                                                              void* __restrict__ dlogits, int B, int C) {
  const long long row = blockIdx.x;
  const long long base = row * C;
  const float g = __ldg(grad_out) / __ldg(denom);
  const float l = lse[row], ts = tsum[row];
  const long long t = soft != nullptr ? -1 : target[row];
  const float uni = smoothing / static_cast<float>(C);
  for (int c = threadIdx.x; c < C; c += kLossThreads) {
    const float p = expf(ldx<XF32>(logits, base + c) - l) * ts;
    float q;
    if (soft != nullptr) q = ldx<TF32>(soft, base + c);
    else q = ts == 0.f ? 0.f : (uni + (c == t ? 1.f - smoothing : 0.f));
    const float d = (p - q) * g;
    if (XF32) reinterpret_cast<float*>(dlogits)[base + c] = d;
    else reinterpret_cast<__nv_bfloat16*>(dlogits)[base + c] = __float2bfloat16_rn(d);
  }
}

static int check_ce(const char* what, const void* logits, int dtype, const int64_t* target, const void* soft, int soft_dtype,
                    double smoothing, int B, int C) {
  B200_CHECK_ARG(logits != nullptr && B > 0 && C > 0, "%s: null logits or empty shape", what);
  B200_CHECK_ARG(dtype == B200_DTYPE_BF16 || dtype == B200_DTYPE_F32, "%s: logits must be bf16 or f32", what);
  B200_CHECK_ARG((target != nullptr) != (soft != nullptr), "%s: give class indices OR a soft target distribution", what);
  B200_CHECK_ARG(soft == nullptr || soft_dtype == B200_DTYPE_BF16 || soft_dtype == B200_DTYPE_F32, "%s: soft targets must be bf16 or f32", what);
  B200_CHECK_ARG(smoothing >= 0.0 && smoothing < 1.0 && (soft == nullptr || smoothing == 0.0),
                 "%s: smoothing must be in [0, 1) and 0 with soft targets", what);
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_cross_entropy_fwd(const void* logits, int32_t logits_dtype, const int64_t* target, const void* soft_target,
                                      int32_t soft_dtype, double smoothing, int64_t ignore_index, float* loss_rows, float* lse,
                                      float* tsum, float* loss_mean, int32_t B, int32_t C, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_ce("b200_cross_entropy_fwd", logits, logits_dtype, target, soft_target, soft_dtype, smoothing, B, C)) return rc;
  B200_CHECK_ARG(loss_rows && lse && tsum && loss_mean, "b200_cross_entropy_fwd: null output");
  const float s = static_cast<float>(smoothing);
  const long long ig = static_cast<long long>(ignore_index);
  const long long* tg = reinterpret_cast<const long long*>(target);
  const bool xf = logits_dtype == B200_DTYPE_F32, tf = soft_dtype == B200_DTYPE_F32;
  if (xf && tf) ce_fwd_kernel<true, true><<<B, kLossThreads, 0, st>>>(logits, tg, soft_target, s, ig, loss_rows, lse, tsum, C);
  else if (xf) ce_fwd_kernel<true, false><<<B, kLossThreads, 0, st>>>(logits, tg, soft_target, s, ig, loss_rows, lse, tsum, C);
  else if (tf) ce_fwd_kernel<false, true><<<B, kLossThreads, 0, st>>>(logits, tg, soft_target, s, ig, loss_rows, lse, tsum, C);
  else ce_fwd_kernel<false, false><<<B, kLossThreads, 0, st>>>(logits, tg, soft_target, s, ig, loss_rows, lse, tsum, C);
  B200_CUDA_OK(cudaGetLastError());
  ce_mean_kernel<<<1, kLossThreads, 0, st>>>(loss_rows, tsum, target != nullptr ? 1 : 0, loss_mean, B);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_cross_entropy_bwd(const void* logits, int32_t logits_dtype, const int64_t* target, const void* soft_target,
                                      int32_t soft_dtype, double smoothing, const float* lse, const float* tsum,
                                      const float* grad_out, const float* denom, void* dlogits, int32_t B, int32_t C,
                                      void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_ce("b200_cross_entropy_bwd", logits, logits_dtype, target, soft_target, soft_dtype, smoothing, B, C)) return rc;
  B200_CHECK_ARG(lse && tsum && grad_out && denom && dlogits, "b200_cross_entropy_bwd: null pointer");
  const float s = static_cast<float>(smoothing);
  const long long* tg = reinterpret_cast<const long long*>(target);
  const bool xf = logits_dtype == B200_DTYPE_F32, tf = soft_dtype == B200_DTYPE_F32;
  if (xf && tf) ce_bwd_kernel<true, true><<<B, kLossThreads, 0, st>>>(logits, tg, soft_target, s, lse, tsum, grad_out, denom, dlogits, B, C);
  else if (xf) ce_bwd_kernel<true, false><<<B, kLossThreads, 0, st>>>(logits, tg, soft_target, s, lse, tsum, grad_out, denom, dlogits, B, C);
  else if (tf) ce_bwd_kernel<false, true><<<B, kLossThreads, 0, st>>>(logits, tg, soft_target, s, lse, tsum, grad_out, denom, dlogits, B, C);
  else ce_bwd_kernel<false, false><<<B, kLossThreads, 0, st>>>(logits, tg, soft_target, s, lse, tsum, grad_out, denom, dlogits, B, C);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}
