// Multi-tensor optimizer step: the pass over parameters, gradients and optimizer state that follows
// the backward in timm's training loop (train.py:1262-1269 -> torch.optim.SGD / torch.optim.AdamW created by
// timm/optim/_optim_factory.py:355-404, gradient clipping through timm/utils/clip_grad.py:18).
//
// One launch covers a whole parameter group: the per-tensor pointer table travels BY VALUE in the kernel's parameter
// space (40 bytes per tensor, up to kMaxTensors per launch; sm_100 accepts 32 KB of kernel parameters), so nothing is
// staged through device memory and fresh .grad pointers every step cost nothing.  A persistent grid walks 4096-element
// chunks; a chunk finds its tensor by binary search over the chunk prefix sums.  Parameters / gradients are bf16 or
// fp32, optimizer state is always fp32, arithmetic is fp32: 14 B (SGD) / 22 B (AdamW) of HBM traffic per bf16 element.
#include "common.h"
#include <math.h>
#include "ptx.cuh"

namespace b200 {

constexpr int kMaxTensors = 512;
constexpr int kChunk = 4096;
constexpr int kThreads = 256;

struct OptimArgs {
  b200_optim_tensor t[kMaxTensors];
  int chunk_start[kMaxTensors + 1];      // prefix sums of ceil(numel / kChunk)
  int n;
};
static_assert(sizeof(OptimArgs) <= 32000, "kernel parameter space");

struct SgdParams { float lr, momentum, one_minus_dampening, weight_decay; int nesterov, first_step; };
struct AdamParams { float one_minus_lr_wd, one_minus_beta1, beta2, one_minus_beta2, eps, weight_decay, step_size, bc2_sqrt; int decoupled; };

__device__ __forceinline__ int find_tensor(const OptimArgs& a, int chunk) {
  int lo = 0, hi = a.n;                  // largest t with chunk_start[t] <= chunk
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (a.chunk_start[mid] <= chunk) lo = mid; else hi = mid;
  }
  return lo;
}

template <bool F32>
__device__ __forceinline__ float ld1(const void* p, long long i) {
  return F32 ? reinterpret_cast<const float*>(p)[i] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
template <bool F32>
__device__ __forceinline__ void st1(void* p, long long i, float v) {
  if (F32) reinterpret_cast<float*>(p)[i] = v;
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}
// 8 consecutive elements from a 16-byte aligned address
template <bool F32>
__device__ __forceinline__ void ld8v(const void* p, long long i, float (&v)[8]) {
  if (F32) {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
    const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 q = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p) + i);
    v[0] = bf16_lo(q.x); v[1] = bf16_hi(q.x); v[2] = bf16_lo(q.y); v[3] = bf16_hi(q.y);
    v[4] = bf16_lo(q.z); v[5] = bf16_hi(q.z); v[6] = bf16_lo(q.w); v[7] = bf16_hi(q.w);
  }
}
template <bool F32>
__device__ __forceinline__ void st8v(void* p, long long i, const float (&v)[8]) {
  if (F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 q;
    q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]);
    q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p) + i) = q;
  }
}

// torch.optim.SGD (_single_tensor_sgd): g += wd*p; buf = first ? g : momentum*buf + (1-dampening)*g;
// g = nesterov ? g + momentum*buf : buf; p -= lr*g.  `scale` is the gradient-clipping coefficient (1 without clipping).
__device__ __forceinline__ void sgd_elem(float& p, float g, float& buf, const SgdParams& h, float scale, bool has_buf) {
  g *= scale;
  if (h.weight_decay != 0.f) g = fmaf(h.weight_decay, p, g);
  if (has_buf) {
    buf = h.first_step ? g : fmaf(h.momentum, buf, h.one_minus_dampening * g);
    g = h.nesterov ? fmaf(h.momentum, buf, g) : buf;
  }
  p = fmaf(-h.lr, g, p);
}

// torch.optim.AdamW / Adam (_single_tensor_adam): decoupled: p *= 1 - lr*wd, else g += wd*p;
// m = lerp(m, g, 1-beta1); v = beta2*v + (1-beta2)*g*g; p -= step_size * m / (sqrt(v)/sqrt(bc2) + eps).
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamParams& h, float scale) {
  g *= scale;
  if (h.weight_decay != 0.f) {
    if (h.decoupled) p *= h.one_minus_lr_wd;
    else g = fmaf(h.weight_decay, p, g);
  }
  m = fmaf(h.one_minus_beta1, g - m, m);
  v = fmaf(h.beta2, v, h.one_minus_beta2 * g * g);
  const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
  p = fmaf(-h.step_size, m / denom, p);
}

template <bool F32, bool ADAM>
__global__ void __launch_bounds__(kThreads) optim_step_kernel(const __grid_constant__ OptimArgs a, const SgdParams sg,
                                                              const AdamParams ad, const float* __restrict__ grad_scale) {
  const float scale = grad_scale != nullptr ? __ldg(grad_scale) : 1.f;
  const int total = a.chunk_start[a.n];
  for (int chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
    const int ti = find_tensor(a, chunk);
    const b200_optim_tensor& t = a.t[ti];
    const long long base = static_cast<long long>(chunk - a.chunk_start[ti]) * kChunk;
    const long long rem = t.numel - base;
    const int len = rem < kChunk ? static_cast<int>(rem) : kChunk;
    const bool has1 = t.state1 != nullptr;
    const uintptr_t bits = reinterpret_cast<uintptr_t>(t.param) | reinterpret_cast<uintptr_t>(t.grad) |
                           reinterpret_cast<uintptr_t>(t.state1) | reinterpret_cast<uintptr_t>(t.state2);
    const bool aligned = (bits & 15) == 0;      // chunk offsets are multiples of kChunk elements
    int done = 0;
    if (aligned) {
      const int vec = len >> 3;
      for (int i = threadIdx.x; i < vec; i += kThreads) {
        const long long off = base + i * 8;
        float p[8], g[8], s1[8], s2[8];
        ld8v<F32>(t.param, off, p);
        ld8v<F32>(t.grad, off, g);
        if (has1) ld8v<true>(t.state1, off, s1);
        if (ADAM) ld8v<true>(t.state2, off, s2);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (ADAM) adam_elem(p[k], g[k], s1[k], s2[k], ad, scale);
          else sgd_elem(p[k], g[k], s1[k], sg, scale, has1);
        }
        st8v<F32>(t.param, off, p);
        if (has1) st8v<true>(t.state1, off, s1);
        if (ADAM) st8v<true>(t.state2, off, s2);
      }
      done = vec << 3;
    }
    for (int i = done + threadIdx.x; i < len; i += kThreads) {
      const long long off = base + i;
      float p = ld1<F32>(t.param, off), g = ld1<F32>(t.grad, off);
      float s1 = has1 ? t.state1[off] : 0.f, s2 = ADAM ? t.state2[off] : 0.f;
      if (ADAM) adam_elem(p, g, s1, s2, ad, scale);
      else sgd_elem(p, g, s1, sg, scale, has1);
      st1<F32>(t.param, off, p);
      if (has1) t.state1[off] = s1;
      if (ADAM) t.state2[off] = s2;
    }
  }
}

// sum over every gradient element of g^2 (torch.nn.utils.clip_grad_norm_, norm_type 2), accumulated into *out.
template <bool F32>
__global__ void __launch_bounds__(kThreads) grad_sumsq_kernel(const __grid_constant__ OptimArgs a, float* __restrict__ out) {
  float acc = 0.f;
  const int total = a.chunk_start[a.n];
  for (int chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
    const int ti = find_tensor(a, chunk);
    const b200_optim_tensor& t = a.t[ti];
    const long long base = static_cast<long long>(chunk - a.chunk_start[ti]) * kChunk;
    const long long rem = t.numel - base;
    const int len = rem < kChunk ? static_cast<int>(rem) : kChunk;
    int done = 0;
    if ((reinterpret_cast<uintptr_t>(t.grad) & 15) == 0) {
      const int vec = len >> 3;
      for (int i = threadIdx.x; i < vec; i += kThreads) {
        float g[8];
        ld8v<F32>(t.grad, base + i * 8, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(g[k], g[k], acc);
      }
      done = vec << 3;
    }
    for (int i = done + threadIdx.x; i < len; i += kThreads) {
      const float g = ld1<F32>(t.grad, base + i);
      acc = fmaf(g, g, acc);
    }
  }
  __shared__ float part[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kThreads / 32 ? part[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0 && v != 0.f) atomicAdd(out, v);
  }
}

// clip_grad_norm_'s coefficient: total = sqrt(sumsq); coef = min(1, max_norm / (total + 1e-6))   (torch/nn/utils/clip_grad.py)
__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ coef,
                                 float* __restrict__ total_norm) {
  const float total = sqrtf(*sumsq);
  if (total_norm != nullptr) *total_norm = total;
  *coef = fminf(1.f, max_norm / (total + 1e-6f));
}

// in-place g *= *scale over every gradient of the table (the in-place form of clip_grad_norm_)
template <bool F32>
__global__ void __launch_bounds__(kThreads) grad_scale_kernel(const __grid_constant__ OptimArgs a, const float* __restrict__ scale_p) {
  const float scale = __ldg(scale_p);
  if (scale == 1.f) return;
  const int total = a.chunk_start[a.n];
  for (int chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
    const int ti = find_tensor(a, chunk);
    const b200_optim_tensor& t = a.t[ti];
    void* grad = const_cast<void*>(t.grad);
    const long long base = static_cast<long long>(chunk - a.chunk_start[ti]) * kChunk;
    const long long rem = t.numel - base;
    const int len = rem < kChunk ? static_cast<int>(rem) : kChunk;
    int done = 0;
    if ((reinterpret_cast<uintptr_t>(grad) & 15) == 0) {
      const int vec = len >> 3;
      for (int i = threadIdx.x; i < vec; i += kThreads) {
        float g[8];
        ld8v<F32>(grad, base + i * 8, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] *= scale;
        st8v<F32>(grad, base + i * 8, g);
      }
      done = vec << 3;
    }
    for (int i = done + threadIdx.x; i < len; i += kThreads) st1<F32>(grad, base + i, ld1<F32>(grad, base + i) * scale);
  }
}

// ModelEmaV3.apply_update_ (timm/utils/model_ema.py:213-231): ema.lerp_(model, weight) per tensor, torch's lerp formula
// (weight < 0.5: a + w*(b - a); else b - (b - a)*(1 - w)).  Table entries: param = EMA tensor, grad = model tensor.
template <bool EF32, bool MF32>
__global__ void __launch_bounds__(kThreads) ema_update_kernel(const __grid_constant__ OptimArgs a, float w) {
  const int total = a.chunk_start[a.n];
  const bool low = w < 0.5f;
  const float omw = 1.f - w;
  for (int chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
    const int ti = find_tensor(a, chunk);
    const b200_optim_tensor& t = a.t[ti];
    const long long base = static_cast<long long>(chunk - a.chunk_start[ti]) * kChunk;
    const long long rem = t.numel - base;
    const int len = rem < kChunk ? static_cast<int>(rem) : kChunk;
    int done = 0;
    if (((reinterpret_cast<uintptr_t>(t.param) | reinterpret_cast<uintptr_t>(t.grad)) & 15) == 0) {
      const int vec = len >> 3;
      for (int i = threadIdx.x; i < vec; i += kThreads) {
        const long long off = base + i * 8;
        float e[8], m[8];
        ld8v<EF32>(t.param, off, e);
        ld8v<MF32>(t.grad, off, m);
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = low ? fmaf(w, m[k] - e[k], e[k]) : fmaf(-(m[k] - e[k]), omw, m[k]);
        st8v<EF32>(t.param, off, e);
      }
      done = vec << 3;
    }
    for (int i = done + threadIdx.x; i < len; i += kThreads) {
      const float e = ld1<EF32>(t.param, base + i), m = ld1<MF32>(t.grad, base + i);
      st1<EF32>(t.param, base + i, low ? fmaf(w, m - e, e) : fmaf(-(m - e), omw, m));
    }
  }
}

// Fill `a` from tensors[first, first+count) and return the number of chunks.
static int fill_args(OptimArgs& a, const b200_optim_tensor* tensors, int first, int count) {
  a.n = count;
  int chunks = 0;
  for (int i = 0; i < count; ++i) {
    a.t[i] = tensors[first + i];
    a.chunk_start[i] = chunks;
    chunks += static_cast<int>((tensors[first + i].numel + kChunk - 1) / kChunk);
  }
  a.chunk_start[count] = chunks;
  return chunks;
}

// n_state: 0 = gradients only, 1 = parameter + optional state1, 2 = parameter + state1 + state2
static int check_table(const char* what, const b200_optim_tensor* tensors, int n, int dtype, int n_state) {
  B200_CHECK_ARG(tensors != nullptr && n > 0, "%s: empty tensor table", what);
  B200_CHECK_ARG(dtype == B200_DTYPE_BF16 || dtype == B200_DTYPE_F32, "%s: dtype must be bf16 or f32", what);
  const uintptr_t amask = dtype == B200_DTYPE_F32 ? 3 : 1;
  long long total_chunks = 0;
  for (int i = 0; i < n; ++i) {
    const b200_optim_tensor& t = tensors[i];
    B200_CHECK_ARG(t.grad != nullptr && t.numel > 0, "%s: tensor %d: null gradient or numel <= 0", what, i);
    B200_CHECK_ARG(n_state == 0 || t.param != nullptr, "%s: tensor %d: null parameter", what, i);
    B200_CHECK_ARG(n_state < 2 || (t.state1 != nullptr && t.state2 != nullptr), "%s: tensor %d: null state", what, i);
    B200_CHECK_ARG(((reinterpret_cast<uintptr_t>(t.param) | reinterpret_cast<uintptr_t>(t.grad)) & amask) == 0 &&
                       ((reinterpret_cast<uintptr_t>(t.state1) | reinterpret_cast<uintptr_t>(t.state2)) & 3) == 0,
                   "%s: tensor %d: misaligned pointer", what, i);
    total_chunks += (t.numel + kChunk - 1) / kChunk;
  }
  B200_CHECK_ARG(total_chunks < (1ll << 31), "%s: too many elements", what);
  return B200_OK;
}

static int grid_for(int chunks) {
  const int cap = sm_count() * 8;
  return chunks < cap ? chunks : cap;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_sgd_step(const b200_optim_tensor* tensors, int32_t n, int32_t dtype, double lr, double momentum,
                             double dampening, double weight_decay, int32_t nesterov, int32_t first_step,
                             const float* grad_scale, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_table("b200_sgd_step", tensors, n, dtype, 1)) return rc;
  B200_CHECK_ARG(!nesterov || (momentum > 0.0 && dampening == 0.0),
                 "b200_sgd_step: Nesterov momentum requires a momentum and zero dampening");
  for (int i = 0; i < n; ++i)
    B200_CHECK_ARG((momentum != 0.0) == (tensors[i].state1 != nullptr),
                   "b200_sgd_step: tensor %d: momentum buffer must be given exactly when momentum != 0", i);
  SgdParams sg{static_cast<float>(lr), static_cast<float>(momentum), static_cast<float>(1.0 - dampening),
               static_cast<float>(weight_decay), nesterov != 0, first_step != 0};
  AdamParams ad{};
  OptimArgs a;
  for (int first = 0; first < n; first += kMaxTensors) {
    const int count = n - first < kMaxTensors ? n - first : kMaxTensors;
    const int chunks = fill_args(a, tensors, first, count);
    if (dtype == B200_DTYPE_F32) optim_step_kernel<true, false><<<grid_for(chunks), kThreads, 0, st>>>(a, sg, ad, grad_scale);
    else optim_step_kernel<false, false><<<grid_for(chunks), kThreads, 0, st>>>(a, sg, ad, grad_scale);
    B200_CUDA_OK(cudaGetLastError());
  }
  return B200_OK;
}

extern "C" int b200_adam_step(const b200_optim_tensor* tensors, int32_t n, int32_t dtype, double lr, double beta1,
                              double beta2, double eps, double weight_decay, int32_t step, int32_t decoupled,
                              const float* grad_scale, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_table("b200_adam_step", tensors, n, dtype, 2)) return rc;
  B200_CHECK_ARG(step >= 1, "b200_adam_step: step counts from 1");
  B200_CHECK_ARG(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0, "b200_adam_step: bad betas / eps");
  // scalar prework in double, as the Python floats of torch/optim/adam.py (_single_tensor_adam)
  const double bc1 = 1.0 - pow(beta1, step);
  const double bc2 = 1.0 - pow(beta2, step);
  AdamParams ad{static_cast<float>(1.0 - lr * weight_decay), static_cast<float>(1.0 - beta1), static_cast<float>(beta2),
                static_cast<float>(1.0 - beta2), static_cast<float>(eps), static_cast<float>(weight_decay),
                static_cast<float>(lr / bc1), static_cast<float>(sqrt(bc2)), decoupled != 0};
  SgdParams sg{};
  OptimArgs a;
  for (int first = 0; first < n; first += kMaxTensors) {
    const int count = n - first < kMaxTensors ? n - first : kMaxTensors;
    const int chunks = fill_args(a, tensors, first, count);
    if (dtype == B200_DTYPE_F32) optim_step_kernel<true, true><<<grid_for(chunks), kThreads, 0, st>>>(a, sg, ad, grad_scale);
    else optim_step_kernel<false, true><<<grid_for(chunks), kThreads, 0, st>>>(a, sg, ad, grad_scale);
    B200_CUDA_OK(cudaGetLastError());
  }
  return B200_OK;
}

extern "C" int b200_grad_sumsq(const b200_optim_tensor* tensors, int32_t n, int32_t dtype, float* sumsq, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_table("b200_grad_sumsq", tensors, n, dtype, 0)) return rc;
  B200_CHECK_ARG(sumsq != nullptr, "b200_grad_sumsq: null output");
  OptimArgs a;
  for (int first = 0; first < n; first += kMaxTensors) {
    const int count = n - first < kMaxTensors ? n - first : kMaxTensors;
    const int chunks = fill_args(a, tensors, first, count);
    if (dtype == B200_DTYPE_F32) grad_sumsq_kernel<true><<<grid_for(chunks), kThreads, 0, st>>>(a, sumsq);
    else grad_sumsq_kernel<false><<<grid_for(chunks), kThreads, 0, st>>>(a, sumsq);
    B200_CUDA_OK(cudaGetLastError());
  }
  return B200_OK;
}

extern "C" int b200_clip_coef(const float* sumsq, double max_norm, float* coef, float* total_norm, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(sumsq != nullptr && coef != nullptr && max_norm >= 0.0, "b200_clip_coef: bad arguments");
  clip_coef_kernel<<<1, 1, 0, st>>>(sumsq, static_cast<float>(max_norm), coef, total_norm);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_grad_scale(const b200_optim_tensor* tensors, int32_t n, int32_t dtype, const float* scale, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_table("b200_grad_scale", tensors, n, dtype, 0)) return rc;
  B200_CHECK_ARG(scale != nullptr, "b200_grad_scale: null scale");
  OptimArgs a;
  for (int first = 0; first < n; first += kMaxTensors) {
    const int count = n - first < kMaxTensors ? n - first : kMaxTensors;
    const int chunks = fill_args(a, tensors, first, count);
    if (dtype == B200_DTYPE_F32) grad_scale_kernel<true><<<grid_for(chunks), kThreads, 0, st>>>(a, scale);
    else grad_scale_kernel<false><<<grid_for(chunks), kThreads, 0, st>>>(a, scale);
    B200_CUDA_OK(cudaGetLastError());
  }
  return B200_OK;
}

extern "C" int b200_ema_update(const b200_optim_tensor* tensors, int32_t n, int32_t ema_dtype, int32_t model_dtype,
                               double weight, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(tensors != nullptr && n > 0, "b200_ema_update: empty tensor table");
  B200_CHECK_ARG((ema_dtype == B200_DTYPE_BF16 || ema_dtype == B200_DTYPE_F32) &&
                     (model_dtype == B200_DTYPE_BF16 || model_dtype == B200_DTYPE_F32), "b200_ema_update: dtype must be bf16 or f32");
  B200_CHECK_ARG(weight >= 0.0 && weight <= 1.0, "b200_ema_update: weight (1 - decay) must be in [0, 1]");
  long long total_chunks = 0;
  for (int i = 0; i < n; ++i) {
    const b200_optim_tensor& t = tensors[i];
    B200_CHECK_ARG(t.param != nullptr && t.grad != nullptr && t.numel > 0, "b200_ema_update: tensor %d: null pointer or numel <= 0", i);
    B200_CHECK_ARG((reinterpret_cast<uintptr_t>(t.param) & (ema_dtype == B200_DTYPE_F32 ? 3 : 1)) == 0 &&
                       (reinterpret_cast<uintptr_t>(t.grad) & (model_dtype == B200_DTYPE_F32 ? 3 : 1)) == 0,
                   "b200_ema_update: tensor %d: misaligned pointer", i);
    total_chunks += (t.numel + kChunk - 1) / kChunk;
  }
  B200_CHECK_ARG(total_chunks < (1ll << 31), "b200_ema_update: too many elements");
  const float w = static_cast<float>(weight);
  const bool ef = ema_dtype == B200_DTYPE_F32, mf = model_dtype == B200_DTYPE_F32;
  OptimArgs a;
  for (int first = 0; first < n; first += kMaxTensors) {
    const int count = n - first < kMaxTensors ? n - first : kMaxTensors;
    const int g = grid_for(fill_args(a, tensors, first, count));
    if (ef && mf) ema_update_kernel<true, true><<<g, kThreads, 0, st>>>(a, w);
    else if (ef) ema_update_kernel<true, false><<<g, kThreads, 0, st>>>(a, w);
    else if (mf) ema_update_kernel<false, true><<<g, kThreads, 0, st>>>(a, w);
    else ema_update_kernel<false, false><<<g, kThreads, 0, st>>>(a, w);
    B200_CUDA_OK(cudaGetLastError());
  }
  return B200_OK;
}
