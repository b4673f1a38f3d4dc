// Per-head normalisation of q and k between the QKV projection and attention: `q, k = self.q_norm(q), self.k_norm(k)`
// with norm_layer(head_dim) = LayerNorm or RmsNorm, timm/layers/attention.py:101-102,117 (head_dim 64).  HBM-bound:
// the qkv buffer [T, 3, H, 64] is normalised IN PLACE (v untouched); training also writes the pre-norm q / k rows to a
// side buffer [T, 2, H, 64] (the attention backward needs the normalised q / k, this backward the raw ones) and the
// row statistics.  One row = 64 bf16 = 8 threads x 16 bytes; statistics through 3 shuffles inside the 8-lane group.
#include "common.h"
#include "ptx.cuh"

namespace b200 {

constexpr int kQkThreads = 256;

__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

__device__ __forceinline__ void ldp8(const void* p, bool f32, int off, float (&v)[8]) {
  if (p == nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
  } else if (f32) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + off));
    const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + off + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p) + off));
    v[0] = bf16_lo(q.x); v[1] = bf16_hi(q.x); v[2] = bf16_lo(q.y); v[3] = bf16_hi(q.y);
    v[4] = bf16_lo(q.z); v[5] = bf16_hi(q.z); v[6] = bf16_lo(q.w); v[7] = bf16_hi(q.w);
  }
}

__device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
  v[0] = bf16_lo(q.x); v[1] = bf16_hi(q.x); v[2] = bf16_lo(q.y); v[3] = bf16_hi(q.y);
  v[4] = bf16_lo(q.z); v[5] = bf16_hi(q.z); v[6] = bf16_lo(q.w); v[7] = bf16_hi(q.w);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 q;
  q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]);
  q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
  return q;
}

// grid.y = 0 (q rows) / 1 (k rows); row r of a kind = (token t, head h), r = t * H + h
template <bool RMS>
__global__ void __launch_bounds__(kQkThreads) qk_norm_fwd_kernel(__nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ save,
                                                                 float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                                 const void* __restrict__ qw, const void* __restrict__ qb,
                                                                 const void* __restrict__ kw, const void* __restrict__ kb,
                                                                 int param_f32, long long T, int H, float eps) {
  const int s = blockIdx.y;
  const int sub = threadIdx.x & 7;
  float w[8], b[8];
  ldp8(s ? kw : qw, param_f32, sub * 8, w);
  ldp8(s ? kb : qb, param_f32, sub * 8, b);
  const long long rows = T * H;
  // the loop condition is warp-uniform (the shuffles below need all 32 lanes): rows past the end are clamped and masked
  for (long long r0 = static_cast<long long>(blockIdx.x) * (kQkThreads / 8) + (threadIdx.x >> 5) * 4; r0 < rows;
       r0 += static_cast<long long>(gridDim.x) * (kQkThreads / 8)) {
    const bool valid = r0 + ((threadIdx.x & 31) >> 3) < rows;
    const long long r = valid ? r0 + ((threadIdx.x & 31) >> 3) : rows - 1;
    const long long t = r / H;
    const int h = static_cast<int>(r - t * H);
    __nv_bfloat16* row = qkv + (t * 3 + s) * H * 64 + h * 64 + sub * 8;
    const uint4 raw = *reinterpret_cast<const uint4*>(row);
    float x[8];
    unpack8(raw, x);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += x[i];
    const float mean = RMS ? 0.f : group8_sum(sum) * (1.f / 64.f);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sq = fmaf(x[i] - mean, x[i] - mean, sq);
    const float rstd = rsqrtf(group8_sum(sq) * (1.f / 64.f) + eps);
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = fmaf((x[i] - mean) * rstd, w[i], b[i]);
    if (!valid) continue;
    *reinterpret_cast<uint4*>(row) = pack8(y);
    const long long srow = (t * 2 + s) * H + h;
    if (save != nullptr) *reinterpret_cast<uint4*>(save + srow * 64 + sub * 8) = raw;
    if (rstd_out != nullptr && sub == 0) {
      rstd_out[srow] = rstd;
      if (!RMS) mean_out[srow] = mean;
    }
  }
}

// In place on dqkv: d(q), d(k) w.r.t. the normalised rows -> w.r.t. the raw rows.  dparams: fp32 [4][64] =
// dq_weight, dq_bias, dk_weight, dk_bias, accumulated with atomics (zero it first).
template <bool RMS>
__global__ void __launch_bounds__(kQkThreads) qk_norm_bwd_kernel(__nv_bfloat16* __restrict__ dqkv, const __nv_bfloat16* __restrict__ save,
                                                                 const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                                 const void* __restrict__ qw, const void* __restrict__ kw,
                                                                 int param_f32, float* __restrict__ dparams, long long T, int H) {
  __shared__ float acc[128];
  const int s = blockIdx.y;
  const int sub = threadIdx.x & 7;
  if (threadIdx.x < 128) acc[threadIdx.x] = 0.f;
  __syncthreads();
  float w[8];
  ldp8(s ? kw : qw, param_f32, sub * 8, w);
  float dw[8], db[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) dw[i] = db[i] = 0.f;
  const long long rows = T * H;
  // the loop condition is warp-uniform (the shuffles below need all 32 lanes): rows past the end are clamped and masked
  for (long long r0 = static_cast<long long>(blockIdx.x) * (kQkThreads / 8) + (threadIdx.x >> 5) * 4; r0 < rows;
       r0 += static_cast<long long>(gridDim.x) * (kQkThreads / 8)) {
    const bool valid = r0 + ((threadIdx.x & 31) >> 3) < rows;
    const long long r = valid ? r0 + ((threadIdx.x & 31) >> 3) : rows - 1;
    const long long t = r / H;
    const int h = static_cast<int>(r - t * H);
    __nv_bfloat16* drow = dqkv + (t * 3 + s) * H * 64 + h * 64 + sub * 8;
    const long long srow = (t * 2 + s) * H + h;
    float dy[8], x[8];
    unpack8(*reinterpret_cast<const uint4*>(drow), dy);
    unpack8(*reinterpret_cast<const uint4*>(save + srow * 64 + sub * 8), x);
    const float rstd = rstd_in[srow];
    const float mean = RMS ? 0.f : mean_in[srow];
    float c1 = 0.f, c2 = 0.f, xh[8], g[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      xh[i] = (x[i] - mean) * rstd;
      g[i] = dy[i] * w[i];
      c1 += g[i];
      c2 = fmaf(g[i], xh[i], c2);
      if (valid) {
        dw[i] = fmaf(dy[i], xh[i], dw[i]);
        db[i] += dy[i];
      }
    }
    c1 = RMS ? 0.f : group8_sum(c1) * (1.f / 64.f);
    c2 = group8_sum(c2) * (1.f / 64.f);
    float dx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dx[i] = rstd * (g[i] - c1 - xh[i] * c2);
    if (valid) *reinterpret_cast<uint4*>(drow) = pack8(dx);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    atomicAdd(&acc[sub * 8 + i], dw[i]);
    atomicAdd(&acc[64 + sub * 8 + i], db[i]);
  }
  __syncthreads();
  if (threadIdx.x < 128 && acc[threadIdx.x] != 0.f) atomicAdd(dparams + s * 128 + threadIdx.x, acc[threadIdx.x]);
}

static int qk_grid(long long rows) {
  long long blocks = (rows + kQkThreads / 8 - 1) / (kQkThreads / 8);
  const long long cap = static_cast<long long>(sm_count()) * 8;
  return static_cast<int>(blocks < cap ? blocks : cap);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_qk_norm_fwd(void* qkv, void* save, float* mean, float* rstd, const void* q_weight, const void* q_bias,
                                const void* k_weight, const void* k_bias, int32_t param_dtype, int64_t T, int32_t H,
                                int32_t head_dim, float eps, int32_t rms, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(qkv && q_weight && k_weight && T > 0 && H > 0, "b200_qk_norm_fwd: null pointer or empty shape");
  B200_CHECK_ARG(head_dim == 64, "b200_qk_norm_fwd: head_dim %d is not supported (64 only)", head_dim);
  B200_CHECK_ARG(param_dtype == B200_DTYPE_BF16 || param_dtype == B200_DTYPE_F32, "b200_qk_norm_fwd: bad parameter dtype");
  B200_CHECK_ARG(rms ? (q_bias == nullptr && k_bias == nullptr) : true, "b200_qk_norm_fwd: RMSNorm has no bias");
  B200_CHECK_ARG((save == nullptr) == (rstd == nullptr) && (rms || (rstd == nullptr) == (mean == nullptr)),
                 "b200_qk_norm_fwd: save / mean / rstd are given together (training) or not at all");
  dim3 grid(qk_grid(static_cast<long long>(T) * H), 2);
  auto* p = reinterpret_cast<__nv_bfloat16*>(qkv);
  auto* sv = reinterpret_cast<__nv_bfloat16*>(save);
  const int pf = param_dtype == B200_DTYPE_F32;
  if (rms) qk_norm_fwd_kernel<true><<<grid, kQkThreads, 0, st>>>(p, sv, mean, rstd, q_weight, q_bias, k_weight, k_bias, pf, T, H, eps);
  else qk_norm_fwd_kernel<false><<<grid, kQkThreads, 0, st>>>(p, sv, mean, rstd, q_weight, q_bias, k_weight, k_bias, pf, T, H, eps);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_qk_norm_bwd(void* dqkv, const void* save, const float* mean, const float* rstd, const void* q_weight,
                                const void* k_weight, int32_t param_dtype, float* dparams, int64_t T, int32_t H,
                                int32_t head_dim, int32_t rms, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(dqkv && save && rstd && q_weight && k_weight && dparams && T > 0 && H > 0,
                 "b200_qk_norm_bwd: null pointer or empty shape");
  B200_CHECK_ARG(rms || mean != nullptr, "b200_qk_norm_bwd: LayerNorm needs the saved means");
  B200_CHECK_ARG(head_dim == 64, "b200_qk_norm_bwd: head_dim %d is not supported (64 only)", head_dim);
  B200_CHECK_ARG(param_dtype == B200_DTYPE_BF16 || param_dtype == B200_DTYPE_F32, "b200_qk_norm_bwd: bad parameter dtype");
  dim3 grid(qk_grid(static_cast<long long>(T) * H), 2);
  auto* p = reinterpret_cast<__nv_bfloat16*>(dqkv);
  auto* sv = reinterpret_cast<const __nv_bfloat16*>(save);
  const int pf = param_dtype == B200_DTYPE_F32;
  if (rms) qk_norm_bwd_kernel<true><<<grid, kQkThreads, 0, st>>>(p, sv, mean, rstd, q_weight, k_weight, pf, dparams, T, H);
  else qk_norm_bwd_kernel<false><<<grid, kQkThreads, 0, st>>>(p, sv, mean, rstd, q_weight, k_weight, pf, dparams, T, H);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}
