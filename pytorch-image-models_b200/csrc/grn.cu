// Global Response Norm of ConvNeXt-V2's MLP (timm/layers/grn.py:17-49 inside GlobalResponseNormMlp, mlp.py:251-295),
// channels-last rows h [B, HW, C] (C = the MLP's hidden width):
//   gx[b,c] = ||h[b,:,c]||_2,  nx = gx / (mean_c gx + eps),  h' = h + bias + weight * h * nx = h * s[b,c] + bias,
//   s = 1 + weight * nx.
// HBM-bound: forward = one read of h for the per-image column sums of squares, a [B, C] coefficient kernel, one
// read-modify-write pass that applies s in place.  Backward = one pass for A[b,c] = sum_hw dh' * h with h = GELU(pre)
// recomputed from the saved pre-activation, a [B, C] coefficient kernel, and one pass that writes
// dpre = (dh' * s + h * t) * GELU'(pre) -- the GELU backward that the fc2-dgrad epilogue performs in ConvNeXt-V1 rides here.
#include "common.h"
#include "gelu.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int kGrnThreads = 256;

__device__ __forceinline__ void unpack8f(const uint4& q, float (&v)[8]) {
  v[0] = bf16_lo(q.x); v[1] = bf16_hi(q.x); v[2] = bf16_lo(q.y); v[3] = bf16_hi(q.y);
  v[4] = bf16_lo(q.z); v[5] = bf16_hi(q.z); v[6] = bf16_lo(q.w); v[7] = bf16_hi(q.w);
}
__device__ __forceinline__ void gelu8(float (&v)[8]) {
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    float a, b;
    upk(gelu_erf2(pk(v[i], v[i + 1])), a, b);
    v[i] = a; v[i + 1] = b;
  }
}

// out[b, c] += sum over a slice of the HW rows of a[b, r, c] * b[b, r, c];  GELU_B: the second operand is GELU(b).
// grid = (channel-vector blocks, B, row slices); out is zeroed by the caller.
template <bool SQUARE, bool GELU_B>
__global__ void __launch_bounds__(kGrnThreads) grn_dot_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                              float* __restrict__ out, int HW, int C, int rows_per_slice) {
  const int cv = blockIdx.x * kGrnThreads + threadIdx.x;
  if (cv >= (C >> 3)) return;
  const int img = blockIdx.y;
  const int r0 = blockIdx.z * rows_per_slice;
  const int r1 = min(HW, r0 + rows_per_slice);
  const long long base = static_cast<long long>(img) * HW * C + cv * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int r = r0; r < r1; ++r) {
    float x[8], y[8];
    unpack8f(__ldg(reinterpret_cast<const uint4*>(a + base + static_cast<long long>(r) * C)), x);
    if (SQUARE) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(x[i], x[i], acc[i]);
    } else {
      unpack8f(__ldg(reinterpret_cast<const uint4*>(b + base + static_cast<long long>(r) * C)), y);
      if (GELU_B) gelu8(y);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(x[i], y[i], acc[i]);
    }
  }
  float* o = out + static_cast<long long>(img) * C + cv * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) atomicAdd(o + i, acc[i]);
}

__device__ __forceinline__ float block_sum(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < kGrnThreads / 32; ++i) r += scratch[i];
  return r;
}

__device__ __forceinline__ float ldw(const void* p, int f32, int c) {
  return f32 ? __ldg(reinterpret_cast<const float*>(p) + c) : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[c]);
}

// one block per image: scale[b, c] = 1 + weight[c] * gx / (mean_c gx + eps),  gx = sqrt(gsq)
__global__ void __launch_bounds__(kGrnThreads) grn_coef_fwd_kernel(const float* __restrict__ gsq, const void* __restrict__ weight,
                                                                   int w_f32, float* __restrict__ scale, int C, float eps) {
  __shared__ float scratch[kGrnThreads / 32];
  const long long base = static_cast<long long>(blockIdx.x) * C;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += kGrnThreads) s += sqrtf(gsq[base + c]);
  const float inv = 1.f / (block_sum(s, scratch) / static_cast<float>(C) + eps);
  for (int c = threadIdx.x; c < C; c += kGrnThreads) scale[base + c] = fmaf(ldw(weight, w_f32, c), sqrtf(gsq[base + c]) * inv, 1.f);
}

// h[b, r, c] = h * scale[b, c] + bias[c], in place
__global__ void __launch_bounds__(kGrnThreads) grn_apply_kernel(__nv_bfloat16* __restrict__ h, const float* __restrict__ scale,
                                                                const void* __restrict__ bias, int b_f32, long long total_vecs,
                                                                int HW, int C) {
  const int cvecs = C >> 3;
  for (long long idx = static_cast<long long>(blockIdx.x) * kGrnThreads + threadIdx.x; idx < total_vecs;
       idx += static_cast<long long>(gridDim.x) * kGrnThreads) {
    const int cv = static_cast<int>(idx % cvecs);
    const long long row = idx / cvecs;
    const int img = static_cast<int>(row / HW);
    float x[8];
    uint4* p = reinterpret_cast<uint4*>(h + row * C + cv * 8);
    unpack8f(*p, x);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + static_cast<long long>(img) * C + cv * 8));
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale + static_cast<long long>(img) * C + cv * 8) + 1);
    const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    float bv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (bias != nullptr) {
      if (b_f32) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(bias) + cv * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(bias) + cv * 8) + 1);
        bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
      } else {
        unpack8f(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(bias) + cv * 8)), bv);
      }
    }
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = fmaf(x[i], s[i], bv[i]);
    uint4 q;
    q.x = pack_bf16(y[0], y[1]); q.y = pack_bf16(y[2], y[3]);
    q.z = pack_bf16(y[4], y[5]); q.w = pack_bf16(y[6], y[7]);
    *p = q;
  }
}

// one block per image.  A[b,c] = sum_hw dh' * h = d loss / d s[b,c].  With dnx = A * w, S = sum_c dnx * gx, m = mean_c gx:
//   d gx = dnx / (m + eps) - S / (C (m + eps)^2);   t[b,c] = d gx / gx  (the coefficient of h in dh; 0 where gx = 0)
//   dweight[c] += A[b,c] * nx[b,c]
__global__ void __launch_bounds__(kGrnThreads) grn_coef_bwd_kernel(const float* __restrict__ A, const float* __restrict__ gsq,
                                                                   const void* __restrict__ weight, int w_f32,
                                                                   float* __restrict__ t, float* __restrict__ dweight, int C,
                                                                   float eps) {
  __shared__ float scratch[kGrnThreads / 32];
  const long long base = static_cast<long long>(blockIdx.x) * C;
  float sg = 0.f, sd = 0.f;
  for (int c = threadIdx.x; c < C; c += kGrnThreads) {
    const float gx = sqrtf(gsq[base + c]);
    sg += gx;
    sd = fmaf(A[base + c] * ldw(weight, w_f32, c), gx, sd);
  }
  const float m = block_sum(sg, scratch) / static_cast<float>(C) + eps;
  const float S = block_sum(sd, scratch);
  const float inv = 1.f / m;
  const float corr = S * inv * inv / static_cast<float>(C);
  for (int c = threadIdx.x; c < C; c += kGrnThreads) {
    const float gx = sqrtf(gsq[base + c]);
    const float a = A[base + c];
    const float dgx = a * ldw(weight, w_f32, c) * inv - corr;
    t[base + c] = gx > 0.f ? dgx / gx : 0.f;
    atomicAdd(dweight + c, a * gx * inv);
  }
}

// dpre = (dh' * scale[b,c] + GELU(pre) * t[b,c]) * GELU'(pre)
__global__ void __launch_bounds__(kGrnThreads) grn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dh2, const __nv_bfloat16* __restrict__ pre,
                                                                    const float* __restrict__ scale, const float* __restrict__ t,
                                                                    __nv_bfloat16* __restrict__ dpre, long long total_vecs, int HW,
                                                                    int C) {
  const int cvecs = C >> 3;
  for (long long idx = static_cast<long long>(blockIdx.x) * kGrnThreads + threadIdx.x; idx < total_vecs;
       idx += static_cast<long long>(gridDim.x) * kGrnThreads) {
    const int cv = static_cast<int>(idx % cvecs);
    const long long row = idx / cvecs;
    const long long co = (row / HW) * C + cv * 8;
    float d[8], x[8];
    unpack8f(__ldg(reinterpret_cast<const uint4*>(dh2 + row * C + cv * 8)), d);
    unpack8f(__ldg(reinterpret_cast<const uint4*>(pre + row * C + cv * 8)), x);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + co)), s1 = __ldg(reinterpret_cast<const float4*>(scale + co) + 1);
    const float4 t0 = __ldg(reinterpret_cast<const float4*>(t + co)), t1 = __ldg(reinterpret_cast<const float4*>(t + co) + 1);
    const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float tt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      f32x2 gp, dgp;
      gelu_and_dgelu2(pk(x[i], x[i + 1]), gp, dgp);
      float g0, g1, dg0, dg1;
      upk(gp, g0, g1);
      upk(dgp, dg0, dg1);
      y[i] = fmaf(d[i], s[i], g0 * tt[i]) * dg0;
      y[i + 1] = fmaf(d[i + 1], s[i + 1], g1 * tt[i + 1]) * dg1;
    }
    uint4 q;
    q.x = pack_bf16(y[0], y[1]); q.y = pack_bf16(y[2], y[3]);
    q.z = pack_bf16(y[4], y[5]); q.w = pack_bf16(y[6], y[7]);
    *reinterpret_cast<uint4*>(dpre + row * C + cv * 8) = q;
  }
}

static void dot_grid(int B, int HW, int C, dim3* grid, int* rows_per_slice) {
  const int cblocks = ((C >> 3) + kGrnThreads - 1) / kGrnThreads;
  int slices = (HW + 63) / 64;                       // ~64 rows per thread
  if (slices > 65535) slices = 65535;
  *rows_per_slice = (HW + slices - 1) / slices;
  slices = (HW + *rows_per_slice - 1) / *rows_per_slice;
  *grid = dim3(cblocks, B, slices);
}

static int elementwise_grid(long long total_vecs) {
  long long blocks = (total_vecs + kGrnThreads - 1) / kGrnThreads;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  return static_cast<int>(blocks < cap ? blocks : cap);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_grn_fwd(void* h, float* gsq, float* scale, const void* weight, const void* bias, int32_t param_dtype,
                            int32_t B, int32_t HW, int32_t C, float eps, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(h && gsq && scale && weight, "b200_grn_fwd: null pointer");
  B200_CHECK_ARG(B > 0 && B <= 65535 && HW > 0 && C > 0 && C % 8 == 0, "b200_grn_fwd: bad shape (C %% 8 == 0, B <= 65535)");
  B200_CHECK_ARG(param_dtype == B200_DTYPE_BF16 || param_dtype == B200_DTYPE_F32, "b200_grn_fwd: bad parameter dtype");
  const int pf = param_dtype == B200_DTYPE_F32;
  auto* hp = reinterpret_cast<__nv_bfloat16*>(h);
  dim3 grid;
  int rps;
  dot_grid(B, HW, C, &grid, &rps);
  grn_dot_kernel<true, false><<<grid, kGrnThreads, 0, st>>>(hp, nullptr, gsq, HW, C, rps);
  B200_CUDA_OK(cudaGetLastError());
  grn_coef_fwd_kernel<<<B, kGrnThreads, 0, st>>>(gsq, weight, pf, scale, C, eps);
  B200_CUDA_OK(cudaGetLastError());
  const long long total = static_cast<long long>(B) * HW * (C / 8);
  grn_apply_kernel<<<elementwise_grid(total), kGrnThreads, 0, st>>>(hp, scale, bias, pf, total, HW, C);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_grn_bwd(const void* dh2, const void* pre, const float* gsq, const float* scale, const void* weight,
                            int32_t param_dtype, float* a_ws, float* t_ws, float* dweight, void* dpre, int32_t B, int32_t HW,
                            int32_t C, float eps, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(dh2 && pre && gsq && scale && weight && a_ws && t_ws && dweight && dpre, "b200_grn_bwd: null pointer");
  B200_CHECK_ARG(B > 0 && B <= 65535 && HW > 0 && C > 0 && C % 8 == 0, "b200_grn_bwd: bad shape (C %% 8 == 0, B <= 65535)");
  B200_CHECK_ARG(param_dtype == B200_DTYPE_BF16 || param_dtype == B200_DTYPE_F32, "b200_grn_bwd: bad parameter dtype");
  const int pf = param_dtype == B200_DTYPE_F32;
  auto* dp = reinterpret_cast<const __nv_bfloat16*>(dh2);
  auto* pp = reinterpret_cast<const __nv_bfloat16*>(pre);
  dim3 grid;
  int rps;
  dot_grid(B, HW, C, &grid, &rps);
  grn_dot_kernel<false, true><<<grid, kGrnThreads, 0, st>>>(dp, pp, a_ws, HW, C, rps);
  B200_CUDA_OK(cudaGetLastError());
  grn_coef_bwd_kernel<<<B, kGrnThreads, 0, st>>>(a_ws, gsq, weight, pf, t_ws, dweight, C, eps);
  B200_CUDA_OK(cudaGetLastError());
  const long long total = static_cast<long long>(B) * HW * (C / 8);
  grn_bwd_apply_kernel<<<elementwise_grid(total), kGrnThreads, 0, st>>>(dp, pp, scale, t_ws, reinterpret_cast<__nv_bfloat16*>(dpre),
                                                                        total, HW, C);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}
