// LayerNorm forward / backward over the last dimension, one warp per row, row held in registers.
// HBM-bound: forward moves (read x + write y) bytes once; backward reads dy, x and writes dx once,
// dweight/dbias via per-CTA partials + a second reduction kernel.
// Replaces F.layer_norm at timm/layers/norm.py:85-90 (LayerNorm.forward) and its autograd backward.
#include "common.h"
#include "ptx.cuh"

namespace b200 {

constexpr int kLnWarps = 8;

template <bool F32>
__device__ __forceinline__ void load8(const void* base, long long idx, float (&v)[8]) {
  if constexpr (F32) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
    float4 a = p[0], b = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    uint4 q = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
    v[0] = bf16_lo(q.x); v[1] = bf16_hi(q.x); v[2] = bf16_lo(q.y); v[3] = bf16_hi(q.y);
    v[4] = bf16_lo(q.z); v[5] = bf16_hi(q.z); v[6] = bf16_lo(q.w); v[7] = bf16_hi(q.w);
  }
}
template <bool F32>
__device__ __forceinline__ void store8(void* base, long long idx, const float (&v)[8]) {
  if constexpr (F32) {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 q;
    q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]);
    q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = q;
  }
}
__device__ __forceinline__ void load8_rt(const void* base, bool f32, long long idx, float (&v)[8]) {
  if (f32) load8<true>(base, idx, v); else load8<false>(base, idx, v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// VPL: 8-element vectors per lane (cols <= VPL * 256).
template <int VPL, bool XF32, bool YF32>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_fwd_kernel(const void* __restrict__ x, const void* __restrict__ w, const void* __restrict__ b,
                     void* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                     long long rows, int cols, float eps, int param_f32) {
  // param_f32: bit 0 = fp32 parameters, bit 1 = RMS mode (timm RmsNorm, norm.py:202-248 / fast_norm.py:119-139:
  // y = x * rsqrt(mean(x^2) + eps) * w -- the same kernel with the mean pinned to zero and no bias)
  const bool rms = (param_f32 & 2) != 0;
  param_f32 &= 1;
  griddep_launch();     // programmatic dependent launch (ptx.cuh): only the launch latency overlaps here
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const long long warp_global = static_cast<long long>(blockIdx.x) * kLnWarps + (threadIdx.x >> 5);
  const long long warp_stride = static_cast<long long>(gridDim.x) * kLnWarps;
  const int nvec = cols >> 3;
  const float inv_n = 1.0f / static_cast<float>(cols);
  for (long long row = warp_global; row < rows; row += warp_stride) {
    float v[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vec = lane + i * 32;
      if (vec < nvec) {
        load8<XF32>(x, row * cols + vec * 8, v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
      }
    }
    const float mean = rms ? 0.f : warp_sum(s) * inv_n;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (lane + i * 32 < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; sq = fmaf(d, d, sq); }
      }
    }
    const float var = warp_sum(sq) * inv_n;
    const float rstd = rsqrtf(var + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vec = lane + i * 32;
      if (vec < nvec) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd;
        if (w != nullptr) {
          float wv[8];
          load8_rt(w, param_f32, vec * 8, wv);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] *= wv[j];
        }
        if (b != nullptr) {
          float bv[8];
          load8_rt(b, param_f32, vec * 8, bv);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += bv[j];
        }
        store8<YF32>(y, row * cols + vec * 8, o);
      }
    }
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
  }
}

// Backward. dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w,  xhat = (x - mean) * rstd.
// Per-CTA partial sums of dy*xhat (dweight) and dy (dbias) go to ws[blockIdx.x][2*cols].
template <int VPL, bool XF32, bool DYF32>
__global__ void __launch_bounds__(kLnWarps * 32, VPL <= 4 ? 2 : 1)
layernorm_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ x, const void* __restrict__ w,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                     const void* __restrict__ dres, void* __restrict__ dx,
                     float* __restrict__ ws, long long rows, int cols, int param_f32) {
  extern __shared__ float red[];  // [kLnWarps][2*cols]
  const bool rms = (param_f32 & 2) != 0;   // see layernorm_fwd_kernel
  param_f32 &= 1;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long warp_global = static_cast<long long>(blockIdx.x) * kLnWarps + warp;
  const long long warp_stride = static_cast<long long>(gridDim.x) * kLnWarps;
  const int nvec = cols >> 3;
  const float inv_n = 1.0f / static_cast<float>(cols);
  float dw_acc[VPL][8], db_acc[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { dw_acc[i][j] = 0.f; db_acc[i][j] = 0.f; }

  for (long long row = warp_global; row < rows; row += warp_stride) {
    const float mean = rms ? 0.f : mean_in[row], rstd = rstd_in[row];
    // pass 1: row statistics of g = dy * w and the dgamma / dbeta partial sums.  x / dy are NOT kept in registers
    // (that is what capped the kernel at 8 warps per SM): pass 2 re-reads them, and those reads hit L1.
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vec = lane + i * 32;
      if (vec < nvec) {
        float xv[8], dyv[8], wv[8];
        load8<XF32>(x, row * cols + vec * 8, xv);
        load8<DYF32>(dy, row * cols + vec * 8, dyv);
        if (w != nullptr) load8_rt(w, param_f32, vec * 8, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          const float g = (w != nullptr) ? dyv[j] * wv[j] : dyv[j];
          s1 += g;
          s2 = fmaf(g, xh, s2);
          dw_acc[i][j] = fmaf(dyv[j], xh, dw_acc[i][j]);
          db_acc[i][j] += dyv[j];
        }
      }
    }
    const float m1 = rms ? 0.f : warp_sum(s1) * inv_n;     // RMS mode: no mean was subtracted in the forward
    const float m2 = warp_sum(s2) * inv_n;
    // pass 2: dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) (+ residual-branch gradient)
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vec = lane + i * 32;
      if (vec < nvec) {
        float xv[8], dyv[8], wv[8], o[8];
        load8<XF32>(x, row * cols + vec * 8, xv);
        load8<DYF32>(dy, row * cols + vec * 8, dyv);
        if (w != nullptr) load8_rt(w, param_f32, vec * 8, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          const float g = (w != nullptr) ? dyv[j] * wv[j] : dyv[j];
          o[j] = rstd * (g - m1 - xh * m2);
        }
        if (dres != nullptr) {   // gradient arriving through the residual branch around this LayerNorm
          float rv[8];
          load8<XF32>(dres, row * cols + vec * 8, rv);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += rv[j];
        }
        store8<XF32>(dx, row * cols + vec * 8, o);
      }
    }
  }
  // CTA reduction of the per-warp partials, then one row of the workspace per CTA.
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = lane + i * 32;
    if (vec < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        red[warp * 2 * cols + vec * 8 + j] = dw_acc[i][j];
        red[warp * 2 * cols + cols + vec * 8 + j] = db_acc[i][j];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * cols; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kLnWarps; ++k) s += red[k * 2 * cols + c];
    ws[static_cast<long long>(blockIdx.x) * 2 * cols + c] = s;
  }
}

// bf16 activations (the training path): ONE pass.  All of a row's x / dy / residual-gradient vectors are requested up
// front as raw 16-byte words and stay packed in registers (9 loads in flight per lane, 16 warps per SM ~ 70 KB in
// flight per SM -- enough to cover HBM latency), the statistics and the dx formula both run from those registers.
__device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
  v[0] = bf16_lo(q.x); v[1] = bf16_hi(q.x); v[2] = bf16_lo(q.y); v[3] = bf16_hi(q.y);
  v[4] = bf16_lo(q.z); v[5] = bf16_hi(q.z); v[6] = bf16_lo(q.w); v[7] = bf16_hi(q.w);
}

template <int VPL>
__global__ void __launch_bounds__(kLnWarps * 32, VPL <= 3 ? 2 : 1)
layernorm_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                          const void* __restrict__ w, const float* __restrict__ mean_in,
                          const float* __restrict__ rstd_in, const __nv_bfloat16* __restrict__ dres,
                          __nv_bfloat16* __restrict__ dx, float* __restrict__ ws, long long rows, int cols,
                          int param_f32) {
  extern __shared__ float red[];  // [kLnWarps][2*cols]
  const bool rms = (param_f32 & 2) != 0;   // see layernorm_fwd_kernel
  param_f32 &= 1;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long warp_global = static_cast<long long>(blockIdx.x) * kLnWarps + warp;
  const long long warp_stride = static_cast<long long>(gridDim.x) * kLnWarps;
  const int nvec = cols >> 3;
  const float inv_n = 1.0f / static_cast<float>(cols);
  float dw_acc[VPL][8], db_acc[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { dw_acc[i][j] = 0.f; db_acc[i][j] = 0.f; }

  for (long long row = warp_global; row < rows; row += warp_stride) {
    uint4 xr[VPL], dyr[VPL], rr[VPL];
    const long long base = row * cols;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vec = lane + i * 32;
      if (vec < nvec) {
        xr[i] = *reinterpret_cast<const uint4*>(x + base + vec * 8);
        dyr[i] = *reinterpret_cast<const uint4*>(dy + base + vec * 8);
        if (dres != nullptr) rr[i] = *reinterpret_cast<const uint4*>(dres + base + vec * 8);
      }
    }
    const float mean = rms ? 0.f : mean_in[row], rstd = rstd_in[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vec = lane + i * 32;
      if (vec < nvec) {
        float xv[8], dyv[8], wv[8];
        unpack8(xr[i], xv);
        unpack8(dyr[i], dyv);
        if (w != nullptr) load8_rt(w, param_f32, vec * 8, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          const float g = (w != nullptr) ? dyv[j] * wv[j] : dyv[j];
          s1 += g;
          s2 = fmaf(g, xh, s2);
          dw_acc[i][j] = fmaf(dyv[j], xh, dw_acc[i][j]);
          db_acc[i][j] += dyv[j];
        }
      }
    }
    const float m1 = rms ? 0.f : warp_sum(s1) * inv_n;     // RMS mode: no mean was subtracted in the forward
    const float m2 = warp_sum(s2) * inv_n;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vec = lane + i * 32;
      if (vec < nvec) {
        float xv[8], dyv[8], wv[8], o[8];
        unpack8(xr[i], xv);
        unpack8(dyr[i], dyv);
        if (w != nullptr) load8_rt(w, param_f32, vec * 8, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          const float g = (w != nullptr) ? dyv[j] * wv[j] : dyv[j];
          o[j] = rstd * (g - m1 - xh * m2);
        }
        if (dres != nullptr) {   // gradient arriving through the residual branch around this LayerNorm
          float rv[8];
          unpack8(rr[i], rv);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += rv[j];
        }
        store8<false>(dx, base + vec * 8, o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = lane + i * 32;
    if (vec < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        red[warp * 2 * cols + vec * 8 + j] = dw_acc[i][j];
        red[warp * 2 * cols + cols + vec * 8 + j] = db_acc[i][j];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * cols; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kLnWarps; ++k) s += red[k * 2 * cols + c];
    ws[static_cast<long long>(blockIdx.x) * 2 * cols + c] = s;
  }
}

// bf16 activations, rows <= 1024 columns: the training path's kernel.  The one-pass kernel above keeps a row's vectors in
// registers, so the bytes in flight per SM are capped by registers (16 warps x one row ~ 70 KB) and ncu shows it waiting on
// loads (long_scoreboard, 57 % of the HBM peak).  Here every warp owns a RING of row slots in shared memory (kLnRing) filled
// by 1-D bulk copies (cp.async.bulk -> mbarrier complete_tx): a warp has kLnRing rows (x, dy and the residual gradient:
// 3 x cols x 2 bytes each) in flight while it computes one -- 4 x 8 warps x 4.6 KB ~ 147 KB in flight per SM at cols = 768 --
// and the registers only hold the accumulators.  Everything else (formulas, per-CTA partials, second-stage reduction) is
// the one-pass kernel's.

template <int VPL>
__global__ void __launch_bounds__(kLnWarps * 32, 1)
layernorm_bwd_ring_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                          const void* __restrict__ w, const float* __restrict__ mean_in,
                          const float* __restrict__ rstd_in, const __nv_bfloat16* __restrict__ dres,
                          __nv_bfloat16* __restrict__ dx, float* __restrict__ ws, long long rows, int cols,
                          int param_f32, int kLnRing) {
  extern __shared__ __align__(128) uint8_t ring_raw[];
  const bool rms = (param_f32 & 2) != 0;
  param_f32 &= 1;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int narr = dres != nullptr ? 3 : 2;
  const uint32_t row_bytes = static_cast<uint32_t>(cols) * 2u;
  const uint32_t slot_bytes = 3u * row_bytes;
  uint8_t* my_ring = ring_raw + static_cast<size_t>(warp) * kLnRing * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_raw + static_cast<size_t>(kLnWarps) * kLnRing * slot_bytes) + warp * kLnRing;
  const long long warp_global = static_cast<long long>(blockIdx.x) * kLnWarps + warp;
  const long long warp_stride = static_cast<long long>(gridDim.x) * kLnWarps;
  const int nvec = cols >> 3;
  const float inv_n = 1.0f / static_cast<float>(cols);
  if (lane == 0) {
    for (int sidx = 0; sidx < kLnRing; ++sidx) mbar_init(&bars[sidx], 1);
    fence_barrier_init();
  }
  __syncwarp();
  auto issue = [&](long long row, int slot) {       // lane 0: request one row's x | dy | dres into ring slot `slot`
    uint8_t* dst = my_ring + static_cast<size_t>(slot) * slot_bytes;
    mbar_arrive_expect_tx(&bars[slot], narr * row_bytes);
    bulk_load_1d(dst, x + row * cols, row_bytes, &bars[slot]);
    bulk_load_1d(dst + row_bytes, dy + row * cols, row_bytes, &bars[slot]);
    if (narr == 3) bulk_load_1d(dst + 2 * row_bytes, dres + row * cols, row_bytes, &bars[slot]);
  };
  if (lane == 0) {
    for (int sidx = 0; sidx < kLnRing; ++sidx) {
      const long long r = warp_global + sidx * warp_stride;
      if (r < rows) issue(r, sidx);
    }
  }
  // Packed fp32 pairs (FFMA2 / FADD2 / FMUL2: two lanes per instruction) for all per-element math; xhat and g = dy * w of the
  // row stay in registers between the statistics pass and the dx pass (one CTA per SM: 255 registers per thread are there).
  f32x2 dw_acc[VPL][4], db_acc[VPL][4], wv[VPL][4];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = lane + i * 32;
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = 1.f;
    if (w != nullptr && vec < nvec) load8_rt(w, param_f32, vec * 8, t);
#pragma unroll
    for (int j = 0; j < 4; ++j) { dw_acc[i][j] = splat(0.f); db_acc[i][j] = splat(0.f); wv[i][j] = pk(t[2 * j], t[2 * j + 1]); }
  }
  int slot = 0;
  uint32_t phase = 0;
  for (long long row = warp_global; row < rows; row += warp_stride) {
    const float mean = rms ? 0.f : __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
    mbar_wait(&bars[slot], phase);
    const uint8_t* src = my_ring + static_cast<size_t>(slot) * slot_bytes;
    uint4 xr[VPL], dyr[VPL], rr[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vec = lane + i * 32;
      if (vec < nvec) {
        xr[i] = *reinterpret_cast<const uint4*>(src + vec * 16);
        dyr[i] = *reinterpret_cast<const uint4*>(src + row_bytes + vec * 16);
        if (narr == 3) rr[i] = *reinterpret_cast<const uint4*>(src + 2 * row_bytes + vec * 16);
      }
    }
    // the slot's data are in registers: hand it back to the copy engine for the row kLnRing iterations ahead
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      const long long r = row + kLnRing * warp_stride;
      if (r < rows) issue(r, slot);
    }
    const f32x2 rstd2 = splat(rstd), nmr2 = splat(-mean * rstd);
    f32x2 s1 = splat(0.f), s2 = splat(0.f);
    f32x2 xh[VPL][4], gg[VPL][4];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (lane + i * 32 < nvec) {
        const uint32_t xw[4] = {xr[i].x, xr[i].y, xr[i].z, xr[i].w}, dw_[4] = {dyr[i].x, dyr[i].y, dyr[i].z, dyr[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const f32x2 dyv = bf16x2_to_f32x2(dw_[j]);
          xh[i][j] = fma2(bf16x2_to_f32x2(xw[j]), rstd2, nmr2);          // (x - mean) * rstd
          gg[i][j] = mul2(dyv, wv[i][j]);
          s1 = add2(s1, gg[i][j]);
          s2 = fma2(gg[i][j], xh[i][j], s2);
          dw_acc[i][j] = fma2(dyv, xh[i][j], dw_acc[i][j]);
          db_acc[i][j] = add2(db_acc[i][j], dyv);
        }
      }
    }
    float s1a, s1b, s2a, s2b;
    upk(s1, s1a, s1b);
    upk(s2, s2a, s2b);
    const float m1 = rms ? 0.f : warp_sum(s1a + s1b) * inv_n;
    const float m2 = warp_sum(s2a + s2b) * inv_n;
    // dx = rstd * (g - m1 - xhat * m2) = g * rstd + (xhat * (-m2 * rstd) - m1 * rstd)
    const f32x2 a2 = splat(-m2 * rstd), c2 = splat(-m1 * rstd);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vec = lane + i * 32;
      if (vec < nvec) {
        const uint32_t rw[4] = {rr[i].x, rr[i].y, rr[i].z, rr[i].w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f32x2 v = fma2(gg[i][j], rstd2, fma2(xh[i][j], a2, c2));
          if (narr == 3) v = add2(v, bf16x2_to_f32x2(rw[j]));
          float lo, hi;
          upk(v, lo, hi);
          o[j] = pack_bf16(lo, hi);
        }
        *reinterpret_cast<uint4*>(dx + row * cols + vec * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    if (++slot == kLnRing) { slot = 0; phase ^= 1; }
  }
  // per-CTA partial sums of dgamma / dbeta: the ring memory is dead once every warp has left its loop
  __syncthreads();
  float* red = reinterpret_cast<float*>(ring_raw);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = lane + i * 32;
    if (vec < nvec) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float lo, hi;
        upk(dw_acc[i][j], lo, hi);
        red[warp * 2 * cols + vec * 8 + 2 * j] = lo;
        red[warp * 2 * cols + vec * 8 + 2 * j + 1] = hi;
        upk(db_acc[i][j], lo, hi);
        red[warp * 2 * cols + cols + vec * 8 + 2 * j] = lo;
        red[warp * 2 * cols + cols + vec * 8 + 2 * j + 1] = hi;
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * cols; c += blockDim.x) {
    float sacc = 0.f;
#pragma unroll
    for (int kk = 0; kk < kLnWarps; ++kk) sacc += red[kk * 2 * cols + c];
    ws[static_cast<long long>(blockIdx.x) * 2 * cols + c] = sacc;
  }
}

// Two CTAs per SM (16 warps hide the per-row dependency chain better than 8): <= 128 registers per thread, so xhat / g are
// recomputed in the dx pass and the weight vector is re-read (L1) per row instead of living in registers.
template <int VPL>
__global__ void __launch_bounds__(kLnWarps * 32, 2)
layernorm_bwd_ring2_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                          const void* __restrict__ w, const float* __restrict__ mean_in,
                          const float* __restrict__ rstd_in, const __nv_bfloat16* __restrict__ dres,
                          __nv_bfloat16* __restrict__ dx, float* __restrict__ ws, long long rows, int cols,
                          int param_f32, int kLnRing) {
  extern __shared__ __align__(128) uint8_t ring_raw[];
  const bool rms = (param_f32 & 2) != 0;
  param_f32 &= 1;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int narr = dres != nullptr ? 3 : 2;
  const uint32_t row_bytes = static_cast<uint32_t>(cols) * 2u;
  const uint32_t slot_bytes = 3u * row_bytes;
  uint8_t* my_ring = ring_raw + static_cast<size_t>(warp) * kLnRing * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_raw + static_cast<size_t>(kLnWarps) * kLnRing * slot_bytes) + warp * kLnRing;
  const long long warp_global = static_cast<long long>(blockIdx.x) * kLnWarps + warp;
  const long long warp_stride = static_cast<long long>(gridDim.x) * kLnWarps;
  const int nvec = cols >> 3;
  const float inv_n = 1.0f / static_cast<float>(cols);
  if (lane == 0) {
    for (int sidx = 0; sidx < kLnRing; ++sidx) mbar_init(&bars[sidx], 1);
    fence_barrier_init();
  }
  __syncwarp();
  auto issue = [&](long long row, int slot) {       // lane 0: request one row's x | dy | dres into ring slot `slot`
    uint8_t* dst = my_ring + static_cast<size_t>(slot) * slot_bytes;
    mbar_arrive_expect_tx(&bars[slot], narr * row_bytes);
    bulk_load_1d(dst, x + row * cols, row_bytes, &bars[slot]);
    bulk_load_1d(dst + row_bytes, dy + row * cols, row_bytes, &bars[slot]);
    if (narr == 3) bulk_load_1d(dst + 2 * row_bytes, dres + row * cols, row_bytes, &bars[slot]);
  };
  if (lane == 0) {
    for (int sidx = 0; sidx < kLnRing; ++sidx) {
      const long long r = warp_global + sidx * warp_stride;
      if (r < rows) issue(r, sidx);
    }
  }
  // Packed fp32 pairs (FFMA2 / FADD2 / FMUL2: two lanes per instruction) for all per-element math; xhat and g = dy * w of the
  // row stay in registers between the statistics pass and the dx pass (one CTA per SM: 255 registers per thread are there).
  f32x2 dw_acc[VPL][4], db_acc[VPL][4];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = lane + i * 32;
    (void)vec;
#pragma unroll
    for (int j = 0; j < 4; ++j) { dw_acc[i][j] = splat(0.f); db_acc[i][j] = splat(0.f); }
  }
  int slot = 0;
  uint32_t phase = 0;
  for (long long row = warp_global; row < rows; row += warp_stride) {
    const float mean = rms ? 0.f : __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
    mbar_wait(&bars[slot], phase);
    const uint8_t* src = my_ring + static_cast<size_t>(slot) * slot_bytes;
    uint4 xr[VPL], dyr[VPL], rr[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vec = lane + i * 32;
      if (vec < nvec) {
        xr[i] = *reinterpret_cast<const uint4*>(src + vec * 16);
        dyr[i] = *reinterpret_cast<const uint4*>(src + row_bytes + vec * 16);
        if (narr == 3) rr[i] = *reinterpret_cast<const uint4*>(src + 2 * row_bytes + vec * 16);
      }
    }
    // the slot's data are in registers: hand it back to the copy engine for the row kLnRing iterations ahead
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      const long long r = row + kLnRing * warp_stride;
      if (r < rows) issue(r, slot);
    }
    const f32x2 rstd2 = splat(rstd), nmr2 = splat(-mean * rstd);
    f32x2 s1 = splat(0.f), s2 = splat(0.f);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vec = lane + i * 32;
      if (vec < nvec) {
        const uint32_t xw[4] = {xr[i].x, xr[i].y, xr[i].z, xr[i].w}, dw_[4] = {dyr[i].x, dyr[i].y, dyr[i].z, dyr[i].w};
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = 1.f;
        if (w != nullptr) load8_rt(w, param_f32, vec * 8, t);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const f32x2 dyv = bf16x2_to_f32x2(dw_[j]);
          const f32x2 xh = fma2(bf16x2_to_f32x2(xw[j]), rstd2, nmr2);    // (x - mean) * rstd
          const f32x2 gg = mul2(dyv, pk(t[2 * j], t[2 * j + 1]));
          s1 = add2(s1, gg);
          s2 = fma2(gg, xh, s2);
          dw_acc[i][j] = fma2(dyv, xh, dw_acc[i][j]);
          db_acc[i][j] = add2(db_acc[i][j], dyv);
        }
      }
    }
    float s1a, s1b, s2a, s2b;
    upk(s1, s1a, s1b);
    upk(s2, s2a, s2b);
    const float m1 = rms ? 0.f : warp_sum(s1a + s1b) * inv_n;
    const float m2 = warp_sum(s2a + s2b) * inv_n;
    // dx = rstd * (g - m1 - xhat * m2) = g * rstd + (xhat * (-m2 * rstd) - m1 * rstd)
    const f32x2 a2 = splat(-m2 * rstd), c2 = splat(-m1 * rstd);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vec = lane + i * 32;
      if (vec < nvec) {
        const uint32_t rw[4] = {rr[i].x, rr[i].y, rr[i].z, rr[i].w};
        const uint32_t xw[4] = {xr[i].x, xr[i].y, xr[i].z, xr[i].w}, dw_[4] = {dyr[i].x, dyr[i].y, dyr[i].z, dyr[i].w};
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = 1.f;
        if (w != nullptr) load8_rt(w, param_f32, vec * 8, t);
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const f32x2 xh = fma2(bf16x2_to_f32x2(xw[j]), rstd2, nmr2);
          const f32x2 gg = mul2(bf16x2_to_f32x2(dw_[j]), pk(t[2 * j], t[2 * j + 1]));
          f32x2 v = fma2(gg, rstd2, fma2(xh, a2, c2));
          if (narr == 3) v = add2(v, bf16x2_to_f32x2(rw[j]));
          float lo, hi;
          upk(v, lo, hi);
          o[j] = pack_bf16(lo, hi);
        }
        *reinterpret_cast<uint4*>(dx + row * cols + vec * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    if (++slot == kLnRing) { slot = 0; phase ^= 1; }
  }
  // per-CTA partial sums of dgamma / dbeta: the ring memory is dead once every warp has left its loop
  __syncthreads();
  float* red = reinterpret_cast<float*>(ring_raw);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = lane + i * 32;
    if (vec < nvec) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float lo, hi;
        upk(dw_acc[i][j], lo, hi);
        red[warp * 2 * cols + vec * 8 + 2 * j] = lo;
        red[warp * 2 * cols + vec * 8 + 2 * j + 1] = hi;
        upk(db_acc[i][j], lo, hi);
        red[warp * 2 * cols + cols + vec * 8 + 2 * j] = lo;
        red[warp * 2 * cols + cols + vec * 8 + 2 * j + 1] = hi;
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * cols; c += blockDim.x) {
    float sacc = 0.f;
#pragma unroll
    for (int kk = 0; kk < kLnWarps; ++kk) sacc += red[kk * 2 * cols + c];
    ws[static_cast<long long>(blockIdx.x) * 2 * cols + c] = sacc;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Wide rows (2048 < cols <= 8192: the inner norm of an Mlp / gated Mlp over the HIDDEN width -- Mlp(norm_layer=...),
// `scale_mlp_norm`, EVA-02's SwiGLU norm; timm/layers/mlp.py:44, 100, 148): one CTA of 256 threads per row, the row's
// 16-byte vectors packed in registers (VB = ceil(cols / 2048) per thread), statistics through a two-level (warp shuffle +
// shared memory) reduction.  bf16 activations only.  Same formulas as the warp-per-row kernels above.
__device__ __forceinline__ float block_sum256(float v, float* scratch) {
  v = warp_sum(v);
  __syncthreads();                      // scratch may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += scratch[i];
  return r;
}

template <int VB>
__global__ void __launch_bounds__(256)
layernorm_wide_fwd_kernel(const __nv_bfloat16* __restrict__ x, const void* __restrict__ w, const void* __restrict__ b,
                          __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                          long long rows, int cols, float eps, int param_f32) {
  __shared__ float scratch[8];
  const bool rms = (param_f32 & 2) != 0;
  param_f32 &= 1;
  const int nvec = cols >> 3;
  const float inv_n = 1.0f / static_cast<float>(cols);
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    uint4 xr[VB];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VB; ++i) {
      const int vec = threadIdx.x + i * 256;
      if (vec < nvec) {
        xr[i] = *reinterpret_cast<const uint4*>(x + row * cols + vec * 8);
        float v[8];
        unpack8(xr[i], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[j];
      }
    }
    const float mean = rms ? 0.f : block_sum256(s, scratch) * inv_n;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VB; ++i) {
      if (threadIdx.x + i * 256 < nvec) {
        float v[8];
        unpack8(xr[i], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; sq = fmaf(d, d, sq); }
      }
    }
    const float rstd = rsqrtf(block_sum256(sq, scratch) * inv_n + eps);
#pragma unroll
    for (int i = 0; i < VB; ++i) {
      const int vec = threadIdx.x + i * 256;
      if (vec < nvec) {
        float v[8], o[8];
        unpack8(xr[i], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rstd;
        if (w != nullptr) {
          float wv[8];
          load8_rt(w, param_f32, vec * 8, wv);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] *= wv[j];
        }
        if (b != nullptr) {
          float bv[8];
          load8_rt(b, param_f32, vec * 8, bv);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += bv[j];
        }
        store8<false>(y, row * cols + vec * 8, o);
      }
    }
    if (threadIdx.x == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
  }
}

// ws[blockIdx.x][2 * cols]: this CTA's partial sums of dy * xhat | dy over its rows (finished by reduce_rows_kernel).
template <int VB>
__global__ void __launch_bounds__(256)
layernorm_wide_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, const void* __restrict__ w,
                          const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                          const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx, float* __restrict__ ws,
                          long long rows, int cols, int param_f32) {
  __shared__ float scratch[8];
  const bool rms = (param_f32 & 2) != 0;
  param_f32 &= 1;
  const int nvec = cols >> 3;
  const float inv_n = 1.0f / static_cast<float>(cols);
  float dw_acc[VB][8], db_acc[VB][8];
#pragma unroll
  for (int i = 0; i < VB; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { dw_acc[i][j] = 0.f; db_acc[i][j] = 0.f; }
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    uint4 xr[VB], dyr[VB];
    const float mean = rms ? 0.f : mean_in[row], rstd = rstd_in[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VB; ++i) {
      const int vec = threadIdx.x + i * 256;
      if (vec < nvec) {
        xr[i] = *reinterpret_cast<const uint4*>(x + row * cols + vec * 8);
        dyr[i] = *reinterpret_cast<const uint4*>(dy + row * cols + vec * 8);
        float xv[8], dyv[8], wv[8];
        unpack8(xr[i], xv);
        unpack8(dyr[i], dyv);
        if (w != nullptr) load8_rt(w, param_f32, vec * 8, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          const float g = (w != nullptr) ? dyv[j] * wv[j] : dyv[j];
          s1 += g;
          s2 = fmaf(g, xh, s2);
          dw_acc[i][j] = fmaf(dyv[j], xh, dw_acc[i][j]);
          db_acc[i][j] += dyv[j];
        }
      }
    }
    const float m1 = rms ? 0.f : block_sum256(s1, scratch) * inv_n;
    const float m2 = block_sum256(s2, scratch) * inv_n;
#pragma unroll
    for (int i = 0; i < VB; ++i) {
      const int vec = threadIdx.x + i * 256;
      if (vec < nvec) {
        float xv[8], dyv[8], wv[8], o[8];
        unpack8(xr[i], xv);
        unpack8(dyr[i], dyv);
        if (w != nullptr) load8_rt(w, param_f32, vec * 8, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          const float g = (w != nullptr) ? dyv[j] * wv[j] : dyv[j];
          o[j] = rstd * (g - m1 - xh * m2);
        }
        if (dres != nullptr) {
          float rv[8];
          unpack8(*reinterpret_cast<const uint4*>(dres + row * cols + vec * 8), rv);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += rv[j];
        }
        store8<false>(dx, row * cols + vec * 8, o);
      }
    }
  }
  float* out = ws + static_cast<long long>(blockIdx.x) * 2 * cols;
#pragma unroll
  for (int i = 0; i < VB; ++i) {
    const int vec = threadIdx.x + i * 256;
    if (vec < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        out[vec * 8 + j] = dw_acc[i][j];
        out[cols + vec * 8 + j] = db_acc[i][j];
      }
    }
  }
}

constexpr int kLnMaxCols = 8192;
static int ln_wide_grid(long long rows) {
  const long long cap = static_cast<long long>(sm_count()) * 4;
  return static_cast<int>(rows < cap ? rows : cap);
}

// Second stage: out[c] = sum_r ws[r][c] over the per-CTA partial rows.  blockDim = (32 columns, 8 row slices):
// each warp streams its row slice with coalesced 128-byte reads, then the 8 slices are combined through smem, so
// the reduction is a handful of independent loads per thread instead of one serial chain over all rows.
__global__ void __launch_bounds__(256)
reduce_rows_kernel(const float* __restrict__ ws, float* __restrict__ out0, float* __restrict__ out1, int nrows,
                   int cols) {
  __shared__ float part[8][33];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (c < 2 * cols) {
#pragma unroll 4
    for (int r = slice; r < nrows; r += 8) s += ws[static_cast<long long>(r) * 2 * cols + c];
  }
  part[slice][lane] = s;
  __syncthreads();
  if (slice == 0 && c < 2 * cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][lane];
    if (c < cols) { if (out0) out0[c] = t; }
    else          { if (out1) out1[c - cols] = t; }
  }
}

static int ln_grid(long long rows) {
  long long want = (rows + kLnWarps - 1) / kLnWarps;
  long long cap = static_cast<long long>(sm_count()) * 8;
  return static_cast<int>(want < cap ? want : cap);
}
static int ln_bwd_grid(long long rows) {
  long long want = (rows + kLnWarps - 1) / kLnWarps;
  long long cap = static_cast<long long>(sm_count()) * 2;
  return static_cast<int>(want < cap ? want : cap);
}
static int ln_ring_mode() {       // 0 = register kernel, 1 = ring (1 CTA / SM, deep ring), 2 = ring (2 CTAs / SM, 2 slots; measured equal: 71.9 vs 73.4 us); default 1
  static const int m = [] { const char* e = getenv("B200_LN_BWD_RING"); return e == nullptr ? 1 : atoi(e); }();
  return m;
}
static bool ln_use_ring() { return ln_ring_mode() != 0; }
static bool ln_ring_lean(int cols) { return ln_ring_mode() == 2 && cols <= 768; }    // 2 x (8 warps x 2 slots x 3 rows) fits up to 768 columns
static int ln_ring_grid(long long rows, int cols = 1024) {       // CTAs per SM: 1 (deep ring) or 2 (lean)
  long long want = (rows + kLnWarps - 1) / kLnWarps;
  long long cap = static_cast<long long>(sm_count()) * (ln_ring_lean(cols) ? 2 : 1);
  return static_cast<int>(want < cap ? want : cap);
}
static int ln_ring_depth(int cols) {        // as many row slots per warp as ~220 KB of shared memory hold (768 columns: 6)
  static const int forced = [] { const char* e = getenv("B200_LN_RING_DEPTH"); return e ? atoi(e) : 0; }();
  int d = (220 * 1024) / (kLnWarps * 3 * cols * 2);
  if (ln_ring_lean(cols)) d = 2;
  if (forced > 0 && forced < d) d = forced;
  return d < 2 ? 2 : (d > 8 ? 8 : d);
}
static size_t ln_ring_smem(int cols) {
  return static_cast<size_t>(kLnWarps) * ln_ring_depth(cols) * 3 * cols * 2 + kLnWarps * ln_ring_depth(cols) * 8;
}

template <int VPL>
static void launch_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, long long rows,
                       int cols, float eps, int xf32, int yf32, int pf32, cudaStream_t st) {
  const int grid = ln_grid(rows);
  const int threads = kLnWarps * 32;
  const dim3 g(grid, 1, 1), t(threads, 1, 1);
  const long long rows_ll = rows;
  if (xf32 && yf32) launch_pdl(layernorm_fwd_kernel<VPL, true, true>, g, t, 0, st, 1, x, w, b, y, mean, rstd, rows_ll, cols, eps, pf32);
  else if (xf32) launch_pdl(layernorm_fwd_kernel<VPL, true, false>, g, t, 0, st, 1, x, w, b, y, mean, rstd, rows_ll, cols, eps, pf32);
  else if (yf32) launch_pdl(layernorm_fwd_kernel<VPL, false, true>, g, t, 0, st, 1, x, w, b, y, mean, rstd, rows_ll, cols, eps, pf32);
  else launch_pdl(layernorm_fwd_kernel<VPL, false, false>, g, t, 0, st, 1, x, w, b, y, mean, rstd, rows_ll, cols, eps, pf32);
}

template <int VPL>
static int launch_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd,
                      const void* dres, void* dx, float* ws, long long rows, int cols, int xf32, int dyf32, int pf32, cudaStream_t st) {
  const int grid = ln_bwd_grid(rows);
  const int threads = kLnWarps * 32;
  const size_t smem = static_cast<size_t>(kLnWarps) * 2 * cols * sizeof(float);
  {
    // opt in to > 48 KB where needed and ask for the largest shared-memory carve-out so two CTAs fit per SM
    auto set = [&](auto kern) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      if (e != cudaSuccess) return e;
      return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024));
    };
    if (xf32 && dyf32) B200_CUDA_OK(set(layernorm_bwd_kernel<VPL, true, true>));
    else if (xf32) B200_CUDA_OK(set(layernorm_bwd_kernel<VPL, true, false>));
    else if (dyf32) B200_CUDA_OK(set(layernorm_bwd_kernel<VPL, false, true>));
    else if (VPL <= 4) B200_CUDA_OK(set(layernorm_bwd_bf16_kernel<VPL>));
    else B200_CUDA_OK(set(layernorm_bwd_kernel<VPL, false, false>));
  }
  if (xf32 && dyf32) layernorm_bwd_kernel<VPL, true, true><<<grid, threads, smem, st>>>(dy, x, w, mean, rstd, dres, dx, ws, rows, cols, pf32);
  else if (xf32) layernorm_bwd_kernel<VPL, true, false><<<grid, threads, smem, st>>>(dy, x, w, mean, rstd, dres, dx, ws, rows, cols, pf32);
  else if (dyf32) layernorm_bwd_kernel<VPL, false, true><<<grid, threads, smem, st>>>(dy, x, w, mean, rstd, dres, dx, ws, rows, cols, pf32);
  else if (VPL <= 4 && ln_use_ring() && (cols * 2) % 16 == 0) {
    // the shared-memory ring kernel (bulk copies keep 4 rows per warp in flight)
    const size_t rsmem = ln_ring_smem(cols);
    auto kern = ln_ring_lean(cols) ? layernorm_bwd_ring2_kernel<VPL> : layernorm_bwd_ring_kernel<VPL>;
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    kern<<<ln_ring_grid(rows, cols), threads, rsmem, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(x), w, mean, rstd,
        reinterpret_cast<const __nv_bfloat16*>(dres), reinterpret_cast<__nv_bfloat16*>(dx), ws, rows, cols, pf32, ln_ring_depth(cols));
  } else if (VPL <= 4)   // wider rows would not fit their packed vectors + accumulators in 128 registers
    layernorm_bwd_bf16_kernel<VPL><<<grid, threads, smem, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(x), w, mean, rstd,
        reinterpret_cast<const __nv_bfloat16*>(dres), reinterpret_cast<__nv_bfloat16*>(dx), ws, rows, cols, pf32);
  else layernorm_bwd_kernel<VPL, false, false><<<grid, threads, smem, st>>>(dy, x, w, mean, rstd, dres, dx, ws, rows, cols, pf32);
  return B200_OK;
}

}  // namespace b200

using namespace b200;

static int norm_fwd(const void* x, const void* weight, const void* bias, void* y, float* mean, float* rstd, int64_t rows,
                    int32_t cols, float eps, int32_t x_dtype, int32_t y_dtype, int32_t param_dtype, int rms, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(x && y, "b200_layernorm_fwd: null x/y");
  B200_CHECK_ARG(rows >= 0 && cols > 0, "b200_layernorm_fwd: bad shape rows=%lld cols=%d", (long long)rows, cols);
  if (cols % 8 != 0 || cols > kLnMaxCols)
    return fail(B200_ERR_UNSUPPORTED, "b200_layernorm_fwd: cols=%d must be a multiple of 8 and <= %d", cols, kLnMaxCols);
  if (rows == 0) return B200_OK;
  const int vpl = (cols + 255) / 256;
  const int xf = x_dtype == B200_DTYPE_F32, yf = y_dtype == B200_DTYPE_F32;
  const int pf = (param_dtype == B200_DTYPE_F32 ? 1 : 0) | (rms ? 2 : 0);
  if (cols > 2048) {
    if (xf || yf) return fail(B200_ERR_UNSUPPORTED, "b200_layernorm_fwd: rows wider than 2048 need bf16 activations");
    const int grid = ln_wide_grid(rows);
    auto* xp = reinterpret_cast<const __nv_bfloat16*>(x);
    auto* yp = reinterpret_cast<__nv_bfloat16*>(y);
    switch ((cols + 2047) / 2048) {
      case 2: layernorm_wide_fwd_kernel<2><<<grid, 256, 0, st>>>(xp, weight, bias, yp, mean, rstd, rows, cols, eps, pf); break;
      case 3: layernorm_wide_fwd_kernel<3><<<grid, 256, 0, st>>>(xp, weight, bias, yp, mean, rstd, rows, cols, eps, pf); break;
      default: layernorm_wide_fwd_kernel<4><<<grid, 256, 0, st>>>(xp, weight, bias, yp, mean, rstd, rows, cols, eps, pf); break;
    }
    B200_CUDA_OK(cudaGetLastError());
    return B200_OK;
  }
  switch (vpl) {
    case 1: launch_fwd<1>(x, weight, bias, y, mean, rstd, rows, cols, eps, xf, yf, pf, st); break;
    case 2: launch_fwd<2>(x, weight, bias, y, mean, rstd, rows, cols, eps, xf, yf, pf, st); break;
    case 3: launch_fwd<3>(x, weight, bias, y, mean, rstd, rows, cols, eps, xf, yf, pf, st); break;
    case 4: launch_fwd<4>(x, weight, bias, y, mean, rstd, rows, cols, eps, xf, yf, pf, st); break;
    case 5: case 6: launch_fwd<6>(x, weight, bias, y, mean, rstd, rows, cols, eps, xf, yf, pf, st); break;
    default: launch_fwd<8>(x, weight, bias, y, mean, rstd, rows, cols, eps, xf, yf, pf, st); break;
  }
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_layernorm_fwd(const void* x, const void* weight, const void* bias, void* y, float* mean,
                                  float* rstd, int64_t rows, int32_t cols, float eps, int32_t x_dtype,
                                  int32_t y_dtype, int32_t param_dtype, void* stream_) {
  return norm_fwd(x, weight, bias, y, mean, rstd, rows, cols, eps, x_dtype, y_dtype, param_dtype, 0, stream_);
}

extern "C" int b200_rmsnorm_fwd(const void* x, const void* weight, void* y, float* rstd, int64_t rows, int32_t cols, float eps,
                                int32_t x_dtype, int32_t y_dtype, int32_t param_dtype, void* stream_) {
  return norm_fwd(x, weight, nullptr, y, nullptr, rstd, rows, cols, eps, x_dtype, y_dtype, param_dtype, 1, stream_);
}

extern "C" int b200_layernorm_bwd_ws_rows(int64_t rows, int32_t cols) {
  return cols > 2048 ? ln_wide_grid(rows > 0 ? rows : 1) : ln_bwd_grid(rows > 0 ? rows : 1);
}

static int norm_bwd(const void* dy, const void* x, const void* weight, const float* mean, const float* rstd, const void* dres,
                    void* dx, float* dweight, float* dbias, float* workspace, int64_t rows, int32_t cols, int32_t x_dtype,
                    int32_t dy_dtype, int32_t param_dtype, int rms, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(dy && x && (mean || rms) && rstd && dx && workspace, "b200_layernorm_bwd: null pointer");
  B200_CHECK_ARG(rows > 0 && cols > 0, "b200_layernorm_bwd: bad shape rows=%lld cols=%d", (long long)rows, cols);
  if (cols % 8 != 0 || cols > kLnMaxCols)
    return fail(B200_ERR_UNSUPPORTED, "b200_layernorm_bwd: cols=%d must be a multiple of 8 and <= %d", cols, kLnMaxCols);
  const int vpl = (cols + 255) / 256;
  const int xf = x_dtype == B200_DTYPE_F32, dyf = dy_dtype == B200_DTYPE_F32;
  const int pf = (param_dtype == B200_DTYPE_F32 ? 1 : 0) | (rms ? 2 : 0);
  if (cols > 2048) {
    if (xf || dyf) return fail(B200_ERR_UNSUPPORTED, "b200_layernorm_bwd: rows wider than 2048 need bf16 activations");
    const int grid = ln_wide_grid(rows);
    auto* dyp = reinterpret_cast<const __nv_bfloat16*>(dy);
    auto* xp = reinterpret_cast<const __nv_bfloat16*>(x);
    auto* rp = reinterpret_cast<const __nv_bfloat16*>(dres);
    auto* dxp = reinterpret_cast<__nv_bfloat16*>(dx);
    switch ((cols + 2047) / 2048) {
      case 2: layernorm_wide_bwd_kernel<2><<<grid, 256, 0, st>>>(dyp, xp, weight, mean, rstd, rp, dxp, workspace, rows, cols, pf); break;
      case 3: layernorm_wide_bwd_kernel<3><<<grid, 256, 0, st>>>(dyp, xp, weight, mean, rstd, rp, dxp, workspace, rows, cols, pf); break;
      default: layernorm_wide_bwd_kernel<4><<<grid, 256, 0, st>>>(dyp, xp, weight, mean, rstd, rp, dxp, workspace, rows, cols, pf); break;
    }
    B200_CUDA_OK(cudaGetLastError());
    if (dweight || dbias) {
      reduce_rows_kernel<<<(2 * cols + 31) / 32, 256, 0, st>>>(workspace, dweight, dbias, grid, cols);
      B200_CUDA_OK(cudaGetLastError());
    }
    return B200_OK;
  }
  int rc;
  switch (vpl) {
    case 1: rc = launch_bwd<1>(dy, x, weight, mean, rstd, dres, dx, workspace, rows, cols, xf, dyf, pf, st); break;
    case 2: rc = launch_bwd<2>(dy, x, weight, mean, rstd, dres, dx, workspace, rows, cols, xf, dyf, pf, st); break;
    case 3: rc = launch_bwd<3>(dy, x, weight, mean, rstd, dres, dx, workspace, rows, cols, xf, dyf, pf, st); break;
    case 4: rc = launch_bwd<4>(dy, x, weight, mean, rstd, dres, dx, workspace, rows, cols, xf, dyf, pf, st); break;
    case 5: case 6: rc = launch_bwd<6>(dy, x, weight, mean, rstd, dres, dx, workspace, rows, cols, xf, dyf, pf, st); break;
    default: rc = launch_bwd<8>(dy, x, weight, mean, rstd, dres, dx, workspace, rows, cols, xf, dyf, pf, st); break;
  }
  if (rc) return rc;
  B200_CUDA_OK(cudaGetLastError());
  if (dweight || dbias) {
    const bool ring = !xf && !dyf && vpl <= 4 && ln_use_ring() && (cols * 2) % 16 == 0;     // as launch_bwd chooses
    const int nrows = ring ? ln_ring_grid(rows, cols) : ln_bwd_grid(rows);
    reduce_rows_kernel<<<(2 * cols + 31) / 32, 256, 0, st>>>(workspace, dweight, dbias, nrows, cols);
    B200_CUDA_OK(cudaGetLastError());
  }
  return B200_OK;
}

extern "C" int b200_layernorm_bwd(const void* dy, const void* x, const void* weight, const float* mean,
                                  const float* rstd, const void* dres, void* dx, float* dweight, float* dbias,
                                  float* workspace,
                                  int64_t rows, int32_t cols, int32_t x_dtype, int32_t dy_dtype,
                                  int32_t param_dtype, void* stream_) {
  return norm_bwd(dy, x, weight, mean, rstd, dres, dx, dweight, dbias, workspace, rows, cols, x_dtype, dy_dtype, param_dtype, 0,
                  stream_);
}

extern "C" int b200_rmsnorm_bwd(const void* dy, const void* x, const void* weight, const float* rstd, const void* dres,
                                void* dx, float* dweight, float* workspace, int64_t rows, int32_t cols, int32_t x_dtype,
                                int32_t dy_dtype, int32_t param_dtype, void* stream_) {
  return norm_bwd(dy, x, weight, nullptr, rstd, dres, dx, dweight, nullptr, workspace, rows, cols, x_dtype, dy_dtype, param_dtype,
                  1, stream_);
}
