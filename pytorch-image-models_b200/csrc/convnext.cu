// ConvNeXt-specific kernels on channels-last (NHWC == [B*H*W, C] token-major) activations:
//   dwconv_nhwc      depthwise k x k convolution (k odd <= 7, stride 1, zero padding k/2) + bias
//                    replaces F.conv2d(groups=C) at timm/models/convnext.py:173-182,200 (and the permutes of :205,208:
//                    the whole network stays NHWC so the following LayerNorm / Linear read contiguous rows)
//   patchify_nhwc    non-overlapping p x p gather for the 2x2/s2 downsample conv (convnext.py:263-275)
//   avgpool_nhwc     global average pool over H*W (classifier.py:211-213)
// The depthwise conv is FMA-bound in fp32 (49 MACs per output element, 0.23 GMAC/img for ConvNeXt-B), not HBM-bound;
// each thread keeps a strip of TW output pixels x 8 channels in registers and slides along the row so every loaded
// input vector feeds up to 7 taps; weights are consumed pre-transposed ([k*k][C]) so a tap is one 16-byte load.
#include <cstdlib>
#include "common.h"
#include "ptx.cuh"

namespace b200 {

constexpr int kDwTW = 7;   // output pixels per thread along W

template <int KS>
__global__ void __launch_bounds__(256)
dwconv_nhwc_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ wt,
                   const void* __restrict__ bias, const __nv_bfloat16* __restrict__ residual,
                   __nv_bfloat16* __restrict__ y, int B, int H, int W, int C, int bias_f32) {
  // wt: weights pre-transposed to [KS*KS][C] so the 8 channels of a thread are one 16-byte load per tap and
  // consecutive threads (consecutive channel groups) read consecutive addresses (L1-resident: 49*C*2 bytes).
  constexpr int PAD = KS / 2;
  const int cvecs = C >> 3;
  const int strips = (W + kDwTW - 1) / kDwTW;
  const long long total = static_cast<long long>(B) * H * strips * cvecs;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % cvecs);
    long long item = idx / cvecs;
    const int strip = static_cast<int>(item % strips);
    item /= strips;
    const int oy = static_cast<int>(item % H);
    const int b = static_cast<int>(item / H);
    const int ox0 = strip * kDwTW;
    f32x2 acc[kDwTW][4];
    {
      f32x2 b0 = 0ull, b1 = 0ull, b2 = 0ull, b3 = 0ull;
      if (bias != nullptr) {
        if (bias_f32) {
          const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(bias) + cv * 8);
          const float4 a = __ldg(p), c2 = __ldg(p + 1);
          b0 = pk(a.x, a.y); b1 = pk(a.z, a.w); b2 = pk(c2.x, c2.y); b3 = pk(c2.z, c2.w);
        } else {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(bias) + cv * 8));
          b0 = bf16x2_to_f32x2(q.x); b1 = bf16x2_to_f32x2(q.y); b2 = bf16x2_to_f32x2(q.z); b3 = bf16x2_to_f32x2(q.w);
        }
      }
#pragma unroll
      for (int t = 0; t < kDwTW; ++t) { acc[t][0] = b0; acc[t][1] = b1; acc[t][2] = b2; acc[t][3] = b3; }
      if (residual != nullptr) {   // backward use: dx = dwconv(dy, flipped w) + gradient of the shortcut branch
        const __nv_bfloat16* rrow = residual + ((static_cast<long long>(b) * H + oy) * W) * C + cv * 8;
#pragma unroll
        for (int t = 0; t < kDwTW; ++t) {
          if (ox0 + t < W) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(rrow + static_cast<long long>(ox0 + t) * C));
            acc[t][0] = add2(acc[t][0], bf16x2_to_f32x2(q.x)); acc[t][1] = add2(acc[t][1], bf16x2_to_f32x2(q.y));
            acc[t][2] = add2(acc[t][2], bf16x2_to_f32x2(q.z)); acc[t][3] = add2(acc[t][3], bf16x2_to_f32x2(q.w));
          }
        }
      }
    }
#pragma unroll 1
    for (int ky = 0; ky < KS; ++ky) {
      const int iy = oy + ky - PAD;
      if (iy < 0 || iy >= H) continue;
      // issue every load of this kernel row first (branch-free, predicated), then do the arithmetic: the 7 weight
      // vectors and the TW + KS - 1 input vectors are all in flight together instead of one L2 round trip each
      uint4 wq[KS];
#pragma unroll
      for (int kx = 0; kx < KS; ++kx)
        wq[kx] = __ldg(reinterpret_cast<const uint4*>(wt + static_cast<long long>(ky * KS + kx) * C + cv * 8));
      const __nv_bfloat16* row = x + ((static_cast<long long>(b) * H + iy) * W) * C + cv * 8;
      uint4 in[kDwTW + KS - 1];
#pragma unroll
      for (int i = 0; i < kDwTW + KS - 1; ++i) {
        const int ix = ox0 + i - PAD;
        const bool ok = ix >= 0 && ix < W;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(row + static_cast<long long>(ok ? ix : 0) * C));
        in[i] = ok ? q : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int i = 0; i < kDwTW + KS - 1; ++i) {
        const f32x2 v0 = bf16x2_to_f32x2(in[i].x), v1 = bf16x2_to_f32x2(in[i].y), v2 = bf16x2_to_f32x2(in[i].z),
                    v3 = bf16x2_to_f32x2(in[i].w);
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          const int t = i - kx;   // output pixel (within the strip) this input contributes to through tap kx
          if (t >= 0 && t < kDwTW) {
            acc[t][0] = fma2(v0, bf16x2_to_f32x2(wq[kx].x), acc[t][0]);
            acc[t][1] = fma2(v1, bf16x2_to_f32x2(wq[kx].y), acc[t][1]);
            acc[t][2] = fma2(v2, bf16x2_to_f32x2(wq[kx].z), acc[t][2]);
            acc[t][3] = fma2(v3, bf16x2_to_f32x2(wq[kx].w), acc[t][3]);
          }
        }
      }
    }
    __nv_bfloat16* orow = y + ((static_cast<long long>(b) * H + oy) * W) * C + cv * 8;
#pragma unroll
    for (int t = 0; t < kDwTW; ++t) {
      if (ox0 + t < W) {
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float lo, hi;
          upk(acc[t][j], lo, hi);
          o[j] = pack_bf16(lo, hi);
        }
        *reinterpret_cast<uint4*>(orow + static_cast<long long>(ox0 + t) * C) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// Depthwise-conv weight gradient: dw_t[ky][kx][c] += sum_{b,oy,ox} dy[b,oy,ox,c] * x[b, oy+ky-PAD, ox+kx-PAD, c].
// A thread owns (8 channels, one kernel row ky) and walks a slab of output strips, keeping the KS x 8 partial sums in
// registers; one round of fp32 atomics per thread at the end (dw_t is zero-initialised by the caller).
template <int KS>
__global__ void __launch_bounds__(256)
dwconv_nhwc_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                         float* __restrict__ dwt, int B, int H, int W, int C, int num_slabs) {
  constexpr int PAD = KS / 2;
  const int cvecs = C >> 3;
  const long long unit = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int cv = static_cast<int>(unit % cvecs);
  const int ky = static_cast<int>((unit / cvecs) % KS);
  const long long slab = unit / (static_cast<long long>(cvecs) * KS);
  if (slab >= num_slabs) return;
  const int strips = (W + kDwTW - 1) / kDwTW;
  const long long items = static_cast<long long>(B) * H * strips;
  f32x2 acc[KS][4];
#pragma unroll
  for (int kx = 0; kx < KS; ++kx) { acc[kx][0] = 0ull; acc[kx][1] = 0ull; acc[kx][2] = 0ull; acc[kx][3] = 0ull; }
  for (long long item = slab; item < items; item += num_slabs) {
    const int strip = static_cast<int>(item % strips);
    const int oy = static_cast<int>((item / strips) % H);
    const int b = static_cast<int>(item / (static_cast<long long>(strips) * H));
    const int iy = oy + ky - PAD;
    if (iy < 0 || iy >= H) continue;
    const int ox0 = strip * kDwTW;
    const __nv_bfloat16* grow = dy + ((static_cast<long long>(b) * H + oy) * W) * C + cv * 8;
    const __nv_bfloat16* xrow = x + ((static_cast<long long>(b) * H + iy) * W) * C + cv * 8;
    uint4 g[kDwTW], in[kDwTW + KS - 1];
#pragma unroll
    for (int t = 0; t < kDwTW; ++t) {
      const bool ok = ox0 + t < W;
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(grow + static_cast<long long>(ok ? ox0 + t : 0) * C));
      g[t] = ok ? q : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < kDwTW + KS - 1; ++i) {
      const int ix = ox0 + i - PAD;
      const bool ok = ix >= 0 && ix < W;
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(xrow + static_cast<long long>(ok ? ix : 0) * C));
      in[i] = ok ? q : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int t = 0; t < kDwTW; ++t) {
      const f32x2 g0 = bf16x2_to_f32x2(g[t].x), g1 = bf16x2_to_f32x2(g[t].y), g2 = bf16x2_to_f32x2(g[t].z),
                  g3 = bf16x2_to_f32x2(g[t].w);
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        acc[kx][0] = fma2(g0, bf16x2_to_f32x2(in[t + kx].x), acc[kx][0]);
        acc[kx][1] = fma2(g1, bf16x2_to_f32x2(in[t + kx].y), acc[kx][1]);
        acc[kx][2] = fma2(g2, bf16x2_to_f32x2(in[t + kx].z), acc[kx][2]);
        acc[kx][3] = fma2(g3, bf16x2_to_f32x2(in[t + kx].w), acc[kx][3]);
      }
    }
  }
#pragma unroll
  for (int kx = 0; kx < KS; ++kx) {
    float* dst = dwt + static_cast<long long>(ky * KS + kx) * C + cv * 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float lo, hi;
      upk(acc[kx][j], lo, hi);
      atomicAdd(dst + 2 * j, lo);
      atomicAdd(dst + 2 * j + 1, hi);
    }
  }
}

// inverse of patchify_nhwc: out[b, oy*p + ky, ox*p + kx, c] = patches[(b, oy, ox)][(ky*p + kx)*C + c]
__global__ void unpatchify_nhwc_kernel(const __nv_bfloat16* __restrict__ patches, __nv_bfloat16* __restrict__ out, int B,
                                       int H, int W, int C, int p) {
  const int cvecs = C >> 3;
  const int oh = H / p, ow = W / p;
  const long long total = static_cast<long long>(B) * oh * ow * p * p * cvecs;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cvi = static_cast<int>(idx % cvecs);
    long long r = idx / cvecs;
    const int kx = static_cast<int>(r % p); r /= p;
    const int ky = static_cast<int>(r % p); r /= p;
    const int ox = static_cast<int>(r % ow); r /= ow;
    const int oy = static_cast<int>(r % oh);
    const int b = static_cast<int>(r / oh);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(patches + idx * 8));
    *reinterpret_cast<uint4*>(out + ((static_cast<long long>(b) * H + oy * p + ky) * W + ox * p + kx) * C + cvi * 8) = v;
  }
}

// out[(b, oy, ox)][(ky*p + kx)*C + c] = x[b, oy*p + ky, ox*p + kx, c]   (16-byte vectors along c)
__global__ void patchify_nhwc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int H,
                                     int W, int C, int p) {
  const int cvecs = C >> 3;
  const int oh = H / p, ow = W / p;
  const long long total = static_cast<long long>(B) * oh * ow * p * p * cvecs;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cvi = static_cast<int>(idx % cvecs);
    long long r = idx / cvecs;
    const int kx = static_cast<int>(r % p); r /= p;
    const int ky = static_cast<int>(r % p); r /= p;
    const int ox = static_cast<int>(r % ow); r /= ow;
    const int oy = static_cast<int>(r % oh);
    const int b = static_cast<int>(r / oh);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(
        x + ((static_cast<long long>(b) * H + oy * p + ky) * W + ox * p + kx) * C + cvi * 8));
    *reinterpret_cast<uint4*>(out + idx * 8) = v;   // idx enumerates the output in storage order
  }
}

// out[b][c] = mean over hw of x[b][hw][c]   (fp32 accumulation), one thread per (b, 8 channels)
__global__ void avgpool_nhwc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int HW,
                                    int C) {
  const int cvecs = C >> 3;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(B) * cvecs) return;
  const int cvi = static_cast<int>(idx % cvecs);
  const int b = static_cast<int>(idx / cvecs);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const __nv_bfloat16* p = x + static_cast<long long>(b) * HW * C + cvi * 8;
  for (int i = 0; i < HW; ++i) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p + static_cast<long long>(i) * C));
    acc[0] += bf16_lo(q.x); acc[1] += bf16_hi(q.x); acc[2] += bf16_lo(q.y); acc[3] += bf16_hi(q.y);
    acc[4] += bf16_lo(q.z); acc[5] += bf16_hi(q.z); acc[6] += bf16_lo(q.w); acc[7] += bf16_hi(q.w);
  }
  const float s = 1.0f / static_cast<float>(HW);
  *reinterpret_cast<uint4*>(out + static_cast<long long>(b) * C + cvi * 8) =
      make_uint4(pack_bf16(acc[0] * s, acc[1] * s), pack_bf16(acc[2] * s, acc[3] * s),
                 pack_bf16(acc[4] * s, acc[5] * s), pack_bf16(acc[6] * s, acc[7] * s));
}

template <int KS>
static int launch_dwconv(const void* x, const void* wt, const void* bias, const void* residual, void* y, int B, int H,
                         int W, int C, int bias_f32, cudaStream_t st) {
  const long long total = static_cast<long long>(B) * H * ((W + kDwTW - 1) / kDwTW) * (C / 8);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  dwconv_nhwc_kernel<KS><<<static_cast<int>(blocks), threads, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(wt), bias,
      reinterpret_cast<const __nv_bfloat16*>(residual), reinterpret_cast<__nv_bfloat16*>(y), B, H, W, C, bias_f32);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200

namespace b200 {
int launch_dwconv_tma(const void* x, const void* wt, const void* bias, const void* residual, void* y, int B, int H, int W,
                      int C, int ksize, int bias_f32, cudaStream_t st);
int launch_dwconv_wgrad_tma(const void* x, const void* dy, float* dwt, int B, int H, int W, int C, int ksize,
                            cudaStream_t st);
int launch_dwconv_tma_ln(const void* x, const void* wt, const void* bias, void* y, float* ln_stats, int B, int H, int W, int C,
                         int ksize, int bias_f32, cudaStream_t st);
}
using namespace b200;

extern "C" int b200_dwconv_nhwc_ln(const void* x, const void* weight_t, const void* bias, void* y, float* ln_stats_out,
                                   int32_t B, int32_t H, int32_t W, int32_t C, int32_t ksize, int32_t bias_dtype, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(x && weight_t && y && ln_stats_out, "b200_dwconv_nhwc_ln: null pointer");
  B200_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0, "b200_dwconv_nhwc_ln: bad shape");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(ln_stats_out) & 7) == 0,
                 "b200_dwconv_nhwc_ln: x must be 16-byte, ln_stats_out 8-byte aligned");
  const int rc = launch_dwconv_tma_ln(x, weight_t, bias, y, ln_stats_out, B, H, W, C, ksize, bias_dtype == B200_DTYPE_F32, st);
  if (rc == B200_ERR_UNSUPPORTED)
    return fail(B200_ERR_UNSUPPORTED, "b200_dwconv_nhwc_ln: needs C %% 64 == 0 and a 7 x 7 kernel (got C=%d k=%d)", C, ksize);
  return rc;
}

extern "C" int b200_dwconv_nhwc(const void* x, const void* weight_t, const void* bias, const void* residual, void* y,
                                int32_t B, int32_t H, int32_t W, int32_t C, int32_t ksize, int32_t bias_dtype,
                                void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(x && weight_t && y, "b200_dwconv_nhwc: null pointer");
  B200_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "b200_dwconv_nhwc: bad shape (C %% 8 == 0)");
  const int bf = bias_dtype == B200_DTYPE_F32;
  {
    // C %% 64 == 0, 7 x 7: the TMA-staged kernel (dwconv_tma.cu); B200_DWCONV_DIRECT=1 keeps the direct-load kernel
    static const bool direct = [] { const char* e = getenv("B200_DWCONV_DIRECT"); return e != nullptr && e[0] == '1'; }();
    if (!direct && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
      const int rc = launch_dwconv_tma(x, weight_t, bias, residual, y, B, H, W, C, ksize, bf, st);
      if (rc != B200_ERR_UNSUPPORTED) return rc;
    }
  }
  switch (ksize) {
    case 7: return launch_dwconv<7>(x, weight_t, bias, residual, y, B, H, W, C, bf, st);
    case 5: return launch_dwconv<5>(x, weight_t, bias, residual, y, B, H, W, C, bf, st);
    case 3: return launch_dwconv<3>(x, weight_t, bias, residual, y, B, H, W, C, bf, st);
    default: return fail(B200_ERR_UNSUPPORTED, "b200_dwconv_nhwc: kernel size %d not supported (3, 5, 7)", ksize);
  }
}

template <int KS>
static int launch_dwconv_wgrad(const void* x, const void* dy, float* dwt, int B, int H, int W, int C, cudaStream_t st) {
  const int cvecs = C / 8;
  const long long items = static_cast<long long>(B) * H * ((W + kDwTW - 1) / kDwTW);
  const long long per_slab = static_cast<long long>(cvecs) * KS;
  long long slabs = (static_cast<long long>(sm_count()) * 8 * 256 + per_slab - 1) / per_slab;
  if (slabs > items) slabs = items;
  if (slabs < 1) slabs = 1;
  const long long threads_total = slabs * per_slab;
  dwconv_nhwc_wgrad_kernel<KS><<<static_cast<unsigned>((threads_total + 255) / 256), 256, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dy), dwt, B, H, W, C,
      static_cast<int>(slabs));
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_dwconv_nhwc_wgrad(const void* x, const void* dy, float* dweight_t, int32_t B, int32_t H, int32_t W,
                                      int32_t C, int32_t ksize, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(x && dy && dweight_t, "b200_dwconv_nhwc_wgrad: null pointer");
  B200_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "b200_dwconv_nhwc_wgrad: bad shape (C %% 8 == 0)");
  {
    static const bool direct = [] { const char* e = getenv("B200_DWCONV_DIRECT"); return e != nullptr && e[0] == '1'; }();
    if (!direct && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0) {
      const int rc = launch_dwconv_wgrad_tma(x, dy, dweight_t, B, H, W, C, ksize, st);
      if (rc != B200_ERR_UNSUPPORTED) return rc;
    }
  }
  switch (ksize) {
    case 7: return launch_dwconv_wgrad<7>(x, dy, dweight_t, B, H, W, C, st);
    case 5: return launch_dwconv_wgrad<5>(x, dy, dweight_t, B, H, W, C, st);
    case 3: return launch_dwconv_wgrad<3>(x, dy, dweight_t, B, H, W, C, st);
    default: return fail(B200_ERR_UNSUPPORTED, "b200_dwconv_nhwc_wgrad: kernel size %d not supported (3, 5, 7)", ksize);
  }
}

extern "C" int b200_unpatchify_nhwc(const void* patches, void* out, int32_t B, int32_t H, int32_t W, int32_t C, int32_t p,
                                    void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(patches && out, "b200_unpatchify_nhwc: null pointer");
  B200_CHECK_ARG(B > 0 && C > 0 && C % 8 == 0 && p > 0 && H % p == 0 && W % p == 0,
                 "b200_unpatchify_nhwc: bad shape (C %% 8 == 0, H, W divisible by p)");
  const long long total = static_cast<long long>(B) * H * W * (C / 8);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  unpatchify_nhwc_kernel<<<static_cast<int>(blocks), threads, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(patches),
                                                                        reinterpret_cast<__nv_bfloat16*>(out), B, H, W, C, p);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_patchify_nhwc(const void* x, void* out, int32_t B, int32_t H, int32_t W, int32_t C, int32_t p,
                                  void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(x && out, "b200_patchify_nhwc: null pointer");
  B200_CHECK_ARG(B > 0 && C > 0 && C % 8 == 0 && p > 0 && H % p == 0 && W % p == 0,
                 "b200_patchify_nhwc: bad shape (C %% 8 == 0, H, W divisible by p)");
  const long long total = static_cast<long long>(B) * H * W * (C / 8);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  patchify_nhwc_kernel<<<static_cast<int>(blocks), threads, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                      reinterpret_cast<__nv_bfloat16*>(out), B, H, W, C, p);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_avgpool_nhwc(const void* x, void* out, int32_t B, int32_t HW, int32_t C, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(x && out, "b200_avgpool_nhwc: null pointer");
  B200_CHECK_ARG(B > 0 && HW > 0 && C > 0 && C % 8 == 0, "b200_avgpool_nhwc: bad shape (C %% 8 == 0)");
  const long long total = static_cast<long long>(B) * (C / 8);
  const int threads = 128;
  avgpool_nhwc_kernel<<<static_cast<int>((total + threads - 1) / threads), threads, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(out), B, HW, C);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}
