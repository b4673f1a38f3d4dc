// ConvNeXt-specific kernels on channels-last (NHWC == [B*H*W, C] token-major) activations:
//   dwconv_nhwc      depthwise k x k convolution (k odd <= 7, stride 1, zero padding k/2) + bias
//                    replaces F.conv2d(groups=C) at timm/models/convnext.py:173-182,200 (and the permutes of :205,208:
//                    the whole network stays NHWC so the following LayerNorm / Linear read contiguous rows)
//   patchify_nhwc    non-overlapping p x p gather for the 2x2/s2 downsample conv (convnext.py:263-275)
//   avgpool_nhwc     global average pool over H*W (classifier.py:211-213)
// The depthwise conv is FMA-bound in fp32 (49 MACs per output element, 0.23 GMAC/img for ConvNeXt-B), not HBM-bound;
// each thread keeps a strip of TW output pixels x 8 channels in registers and slides along the row so every loaded
// input vector feeds up to 7 taps.
#include "common.h"
#include "ptx.cuh"

namespace b200 {

constexpr int kDwTW = 7;   // output pixels per thread along W

template <int KS>
__global__ void __launch_bounds__(64)
dwconv_nhwc_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                   const void* __restrict__ bias, __nv_bfloat16* __restrict__ y, int B, int H, int W, int C,
                   int bias_f32) {
  // weights of this CTA's 8-channel groups, transposed to [tap][channel] so a thread reads 16 B per tap
  extern __shared__ __align__(16) __nv_bfloat16 sw[];   // [KS*KS][cvec_per_cta * 8]
  constexpr int PAD = KS / 2;
  const int cvecs = C >> 3;
  const int cv_per_cta = blockDim.x;                       // one 8-channel group per thread in x
  const int cv0 = blockIdx.x * cv_per_cta;
  const int cv = cv0 + threadIdx.x;
  for (int i = threadIdx.x; i < KS * KS * cv_per_cta * 8; i += blockDim.x) {
    const int tap = i / (cv_per_cta * 8);
    const int cl = i - tap * (cv_per_cta * 8);
    const int c = cv0 * 8 + cl;
    sw[i] = (c < C) ? w[static_cast<long long>(c) * KS * KS + tap] : __float2bfloat16(0.f);
  }
  __syncthreads();
  if (cv >= cvecs) return;
  const int strips = (W + kDwTW - 1) / kDwTW;
  const long long total = static_cast<long long>(B) * H * strips;
  float bv[8];
  if (bias != nullptr) {
    if (bias_f32) {
      const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(bias) + cv * 8);
      const float4 a = __ldg(p), b2 = __ldg(p + 1);
      bv[0] = a.x; bv[1] = a.y; bv[2] = a.z; bv[3] = a.w; bv[4] = b2.x; bv[5] = b2.y; bv[6] = b2.z; bv[7] = b2.w;
    } else {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(bias) + cv * 8));
      bv[0] = bf16_lo(q.x); bv[1] = bf16_hi(q.x); bv[2] = bf16_lo(q.y); bv[3] = bf16_hi(q.y);
      bv[4] = bf16_lo(q.z); bv[5] = bf16_hi(q.z); bv[6] = bf16_lo(q.w); bv[7] = bf16_hi(q.w);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) bv[j] = 0.f;
  }
  for (long long item = blockIdx.y; item < total; item += gridDim.y) {
    const int strip = static_cast<int>(item % strips);
    const int oy = static_cast<int>((item / strips) % H);
    const int b = static_cast<int>(item / (static_cast<long long>(strips) * H));
    const int ox0 = strip * kDwTW;
    f32x2 acc[kDwTW][4];
#pragma unroll
    for (int t = 0; t < kDwTW; ++t) {
      acc[t][0] = pk(bv[0], bv[1]); acc[t][1] = pk(bv[2], bv[3]);
      acc[t][2] = pk(bv[4], bv[5]); acc[t][3] = pk(bv[6], bv[7]);
    }
#pragma unroll 1
    for (int ky = 0; ky < KS; ++ky) {
      const int iy = oy + ky - PAD;
      if (iy < 0 || iy >= H) continue;
      f32x2 wt[KS][4];
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        const uint4 q = *reinterpret_cast<const uint4*>(sw + (ky * KS + kx) * (cv_per_cta * 8) + threadIdx.x * 8);
        wt[kx][0] = bf16x2_to_f32x2(q.x); wt[kx][1] = bf16x2_to_f32x2(q.y);
        wt[kx][2] = bf16x2_to_f32x2(q.z); wt[kx][3] = bf16x2_to_f32x2(q.w);
      }
      const __nv_bfloat16* row = x + ((static_cast<long long>(b) * H + iy) * W) * C + cv * 8;
#pragma unroll
      for (int i = 0; i < kDwTW + KS - 1; ++i) {
        const int ix = ox0 + i - PAD;
        if (ix < 0 || ix >= W) continue;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(row + static_cast<long long>(ix) * C));
        const f32x2 v0 = bf16x2_to_f32x2(q.x), v1 = bf16x2_to_f32x2(q.y), v2 = bf16x2_to_f32x2(q.z), v3 = bf16x2_to_f32x2(q.w);
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          const int t = i - kx;   // output pixel (within the strip) this input contributes to through tap kx
          if (t >= 0 && t < kDwTW) {
            acc[t][0] = fma2(v0, wt[kx][0], acc[t][0]);
            acc[t][1] = fma2(v1, wt[kx][1], acc[t][1]);
            acc[t][2] = fma2(v2, wt[kx][2], acc[t][2]);
            acc[t][3] = fma2(v3, wt[kx][3], acc[t][3]);
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
static int launch_dwconv(const void* x, const void* w, const void* bias, void* y, int B, int H, int W, int C,
                         int bias_f32, cudaStream_t st) {
  const int cvecs = C / 8;
  int threads = cvecs < 64 ? ((cvecs + 31) / 32) * 32 : 64;
  const int gx = (cvecs + threads - 1) / threads;
  const long long items = static_cast<long long>(B) * H * ((W + kDwTW - 1) / kDwTW);
  long long gy = items;
  const long long cap = static_cast<long long>(sm_count()) * 16 / gx + 1;
  if (gy > cap) gy = cap;
  if (gy > 65535) gy = 65535;
  const size_t smem = static_cast<size_t>(KS) * KS * threads * 8 * sizeof(__nv_bfloat16);
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(dwconv_nhwc_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      KS * KS * 64 * 8 * static_cast<int>(sizeof(__nv_bfloat16))));
    attr_set = true;
  }
  dwconv_nhwc_kernel<KS><<<dim3(gx, static_cast<unsigned>(gy)), threads, smem, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(w), bias,
      reinterpret_cast<__nv_bfloat16*>(y), B, H, W, C, bias_f32);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_dwconv_nhwc(const void* x, const void* weight, const void* bias, void* y, int32_t B, int32_t H,
                                int32_t W, int32_t C, int32_t ksize, int32_t bias_dtype, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(x && weight && y, "b200_dwconv_nhwc: null pointer");
  B200_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "b200_dwconv_nhwc: bad shape (C %% 8 == 0)");
  const int bf = bias_dtype == B200_DTYPE_F32;
  switch (ksize) {
    case 7: return launch_dwconv<7>(x, weight, bias, y, B, H, W, C, bf, st);
    case 5: return launch_dwconv<5>(x, weight, bias, y, B, H, W, C, bf, st);
    case 3: return launch_dwconv<3>(x, weight, bias, y, B, H, W, C, bf, st);
    default: return fail(B200_ERR_UNSUPPORTED, "b200_dwconv_nhwc: kernel size %d not supported (3, 5, 7)", ksize);
  }
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
