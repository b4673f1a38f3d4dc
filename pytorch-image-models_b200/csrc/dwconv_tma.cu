// Depthwise KS x KS convolution on channels-last rows, TMA-staged (the ConvNeXtBlock token mixer,
// timm/models/convnext.py:142-149, 201: nn.Conv2d(dim, dim, 7, padding=3, groups=dim)).
//
// Persistent CTAs walk (channel block of 64, image, 14 x 14 output tile) work items.  One thread issues a 4-D TMA
// load {64 channels, 20 x 20 input pixels, 1 image} of the NEXT item into the other shared-memory buffer while the
// current one is computed; the halo outside the image arrives as zeros from the TMA unit (negative / past-the-end
// coordinates), so the arithmetic carries no border predicates at all.  The 49 x 64 weights of the channel block
// sit in shared memory as fp32.  A thread owns 8 channels x 7 consecutive output pixels of one row: per kernel row it
// reads 13 input vectors (16 bytes = 8 channels) from shared memory, widens them once and feeds up to 7 taps each
// (196 packed-fp32 FFMA2 per kernel row); quarter-warps read the 8 channel vectors of one pixel = 128 contiguous
// bytes, so the reads are bank-conflict free.  The kernel is FMA-issue-bound (49 MACs per output element; HBM
// traffic is one read + one write of the activation), the staging only keeps the FMA pipe fed.
//
// With `residual` the result is added to it: the data gradient of the conv is the same kernel on the spatially
// flipped weights, with the shortcut-branch gradient folded in (convnext.py:213).
#include "common.h"
#include "ptx.cuh"

namespace b200 {

namespace {
constexpr int kCB = 64;                 // channels per work item (128 B per pixel in shared memory)
constexpr int kTH = 14, kTWt = 14;      // output tile
constexpr int kSW = 7;                  // output pixels per thread along W
constexpr int kThreads = kTH * (kTWt / kSW) * (kCB / 8);   // 224
}  // namespace

template <int KS>
struct DwTmaSmem {
  static constexpr int kIH = kTH + KS - 1, kIW = kTWt + KS - 1;
  static constexpr int kInBytes = kIH * kIW * kCB * 2;                    // 51200 for KS = 7
  static constexpr int kInStride = (kInBytes + 1023) / 1024 * 1024;
  static constexpr int kWOff = 2 * kInStride;
  static constexpr int kWBytes = KS * KS * kCB * 4;                       // fp32 weights of the channel block
  static constexpr int kBarOff = kWOff + (kWBytes + 15) / 16 * 16;
  static constexpr int kTotal = kBarOff + 64 + 128;   // two CTAs per SM: 2 x (115 KB + 1 KB reserved) <= 228 KB
};

struct DwTmaParams {
  const __nv_bfloat16* wt;        // [KS*KS][C]
  const void* bias;               // [C] or null
  const __nv_bfloat16* residual;  // [B,H,W,C] or null
  __nv_bfloat16* y;
  int B, H, W, C, bias_f32;
  int tiles_y, tiles_x, cblocks;
  long long num_items;
  // STATS instantiation (b200_dwconv_nhwc_ln): fp32 [C / 64 + 1][B*H*W][2] -- per pixel and 64-channel block the sum and
  // the sum of squares of the bf16 values stored, the producer side of a LayerNorm folded into the next GEMM (gemm.cu)
  float* ln_stats;
};

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 q;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(addr));
  return q;
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <int KS, bool STATS = false>
__global__ void __launch_bounds__(kThreads, 2)
dwconv_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const DwTmaParams p) {
  using L = DwTmaSmem<KS>;
  constexpr int PAD = KS / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  float* sw = reinterpret_cast<float*>(smem + L::kWOff);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* bar_empty = bar_full + 2;     // every thread has finished reading the buffer (count = kThreads)

  const int tid = threadIdx.x;
  const int cv = tid & 7;                    // 8-channel vector inside the channel block
  const int strip = (tid >> 3) & 1;          // which 7-pixel half of the tile row
  const int ty = tid >> 4;                   // output row inside the tile

  if (tid == 0) {
    tma_prefetch_desc(&tm_x);
    mbar_init(&bar_full[0], 1);
    mbar_init(&bar_full[1], 1);
    mbar_init(&bar_empty[0], kThreads);
    mbar_init(&bar_empty[1], kThreads);
    fence_barrier_init();
  }
  __syncthreads();

  // item -> (channel block, image, tile row, tile column); channel block outermost so the weights are staged rarely
  const long long per_cblock = static_cast<long long>(p.B) * p.tiles_y * p.tiles_x;
  auto decode = [&](long long item, int& cb, int& b, int& oy0, int& ox0) {
    cb = static_cast<int>(item / per_cblock);
    long long r = item - cb * per_cblock;
    b = static_cast<int>(r / (p.tiles_y * p.tiles_x));
    r -= static_cast<long long>(b) * (p.tiles_y * p.tiles_x);
    const int t_y = static_cast<int>(r / p.tiles_x);
    oy0 = t_y * kTH;
    ox0 = static_cast<int>(r - static_cast<long long>(t_y) * p.tiles_x) * kTWt;
  };
  auto issue = [&](long long item, int buf) {
    int cb, b, oy0, ox0;
    decode(item, cb, b, oy0, ox0);
    mbar_arrive_expect_tx(&bar_full[buf], L::kInBytes);
    tma_load_4d(smem + buf * L::kInStride, &tm_x, &bar_full[buf], cb * kCB, ox0 - PAD, oy0 - PAD, b);
  };

  const long long first = blockIdx.x, step = gridDim.x;
  if (tid == 0) {
    if (first < p.num_items) issue(first, 0);
    if (first + step < p.num_items) issue(first + step, 1);
  }

  int cur_cb = -1;
  int it = 0;
  for (long long item = first; item < p.num_items; item += step, ++it) {
    const int buf = it & 1;
    int cb, b, oy0, ox0;
    decode(item, cb, b, oy0, ox0);
    if (cb != cur_cb) {
      // (re)stage the channel block's weights as fp32 once every thread is done with the previous block's
      if (cur_cb >= 0) __syncthreads();
      for (int i = tid; i < KS * KS * (kCB / 8); i += kThreads) {
        const int tap = i / (kCB / 8), v = i - tap * (kCB / 8);
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(p.wt + static_cast<long long>(tap) * p.C + cb * kCB + v * 8));
        // layout [tap][half][vector][4 floats]: the 8 vectors' first (second) halves are 128 contiguous bytes, so a
        // quarter-warp's 16-byte reads touch every bank once (a plain [tap][64] layout is a 2-way conflict)
        float4* dst = reinterpret_cast<float4*>(sw + tap * kCB) + v;
        dst[0] = make_float4(bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y));
        dst[8] = make_float4(bf16_lo(q.z), bf16_hi(q.z), bf16_lo(q.w), bf16_hi(q.w));
      }
      cur_cb = cb;
      __syncthreads();
    }
    mbar_wait(&bar_full[buf], (it >> 1) & 1);

    const int c = cb * kCB + cv * 8;
    const int oy = oy0 + ty;
    const int oxs = ox0 + strip * kSW;
    f32x2 acc[kSW][4];
    {
      f32x2 b0 = 0ull, b1 = 0ull, b2 = 0ull, b3 = 0ull;
      if (p.bias != nullptr) {
        if (p.bias_f32) {
          const float4* bp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.bias) + c);
          const float4 a = __ldg(bp), a2 = __ldg(bp + 1);
          b0 = pk(a.x, a.y); b1 = pk(a.z, a.w); b2 = pk(a2.x, a2.y); b3 = pk(a2.z, a2.w);
        } else {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.bias) + c));
          b0 = bf16x2_to_f32x2(q.x); b1 = bf16x2_to_f32x2(q.y); b2 = bf16x2_to_f32x2(q.z); b3 = bf16x2_to_f32x2(q.w);
        }
      }
#pragma unroll
      for (int t = 0; t < kSW; ++t) { acc[t][0] = b0; acc[t][1] = b1; acc[t][2] = b2; acc[t][3] = b3; }
    }
    const uint32_t tile_addr = smem_u32(smem + buf * L::kInStride);
    const uint32_t sw_addr = smem_u32(sw);
#pragma unroll 1
    for (int ky = 0; ky < KS; ++ky) {
      f32x2 wv[KS][4];
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        const uint32_t wa = sw_addr + ((ky * KS + kx) * kCB + cv * 4) * 4;
        const uint4 a = lds128(wa), a2 = lds128(wa + 128);
        wv[kx][0] = pk(__uint_as_float(a.x), __uint_as_float(a.y)); wv[kx][1] = pk(__uint_as_float(a.z), __uint_as_float(a.w));
        wv[kx][2] = pk(__uint_as_float(a2.x), __uint_as_float(a2.y)); wv[kx][3] = pk(__uint_as_float(a2.z), __uint_as_float(a2.w));
      }
      const uint32_t row = tile_addr + (((ty + ky) * L::kIW + strip * kSW) * kCB + cv * 8) * 2;
#pragma unroll
      for (int i = 0; i < kSW + KS - 1; ++i) {
        const uint4 q = lds128(row + i * (kCB * 2));
        const f32x2 v0 = bf16x2_to_f32x2(q.x), v1 = bf16x2_to_f32x2(q.y), v2 = bf16x2_to_f32x2(q.z), v3 = bf16x2_to_f32x2(q.w);
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          const int t = i - kx;   // output pixel (within the strip) this input feeds through tap kx
          if (t >= 0 && t < kSW) {
            acc[t][0] = fma2(v0, wv[kx][0], acc[t][0]);
            acc[t][1] = fma2(v1, wv[kx][1], acc[t][1]);
            acc[t][2] = fma2(v2, wv[kx][2], acc[t][2]);
            acc[t][3] = fma2(v3, wv[kx][3], acc[t][3]);
          }
        }
      }
    }
    // this thread is done with the buffer; thread 0 refills it with the item after next once everyone is (the other
    // warps go straight on to their stores and the next item instead of meeting at a CTA barrier)
    mbar_arrive(&bar_empty[buf]);
    if (tid == 0 && item + 2 * step < p.num_items) {
      mbar_wait(&bar_empty[buf], (it >> 1) & 1);
      issue(item + 2 * step, buf);
    }

    float st_s[kSW], st_q[kSW];      // STATS: this thread's 8 channels of each of its pixels
    if constexpr (STATS) {
#pragma unroll
      for (int t = 0; t < kSW; ++t) st_s[t] = st_q[t] = 0.f;
    }
    if (oy < p.H) {
      const long long rowoff = ((static_cast<long long>(b) * p.H + oy) * p.W) * p.C + c;
#pragma unroll
      for (int t = 0; t < kSW; ++t) {
        const int ox = oxs + t;
        if (ox < p.W) {
          const long long off = rowoff + static_cast<long long>(ox) * p.C;
          if (p.residual != nullptr) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(p.residual + off));
            acc[t][0] = add2(acc[t][0], bf16x2_to_f32x2(q.x)); acc[t][1] = add2(acc[t][1], bf16x2_to_f32x2(q.y));
            acc[t][2] = add2(acc[t][2], bf16x2_to_f32x2(q.z)); acc[t][3] = add2(acc[t][3], bf16x2_to_f32x2(q.w));
          }
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float lo, hi;
            upk(acc[t][j], lo, hi);
            o[j] = pack_bf16(lo, hi);
          }
          *reinterpret_cast<uint4*>(p.y + off) = make_uint4(o[0], o[1], o[2], o[3]);
          if constexpr (STATS) {
            f32x2 s2 = splat(0.f), q2 = splat(0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const f32x2 v = bf16x2_to_f32x2(o[j]);        // exactly the values the LayerNorm would read
              s2 = add2(s2, v);
              q2 = fma2(v, v, q2);
            }
            float a0, a1, b0, b1;
            upk(s2, a0, a1);
            upk(q2, b0, b1);
            st_s[t] = a0 + a1;
            st_q[t] = b0 + b1;
          }
        }
      }
    }
    if constexpr (STATS) {
      // the 8 lanes cv = 0..7 of a (row, strip) hold the 64 channels of the same pixels: butterfly over them (all 32 lanes
      // take part: kThreads is a whole number of warps), lane cv == 0 stores the block's (sum, sum of squares) per pixel
      const long long rows_total = static_cast<long long>(p.B) * p.H * p.W;
      float2* dst = reinterpret_cast<float2*>(p.ln_stats) + static_cast<long long>(cb) * rows_total +
                    (static_cast<long long>(b) * p.H + oy) * p.W;
#pragma unroll
      for (int t = 0; t < kSW; ++t) {
        float sv = st_s[t], qv = st_q[t];
#pragma unroll
        for (int m = 1; m < 8; m <<= 1) {
          sv += __shfl_xor_sync(0xffffffffu, sv, m);
          qv += __shfl_xor_sync(0xffffffffu, qv, m);
        }
        if (cv == 0 && oy < p.H && oxs + t < p.W) dst[oxs + t] = make_float2(sv, qv);
      }
    }
  }
}

template <int KS, bool STATS = false>
static int launch_dwconv_tma_ks(const void* x, const void* wt, const void* bias, const void* residual, void* y, int B, int H,
                                int W, int C, int bias_f32, cudaStream_t st, float* ln_stats = nullptr) {
  using L = DwTmaSmem<KS>;
  CUtensorMap tm;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    uint64_t strides[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(W) * C * 2, static_cast<uint64_t>(H) * W * C * 2};
    uint32_t box[4] = {kCB, static_cast<uint32_t>(L::kIW), static_cast<uint32_t>(L::kIH), 1};
    int rc = encode_tmap_bf16(&tm, x, 4, dims, strides, box, 0);
    if (rc) return rc;
  }
  DwTmaParams p;
  p.wt = reinterpret_cast<const __nv_bfloat16*>(wt); p.bias = bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual); p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.B = B; p.H = H; p.W = W; p.C = C; p.bias_f32 = bias_f32;
  p.tiles_y = (H + kTH - 1) / kTH; p.tiles_x = (W + kTWt - 1) / kTWt; p.cblocks = C / kCB;
  p.num_items = static_cast<long long>(p.cblocks) * B * p.tiles_y * p.tiles_x;
  p.ln_stats = ln_stats;
  auto kern = dwconv_tma_kernel<KS, STATS>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  long long grid = static_cast<long long>(sm_count()) * 2;
  if (grid > p.num_items) grid = p.num_items;
  kern<<<static_cast<int>(grid), kThreads, L::kTotal, st>>>(tm, p);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

// Returns B200_ERR_UNSUPPORTED (without setting the error string) when the shape is outside this kernel's cover
// (C %% 64 != 0 or a kernel size other than 7): the caller falls back to the direct-load kernel in convnext.cu.
int launch_dwconv_tma(const void* x, const void* wt, const void* bias, const void* residual, void* y, int B, int H, int W,
                      int C, int ksize, int bias_f32, cudaStream_t st) {
  if (C % kCB != 0 || ksize != 7) return B200_ERR_UNSUPPORTED;
  return launch_dwconv_tma_ks<7>(x, wt, bias, residual, y, B, H, W, C, bias_f32, st);
}

// The same convolution + per-pixel partial LayerNorm statistics (ln_stats: fp32 [C / 64 + 1][B*H*W][2], slot = channel block).
int launch_dwconv_tma_ln(const void* x, const void* wt, const void* bias, void* y, float* ln_stats, int B, int H, int W, int C,
                         int ksize, int bias_f32, cudaStream_t st) {
  if (C % kCB != 0 || ksize != 7) return B200_ERR_UNSUPPORTED;
  return launch_dwconv_tma_ks<7, true>(x, wt, bias, nullptr, y, B, H, W, C, bias_f32, st, ln_stats);
}


// ------------------------------------------------------------------------------------------------ weight gradient
// dw_t[ky][kx][c] += sum_{b,oy,ox} dy[b,oy,ox,c] * x[b, oy+ky-3, ox+kx-3, c]   (fp32, atomically accumulated).
// Same work items as the forward kernel; per item two TMA loads (the 20 x 20 x-tile with its zero halo and the
// 14 x 14 dy-tile, whose part past the image edge also arrives as zeros -- no border predicates).  A thread owns
// (8 channels, one kernel row ky, two or three output rows of the tile) and keeps the 7 x 8 partial sums of its kernel row in
// registers across ALL items of a channel block: per output row it reads 14 dy vectors and 20 x vectors from shared
// memory and issues 392 FFMA2 (a 7-wide sliding window of widened x vectors, fully unrolled so the window lives in
// registers).  At a channel-block change the partials are combined through shared-memory atomics and leave as one
// global fp32 atomic per weight and CTA.
namespace {
constexpr int kWgGroups = 6;                             // row groups: rows {0-2, 3-5, 6-7, 8-9, 10-11, 12-13}
constexpr int kWgThreads = kWgGroups * 7 * (kCB / 8);    // 336 (KS = 7): 11 warps -> a 168-register budget, no spills
}  // namespace

struct DwWgSmem {
  static constexpr int kXBytes = 20 * 20 * kCB * 2;      // 51200
  static constexpr int kDyBytes = kTH * kTWt * kCB * 2;  // 25088
  static constexpr int kStage = (kXBytes + kDyBytes + 1023) / 1024 * 1024;   // 76800
  static constexpr int kAccOff = 2 * kStage;
  static constexpr int kAccBytes = 49 * kCB * 4;
  static constexpr int kBarOff = kAccOff + kAccBytes;
  static constexpr int kTotal = kBarOff + 64 + 128;
};

struct DwWgParams {
  float* dwt;   // [49][C] fp32
  int B, H, W, C;
  int tiles_y, tiles_x, cblocks;
  long long num_items;
};

__global__ void __launch_bounds__(kWgThreads, 1)
dwconv_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dy,
                        const DwWgParams p) {
  using L = DwWgSmem;
  constexpr int KS = 7, PAD = 3, IW = 20;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  float* sacc = reinterpret_cast<float*>(smem + L::kAccOff);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* bar_empty = bar_full + 2;

  const int tid = threadIdx.x;
  const int cv = tid & 7;
  const int ky = (tid >> 3) % KS;
  const int grp = tid / (8 * KS);
  const int row_begin = grp < 2 ? 3 * grp : 2 * grp + 2;     // first output row of this group
  const int row_count = grp < 2 ? 3 : 2;

  if (tid == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_dy);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], kWgThreads);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < 49 * kCB; i += kWgThreads) sacc[i] = 0.f;
  __syncthreads();

  const long long per_cblock = static_cast<long long>(p.B) * p.tiles_y * p.tiles_x;
  auto decode = [&](long long item, int& cb, int& b, int& oy0, int& ox0) {
    cb = static_cast<int>(item / per_cblock);
    long long r = item - cb * per_cblock;
    b = static_cast<int>(r / (p.tiles_y * p.tiles_x));
    r -= static_cast<long long>(b) * (p.tiles_y * p.tiles_x);
    const int t_y = static_cast<int>(r / p.tiles_x);
    oy0 = t_y * kTH;
    ox0 = static_cast<int>(r - static_cast<long long>(t_y) * p.tiles_x) * kTWt;
  };
  auto issue = [&](long long item, int buf) {
    int cb, b, oy0, ox0;
    decode(item, cb, b, oy0, ox0);
    uint8_t* st = smem + buf * L::kStage;
    mbar_arrive_expect_tx(&bar_full[buf], L::kXBytes + L::kDyBytes);
    tma_load_4d(st, &tm_x, &bar_full[buf], cb * kCB, ox0 - PAD, oy0 - PAD, b);
    tma_load_4d(st + L::kXBytes, &tm_dy, &bar_full[buf], cb * kCB, ox0, oy0, b);
  };
  // per-thread partial sums -> shared accumulator -> one global atomic per weight
  f32x2 acc[KS][4];
#define DW_CLEAR_ACC()                                   \
  _Pragma("unroll") for (int kx_ = 0; kx_ < KS; ++kx_)   \
      _Pragma("unroll") for (int j_ = 0; j_ < 4; ++j_) acc[kx_][j_] = 0ull;
#define DW_FLUSH(cb_)                                                                              \
  {                                                                                                \
    _Pragma("unroll") for (int kx_ = 0; kx_ < KS; ++kx_) {                                         \
      float* dst_ = sacc + (ky * KS + kx_) * kCB + cv * 8;                                         \
      _Pragma("unroll") for (int j_ = 0; j_ < 4; ++j_) {                                           \
        float lo_, hi_;                                                                            \
        upk(acc[kx_][j_], lo_, hi_);                                                               \
        atomicAdd(dst_ + 2 * j_, lo_);                                                             \
        atomicAdd(dst_ + 2 * j_ + 1, hi_);                                                         \
      }                                                                                            \
    }                                                                                              \
    __syncthreads();                                                                               \
    for (int i_ = tid; i_ < 49 * kCB; i_ += kWgThreads) {                                          \
      const int tap_ = i_ / kCB, c_ = i_ - tap_ * kCB;                                             \
      atomicAdd(p.dwt + static_cast<long long>(tap_) * p.C + (cb_) * kCB + c_, sacc[i_]);          \
      sacc[i_] = 0.f;                                                                              \
    }                                                                                              \
    __syncthreads();                                                                               \
  }

  const long long first = blockIdx.x, step = gridDim.x;
  if (tid == 0) {
    if (first < p.num_items) issue(first, 0);
    if (first + step < p.num_items) issue(first + step, 1);
  }
  DW_CLEAR_ACC();
  int cur_cb = -1;
  int it = 0;
  for (long long item = first; item < p.num_items; item += step, ++it) {
    const int buf = it & 1;
    int cb, b, oy0, ox0;
    decode(item, cb, b, oy0, ox0);
    if (cb != cur_cb) {
      if (cur_cb >= 0) {
        DW_FLUSH(cur_cb);
        DW_CLEAR_ACC();
      }
      cur_cb = cb;
    }
    mbar_wait(&bar_full[buf], (it >> 1) & 1);
    const uint32_t xt = smem_u32(smem + buf * L::kStage);
    const uint32_t dyt = xt + L::kXBytes;
#pragma unroll 1
    for (int rr = 0; rr < row_count; ++rr) {
      const int orow = row_begin + rr;
      const uint32_t xrow = xt + (((orow + ky) * IW) * kCB + cv * 8) * 2;
      const uint32_t dyrow = dyt + ((orow * kTWt) * kCB + cv * 8) * 2;
      f32x2 win[KS][4];   // widened x vectors at columns t .. t+6 of the input row
#pragma unroll
      for (int i = 0; i < KS - 1; ++i) {
        const uint4 q = lds128(xrow + i * (kCB * 2));
        win[i][0] = bf16x2_to_f32x2(q.x); win[i][1] = bf16x2_to_f32x2(q.y);
        win[i][2] = bf16x2_to_f32x2(q.z); win[i][3] = bf16x2_to_f32x2(q.w);
      }
#pragma unroll 1
      for (int t0 = 0; t0 < kTWt; t0 += KS)
#pragma unroll
      for (int tt = 0; tt < KS; ++tt) {
        const int t = t0 + tt;      // t % KS == tt: the window slots below are compile-time indices
        {
          const uint4 q = lds128(xrow + (t + KS - 1) * (kCB * 2));
          win[(tt + KS - 1) % KS][0] = bf16x2_to_f32x2(q.x); win[(tt + KS - 1) % KS][1] = bf16x2_to_f32x2(q.y);
          win[(tt + KS - 1) % KS][2] = bf16x2_to_f32x2(q.z); win[(tt + KS - 1) % KS][3] = bf16x2_to_f32x2(q.w);
        }
        const uint4 d = lds128(dyrow + t * (kCB * 2));
        const f32x2 d0 = bf16x2_to_f32x2(d.x), d1 = bf16x2_to_f32x2(d.y), d2 = bf16x2_to_f32x2(d.z), d3 = bf16x2_to_f32x2(d.w);
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          acc[kx][0] = fma2(d0, win[(tt + kx) % KS][0], acc[kx][0]);
          acc[kx][1] = fma2(d1, win[(tt + kx) % KS][1], acc[kx][1]);
          acc[kx][2] = fma2(d2, win[(tt + kx) % KS][2], acc[kx][2]);
          acc[kx][3] = fma2(d3, win[(tt + kx) % KS][3], acc[kx][3]);
        }
      }
    }
    mbar_arrive(&bar_empty[buf]);
    if (tid == 0 && item + 2 * step < p.num_items) {
      mbar_wait(&bar_empty[buf], (it >> 1) & 1);
      issue(item + 2 * step, buf);
    }
  }
  if (cur_cb >= 0) DW_FLUSH(cur_cb);
#undef DW_CLEAR_ACC
#undef DW_FLUSH
}

// Returns B200_ERR_UNSUPPORTED when the shape is outside this kernel's cover (see launch_dwconv_tma).
int launch_dwconv_wgrad_tma(const void* x, const void* dy, float* dwt, int B, int H, int W, int C, int ksize,
                            cudaStream_t st) {
  if (C % kCB != 0 || ksize != 7) return B200_ERR_UNSUPPORTED;
  using L = DwWgSmem;
  CUtensorMap tx, tdy;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    uint64_t strides[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(W) * C * 2, static_cast<uint64_t>(H) * W * C * 2};
    uint32_t bx[4] = {kCB, 20, 20, 1}, bd[4] = {kCB, kTWt, kTH, 1};
    int rc = encode_tmap_bf16(&tx, x, 4, dims, strides, bx, 0);
    if (rc) return rc;
    rc = encode_tmap_bf16(&tdy, dy, 4, dims, strides, bd, 0);
    if (rc) return rc;
  }
  DwWgParams p;
  p.dwt = dwt; p.B = B; p.H = H; p.W = W; p.C = C;
  p.tiles_y = (H + kTH - 1) / kTH; p.tiles_x = (W + kTWt - 1) / kTWt; p.cblocks = C / kCB;
  p.num_items = static_cast<long long>(p.cblocks) * B * p.tiles_y * p.tiles_x;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(dwconv_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  long long grid = sm_count();
  if (grid > p.num_items) grid = p.num_items;
  dwconv_wgrad_tma_kernel<<<static_cast<int>(grid), kWgThreads, L::kTotal, st>>>(tx, tdy, p);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200
