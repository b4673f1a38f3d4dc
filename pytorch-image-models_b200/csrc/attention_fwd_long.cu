// Attention forward for long sequences (N > 256 tokens: ViT-L/14 @336 = 577, NaFlex batches), head_dim 64, and for
// head_dim 128 at any N (NH = 2: ViT-H 80 / SO400M 72 / 96-wide heads arrive zero-padded to 128 through the weights, see
// functional.head_pad).  A 128-wide head is handled as two 64-wide halves so that every operand sub-tile stays one
// 128-byte swizzle atom per token row: S = Q0 K0^T + Q1 K1^T (8 instead of 4 k-steps into the same accumulator),
// O = [P V0 | P V1] (two N = 64 MMAs into neighbouring TMEM columns); one CTA per SM (192 KB of shared memory).
// One CTA (128 threads = 128 query rows) per (query tile, head, image), two CTAs per SM; online softmax over
// 128-key blocks (flash-attention recurrence), same operand / result layout as attention_fwd.cu.  Differences to the
// single-buffer kernel it replaces on this path:
//   * K / V blocks are double buffered: block j+1 is requested by TMA before block j is consumed;
//   * every key block except the last is processed by an instantiation without any key-mask code (with a run-time
//     mask the compiler emits a compare + select per score element); only the last block masks;
//   * 6 of every 16 column pairs take exp2 on the FMA pipe (Cody-Waite + degree-3 minimax, see attention_fwd2.cu);
//   * the output tile leaves through a swizzled staging tile (the dead Q tile) and one TMA store per warp -- a 3-D map
//     {H*64, N, B} clips the rows past N -- instead of per-lane 16-byte row stores.
// Replaces timm/layers/attention.py:115-135 for these shapes; key-padding masks (naflexvit.py:964-1029) arrive as
// per-sample key counts.
#include <cstdlib>
#include <type_traits>
#include "common.h"
#include "ptx.cuh"

namespace b200 {

namespace {
constexpr int kHd = 64;
constexpr int kQT = 128;                    // query rows per CTA
constexpr int kKB = 128;                    // keys per block
constexpr int kQBytes = kQT * kHd * 2;      // 16384
constexpr int kKBytes = kKB * kHd * 2;      // 16384
constexpr int kPBytes = (kKB / 8) * (kQT * 16);   // 32768: P as bf16 K-major core matrices
// shared-memory layout for NH 64-wide halves of the head: Q (NH sub-tiles), two K buffers, two V buffers (NH sub-tiles each), P
template <int NH>
struct FwdLongSmem {
  static constexpr int kOffQ = 0;
  static constexpr int kOffK = NH * kQBytes;
  static constexpr int kOffV = kOffK + 2 * NH * kKBytes;
  static constexpr int kOffP = kOffV + 2 * NH * kKBytes;
  static constexpr int kOffBar = kOffP + kPBytes;
  static constexpr int kSmem = kOffBar + 64;  // NH = 1: 114752 -- two CTAs per SM need 2 x (kSmem + 1 KB reserved) <= 228 KB, so
                                              // the buffer is declared 1024-byte aligned instead of carrying alignment slack
};
}  // namespace

struct AttnLongParams {
  float* lse;
  const int* kv_len;
  int B, N, H;
  float scale, scale_log2e;
  AttnBias bias;      // BIAS kernels: additive mask / bias on the scaled scores (common.h)
  float inv_scale;
};

__device__ __forceinline__ void exp2_poly_pair_l(f32x2 tt, float& e0, float& e1) {
  float t0, t1;
  upk(tt, t0, t1);
  tt = pk(fmaxf(t0, -126.0f), fmaxf(t1, -126.0f));
  const f32x2 r = add2(tt, splat(12582912.0f));
  const f32x2 f = add2(tt, add2(splat(12582912.0f), mul2(r, splat(-1.0f))));
  f32x2 q = fma2(f, splat(0.05520550534129143f), splat(0.24261397123336792f));
  q = fma2(q, f, splat(0.6932547688484192f));
  q = fma2(q, f, splat(0.9999276995658875f));
  float q0, q1, r0, r1;
  upk(q, q0, q1);
  upk(r, r0, r1);
  e0 = __int_as_float(__float_as_int(q0) + (__float_as_int(r0) << 23));
  e1 = __int_as_float(__float_as_int(q1) + (__float_as_int(r1) << 23));
}

// kPolyMask: which of the 16 column pairs of a 32-column chunk take exp2 on the FMA pipe (0 = all on the MUFU)
// BIAS = true: an additive fp32 bias (causal triangle / general attn_mask, attention.py:21-40) is added to every score tile
// in TMEM before the online-softmax passes read it; a key block that is entirely -inf for a row leaves the row's running
// statistics untouched (P = 0, rescale factor from the old maximum).
template <uint32_t kPolyMask, bool BIAS = false, int NH = 1>
__global__ void __launch_bounds__(128, 3 - NH)
attn_fwd_long_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                     const __grid_constant__ CUtensorMap tm_out, const AttnLongParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using L = FwdLongSmem<NH>;
  constexpr int kOffQ = L::kOffQ, kOffK = L::kOffK, kOffV = L::kOffV, kOffP = L::kOffP, kOffBar = L::kOffBar;
  constexpr int kHD = NH * kHd;               // the (padded) head width
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sP = smem + kOffP;
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* bar_kv = bar_q + 1;     // [2]
  uint64_t* bar_s = bar_q + 3;
  uint64_t* bar_o = bar_q + 4;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_q + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int N = p.N, H = p.H;
  int kvlen = N;
  if (p.kv_len != nullptr) kvlen = max(1, min(p.kv_len[b], N));
  const int nblk = (kvlen + kKB - 1) / kKB;
  const int row_in_img = qt * kQT + tid;
  const int col_q = h * kHD, col_k = (H + h) * kHD, col_v = (2 * H + h) * kHD;

  if (tid == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_out);
    mbar_init(bar_q, 1);
    mbar_init(&bar_kv[0], 1);
    mbar_init(&bar_kv[1], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  constexpr int kOCol = 128;        // S in columns [0,128), the block's O in [128, 128 + 64 NH)

  auto load_kv = [&](int blk) {
    const int buf = blk & 1;
    mbar_arrive_expect_tx(&bar_kv[buf], 2 * NH * kKBytes);
#pragma unroll
    for (int i = 0; i < NH; ++i) {
      tma_load_2d(smem + kOffK + (buf * NH + i) * kKBytes, &tm_kv, &bar_kv[buf], col_k + i * kHd, b * N + blk * kKB);
      tma_load_2d(smem + kOffV + (buf * NH + i) * kKBytes, &tm_kv, &bar_kv[buf], col_v + i * kHd, b * N + blk * kKB);
    }
  };
  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, NH * kQBytes);
#pragma unroll
    for (int i = 0; i < NH; ++i) tma_load_2d(sQ + i * kQBytes, &tm_q, bar_q, col_q + i * kHd, b * N + qt * kQT);
    load_kv(0);
    if (nblk > 1) load_kv(1);
  }

  constexpr uint32_t idesc_s = make_idesc_bf16(kQT, kKB, false, false);
  constexpr uint32_t idesc_o = make_idesc_bf16(kQT, kHd, false, true);
  float m_run = -INFINITY, l_run = 0.f;
  f32x2 acc[kHD / 2];
#pragma unroll
  for (int i = 0; i < kHD / 2; ++i) acc[i] = 0ull;
  const f32x2 sc2 = splat(p.scale_log2e);
  uint8_t* prow = sP + tid * 16;

  // One key block of the online softmax: row max, rescale factor, P = exp2(...) -> smem, row sum.
  auto softmax_block = [&](auto masked_tag, int nvalid, float& alpha) {
    constexpr bool MASKED = decltype(masked_tag)::value;
    float mb0 = -INFINITY, mb1 = -INFINITY;
    {
      uint32_t r[2][32];
      tmem_ld_x32(tmem_row, r[0]);
#pragma unroll
      for (int c = 0; c < kKB / 32; ++c) {
        tmem_ld_wait_x32(r[c & 1]);
        if (c + 1 < kKB / 32) tmem_ld_x32(tmem_row + (c + 1) * 32, r[(c + 1) & 1]);
        if (!MASKED || (c + 1) * 32 <= nvalid) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            mb0 = max3(mb0, __uint_as_float(r[c & 1][i]), __uint_as_float(r[c & 1][i + 1]));
            mb1 = max3(mb1, __uint_as_float(r[c & 1][i + 2]), __uint_as_float(r[c & 1][i + 3]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < nvalid) mb0 = fmaxf(mb0, __uint_as_float(r[c & 1][i]));
        }
      }
    }
    const float m_new = fmaxf(m_run, fmaxf(mb0, mb1));
    const float neg_mc = (BIAS && m_new == -INFINITY) ? 0.f : -m_new * p.scale_log2e;
    alpha = ex2_approx(fmaf(m_run, p.scale_log2e, neg_mc));
    const f32x2 nm2 = splat(neg_mc);
    f32x2 lsum2 = 0ull;
    {
      uint32_t r[2][32];
      tmem_ld_x32(tmem_row, r[0]);
#pragma unroll
      for (int c = 0; c < kKB / 32; ++c) {
        tmem_ld_wait_x32(r[c & 1]);
        if (c + 1 < kKB / 32) tmem_ld_x32(tmem_row + (c + 1) * 32, r[(c + 1) & 1]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const f32x2 tt = fma2(pk(__uint_as_float(r[c & 1][g * 8 + 2 * i]), __uint_as_float(r[c & 1][g * 8 + 2 * i + 1])), sc2, nm2);
            float e0, e1;
            if (!MASKED && ((kPolyMask >> (g * 4 + i)) & 1u)) {
              exp2_poly_pair_l(tt, e0, e1);
            } else {
              float t0, t1;
              upk(tt, t0, t1);
              e0 = ex2_approx(t0);
              e1 = ex2_approx(t1);
            }
            if (MASKED) {
              const int k = c * 32 + g * 8 + 2 * i;
              e0 = (k < nvalid) ? e0 : 0.f;
              e1 = (k + 1 < nvalid) ? e1 : 0.f;
            }
            lsum2 = add2(lsum2, pk(e0, e1));
            o[i] = pack_bf16(e0, e1);
          }
          *reinterpret_cast<uint4*>(prow + (c * 4 + g) * (kQT * 16)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    float s0, s1;
    upk(lsum2, s0, s1);
    l_run = fmaf(l_run, alpha, s0 + s1);
    m_run = m_new;
  };

  for (int j = 0; j < nblk; ++j) {
    const int buf = j & 1;
    const uint32_t ph = j & 1;
    if (tid == 0) {
      if (j == 0) mbar_wait(bar_q, 0);
      mbar_wait(&bar_kv[buf], (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int i = 0; i < NH; ++i) {
        const uint64_t qd = make_smem_desc(smem_u32(sQ + i * kQBytes), 16, 1024, kLayoutSw128);
        const uint64_t kd = make_smem_desc(smem_u32(smem + kOffK + (buf * NH + i) * kKBytes), 16, 1024, kLayoutSw128);
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k) umma_ss(tmem_base, qd + 2 * k, kd + 2 * k, idesc_s, (i > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_s, ph);
    tc_fence_after();

    if constexpr (BIAS) {
      const float* brow = row_in_img < N ? p.bias.ptr + b * p.bias.bs + h * p.bias.hs + static_cast<long long>(row_in_img) * p.bias.ld
                                         : nullptr;
      tmem_add_bias_row<kKB>(tmem_row, brow, j * kKB, p.bias.ld, p.inv_scale);
    }
    const int nvalid = kvlen - j * kKB;
    float alpha;
    if (nvalid >= kKB) softmax_block(std::false_type{}, nvalid, alpha);
    else softmax_block(std::true_type{}, nvalid, alpha);

    fence_proxy_async_smem();   // P (generic-proxy stores) -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      // P: K-major, no swizzle: core matrices 8 rows x 16 B, row groups 128 B apart (SBO), the two 8-key chunks of a
      //    K=16 step 2048 B apart (LBO).   V: MN-major SW128, 16 keys = 2048 B.
      const uint64_t pd = make_smem_desc(smem_u32(sP), kQT * 16, 128, kLayoutNone);
#pragma unroll
      for (int i = 0; i < NH; ++i) {      // one N = 64 MMA chain per 64-wide half of V
        const uint64_t vd = make_smem_desc(smem_u32(smem + kOffV + (buf * NH + i) * kKBytes), 1024, 1024, kLayoutSw128);
#pragma unroll
        for (int k = 0; k < kKB / 16; ++k)
          umma_ss(tmem_base + kOCol + i * kHd, pd + static_cast<uint64_t>(k * ((2 * kQT * 16) >> 4)),
                  vd + static_cast<uint64_t>(k * (2048 >> 4)), idesc_o, k > 0 ? 1u : 0u);
      }
      umma_commit(bar_o);
    }
    __syncwarp();
    mbar_wait(bar_o, ph);
    tc_fence_after();
    // the K / V buffer of block j is free: request block j+2 into it while block j+1 (already in flight) is consumed
    if (tid == 0 && j + 2 < nblk) load_kv(j + 2);
    const f32x2 al2 = splat(alpha);
#pragma unroll
    for (int c = 0; c < kHD; c += 32) {
      uint32_t r[32];
      tmem_ld_x32(tmem_row + kOCol + c, r);
      tmem_ld_wait_x32(r);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        acc[c / 2 + i] = fma2(acc[c / 2 + i], al2, pk(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])));
    }
    // no CTA barrier here: the next block's S MMA only needs every thread to be past its S reads (the barrier before
    // the PV MMA above), the next P writes wait for bar_o, and the next PV MMA is issued behind the next barrier,
    // i.e. after every thread finished reading this block's O
    tc_fence_before();
  }

  {
    // ---- epilogue: O / l -> bf16 -> swizzled staging (the dead Q tile) -> one TMA store per warp
    const f32x2 il2 = splat(1.0f / l_run);
    const uint32_t srow = smem_u32(sQ) + tid * 128;
    const uint32_t sw = static_cast<uint32_t>(tid & 7);
#pragma unroll
    for (int g = 0; g < 8 * NH; ++g) {     // half g / 8 of the head -> Q sub-tile g / 8
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float lo, hi;
        upk(mul2(acc[g * 4 + i], il2), lo, hi);
        o[i] = pack_bf16(lo, hi);
      }
      const uint32_t addr = srow + (g >> 3) * kQBytes + ((static_cast<uint32_t>(g & 7) ^ sw) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && qt * kQT + warp * 32 < N) {
#pragma unroll
      for (int i = 0; i < NH; ++i) tma_store_3d(&tm_out, sQ + i * kQBytes + warp * (32 * 128), col_q + i * kHd, qt * kQT + warp * 32, b);
      tma_store_commit();
      tma_store_wait_read<0>();   // the staging rows have been read; the global writes complete on their own
    }
    if (row_in_img < N && p.lse != nullptr)
      p.lse[(static_cast<long long>(b) * H + h) * N + row_in_img] = fmaf(m_run, p.scale, __logf(l_run));
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// hd = 64 or 128 (the padded head width the qkv buffer carries)
int launch_attn_fwd_long(const void* qkv, void* out, float* lse, const int32_t* kv_len, const AttnBias* bias, int B, int N, int H,
                         int hd, float scale, cudaStream_t st) {
  CUtensorMap tq, tkv, tout;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(3) * H * hd, static_cast<uint64_t>(B) * N};
    uint64_t strides[1] = {static_cast<uint64_t>(3) * H * hd * 2};
    uint32_t box[2] = {kHd, kQT};
    int rc = encode_tmap_bf16(&tq, qkv, 2, dims, strides, box, true);
    if (rc) return rc;
    rc = encode_tmap_bf16(&tkv, qkv, 2, dims, strides, box, true);
    if (rc) return rc;
    uint64_t odims[3] = {static_cast<uint64_t>(H) * hd, static_cast<uint64_t>(N), static_cast<uint64_t>(B)};
    uint64_t ostr[2] = {static_cast<uint64_t>(H) * hd * 2, static_cast<uint64_t>(N) * H * hd * 2};
    uint32_t obox[3] = {kHd, 32, 1};
    rc = encode_tmap_bf16(&tout, out, 3, odims, ostr, obox, 128);
    if (rc) return rc;
  }
  AttnLongParams p;
  p.lse = lse; p.kv_len = kv_len; p.B = B; p.N = N; p.H = H;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  static const int poly = [] { const char* e = getenv("B200_ATTN_LONG_POLY"); return e ? atoi(e) : 0; }();
  p.bias = bias != nullptr ? *bias : AttnBias{nullptr, 0, 0, 0};
  p.inv_scale = 1.0f / scale;
  auto kern = poly == 6 ? attn_fwd_long_kernel<0x2525u> : (poly == 4 ? attn_fwd_long_kernel<0x1111u> : attn_fwd_long_kernel<0u>);
  if (p.bias.ptr != nullptr) kern = attn_fwd_long_kernel<0u, true>;
  int smem_bytes = FwdLongSmem<1>::kSmem;
  if (hd == 2 * kHd) {
    kern = p.bias.ptr != nullptr ? attn_fwd_long_kernel<0u, true, 2> : attn_fwd_long_kernel<0u, false, 2>;
    smem_bytes = FwdLongSmem<2>::kSmem;
  }
  static decltype(kern) attr_set_for = nullptr;
  if (attr_set_for != kern) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set_for = kern;
  }
  dim3 grid((N + kQT - 1) / kQT, H, B);
  kern<<<grid, 128, smem_bytes, st>>>(tq, tkv, tout, p);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200
