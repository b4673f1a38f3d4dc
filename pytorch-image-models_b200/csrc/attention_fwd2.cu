// Attention forward v2 for N <= 208 tokens (ViT @224/16: 197), head_dim 64 -- persistent, warp-specialised, P in TMEM.
//
// One CTA per SM loops over (image, head) items; the two 128-row query tiles of an item are two INDEPENDENT pipelines
// that only share the item's Q/K/V stage, so the tensor-core phases of one overlap the exp phase of the other:
//   warp 8 (one thread)   TMA producer: keeps both smem stages filled (item i+1 lands while item i computes)
//   warp 9 / 10 (1 thr)   tcgen05 issuer of query tile 0 / 1: S_t = Q_t K^T (SS MMA, 128 x 208 x 16 x4), then
//                         O_t = P_t V (TS MMA: A = P read from TENSOR MEMORY, 128 x 64 x 16 x13, V MN-major in place)
//   warps 0-3 / 4-7       softmax warpgroup of query tile 0 / 1: one thread per query row, two pipelined passes over
//                         the S columns in TMEM (max, then exp2 + row sum); P is written back INTO the S columns as
//                         packed bf16 pairs (tcgen05.st), so no shared memory and no smem round trip for P;
//                         epilogue O / l -> bf16 -> 128B-swizzled staging tile -> TMA store into the merged-head
//                         [B, N, H*64] layout (a 3-D map clips the rows past N) (+ optional log-sum-exp)
// Shared memory: the two TMA stages (2 x 84 KB) + 2 x 16 KB output staging; TMEM: two 256-column regions (S/P/O).
// Replaces timm/layers/attention.py:115-135 exactly like attention_fwd.cu (which remains the path for N > 208).
#include <cstdlib>
#include "common.h"
#include "ptx.cuh"

namespace b200 {

namespace {
constexpr int kHd = 64;
constexpr int kKB = 208;                       // keys per item (MMA N of S, K of PV)
constexpr int kQBytes = 256 * kHd * 2;         // 32768: both query tiles in one TMA box
constexpr int kKBytes = kKB * kHd * 2;         // 26624
constexpr int kStageBytes = kQBytes + 2 * kKBytes;   // 86016 (multiple of 1024)
constexpr int kOutOff = 2 * kStageBytes;        // 2 query tiles x 128 rows x 128 B output staging (SW128)
constexpr int kBarOff = kOutOff + 2 * 128 * 128;
constexpr int kSmemTotal = kBarOff + 128 + 1024;
constexpr int kThreads = 11 * 32;
constexpr int kOCol = 128;                     // O accumulator inside a tile's 256-column TMEM region
constexpr int kFull = kKB / 32;                // 6 chunks of 32 columns + one of 16
}  // namespace

// 2^t for a pair of non-positive exponents WITHOUT the MUFU: t = j + f (j = round(t), |f| <= 0.5),
// 2^f by a degree-3 minimax polynomial (max relative error 7.6e-5, far inside the bf16 rounding of P), 2^j by adding
// j to the exponent field (the magic-constant add leaves j in the low mantissa bits of r).  The exp pass is bound by
// the 16-lane MUFU; running part of the columns through the FMA pipe instead shortens it (cf. FlashAttention-4).
__device__ __forceinline__ void exp2_poly_pair(f32x2 tt, float& e0, float& e1) {
  float t0, t1;
  upk(tt, t0, t1);
  tt = pk(fmaxf(t0, -126.0f), fmaxf(t1, -126.0f));
  const f32x2 r = add2(tt, splat(12582912.0f));
  const f32x2 f = add2(tt, add2(splat(12582912.0f), mul2(r, splat(-1.0f))));   // t - (r - magic)
  f32x2 q = fma2(f, splat(0.05520550534129143f), splat(0.24261397123336792f));
  q = fma2(q, f, splat(0.6932547688484192f));
  q = fma2(q, f, splat(0.9999276995658875f));
  float q0, q1, r0, r1;
  upk(q, q0, q1);
  upk(r, r0, r1);
  e0 = __int_as_float(__float_as_int(q0) + (__float_as_int(r0) << 23));
  e1 = __int_as_float(__float_as_int(q1) + (__float_as_int(r1) << 23));
}
// kPolyMask (template parameter): which of the 16 column pairs of a 32-column chunk take the polynomial path (the
// rest use MUFU.EX2)

struct AttnFwd2Params {
  __nv_bfloat16* out;
  float* lse;
  const int* kv_len;
  int B, N, H, num_items;
  float scale, scale_log2e;
  int stagger;        // cycles the second query tile's pipeline starts behind the first (anti-phase exp passes)
  long long* trace;   // debug (b200_set_trace): 16 cycle counters per CTA, nullptr in production
  AttnBias bias;      // BIAS kernels: additive mask / bias on the scaled scores (common.h)
  float inv_scale;
};

// MASKED = false: no key-padding lengths and N > 192, so the six full 32-column chunks carry no masking code at all
// (with a run-time condition the compiler if-converts the mask into a compare + select per element of every chunk).
// BIAS = true: an additive fp32 bias (causal triangle / general attn_mask, attention.py:21-40,129-131) is added to the
// score tile in TMEM before the two softmax passes read it (tmem_add_bias_row); everything after that is unchanged.
template <bool TRACE, uint32_t kPolyMask, bool MASKED, bool BIAS = false>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                 const __grid_constant__ CUtensorMap tm_out, const AttnFwd2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* bar_full = bars;            // [2] TMA landed (per stage)
  uint64_t* bar_sfree = bars + 2;       // [2] every MMA (of both query tiles) that read the stage has completed
  uint64_t* bar_s = bars + 4;           // [2] S_t in TMEM (per query tile)
  uint64_t* bar_p = bars + 6;           // [2] P_t written to TMEM by the 128 threads of the tile
  uint64_t* bar_o = bars + 8;           // [2] O_t in TMEM
  uint64_t* bar_tfree = bars + 10;      // [2] the tile's 128 threads have read O: its TMEM region is reusable
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, H = p.H;

  if (tid == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_out);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_sfree[i], 2);
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_p[i], 128);
      mbar_init(&bar_o[i], 1);
      mbar_init(&bar_tfree[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  griddep_launch();     // programmatic dependent launch (ptx.cuh): set-up above overlaps the QKV GEMM's tail
  griddep_wait();

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        const int stage = it & 1;
        if (it >= 2) mbar_wait(&bar_sfree[stage], ((it - 2) >> 1) & 1);   // both tiles' MMAs of item it-2 are done
        const int b = item / H, h = item - b * H;
        uint8_t* st = smem + stage * kStageBytes;
        mbar_arrive_expect_tx(&bar_full[stage], kStageBytes);
        tma_load_2d(st, &tm_q, &bar_full[stage], h * kHd, b * N);
        tma_load_2d(st + kQBytes, &tm_kv, &bar_full[stage], (H + h) * kHd, b * N);
        tma_load_2d(st + kQBytes + kKBytes, &tm_kv, &bar_full[stage], (2 * H + h) * kHd, b * N);
      }
    }
  } else if (warp >= 9) {
    // ------------------------------------------------------------------ MMA issuer of query tile t
    if (lane == 0) {
      const int t = warp - 9;
      constexpr uint32_t idesc_s = make_idesc_bf16(128, kKB, false, false);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, kHd, false, true);
      const uint32_t tmem_t = tmem_base + t * 256;
      long long w_full = 0, w_tfree = 0, w_p = 0;
      const long long c_begin = TRACE ? clock64() : 0;
      if (t == 1 && p.stagger > 0 && static_cast<int>(blockIdx.x) < p.num_items) {
        // start half a period late (counted from the arrival of the first operands): both tiles' chains are equally
        // long, so the offset persists and one tile's exp pass (MUFU) runs under the other's MMA / max / epilogue
        // phases instead of competing with its exp pass
        mbar_wait(&bar_full[0], 0);
        const long long c = clock64();
        while (clock64() - c < p.stagger) { }
      }
      int it = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        const int stage = it & 1;
        const uint32_t sbase = smem_u32(smem + stage * kStageBytes);
        long long c0 = TRACE ? clock64() : 0;
        mbar_wait(&bar_full[stage], (it >> 1) & 1);
        long long c1 = TRACE ? clock64() : 0;
        w_full += c1 - c0;
        if (it > 0) mbar_wait(&bar_tfree[t], (it - 1) & 1);   // the tile's previous O has been read out of TMEM
        if (TRACE) w_tfree += clock64() - c1;
        tc_fence_after();
        const uint64_t kd = make_smem_desc(sbase + kQBytes, 16, 1024, kLayoutSw128);
        const uint64_t qd = make_smem_desc(sbase + t * (128 * kHd * 2), 16, 1024, kLayoutSw128);
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k) umma_ss(tmem_t, qd + 2 * k, kd + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&bar_s[t]);
        const uint64_t vd = make_smem_desc(sbase + kQBytes + kKBytes, 1024, 1024, kLayoutSw128);
        if (TRACE) c0 = clock64();
        mbar_wait(&bar_p[t], it & 1);
        if (TRACE) w_p += clock64() - c0;
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kKB / 16; ++k)
          umma_ts(tmem_t + kOCol, tmem_t + k * 8, vd + static_cast<uint64_t>(k * (2048 >> 4)), idesc_o, k > 0 ? 1u : 0u);
        umma_commit(&bar_o[t]);
        umma_commit(&bar_sfree[stage]);
      }
      if (TRACE && p.trace) {
        long long* tr = p.trace + blockIdx.x * 16;
        tr[5 + 3 * t] = w_full; tr[6 + 3 * t] = w_tfree; tr[7 + 3 * t] = w_p;
        if (t == 0) { tr[14] = clock64() - c_begin; tr[15] = it; }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int t = warp >> 2;                                  // query tile
    const int rowt = (warp & 3) * 32 + lane;                  // row inside the tile == TMEM lane
    const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + t * 256;
    const int row_in_img = t * 128 + rowt;
    uint8_t* ostg = smem + kOutOff + t * (128 * 128) + (warp & 3) * (32 * 128);   // this warp's 32 rows x 128 B
    const uint32_t ostg_row = smem_u32(ostg) + lane * 128;
    const uint32_t osw = static_cast<uint32_t>(lane & 7);
    const f32x2 sc2 = splat(p.scale_log2e);
    int it = 0;
    long long ph[5] = {0, 0, 0, 0, 0};
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      const int b = item / H, h = item - b * H;
      int kvlen = N;
      if (p.kv_len != nullptr) kvlen = max(1, min(p.kv_len[b], N));
      long long ck = TRACE ? clock64() : 0;
      mbar_wait(&bar_s[t], it & 1);
      tc_fence_after();
      if (TRACE) { const long long c = clock64(); ph[0] += c - ck; ck = c; }
      if constexpr (BIAS) {
        const float* brow = row_in_img < N ? p.bias.ptr + b * p.bias.bs + h * p.bias.hs + static_cast<long long>(row_in_img) * p.bias.ld
                                           : nullptr;
        tmem_add_bias_row<kFull * 32 + 16>(tmem_row, brow, 0, p.bias.ld, p.inv_scale);
      }
      // ---- pass 1: row max over the valid keys
      float mb0 = -INFINITY, mb1 = -INFINITY;
      {
        uint32_t r[2][32];
        tmem_ld_x32(tmem_row, r[0]);
#pragma unroll
        for (int c = 0; c < kFull; ++c) {
          tmem_ld_wait_x32(r[c & 1]);
          if (c + 1 < kFull) tmem_ld_x32(tmem_row + (c + 1) * 32, r[(c + 1) & 1]);
          else tmem_ld_x16(tmem_row + kFull * 32, reinterpret_cast<uint32_t(&)[16]>(r[(c + 1) & 1]));
          if (!MASKED || (c + 1) * 32 <= kvlen) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              mb0 = max3(mb0, __uint_as_float(r[c & 1][i]), __uint_as_float(r[c & 1][i + 1]));
              mb1 = max3(mb1, __uint_as_float(r[c & 1][i + 2]), __uint_as_float(r[c & 1][i + 3]));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < kvlen) mb0 = fmaxf(mb0, __uint_as_float(r[c & 1][i]));
          }
        }
        uint32_t(&tl)[16] = reinterpret_cast<uint32_t(&)[16]>(r[kFull & 1]);
        tmem_ld_wait_x16(tl);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (kFull * 32 + i < kvlen) mb0 = fmaxf(mb0, __uint_as_float(tl[i]));
      }
      const float m = fmaxf(mb0, mb1);
      if (TRACE) { const long long c = clock64(); ph[1] += c - ck; ck = c; }
      const f32x2 nm2 = splat(-m * p.scale_log2e);
      f32x2 lsum2 = 0ull;
      // ---- pass 2: p = exp2((s - m) * scale * log2e); P chunk c (32 S columns) -> TMEM columns [16c, 16c + 16)
      {
        uint32_t r[2][32];
        tmem_ld_x32(tmem_row, r[0]);
#pragma unroll
        for (int c = 0; c < kFull; ++c) {
          tmem_ld_wait_x32(r[c & 1]);
          if (c + 1 < kFull) tmem_ld_x32(tmem_row + (c + 1) * 32, r[(c + 1) & 1]);
          else tmem_ld_x16(tmem_row + kFull * 32, reinterpret_cast<uint32_t(&)[16]>(r[(c + 1) & 1]));
          const bool full = !MASKED || (c + 1) * 32 <= kvlen;
          uint32_t pt[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const f32x2 tt = fma2(pk(__uint_as_float(r[c & 1][2 * i]), __uint_as_float(r[c & 1][2 * i + 1])), sc2, nm2);
            float e0, e1;
            if ((kPolyMask >> i) & 1u) {
              exp2_poly_pair(tt, e0, e1);
            } else {
              float t0, t1;
              upk(tt, t0, t1);
              e0 = ex2_approx(t0);
              e1 = ex2_approx(t1);
            }
            if (!full) {
              e0 = (c * 32 + 2 * i < kvlen) ? e0 : 0.f;
              e1 = (c * 32 + 2 * i + 1 < kvlen) ? e1 : 0.f;
            }
            lsum2 = add2(lsum2, pk(e0, e1));
            pt[i] = pack_bf16(e0, e1);
          }
          tmem_st_x16(tmem_row + c * 16, pt);
        }
        uint32_t(&tl)[16] = reinterpret_cast<uint32_t(&)[16]>(r[kFull & 1]);
        tmem_ld_wait_x16(tl);
        uint32_t pt8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float t0, t1;
          upk(fma2(pk(__uint_as_float(tl[2 * i]), __uint_as_float(tl[2 * i + 1])), sc2, nm2), t0, t1);
          const float e0 = (kFull * 32 + 2 * i < kvlen) ? ex2_approx(t0) : 0.f;
          const float e1 = (kFull * 32 + 2 * i + 1 < kvlen) ? ex2_approx(t1) : 0.f;
          lsum2 = add2(lsum2, pk(e0, e1));
          pt8[i] = pack_bf16(e0, e1);
        }
        tmem_st_x8(tmem_row + kFull * 16, pt8);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bar_p[t]);
      if (TRACE) { const long long c = clock64(); ph[2] += c - ck; ck = c; }
      float l0, l1;
      upk(lsum2, l0, l1);
      const float l = l0 + l1;
      // ---- epilogue: O / l -> bf16 rows of the merged-head output
      mbar_wait(&bar_o[t], it & 1);
      tc_fence_after();
      if (TRACE) { const long long c = clock64(); ph[3] += c - ck; ck = c; }
      const f32x2 il2 = splat(1.0f / l);
      uint32_t ro[2][32];
      tmem_ld_x32(tmem_row + kOCol, ro[0]);
      tmem_ld_x32(tmem_row + kOCol + 32, ro[1]);
      tmem_ld_wait_x32(ro[0]);
      tmem_ld_wait_x32(ro[1]);
      tc_fence_before();
      mbar_arrive(&bar_tfree[t]);                   // O is in registers: the issuer may start the next item's S
      if (lane == 0) tma_store_wait_read<0>();      // the previous item's store has left the staging rows
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float lo, hi;
            upk(mul2(pk(__uint_as_float(ro[c][g * 8 + 2 * i]), __uint_as_float(ro[c][g * 8 + 2 * i + 1])), il2), lo, hi);
            o[i] = pack_bf16(lo, hi);
          }
          const uint32_t addr = ostg_row + (((static_cast<uint32_t>(c * 4 + g)) ^ osw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3])
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && t * 128 + (warp & 3) * 32 < N) {
        tma_store_3d(&tm_out, ostg, h * kHd, t * 128 + (warp & 3) * 32, b);   // rows past N are clipped by the map
        tma_store_commit();
      }
      if (row_in_img < N && p.lse != nullptr)
        p.lse[(static_cast<long long>(b) * H + h) * N + row_in_img] = fmaf(m, p.scale, __logf(l));
      if (TRACE) { const long long c = clock64(); ph[4] += c - ck; ck = c; }
    }
    if (lane == 0) tma_store_wait<0>();
    if (TRACE && p.trace && (warp & 3) == 0 && lane == 0 && t == 0) {
      long long* tr = p.trace + blockIdx.x * 16;
      for (int i = 0; i < 5; ++i) tr[i] = ph[i];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

long long* g_attn_fwd_trace = nullptr;

int launch_attn_fwd2(const void* qkv, void* out, float* lse, const int32_t* kv_len, const AttnBias* bias, int B, int N, int H,
                     float scale, cudaStream_t st) {
  CUtensorMap tq, tkv;
  uint64_t dims[2] = {static_cast<uint64_t>(3) * H * kHd, static_cast<uint64_t>(B) * N};
  uint64_t strides[1] = {static_cast<uint64_t>(3) * H * kHd * 2};
  uint32_t box_q[2] = {kHd, 256};
  uint32_t box_kv[2] = {kHd, kKB};
  int rc = encode_tmap_bf16(&tq, qkv, 2, dims, strides, box_q, true);
  if (rc) return rc;
  rc = encode_tmap_bf16(&tkv, qkv, 2, dims, strides, box_kv, true);
  if (rc) return rc;
  CUtensorMap tout;
  {
    uint64_t odims[3] = {static_cast<uint64_t>(H) * kHd, static_cast<uint64_t>(N), static_cast<uint64_t>(B)};
    uint64_t ostr[2] = {static_cast<uint64_t>(H) * kHd * 2, static_cast<uint64_t>(N) * H * kHd * 2};
    uint32_t obox[3] = {kHd, 32, 1};
    rc = encode_tmap_bf16(&tout, out, 3, odims, ostr, obox, 128);
    if (rc) return rc;
  }
  AttnFwd2Params p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse; p.kv_len = kv_len; p.B = B; p.N = N; p.H = H; p.num_items = B * H;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.trace = g_attn_fwd_trace;
  p.bias = bias != nullptr ? *bias : AttnBias{nullptr, 0, 0, 0};
  p.inv_scale = 1.0f / scale;
  static const int stagger = [] { const char* e = getenv("B200_ATTN_STAGGER"); return e ? atoi(e) : 0; }();
  p.stagger = stagger;
  // B200_ATTN_POLY = how many of every 16 column pairs compute exp2 on the FMA pipe (0, 4, 6, 8, 10, 12)
  static const int poly = [] { const char* e = getenv("B200_ATTN_POLY"); return e ? atoi(e) : 6; }();
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnFwd2Params);
  const bool masked = kv_len != nullptr || N <= 192;
  Kern kern;
  if (p.bias.ptr != nullptr) {
    kern = masked ? attn_fwd2_kernel<false, 0u, true, true> : attn_fwd2_kernel<false, 0u, false, true>;
  } else if (masked) {
    kern = p.trace ? attn_fwd2_kernel<true, 0u, true> : attn_fwd2_kernel<false, 0u, true>;
  } else if (p.trace) {
    kern = poly == 0 ? attn_fwd2_kernel<true, 0u, false> : attn_fwd2_kernel<true, 0x5555u, false>;
  } else {
    kern = attn_fwd2_kernel<false, 0u, false>;
    if (poly == 4) kern = attn_fwd2_kernel<false, 0x1111u, false>;
    else if (poly == 6) kern = attn_fwd2_kernel<false, 0x2525u, false>;
    else if (poly == 8) kern = attn_fwd2_kernel<false, 0x5555u, false>;
    else if (poly == 10) kern = attn_fwd2_kernel<false, 0x5757u, false>;
    else if (poly == 12) kern = attn_fwd2_kernel<false, 0x7777u, false>;
  }
  static Kern attr_set_for = nullptr;
  if (attr_set_for != kern) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    attr_set_for = kern;
  }
  int grid = sm_count();
  if (grid > p.num_items) grid = p.num_items;
  B200_CUDA_OK(launch_pdl(kern, dim3(grid, 1, 1), dim3(kThreads, 1, 1), kSmemTotal, st, 1, tq, tkv, tout, p));
  return B200_OK;
}

}  // namespace b200
