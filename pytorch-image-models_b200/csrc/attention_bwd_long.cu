// Attention backward for long sequences (208 < N <= 2240 tokens: ViT-L/14 @336 = 577, DINOv2 @518 = 1370, NaFlex batches), head_dim 64, and for
// head_dim 128 at any N <= 1120 (NH = 2, see the end of this comment).
// Same transposed-domain mathematics as attention_bwd2.cu, re-tiled so that nothing scales with N on chip:
//   * persistent CTAs loop over (image, head) items; per item an outer loop over 128-key tiles and an inner loop over
//     112-query chunks.  K / V of the key tile stay in shared memory, Q / dO chunks stream through a two-stage TMA ring;
//   * TMEM holds S^T and dP^T of the step (2 x 112 columns), the dV / dK accumulators of the key tile (accumulated over
//     the query chunks) and ONE 64-column dQ partial of the step: 448 columns for any N;
//   * the dQ partial of a (key tile, query chunk) step leaves as fp32 through a swizzled staging tile and a TMA
//     reduce-add into a zero-initialised fp32 dq buffer [B, N, H*64] (the sum over key tiles happens in L2);
//     attn_dq_convert_kernel writes it into the q third of dqkv afterwards.  dK / dV tiles leave as bf16 TMA stores.
// head_dim 128 (NH = 2; 72 / 80 / 96-wide heads arrive zero-padded through the weights): every operand tile is kept as two
// 64-wide sub-tiles (one 128-byte swizzle atom per token row each).  TMEM cannot hold both halves of dV / dK / dQ next to
// S^T / dP^T (608 columns), so a work item is (image, head, HALF): S^T and dP^T contract over the full 128-wide head (8
// k-steps over the two sub-tiles), the three accumulations run for the item's half only and the other half is a second item
// (S^T / dP^T are computed twice: 14 instead of 10 64-wide MMA units per head).  The Q / dO ring has one stage (shared memory).
// The reference has no backward code (gradients of timm/layers/attention.py:115-135 come from autograd's SDPA backward);
// the kernel is validated against autograd of the oracle, like the short-sequence kernels.
#include "common.h"
#include "ptx.cuh"

namespace b200 {

namespace {
constexpr int kHd = 64;
constexpr int kQH = 112;                        // queries per chunk (MMA N of S^T / dP^T)
constexpr int kKT = 128;                        // keys per tile
constexpr int kRowB = kHd * 2;                  // 128 B per token row
constexpr int kKVB = kKT * kRowB;               // 16384
constexpr int kQB = kQH * kRowB;                // 14336 (14 x 1024: SW128 atoms stay aligned)
constexpr int kPB = (kQH / 8) * 2048;           // 28672
constexpr int kDSB = 16 * 2048;
// shared-memory layout for NH 64-wide halves of the head (offsets in the comments: NH = 1)
template <int NH>
struct BwdLongSmem {
  static constexpr int kStages = NH == 1 ? 2 : 1;         // Q / dO ring depth (NH = 2: one stage is all that fits)
  static constexpr int kMaxChunks = NH == 1 ? 20 : 10;    // query chunks whose lse / delta fit: N <= 2240 (DINOv2 @518: 1370) / 1120
  static constexpr int kOffK = 0, kOffV = NH * kKVB;      // NH sub-tiles each
  static constexpr int kOffQ = 2 * NH * kKVB;             // stage s: Q chunk (NH sub-tiles) at kOffQ + s * kStage, dO chunk behind it
  static constexpr int kStage = 2 * NH * kQB;             // 28672
  static constexpr int kOffP = kOffQ + kStages * kStage;  // 90112: P^T, 14 chunks of [128 keys x 8 queries] bf16
  static constexpr int kOffDS = kOffP + kPB;              // 118784: dS^T, 16 chunks (the dQ MMA reads a full 128-query M tile)
  static constexpr int kOffOut = kOffDS + kDSB;           // 151552: 8 warps x (32 rows x 128 B) output staging
  static constexpr int kOffLse = kOffOut + 8 * 4096;      // 184320
  static constexpr int kOffDelta = kOffLse + kMaxChunks * kQH * 4;
  static constexpr int kOffBar = kOffDelta + kMaxChunks * kQH * 4;   // 202240
  static constexpr int kSmemTotal = kOffBar + 128;
};
static_assert(BwdLongSmem<2>::kSmemTotal <= 232448, "head_dim 128: shared memory");
constexpr int kColS = 0, kColDP = 128, kColDV = 256, kColDK = 320, kColDQ = 384;
constexpr int kComputeThreads = 256;
}  // namespace

struct AttnBwdLongParams {
  const float* lse;
  const float* delta;
  const int* kv_len;
  __nv_bfloat16* dqkv;
  int B, N, H;
  float scale, scale_log2e;
  AttnBias bias;         // BIAS kernel: additive mask / bias on the scaled scores (common.h)
  float inv_scale;
};

__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// one group of 8 queries: p = exp2(s*c - lse2), ds = p * (dp - delta) * scale  -> 16 B of P^T and dS^T each
__device__ __forceinline__ void pds8_l(const uint32_t* s, const uint32_t* dp, const float* lse2, const float* delta, f32x2 sc2,
                                       f32x2 scale2, bool key_ok, uint8_t* pdst, uint8_t* dsdst) {
  uint32_t po[4], dso[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t0, t1;
    upk(fma2(pk(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), sc2, pk(-lse2[2 * i], -lse2[2 * i + 1])), t0, t1);
    const float p0 = key_ok ? ex2_approx(t0) : 0.f, p1 = key_ok ? ex2_approx(t1) : 0.f;
    const f32x2 d = add2(pk(__uint_as_float(dp[2 * i]), __uint_as_float(dp[2 * i + 1])), pk(-delta[2 * i], -delta[2 * i + 1]));
    float d0, d1;
    upk(mul2(mul2(pk(p0, p1), d), scale2), d0, d1);
    po[i] = pack_bf16(p0, p1);
    dso[i] = key_ok ? pack_bf16(d0, d1) : 0u;
  }
  *reinterpret_cast<uint4*>(pdst) = make_uint4(po[0], po[1], po[2], po[3]);
  *reinterpret_cast<uint4*>(dsdst) = make_uint4(dso[0], dso[1], dso[2], dso[3]);
}

// BIAS = true: scores + additive bias (causal triangle / general attn_mask), as in attention_bwd2.cu.
template <bool BIAS, int NH = 1>
__global__ void __launch_bounds__(kComputeThreads + 32, 1)
attn_bwd_long_kernel(const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_q,
                     const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_out32,
                     const __grid_constant__ CUtensorMap tm_dq32, const __grid_constant__ CUtensorMap tm_dq16,
                     const AttnBwdLongParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using L = BwdLongSmem<NH>;
  constexpr int kStages = L::kStages, kOffK = L::kOffK, kOffV = L::kOffV, kOffQ = L::kOffQ, kStage = L::kStage, kOffP = L::kOffP,
                kOffDS = L::kOffDS, kOffOut = L::kOffOut, kOffLse = L::kOffLse, kOffDelta = L::kOffDelta, kOffBar = L::kOffBar;
  static_assert(L::kSmemTotal <= 232448, "shared memory");
  constexpr int kHD = NH * kHd;                            // the (padded) head width
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  uint8_t* sP = smem + kOffP;
  uint8_t* sDS = smem + kOffDS;
  float* sLse = reinterpret_cast<float*>(smem + kOffLse);
  float* sDelta = reinterpret_cast<float*>(smem + kOffDelta);
  uint64_t* bar_kv = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* bar_q = bar_kv + 1;        // [2] Q / dO chunk of the stage landed
  uint64_t* bar_qfree = bar_kv + 3;    // [2] the MMAs that read the stage have completed
  uint64_t* bar_sd = bar_kv + 5;       // S^T / dP^T of the step are in TMEM
  uint64_t* bar_p = bar_kv + 6;        // P^T / dS^T of the step are in smem (count 256)
  uint64_t* bar_dq = bar_kv + 7;       // dV / dK / dQ MMAs of the step have completed
  uint64_t* bar_dqfree = bar_kv + 8;   // the dQ partial has been read out of TMEM (count 256)
  uint64_t* bar_kvdone = bar_kv + 9;   // every MMA of the key tile has completed
  uint64_t* bar_tail = bar_kv + 10;    // dV / dK of the key tile have been read out of TMEM (count 256)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_kv + 11);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, H = p.H;
  const int num_items = p.B * H * NH;                      // item = ((image, head), half of the head)
  const int nq = (N + kQH - 1) / kQH;                      // query chunks per item
  auto item_kvlen = [&](int item) { return p.kv_len != nullptr ? max(1, min(p.kv_len[item / (H * NH)], N)) : N; };

  if (tid == 0) {
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_out32);
    tma_prefetch_desc(&tm_dq32);
    tma_prefetch_desc(&tm_dq16);
    mbar_init(bar_kv, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_q[i], 1);
      mbar_init(&bar_qfree[i], 1);
    }
    mbar_init(bar_sd, 1);
    mbar_init(bar_p, kComputeThreads);
    mbar_init(bar_dq, 1);
    mbar_init(bar_dqfree, kComputeThreads);
    mbar_init(bar_kvdone, 1);
    mbar_init(bar_tail, kComputeThreads);
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

  if (warp == 8) {
    // ------------------------------------------------------------------ issuer (TMA + tcgen05.mma), one thread
    if (lane == 0) {
      constexpr uint32_t idesc_sd = make_idesc_bf16(128, kQH, false, false);
      constexpr uint32_t idesc_kv = make_idesc_bf16(128, kHd, false, true);
      constexpr uint32_t idesc_dq = make_idesc_bf16(128, kHd, true, true);
      uint32_t gs = 0, gk = 0;                 // global step / key-tile counters (barrier phases)
      uint32_t loads[2] = {0, 0}, uses[2] = {0, 0};
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int bh = item / NH, hf = item - bh * NH;     // hf: which 64-wide half this item accumulates
        const int b = bh / H, h = bh - b * H;
        const int nkt = (item_kvlen(item) + kKT - 1) / kKT;
        const int total = nkt * nq;
        auto load_chunk = [&](int qh, int st) {
          if (loads[st] > 0) mbar_wait(&bar_qfree[st], (loads[st] - 1) & 1);   // the stage's previous readers are done
          uint8_t* dst = smem + kOffQ + st * kStage;
          mbar_arrive_expect_tx(&bar_q[st], 2 * NH * kQB);
#pragma unroll
          for (int i = 0; i < NH; ++i) {
            tma_load_2d(dst + i * kQB, &tm_q, &bar_q[st], h * kHD + i * kHd, b * N + qh * kQH);
            tma_load_2d(dst + (NH + i) * kQB, &tm_do, &bar_q[st], h * kHD + i * kHd, b * N + qh * kQH);
          }
          ++loads[st];
        };
        load_chunk(0, gs % kStages);
        for (int s = 0; s < total; ++s, ++gs) {
          const int kt = s / nq, qh = s - kt * nq;
          const int st = gs % kStages;
          if (qh == 0) {
            if (gk > 0) mbar_wait(bar_kvdone, (gk - 1) & 1);       // every MMA reading the previous K / V tile is done
            mbar_arrive_expect_tx(bar_kv, 2 * NH * kKVB);
#pragma unroll
            for (int i = 0; i < NH; ++i) {
              tma_load_2d(sK + i * kKVB, &tm_kv, bar_kv, (H + h) * kHD + i * kHd, b * N + kt * kKT);
              tma_load_2d(sV + i * kKVB, &tm_kv, bar_kv, (2 * H + h) * kHD + i * kHd, b * N + kt * kKT);
            }
            mbar_wait(bar_kv, gk & 1);
          }
          mbar_wait(&bar_q[st], uses[st] & 1);
          ++uses[st];
          tc_fence_after();
          const uint32_t qs = smem_u32(smem + kOffQ + st * kStage), dos = qs + NH * kQB;
          {
#pragma unroll
            for (int i = 0; i < NH; ++i) {      // S^T = K Q^T: the contraction runs over every 64-wide half of the head
              const uint64_t kd = make_smem_desc(smem_u32(sK + i * kKVB), 16, 1024, kLayoutSw128);
              const uint64_t qd = make_smem_desc(qs + i * kQB, 16, 1024, kLayoutSw128);
#pragma unroll
              for (int k = 0; k < kHd / 16; ++k)
                umma_ss(tmem_base + kColS, kd + 2 * k, qd + 2 * k, idesc_sd, (i > 0 || k > 0) ? 1u : 0u);
            }
#pragma unroll
            for (int i = 0; i < NH; ++i) {      // dP^T = V dO^T
              const uint64_t vd = make_smem_desc(smem_u32(sV + i * kKVB), 16, 1024, kLayoutSw128);
              const uint64_t dod = make_smem_desc(dos + i * kQB, 16, 1024, kLayoutSw128);
#pragma unroll
              for (int k = 0; k < kHd / 16; ++k)
                umma_ss(tmem_base + kColDP, vd + 2 * k, dod + 2 * k, idesc_sd, (i > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(bar_sd);
          }
          if (kStages == 2 && s + 1 < total) load_chunk((s + 1) % nq, (gs + 1) % kStages);   // next chunk streams in under this step
          mbar_wait(bar_p, gs & 1);
          if (gs > 0) mbar_wait(bar_dqfree, (gs - 1) & 1);          // the previous dQ partial has left TMEM
          if (qh == 0 && gk > 0) mbar_wait(bar_tail, (gk - 1) & 1);  // the previous key tile's dV / dK have left TMEM
          tc_fence_after();
          const uint64_t pd = make_smem_desc(smem_u32(sP), 2048, 128, kLayoutNone);
          const uint64_t dsd = make_smem_desc(smem_u32(sDS), 2048, 128, kLayoutNone);
          // the accumulations of this item's half hf: dV[:, hf] / dK[:, hf] / dQ[:, hf] read sub-tile hf of dO / Q / K
          const uint64_t dob = make_smem_desc(dos + hf * kQB, 1024, 1024, kLayoutSw128);
          const uint64_t qb = make_smem_desc(qs + hf * kQB, 1024, 1024, kLayoutSw128);
#pragma unroll
          for (int k = 0; k < kQH / 16; ++k)     // dV_kt (+)= P^T dO_qh
            umma_ss(tmem_base + kColDV, pd + static_cast<uint64_t>(k * (4096 >> 4)), dob + static_cast<uint64_t>(k * (2048 >> 4)),
                    idesc_kv, (qh > 0 || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < kQH / 16; ++k)     // dK_kt (+)= dS^T Q_qh
            umma_ss(tmem_base + kColDK, dsd + static_cast<uint64_t>(k * (4096 >> 4)), qb + static_cast<uint64_t>(k * (2048 >> 4)),
                    idesc_kv, (qh > 0 || k > 0) ? 1u : 0u);
          const uint64_t ad = make_smem_desc(smem_u32(sDS), 128, 2048, kLayoutNone);
          const uint64_t kb = make_smem_desc(smem_u32(sK + hf * kKVB), 1024, 1024, kLayoutSw128);
#pragma unroll
          for (int k = 0; k < kKT / 16; ++k)     // dQ partial = dS K_kt   (A: dS^T read MN-major; B: K_kt MN-major in place)
            umma_ss(tmem_base + kColDQ, ad + static_cast<uint64_t>(k * (256 >> 4)), kb + static_cast<uint64_t>(k * (2048 >> 4)),
                    idesc_dq, k > 0 ? 1u : 0u);
          umma_commit(bar_dq);
          umma_commit(&bar_qfree[st]);
          if (qh == nq - 1) {
            umma_commit(bar_kvdone);
            ++gk;
          }
          // one-stage ring: the next chunk can only be requested once this step's MMAs have released the stage
          if (kStages == 1 && s + 1 < total) load_chunk((s + 1) % nq, 0);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ compute warps: thread = key row, warpgroup = column half
    const int wg = warp >> 2;                       // 0 / 1: which 56 of the 112 query columns (S / dP), which 32 of the dQ columns
    const int rowt = (warp & 3) * 32 + lane;        // key row inside the tile == TMEM lane
    const int wrow = (warp & 3) * 32;
    const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(wrow) << 16);
    const f32x2 sc2 = splat(p.scale_log2e), scale2 = splat(p.scale);
    uint8_t* ostg = smem + kOffOut + warp * 4096;
    const uint32_t ostg_row = smem_u32(ostg) + lane * 128;
    const uint32_t osw = static_cast<uint32_t>(lane & 7);
    const long long tok_stride = 3LL * H * kHD;
    uint32_t gs = 0, gk = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int bh = item / NH, hf = item - bh * NH;
      const int b = bh / H, h = bh - b * H;
      const int kvlen = item_kvlen(item);
      const int nkt = (kvlen + kKT - 1) / kKT;
      // softmax statistics of the item's queries (lse in the log2 domain; +inf -> P = 0 for padding queries)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int q = tid; q < nq * kQH; q += kComputeThreads) {
        const long long idx = static_cast<long long>(bh) * N + q;
        sLse[q] = (q < N) ? p.lse[idx] * 1.4426950408889634f : INFINITY;
        sDelta[q] = (q < N) ? p.delta[idx] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int kt = 0; kt < nkt; ++kt) {
        const bool key_ok = kt * kKT + rowt < kvlen;
        for (int qh = 0; qh < nq; ++qh, ++gs) {
          mbar_wait(bar_sd, gs & 1);
          tc_fence_after();
          const int c0 = wg * 56;
          const float* l2 = sLse + qh * kQH + c0;
          const float* dl = sDelta + qh * kQH + c0;
          uint8_t* prow = sP + (c0 >> 3) * 2048 + rowt * 16;
          uint8_t* dsrow = sDS + (c0 >> 3) * 2048 + rowt * 16;
          const int q_a = qh * kQH + c0;
          const float* bkey = nullptr;       // bias[q_a, key] of this item
          if constexpr (BIAS)
            if (kt * kKT + rowt < N && q_a < N)
              bkey = p.bias.ptr + b * p.bias.bs + h * p.bias.hs + static_cast<long long>(q_a) * p.bias.ld + kt * kKT + rowt;
          {
            uint32_t s[32], d[32];
            tmem_ld_x32(tmem_row + kColS + c0, s);
            tmem_ld_x32(tmem_row + kColDP + c0, d);
            tmem_ld_wait_x32(s);
            tmem_ld_wait_x32(d);
            if constexpr (BIAS) add_bias_transposed<32>(s, bkey, p.bias.ld, N - q_a, p.inv_scale);
#pragma unroll
            for (int g = 0; g < 4; ++g)
              pds8_l(s + g * 8, d + g * 8, l2 + g * 8, dl + g * 8, sc2, scale2, key_ok, prow + g * 2048, dsrow + g * 2048);
          }
          {
            uint32_t s[16], d[16];
            tmem_ld_x16(tmem_row + kColS + c0 + 32, s);
            tmem_ld_x16(tmem_row + kColDP + c0 + 32, d);
            tmem_ld_wait_x16(s);
            tmem_ld_wait_x16(d);
            if constexpr (BIAS)
              add_bias_transposed<16>(s, bkey ? bkey + 32LL * p.bias.ld : nullptr, p.bias.ld, N - q_a - 32, p.inv_scale);
#pragma unroll
            for (int g = 0; g < 2; ++g)
              pds8_l(s + g * 8, d + g * 8, l2 + 32 + g * 8, dl + 32 + g * 8, sc2, scale2, key_ok, prow + (4 + g) * 2048,
                     dsrow + (4 + g) * 2048);
          }
          {
            uint32_t s[8], d[8];
            tmem_ld_x8(tmem_row + kColS + c0 + 48, s);
            tmem_ld_x8(tmem_row + kColDP + c0 + 48, d);
            tmem_ld_wait_x8(s);
            tmem_ld_wait_x8(d);
            if constexpr (BIAS)
              add_bias_transposed<8>(s, bkey ? bkey + 48LL * p.bias.ld : nullptr, p.bias.ld, N - q_a - 48, p.inv_scale);
            pds8_l(s, d, l2 + 48, dl + 48, sc2, scale2, key_ok, prow + 6 * 2048, dsrow + 6 * 2048);
          }
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(bar_p);
          // ---- dQ partial of this step: rows = queries of the chunk (TMEM lanes), this warp's 32 rows x 32 fp32 columns
          mbar_wait(bar_dq, gs & 1);
          tc_fence_after();
          {
            uint32_t r[32];
            tmem_ld_x32(tmem_row + kColDQ + wg * 32, r);
            tmem_ld_wait_x32(r);
            tc_fence_before();
            mbar_arrive(bar_dqfree);
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const uint32_t addr = ostg_row + ((static_cast<uint32_t>(g) ^ osw) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[4 * g]), "r"(r[4 * g + 1]), "r"(r[4 * g + 2]),
                           "r"(r[4 * g + 3])
                           : "memory");
            }
            fence_proxy_async_smem();
            __syncwarp();
            const int q0 = qh * kQH + wrow;                      // first query row of this warp's tile
            if (lane == 0 && wrow < kQH && q0 < N) {
              // the last quadrant of a 112-query chunk holds 16 rows only
              tma_reduce_add_3d(wrow + 32 <= kQH ? &tm_dq32 : &tm_dq16, ostg, h * kHD + hf * kHd + wg * 32, q0, b);
              tma_store_commit();
            }
          }
        }
        // ---- dV / dK of this key tile (every query chunk accumulated)
        mbar_wait(bar_kvdone, gk & 1);
        tc_fence_after();
        {
          const int prow0 = kt * kKT + wrow;
          const uint32_t taddr = tmem_row + (wg == 0 ? kColDV : kColDK);
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (int c = 0; c < kHd; c += 32) {
            uint32_t r[32];
            tmem_ld_x32(taddr + c, r);
            tmem_ld_wait_x32(r);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t addr = ostg_row + ((static_cast<uint32_t>(c / 8 + g) ^ osw) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                           "r"(pack_bf16(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1]))),
                           "r"(pack_bf16(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3]))),
                           "r"(pack_bf16(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5]))),
                           "r"(pack_bf16(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7])))
                           : "memory");
            }
          }
          tc_fence_before();
          mbar_arrive(bar_tail);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && prow0 < N) {
            tma_store_3d(&tm_out32, ostg, ((wg == 0 ? 2 : 1) * H + h) * kHD + hf * kHd, prow0, b);
            tma_store_commit();
          }
          ++gk;
        }
      }
      // keys beyond kv_len (padding keys of a NaFlex batch) never entered a tile: their dK / dV are zero
      for (int key = nkt * kKT + tid; key < N; key += kComputeThreads) {
        for (int which = 1; which <= 2; ++which) {
          uint4* dst = reinterpret_cast<uint4*>(p.dqkv + (static_cast<long long>(b) * N + key) * tok_stride + (which * H + h) * kHD +
                                                hf * kHd);
#pragma unroll
          for (int g = 0; g < 8; ++g) dst[g] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dqkv[b, n, 0 : H*64] = bf16(dq32[b, n, :])
__global__ void __launch_bounds__(256)
attn_dq_convert_kernel(const float* __restrict__ dq32, __nv_bfloat16* __restrict__ dqkv, long long rows, int HD) {
  const long long total = rows * (HD / 8);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / (HD / 8);
    const int c = static_cast<int>(i - r * (HD / 8)) * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(dq32 + r * HD + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(dq32 + r * HD + c + 4));
    *reinterpret_cast<uint4*>(dqkv + r * 3 * HD + c) =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
}

// dq32: fp32 workspace [B, N, H*hd] (zeroed here); hd = 64 or 128.  Returns B200_ERR_UNSUPPORTED for N > 2240 (1120 at hd 128).
int launch_attn_bwd_long(const void* qkv, const void* dout, const float* lse, const float* delta, float* dq32, void* dqkv,
                         const int32_t* kv_len, const AttnBias* bias, int B, int N, int H, int hd, float scale, cudaStream_t st) {
  const int NH = hd / kHd;
  const int max_n = (NH == 1 ? BwdLongSmem<1>::kMaxChunks : BwdLongSmem<2>::kMaxChunks) * kQH;
  if (N > max_n) return fail(B200_ERR_UNSUPPORTED, "b200_attn_bwd: N=%d tokens not supported (<= %d at head_dim %d)", N, max_n, hd);
  const long long HD = static_cast<long long>(H) * hd;
  B200_CUDA_OK(cudaMemsetAsync(dq32, 0, static_cast<size_t>(B) * N * HD * sizeof(float), st));
  CUtensorMap tkv, tq, tdo, to32, tdq32, tdq16;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(3) * HD, static_cast<uint64_t>(B) * N};
    uint64_t strides[1] = {static_cast<uint64_t>(3) * HD * 2};
    uint32_t bkv[2] = {kHd, kKT}, bq[2] = {kHd, kQH};
    int rc = encode_tmap_bf16(&tkv, qkv, 2, dims, strides, bkv, 128);
    if (rc) return rc;
    rc = encode_tmap_bf16(&tq, qkv, 2, dims, strides, bq, 128);
    if (rc) return rc;
    uint64_t ddims[2] = {static_cast<uint64_t>(HD), static_cast<uint64_t>(B) * N};
    uint64_t dstr[1] = {static_cast<uint64_t>(HD) * 2};
    rc = encode_tmap_bf16(&tdo, dout, 2, ddims, dstr, bq, 128);
    if (rc) return rc;
    uint64_t odims[3] = {static_cast<uint64_t>(3) * HD, static_cast<uint64_t>(N), static_cast<uint64_t>(B)};
    uint64_t ostr[2] = {static_cast<uint64_t>(3) * HD * 2, static_cast<uint64_t>(N) * 3 * HD * 2};
    uint32_t b32[3] = {kHd, 32, 1};
    rc = encode_tmap_bf16(&to32, dqkv, 3, odims, ostr, b32, 128);
    if (rc) return rc;
    uint64_t qdims[3] = {static_cast<uint64_t>(HD), static_cast<uint64_t>(N), static_cast<uint64_t>(B)};
    uint64_t qstr[2] = {static_cast<uint64_t>(HD) * 4, static_cast<uint64_t>(N) * HD * 4};
    uint32_t q32[3] = {32, 32, 1}, q16[3] = {32, 16, 1};
    rc = encode_tmap_f32(&tdq32, dq32, 3, qdims, qstr, q32, 128);
    if (rc) return rc;
    rc = encode_tmap_f32(&tdq16, dq32, 3, qdims, qstr, q16, 128);
    if (rc) return rc;
  }
  AttnBwdLongParams p;
  p.lse = lse; p.delta = delta; p.kv_len = kv_len; p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  p.B = B; p.N = N; p.H = H; p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.bias = bias != nullptr ? *bias : AttnBias{nullptr, 0, 0, 0};
  p.inv_scale = 1.0f / scale;
  auto kern = p.bias.ptr != nullptr ? attn_bwd_long_kernel<true> : attn_bwd_long_kernel<false>;
  int smem_bytes = BwdLongSmem<1>::kSmemTotal;
  if (NH == 2) {
    kern = p.bias.ptr != nullptr ? attn_bwd_long_kernel<true, 2> : attn_bwd_long_kernel<false, 2>;
    smem_bytes = BwdLongSmem<2>::kSmemTotal;
  }
  static decltype(kern) attr_set_for = nullptr;
  if (attr_set_for != kern) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set_for = kern;
  }
  int grid = sm_count();
  if (grid > B * H * NH) grid = B * H * NH;
  kern<<<grid, kComputeThreads + 32, smem_bytes, st>>>(tkv, tq, tdo, to32, tdq32, tdq16, p);
  B200_CUDA_OK(cudaGetLastError());
  const long long rows = static_cast<long long>(B) * N;
  long long blocks = (rows * (HD / 8) + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  attn_dq_convert_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(dq32, reinterpret_cast<__nv_bfloat16*>(dqkv), rows,
                                                                  static_cast<int>(HD));
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200
