// Attention backward v2 for sm_100a (head_dim 64, N <= 208 tokens) -- the kernel b200_attn_bwd launches.
// Same mathematics and transposed-domain formulation as attention_bwd.cu (kept as the v1 reference path), but
// organised so the tensor core and the 256 softmax threads overlap:
//   * queries are processed in two halves of 112 (MMA N = 112) so that S^T and dP^T each get their OWN TMEM region
//     (2 x 112 columns) next to the five persistent accumulators dV, dK (current key tile), dQ (both halves):
//     112 + 112(+pad) + 64*4 = 512 columns exactly;
//   * S^T and dP^T of a (key tile, query half) step are issued together, one pass over both produces P^T and dS^T
//     (bf16 core matrices in smem), split by columns across two warpgroups (8 warps).  The pass is bound by the TMEM read
//     bandwidth (~64 B/clk/SM: 115 KB of fp32 S^T / dP^T per step ~ 1800 of its ~2450 cycles, tools/attn_bwd_trace.py) --
//     a 16-warp variant (4 column groups) was measured and is NOT faster (319 vs 285 us per launch);
//   * a dedicated issuer warp queues [S^T, dP^T] of step i+1 FIRST and [dV += P^T dO, dK += dS^T Q, dQ += dS K] of step i
//     behind them: the compute warps start step i+1 (TMEM reads of S^T / dP^T at ~64 B/clk are the longest part of a
//     step) while the tensor core is still busy with step i's accumulations; a "smem free" barrier committed behind
//     those accumulations tells the compute warps when P^T / dS^T of step i may be overwritten.
// No atomics: dK / dV of a key tile accumulate over the query halves, dQ of a query half over the key tiles, all in TMEM.
// Persistent: one CTA per SM loops over (image, head) items; as soon as an item's last MMA has completed the issuer
// requests the NEXT item's Q / K / V / dO, so those loads run under the current item's dV / dK / dQ read-out and stores.
#include "common.h"
#include "ptx.cuh"

namespace b200 {

namespace {
constexpr int kHd = 64;
constexpr int kNQ = 208;                         // tokens loaded per operand tile (TMA box rows)
constexpr int kQH = 112;                         // queries per half (MMA N)
constexpr int kRowB = kHd * 2;                   // 128 B per token row
constexpr int kTileB = kNQ * kRowB;              // 26624
constexpr int kOffQ = 0, kOffK = kTileB, kOffV = 2 * kTileB, kOffDO = 3 * kTileB;
constexpr int kOffPad = 4 * kTileB;              // 2 KB of zeros: the q-half-1 over-read of dO (rows 208..223)
constexpr int kOffP = kOffPad + 2048;            // P^T : 14 chunks of [128 keys x 8 queries] bf16
constexpr int kPB = (kQH / 8) * 2048;            // 28672
constexpr int kOffDS = kOffP + kPB;              // dS^T: 16 chunks (the dQ MMA reads a full 128-query M tile)
constexpr int kDSB = 16 * 2048;
constexpr int kOffLse = kOffDS + kDSB;
constexpr int kOffDelta = kOffLse + 2 * 2 * kQH * 4;     // lse / delta: two buffers each (current item | next item)
constexpr int kOffOut = (kOffDelta + 2 * 2 * kQH * 4 + 1023) / 1024 * 1024;   // 8 warps x (32 rows x 128 B) output staging   // 8 warps x (32 rows x 128 B) output staging
constexpr int kOffBar = kOffOut + 8 * 4096;
constexpr int kSmemTotal = kOffBar + 64 + 1024;
constexpr int kColS = 0, kColDP = 128, kColDV = 256, kColDK = 320, kColDQ0 = 384, kColDQ1 = 448;
constexpr int kComputeThreads = 256;
}  // namespace

struct AttnBwd2Params {
  const float* lse;
  const float* delta;
  float* dbias;          // optional fp32 [3 * H * 64]: += column sums of dqkv (the qkv bias gradient), from the staging tiles
  __nv_bfloat16* dqkv;
  long long* trace;      // debug (b200_set_trace(2, buf)): 16 cycle counters per CTA, nullptr in production
  const int* kv_len;
  int B, N, H;
  float scale, scale_log2e;
  AttnBias bias;         // BIAS kernel: additive mask / bias on the scaled scores (common.h)
  float inv_scale;
};

// One warp's 32 accumulator rows x 64 columns: TMEM -> bf16 -> 128B-swizzled staging tile -> one TMA store into
// dqkv[b, row0 .. row0+31 (clipped at N by the 3-D map), col0 .. col0+63].  (Per-lane 16-byte global stores of 32
// different rows cost one LSU transaction each -- ~2000 cycles per 128 x 64 tile; the TMA store is one instruction.)
// dbias != nullptr: the column sums of the staged bf16 tile (rows < nrows; rows past the sequence are exact zeros: P = 0
// for padding queries / keys) are added to dbias[col0 ..] -- lane l owns columns 2l, 2l+1, reads them bank-conflict free
// through the 128-byte swizzle, one vector reduction per lane and tile.
__device__ __forceinline__ void store_tile64(uint8_t* stg, uint32_t taddr, const CUtensorMap* map, int col0, int row0, int b,
                                             int lane, bool any_row, float* dbias = nullptr, int nrows = 32) {
  if (lane == 0) tma_store_wait_read<0>();
  __syncwarp();
  const uint32_t srow = smem_u32(stg) + lane * 128;
  const uint32_t sw = static_cast<uint32_t>(lane & 7);
#pragma unroll
  for (int c = 0; c < kHd; c += 32) {
    uint32_t r[32];
    tmem_ld_x32(taddr + c, r);
    tmem_ld_wait_x32(r);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint32_t addr = srow + ((static_cast<uint32_t>(c / 8 + g) ^ sw) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                   "r"(pack_bf16(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1]))),
                   "r"(pack_bf16(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3]))),
                   "r"(pack_bf16(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5]))),
                   "r"(pack_bf16(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7])))
                   : "memory");
    }
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (dbias != nullptr && any_row) {
    const uint32_t base = smem_u32(stg) + static_cast<uint32_t>(lane & 3) * 4u;
    const uint32_t chunk = static_cast<uint32_t>(lane >> 2);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      uint32_t w;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(base + r * 128u + ((chunk ^ static_cast<uint32_t>(r & 7)) << 4)));
      if (r < nrows) {
        s0 += bf16_lo(w);
        s1 += bf16_hi(w);
      }
    }
    asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dbias + col0 + 2 * lane), "f"(s0), "f"(s1) : "memory");
  }
  if (lane == 0 && any_row) {
    tma_store_3d(map, stg, col0, row0, b);
    tma_store_commit();
  }
}

// one group of 8 queries: p = exp2(s*c - lse2), ds = p * (dp - delta) * scale  -> 16 B of P^T and dS^T each
__device__ __forceinline__ void pds8(const uint32_t* s, const uint32_t* dp, const float* lse2, const float* delta,
                                     f32x2 sc2, f32x2 scale2, bool key_ok, uint8_t* pdst, uint8_t* dsdst) {
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

// BIAS = true: P is recomputed from scores + additive bias (causal triangle / general attn_mask), added to the S^T registers
// right after they leave TMEM (add_bias_transposed); the bias itself receives no gradient (a mask, attention.py:21-40).
template <bool BIAS>
__global__ void __launch_bounds__(kComputeThreads + 32, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                 const __grid_constant__ CUtensorMap tm_out32, const __grid_constant__ CUtensorMap tm_out16,
                 const AttnBwd2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  uint8_t* sDO = smem + kOffDO;
  uint8_t* sP = smem + kOffP;
  uint8_t* sDS = smem + kOffDS;
  float* sLse = reinterpret_cast<float*>(smem + kOffLse);
  float* sDelta = reinterpret_cast<float*>(smem + kOffDelta);
  uint64_t* bar_ld = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* bar_sd = bar_ld + 1;      // issuer -> compute: S^T / dP^T of the step are in TMEM
  uint64_t* bar_p = bar_ld + 2;       // compute -> issuer: P^T / dS^T of the step are in smem (count 256)
  uint64_t* bar_fin = bar_ld + 3;     // issuer -> compute: every MMA of the item has completed
  uint64_t* bar_tail = bar_ld + 4;    // compute -> issuer: the item's dV / dK / dQ accumulators have been read (count 256)
  uint64_t* bar_free = bar_ld + 5;    // issuer -> compute: the step's dV / dK / dQ MMAs are done (P^T / dS^T smem reusable)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_ld + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, H = p.H;
  const int num_items = p.B * H;
  auto item_kvlen = [&](int item) {
    return p.kv_len != nullptr ? max(1, min(p.kv_len[item / H], N)) : N;
  };
  auto load_stats = [&](int item, int buf, int t0, int nthreads) {   // lse (log2 domain) / delta rows of an item -> smem
    for (int q = t0; q < 2 * kQH; q += nthreads) {
      const long long idx = static_cast<long long>(item) * N + q;      // item = b * H + h
      sLse[buf * 2 * kQH + q] = (q < N) ? p.lse[idx] * 1.4426950408889634f : INFINITY;   // +inf -> P = 0 for padding queries
      sDelta[buf * 2 * kQH + q] = (q < N) ? p.delta[idx] : 0.f;
    }
  };

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_out32);
    tma_prefetch_desc(&tm_out16);
    mbar_init(bar_ld, 1);
    mbar_init(bar_sd, 1);
    mbar_init(bar_p, kComputeThreads);
    mbar_init(bar_fin, 1);
    mbar_init(bar_tail, kComputeThreads);
    mbar_init(bar_free, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  if (static_cast<int>(blockIdx.x) < num_items) load_stats(blockIdx.x, 0, tid, blockDim.x);
  for (int i = tid; i < 2048 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem + kOffPad)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == kComputeThreads / 32) {
    // ------------------------------------------------------------------ issuer (TMA + tcgen05.mma), one thread
    if (lane == 0) {
      constexpr uint32_t idesc_sd = make_idesc_bf16(128, kQH, false, false);
      constexpr uint32_t idesc_kv = make_idesc_bf16(128, kHd, false, true);
      constexpr uint32_t idesc_dq = make_idesc_bf16(128, kHd, true, true);
      auto load_item = [&](int item) {
        const int b = item / H, h = item - b * H;
        mbar_arrive_expect_tx(bar_ld, 4 * kTileB);
        tma_load_2d(sQ, &tm_qkv, bar_ld, h * kHd, b * N);
        tma_load_2d(sK, &tm_qkv, bar_ld, (H + h) * kHd, b * N);
        tma_load_2d(sV, &tm_qkv, bar_ld, (2 * H + h) * kHd, b * N);
        tma_load_2d(sDO, &tm_do, bar_ld, h * kHd, b * N);
      };
      auto issue_sd = [&](int step) {
        const int kt = step >> 1, qh = step & 1;
        const uint64_t kd = make_smem_desc(smem_u32(sK) + kt * 128 * kRowB, 16, 1024, kLayoutSw128);
        const uint64_t vd = make_smem_desc(smem_u32(sV) + kt * 128 * kRowB, 16, 1024, kLayoutSw128);
        const uint64_t qd = make_smem_desc(smem_u32(sQ) + qh * kQH * kRowB, 16, 1024, kLayoutSw128);
        const uint64_t dod = make_smem_desc(smem_u32(sDO) + qh * kQH * kRowB, 16, 1024, kLayoutSw128);
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k) umma_ss(tmem_base + kColS, kd + 2 * k, qd + 2 * k, idesc_sd, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k) umma_ss(tmem_base + kColDP, vd + 2 * k, dod + 2 * k, idesc_sd, k > 0 ? 1u : 0u);
      };
      if (static_cast<int>(blockIdx.x) < num_items) load_item(blockIdx.x);
      int it = 0;
      uint32_t gs = 0;                       // global step counter: phase of bar_sd / bar_p
      const bool tr = p.trace != nullptr;
      long long w_ld = 0, w_p = 0, w_tail = 0, w_fin = 0, c0 = 0;
      const long long c_begin = tr ? clock64() : 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int num_steps = ((item_kvlen(item) + 127) / 128) * 2;
        if (tr) c0 = clock64();
        mbar_wait(bar_ld, it & 1);
        if (tr) w_ld += clock64() - c0;
        if (item + static_cast<int>(gridDim.x) < num_items) {      // the next item's operands start moving HBM -> L2 now
          const int nb = (item + gridDim.x) / H, nh = (item + gridDim.x) - nb * H;
          tma_prefetch_2d(&tm_qkv, nh * kHd, nb * N);
          tma_prefetch_2d(&tm_qkv, (H + nh) * kHd, nb * N);
          tma_prefetch_2d(&tm_qkv, (2 * H + nh) * kHd, nb * N);
          tma_prefetch_2d(&tm_do, nh * kHd, nb * N);
        }
        tc_fence_after();
        issue_sd(0);
        umma_commit(bar_sd);
        for (int step = 0; step < num_steps; ++step, ++gs) {
          const int kt = step >> 1, qh = step & 1;
          if (tr) c0 = clock64();
          mbar_wait(bar_p, gs & 1);
          if (tr) { const long long c = clock64(); w_p += c - c0; c0 = c; }
          if (step == 0 && it > 0) mbar_wait(bar_tail, (it - 1) & 1);   // previous item's accumulators were read out
          if (tr) w_tail += clock64() - c0;
          tc_fence_after();
          // the compute warps are done with S^T / dP^T of this step: the next step's go first, so that their softmax pass
          // (bound by the TMEM reads) runs under this step's accumulations
          if (step + 1 < num_steps) {
            issue_sd(step + 1);
            umma_commit(bar_sd);
          }
          // dV_kt (+)= P^T dO_qh ; dK_kt (+)= dS^T Q_qh   (A: K-major core matrices, 8-query chunks 2048 B apart)
          const uint64_t pd = make_smem_desc(smem_u32(sP), 2048, 128, kLayoutNone);
          const uint64_t dsd = make_smem_desc(smem_u32(sDS), 2048, 128, kLayoutNone);
          const uint64_t dob = make_smem_desc(smem_u32(sDO) + qh * kQH * kRowB, 1024, 1024, kLayoutSw128);
          const uint64_t qb = make_smem_desc(smem_u32(sQ) + qh * kQH * kRowB, 1024, 1024, kLayoutSw128);
#pragma unroll
          for (int k = 0; k < kQH / 16; ++k)
            umma_ss(tmem_base + kColDV, pd + static_cast<uint64_t>(k * (4096 >> 4)), dob + static_cast<uint64_t>(k * (2048 >> 4)),
                    idesc_kv, (qh > 0 || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < kQH / 16; ++k)
            umma_ss(tmem_base + kColDK, dsd + static_cast<uint64_t>(k * (4096 >> 4)), qb + static_cast<uint64_t>(k * (2048 >> 4)),
                    idesc_kv, (qh > 0 || k > 0) ? 1u : 0u);
          // dQ_qh (+)= dS K_kt   (A: the same dS^T tile read MN-major; B: K_kt MN-major in place)
          const uint64_t ad = make_smem_desc(smem_u32(sDS), 128, 2048, kLayoutNone);
          const uint64_t kb = make_smem_desc(smem_u32(sK) + kt * 128 * kRowB, 1024, 1024, kLayoutSw128);
#pragma unroll
          for (int k = 0; k < 128 / 16; ++k)
            umma_ss(tmem_base + (qh == 0 ? kColDQ0 : kColDQ1), ad + static_cast<uint64_t>(k * (256 >> 4)),
                    kb + static_cast<uint64_t>(k * (2048 >> 4)), idesc_dq, (kt > 0 || k > 0) ? 1u : 0u);
          umma_commit(bar_free);              // P^T / dS^T of this step have been consumed
          if (step + 1 == num_steps) umma_commit(bar_fin);
        }
        // every MMA of this item has completed -> the operand tiles are free: fetch the next item under the stores
        const int next = item + gridDim.x;
        if (next < num_items) {
          if (tr) c0 = clock64();
          mbar_wait(bar_fin, it & 1);
          if (tr) w_fin += clock64() - c0;
          load_item(next);
        }
      }
      if (tr) {
        long long* t = p.trace + blockIdx.x * 16;
        t[8] = w_ld; t[9] = w_p; t[10] = w_tail; t[11] = w_fin; t[12] = clock64() - c_begin; t[15] = it;
      }
    }
  } else {
    // ------------------------------------------------------------------ compute warps: thread = key row, warpgroup = column half
    const int wg = warp >> 2;                       // 0 / 1: which 56 of the 112 query columns
    const int rowt = (warp & 3) * 32 + lane;        // key row inside the tile == TMEM lane
    const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const f32x2 sc2 = splat(p.scale_log2e), scale2 = splat(p.scale);
    const long long tok_stride = 3LL * H * kHd;
    uint8_t* ostg = smem + kOffOut + warp * 4096;
    const int wrow = (warp & 3) * 32;                 // first accumulator row of this warp
    const int c0 = wg * 56;                           // first query column (within the half) of this warpgroup
    int it = 0;
    uint32_t gs = 0;
    const bool tr = p.trace != nullptr && tid == 0;
    long long ph[6] = {0, 0, 0, 0, 0, 0}, ck = tr ? clock64() : 0;   // wait S/dP | prev-tile store | softmax pass | wait fin | tail stores | stats
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
    const int b = item / H, h = item - b * H;
    const int kvlen = item_kvlen(item);
    const int num_kt = (kvlen + 127) / 128;
    const int num_steps = num_kt * 2;
    const int sb = it & 1;                          // which lse / delta buffer holds this item's statistics
    // the next item's statistics: requested now (one entry per thread), parked in registers, stored at the item's end
    const int nitem = item + static_cast<int>(gridDim.x);
    float n_lse = INFINITY, n_delta = 0.f;
    if (nitem < num_items && tid < 2 * kQH && tid < N) {
      n_lse = __ldg(p.lse + static_cast<long long>(nitem) * N + tid);
      n_delta = __ldg(p.delta + static_cast<long long>(nitem) * N + tid);
    }
    for (int step = 0; step < num_steps; ++step, ++gs) {
      const int kt = step >> 1, qh = step & 1;
      const int key = kt * 128 + rowt;
      const bool key_ok = key < kvlen;
      mbar_wait(bar_sd, gs & 1);
      tc_fence_after();
      if (tr) { const long long c = clock64(); ph[0] += c - ck; ck = c; }
      bool freed = false;               // waited for the previous step's accumulations (P^T / dS^T smem free, dV / dK final)
      if (qh == 0 && kt > 0) {
        // dV / dK of the previous key tile: their MMAs are queued BEHIND this step's S^T / dP^T -> wait for them
        mbar_wait(bar_free, (gs - 1) & 1);
        tc_fence_after();
        freed = true;
        const int prow0 = (kt - 1) * 128 + wrow;
        store_tile64(ostg, tmem_row + (wg == 0 ? kColDV : kColDK), &tm_out32, ((wg == 0 ? 2 : 1) * H + h) * kHd, prow0, b,
                     lane, prow0 < N, p.dbias);
      }
      if (tr) { const long long c = clock64(); ph[1] += c - ck; ck = c; }
      const float* l2 = sLse + sb * 2 * kQH + qh * kQH + c0;
      const float* dl = sDelta + sb * 2 * kQH + qh * kQH + c0;
      uint8_t* prow = sP + (c0 >> 3) * 2048 + rowt * 16;
      uint8_t* dsrow = sDS + (c0 >> 3) * 2048 + rowt * 16;
      // Dead work is skipped instead of computed: the warp's 32 key rows all lie past kv_len (second key tile of N = 197:
      // keys 224..255), or a chunk's queries all lie past N (second query half: 200..223) -> P^T = dS^T = 0 without
      // touching TMEM.  The pass is bound by the TMEM read bandwidth, so this is ~22 % of it at N = 197.
      const bool rows_dead = kt * 128 + (warp & 3) * 32 >= kvlen;
      const int q_a = qh * kQH + c0;
      const float* bkey = nullptr;       // bias[q_a, key] of this item
      if constexpr (BIAS)
        if (key < N) bkey = p.bias.ptr + b * p.bias.bs + h * p.bias.hs + static_cast<long long>(q_a) * p.bias.ld + key;
      auto zero_groups = [&](int g0, int ng) {
        if (!freed && gs > 0) { mbar_wait(bar_free, (gs - 1) & 1); freed = true; }
        for (int g = g0; g < g0 + ng; ++g) {
          *reinterpret_cast<uint4*>(prow + g * 2048) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(dsrow + g * 2048) = make_uint4(0u, 0u, 0u, 0u);
        }
      };
      if (rows_dead || q_a >= N) {
        zero_groups(0, 4);
      } else {
        uint32_t s[32], d[32];
        tmem_ld_x32(tmem_row + kColS + c0, s);
        tmem_ld_x32(tmem_row + kColDP + c0, d);
        tmem_ld_wait_x32(s);
        tmem_ld_wait_x32(d);
        // first write into P^T / dS^T: the previous step's dV / dK / dQ MMAs must have finished reading them (their
        // completion normally lies well inside the TMEM reads above)
        if (!freed && gs > 0) { mbar_wait(bar_free, (gs - 1) & 1); freed = true; }
        if constexpr (BIAS) add_bias_transposed<32>(s, bkey, p.bias.ld, N - q_a, p.inv_scale);
#pragma unroll
        for (int g = 0; g < 4; ++g)
          pds8(s + g * 8, d + g * 8, l2 + g * 8, dl + g * 8, sc2, scale2, key_ok, prow + g * 2048, dsrow + g * 2048);
      }
      if (rows_dead || q_a + 32 >= N) {
        zero_groups(4, 2);
      } else {
        uint32_t s[16], d[16];
        tmem_ld_x16(tmem_row + kColS + c0 + 32, s);
        tmem_ld_x16(tmem_row + kColDP + c0 + 32, d);
        tmem_ld_wait_x16(s);
        tmem_ld_wait_x16(d);
        if constexpr (BIAS)
          add_bias_transposed<16>(s, bkey ? bkey + 32LL * p.bias.ld : nullptr, p.bias.ld, N - q_a - 32, p.inv_scale);
#pragma unroll
        for (int g = 0; g < 2; ++g)
          pds8(s + g * 8, d + g * 8, l2 + 32 + g * 8, dl + 32 + g * 8, sc2, scale2, key_ok, prow + (4 + g) * 2048,
               dsrow + (4 + g) * 2048);
      }
      if (rows_dead || q_a + 48 >= N) {
        zero_groups(6, 1);
      } else {
        uint32_t s[8], d[8];
        tmem_ld_x8(tmem_row + kColS + c0 + 48, s);
        tmem_ld_x8(tmem_row + kColDP + c0 + 48, d);
        tmem_ld_wait_x8(s);
        tmem_ld_wait_x8(d);
        if constexpr (BIAS)
          add_bias_transposed<8>(s, bkey ? bkey + 48LL * p.bias.ld : nullptr, p.bias.ld, N - q_a - 48, p.inv_scale);
        pds8(s, d, l2 + 48, dl + 48, sc2, scale2, key_ok, prow + 6 * 2048, dsrow + 6 * 2048);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);
      if (tr) { const long long c = clock64(); ph[2] += c - ck; ck = c; }
    }
    // ---- tail: last key tile's dV / dK, then dQ
    mbar_wait(bar_fin, it & 1);
    tc_fence_after();
    if (tr) { const long long c = clock64(); ph[3] += c - ck; ck = c; }
    {
      const int prow0 = (num_kt - 1) * 128 + wrow;
      store_tile64(ostg, tmem_row + (wg == 0 ? kColDV : kColDK), &tm_out32, ((wg == 0 ? 2 : 1) * H + h) * kHd, prow0, b, lane,
                   prow0 < N, p.dbias);
    }
    {
      // warpgroup 0 -> query half 0, warpgroup 1 -> half 1; a half has 112 rows, so the last warp stores 16 rows only
      const int q0 = wg * kQH + wrow;
      store_tile64(ostg, tmem_row + (wg == 0 ? kColDQ0 : kColDQ1), wrow + 32 <= kQH ? &tm_out32 : &tm_out16, h * kHd, q0, b,
                   lane, q0 < N, p.dbias, wrow + 32 <= kQH ? 32 : 16);
    }
    tc_fence_before();
    mbar_arrive(bar_tail);            // accumulators read: the issuer may start the next item's dV / dK / dQ MMAs
    if (tr) { const long long c = clock64(); ph[4] += c - ck; ck = c; }
    // keys beyond kv_len (padding keys of a NaFlex batch) never entered a tile: their dK / dV are zero
    for (int key = num_kt * 128 + tid; key < N; key += kComputeThreads) {
      for (int which = 1; which <= 2; ++which) {
        uint4* dst = reinterpret_cast<uint4*>(p.dqkv + (static_cast<long long>(b) * N + key) * tok_stride + (which * H + h) * kHd);
#pragma unroll
        for (int g = 0; g < 8; ++g) dst[g] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    // the next item's softmax statistics (every compute thread is past its last use of this item's)
    if (nitem < num_items) {
      if (tid < 2 * kQH) {
        sLse[(sb ^ 1) * 2 * kQH + tid] = n_lse * 1.4426950408889634f;     // +inf stays +inf (padding queries -> P = 0)
        sDelta[(sb ^ 1) * 2 * kQH + tid] = n_delta;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");     // visible to every compute thread before the next item reads them
    }
    if (tr) { const long long c = clock64(); ph[5] += c - ck; ck = c; }
    }   // items
    if (tr) {
      long long* t = p.trace + blockIdx.x * 16;
      for (int i = 0; i < 6; ++i) t[i] = ph[i];
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

long long* g_attn_bwd_trace = nullptr;

int launch_attn_bwd2(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv,
                     const int32_t* kv_len, const AttnBias* bias, int B, int N, int H, float scale, float* dbias, cudaStream_t st) {
  CUtensorMap tq, tdo;
  uint32_t box[2] = {kHd, kNQ};
  {
    uint64_t dims[2] = {static_cast<uint64_t>(3) * H * kHd, static_cast<uint64_t>(B) * N};
    uint64_t strides[1] = {static_cast<uint64_t>(3) * H * kHd * 2};
    int rc = encode_tmap_bf16(&tq, qkv, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(H) * kHd, static_cast<uint64_t>(B) * N};
    uint64_t strides[1] = {static_cast<uint64_t>(H) * kHd * 2};
    int rc = encode_tmap_bf16(&tdo, dout, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  CUtensorMap to32, to16;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(3) * H * kHd, static_cast<uint64_t>(N), static_cast<uint64_t>(B)};
    uint64_t strides[2] = {static_cast<uint64_t>(3) * H * kHd * 2, static_cast<uint64_t>(N) * 3 * H * kHd * 2};
    uint32_t b32[3] = {kHd, 32, 1}, b16[3] = {kHd, 16, 1};
    int rc = encode_tmap_bf16(&to32, dqkv, 3, dims, strides, b32, 128);
    if (rc) return rc;
    rc = encode_tmap_bf16(&to16, dqkv, 3, dims, strides, b16, 128);
    if (rc) return rc;
  }
  AttnBwd2Params p;
  p.lse = lse; p.delta = delta; p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv); p.kv_len = kv_len;
  p.dbias = dbias;
  p.trace = g_attn_bwd_trace;
  p.B = B; p.N = N; p.H = H; p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.bias = bias != nullptr ? *bias : AttnBias{nullptr, 0, 0, 0};
  p.inv_scale = 1.0f / scale;
  auto kern = p.bias.ptr != nullptr ? attn_bwd2_kernel<true> : attn_bwd2_kernel<false>;
  static decltype(kern) attr_set_for = nullptr;
  if (attr_set_for != kern) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    attr_set_for = kern;
  }
  int grid = sm_count();
  if (grid > B * H) grid = B * H;
  kern<<<grid, kComputeThreads + 32, kSmemTotal, st>>>(tq, tdo, to32, to16, p);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200
