// Attention forward v2 for N <= 208 tokens (ViT @224/16: 197), head_dim 64 -- persistent, warp-specialised, P in TMEM.
//
// One CTA per SM loops over (image, head) items.  Per item both 128-row query tiles are processed concurrently:
//   warp 8 (one thread)   TMA producer + tcgen05 issuer: prefetches the NEXT item's Q/K/V into the other smem stage
//                         while the current item computes; issues S_t = Q_t K^T (SS MMA, 128 x 208 x 16 x4) and
//                         O_t = P_t V (TS MMA: A = P read from TENSOR MEMORY, 128 x 64 x 16 x13, V MN-major in place)
//   warps 0-3 / 4-7       softmax warpgroup of query tile 0 / 1: one thread per query row, two pipelined passes over
//                         the S columns in TMEM (max, then exp2 + row sum); P is written back INTO the S columns as
//                         packed bf16 pairs (tcgen05.st), so no shared memory and no smem round trip for P;
//                         epilogue O / l -> bf16 -> merged-head [B, N, H*64] rows (+ optional log-sum-exp)
// Shared memory is only the two TMA stages (2 x 85 KB); TMEM: two 256-column regions (S/P/O per query tile).
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
constexpr int kBarOff = 2 * kStageBytes;
constexpr int kSmemTotal = kBarOff + 128 + 1024;
constexpr int kOCol = 128;                     // O accumulator inside a tile's 256-column TMEM region
constexpr int kFull = kKB / 32;                // 6 chunks of 32 columns + one of 16
}  // namespace

struct AttnFwd2Params {
  __nv_bfloat16* out;
  float* lse;
  const int* kv_len;
  int B, N, H, num_items;
  float scale, scale_log2e;
};

__global__ void __launch_bounds__(288, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                 const AttnFwd2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* bar_full = bars;            // [2] TMA landed (per stage)
  uint64_t* bar_sfree = bars + 2;       // [2] every MMA that read the stage has completed
  uint64_t* bar_s = bars + 4;           // [2] S_t in TMEM (per query tile)
  uint64_t* bar_p = bars + 6;           // [2] P_t written to TMEM by the 128 threads of the tile
  uint64_t* bar_o = bars + 8;           // [2] O_t in TMEM
  uint64_t* bar_tfree = bars + 10;      // all 256 compute threads have read O: TMEM regions reusable
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, H = p.H;

  if (tid == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_sfree[i], 1);
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_p[i], 128);
      mbar_init(&bar_o[i], 1);
    }
    mbar_init(bar_tfree, 256);
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
    // ------------------------------------------------------------------ producer + MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, kKB, false, false);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, kHd, false, true);
      auto load_item = [&](int item, int stage) {
        const int b = item / H, h = item - b * H;
        uint8_t* st = smem + stage * kStageBytes;
        mbar_arrive_expect_tx(&bar_full[stage], kStageBytes);
        tma_load_2d(st, &tm_q, &bar_full[stage], h * kHd, b * N);
        tma_load_2d(st + kQBytes, &tm_kv, &bar_full[stage], (H + h) * kHd, b * N);
        tma_load_2d(st + kQBytes + kKBytes, &tm_kv, &bar_full[stage], (2 * H + h) * kHd, b * N);
      };
      int it = 0;
      if (static_cast<int>(blockIdx.x) < p.num_items) load_item(blockIdx.x, 0);
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        const int stage = it & 1;
        const uint32_t sbase = smem_u32(smem + stage * kStageBytes);
        mbar_wait(&bar_full[stage], (it >> 1) & 1);
        if (it > 0) mbar_wait(bar_tfree, (it - 1) & 1);     // previous item's O has been read out of TMEM
        tc_fence_after();
        const uint64_t kd = make_smem_desc(sbase + kQBytes, 16, 1024, kLayoutSw128);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint64_t qd = make_smem_desc(sbase + t * (128 * kHd * 2), 16, 1024, kLayoutSw128);
#pragma unroll
          for (int k = 0; k < kHd / 16; ++k) umma_ss(tmem_base + t * 256, qd + 2 * k, kd + 2 * k, idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&bar_s[t]);
        }
        // prefetch the next item into the other stage (free once the previous item's PV MMAs have completed)
        const int next = item + gridDim.x;
        if (next < p.num_items) {
          if (it > 0) mbar_wait(&bar_sfree[stage ^ 1], ((it - 1) >> 1) & 1);
          load_item(next, stage ^ 1);
        }
        const uint64_t vd = make_smem_desc(sbase + kQBytes + kKBytes, 1024, 1024, kLayoutSw128);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&bar_p[t], it & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kKB / 16; ++k)
            umma_ts(tmem_base + t * 256 + kOCol, tmem_base + t * 256 + k * 8, vd + static_cast<uint64_t>(k * (2048 >> 4)),
                    idesc_o, k > 0 ? 1u : 0u);
          umma_commit(&bar_o[t]);
        }
        umma_commit(&bar_sfree[stage]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int t = warp >> 2;                                  // query tile
    const int rowt = (warp & 3) * 32 + lane;                  // row inside the tile == TMEM lane
    const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + t * 256;
    const int row_in_img = t * 128 + rowt;
    const f32x2 sc2 = splat(p.scale_log2e);
    int it = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      const int b = item / H, h = item - b * H;
      int kvlen = N;
      if (p.kv_len != nullptr) kvlen = max(1, min(p.kv_len[b], N));
      mbar_wait(&bar_s[t], it & 1);
      tc_fence_after();
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
          if ((c + 1) * 32 <= kvlen) {
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
          const bool full = (c + 1) * 32 <= kvlen;
          uint32_t pt[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float t0, t1;
            upk(fma2(pk(__uint_as_float(r[c & 1][2 * i]), __uint_as_float(r[c & 1][2 * i + 1])), sc2, nm2), t0, t1);
            float e0 = ex2_approx(t0), e1 = ex2_approx(t1);
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
      float l0, l1;
      upk(lsum2, l0, l1);
      const float l = l0 + l1;
      // ---- epilogue: O / l -> bf16 rows of the merged-head output
      mbar_wait(&bar_o[t], it & 1);
      tc_fence_after();
      const f32x2 il2 = splat(1.0f / l);
      __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * N + row_in_img) * (H * kHd) + h * kHd;
#pragma unroll
      for (int c = 0; c < kHd; c += 32) {
        uint32_t r[32];
        tmem_ld_x32(tmem_row + kOCol + c, r);
        tmem_ld_wait_x32(r);
        if (row_in_img < N) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float lo, hi;
              upk(mul2(pk(__uint_as_float(r[g * 8 + 2 * i]), __uint_as_float(r[g * 8 + 2 * i + 1])), il2), lo, hi);
              o[i] = pack_bf16(lo, hi);
            }
            reinterpret_cast<uint4*>(orow + c)[g] = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tfree);
      if (row_in_img < N && p.lse != nullptr)
        p.lse[(static_cast<long long>(b) * H + h) * N + row_in_img] = fmaf(m, p.scale, __logf(l));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_attn_fwd2(const void* qkv, void* out, float* lse, const int32_t* kv_len, int B, int N, int H, float scale,
                     cudaStream_t st) {
  CUtensorMap tq, tkv;
  uint64_t dims[2] = {static_cast<uint64_t>(3) * H * kHd, static_cast<uint64_t>(B) * N};
  uint64_t strides[1] = {static_cast<uint64_t>(3) * H * kHd * 2};
  uint32_t box_q[2] = {kHd, 256};
  uint32_t box_kv[2] = {kHd, kKB};
  int rc = encode_tmap_bf16(&tq, qkv, 2, dims, strides, box_q, true);
  if (rc) return rc;
  rc = encode_tmap_bf16(&tkv, qkv, 2, dims, strides, box_kv, true);
  if (rc) return rc;
  AttnFwd2Params p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse; p.kv_len = kv_len; p.B = B; p.N = N; p.H = H; p.num_items = B * H;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(attn_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    attr_set = true;
  }
  int grid = sm_count();
  if (grid > p.num_items) grid = p.num_items;
  attn_fwd2_kernel<<<grid, 288, kSmemTotal, st>>>(tq, tkv, p);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200
