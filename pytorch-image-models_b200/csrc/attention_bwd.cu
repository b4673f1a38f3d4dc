// Attention backward for sm_100a (head_dim 64, N <= 208 tokens): dqkv [B,N,3,H,64] from dout [B,N,H*64].
//
// The reference has no backward code: gradients of timm/layers/attention.py:115-135 come from PyTorch autograd
// (SDPA backward).  This kernel owns that definition; it is validated against autograd of the oracle.
//
// One CTA per (head, image) holds Q, K, V, dO of that head in shared memory (TMA, read straight out of the
// [B,N,3,H,64] qkv buffer / [B,N,H*64] dout buffer) and works in the TRANSPOSED domain, one thread per KEY row:
//     S^T  = K Q^T                       tcgen05.mma 128 x 208 x 16 (x4)        -> TMEM
//     P^T  = exp2(S^T * c - lse2[q])     thread-local, bf16 core matrices       -> smem
//     dP^T = V dO^T                      tcgen05.mma (same TMEM columns as S^T) -> TMEM
//     dS^T = P^T * (dP^T - delta[q]) * scale                                    -> smem
//     dV   = P^T dO,   dK = dS^T Q       tcgen05.mma 128 x 64 x 16 (x13), A from smem, B MN-major in place
//     dQ  += dS K                        the same dS^T tile read MN-major as A; accumulated in TMEM over key tiles
// so dK / dV / dQ of a head are complete inside the CTA: no atomics, no second pass.
// delta[q] = sum_d dO[q,d] * O[q,d] comes from attn_bwd_delta_kernel.
#include <cstdlib>
#include "common.h"
#include "ptx.cuh"

namespace b200 {

namespace {
constexpr int kHd = 64;
constexpr int kNQ = 208;                       // max tokens per image handled (queries = MMA N dimension)
constexpr int kRowB = kHd * 2;                 // 128 B per token row
constexpr int kTileB = kNQ * kRowB;            // 26624
constexpr int kPB = (kNQ / 8) * 2048;          // P^T: 26 chunks of [128 keys x 16 B]
constexpr int kDSB = 32 * 2048;                // dS^T: 32 chunks so the q >= 208 half-tile of dQ reads inside the buffer
constexpr int kOffQ = 0, kOffK = kTileB, kOffV = 2 * kTileB, kOffDO = 3 * kTileB;
constexpr int kOffP = 4 * kTileB;              // 106496 (1024-aligned)
constexpr int kOffDS = kOffP + kPB;            // 159744
constexpr int kOffLse = kOffDS + kDSB;         // 225280
constexpr int kOffDelta = kOffLse + kNQ * 4;
constexpr int kOffBar = kOffDelta + kNQ * 4;   // 226944
constexpr int kSmemTotal = kOffBar + 64 + 1024;
// TMEM columns
constexpr int kColSD = 0, kColDV = 208, kColDK = 272, kColDQ0 = 336, kColDQ1 = 400;
}  // namespace

struct AttnBwdParams {
  const float* lse;
  const float* delta;
  __nv_bfloat16* dqkv;
  const int* kv_len;
  int B, N, H;
  float scale, scale_log2e;
};

// delta[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d]   (fp32); one warp per (b, q) row of H*hd elements (LPH = hd / 8 lanes per head).
template <int LPH>
__global__ void attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                      float* __restrict__ delta, int B, int N, int H) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= static_cast<long long>(B) * N) return;
  const int b = static_cast<int>(row / N), q = static_cast<int>(row % N);
  constexpr int hd = LPH * 8;
  // a head's hd elements = LPH lanes x 8 elements; 32 / LPH heads per warp pass
  for (int h0 = 0; h0 < H; h0 += 32 / LPH) {
    const int h = h0 + lane / LPH;
    float s = 0.f;
    if (h < H) {
      const long long off = row * (H * hd) + h * hd + (lane % LPH) * 8;
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(o + off));
      const uint4 g = __ldg(reinterpret_cast<const uint4*>(dout + off));
      s = bf16_lo(a.x) * bf16_lo(g.x) + bf16_hi(a.x) * bf16_hi(g.x) + bf16_lo(a.y) * bf16_lo(g.y) +
          bf16_hi(a.y) * bf16_hi(g.y) + bf16_lo(a.z) * bf16_lo(g.z) + bf16_hi(a.z) * bf16_hi(g.z) +
          bf16_lo(a.w) * bf16_lo(g.w) + bf16_hi(a.w) * bf16_hi(g.w);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (LPH == 16) s += __shfl_xor_sync(0xffffffffu, s, 8);
    if (h < H && lane % LPH == 0) delta[(static_cast<long long>(b) * H + h) * N + q] = s;
  }
}

__global__ void __launch_bounds__(128, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                const AttnBwdParams p) {
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
  uint64_t* bar_s = bar_ld + 1;
  uint64_t* bar_dp = bar_ld + 2;
  uint64_t* bar_out = bar_ld + 3;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_ld + 4);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x, b = blockIdx.y;
  const int N = p.N, H = p.H;
  int kvlen = N;
  if (p.kv_len != nullptr) kvlen = max(1, min(p.kv_len[b], N));

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    mbar_init(bar_ld, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_dp, 1);
    mbar_init(bar_out, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  // per-query softmax statistics of this head: lse (log2 domain; +inf for padding queries -> P = 0) and delta
  for (int q = tid; q < kNQ; q += 128) {
    const long long idx = (static_cast<long long>(b) * H + h) * N + q;
    sLse[q] = (q < N) ? p.lse[idx] * 1.4426950408889634f : INFINITY;
    sDelta[q] = (q < N) ? p.delta[idx] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_ld, 4 * kTileB);
    tma_load_2d(sQ, &tm_qkv, bar_ld, h * kHd, b * N);
    tma_load_2d(sK, &tm_qkv, bar_ld, (H + h) * kHd, b * N);
    tma_load_2d(sV, &tm_qkv, bar_ld, (2 * H + h) * kHd, b * N);
    tma_load_2d(sDO, &tm_do, bar_ld, h * kHd, b * N);
  }

  constexpr uint32_t idesc_sd = make_idesc_bf16(128, kNQ, false, false);   // S^T, dP^T
  constexpr uint32_t idesc_kv = make_idesc_bf16(128, kHd, false, true);    // dV, dK  (A K-major, B MN-major)
  constexpr uint32_t idesc_dq = make_idesc_bf16(128, kHd, true, true);     // dQ      (A MN-major, B MN-major)
  const int num_kt = (kvlen + 127) / 128;
  const f32x2 sc2 = splat(p.scale_log2e);
  const long long tok_stride = 3LL * H * kHd;

  for (int kt = 0; kt < num_kt; ++kt) {
    const uint32_t ph = kt & 1;
    const int key = kt * 128 + tid;
    const bool key_ok = key < kvlen;
    const uint32_t sk_t = smem_u32(sK) + kt * 128 * kRowB;
    const uint32_t sv_t = smem_u32(sV) + kt * 128 * kRowB;

    if (tid == 0) {
      if (kt == 0) mbar_wait(bar_ld, 0);
      tc_fence_after();
      const uint64_t ad = make_smem_desc(sk_t, 16, 1024, kLayoutSw128);
      const uint64_t bd = make_smem_desc(smem_u32(sQ), 16, 1024, kLayoutSw128);
#pragma unroll
      for (int k = 0; k < kHd / 16; ++k) umma_ss(tmem_base + kColSD, ad + 2 * k, bd + 2 * k, idesc_sd, k > 0 ? 1u : 0u);
      umma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_s, ph);
    tc_fence_after();

    // ---- pass A: P^T[key, q] = exp2(S^T * c - lse2[q])  -> sP (bf16 core matrices, own row)
    uint8_t* prow = sP + tid * 16;
#pragma unroll 1
    for (int c = 0; c < kNQ; c += 32) {
      if (c + 32 <= kNQ) {
        uint32_t r[32];
        tmem_ld_x32(tmem_row + kColSD + c, r);
        tmem_ld_wait_x32(r);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int q = c + g * 8 + 2 * i;
            float t0, t1;
            upk(fma2(pk(__uint_as_float(r[g * 8 + 2 * i]), __uint_as_float(r[g * 8 + 2 * i + 1])), sc2,
                     pk(-sLse[q], -sLse[q + 1])), t0, t1);
            o[i] = key_ok ? pack_bf16(ex2_approx(t0), ex2_approx(t1)) : 0u;
          }
          *reinterpret_cast<uint4*>(prow + ((c >> 3) + g) * 2048) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      } else {
        uint32_t r[16];
        tmem_ld_x16(tmem_row + kColSD + c, r);
        tmem_ld_wait_x16(r);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int q = c + g * 8 + 2 * i;
            float t0, t1;
            upk(fma2(pk(__uint_as_float(r[g * 8 + 2 * i]), __uint_as_float(r[g * 8 + 2 * i + 1])), sc2,
                     pk(-sLse[q], -sLse[q + 1])), t0, t1);
            o[i] = key_ok ? pack_bf16(ex2_approx(t0), ex2_approx(t1)) : 0u;
          }
          *reinterpret_cast<uint4*>(prow + ((c >> 3) + g) * 2048) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    if (tid == 0) {
      tc_fence_after();
      // dP^T = V_kt dO^T  (overwrites the S^T columns: every thread has consumed its S^T row)
      const uint64_t ad = make_smem_desc(sv_t, 16, 1024, kLayoutSw128);
      const uint64_t bd = make_smem_desc(smem_u32(sDO), 16, 1024, kLayoutSw128);
#pragma unroll
      for (int k = 0; k < kHd / 16; ++k) umma_ss(tmem_base + kColSD, ad + 2 * k, bd + 2 * k, idesc_sd, k > 0 ? 1u : 0u);
      umma_commit(bar_dp);
      // dV_kt = P^T dO : A = sP (K-major core matrices: 8-query chunks 2048 B apart, 8-key groups 128 B apart),
      //                  B = dO consumed MN-major in place (16 queries = 2048 B per K step)
      const uint64_t pd = make_smem_desc(smem_u32(sP), 2048, 128, kLayoutNone);
      const uint64_t dod = make_smem_desc(smem_u32(sDO), 1024, 1024, kLayoutSw128);
#pragma unroll
      for (int k = 0; k < kNQ / 16; ++k)
        umma_ss(tmem_base + kColDV, pd + static_cast<uint64_t>(k * (4096 >> 4)), dod + static_cast<uint64_t>(k * (2048 >> 4)),
                idesc_kv, k > 0 ? 1u : 0u);
    }
    __syncwarp();
    mbar_wait(bar_dp, ph);
    tc_fence_after();

    // ---- pass B: dS^T = P^T * (dP^T - delta[q]) * scale  -> sDS
    uint8_t* dsrow = sDS + tid * 16;
    const f32x2 s2 = splat(p.scale);
#pragma unroll 1
    for (int c = 0; c < kNQ; c += 32) {
      if (c + 32 <= kNQ) {
        uint32_t r[32];
        tmem_ld_x32(tmem_row + kColSD + c, r);
        tmem_ld_wait_x32(r);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 pq = *reinterpret_cast<const uint4*>(prow + ((c >> 3) + g) * 2048);
          const uint32_t pw[4] = {pq.x, pq.y, pq.z, pq.w};
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int q = c + g * 8 + 2 * i;
            const f32x2 dp = add2(pk(__uint_as_float(r[g * 8 + 2 * i]), __uint_as_float(r[g * 8 + 2 * i + 1])),
                                  pk(-sDelta[q], -sDelta[q + 1]));
            float d0, d1;
            upk(mul2(mul2(bf16x2_to_f32x2(pw[i]), dp), s2), d0, d1);
            o[i] = pack_bf16(d0, d1);
          }
          *reinterpret_cast<uint4*>(dsrow + ((c >> 3) + g) * 2048) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      } else {
        uint32_t r[16];
        tmem_ld_x16(tmem_row + kColSD + c, r);
        tmem_ld_wait_x16(r);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const uint4 pq = *reinterpret_cast<const uint4*>(prow + ((c >> 3) + g) * 2048);
          const uint32_t pw[4] = {pq.x, pq.y, pq.z, pq.w};
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int q = c + g * 8 + 2 * i;
            const f32x2 dp = add2(pk(__uint_as_float(r[g * 8 + 2 * i]), __uint_as_float(r[g * 8 + 2 * i + 1])),
                                  pk(-sDelta[q], -sDelta[q + 1]));
            float d0, d1;
            upk(mul2(mul2(bf16x2_to_f32x2(pw[i]), dp), s2), d0, d1);
            o[i] = pack_bf16(d0, d1);
          }
          *reinterpret_cast<uint4*>(dsrow + ((c >> 3) + g) * 2048) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    if (tid == 0) {
      tc_fence_after();
      // dK_kt = dS^T Q : A = sDS K-major, B = Q MN-major in place
      const uint64_t dsd = make_smem_desc(smem_u32(sDS), 2048, 128, kLayoutNone);
      const uint64_t qd = make_smem_desc(smem_u32(sQ), 1024, 1024, kLayoutSw128);
#pragma unroll
      for (int k = 0; k < kNQ / 16; ++k)
        umma_ss(tmem_base + kColDK, dsd + static_cast<uint64_t>(k * (4096 >> 4)), qd + static_cast<uint64_t>(k * (2048 >> 4)),
                idesc_kv, k > 0 ? 1u : 0u);
      // dQ += dS K_kt : A = the same sDS tile read MN-major (M = queries: 8-query groups 2048 B apart (SBO), 8-key
      //                 groups 128 B apart (LBO)), K = the 128 keys of this tile; B = K_kt MN-major in place
      const uint64_t kd = make_smem_desc(sk_t, 1024, 1024, kLayoutSw128);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const uint64_t ad = make_smem_desc(smem_u32(sDS) + mt * 16 * 2048, 128, 2048, kLayoutNone);
#pragma unroll
        for (int k = 0; k < 128 / 16; ++k)
          umma_ss(tmem_base + (mt == 0 ? kColDQ0 : kColDQ1), ad + static_cast<uint64_t>(k * (256 >> 4)),
                  kd + static_cast<uint64_t>(k * (2048 >> 4)), idesc_dq, (kt > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(bar_out);   // covers dV (issued earlier), dK and dQ
    }
    __syncwarp();
    mbar_wait(bar_out, ph);
    tc_fence_after();

    // ---- dV, dK rows of this key tile -> dqkv[b, key, 2 / 1, h, :]
#pragma unroll
    for (int which = 0; which < 2; ++which) {   // 0: dV, 1: dK
      __nv_bfloat16* dst = p.dqkv + (static_cast<long long>(b) * N + key) * tok_stride + ((which == 0 ? 2 : 1) * H + h) * kHd;
#pragma unroll
      for (int c = 0; c < kHd; c += 32) {
        uint32_t r[32];
        tmem_ld_x32(tmem_row + (which == 0 ? kColDV : kColDK) + c, r);
        tmem_ld_wait_x32(r);
        if (key < N) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            reinterpret_cast<uint4*>(dst + c)[g] =
                make_uint4(pack_bf16(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1])),
                           pack_bf16(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3])),
                           pack_bf16(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5])),
                           pack_bf16(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7])));
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // TMEM S/dV/dK columns and sP / sDS are free for the next key tile
  }

  // keys beyond kv_len (padding keys of a NaFlex batch) never entered a tile: their dK / dV are zero
  for (int key = num_kt * 128 + tid; key < N; key += 128) {
    for (int which = 1; which <= 2; ++which) {
      uint4* dst = reinterpret_cast<uint4*>(p.dqkv + (static_cast<long long>(b) * N + key) * tok_stride + (which * H + h) * kHd);
#pragma unroll
      for (int g = 0; g < 8; ++g) dst[g] = make_uint4(0u, 0u, 0u, 0u);
    }
  }

  // ---- dQ rows: thread t owns queries t (tile 0) and 128 + t (tile 1)
  tc_fence_after();
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int q = mt * 128 + tid;
    __nv_bfloat16* dst = p.dqkv + (static_cast<long long>(b) * N + q) * tok_stride + h * kHd;
#pragma unroll
    for (int c = 0; c < kHd; c += 32) {
      uint32_t r[32];
      tmem_ld_x32(tmem_row + (mt == 0 ? kColDQ0 : kColDQ1) + c, r);
      tmem_ld_wait_x32(r);
      if (q < N) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          reinterpret_cast<uint4*>(dst + c)[g] =
              make_uint4(pack_bf16(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1])),
                         pack_bf16(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3])),
                         pack_bf16(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5])),
                         pack_bf16(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7])));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace b200

namespace b200 {
int launch_attn_bwd2(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv,
                     const int32_t* kv_len, const AttnBias* bias, int B, int N, int H, float scale, float* dbias, cudaStream_t st);
int launch_attn_bwd_long(const void* qkv, const void* dout, const float* lse, const float* delta, float* dq32, void* dqkv,
                         const int32_t* kv_len, const AttnBias* bias, int B, int N, int H, int hd, float scale, cudaStream_t st);
}
using namespace b200;

extern "C" int64_t b200_attn_bwd_ws_floats_hd(int32_t B, int32_t N, int32_t H, int32_t hd) {
  const long long d = (static_cast<long long>(B) * H * N + 3) / 4 * 4;
  return (N > kNQ || hd > kHd) ? d + static_cast<long long>(B) * N * H * hd : d;     // + the long kernel's fp32 dQ buffer
}

extern "C" int64_t b200_attn_bwd_ws_floats(int32_t B, int32_t N, int32_t H) { return b200_attn_bwd_ws_floats_hd(B, N, H, kHd); }

static int attn_bwd_v1(const void* qkv, const void* dout, const float* lse, float* delta, void* dqkv, const int32_t* kv_len,
                       int32_t B, int32_t N, int32_t H, float scale, cudaStream_t st);

static int attn_bwd_dispatch(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv,
                             const int32_t* kv_len, const AttnBias* bias, int32_t B, int32_t N, int32_t H, int32_t hd,
                             float scale, float* dbias, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(qkv && out && dout && lse && delta && dqkv, "b200_attn_bwd: null pointer");
  B200_CHECK_ARG(B > 0 && N > 0 && H > 0 && B <= 65535, "b200_attn_bwd: bad shape B=%d N=%d H=%d", B, N, H);
  if (hd != kHd && hd != 2 * kHd)
    return fail(B200_ERR_UNSUPPORTED, "b200_attn_bwd: head_dim %d not supported (64 or 128; other widths are zero-padded by the caller)", hd);
  {
    const long long rows = static_cast<long long>(B) * N;
    const int wpb = 8;
    auto dk = hd == kHd ? attn_bwd_delta_kernel<8> : attn_bwd_delta_kernel<16>;
    dk<<<static_cast<unsigned>((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(dout), delta, B, N, H);
    B200_CUDA_OK(cudaGetLastError());
  }
  static const bool use_v1 = [] { const char* e = getenv("B200_ATTN_BWD_V1"); return e != nullptr && e[0] == '1'; }();
  if (dbias != nullptr && (N > kNQ || hd != kHd || (use_v1 && bias == nullptr)))
    return fail(B200_ERR_UNSUPPORTED, "b200_attn_bwd: the fused qkv-bias gradient is available for N <= %d, head_dim 64 (v2 kernel) only", kNQ);
  if (N > kNQ || hd != kHd) {
    // long sequences and 128-wide heads (attention_bwd_long.cu): the workspace continues behind delta with the fp32 dq buffer
    float* dq32 = delta + (static_cast<long long>(B) * H * N + 3) / 4 * 4;
    return launch_attn_bwd_long(qkv, dout, lse, delta, dq32, dqkv, kv_len, bias, B, N, H, hd, scale, st);
  }
  // v2 (attention_bwd2.cu: two query halves, 8 compute warps + issuer warp) is the default; B200_ATTN_BWD_V1=1 keeps
  // the simpler v1 kernel of this file selectable for A/B comparison and as a cross-check in the tests (no bias variant).
  if (!use_v1 || bias != nullptr) return launch_attn_bwd2(qkv, dout, lse, delta, dqkv, kv_len, bias, B, N, H, scale, dbias, st);
  return attn_bwd_v1(qkv, dout, lse, delta, dqkv, kv_len, B, N, H, scale, st);
}

extern "C" int b200_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                             void* dqkv, const int32_t* kv_len, int32_t B, int32_t N, int32_t H, int32_t hd,
                             float scale, float* dbias, void* stream_) {
  return attn_bwd_dispatch(qkv, out, dout, lse, delta, dqkv, kv_len, nullptr, B, N, H, hd, scale, dbias, stream_);
}

extern "C" int b200_attn_bwd_bias(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                                  void* dqkv, const float* bias, int64_t bias_batch_stride, int64_t bias_head_stride,
                                  int32_t bias_ld, int32_t B, int32_t N, int32_t H, int32_t hd, float scale, float* dbias,
                                  void* stream_) {
  B200_CHECK_ARG(bias != nullptr, "b200_attn_bwd_bias: null bias");
  B200_CHECK_ARG(bias_ld >= N && bias_ld % 4 == 0 && bias_batch_stride >= 0 && bias_head_stride >= 0,
                 "b200_attn_bwd_bias: bias row pitch %d must be a multiple of 4 and >= N = %d", bias_ld, N);
  const AttnBias ab{bias, bias_batch_stride, bias_head_stride, bias_ld};
  return attn_bwd_dispatch(qkv, out, dout, lse, delta, dqkv, nullptr, &ab, B, N, H, hd, scale, dbias, stream_);
}

static int attn_bwd_v1(const void* qkv, const void* dout, const float* lse, float* delta, void* dqkv, const int32_t* kv_len,
                       int32_t B, int32_t N, int32_t H, float scale, cudaStream_t st) {
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
  AttnBwdParams p;
  p.lse = lse; p.delta = delta; p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv); p.kv_len = kv_len;
  p.B = B; p.N = N; p.H = H; p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    attr_set = true;
  }
  attn_bwd_kernel<<<dim3(H, B), 128, kSmemTotal, st>>>(tq, tdo, p);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}
