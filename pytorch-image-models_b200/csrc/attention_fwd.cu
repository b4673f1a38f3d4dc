// Fused scaled-dot-product attention forward for sm_100a, head_dim 64, reading q/k/v straight out of the
// QKV projection's [B, N, 3, H, 64] output through TMA (no permute / unbind copies) and writing the
// merged-head [B, N, H*64] layout the output projection consumes.
//
//   S = Q K^T   tcgen05.mma 128 x KB x 16 (x4), operands 128B-swizzled smem, S in TMEM (fp32)
//   softmax     one thread per query row (tcgen05.ld 32x32b), exp2 with the scale folded in, online
//               across key blocks; P written to smem as bf16 core matrices
//   O = P V     tcgen05.mma 128 x 64 x 16 (x KB/16), V consumed MN-major exactly as it sits in qkv
//
// Replaces timm/layers/attention.py:115-135 (reshape/permute/unbind + F.scaled_dot_product_attention +
// transpose/reshape); key-padding masks (naflexvit.py:964-1029) are given as per-sample key counts.
#include "common.h"
#include <cstdlib>
#include "ptx.cuh"

namespace b200 {

constexpr int kHd = 64;
constexpr int kQTile = 128;

template <int KB, bool OVERLAY>
struct AttnSmem {
  static constexpr int kQBytes = kQTile * kHd * 2;          // 16384
  static constexpr int kKBytes = KB * kHd * 2;              // KB * 128
  static constexpr int kPBytes = (KB / 8) * (kQTile * 16);  // KB/8 chunks of [128 rows x 16 B]
  static constexpr int kQOff = 0;
  static constexpr int kKOff = kQBytes;
  static constexpr int kPOff = OVERLAY ? 0 : kQBytes + kKBytes;
  static constexpr int kQKPEnd = OVERLAY ? (kPBytes > kQBytes + kKBytes ? kPBytes : kQBytes + kKBytes)
                                         : kQBytes + kKBytes + kPBytes;
  static constexpr int kVOff = (kQKPEnd + 1023) & ~1023;
  static constexpr int kBarOff = kVOff + kKBytes;
  static constexpr int kTotal = kBarOff + 64 + 1024;
};

struct AttnParams {
  __nv_bfloat16* out;
  float* lse;
  const int* kv_len;
  int B, N, H;
  float scale, scale_log2e;
};

template <int KB, bool OVERLAY, bool PTMEM>
__global__ void __launch_bounds__(128, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                const AttnParams p) {
  using L = AttnSmem<KB, OVERLAY>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + L::kQOff;
  uint8_t* sK = smem + L::kKOff;
  uint8_t* sV = smem + L::kVOff;
  uint8_t* sP = smem + L::kPOff;
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* bar_kv = bar_q + 1;
  uint64_t* bar_s = bar_q + 2;
  uint64_t* bar_o = bar_q + 3;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_q + 4);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int N = p.N, H = p.H;
  int kvlen = N;
  if (p.kv_len != nullptr) kvlen = max(1, min(p.kv_len[b], N));
  const int nblk = (kvlen + KB - 1) / KB;
  const int row_in_img = qt * kQTile + tid;
  const int q_row0 = b * N + qt * kQTile;
  const int col_q = h * kHd, col_k = (H + h) * kHd, col_v = (2 * H + h) * kHd;

  if (tid == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    mbar_init(bar_q, 1);
    mbar_init(bar_kv, 1);
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

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, L::kQBytes);
    tma_load_2d(sQ, &tm_q, bar_q, col_q, q_row0);
    mbar_arrive_expect_tx(bar_kv, 2 * L::kKBytes);
    tma_load_2d(sK, &tm_kv, bar_kv, col_k, b * N);
    tma_load_2d(sV, &tm_kv, bar_kv, col_v, b * N);
  }

  constexpr uint32_t idesc_s = make_idesc_bf16(kQTile, KB, false, false);
  constexpr uint32_t idesc_o = make_idesc_bf16(kQTile, kHd, false, true);

  float m_run = -INFINITY, l_run = 0.f;
  constexpr int kAcc = OVERLAY ? 1 : kHd / 2;   // multi-block variant keeps the running O in registers
  f32x2 acc[kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; ++i) acc[i] = 0ull;
  constexpr int kFull = KB / 32;                // 32-column chunks
  constexpr bool kTail16 = (KB % 32) != 0;      // + one 16-column chunk (KB = 208)

  for (int j = 0; j < nblk; ++j) {
    const uint32_t ph = j & 1;
    if (tid == 0) {
      if (j == 0) mbar_wait(bar_q, 0);
      mbar_wait(bar_kv, ph);
      tc_fence_after();
      const uint64_t qd = make_smem_desc(smem_u32(sQ), 16, 1024, kLayoutSw128);
      const uint64_t kd = make_smem_desc(smem_u32(sK), 16, 1024, kLayoutSw128);
#pragma unroll
      for (int k = 0; k < kHd / 16; ++k) umma_ss(tmem_base, qd + 2 * k, kd + 2 * k, idesc_s, k > 0 ? 1u : 0u);
      umma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_s, ph);
    tc_fence_after();

    const int key0 = j * KB;
    const int nvalid = kvlen - key0;            // keys of this block that are not padding (may exceed KB)

    // ---- pass 1: row max over the valid keys (software pipelined: load chunk c+1 while reducing chunk c)
    float mb0 = -INFINITY, mb1 = -INFINITY;
    {
      uint32_t r[2][32];
      tmem_ld_x32(tmem_row, r[0]);
#pragma unroll
      for (int c = 0; c < kFull; ++c) {
        tmem_ld_wait_x32(r[c & 1]);
        if (c + 1 < kFull) tmem_ld_x32(tmem_row + (c + 1) * 32, r[(c + 1) & 1]);
        else if (kTail16) tmem_ld_x16(tmem_row + kFull * 32, reinterpret_cast<uint32_t(&)[16]>(r[(c + 1) & 1]));
        if ((c + 1) * 32 <= nvalid) {
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
      if (kTail16) {
        uint32_t(&t)[16] = reinterpret_cast<uint32_t(&)[16]>(r[kFull & 1]);
        tmem_ld_wait_x16(t);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (kFull * 32 + i < nvalid) mb0 = fmaxf(mb0, __uint_as_float(t[i]));
      }
    }
    const float m_new = fmaxf(m_run, fmaxf(mb0, mb1));
    const float neg_mc = -m_new * p.scale_log2e;
    const float alpha = ex2_approx(fmaf(m_run, p.scale_log2e, neg_mc));
    const f32x2 sc2 = splat(p.scale_log2e), nm2 = splat(neg_mc);
    f32x2 lsum2 = 0ull;

    // ---- pass 2: p = exp2((s - m) * scale * log2e) -> bf16 core matrices in smem
    uint8_t* prow = sP + tid * 16;
    {
      uint32_t r[2][32];
      tmem_ld_x32(tmem_row, r[0]);
#pragma unroll
      for (int c = 0; c < kFull; ++c) {
        tmem_ld_wait_x32(r[c & 1]);
        if (c + 1 < kFull) tmem_ld_x32(tmem_row + (c + 1) * 32, r[(c + 1) & 1]);
        else if (kTail16) tmem_ld_x16(tmem_row + kFull * 32, reinterpret_cast<uint32_t(&)[16]>(r[(c + 1) & 1]));
        const bool full = (c + 1) * 32 <= nvalid;
        uint32_t pt[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float t0, t1;
            upk(fma2(pk(__uint_as_float(r[c & 1][g * 8 + 2 * i]), __uint_as_float(r[c & 1][g * 8 + 2 * i + 1])), sc2, nm2), t0, t1);
            float e0 = ex2_approx(t0), e1 = ex2_approx(t1);
            if (!full) {
              const int k = c * 32 + g * 8 + 2 * i;
              e0 = (k < nvalid) ? e0 : 0.f;
              e1 = (k + 1 < nvalid) ? e1 : 0.f;
            }
            lsum2 = add2(lsum2, pk(e0, e1));
            o[i] = pack_bf16(e0, e1);
          }
          if constexpr (PTMEM) {
#pragma unroll
            for (int i = 0; i < 4; ++i) pt[g * 4 + i] = o[i];
          } else {
            *reinterpret_cast<uint4*>(prow + (c * 4 + g) * (kQTile * 16)) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
        if constexpr (PTMEM) tmem_st_x16(tmem_row + c * 16, pt);   // P chunk c: bf16 pairs -> TMEM columns [16c, 16c+16)
      }
      if (kTail16) {
        uint32_t(&t)[16] = reinterpret_cast<uint32_t(&)[16]>(r[kFull & 1]);
        tmem_ld_wait_x16(t);
        uint32_t pt8[8];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float t0, t1;
            upk(fma2(pk(__uint_as_float(t[g * 8 + 2 * i]), __uint_as_float(t[g * 8 + 2 * i + 1])), sc2, nm2), t0, t1);
            const int k = kFull * 32 + g * 8 + 2 * i;
            const float e0 = (k < nvalid) ? ex2_approx(t0) : 0.f;
            const float e1 = (k + 1 < nvalid) ? ex2_approx(t1) : 0.f;
            lsum2 = add2(lsum2, pk(e0, e1));
            o[i] = pack_bf16(e0, e1);
          }
          if constexpr (PTMEM) {
#pragma unroll
            for (int i = 0; i < 4; ++i) pt8[g * 4 + i] = o[i];
          } else {
            *reinterpret_cast<uint4*>(prow + (kFull * 4 + g) * (kQTile * 16)) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
        if constexpr (PTMEM) tmem_st_x8(tmem_row + kFull * 16, pt8);
      }
    }
    {
      float s0, s1;
      upk(lsum2, s0, s1);
      l_run = fmaf(l_run, alpha, s0 + s1);
    }
    m_run = m_new;

    if constexpr (PTMEM) tmem_st_wait();
    else fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();

    constexpr int kOCol = PTMEM ? 128 : 0;   // P occupies TMEM columns [0, KB/2): O goes above it
    if (tid == 0 && PTMEM) {
      tc_fence_after();
      // O = P V with P read from TMEM (bf16 pairs, 8 columns per K=16 step), V MN-major from smem
      const uint64_t vd = make_smem_desc(smem_u32(sV), 1024, 1024, kLayoutSw128);
#pragma unroll
      for (int k = 0; k < KB / 16; ++k)
        umma_ts(tmem_base + kOCol, tmem_base + k * 8, vd + static_cast<uint64_t>(k * (2048 >> 4)), idesc_o, k > 0 ? 1u : 0u);
      umma_commit(bar_o);
    }
    if (tid == 0 && !PTMEM) {
      tc_fence_after();
      // P: K-major, no swizzle: core matrices 8 rows x 16 B; row groups 128 B apart (SBO), the two
      //    8-key chunks of one K=16 step 2048 B apart (LBO).   V: MN-major SW128, 16 keys = 2048 B.
      const uint64_t pd = make_smem_desc(smem_u32(sP), kQTile * 16, 128, kLayoutNone);
      const uint64_t vd = make_smem_desc(smem_u32(sV), 1024, 1024, kLayoutSw128);
#pragma unroll
      for (int k = 0; k < KB / 16; ++k)
        umma_ss(tmem_base, pd + static_cast<uint64_t>(k * ((2 * kQTile * 16) >> 4)),
                vd + static_cast<uint64_t>(k * (2048 >> 4)), idesc_o, k > 0 ? 1u : 0u);
      umma_commit(bar_o);
    }
    __syncwarp();
    mbar_wait(bar_o, ph);
    tc_fence_after();
    if constexpr (!OVERLAY) {
      const f32x2 al2 = splat(alpha);
#pragma unroll
      for (int c = 0; c < kHd; c += 32) {
        uint32_t r[32];
        tmem_ld_x32(tmem_row + kOCol + c, r);
        tmem_ld_wait_x32(r);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          acc[c / 2 + i] = fma2(acc[c / 2 + i], al2, pk(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])));
      }
      tc_fence_before();
      __syncthreads();  // S/O columns of TMEM and the K/V/P smem are free again
      if (tid == 0 && j + 1 < nblk) {
        mbar_arrive_expect_tx(bar_kv, 2 * L::kKBytes);
        tma_load_2d(sK, &tm_kv, bar_kv, col_k, b * N + (j + 1) * KB);
        tma_load_2d(sV, &tm_kv, bar_kv, col_v, b * N + (j + 1) * KB);
      }
    }
  }

  {
    // ---- epilogue: O / l -> bf16, merged-head layout; single-block variants read O straight from TMEM
    const f32x2 il2 = splat(1.0f / l_run);
    __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * N + row_in_img) * (H * kHd) + h * kHd;
#pragma unroll
    for (int c = 0; c < kHd; c += 32) {
      uint32_t o[16];
      if constexpr (OVERLAY) {
        uint32_t r[32];
        tmem_ld_x32(tmem_row + (PTMEM ? 128 : 0) + c, r);
        tmem_ld_wait_x32(r);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float lo, hi;
          upk(mul2(pk(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), il2), lo, hi);
          o[i] = pack_bf16(lo, hi);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float lo, hi;
          upk(mul2(acc[c / 2 + i], il2), lo, hi);
          o[i] = pack_bf16(lo, hi);
        }
      }
      if (row_in_img < N) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          reinterpret_cast<uint4*>(orow + c)[g] = make_uint4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
      }
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

template <int KB, bool OVERLAY, bool PTMEM>
static int launch_attn_fwd(const void* qkv, void* out, float* lse, const int32_t* kv_len, int B, int N, int H,
                           float scale, cudaStream_t st) {
  using L = AttnSmem<KB, OVERLAY>;
  CUtensorMap tq, tkv;
  uint64_t dims[2] = {static_cast<uint64_t>(3) * H * kHd, static_cast<uint64_t>(B) * N};
  uint64_t strides[1] = {static_cast<uint64_t>(3) * H * kHd * 2};
  uint32_t box_q[2] = {kHd, kQTile};
  uint32_t box_kv[2] = {kHd, KB};
  int rc = encode_tmap_bf16(&tq, qkv, 2, dims, strides, box_q, true);
  if (rc) return rc;
  rc = encode_tmap_bf16(&tkv, qkv, 2, dims, strides, box_kv, true);
  if (rc) return rc;
  AttnParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse; p.kv_len = kv_len; p.B = B; p.N = N; p.H = H;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  auto kern = attn_fwd_kernel<KB, OVERLAY, PTMEM>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  dim3 grid((N + kQTile - 1) / kQTile, H, B);
  kern<<<grid, 128, L::kTotal, st>>>(tq, tkv, p);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200

namespace b200 {
int launch_attn_fwd2(const void* qkv, void* out, float* lse, const int32_t* kv_len, const AttnBias* bias, int B, int N, int H,
                     float scale, cudaStream_t st);
int launch_attn_fwd_long(const void* qkv, void* out, float* lse, const int32_t* kv_len, const AttnBias* bias, int B, int N, int H,
                         int hd, float scale, cudaStream_t st);
}

using namespace b200;

static int attn_fwd_dispatch(const void* qkv, void* out, float* lse, const int32_t* kv_len, const AttnBias* bias, int32_t B,
                             int32_t N, int32_t H, int32_t hd, float scale, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(qkv && out, "b200_attn_fwd: null qkv/out");
  B200_CHECK_ARG(B > 0 && N > 0 && H > 0, "b200_attn_fwd: bad shape B=%d N=%d H=%d", B, N, H);
  if (hd != kHd && hd != 2 * kHd)
    return fail(B200_ERR_UNSUPPORTED, "b200_attn_fwd: head_dim %d not supported (64 or 128; other widths are zero-padded by the caller)", hd);
  B200_CHECK_ARG(B <= 65535 && H <= 65535, "b200_attn_fwd: B/H exceed the grid limits");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "b200_attn_fwd: qkv/out must be 16-byte aligned");
  // N <= 208 (ViT @224/16): the persistent, P-in-TMEM kernel of attention_fwd2.cu; B200_ATTN_FWD_V1=1 keeps the
  // one-CTA-per-tile kernels of this file selectable for A/B comparison (they remain the path for 208 < N <= 256).
  static const bool use_v1 = [] { const char* e = getenv("B200_ATTN_FWD_V1"); return e != nullptr && e[0] == '1'; }();
  // 128-wide heads: the online-softmax kernel of attention_fwd_long.cu with two 64-wide halves per operand, at any N
  if (hd == 2 * kHd) return launch_attn_fwd_long(qkv, out, lse, kv_len, bias, B, N, H, hd, scale, st);
  if (bias != nullptr) {     // additive bias: the two production kernels only
    if (N <= 208) return launch_attn_fwd2(qkv, out, lse, kv_len, bias, B, N, H, scale, st);
    return launch_attn_fwd_long(qkv, out, lse, kv_len, bias, B, N, H, kHd, scale, st);
  }
  if (!use_v1 && N <= 208) return launch_attn_fwd2(qkv, out, lse, kv_len, nullptr, B, N, H, scale, st);
  static const bool ptmem = [] { const char* e = getenv("B200_ATTN_PTMEM"); return e != nullptr && e[0] == '1'; }();
  if (ptmem && N <= 208) return launch_attn_fwd<208, true, true>(qkv, out, lse, kv_len, B, N, H, scale, st);
  if (N <= 208) return launch_attn_fwd<208, true, false>(qkv, out, lse, kv_len, B, N, H, scale, st);
  if (N <= 256) return launch_attn_fwd<256, true, false>(qkv, out, lse, kv_len, B, N, H, scale, st);
  // longer sequences: the double-buffered online-softmax kernel of attention_fwd_long.cu
  if (!use_v1) return launch_attn_fwd_long(qkv, out, lse, kv_len, nullptr, B, N, H, kHd, scale, st);
  return launch_attn_fwd<128, false, false>(qkv, out, lse, kv_len, B, N, H, scale, st);
}

extern "C" int b200_attn_fwd(const void* qkv, void* out, float* lse, const int32_t* kv_len, int32_t B, int32_t N,
                             int32_t H, int32_t hd, float scale, void* stream_) {
  return attn_fwd_dispatch(qkv, out, lse, kv_len, nullptr, B, N, H, hd, scale, stream_);
}

// Checks shared with b200_attn_bwd_bias (attention_bwd.cu has its own copy of the test).
static int check_bias(const char* who, const float* bias, int64_t bs, int64_t hs, int32_t ld, int32_t N) {
  B200_CHECK_ARG(bias != nullptr, "%s: null bias", who);
  B200_CHECK_ARG(ld >= N && ld % 4 == 0 && bs >= 0 && hs >= 0 && bs % 4 == 0 && hs % 4 == 0,
                 "%s: bias row pitch %d (>= N = %d) and the batch / head strides must be multiples of 4 elements", who, ld, N);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(bias) & 15) == 0, "%s: bias must be 16-byte aligned", who);
  return B200_OK;
}

extern "C" int b200_attn_fwd_bias(const void* qkv, void* out, float* lse, const float* bias, int64_t bias_batch_stride,
                                  int64_t bias_head_stride, int32_t bias_ld, int32_t B, int32_t N, int32_t H, int32_t hd,
                                  float scale, void* stream_) {
  const int rc = check_bias("b200_attn_fwd_bias", bias, bias_batch_stride, bias_head_stride, bias_ld, N);
  if (rc) return rc;
  const AttnBias ab{bias, bias_batch_stride, bias_head_stride, bias_ld};
  return attn_fwd_dispatch(qkv, out, lse, nullptr, &ab, B, N, H, hd, scale, stream_);
}
