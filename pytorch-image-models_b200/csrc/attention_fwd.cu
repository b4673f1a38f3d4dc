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
#include "ptx.cuh"

namespace b200 {

constexpr int kHd = 64;
constexpr int kQTile = 128;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

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

template <int KB, bool OVERLAY>
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
  float acc[kHd];
#pragma unroll
  for (int i = 0; i < kHd; ++i) acc[i] = 0.f;

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

    // ---- pass 1: block row max over the valid keys
    const int key0 = j * KB;
    float mb = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < KB; c += 32) {
      if (c + 32 <= KB) {
        uint32_t r[32];
        tmem_ld_x32(tmem_row + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (key0 + c + i < kvlen) mb = fmaxf(mb, __uint_as_float(r[i]));
      } else {
        uint32_t r[16];
        tmem_ld_x16(tmem_row + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (key0 + c + i < kvlen) mb = fmaxf(mb, __uint_as_float(r[i]));
      }
    }
    const float m_new = fmaxf(m_run, mb);
    const float neg_mc = -m_new * p.scale_log2e;
    const float alpha = ex2(fmaf(m_run, p.scale_log2e, neg_mc));
    float lsum = 0.f;

    // ---- pass 2: p = exp2((s - m) * scale * log2e) -> bf16 core matrices in smem
    uint8_t* prow = sP + tid * 16;
#pragma unroll 1
    for (int c = 0; c < KB; c += 32) {
      if (c + 32 <= KB) {
        uint32_t r[32];
        tmem_ld_x32(tmem_row + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float s = __uint_as_float(r[g * 8 + i]);
            e[i] = (key0 + c + g * 8 + i < kvlen) ? ex2(fmaf(s, p.scale_log2e, neg_mc)) : 0.f;
            lsum += e[i];
          }
          uint4 q;
          q.x = pack_bf16(e[0], e[1]); q.y = pack_bf16(e[2], e[3]);
          q.z = pack_bf16(e[4], e[5]); q.w = pack_bf16(e[6], e[7]);
          *reinterpret_cast<uint4*>(prow + ((c >> 3) + g) * (kQTile * 16)) = q;
        }
      } else {
        uint32_t r[16];
        tmem_ld_x16(tmem_row + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float s = __uint_as_float(r[g * 8 + i]);
            e[i] = (key0 + c + g * 8 + i < kvlen) ? ex2(fmaf(s, p.scale_log2e, neg_mc)) : 0.f;
            lsum += e[i];
          }
          uint4 q;
          q.x = pack_bf16(e[0], e[1]); q.y = pack_bf16(e[2], e[3]);
          q.z = pack_bf16(e[4], e[5]); q.w = pack_bf16(e[6], e[7]);
          *reinterpret_cast<uint4*>(prow + ((c >> 3) + g) * (kQTile * 16)) = q;
        }
      }
    }
    l_run = fmaf(l_run, alpha, lsum);
    m_run = m_new;

    fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();

    if (tid == 0) {
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
#pragma unroll
    for (int c = 0; c < kHd; c += 32) {
      uint32_t r[32];
      tmem_ld_x32(tmem_row + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[c + i] = fmaf(acc[c + i], alpha, __uint_as_float(r[i]));
    }
    tc_fence_before();
    __syncthreads();  // S/O columns of TMEM and the K/V/P smem are free again
    if (tid == 0 && j + 1 < nblk) {
      mbar_arrive_expect_tx(bar_kv, 2 * L::kKBytes);
      tma_load_2d(sK, &tm_kv, bar_kv, col_k, b * N + (j + 1) * KB);
      tma_load_2d(sV, &tm_kv, bar_kv, col_v, b * N + (j + 1) * KB);
    }
  }

  if (row_in_img < N) {
    const float inv_l = 1.0f / l_run;
    __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * N + row_in_img) * (H * kHd) + h * kHd;
#pragma unroll
    for (int g = 0; g < kHd / 8; ++g) {
      uint4 q;
      q.x = pack_bf16(acc[g * 8 + 0] * inv_l, acc[g * 8 + 1] * inv_l);
      q.y = pack_bf16(acc[g * 8 + 2] * inv_l, acc[g * 8 + 3] * inv_l);
      q.z = pack_bf16(acc[g * 8 + 4] * inv_l, acc[g * 8 + 5] * inv_l);
      q.w = pack_bf16(acc[g * 8 + 6] * inv_l, acc[g * 8 + 7] * inv_l);
      reinterpret_cast<uint4*>(orow)[g] = q;
    }
    if (p.lse != nullptr)
      p.lse[(static_cast<long long>(b) * H + h) * N + row_in_img] = fmaf(m_run, p.scale, __logf(l_run));
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <int KB, bool OVERLAY>
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
  auto kern = attn_fwd_kernel<KB, OVERLAY>;
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

using namespace b200;

extern "C" int b200_attn_fwd(const void* qkv, void* out, float* lse, const int32_t* kv_len, int32_t B, int32_t N,
                             int32_t H, int32_t hd, float scale, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(qkv && out, "b200_attn_fwd: null qkv/out");
  B200_CHECK_ARG(B > 0 && N > 0 && H > 0, "b200_attn_fwd: bad shape B=%d N=%d H=%d", B, N, H);
  if (hd != kHd) return fail(B200_ERR_UNSUPPORTED, "b200_attn_fwd: head_dim %d not supported (64 only)", hd);
  B200_CHECK_ARG(B <= 65535 && H <= 65535, "b200_attn_fwd: B/H exceed the grid limits");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "b200_attn_fwd: qkv/out must be 16-byte aligned");
  if (N <= 208) return launch_attn_fwd<208, true>(qkv, out, lse, kv_len, B, N, H, scale, st);
  if (N <= 256) return launch_attn_fwd<256, true>(qkv, out, lse, kv_len, B, N, H, scale, st);
  return launch_attn_fwd<128, false>(qkv, out, lse, kv_len, B, N, H, scale, st);
}
