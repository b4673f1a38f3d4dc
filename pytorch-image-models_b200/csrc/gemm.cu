// Persistent, warp-specialised bf16 GEMM for sm_100a:  out[M,N] = epilogue(A[M,K] * B[N,K]^T)
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, STAGES deep)
//   warp 1      : MMA issuer     (one thread issues tcgen05.mma 128 x BN x 16, accumulators in TMEM,
//                                 two accumulator stages so the epilogue overlaps the next tile)
//   warp 2      : TMEM allocator / deallocator
//   warps 4..11 : epilogue       (tcgen05.ld -> bias / GELU / residual+LayerScale / dGELU -> global)
//
// Replaces F.linear (+ the separate GELU / residual / LayerScale ATen passes) of
// timm/layers/attention.py:115,137, timm/layers/mlp.py:48-52, timm/models/vision_transformer.py:212-213.
// Operands may be K-major (row-major [rows, K]) or MN-major (stored [K, rows]) so the same kernel
// serves forward, dgrad (B MN-major) and wgrad (A and B MN-major) without transposes.
#include "common.h"
#include "ptx.cuh"

namespace b200 {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kNumEpiWarps = 8;
constexpr int kNumThreads = 128 + kNumEpiWarps * 32;

struct GemmKernelParams {
  void* out;
  const void* bias;
  const __nv_bfloat16* aux;
  const void* gamma;
  long long ldo, ldaux;
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;  // num_k_blocks: total over K
  int split_k, k_blocks_per_split;
  int epilogue, out_f32, bias_f32, gamma_f32, accumulate;
};

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarrierOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarrierOffset + 256 + 1024;  // + barriers + alignment slack
};

__device__ __forceinline__ float gelu_erf(float x) {
  // 0.5 x (1 + erf(x / sqrt2)); erf by Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7), 2 MUFU + 8 FMA.
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = exp2f(-1.4426950408889634f * z * z);
  const float erf_abs = fmaf(-p, e, 1.0f);
  const float erf_v = copysignf(erf_abs, x);
  return 0.5f * x * (1.0f + erf_v);
}

__device__ __forceinline__ float dgelu_erf(float x) {
  // d/dx [x Phi(x)] = Phi(x) + x phi(x)
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = exp2f(-1.4426950408889634f * z * z);  // exp(-x^2/2)
  const float erf_v = copysignf(fmaf(-p, e, 1.0f), x);
  const float cdf = 0.5f * (1.0f + erf_v);
  return fmaf(x * 0.3989422804014327f, e, cdf);
}

__device__ __forceinline__ void load8_param(const void* base, int f32, int col, float (&v)[8]) {
  if (f32) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + col);
    float4 a = __ldg(p), b = __ldg(p + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + col));
    v[0] = bf16_lo(q.x); v[1] = bf16_hi(q.x); v[2] = bf16_lo(q.y); v[3] = bf16_hi(q.y);
    v[4] = bf16_lo(q.z); v[5] = bf16_hi(q.z); v[6] = bf16_lo(q.w); v[7] = bf16_hi(q.w);
  }
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmKernelParams p) {
  using L = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarrierOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kNumEpiWarps * 32);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr_smem, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int tiles_mn = p.num_m_blocks * p.num_n_blocks;
  const int num_tiles = tiles_mn * p.split_k;

  if (warp_idx == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int ks = tile / tiles_mn;
      const int mn = tile - ks * tiles_mn;
      const int m_blk = mn / p.num_n_blocks;
      const int n_blk = mn - m_blk * p.num_n_blocks;
      const int kb0 = ks * p.k_blocks_per_split;
      const int kb1 = min(kb0 + p.k_blocks_per_split, p.num_k_blocks);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * L::kStageBytes;
        uint8_t* sb = sa + L::kABytes;
        mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
        if constexpr (!A_MN) {
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
        } else {
#pragma unroll
          for (int i = 0; i < BM / 64; ++i)
            tma_load_2d(sa + i * (64 * BK * 2), &tmap_a, &full_bar[stage], m_blk * BM + i * 64, kb * BK);
        }
        if constexpr (!B_MN) {
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN);
        } else {
#pragma unroll
          for (int i = 0; i < BN / 64; ++i)
            tma_load_2d(sb + i * (64 * BK * 2), &tmap_b, &full_bar[stage], n_blk * BN + i * 64, kb * BK);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp_idx == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
    // K-major SW128: rows of 128 B, 8-row groups 1024 B apart (SBO); one UMMA_K=16 step = +32 B.
    // MN-major SW128: 64-element (128 B) runs along MN, 8 k-rows per 1024 B (SBO), next 64 MN elements
    //                 64*BK*2 B further (LBO); one UMMA_K=16 step = 16 k-rows = +2048 B.
    constexpr uint32_t a_lbo = A_MN ? 64 * BK * 2 : 16, b_lbo = B_MN ? 64 * BK * 2 : 16;
    constexpr uint32_t a_kstep = (A_MN ? 2048 : 32) >> 4, b_kstep = (B_MN ? 2048 : 32) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int ks = tile / tiles_mn;
      const int kb0 = ks * p.k_blocks_per_split;
      const int kb1 = min(kb0 + p.k_blocks_per_split, p.num_k_blocks);
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
        const uint32_t sb = sa + L::kABytes;
        const uint64_t adesc = make_smem_desc(sa, a_lbo, 1024, kLayoutSw128);
        const uint64_t bdesc = make_smem_desc(sb, b_lbo, 1024, kLayoutSw128);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          umma_ss(tmem_d, adesc + static_cast<uint64_t>(k * a_kstep), bdesc + static_cast<uint64_t>(k * b_kstep),
                  idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tfull_bar[as]);  // accumulator ready for the epilogue
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else if (warp_idx >= 4) {
    // ------------------------------------------------------------ epilogue
    const int ew = warp_idx - 4;
    const int quad = warp_idx & 3;          // TMEM lane quadrant this warp may access
    const int half = ew >> 2;               // which half of the BN columns
    constexpr int kColsPerWarp = BN / 2;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int ks = tile / tiles_mn;
      const int mn = tile - ks * tiles_mn;
      const int m_blk = mn / p.num_n_blocks;
      const int n_blk = mn - m_blk * p.num_n_blocks;
      const int row = m_blk * BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + half * kColsPerWarp;
#pragma unroll 1
      for (int c = 0; c < kColsPerWarp; c += 32) {
        uint32_t r[32];
        tmem_ld_x32(taddr0 + c, r);
        tmem_ld_wait();
        const int col0 = n_blk * BN + half * kColsPerWarp + c;
        if (row_ok && col0 < p.N) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = col0 + g * 8;
            if (col < p.N) {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
              if (p.bias != nullptr) {
                float bv[8];
                load8_param(p.bias, p.bias_f32, col, bv);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += bv[j];
              }
              if (p.epilogue == B200_EPI_BIAS_GELU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
              } else if (p.epilogue == B200_EPI_BIAS_RESIDUAL) {
                if (p.gamma != nullptr) {
                  float gv[8];
                  load8_param(p.gamma, p.gamma_f32, col, gv);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] *= gv[j];
                }
                float av[8];
                load8_param(p.aux + static_cast<long long>(row) * p.ldaux, 0, col, av);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += av[j];
              } else if (p.epilogue == B200_EPI_DGELU) {
                float av[8];
                load8_param(p.aux + static_cast<long long>(row) * p.ldaux, 0, col, av);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] *= dgelu_erf(av[j]);
              }
              if (p.out_f32) {
                float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col;
                if (p.split_k > 1) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) atomicAdd(o + j, v[j]);
                } else {
                  float4* o4 = reinterpret_cast<float4*>(o);
                  if (p.accumulate) {
                    float4 a = o4[0], b = o4[1];
                    v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
                    v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
                  }
                  o4[0] = make_float4(v[0], v[1], v[2], v[3]);
                  o4[1] = make_float4(v[4], v[5], v[6], v[7]);
                }
              } else {
                uint4 q;
                q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]);
                q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) +
                                          static_cast<long long>(row) * p.ldo + col) = q;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
static int launch_gemm(const b200_gemm_desc& d, cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES>;
  CUtensorMap ta, tb;
  {
    // A: K-major -> dims {K, M}, box {64, 128}; MN-major -> dims {M, K}, box {64, 64}
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    if (!A_MN) { dims[0] = d.K; dims[1] = d.M; box[0] = BK; box[1] = BM; }
    else       { dims[0] = d.M; dims[1] = d.K; box[0] = 64; box[1] = BK; }
    strides[0] = static_cast<uint64_t>(d.lda) * 2;
    int rc = encode_tmap_bf16(&ta, d.a, 2, dims, strides, box, true);
    if (rc) return rc;
    if (!B_MN) { dims[0] = d.K; dims[1] = d.N; box[0] = BK; box[1] = BN; }
    else       { dims[0] = d.N; dims[1] = d.K; box[0] = 64; box[1] = BK; }
    strides[0] = static_cast<uint64_t>(d.ldb) * 2;
    rc = encode_tmap_bf16(&tb, d.b, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  GemmKernelParams p;
  p.out = d.out; p.bias = d.bias; p.aux = reinterpret_cast<const __nv_bfloat16*>(d.aux); p.gamma = d.gamma;
  p.ldo = d.ldo; p.ldaux = d.ldaux;
  p.M = d.M; p.N = d.N; p.K = d.K;
  p.num_m_blocks = (d.M + BM - 1) / BM;
  p.num_n_blocks = (d.N + BN - 1) / BN;
  p.num_k_blocks = (d.K + BK - 1) / BK;
  int split = d.split_k > 1 ? d.split_k : 1;
  if (split > p.num_k_blocks) split = p.num_k_blocks;
  p.k_blocks_per_split = (p.num_k_blocks + split - 1) / split;
  p.split_k = (p.num_k_blocks + p.k_blocks_per_split - 1) / p.k_blocks_per_split;
  p.epilogue = d.epilogue; p.out_f32 = d.out_dtype == B200_DTYPE_F32;
  p.bias_f32 = d.bias_dtype == B200_DTYPE_F32; p.gamma_f32 = d.gamma_dtype == B200_DTYPE_F32;
  p.accumulate = d.accumulate;

  auto kern = gemm_bf16_kernel<BN, STAGES, A_MN, B_MN>;
  static bool attr_set = false;  // idempotent; a benign race sets the same value twice
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  const int num_tiles = p.num_m_blocks * p.num_n_blocks * p.split_k;
  int grid = sm_count();
  if (grid > num_tiles) grid = num_tiles;
  kern<<<grid, kNumThreads, L::kTotal, stream>>>(ta, tb, p);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_gemm(const b200_gemm_desc* dp, void* stream_) {
  if (dp == nullptr) return fail(B200_ERR_INVALID_ARG, "b200_gemm: null descriptor");
  const b200_gemm_desc& d = *dp;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(d.a && d.b && d.out, "b200_gemm: null operand pointer");
  B200_CHECK_ARG(d.M > 0 && d.N > 0 && d.K > 0, "b200_gemm: non-positive shape M=%d N=%d K=%d", d.M, d.N, d.K);
  B200_CHECK_ARG(d.N % 8 == 0, "b200_gemm: N=%d must be a multiple of 8", d.N);
  B200_CHECK_ARG(d.ldo % 8 == 0 && d.ldo >= d.N, "b200_gemm: ldo=%lld must be >= N and a multiple of 8",
                 (long long)d.ldo);
  B200_CHECK_ARG(d.lda % 8 == 0 && d.ldb % 8 == 0, "b200_gemm: lda/ldb must be multiples of 8 (16-byte TMA strides)");
  B200_CHECK_ARG(d.lda >= (d.a_mn ? d.M : d.K), "b200_gemm: lda=%lld too small", (long long)d.lda);
  B200_CHECK_ARG(d.ldb >= (d.b_mn ? d.N : d.K), "b200_gemm: ldb=%lld too small", (long long)d.ldb);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(d.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.b) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(d.out) & 15) == 0,
                 "b200_gemm: operand pointers must be 16-byte aligned");
  B200_CHECK_ARG(d.epilogue >= B200_EPI_BIAS && d.epilogue <= B200_EPI_DGELU, "b200_gemm: bad epilogue %d", d.epilogue);
  if (d.epilogue == B200_EPI_BIAS_RESIDUAL || d.epilogue == B200_EPI_DGELU) {
    B200_CHECK_ARG(d.aux != nullptr && d.ldaux % 8 == 0 && d.ldaux >= d.N &&
                       (reinterpret_cast<uintptr_t>(d.aux) & 15) == 0,
                   "b200_gemm: epilogue %d needs a 16-byte aligned aux with ldaux %% 8 == 0", d.epilogue);
  }
  if (d.bias) B200_CHECK_ARG((reinterpret_cast<uintptr_t>(d.bias) & 15) == 0, "b200_gemm: bias must be 16-byte aligned");
  if (d.gamma) B200_CHECK_ARG((reinterpret_cast<uintptr_t>(d.gamma) & 15) == 0, "b200_gemm: gamma must be 16-byte aligned");
  if (d.accumulate) B200_CHECK_ARG(d.out_dtype == B200_DTYPE_F32, "b200_gemm: accumulate needs an fp32 output");
  if (d.split_k > 1)
    B200_CHECK_ARG(d.out_dtype == B200_DTYPE_F32 && d.epilogue == B200_EPI_BIAS && d.bias == nullptr,
                   "b200_gemm: split_k needs fp32 output, EPI_BIAS and no bias");

  // Tile-N selection: 256 wide unless N is small (wasted columns) -- 128 then.
  const bool wide = d.N > 128;
  if (wide) {
    if (!d.a_mn && !d.b_mn) return launch_gemm<256, 4, false, false>(d, stream);
    if (!d.a_mn && d.b_mn) return launch_gemm<256, 4, false, true>(d, stream);
    if (d.a_mn && d.b_mn) return launch_gemm<256, 4, true, true>(d, stream);
    return launch_gemm<256, 4, true, false>(d, stream);
  } else {
    if (!d.a_mn && !d.b_mn) return launch_gemm<128, 6, false, false>(d, stream);
    if (!d.a_mn && d.b_mn) return launch_gemm<128, 6, false, true>(d, stream);
    if (d.a_mn && d.b_mn) return launch_gemm<128, 6, true, true>(d, stream);
    return launch_gemm<128, 6, true, false>(d, stream);
  }
}
