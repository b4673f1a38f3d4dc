// Persistent, warp-specialised bf16 GEMM for sm_100a:  out[M,N] = epilogue(A[M,K] * B[N,K]^T)
//
// Default mode -- CTA pairs (template CTA2): the grid is clusters of two CTAs on one TPC; a pair owns a 256 x 256
// output tile.  Each CTA TMA-loads its own 128 rows of A and HALF of the B tile per k-block (32 KB instead of 48 KB
// through L2 / TMA / shared memory -- single CTAs were L2->SM operand-bound), the leader issues
// tcgen05.mma.cta_group::2 (M = 256) which reads both B halves, and each CTA's TMEM receives its 128 x 256 accumulator.
// Single CTAs (128 x BN tiles) remain for M < 256 / N <= 128 and behind B200_GEMM_2CTA=0.
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, STAGES deep; in a pair both CTAs'
//                                 bytes complete on the leader's mbarrier)
//   warp 1      : MMA issuer     (one thread of the leader issues the MMAs; accumulators in TMEM, two accumulator
//                                 stages so the epilogue overlaps the next tile; tcgen05.commit multicasts the
//                                 "slot free" / "accumulator ready" arrivals to both CTAs)
//   warp 2      : TMEM allocator / deallocator
//   warps 4..19 : epilogue       (tcgen05.ld -> bias / GELU / residual+LayerScale+DropPath / dGELU / RoPE -> swizzled
//                                 smem staging tile -> TMA store, in rounds of RC columns; the residual /
//                                 pre-activation operand comes in through a TMA load into the same staging tile, so
//                                 every global access of the epilogue is a full-line bulk copy; fp32 outputs
//                                 accumulate / split-K through TMA reduce-add; one remote mbarrier arrive per warp
//                                 hands the accumulator back to the leader)
//
// Replaces F.linear (+ the separate GELU / residual / LayerScale ATen passes) of
// timm/layers/attention.py:115,137, timm/layers/mlp.py:48-52, timm/models/vision_transformer.py:212-213.
// Operands may be K-major (row-major [rows, K]) or MN-major (stored [K, rows]) so the same kernel
// serves forward, dgrad (B MN-major) and wgrad (A and B MN-major) without transposes.
#include "common.h"
#include "ptx.cuh"
#include "gelu.cuh"

namespace b200 {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kNumEpiWarps = 16;
constexpr int kNumThreads = 128 + kNumEpiWarps * 32;

struct GemmKernelParams {
  void* out;
  void* out2;
  const float* row_scale;
  int rows_per_scale;
  const void* bias;
  const __nv_bfloat16* aux;
  const void* gamma;
  long long ldo, ldaux;
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;  // num_k_blocks: total over K
  int split_k, k_blocks_per_split;
  int epilogue, out_f32, bias_f32, gamma_f32, accumulate;
  const float* rope;           // EPI_BIAS_ROPE: [rope_tokens - rope_prefix][128] = [sin | cos]
  int rope_tokens, rope_prefix, rope_cols;
  int act;                     // activation of the GELU / DGELU epilogues in the generic kernels: 0 erf GELU, 1 tanh GELU, 2 SiLU
  // Patch-embedding mode (b200_patch_embed_fwd): the A operand comes straight from the NCHW image through a 5-D tensor map
  // {kx % 8, px, (b, c, py), kx / 8, ky} -- no patch matrix in HBM.  An "m-block" is one tile of pe_pyc grid rows of one
  // image (pe_tile_rows = pe_pyc * pe_gw <= 128 patch rows; the accumulator rows above that are never stored), a k-block
  // is 64 / p kernel rows of one channel.  The box lands as the UN-swizzled K-major core-matrix layout (8 rows x 16 bytes
  // per core matrix, row pitch 16 B, K chunks pe_lbo = 16 * pe_tile_rows bytes apart): a swizzled box would need its
  // innermost dimension to span the whole 128-byte row, but a patch row is p * 2 = 32 contiguous bytes of the image
  // (measured: with SWIZZLE_128B every 32-byte inner row is placed on its own 128-byte line).  Output / residual
  // coordinates are per image (3-D output map {D, tokens, B}).
  int patch, pe_ntiles, pe_pyc, pe_gw, pe_gh, pe_C, pe_kb_per_c, pe_ky_per_kb, pe_tile_rows, pe_row_off, pe_pos_off, pe_stage_tx, pe_lbo;
  float* colsum;               // optional fp32 [N]: += column sums of the bf16 output (the bias gradient of the Linear whose
                               // output gradient this GEMM produces), reduced from the staging tile, red.global.add
  // LayerNorm folded into the GEMMs either side of it (inference; b200_gemm_desc.ln_*):
  //   producer (ln_stats_out): every epilogue warp adds up, per accumulator row, the sum and the sum of squares of the bf16
  //     values it stores (its BN/4 columns) and writes the pair to slot n_blk * 4 + part of a [slots][M][2] fp32 buffer --
  //     one coalesced 8-byte store per lane and tile, no atomics, the same bits on every run;
  //   consumer (ln_stats): A is the UN-normalised activation, B = weight * ln_gamma; ln_finalize_kernel first turns the slots
  //     into one (rstd, -mean * rstd) pair per row (slot index ln_slots of the same buffer), the epilogue reads its row's pair
  //     one tile ahead and out = rstd * (acc - mean * c1[n]) + c2[n] replaces the bias add
  //     (c1[n] = sum_k B[n,k], c2[n] = sum_k ln_beta[k] W[n,k] + bias[n] arrive as ln_c1 and as the fp32 `bias`).
  float* ln_stats_out;
  const float* ln_stats;
  const float* ln_c1;
  int ln_slots;
  float ln_eps, ln_inv_k;
  unsigned long long* trace;   // debug: 12 timestamps / wait totals per CTA (b200_set_trace), nullptr in production
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <int BN, int STAGES, bool CTA2 = false, int RC = 32>
struct GemmSmem {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (CTA2 ? BN / 2 : BN) * BK * 2;   // a CTA pair holds half of the B tile each
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingOffset = STAGES * kStageBytes;   // kNumEpiWarps staging tiles of 32 rows x 2*RC bytes
  static constexpr int kStagingBytes = kNumEpiWarps * 32 * RC * 2;
  static constexpr int kBarrierOffset = kStagingOffset + kStagingBytes;
  static constexpr int kTotal = kBarrierOffset + 512 + 1024;  // + barriers + alignment slack
};

// 8 consecutive parameters (bias / LayerScale gamma) starting at column `col`, as 4 fp32 pairs
__device__ __forceinline__ void load8_param2(const void* base, int f32, int col, f32x2 (&v)[4]) {
  if (f32) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + col);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    v[0] = pk(a.x, a.y); v[1] = pk(a.z, a.w); v[2] = pk(b.x, b.y); v[3] = pk(b.z, b.w);
  } else {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + col));
    v[0] = bf16x2_to_f32x2(q.x); v[1] = bf16x2_to_f32x2(q.y); v[2] = bf16x2_to_f32x2(q.z); v[3] = bf16x2_to_f32x2(q.w);
  }
}

// EPI >= 0 fixes the epilogue (and bf16 output) at compile time -- the transcendental epilogues get straight-line code
// without the per-chunk epilogue-type branches; EPI = -1 is the generic kernel that reads p.epilogue at run time.
// ACTRT: the GELU / DGELU epilogues of the generic kernel read the activation kind from p.act (tanh GELU, SiLU); the
// default instantiations (ACTRT = false) are the exact-GELU kernels with no trace of that switch.
// LNF: the folded-LayerNorm epilogues (GemmKernelParams::ln_*) exist only in the instantiations built with LNF = 1 (consumer)
// or LNF = 2 (producer) -- K-major pair kernels of the inference forward -- so that every other kernel is untouched by them.
template <int BN, int STAGES, bool A_MN, bool B_MN, bool CTA2, int RC, int EPI, bool ACTRT, int LNF>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_aux,
                 const __grid_constant__ CUtensorMap tmap_out2, const GemmKernelParams p) {
  using L = GemmSmem<BN, STAGES, CTA2, RC>;
  static_assert(RC == 32 || RC == 64, "staging rows are 64 or 128 bytes");
  static_assert((BN / 4) % RC == 0, "a warp's BN/4 columns are a whole number of rounds");
  // CTA2: the kernel runs as clusters of two CTAs on one TPC.  The pair owns a 256 x BN output tile: each CTA loads
  // its own 128 rows of A and HALF of the B tile, the leader (rank 0) issues tcgen05.mma.cta_group::2 (M = 256) which
  // reads both B halves, and each CTA's TMEM receives its own 128 x BN accumulator -- per SM 32 KB instead of 48 KB
  // of operands per k-block come through L2 / TMA / shared memory.
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarrierOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* aux_bar = tempty_bar + 2;  // one per epilogue warp
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(aux_bar + kNumEpiWarps);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  unsigned long long* const trace = p.trace ? p.trace + blockIdx.x * 12 : nullptr;
  if (trace && threadIdx.x == 0) {
    trace[0] = gtime();
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    trace[11] = smid;
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kNumEpiWarps * (CTA2 ? 2 : 1));   // one arrival per epilogue warp (of both CTAs)
    }
    for (int i = 0; i < kNumEpiWarps; ++i) mbar_init(&aux_bar[i], 1);
    tma_prefetch_desc(&tmap_out);
    tma_prefetch_desc(&tmap_aux);
    tma_prefetch_desc(&tmap_out2);
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    if constexpr (CTA2) {
      tmem_alloc_pair(tmem_ptr_smem, 2 * BN);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_ptr_smem, 2 * BN);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // programmatic dependent launch: the set-up above ran under the previous kernel's tail; from here on global memory
  // written by it is read (operands, residual) or overwritten, so wait for it -- and let our own successor set up early
  griddep_launch();
  griddep_wait();
  if (trace && threadIdx.x == 0) trace[1] = gtime();

  // CTA2: tiles are 256-row pairs of m-blocks, walked by cluster; this CTA's m-block is 2 * pair + rank
  const int num_m_units = CTA2 ? (p.num_m_blocks + 1) / 2 : p.num_m_blocks;
  const int tiles_mn = num_m_units * p.num_n_blocks;
  const int num_tiles = tiles_mn * p.split_k;
  const int tile0 = CTA2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = CTA2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp_idx == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    long long wait_slot = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int ks = tile / tiles_mn;
      const int mn = tile - ks * tiles_mn;
      const int m_unit = mn / p.num_n_blocks;
      const int n_blk = mn - m_unit * p.num_n_blocks;
      const int m_blk = CTA2 ? 2 * m_unit + static_cast<int>(cta_rank) : m_unit;
      const int n_row0 = n_blk * BN + (CTA2 ? static_cast<int>(cta_rank) * (BN / 2) : 0);   // first B row this CTA loads
      const int kb0 = ks * p.k_blocks_per_split;
      const int kb1 = min(kb0 + p.k_blocks_per_split, p.num_k_blocks);
      for (int kb = kb0; kb < kb1; ++kb) {
        if (trace) {
          const long long c0 = clock64();
          mbar_wait(&empty_bar[stage], phase ^ 1);
          wait_slot += clock64() - c0;
        } else {
          mbar_wait(&empty_bar[stage], phase ^ 1);
        }
        uint8_t* sa = smem + stage * L::kStageBytes;
        uint8_t* sb = sa + L::kABytes;
        if constexpr (CTA2) {
          // both CTAs' bytes complete on the LEADER's barrier (the leader alone issues the MMAs)
          if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], p.patch ? p.pe_stage_tx : 2 * L::kStageBytes);
          if constexpr (!A_MN) {
            if (p.patch) {
              const int img = m_blk / p.pe_ntiles, ti = m_blk - img * p.pe_ntiles;
              const int ch = kb / p.pe_kb_per_c, j = kb - ch * p.pe_kb_per_c;
              tma_load_5d_pair(sa, &tmap_a, &full_bar[stage], 0, 0, (img * p.pe_C + ch) * p.pe_gh + ti * p.pe_pyc, 0, j * p.pe_ky_per_kb);
            } else {
              tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
            }
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              tma_load_2d_pair(sa + i * (64 * BK * 2), &tmap_a, &full_bar[stage], m_blk * BM + i * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            tma_load_2d_pair(sb, &tmap_b, &full_bar[stage], kb * BK, n_row0);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 128; ++i)
              tma_load_2d_pair(sb + i * (64 * BK * 2), &tmap_b, &full_bar[stage], n_row0 + i * 64, kb * BK);
          }
        } else {
          mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
          if constexpr (!A_MN) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              tma_load_2d(sa + i * (64 * BK * 2), &tmap_a, &full_bar[stage], m_blk * BM + i * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n_row0);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_2d(sb + i * (64 * BK * 2), &tmap_b, &full_bar[stage], n_row0 + i * 64, kb * BK);
          }
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    if (trace) trace[9] = wait_slot;
  } else if (warp_idx == 1 && lane == 0 && cta_rank == 0) {
    // ------------------------------------------------------------ MMA issuer (the leader CTA of a pair)
    constexpr uint32_t idesc = make_idesc_bf16(CTA2 ? 2 * BM : BM, BN, A_MN, B_MN);
    // K-major SW128: rows of 128 B, 8-row groups 1024 B apart (SBO); one UMMA_K=16 step = +32 B.
    // MN-major SW128: 64-element (128 B) runs along MN, 8 k-rows per 1024 B (SBO), next 64 MN elements
    //                 64*BK*2 B further (LBO); one UMMA_K=16 step = 16 k-rows = +2048 B.
    constexpr uint32_t a_lbo = A_MN ? 64 * BK * 2 : 16, b_lbo = B_MN ? 64 * BK * 2 : 16;
    constexpr uint32_t a_kstep = (A_MN ? 2048 : 32) >> 4, b_kstep = (B_MN ? 2048 : 32) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    long long wait_ops = 0, wait_acc = 0;
    unsigned long long first_ops = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int ks = tile / tiles_mn;
      const int kb0 = ks * p.k_blocks_per_split;
      const int kb1 = min(kb0 + p.k_blocks_per_split, p.num_k_blocks);
      if (trace) {
        const long long c0 = clock64();
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        wait_acc += clock64() - c0;
      } else {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
      }
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (trace) {
          const long long c0 = clock64();
          mbar_wait(&full_bar[stage], phase);
          wait_ops += clock64() - c0;
          if (first_ops == 0) first_ops = gtime();
        } else {
          mbar_wait(&full_bar[stage], phase);
        }
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
        const uint32_t sb = sa + L::kABytes;
        // patch mode: A is the un-swizzled core-matrix layout (K chunks of 8 elements pe_lbo bytes apart, 8-row groups 128 B
        // apart); one UMMA_K = 16 step = two K chunks
        const uint64_t adesc = p.patch ? make_smem_desc(sa, static_cast<uint32_t>(p.pe_lbo), 128, kLayoutNone)
                                       : make_smem_desc(sa, a_lbo, 1024, kLayoutSw128);
        const uint32_t akstep = p.patch ? static_cast<uint32_t>(2 * p.pe_lbo) >> 4 : a_kstep;
        const uint64_t bdesc = make_smem_desc(sb, b_lbo, 1024, kLayoutSw128);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          if constexpr (CTA2)
            umma_ss_pair(tmem_d, adesc + static_cast<uint64_t>(k * akstep), bdesc + static_cast<uint64_t>(k * b_kstep),
                         idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          else
            umma_ss(tmem_d, adesc + static_cast<uint64_t>(k * akstep), bdesc + static_cast<uint64_t>(k * b_kstep),
                    idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
        if constexpr (CTA2) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      // accumulator ready for the epilogue (of both CTAs)
      if constexpr (CTA2) umma_commit_pair(&tfull_bar[as]); else umma_commit(&tfull_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (trace) { trace[7] = wait_ops; trace[8] = wait_acc; trace[2] = first_ops; }
  } else if (warp_idx >= 4) {
    // ------------------------------------------------------------ epilogue
    // Each warp owns 32 accumulator rows x BN/4 columns of the tile and moves them in rounds of RC bf16 columns
    // (RC/2 fp32 columns): TMEM -> registers (32 columns per tcgen05.ld) -> epilogue math -> swizzled staging tile
    // (32 rows x 2*RC bytes, the TMA box layout) -> one TMA store per round.  With RC = BN/4 a tile is a single
    // round per warp: the residual / pre-activation tile is requested before the accumulator is even ready and the
    // staging tile is only reused a whole mainloop later, so no TMA latency sits on the epilogue's critical path.
    const int ew = warp_idx - 4;
    const int quad = warp_idx & 3;          // TMEM lane quadrant this warp may access
    const int part = ew >> 2;               // which quarter of the BN columns
    constexpr int kColsPerWarp = BN / 4;
    constexpr uint32_t kRowBytes = RC * 2;                     // 64 (64B swizzle) or 128 (128B swizzle)
    constexpr uint32_t kStgBytes = 32 * kRowBytes;
    uint8_t* stg = smem + L::kStagingOffset + ew * kStgBytes;
    const uint32_t stg_row = smem_u32(stg) + lane * kRowBytes;
    const uint32_t sw = RC == 32 ? static_cast<uint32_t>((lane >> 1) & 3) : static_cast<uint32_t>(lane & 7);
    const bool dual = p.out2 != nullptr;   // training: out2 <- pre-activation / un-scaled branch, out <- final value
    // residual / pre-activation tile fetched by TMA into the staging tile (the dual residual mode reads it later)
    const int epi = EPI >= 0 ? EPI : p.epilogue;
    const bool out_f32 = EPI >= 0 ? false : p.out_f32 != 0;
    const bool has_aux = (epi == B200_EPI_BIAS_RESIDUAL && !dual) || epi == B200_EPI_DGELU;
    const bool reduce_out = out_f32 && (p.accumulate || p.split_k > 1);
    const int round_cols = out_f32 ? RC / 2 : RC;     // output columns per staging tile
    const int sub_cols = out_f32 ? 16 : 32;           // output columns per 64-byte span (one tcgen05.ld)
    constexpr bool ln_produce = LNF == 2 && EPI < 0;   // row statistics of the stored tile (see GemmKernelParams)
    constexpr bool ln_consume = LNF == 1;
    uint32_t aux_phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    unsigned long long first_acc = 0, last_acc = 0, ntiles = 0;
    // folded LayerNorm, consumer side: (rstd, -mean * rstd) of this lane's row, fetched one tile ahead
    const float2* ln_fin = ln_consume ? reinterpret_cast<const float2*>(p.ln_stats) + static_cast<long long>(p.ln_slots) * p.M : nullptr;
    auto ln_row_of = [&](int t) {
      const int m_unit_t = (t % tiles_mn) / p.num_n_blocks;
      return (CTA2 ? 2 * m_unit_t + static_cast<int>(cta_rank) : m_unit_t) * BM + quad * 32 + lane;
    };
    float2 ln_next = make_float2(0.f, 0.f);
    if (ln_consume && tile0 < num_tiles) {
      const int r = ln_row_of(tile0);
      if (r < p.M) ln_next = __ldg(ln_fin + r);
    }
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int ks = tile / tiles_mn;
      const int mn = tile - ks * tiles_mn;
      const int m_unit = mn / p.num_n_blocks;
      const int n_blk = mn - m_unit * p.num_n_blocks;
      const int m_blk = CTA2 ? 2 * m_unit + static_cast<int>(cta_rank) : m_unit;
      const int row0 = m_blk * BM + quad * 32;
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + part * kColsPerWarp;
      // per-row scale of the residual branch (DropPath keep-mask / keep_prob of the row's sample, drop.py:158-175)
      f32x2 rs2 = splat(1.0f);
      if (p.row_scale != nullptr && row0 + lane < p.M) rs2 = splat(__ldg(p.row_scale + (row0 + lane) / p.rows_per_scale));
      // rotary embedding: this lane's row is token (row % N) of its sample; prefix tokens / rows past M pass through
      const float* rope_row = nullptr;
      if (EPI < 0 && epi == B200_EPI_BIAS_ROPE && row0 + lane < p.M) {
        const int tok = (row0 + lane) % p.rope_tokens - p.rope_prefix;
        if (tok >= 0) rope_row = p.rope + static_cast<long long>(tok) * 128;
      }
      const f32x2 ln_rstd2 = splat(ln_next.x), ln_nmr2 = splat(ln_next.y);
      if (ln_consume && tile + tile_step < num_tiles) {
        const int r = ln_row_of(tile + tile_step);
        ln_next = r < p.M ? __ldg(ln_fin + r) : make_float2(0.f, 0.f);
      }
      f32x2 ln_sum2 = splat(0.f), ln_sq2 = splat(0.f);
      bool waited = false;
#pragma unroll 1
      for (int rc = 0; rc < kColsPerWarp; rc += round_cols) {
        const int col0 = n_blk * BN + part * kColsPerWarp + rc;
        // the staging tile is free once the previous TMA store has finished READING it
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
        if (has_aux && lane == 0) {
          mbar_arrive_expect_tx(&aux_bar[ew], kStgBytes);
          // patch mode: the residual is the position table, indexed by the patch's place inside its image
          const int aux_row = p.patch ? (m_blk % p.pe_ntiles) * p.pe_pyc * p.pe_gw + quad * 32 + p.pe_pos_off : row0;
          tma_load_2d(stg, &tmap_aux, &aux_bar[ew], col0, aux_row);
        }
        if (!waited) {
          mbar_wait(&tfull_bar[as], aphase);
          tc_fence_after();
          waited = true;
          if (trace && ew == 0 && lane == 0) {
            last_acc = gtime();
            if (first_acc == 0) first_acc = last_acc;
            ++ntiles;
          }
        }
        bool aux_waited = false;
#pragma unroll 1
        for (int sub = 0; sub < round_cols; sub += sub_cols) {
          const uint32_t cb = static_cast<uint32_t>(sub / sub_cols) * 4;   // first 16-byte chunk of this 64-byte span
          if (out_f32) {
            // fp32 output (weight gradients): 16 columns = 4 chunks of 4 floats
            uint32_t r[16];
            tmem_ld_x16(taddr0 + rc + sub, r);
            tmem_ld_wait_x16(r);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int col = col0 + sub + g * 4;
              float v0 = __uint_as_float(r[g * 4 + 0]), v1 = __uint_as_float(r[g * 4 + 1]);
              float v2 = __uint_as_float(r[g * 4 + 2]), v3 = __uint_as_float(r[g * 4 + 3]);
              if (p.bias != nullptr && col < p.N) {
                if (p.bias_f32) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.bias) + col));
                  v0 += b4.x; v1 += b4.y; v2 += b4.z; v3 += b4.w;
                } else {
                  const uint2 b2 = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.bias) + col));
                  v0 += bf16_lo(b2.x); v1 += bf16_hi(b2.x); v2 += bf16_lo(b2.y); v3 += bf16_hi(b2.y);
                }
              }
              const uint32_t addr = stg_row + (((cb + static_cast<uint32_t>(g)) ^ sw) << 4);
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
            }
          } else {
            uint32_t r[32];
            tmem_ld_x32(taddr0 + rc + sub, r);
            tmem_ld_wait_x32(r);
            if (has_aux && !aux_waited) {
              mbar_wait(&aux_bar[ew], aux_phase);
              aux_phase ^= 1;
              aux_waited = true;
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int col = col0 + sub + g * 8;
              const uint32_t addr = stg_row + (((cb + static_cast<uint32_t>(g)) ^ sw) << 4);
              f32x2 v[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) v[j] = pk(__uint_as_float(r[g * 8 + 2 * j]), __uint_as_float(r[g * 8 + 2 * j + 1]));
              if (col < p.N) {
                if (ln_consume) {       // LayerNorm(x) W^T + b from x (W * gamma)^T: rstd * (acc - mean * c1) + c2
                  f32x2 c1v[4], bv[4];
                  load8_param2(p.ln_c1, 1, col, c1v);
                  load8_param2(p.bias, 1, col, bv);
#pragma unroll
                  for (int j = 0; j < 4; ++j) v[j] = fma2(v[j], ln_rstd2, fma2(c1v[j], ln_nmr2, bv[j]));
                } else if (p.bias != nullptr) {
                  f32x2 bv[4];
                  load8_param2(p.bias, p.bias_f32, col, bv);
#pragma unroll
                  for (int j = 0; j < 4; ++j) v[j] = add2(v[j], bv[j]);
                }
                if (EPI < 0 && epi == B200_EPI_BIAS_ROPE) {
                  if (rope_row != nullptr && col < p.rope_cols) {
                    // y[2i] = x[2i] cos[2i] - x[2i+1] sin[2i];  y[2i+1] = x[2i+1] cos[2i+1] + x[2i] sin[2i+1]
                    const float4* e = reinterpret_cast<const float4*>(rope_row + (col & 63));
                    const float4 s0 = __ldg(e), s1 = __ldg(e + 1), c0 = __ldg(e + 16), c1 = __ldg(e + 17);
                    const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                    const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      float x0, x1;
                      upk(v[j], x0, x1);
                      const uint32_t xb = pack_bf16(x0, x1);           // the bf16 value nn.Linear would have emitted
                      x0 = bf16_lo(xb);
                      x1 = bf16_hi(xb);
                      v[j] = pk(fmaf(x0, cs[2 * j], -x1 * sn[2 * j]), fmaf(x1, cs[2 * j + 1], x0 * sn[2 * j + 1]));
                    }
                  }
                } else if (epi == B200_EPI_BIAS_GELU && !dual) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) v[j] = (EPI < 0 && ACTRT && p.act) ? act_fwd2(v[j], p.act) : gelu_erf2(v[j]);
                } else if (has_aux && !dual) {
                  uint4 q;
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(addr));
                  const f32x2 av[4] = {bf16x2_to_f32x2(q.x), bf16x2_to_f32x2(q.y), bf16x2_to_f32x2(q.z), bf16x2_to_f32x2(q.w)};
                  if (epi == B200_EPI_BIAS_RESIDUAL) {
                    if (p.gamma != nullptr) {
                      f32x2 gv[4];
                      load8_param2(p.gamma, p.gamma_f32, col, gv);
#pragma unroll
                      for (int j = 0; j < 4; ++j) v[j] = fma2(v[j], mul2(gv[j], rs2), av[j]);
                    } else {
#pragma unroll
                      for (int j = 0; j < 4; ++j) v[j] = fma2(v[j], rs2, av[j]);
                    }
                  } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                      v[j] = mul2(v[j], (EPI < 0 && ACTRT && p.act) ? act_bwd2(av[j], p.act) : dgelu_poly2(av[j]));
                  }
                }
              }
              uint32_t o[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float lo, hi;
                upk(v[j], lo, hi);
                o[j] = pack_bf16(lo, hi);
              }
              if (ln_produce) {         // statistics of exactly the bf16 values the next LayerNorm would read
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const f32x2 xv = bf16x2_to_f32x2(o[j]);
                  ln_sum2 = add2(ln_sum2, xv);
                  ln_sq2 = fma2(xv, xv, ln_sq2);
                }
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3])
                           : "memory");
            }
          }
        }
        if (rc + round_cols >= kColsPerWarp) {  // last TMEM read of this tile: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CTA2) mbar_arrive_leader(&tempty_bar[as]); else mbar_arrive(&tempty_bar[as]);
          }
        }
        if (dual) {
          // the staging tile holds the bf16 pre-activation (GELU) / un-scaled branch output (residual): store it to
          // out2, then turn exactly those bf16 values (what the backward will see) into the final output in place
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (col0 < p.N && row0 < p.M) {
              tma_store_2d(&tmap_out2, stg, col0, row0);
              tma_store_commit();
            }
            tma_store_wait_read<0>();
          }
          __syncwarp();
          const bool row_ok = row0 + lane < p.M;
#pragma unroll 1
          for (int j = 0; j < RC / 8; ++j) {
            const uint32_t addr = stg_row + ((static_cast<uint32_t>(j) ^ sw) << 4);
            uint4 q;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(addr));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
            uint32_t o[4];
            if (epi == B200_EPI_BIAS_GELU) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float lo, hi;
                const f32x2 xv = bf16x2_to_f32x2(w[i]);
                upk((EPI < 0 && ACTRT && p.act) ? act_fwd2(xv, p.act) : gelu_erf2(xv), lo, hi);
                o[i] = pack_bf16(lo, hi);
              }
            } else {
              // residual epilogue: out = aux + row_scale * gamma * branch.  The residual row segment comes straight
              // from global memory here (16 bytes per lane and step, every 32-byte sector fully used over two steps).
              const int col = col0 + j * 8;
              uint4 a4 = make_uint4(0u, 0u, 0u, 0u);
              f32x2 gv[4] = {rs2, rs2, rs2, rs2};
              if (col < p.N) {
                if (row_ok) a4 = __ldg(reinterpret_cast<const uint4*>(p.aux + static_cast<long long>(row0 + lane) * p.ldaux + col));
                if (p.gamma != nullptr) {
                  load8_param2(p.gamma, p.gamma_f32, col, gv);
#pragma unroll
                  for (int i = 0; i < 4; ++i) gv[i] = mul2(gv[i], rs2);
                }
              }
              const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float lo, hi;
                upk(fma2(bf16x2_to_f32x2(w[i]), gv[i], bf16x2_to_f32x2(aw[i])), lo, hi);
                o[i] = pack_bf16(lo, hi);
              }
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
          }
        }
        fence_proxy_async_smem();  // staging writes (generic proxy) -> visible to the TMA store (async proxy)
        __syncwarp();
        if constexpr (RC == 64) {
          // bias gradient of the Linear whose output gradient this tile is (db = column sums of dY, here of exactly the
          // bf16 values being stored): lane l sums bf16 columns 2l, 2l+1 of the warp's 32 staged rows (128-byte rows,
          // 16-byte chunks XOR-swizzled by row & 7: the 32 lanes of a row read hit 32 different banks) and adds the pair
          // to colsum with one vector reduction.  Rows past M are zero (zero-filled operands), columns are checked.
          if (p.colsum != nullptr && !out_f32) {
            const uint32_t base = smem_u32(stg) + static_cast<uint32_t>(lane & 3) * 4u;
            const uint32_t chunk = static_cast<uint32_t>(lane >> 2);
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              uint32_t w;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(base + r * 128u + ((chunk ^ static_cast<uint32_t>(r & 7)) << 4)));
              s0 += bf16_lo(w);
              s1 += bf16_hi(w);
            }
            const int col = col0 + 2 * lane;
            if (col < p.N && row0 < p.M)
              asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p.colsum + col), "f"(s0), "f"(s1) : "memory");
          }
        }
        if (lane == 0 && col0 < p.N && row0 < p.M) {
          if (p.patch) {
            // tokens of image `img`: rows [ti * pyc * gw + quad * 32 ...) + the prefix-token offset; a tile holds
            // pe_tile_rows patch rows, so the last (partial) 32-row group goes through the short-box map and the groups
            // above it are not stored at all; rows past the image's last token are clipped by the 3-D map
            const int img = m_blk / p.pe_ntiles, ti = m_blk - img * p.pe_ntiles;
            const int valid = p.pe_tile_rows - quad * 32;
            const int tok = ti * p.pe_pyc * p.pe_gw + quad * 32 + p.pe_row_off;
            if (valid >= 32) tma_store_3d(&tmap_out, stg, col0, tok, img);
            else if (valid > 0) tma_store_3d(&tmap_out2, stg, col0, tok, img);
          } else if (reduce_out) tma_reduce_add_2d(&tmap_out, stg, col0, row0);
          else tma_store_2d(&tmap_out, stg, col0, row0);
          tma_store_commit();
        }
      }
      if (ln_produce && row0 + lane < p.M) {
        float s0, s1, q0, q1;
        upk(ln_sum2, s0, s1);
        upk(ln_sq2, q0, q1);
        reinterpret_cast<float2*>(p.ln_stats_out)[static_cast<long long>(n_blk * 4 + part) * p.M + row0 + lane] =
            make_float2(s0 + s1, q0 + q1);
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) tma_store_wait<0>();
    if (trace && ew == 0 && lane == 0) { trace[3] = first_acc; trace[4] = last_acc; trace[10] = ntiles; trace[5] = gtime(); }
  }

  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all(); else __syncthreads();   // a pair leaves together: peer smem / barriers stay valid
  if (warp_idx == 2) {
    tc_fence_after();
    if constexpr (CTA2) tmem_dealloc_pair(tmem_base, 2 * BN); else tmem_dealloc(tmem_base, 2 * BN);
  }
  if (trace && threadIdx.x == 0) trace[6] = gtime();
}

// Folded LayerNorm: the producer's [slots][M][2] partial (sum, sum of squares) -> slot `slots`: (rstd, -mean * rstd) per row.
__global__ void __launch_bounds__(256) ln_finalize_kernel(float2* __restrict__ stats, int M, int slots, float inv_k, float eps) {
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= M) return;
  float s = 0.f, q = 0.f;
  for (int i = 0; i < slots; ++i) {
    const float2 v = stats[static_cast<long long>(i) * M + r];
    s += v.x;
    q += v.y;
  }
  const float mean = s * inv_k;
  const float rstd = rsqrtf(fmaxf(fmaf(q, inv_k, -mean * mean), 0.f) + eps);
  stats[static_cast<long long>(slots) * M + r] = make_float2(rstd, -mean * rstd);
}

static unsigned long long* g_gemm_trace = nullptr;

template <int BN, int STAGES, bool A_MN, bool B_MN, bool CTA2 = false, int RC = 32, int EPI = -1, bool ACTRT = false, int LNF = 0>
static int launch_gemm(const b200_gemm_desc& d, cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES, CTA2, RC>;
  static_assert(L::kTotal <= 227 * 1024, "shared memory budget");
  constexpr int kBRows = CTA2 ? BN / 2 : BN;   // B rows one CTA loads per k-block
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
    if (!B_MN) { dims[0] = d.K; dims[1] = d.N; box[0] = BK; box[1] = kBRows; }
    else       { dims[0] = d.N; dims[1] = d.K; box[0] = 64; box[1] = BK; }
    strides[0] = static_cast<uint64_t>(d.ldb) * 2;
    rc = encode_tmap_bf16(&tb, d.b, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  CUtensorMap tout, taux, tout2;
  {
    // out / aux: dims {N, M}; box = one 32-row staging tile with rows of 2*RC bytes (64B / 128B swizzle)
    const bool of32 = d.out_dtype == B200_DTYPE_F32;
    constexpr int kSwz = RC == 32 ? 64 : 128;
    uint64_t dims[2] = {static_cast<uint64_t>(d.N), static_cast<uint64_t>(d.M)};
    uint64_t strides[1] = {static_cast<uint64_t>(d.ldo) * (of32 ? 4 : 2)};
    uint32_t box[2] = {of32 ? static_cast<uint32_t>(RC / 2) : static_cast<uint32_t>(RC), 32u};
    int rc = of32 ? encode_tmap_f32(&tout, d.out, 2, dims, strides, box, kSwz)
                  : encode_tmap_bf16(&tout, d.out, 2, dims, strides, box, kSwz);
    if (rc) return rc;
    if (d.aux != nullptr) {
      uint64_t astr[1] = {static_cast<uint64_t>(d.ldaux) * 2};
      uint32_t abox[2] = {static_cast<uint32_t>(RC), 32u};
      rc = encode_tmap_bf16(&taux, d.aux, 2, dims, astr, abox, kSwz);
      if (rc) return rc;
    } else {
      taux = tout;
    }
    if (d.out2 != nullptr) {
      uint64_t ostr[1] = {static_cast<uint64_t>(d.ldo2) * 2};
      uint32_t obox[2] = {static_cast<uint32_t>(RC), 32u};
      rc = encode_tmap_bf16(&tout2, d.out2, 2, dims, ostr, obox, kSwz);
      if (rc) return rc;
    } else {
      tout2 = tout;
    }
  }
  GemmKernelParams p;
  p.out = d.out; p.out2 = d.out2; p.bias = d.bias;
  p.row_scale = d.row_scale; p.rows_per_scale = d.rows_per_scale > 0 ? d.rows_per_scale : 1; p.aux = reinterpret_cast<const __nv_bfloat16*>(d.aux); p.gamma = d.gamma;
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
  p.rope = d.rope; p.rope_tokens = d.rope_tokens; p.rope_prefix = d.rope_prefix; p.rope_cols = d.rope_cols;
  p.act = d.act;
  p.colsum = d.colsum;
  p.ln_stats_out = d.ln_stats_out; p.ln_stats = d.ln_stats; p.ln_c1 = d.ln_c1; p.ln_slots = d.ln_slots; p.ln_eps = d.ln_eps;
  p.ln_inv_k = 1.0f / static_cast<float>(d.K);
  p.patch = 0;
  p.trace = g_gemm_trace;

  auto kern = gemm_bf16_kernel<BN, STAGES, A_MN, B_MN, CTA2, RC, EPI, ACTRT, LNF>;
  static bool attr_set = false;  // idempotent; a benign race sets the same value twice
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  if constexpr (CTA2) {
    const int num_tiles = ((p.num_m_blocks + 1) / 2) * p.num_n_blocks * p.split_k;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sm_count() & ~1, 1, 1);
    cfg.blockDim = dim3(kNumThreads, 1, 1);
    cfg.dynamicSmemBytes = L::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // persistent grid = the number of CTA pairs the device can hold at once (74 on a full B200: one per TPC)
    static int max_clusters = 0;
    if (max_clusters == 0) {
      int n = 0;
      B200_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
      if (n <= 0) return fail(B200_ERR_CUDA, "b200_gemm: the device cannot host a CTA pair of the GEMM kernel");
      max_clusters = n;
    }
    const int clusters = num_tiles < max_clusters ? num_tiles : max_clusters;
    B200_CUDA_OK(launch_pdl(kern, dim3(2 * clusters, 1, 1), dim3(kNumThreads, 1, 1), L::kTotal, stream, 2, ta, tb, tout, taux,
                            tout2, p));
    return B200_OK;
  } else {
    const int num_tiles = p.num_m_blocks * p.num_n_blocks * p.split_k;
    int grid = sm_count();
    if (grid > num_tiles) grid = num_tiles;
    B200_CUDA_OK(launch_pdl(kern, dim3(grid, 1, 1), dim3(kNumThreads, 1, 1), L::kTotal, stream, 1, ta, tb, tout, taux, tout2, p));
    return B200_OK;
  }
}

// PatchEmbed.forward without a patch matrix: out[b, row_off + g, :] = patch(b, g) . W^T + bias (+ pos[pos_off + g, :]).
static int launch_patch_embed(const void* img, const void* weight, const void* bias, int bias_dtype, const void* pos,
                              int pos_rows, int pos_off, void* out, int B, int C, int H, int W, int ps, int D, int out_tokens,
                              int row_off, cudaStream_t stream) {
  constexpr int BN = 256, STAGES = 6, RC = 32;
  using L = GemmSmem<BN, STAGES, true, RC>;
  const int gw = W / ps, gh = H / ps, K = C * ps * ps;
  int pyc = 128 / gw;
  if (pyc > gh) pyc = gh;
  const int ntiles = (gh + pyc - 1) / pyc;
  pyc = (gh + ntiles - 1) / ntiles;                    // balance the tiles of an image (224 / 16: 7 + 7 grid rows)
  const int tile_rows = pyc * gw;
  CUtensorMap ta, tb, tout, tout2, taux;
  {
    // image [B, C, H, W] as {kx % 8 (stride 1), px: gw (stride ps), R = (b, c, py): B*C*gh (stride ps*W), kx / 8: ps/8 (stride 8),
    // ky: ps (stride W)}; box {8, gw, pyc, ps/8, 64/ps}, no swizzle: in shared memory 8 elements (16 B) of a patch row, then
    // the tile's patch rows at 16-byte pitch, then the eight K chunks of the k-block
    uint64_t dims[5] = {8, static_cast<uint64_t>(gw), static_cast<uint64_t>(B) * C * gh, static_cast<uint64_t>(ps / 8),
                        static_cast<uint64_t>(ps)};
    uint64_t strides[4] = {static_cast<uint64_t>(ps) * 2, static_cast<uint64_t>(ps) * W * 2, 16, static_cast<uint64_t>(W) * 2};
    uint32_t box[5] = {8, static_cast<uint32_t>(gw), static_cast<uint32_t>(pyc), static_cast<uint32_t>(ps / 8),
                       static_cast<uint32_t>(64 / ps)};
    int rc = encode_tmap_bf16(&ta, img, 5, dims, strides, box, 0);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(D)}, strides[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box[2] = {BK, BN / 2};
    int rc = encode_tmap_bf16(&tb, weight, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(D), static_cast<uint64_t>(out_tokens), static_cast<uint64_t>(B)};
    uint64_t strides[2] = {static_cast<uint64_t>(D) * 2, static_cast<uint64_t>(out_tokens) * D * 2};
    uint32_t box[3] = {RC, 32u, 1u};
    int rc = encode_tmap_bf16(&tout, out, 3, dims, strides, box, 64);
    if (rc) return rc;
    box[1] = tile_rows % 32 ? static_cast<uint32_t>(tile_rows % 32) : 32u;
    rc = encode_tmap_bf16(&tout2, out, 3, dims, strides, box, 64);
    if (rc) return rc;
  }
  if (pos != nullptr) {
    uint64_t dims[2] = {static_cast<uint64_t>(D), static_cast<uint64_t>(pos_rows)}, strides[1] = {static_cast<uint64_t>(D) * 2};
    uint32_t box[2] = {RC, 32u};
    int rc = encode_tmap_bf16(&taux, pos, 2, dims, strides, box, 64);
    if (rc) return rc;
  } else {
    taux = tout;
  }
  GemmKernelParams p = {};
  p.out = out; p.bias = bias; p.bias_f32 = bias_dtype == B200_DTYPE_F32;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(pos);
  p.rows_per_scale = 1;
  p.num_m_blocks = B * ntiles;
  p.M = p.num_m_blocks * BM; p.N = D; p.K = K;
  p.num_n_blocks = (D + BN - 1) / BN;
  p.num_k_blocks = K / BK;
  p.split_k = 1; p.k_blocks_per_split = p.num_k_blocks;
  p.epilogue = pos != nullptr ? B200_EPI_BIAS_RESIDUAL : B200_EPI_BIAS;
  p.patch = 1; p.pe_ntiles = ntiles; p.pe_pyc = pyc; p.pe_gw = gw; p.pe_gh = gh; p.pe_C = C;
  p.pe_kb_per_c = ps * ps / BK; p.pe_ky_per_kb = BK / ps; p.pe_tile_rows = tile_rows; p.pe_row_off = row_off;
  p.pe_pos_off = pos_off; p.pe_stage_tx = 2 * (tile_rows * BK * 2 + L::kBBytes); p.pe_lbo = tile_rows * 16;
  p.trace = nullptr;
  auto kern = gemm_bf16_kernel<BN, STAGES, false, false, true, RC, -1, false, 0>;
  B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
  const int num_tiles = ((p.num_m_blocks + 1) / 2) * p.num_n_blocks;
  int clusters = (sm_count() & ~1) / 2;
  if (clusters > num_tiles) clusters = num_tiles;
  B200_CUDA_OK(launch_pdl(kern, dim3(2 * clusters, 1, 1), dim3(kNumThreads, 1, 1), L::kTotal, stream, 2, ta, tb, tout, taux, tout2, p));
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_patch_embed_fwd(const void* img, const void* weight, const void* bias, int32_t bias_dtype, const void* pos,
                                    int32_t pos_rows, int32_t pos_off, void* out, int32_t B, int32_t C, int32_t H, int32_t W,
                                    int32_t patch, int32_t D, int32_t out_tokens, int32_t row_off, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(img && weight && out && B > 0 && C > 0 && D > 0, "b200_patch_embed_fwd: null pointer or empty shape");
  if (patch < 8 || 64 % patch != 0 || H % patch != 0 || W % patch != 0 || W / patch > 128 || D % 8 != 0 || W % 8 != 0 ||
      (C * patch * patch) % 64 != 0)
    return fail(B200_ERR_UNSUPPORTED, "b200_patch_embed_fwd: needs patch in {8, 16, 32, 64}, H, W multiples of it, W / patch <= 128, "
                                      "D %% 8 == 0 (got patch=%d H=%d W=%d D=%d)", patch, H, W, D);
  B200_CHECK_ARG(out_tokens >= row_off + (H / patch) * (W / patch) && row_off >= 0, "b200_patch_embed_fwd: out_tokens too small");
  B200_CHECK_ARG(pos == nullptr || (pos_off >= 0 && pos_rows >= pos_off + (H / patch) * (W / patch)), "b200_patch_embed_fwd: position table too short");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(img) & 15) == 0 && (reinterpret_cast<uintptr_t>(weight) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(pos) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
                 "b200_patch_embed_fwd: pointers must be 16-byte aligned");
  return launch_patch_embed(img, weight, bias, bias_dtype, pos, pos_rows, pos_off, out, B, C, H, W, patch, D, out_tokens, row_off, stream);
}

namespace b200 { extern long long* g_attn_fwd_trace; extern long long* g_attn_bwd_trace; }

extern "C" void b200_set_trace(int32_t kernel, void* device_buf) {
  if (kernel == 0) g_gemm_trace = reinterpret_cast<unsigned long long*>(device_buf);
  else if (kernel == 1) b200::g_attn_fwd_trace = reinterpret_cast<long long*>(device_buf);
  else if (kernel == 2) b200::g_attn_bwd_trace = reinterpret_cast<long long*>(device_buf);
}

// Slots of row statistics a b200_gemm launch with ln_stats_out writes for an [M, N] output: 4 per n-block of the tile width
// the dispatcher below picks (256 columns when N > 128, else 128).
extern "C" int32_t b200_gemm_ln_slots(int32_t M, int32_t N) {
  (void)M;
  const int bn = N > 128 ? 256 : 128;
  return 4 * ((N + bn - 1) / bn);
}

extern "C" int b200_gemm(const b200_gemm_desc* dp, void* stream_) {
  if (dp == nullptr) return fail(B200_ERR_INVALID_ARG, "b200_gemm: null descriptor");
  const b200_gemm_desc& d = *dp;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(d.a && d.b && d.out, "b200_gemm: null operand pointer");
  B200_CHECK_ARG(d.M > 0 && d.N > 0 && d.K > 0, "b200_gemm: non-positive shape M=%d N=%d K=%d", d.M, d.N, d.K);
  B200_CHECK_ARG(d.N % 8 == 0, "b200_gemm: N=%d must be a multiple of 8", d.N);
  B200_CHECK_ARG(d.ldo % (d.out_dtype == B200_DTYPE_F32 ? 4 : 8) == 0 && d.ldo >= d.N,
                 "b200_gemm: ldo=%lld must be >= N and give 16-byte aligned rows", (long long)d.ldo);
  if (d.out_dtype == B200_DTYPE_F32)
    B200_CHECK_ARG(d.epilogue == B200_EPI_BIAS, "b200_gemm: fp32 outputs support only the bias epilogue");
  B200_CHECK_ARG(d.lda % 8 == 0 && d.ldb % 8 == 0, "b200_gemm: lda/ldb must be multiples of 8 (16-byte TMA strides)");
  B200_CHECK_ARG(d.lda >= (d.a_mn ? d.M : d.K), "b200_gemm: lda=%lld too small", (long long)d.lda);
  B200_CHECK_ARG(d.ldb >= (d.b_mn ? d.N : d.K), "b200_gemm: ldb=%lld too small", (long long)d.ldb);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(d.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.b) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(d.out) & 15) == 0,
                 "b200_gemm: operand pointers must be 16-byte aligned");
  B200_CHECK_ARG(d.epilogue >= B200_EPI_BIAS && d.epilogue <= B200_EPI_BIAS_ROPE, "b200_gemm: bad epilogue %d", d.epilogue);
  B200_CHECK_ARG(d.act >= 0 && d.act <= 2, "b200_gemm: bad activation %d (0 erf GELU, 1 tanh GELU, 2 SiLU)", d.act);
  if (d.epilogue == B200_EPI_BIAS_ROPE)
    B200_CHECK_ARG(d.rope != nullptr && (reinterpret_cast<uintptr_t>(d.rope) & 15) == 0 && d.rope_tokens > 0 &&
                       d.rope_prefix >= 0 && d.rope_prefix < d.rope_tokens && d.rope_cols % 64 == 0 && d.rope_cols <= d.N &&
                       d.out_dtype == B200_DTYPE_BF16 && d.out2 == nullptr && !d.a_mn && !d.b_mn,
                   "b200_gemm: EPI_BIAS_ROPE needs a 16-byte aligned rope table, rope_prefix < rope_tokens, rope_cols a "
                   "multiple of 64 (head_dim 64), bf16 output and K-major operands");
  if (d.epilogue == B200_EPI_BIAS_RESIDUAL || d.epilogue == B200_EPI_DGELU) {
    B200_CHECK_ARG(d.aux != nullptr && d.ldaux % 8 == 0 && d.ldaux >= d.N &&
                       (reinterpret_cast<uintptr_t>(d.aux) & 15) == 0,
                   "b200_gemm: epilogue %d needs a 16-byte aligned aux with ldaux %% 8 == 0", d.epilogue);
  }
  if (d.out2 != nullptr)
    B200_CHECK_ARG((d.epilogue == B200_EPI_BIAS_GELU || d.epilogue == B200_EPI_BIAS_RESIDUAL) &&
                       d.out_dtype == B200_DTYPE_BF16 && d.ldo2 % 8 == 0 && d.ldo2 >= d.N &&
                       (reinterpret_cast<uintptr_t>(d.out2) & 15) == 0,
                   "b200_gemm: out2 needs the GELU or residual epilogue, bf16 output and 16-byte aligned rows");
  if (d.colsum != nullptr)
    B200_CHECK_ARG(d.epilogue == B200_EPI_DGELU && d.out_dtype == B200_DTYPE_BF16 && d.M >= 256 && d.N > 128 && !d.a_mn &&
                       (reinterpret_cast<uintptr_t>(d.colsum) & 7) == 0,
                   "b200_gemm: colsum rides in the EPI_DGELU epilogue of the 64-column-round kernels only (bf16 output, "
                   "M >= 256, N > 128, 8-byte aligned fp32 [N])");
  if (d.ln_stats_out != nullptr)
    B200_CHECK_ARG((d.epilogue == B200_EPI_BIAS || d.epilogue == B200_EPI_BIAS_RESIDUAL) && d.out_dtype == B200_DTYPE_BF16 &&
                       d.out2 == nullptr && d.split_k <= 1 && (reinterpret_cast<uintptr_t>(d.ln_stats_out) & 7) == 0,
                   "b200_gemm: ln_stats_out needs the bias / residual epilogue, one bf16 output and an 8-byte aligned buffer");
  if (d.ln_stats != nullptr)
    B200_CHECK_ARG((d.epilogue == B200_EPI_BIAS || d.epilogue == B200_EPI_BIAS_GELU) && d.out_dtype == B200_DTYPE_BF16 &&
                       d.out2 == nullptr && !d.a_mn && d.ln_c1 != nullptr && d.bias != nullptr && d.bias_dtype == B200_DTYPE_F32 &&
                       d.ln_slots > 0 && (reinterpret_cast<uintptr_t>(d.ln_stats) & 7) == 0 &&
                       (reinterpret_cast<uintptr_t>(d.ln_c1) & 15) == 0,
                   "b200_gemm: ln_stats needs the bias / GELU epilogue, a bf16 output, fp32 ln_c1 and fp32 bias (c2), ln_slots > 0");
  if (d.row_scale != nullptr)
    B200_CHECK_ARG(d.epilogue == B200_EPI_BIAS_RESIDUAL && d.rows_per_scale > 0,
                   "b200_gemm: row_scale needs the residual epilogue and rows_per_scale > 0");
  if (d.bias) B200_CHECK_ARG((reinterpret_cast<uintptr_t>(d.bias) & 15) == 0, "b200_gemm: bias must be 16-byte aligned");
  if (d.gamma) B200_CHECK_ARG((reinterpret_cast<uintptr_t>(d.gamma) & 15) == 0, "b200_gemm: gamma must be 16-byte aligned");
  if (d.accumulate) B200_CHECK_ARG(d.out_dtype == B200_DTYPE_F32, "b200_gemm: accumulate needs an fp32 output");
  if (d.split_k > 1)
    B200_CHECK_ARG(d.out_dtype == B200_DTYPE_F32 && d.epilogue == B200_EPI_BIAS && d.bias == nullptr,
                   "b200_gemm: split_k needs fp32 output, EPI_BIAS and no bias");

  // Tile-N selection: 256 wide unless N is small (wasted columns) -- 128 then.
  const bool wide = d.N > 128;
  if (d.ln_stats_out != nullptr || d.ln_stats != nullptr) {
    // folded LayerNorm: the K-major pair kernels of the inference forward (callers fall back to a LayerNorm pass otherwise)
    if (!(wide && d.M >= 256 && !d.a_mn && !d.b_mn && (d.act == 0 || (d.ln_stats != nullptr && d.epilogue == B200_EPI_BIAS_GELU))))
      return fail(B200_ERR_UNSUPPORTED, "b200_gemm: ln_stats / ln_stats_out need M >= 256, N > 128 and K-major operands");
    if (d.ln_stats != nullptr && d.ln_stats_out != nullptr)
      return fail(B200_ERR_UNSUPPORTED, "b200_gemm: one launch is either the producer or the consumer of a folded LayerNorm");
    if (d.ln_stats_out != nullptr) return launch_gemm<256, 6, false, false, true, 32, -1, false, 2>(d, stream);
    ln_finalize_kernel<<<(d.M + 255) / 256, 256, 0, stream>>>(reinterpret_cast<float2*>(d.ln_stats), d.M, d.ln_slots,
                                                             1.0f / static_cast<float>(d.K), d.ln_eps);
    B200_CUDA_OK(cudaGetLastError());
    if (d.epilogue == B200_EPI_BIAS_GELU && d.act != 0)      // tanh GELU / SiLU: the activation-at-run-time kernel as consumer
      return launch_gemm<256, 5, false, false, true, 64, -1, true, 1>(d, stream);
    if (d.epilogue == B200_EPI_BIAS_GELU) return launch_gemm<256, 5, false, false, true, 64, B200_EPI_BIAS_GELU, false, 1>(d, stream);
    return launch_gemm<256, 6, false, false, true, 32, -1, false, 1>(d, stream);
  }
  // CTA pairs (256 x 256 tiles, cta_group::2) once there are enough 256-row tile pairs to fill the GPU
  if (d.act != 0) {
    // tanh GELU / SiLU: the activation-at-run-time instantiations of the two shapes an MLP produces -- fc1 forward
    // (K-major x K-major, EPI_BIAS_GELU) and the fc2 dgrad (K-major x MN-major, EPI_DGELU)
    B200_CHECK_ARG((d.epilogue == B200_EPI_BIAS_GELU || d.epilogue == B200_EPI_DGELU) && !d.a_mn && d.out_dtype == B200_DTYPE_BF16,
                   "b200_gemm: act != 0 applies to the GELU / DGELU epilogues of K-major activations");
    if (wide && d.M >= 256)
      return d.b_mn ? launch_gemm<256, 5, false, true, true, 64, -1, true>(d, stream)
                    : launch_gemm<256, 5, false, false, true, 64, -1, true>(d, stream);
    if (wide)
      return d.b_mn ? launch_gemm<256, 4, false, true, false, 32, -1, true>(d, stream)
                    : launch_gemm<256, 4, false, false, false, 32, -1, true>(d, stream);
    return d.b_mn ? launch_gemm<128, 6, false, true, false, 32, -1, true>(d, stream)
                  : launch_gemm<128, 6, false, false, false, 32, -1, true>(d, stream);
  }
  static const int pair_mode = [] { const char* e = getenv("B200_GEMM_2CTA"); return e ? atoi(e) : 1; }();
  // B200_GEMM_2CTA: 0 = single CTAs; 2 = pairs, 32-column epilogue rounds, 6-stage ring; 3 = pairs, 64-column rounds,
  //                 5-stage ring; 1 (default) = pairs, round width chosen per epilogue
  // default (1): the transcendental epilogues (GELU, GELU') take the 64-column rounds, whose single TMA store per tile
  // keeps 16 busy epilogue warps off the store-completion wait; bias / residual epilogues keep the sixth ring stage
  const bool heavy_epi = d.epilogue == B200_EPI_BIAS_GELU || d.epilogue == B200_EPI_DGELU;
  if (wide && (pair_mode == 2 || (pair_mode == 1 && !heavy_epi)) && d.M >= 256) {
    if (!d.a_mn && !d.b_mn) return launch_gemm<256, 6, false, false, true>(d, stream);
    if (!d.a_mn && d.b_mn) return launch_gemm<256, 6, false, true, true>(d, stream);
    if (d.a_mn && d.b_mn) return launch_gemm<256, 6, true, true, true>(d, stream);
    return launch_gemm<256, 6, true, false, true>(d, stream);
  }
  if (wide && pair_mode && d.M >= 256) {
    // the two shapes the model produces get epilogue-specialised kernels: fc1 forward (K-major x K-major, GELU) and
    // the fc2 dgrad that applies GELU' (K-major x MN-major)
    if (d.out_dtype == B200_DTYPE_BF16 && !d.a_mn && !d.b_mn && d.epilogue == B200_EPI_BIAS_GELU)
      return launch_gemm<256, 5, false, false, true, 64, B200_EPI_BIAS_GELU>(d, stream);
    if (d.out_dtype == B200_DTYPE_BF16 && !d.a_mn && d.b_mn && d.epilogue == B200_EPI_DGELU)
      return launch_gemm<256, 5, false, true, true, 64, B200_EPI_DGELU>(d, stream);
    if (!d.a_mn && !d.b_mn) return launch_gemm<256, 5, false, false, true, 64>(d, stream);
    if (!d.a_mn && d.b_mn) return launch_gemm<256, 5, false, true, true, 64>(d, stream);
    if (d.a_mn && d.b_mn) return launch_gemm<256, 5, true, true, true, 64>(d, stream);
    return launch_gemm<256, 5, true, false, true, 64>(d, stream);
  }
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
