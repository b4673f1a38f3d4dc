// Data-parallel gradient exchange fused into the kernel that finalises a block's gradients (the ONE
// multi-GPU exchange of the path -- the all-reduce(mean) of the parameter gradients that timm gets from
// DistributedDataParallel, timm/task/classification.py:64-65).
//
// The backward of a transformer block leaves the block's parameter gradients in one fp32 workspace (wgrad split-K sums,
// bias / LayerNorm / LayerScale column sums).  Instead of converting that workspace to bf16 locally and handing the
// result to a separate NCCL all-reduce kernel, this kernel reads the workspace ONCE, scales by 1 / world_size, rounds to
// bf16 and issues `multimem.red.add` on the NVLink-multicast address of the flat gradient buffer: the NVSwitch replicates
// every 16-byte packet to all ranks of the node and each rank's memory controller adds it into its own copy.  When every
// rank has done so for all of its blocks, every rank holds the mean gradient -- there is no all-reduce kernel, no SM
// reserved for communication (an NCCL kernel occupies SMs our persistent GEMMs then cannot start CTAs on) and the NVLink
// traffic of block k rides under the backward GEMMs of block k-1.  Egress per rank and step = the gradient bytes once
// (173 MB for ViT-B in bf16, ~0.2 ms of NVLink5 time spread over a 20 ms backward).
//
// Ordering contract (host side, b200timm/grad_sync.py): all ranks zero their buffers -> peer barrier -> pushes ->
// peer barrier -> optimizer.  Each thread ends with a system-scope fence so that its reductions are performed before the
// grid completes (the barrier kernel that follows publishes with release semantics).
#include "common.h"
#include "ptx.cuh"

namespace b200 {

template <bool SRC_F32>
__global__ void __launch_bounds__(256) grad_push_kernel(const void* __restrict__ src, __nv_bfloat16* dst_mc, long long nvec,
                                                        float scale) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    uint32_t o[4];
    if constexpr (SRC_F32) {
      const float4* p = reinterpret_cast<const float4*>(src) + 2 * i;
      const float4 a = __ldg(p), b = __ldg(p + 1);
      o[0] = pack_bf16(a.x * scale, a.y * scale);
      o[1] = pack_bf16(a.z * scale, a.w * scale);
      o[2] = pack_bf16(b.x * scale, b.y * scale);
      o[3] = pack_bf16(b.z * scale, b.w * scale);
    } else {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(src) + i);
      o[0] = pack_bf16(bf16_lo(q.x) * scale, bf16_hi(q.x) * scale);
      o[1] = pack_bf16(bf16_lo(q.y) * scale, bf16_hi(q.y) * scale);
      o[2] = pack_bf16(bf16_lo(q.z) * scale, bf16_hi(q.z) * scale);
      o[3] = pack_bf16(bf16_lo(q.w) * scale, bf16_hi(q.w) * scale);
    }
    asm volatile("multimem.red.relaxed.sys.global.add.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst_mc + 8 * i), "r"(o[0]),
                 "r"(o[1]), "r"(o[2]), "r"(o[3])
                 : "memory");
  }
  asm volatile("fence.acq_rel.sys;" ::: "memory");
}

// Replica consistency.  multimem.red.add applies the ranks' contributions to each copy in ARRIVAL order, and bf16 addition
// is not associative: with more than two ranks the copies can differ in the last bit.  Data-parallel replicas must stay
// bit-identical (DDP's all-reduce guarantees it), so after the pushes every rank re-publishes ITS 1 / world_size slice of
// the buffer to all ranks with multimem.st: slice r of every copy becomes rank r's sum.  16 bytes per thread and step.
__global__ void __launch_bounds__(256) grad_publish_kernel(const uint4* __restrict__ src_local, uint4* dst_mc, long long nvec) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 v = src_local[i];
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_mc + i), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
  }
  asm volatile("fence.acq_rel.sys;" ::: "memory");
}

}  // namespace b200

using namespace b200;

extern "C" int b200_grad_publish(const void* src_local, void* dst_multicast, int64_t bytes, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(src_local != nullptr && dst_multicast != nullptr && bytes > 0, "b200_grad_publish: null pointer or empty range");
  B200_CHECK_ARG(bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(src_local) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(dst_multicast) & 15) == 0,
                 "b200_grad_publish: the range must be 16-byte aligned and a multiple of 16 bytes");
  const long long nvec = bytes / 16;
  long long blocks = (nvec + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 4;
  if (blocks > cap) blocks = cap;
  grad_publish_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(reinterpret_cast<const uint4*>(src_local),
                                                                      reinterpret_cast<uint4*>(dst_multicast), nvec);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_grad_push(const void* src, int32_t src_dtype, void* dst_multicast, int64_t n, float scale, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(src != nullptr && dst_multicast != nullptr && n > 0, "b200_grad_push: null pointer or empty range");
  B200_CHECK_ARG(src_dtype == B200_DTYPE_F32 || src_dtype == B200_DTYPE_BF16, "b200_grad_push: source must be f32 or bf16");
  B200_CHECK_ARG(n % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst_multicast) & 15) == 0,
                 "b200_grad_push: n must be a multiple of 8 and both pointers 16-byte aligned");
  const long long nvec = n / 8;
  long long blocks = (nvec + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 4;      // a few CTAs per SM: the kernel is NVLink-bound
  if (blocks > cap) blocks = cap;
  if (src_dtype == B200_DTYPE_F32)
    grad_push_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(dst_multicast), nvec, scale);
  else
    grad_push_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(dst_multicast), nvec, scale);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}
