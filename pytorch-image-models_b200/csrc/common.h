// Host-side helpers shared by every translation unit of libb200timm.so: error reporting, the
// driver entry point for cuTensorMapEncodeTiled (resolved at run time so the library links against
// cudart only), device-property cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200timm.h"

namespace b200 {

// Records a message retrievable through b200_last_error() and returns `code`.
int fail(int code, const char* fmt, ...);

#define B200_CHECK_ARG(cond, ...)                                   \
  do {                                                              \
    if (!(cond)) return ::b200::fail(B200_ERR_INVALID_ARG, __VA_ARGS__); \
  } while (0)

#define B200_CUDA_OK(expr)                                                                        \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return ::b200::fail(B200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                          __FILE__, __LINE__);                                                    \
  } while (0)

// Additive attention bias (resolve_self_attn_mask of timm/layers/attention.py:21-40: a causal triangle, a bool keep-mask as
// 0 / -inf, or a float mask): fp32 [batch or 1, heads or 1, N, ld], ld >= N a multiple of 4; bs / hs = batch / head strides
// in elements (0 = broadcast).  ptr == nullptr: no bias.
struct AttnBias {
  const float* ptr;
  long long bs, hs;
  int ld;
};

// Number of SMs of the current device (cached per device).
int sm_count();

// B200_PDL=1 switches programmatic dependent launch on (off by default, see pdl_enabled() in common.cu).
bool pdl_enabled();

// Launch `kern` with programmatic stream serialization (see griddep_wait in ptx.cuh) and, optionally, clusters of
// `cluster` CTAs.  The kernel must call griddep_wait() before its first access to global memory.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Encode a tiled tensor map over a bf16 tensor of rank `rank` (dims[0] innermost, strides in BYTES for
// dims 1..rank-1), swizzle = 0 (none) / 64 / anything else (128-byte), zero fill out of bounds.  Returns 0 or an error code.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle);
// Same for an fp32 tensor (GEMM fp32 outputs: weight gradients).
int encode_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle);

}  // namespace b200
