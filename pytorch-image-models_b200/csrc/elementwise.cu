// HBM-bound helper kernels around the GEMMs: patch gather (PatchEmbed), token assembly (cls/pos embed),
// column sums (bias gradients).  All are 16-byte vectorised, coalesced along the innermost dimension.
#include "common.h"
#include <string.h>
#include "ptx.cuh"
#include "gelu.cuh"

namespace b200 {

// patches[(b*gh + py)*gw + px][c*p*p + ky*p + kx] = img[b][c][py*p + ky][px*p + kx]
// One thread per 8 consecutive k.  When p % 8 == 0 (and the image is bf16) the 8 elements are one
// aligned 16-byte load from an image row; otherwise elements are gathered one by one.
// Replaces the im2col implied by nn.Conv2d(k=p, stride=p) at timm/layers/patch_embed.py:136.
template <bool IMG_F32>
__global__ void patchify_kernel(const void* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int C, int H,
                                int W, int p, int gh, int gw, int Kpad) {
  const int kvecs = Kpad >> 3;
  const long long total = static_cast<long long>(B) * gh * gw * kvecs;
  const int K = C * p * p;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kv = static_cast<int>(idx % kvecs);
    const long long patch = idx / kvecs;
    const int px = static_cast<int>(patch % gw);
    const int py = static_cast<int>((patch / gw) % gh);
    const int b = static_cast<int>(patch / (static_cast<long long>(gw) * gh));
    const int k0 = kv * 8;
    uint4 q;
    if (!IMG_F32 && (p & 7) == 0 && k0 + 8 <= K) {
      const int c = k0 / (p * p);
      const int rem = k0 - c * p * p;
      const int ky = rem / p, kx = rem - ky * p;
      const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(img) +
                                 ((static_cast<long long>(b) * C + c) * H + (py * p + ky)) * W + (px * p + kx);
      if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        q = __ldg(reinterpret_cast<const uint4*>(src));
      } else {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = __ldg(reinterpret_cast<const uint32_t*>(src) + i);
        q = make_uint4(w[0], w[1], w[2], w[3]);
      }
    } else {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = k0 + i;
        if (k < K) {
          const int c = k / (p * p);
          const int rem = k - c * p * p;
          const int ky = rem / p, kx = rem - ky * p;
          const long long off = ((static_cast<long long>(b) * C + c) * H + (py * p + ky)) * W + (px * p + kx);
          v[i] = IMG_F32 ? __ldg(reinterpret_cast<const float*>(img) + off)
                         : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(img)[off]);
        } else {
          v[i] = 0.f;
        }
      }
      q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]);
      q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
    }
    *reinterpret_cast<uint4*>(out + patch * Kpad + k0) = q;
  }
}

__device__ __forceinline__ void ld8(const void* base, bool f32, long long idx, float (&v)[8]) {
  if (f32) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
    float4 a = __ldg(p), b = __ldg(p + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx));
    v[0] = bf16_lo(q.x); v[1] = bf16_hi(q.x); v[2] = bf16_lo(q.y); v[3] = bf16_hi(q.y);
    v[4] = bf16_lo(q.z); v[5] = bf16_hi(q.z); v[6] = bf16_lo(q.w); v[7] = bf16_hi(q.w);
  }
}

// The same gather for uint8 images with the loader's normalisation folded in (timm/data/loader.py:137-142:
// next_input.to(float32).sub_(mean * 255).div_(std * 255)).  A uint8 pixel has 256 possible values, so the normalised
// bf16 value comes from a per-channel table lut[c][v] = bf16((float(v) - mean[c]) / std[c]) that the entry point fills on
// the host with the same IEEE fp32 subtract / divide and one round-to-nearest-even -- bit-identical to
// normalise-then-cast, and the kernel has no division (the first version divided on the device: 72 % XU-pipe
// utilisation, 1.3 TB/s; profiles/r01g_kernels.md).  The table travels in the kernel's parameter space and is staged in
// shared memory.
struct NormLut { uint16_t v[8][256]; };
__global__ void patchify_u8_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int C, int H,
                                   int W, int p, int gh, int gw, int Kpad, const __grid_constant__ NormLut lut) {
  __shared__ uint16_t tab[8 * 256];
  for (int i = threadIdx.x; i < C * 256; i += blockDim.x) tab[i] = lut.v[i >> 8][i & 255];
  __syncthreads();
  const int kvecs = Kpad >> 3;
  const long long total = static_cast<long long>(B) * gh * gw * kvecs;
  const int K = C * p * p;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kv = static_cast<int>(idx % kvecs);
    const long long patch = idx / kvecs;
    const int px = static_cast<int>(patch % gw);
    const int py = static_cast<int>((patch / gw) % gh);
    const int b = static_cast<int>(patch / (static_cast<long long>(gw) * gh));
    const int k0 = kv * 8;
    uint32_t h[8];
    bool vec = false;
    if ((p & 7) == 0 && k0 + 8 <= K) {          // 8 consecutive pixels of one image row: one 8-byte load
      const int c = k0 / (p * p);
      const int rem = k0 - c * p * p;
      const int ky = rem / p, kx = rem - ky * p;
      const uint8_t* src = img + ((static_cast<long long>(b) * C + c) * H + (py * p + ky)) * W + (px * p + kx);
      if ((reinterpret_cast<uintptr_t>(src) & 7) == 0) {
        const uint2 q = __ldg(reinterpret_cast<const uint2*>(src));
        const uint16_t* t = tab + c * 256;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          h[i] = t[(q.x >> (8 * i)) & 0xFFu];
          h[4 + i] = t[(q.y >> (8 * i)) & 0xFFu];
        }
        vec = true;
      }
    }
    if (!vec) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = k0 + i;
        if (k < K) {
          const int c = k / (p * p);
          const int rem = k - c * p * p;
          const int ky = rem / p, kx = rem - ky * p;
          const long long off = ((static_cast<long long>(b) * C + c) * H + (py * p + ky)) * W + (px * p + kx);
          h[i] = tab[c * 256 + __ldg(img + off)];
        } else {
          h[i] = 0u;
        }
      }
    }
    uint4 q;
    q.x = h[0] | (h[1] << 16); q.y = h[2] | (h[3] << 16);
    q.z = h[4] | (h[5] << 16); q.w = h[6] | (h[7] << 16);
    *reinterpret_cast<uint4*>(out + patch * Kpad + k0) = q;
  }
}

// tokens[b][j][:] = (j < n_prefix ? prefix[j] : proj[b*G + j - n_prefix]) + pos[j]          (pos_has_prefix)
// tokens[b][j][:] = (j < n_prefix ? prefix[j] : proj[b*G + j - n_prefix] + pos[j - n_prefix]) (otherwise:
// no_embed_class).  timm/models/vision_transformer.py:1038-1075 (_pos_embed: cat cls/reg tokens, add pos_embed).
__global__ void assemble_tokens_kernel(const __nv_bfloat16* __restrict__ proj, const void* __restrict__ prefix,
                                       const void* __restrict__ pos, __nv_bfloat16* __restrict__ tokens, int B, int G,
                                       int n_prefix, int D, int param_f32, int pos_has_prefix) {
  const int dvecs = D >> 3;
  const int ntok = G + n_prefix;
  const long long total = static_cast<long long>(B) * ntok * dvecs;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int dv = static_cast<int>(idx % dvecs);
    const long long t = idx / dvecs;
    const int j = static_cast<int>(t % ntok);
    const int b = static_cast<int>(t / ntok);
    float v[8], pv[8];
    if (j < n_prefix) ld8(prefix, param_f32, static_cast<long long>(j) * D + dv * 8, v);
    else ld8(proj, false, (static_cast<long long>(b) * G + (j - n_prefix)) * D + dv * 8, v);
    const int pj = pos_has_prefix ? j : j - n_prefix;
    if (pos != nullptr && pj >= 0) {
      ld8(pos, param_f32, static_cast<long long>(pj) * D + dv * 8, pv);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += pv[i];
    }
    uint4 q;
    q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]);
    q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(tokens + t * D + dv * 8) = q;
  }
}

// NaFlex token assembly (timm/models/naflexvit.py:857-960, _apply_learned_naflex_pos_embed :567-621):
// tokens[b][j] = prefix[j]                                   j <  n_prefix
//              = proj[b][i] + (i < pos_len[b] ? table[pos_off[b] + i] : 0)      i = j - n_prefix
// where table rows are the position embeddings already interpolated to each sample's (h, w) grid, row-major.
__global__ void naflex_assemble_kernel(const __nv_bfloat16* __restrict__ proj, const void* __restrict__ prefix,
                                       const __nv_bfloat16* __restrict__ table, const int* __restrict__ pos_off,
                                       const int* __restrict__ pos_len, __nv_bfloat16* __restrict__ tokens, int B,
                                       int N, int n_prefix, int D, int param_f32) {
  const int dvecs = D >> 3;
  const int ntok = N + n_prefix;
  const long long total = static_cast<long long>(B) * ntok * dvecs;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int dv = static_cast<int>(idx % dvecs);
    const long long t = idx / dvecs;
    const int j = static_cast<int>(t % ntok);
    const int b = static_cast<int>(t / ntok);
    float v[8];
    if (j < n_prefix) {
      ld8(prefix, param_f32, static_cast<long long>(j) * D + dv * 8, v);
    } else {
      const int i = j - n_prefix;
      ld8(proj, false, (static_cast<long long>(b) * N + i) * D + dv * 8, v);
      if (table != nullptr && i < pos_len[b]) {
        float pv[8];
        ld8(table, false, (static_cast<long long>(pos_off[b]) + i) * D + dv * 8, pv);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] += pv[k];
      }
    }
    uint4 q;
    q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]);
    q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(tokens + t * D + dv * 8) = q;
  }
}

// out[c] += sum over a slab of rows of x[r][c]; blockIdx.x = 256-column group, blockIdx.y = row slab.
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long rows, int cols,
              long long rows_per_slab) {
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_slab;
  const long long r1 = min(rows, r0 + rows_per_slab);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (col < cols) {
    for (long long r = r0 + warp; r < r1; r += 8) {
      uint4 q = __ldg(reinterpret_cast<const uint4*>(x + r * cols + col));
      acc[0] += bf16_lo(q.x); acc[1] += bf16_hi(q.x); acc[2] += bf16_lo(q.y); acc[3] += bf16_hi(q.y);
      acc[4] += bf16_lo(q.z); acc[5] += bf16_hi(q.z); acc[6] += bf16_lo(q.w); acc[7] += bf16_hi(q.w);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][c];
  if (blockIdx.x * 256 + c < cols) atomicAdd(out + blockIdx.x * 256 + c, s);
}

// out[r][c] = x[r][c] * col_scale[c] * row_scale[r / rps]: gradient of a residual branch w.r.t. its un-scaled output
// (LayerScale gamma, timm/layers/layer_scale.py:26-27, and the DropPath factor, drop.py:158-175).
__global__ void scale_rows_cols_kernel(const __nv_bfloat16* __restrict__ x, const void* __restrict__ cs,
                                       const float* __restrict__ rs, __nv_bfloat16* __restrict__ out, long long rows,
                                       int cols, int rps, int cs_f32) {
  const int cvecs = cols >> 3;
  const long long total = rows * cvecs;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % cvecs);
    const long long r = idx / cvecs;
    float v[8], c[8];
    ld8(x, false, r * cols + cv * 8, v);
    const float rscale = rs != nullptr ? __ldg(rs + r / rps) : 1.0f;
    if (cs != nullptr) {
      ld8(cs, cs_f32, cv * 8, c);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= c[i] * rscale;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= rscale;
    }
    *reinterpret_cast<uint4*>(out + r * cols + cv * 8) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

// out[c] += sum_r a[r][c] * b[r][c] * row_scale[r / rps]   (LayerScale gamma gradient); same blocking as colsum_kernel
__global__ void __launch_bounds__(256)
colsum_mul_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, const float* __restrict__ rs,
                  float* __restrict__ out, long long rows, int cols, long long rows_per_slab, int rps) {
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_slab;
  const long long r1 = min(rows, r0 + rows_per_slab);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (col < cols) {
    for (long long r = r0 + warp; r < r1; r += 8) {
      float av[8], bv[8];
      ld8(a, false, r * cols + col, av);
      ld8(b, false, r * cols + col, bv);
      const float s = rs != nullptr ? __ldg(rs + r / rps) : 1.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(av[i] * s, bv[i], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][c];
  if (blockIdx.x * 256 + c < cols) atomicAdd(out + blockIdx.x * 256 + c, s);
}

// dst (bf16 or fp32) (+)= src fp32  -- weight-gradient hand-off from the fp32 split-K workspace to .grad
template <bool DST_F32>
__global__ void convert_kernel(const float* __restrict__ src, void* __restrict__ dst, long long n, int accumulate) {
  const long long nvec = n >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    if (DST_F32) {
      float4* d = reinterpret_cast<float4*>(dst) + i;
      if (accumulate) { const float4 o = *d; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
      *d = v;
    } else {
      uint2* d = reinterpret_cast<uint2*>(dst) + i;
      if (accumulate) {
        const uint2 o = *d;
        v.x += bf16_lo(o.x); v.y += bf16_hi(o.x); v.z += bf16_lo(o.y); v.w += bf16_hi(o.y);
      }
      *d = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
  }
}

// Rotary embedding of the q / k thirds of qkv[B*N, 3*H*64] in place (AttentionRope, timm/layers/attention.py:249-253).
// A thread owns token row r, head slot g (q heads then k heads) and the 8-column chunks at d0 and d0 + 32 of the head:
// both layouts pair columns inside those 16 values (interleaved: neighbours; half: d <-> d + 32).
__global__ void __launch_bounds__(256)
rope_qk_kernel(__nv_bfloat16* __restrict__ qkv, const float* __restrict__ emb, long long rows, int N, int H, int prefix,
               int half_layout, int inverse, int per_sample) {
  const long long total = rows * (2LL * H) * 4;
  const long long ld = 3LL * H * 64;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int d0 = static_cast<int>(idx & 3) * 8;
    const long long t = idx >> 2;
    const int g = static_cast<int>(t % (2 * H));
    const long long r = t / (2 * H);
    const int tok = static_cast<int>(r % N) - prefix;
    if (tok < 0) continue;
    __nv_bfloat16* px = qkv + r * ld + g * 64 + d0;
    // shared table [N - prefix, 128], or one table per sample [B, N - prefix, 128] (NaFlex: rope of shape [B, 1, N, 2 hd])
    const float* e = emb + ((per_sample ? (r / N) * (N - prefix) : 0LL) + tok) * 128 + d0;
    float x[16], sn[16], cs[16], y[16];
    {
      const uint4 a = *reinterpret_cast<const uint4*>(px), b = *reinterpret_cast<const uint4*>(px + 32);
      const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) { x[2 * i] = bf16_lo(w[i]); x[2 * i + 1] = bf16_hi(w[i]); }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(e + 32 * i)), s1 = __ldg(reinterpret_cast<const float4*>(e + 32 * i + 4));
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(e + 64 + 32 * i)), c1 = __ldg(reinterpret_cast<const float4*>(e + 64 + 32 * i + 4));
        sn[8 * i + 0] = s0.x; sn[8 * i + 1] = s0.y; sn[8 * i + 2] = s0.z; sn[8 * i + 3] = s0.w;
        sn[8 * i + 4] = s1.x; sn[8 * i + 5] = s1.y; sn[8 * i + 6] = s1.z; sn[8 * i + 7] = s1.w;
        cs[8 * i + 0] = c0.x; cs[8 * i + 1] = c0.y; cs[8 * i + 2] = c0.z; cs[8 * i + 3] = c0.w;
        cs[8 * i + 4] = c1.x; cs[8 * i + 5] = c1.y; cs[8 * i + 6] = c1.z; cs[8 * i + 7] = c1.w;
      }
    }
    // forward:  y[a] = x[a] cos[a] - x[b] sin[a],  y[b] = x[b] cos[b] + x[a] sin[b]   for each pair (a, b)
    // inverse (transpose): y[a] = x[a] cos[a] + x[b] sin[b],  y[b] = x[b] cos[b] - x[a] sin[a]
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int a = half_layout ? i : (i < 4 ? 2 * i : 8 + 2 * (i - 4));
      const int b = half_layout ? i + 8 : a + 1;
      if (!inverse) {
        y[a] = fmaf(x[a], cs[a], -x[b] * sn[a]);
        y[b] = fmaf(x[b], cs[b], x[a] * sn[b]);
      } else {
        y[a] = fmaf(x[a], cs[a], x[b] * sn[b]);
        y[b] = fmaf(x[b], cs[b], -x[a] * sn[a]);
      }
    }
    *reinterpret_cast<uint4*>(px) = make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
    *reinterpret_cast<uint4*>(px + 32) =
        make_uint4(pack_bf16(y[8], y[9]), pack_bf16(y[10], y[11]), pack_bf16(y[12], y[13]), pack_bf16(y[14], y[15]));
  }
}

// Backward of naflex_assemble_kernel: dproj[b][i] = dtokens[b][n_prefix + i] (bf16 copy, the wgrad GEMM's operand),
// dprefix[j] += dtokens[b][j] and dtable[pos_off[b] + i] += dtokens[b][n_prefix + i] for i < pos_len[b] (fp32, one
// 16-byte red per 4 channels: samples sharing a grid size share table rows).  dprefix / dtable are zeroed by the caller.
__global__ void naflex_assemble_bwd_kernel(const __nv_bfloat16* __restrict__ dtokens, __nv_bfloat16* __restrict__ dproj,
                                           float* __restrict__ dprefix, float* __restrict__ dtable,
                                           const int* __restrict__ pos_off, const int* __restrict__ pos_len, int B,
                                           int N, int n_prefix, int D) {
  const int dvecs = D >> 3;
  const int ntok = N + n_prefix;
  const long long total = static_cast<long long>(B) * ntok * dvecs;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int dv = static_cast<int>(idx % dvecs);
    const long long t = idx / dvecs;
    const int j = static_cast<int>(t % ntok);
    const int b = static_cast<int>(t / ntok);
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(dtokens + t * D + dv * 8));
    float* acc = nullptr;
    if (j < n_prefix) {
      if (dprefix != nullptr) acc = dprefix + static_cast<long long>(j) * D + dv * 8;
    } else {
      const int i = j - n_prefix;
      *reinterpret_cast<uint4*>(dproj + (static_cast<long long>(b) * N + i) * D + dv * 8) = q;
      if (dtable != nullptr && i < pos_len[b]) acc = dtable + (static_cast<long long>(pos_off[b]) + i) * D + dv * 8;
    }
    if (acc != nullptr) {
      atomicAdd(reinterpret_cast<float4*>(acc), make_float4(bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y)));
      atomicAdd(reinterpret_cast<float4*>(acc) + 1, make_float4(bf16_lo(q.z), bf16_hi(q.z), bf16_lo(q.w), bf16_hi(q.w)));
    }
  }
}

// Gated MLP activations (timm/layers/mlp.py:106-108 GluMlp, :156-158 SwiGLU): h = fc1(x) is [M, 2H] with the gate in the
// first (gate_first) or second half of each row; out[m, j] = act(gate[m, j]) * value[m, j].  ACT: 0 SiLU, 1 sigmoid.
template <int ACT>
__device__ __forceinline__ float glu_act(float g) {
  const float sg = 1.f / (1.f + __expf(-g));
  return ACT == 0 ? g * sg : sg;
}
template <int ACT>
__device__ __forceinline__ float glu_dact(float g) {
  const float sg = 1.f / (1.f + __expf(-g));
  return ACT == 0 ? sg * fmaf(g, 1.f - sg, 1.f) : sg * (1.f - sg);
}

template <int ACT>
__global__ void glu_fwd_kernel(const __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ out, long long M, int H,
                               int gate_first) {
  const int hv = H >> 3;
  const long long total = M * hv;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = idx / hv;
    const int j = static_cast<int>(idx - m * hv) * 8;
    float a[8], b[8];
    ld8(h, false, m * 2 * H + j, a);
    ld8(h, false, m * 2 * H + H + j, b);
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = gate_first ? glu_act<ACT>(a[k]) * b[k] : a[k] * glu_act<ACT>(b[k]);
    uint4 q;
    q.x = pack_bf16(o[0], o[1]); q.y = pack_bf16(o[2], o[3]);
    q.z = pack_bf16(o[4], o[5]); q.w = pack_bf16(o[6], o[7]);
    *reinterpret_cast<uint4*>(out + m * H + j) = q;
  }
}

// dh[m, gate half] = dout * value * act'(gate);  dh[m, value half] = dout * act(gate)
template <int ACT>
__global__ void glu_bwd_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ dout,
                               __nv_bfloat16* __restrict__ dh, long long M, int H, int gate_first) {
  const int hv = H >> 3;
  const long long total = M * hv;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = idx / hv;
    const int j = static_cast<int>(idx - m * hv) * 8;
    float a[8], b[8], d[8], da[8], db[8];
    ld8(h, false, m * 2 * H + j, a);
    ld8(h, false, m * 2 * H + H + j, b);
    ld8(dout, false, m * H + j, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float g = gate_first ? a[k] : b[k], v = gate_first ? b[k] : a[k];
      const float dg = d[k] * v * glu_dact<ACT>(g), dv = d[k] * glu_act<ACT>(g);
      da[k] = gate_first ? dg : dv;
      db[k] = gate_first ? dv : dg;
    }
    uint4 q;
    q.x = pack_bf16(da[0], da[1]); q.y = pack_bf16(da[2], da[3]);
    q.z = pack_bf16(da[4], da[5]); q.w = pack_bf16(da[6], da[7]);
    *reinterpret_cast<uint4*>(dh + m * 2 * H + j) = q;
    q.x = pack_bf16(db[0], db[1]); q.y = pack_bf16(db[2], db[3]);
    q.z = pack_bf16(db[4], db[5]); q.w = pack_bf16(db[6], db[7]);
    *reinterpret_cast<uint4*>(dh + m * 2 * H + H + j) = q;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// The NaFlex input tail: batch_patchify (timm/models/naflexvit.py:159-201) and the NaFlexPrefetchLoader
// normalisation (timm/data/naflex_loader.py:126-161) folded into the pass that builds the patch-projection GEMM's A operand.
//   naflex_patchify_kernel : NCHW image (uint8 / bf16 / fp32) -> [B * gh * gw, Kpad] bf16 rows, flat patch layout P-P-C
//                            (channels_last, NaFlex default: k = (ky * p + kx) * C + c) or C-P-P; uint8 pixels are
//                            normalised (v - mean[c]) / std[c] in fp32 and rounded once -- the loader's sub / div and the
//                            autocast / .to(bf16) cast in one step; rows past H / W (pad=True) read as zero pixels, i.e.
//                            (0 - mean) / std for uint8 exactly as normalising a zero-padded image would.
//   patches_u8_kernel      : the loader's uint8 patches [rows, K] (already patchified by the dataset) -> normalised bf16
//                            [rows, Kpad]: 1 byte read + 2 written per value instead of the fp32 round trip.
struct Norm8 { float mean[8], std[8]; };      // per-channel statistics, by value in the kernel's parameter space

template <int SRC>   // 0 = uint8, 1 = bf16, 2 = fp32
__global__ void naflex_patchify_kernel(const void* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int C, int H, int W,
                                       int p, int gh, int gw, int Kpad, int channels_last, const __grid_constant__ Norm8 nrm) {
  const int kvecs = Kpad >> 3;
  const long long total = static_cast<long long>(B) * gh * gw * kvecs;
  const int K = C * p * p;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kv = static_cast<int>(idx % kvecs);
    const long long patch = idx / kvecs;
    const int px = static_cast<int>(patch % gw);
    const int py = static_cast<int>((patch / gw) % gh);
    const int b = static_cast<int>(patch / (static_cast<long long>(gw) * gh));
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kv * 8 + i;
      v[i] = 0.f;
      if (k < K) {
        int c, ky, kx;
        if (channels_last) { c = k % C; const int r = k / C; ky = r / p; kx = r - ky * p; }
        else { c = k / (p * p); const int r = k - c * p * p; ky = r / p; kx = r - ky * p; }
        const int y = py * p + ky, x = px * p + kx;
        float pix = 0.f;
        if (y < H && x < W) {
          const long long off = ((static_cast<long long>(b) * C + c) * H + y) * W + x;
          if (SRC == 0) pix = static_cast<float>(reinterpret_cast<const uint8_t*>(img)[off]);
          else if (SRC == 1) pix = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(img)[off]);
          else pix = __ldg(reinterpret_cast<const float*>(img) + off);
        }
        v[i] = (SRC == 0) ? __fdiv_rn(__fsub_rn(pix, nrm.mean[c]), nrm.std[c]) : pix;
      }
    }
    *reinterpret_cast<uint4*>(out + patch * Kpad + kv * 8) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

__global__ void patches_u8_kernel(const uint8_t* __restrict__ src, __nv_bfloat16* __restrict__ out, long long rows, int K, int Kpad,
                                  int C, int channels_last, const __grid_constant__ Norm8 nrm) {
  const int kvecs = Kpad >> 3;
  const long long total = rows * kvecs;
  const int pp = K / C;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kv = static_cast<int>(idx % kvecs);
    const long long row = idx / kvecs;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kv * 8 + i;
      v[i] = 0.f;
      if (k < K) {
        const int c = channels_last ? k % C : k / pp;
        v[i] = __fdiv_rn(__fsub_rn(static_cast<float>(src[row * K + k]), nrm.mean[c]), nrm.std[c]);
      }
    }
    *reinterpret_cast<uint4*>(out + row * Kpad + kv * 8) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

// dpre = dy * act'(pre): the activation backward of an Mlp whose activation output does not feed fc2 directly (an inner
// norm sits in between: Mlp(norm_layer=...), timm/layers/mlp.py:44-53 `scale_mlp_norm`) -- otherwise GELU' rides in the
// fc2-dgrad GEMM epilogue.  act: 0 exact GELU, 1 tanh GELU, 2 SiLU (gelu.cuh).
__global__ void act_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ pre,
                               __nv_bfloat16* __restrict__ dpre, long long nvec, int act) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(dy) + i), x = __ldg(reinterpret_cast<const uint4*>(pre) + i);
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, xw[4] = {x.x, x.y, x.z, x.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float lo, hi;
      upk(mul2(bf16x2_to_f32x2(gw[k]), act_bwd2(bf16x2_to_f32x2(xw[k]), act)), lo, hi);
      o[k] = pack_bf16(lo, hi);
    }
    reinterpret_cast<uint4*>(dpre)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace b200

using namespace b200;

static Norm8 make_norm8(const float* mean, const float* std, int C) {
  Norm8 n;
  for (int c = 0; c < 8; ++c) { n.mean[c] = (mean && c < C) ? mean[c] : 0.f; n.std[c] = (std && c < C) ? std[c] : 1.f; }
  return n;
}

extern "C" int b200_batch_patchify(const void* img, int32_t img_dtype, void* patches, int32_t B, int32_t C, int32_t H, int32_t W,
                                   int32_t p, int32_t Kpad, int32_t channels_last, const float* mean, const float* std,
                                   void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(img && patches && B > 0 && C > 0 && C <= 8 && p > 0 && H > 0 && W > 0, "b200_batch_patchify: bad arguments (1 <= C <= 8)");
  B200_CHECK_ARG(Kpad % 8 == 0 && Kpad >= C * p * p, "b200_batch_patchify: Kpad=%d must be a multiple of 8 and >= C*p*p", Kpad);
  B200_CHECK_ARG(img_dtype == B200_DTYPE_U8 || img_dtype == B200_DTYPE_BF16 || img_dtype == B200_DTYPE_F32,
                 "b200_batch_patchify: image must be uint8, bf16 or f32");
  if (img_dtype == B200_DTYPE_U8) {
    B200_CHECK_ARG(mean && std, "b200_batch_patchify: uint8 images need mean / std");
    for (int c = 0; c < C; ++c) B200_CHECK_ARG(std[c] != 0.f, "b200_batch_patchify: std[%d] is zero", c);
  }
  const Norm8 nrm = make_norm8(mean, std, C);
  const int gh = (H + p - 1) / p, gw = (W + p - 1) / p;          // pad=True: ragged edges are zero-padded
  const long long total = static_cast<long long>(B) * gh * gw * (Kpad / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  auto* o = reinterpret_cast<__nv_bfloat16*>(patches);
  const int g = static_cast<int>(blocks);
  if (img_dtype == B200_DTYPE_U8) naflex_patchify_kernel<0><<<g, 256, 0, st>>>(img, o, B, C, H, W, p, gh, gw, Kpad, channels_last, nrm);
  else if (img_dtype == B200_DTYPE_BF16) naflex_patchify_kernel<1><<<g, 256, 0, st>>>(img, o, B, C, H, W, p, gh, gw, Kpad, channels_last, nrm);
  else naflex_patchify_kernel<2><<<g, 256, 0, st>>>(img, o, B, C, H, W, p, gh, gw, Kpad, channels_last, nrm);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_patches_u8(const void* patches_u8, void* out, int64_t rows, int32_t K, int32_t Kpad, int32_t C,
                               int32_t channels_last, const float* mean, const float* std, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(patches_u8 && out && mean && std && rows > 0 && K > 0 && C > 0 && C <= 8 && K % C == 0,
                 "b200_patches_u8: bad arguments (1 <= C <= 8, K a multiple of C)");
  B200_CHECK_ARG(Kpad % 8 == 0 && Kpad >= K, "b200_patches_u8: Kpad=%d must be a multiple of 8 and >= K", Kpad);
  for (int c = 0; c < C; ++c) B200_CHECK_ARG(std[c] != 0.f, "b200_patches_u8: std[%d] is zero", c);
  const Norm8 nrm = make_norm8(mean, std, C);
  const long long total = rows * (Kpad / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  patches_u8_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(reinterpret_cast<const uint8_t*>(patches_u8),
                                                              reinterpret_cast<__nv_bfloat16*>(out), rows, K, Kpad, C, channels_last, nrm);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_act_bwd(const void* dy, const void* pre, void* dpre, int64_t n, int32_t act, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(dy && pre && dpre && n > 0 && n % 8 == 0, "b200_act_bwd: null pointer or n not a positive multiple of 8");
  B200_CHECK_ARG(act >= 0 && act <= 2, "b200_act_bwd: bad activation %d", act);
  const long long nvec = n / 8;
  long long blocks = (nvec + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  act_bwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dy),
                                                                 reinterpret_cast<const __nv_bfloat16*>(pre),
                                                                 reinterpret_cast<__nv_bfloat16*>(dpre), nvec, act);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_patchify(const void* img, void* patches, int32_t B, int32_t C, int32_t H, int32_t W, int32_t p,
                             int32_t Kpad, int32_t img_dtype, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(img && patches, "b200_patchify: null pointer");
  B200_CHECK_ARG(B > 0 && C > 0 && p > 0 && H >= p && W >= p, "b200_patchify: bad shape");
  B200_CHECK_ARG(Kpad % 8 == 0 && Kpad >= C * p * p, "b200_patchify: Kpad=%d must be a multiple of 8 and >= C*p*p", Kpad);
  const int gh = H / p, gw = W / p;
  const long long total = static_cast<long long>(B) * gh * gw * (Kpad / 8);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  if (img_dtype == B200_DTYPE_F32)
    patchify_kernel<true><<<static_cast<int>(blocks), threads, 0, st>>>(img, reinterpret_cast<__nv_bfloat16*>(patches), B, C, H, W, p, gh, gw, Kpad);
  else
    patchify_kernel<false><<<static_cast<int>(blocks), threads, 0, st>>>(img, reinterpret_cast<__nv_bfloat16*>(patches), B, C, H, W, p, gh, gw, Kpad);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_patchify_u8(const void* img, void* patches, int32_t B, int32_t C, int32_t H, int32_t W, int32_t p,
                                int32_t Kpad, const float* mean, const float* std, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(img && patches && mean && std, "b200_patchify_u8: null pointer");
  B200_CHECK_ARG(B > 0 && C > 0 && C <= 8 && p > 0 && H >= p && W >= p, "b200_patchify_u8: bad shape (1 <= C <= 8)");
  B200_CHECK_ARG(Kpad % 8 == 0 && Kpad >= C * p * p, "b200_patchify_u8: Kpad=%d must be a multiple of 8 and >= C*p*p", Kpad);
  NormLut lut = {};
  for (int c = 0; c < C; ++c) {
    B200_CHECK_ARG(std[c] != 0.f, "b200_patchify_u8: std[%d] is zero", c);
    for (int v = 0; v < 256; ++v) {
      const volatile float d = static_cast<float>(v) - mean[c];      // fp32 subtract, then fp32 divide (IEEE, no contraction)
      const float f = d / std[c];
      uint32_t u;
      memcpy(&u, &f, 4);
      u += 0x7FFFu + ((u >> 16) & 1u);                               // round to nearest even, as cvt.rn.bf16.f32
      lut.v[c][v] = static_cast<uint16_t>(u >> 16);
    }
  }
  const int gh = H / p, gw = W / p;
  const long long total = static_cast<long long>(B) * gh * gw * (Kpad / 8);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  patchify_u8_kernel<<<static_cast<int>(blocks), threads, 0, st>>>(reinterpret_cast<const uint8_t*>(img),
                                                                    reinterpret_cast<__nv_bfloat16*>(patches), B, C, H, W, p,
                                                                    gh, gw, Kpad, lut);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_assemble_tokens(const void* proj, const void* prefix, const void* pos, void* tokens, int32_t B,
                                    int32_t G, int32_t n_prefix, int32_t D, int32_t param_dtype,
                                    int32_t pos_has_prefix, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(proj && tokens, "b200_assemble_tokens: null pointer");
  B200_CHECK_ARG(n_prefix == 0 || prefix != nullptr, "b200_assemble_tokens: n_prefix > 0 needs prefix tokens");
  B200_CHECK_ARG(B > 0 && G > 0 && n_prefix >= 0 && D > 0 && D % 8 == 0, "b200_assemble_tokens: bad shape (D %% 8 == 0)");
  const long long total = static_cast<long long>(B) * (G + n_prefix) * (D / 8);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  assemble_tokens_kernel<<<static_cast<int>(blocks), threads, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(proj), prefix, pos, reinterpret_cast<__nv_bfloat16*>(tokens), B, G,
      n_prefix, D, param_dtype == B200_DTYPE_F32, pos_has_prefix);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_colsum(const void* x, float* out, int64_t rows, int32_t cols, int32_t accumulate, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(x && out, "b200_colsum: null pointer");
  B200_CHECK_ARG(rows > 0 && cols > 0 && cols % 8 == 0, "b200_colsum: bad shape (cols %% 8 == 0)");
  if (!accumulate) B200_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
  const int gx = (cols + 255) / 256;
  int slabs = (sm_count() * 4 + gx - 1) / gx;
  long long rps = (rows + slabs - 1) / slabs;
  if (rps < 8) rps = 8;
  slabs = static_cast<int>((rows + rps - 1) / rps);
  colsum_kernel<<<dim3(gx, slabs), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), out, rows, cols, rps);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_convert_f32(const float* src, void* dst, int64_t n, int32_t dst_dtype, int32_t accumulate,
                                void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(src && dst, "b200_convert_f32: null pointer");
  B200_CHECK_ARG(n > 0 && n % 4 == 0, "b200_convert_f32: n=%lld must be a positive multiple of 4", (long long)n);
  const int threads = 256;
  long long blocks = (n / 4 + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  if (dst_dtype == B200_DTYPE_F32) convert_kernel<true><<<static_cast<int>(blocks), threads, 0, st>>>(src, dst, n, accumulate);
  else convert_kernel<false><<<static_cast<int>(blocks), threads, 0, st>>>(src, dst, n, accumulate);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_naflex_assemble(const void* proj, const void* prefix, const void* table, const int32_t* pos_off,
                                    const int32_t* pos_len, void* tokens, int32_t B, int32_t N, int32_t n_prefix,
                                    int32_t D, int32_t prefix_dtype, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(proj && tokens, "b200_naflex_assemble: null pointer");
  B200_CHECK_ARG(n_prefix == 0 || prefix != nullptr, "b200_naflex_assemble: n_prefix > 0 needs prefix tokens");
  B200_CHECK_ARG(table == nullptr || (pos_off && pos_len), "b200_naflex_assemble: a position table needs pos_off / pos_len");
  B200_CHECK_ARG(B > 0 && N > 0 && n_prefix >= 0 && D > 0 && D % 8 == 0, "b200_naflex_assemble: bad shape (D %% 8 == 0)");
  const long long total = static_cast<long long>(B) * (N + n_prefix) * (D / 8);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  naflex_assemble_kernel<<<static_cast<int>(blocks), threads, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(proj), prefix, reinterpret_cast<const __nv_bfloat16*>(table), pos_off, pos_len,
      reinterpret_cast<__nv_bfloat16*>(tokens), B, N, n_prefix, D, prefix_dtype == B200_DTYPE_F32);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_naflex_assemble_bwd(const void* dtokens, void* dproj, float* dprefix, float* dtable,
                                        const int32_t* pos_off, const int32_t* pos_len, int32_t B, int32_t N,
                                        int32_t n_prefix, int32_t D, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(dtokens && dproj, "b200_naflex_assemble_bwd: null pointer");
  B200_CHECK_ARG(dtable == nullptr || (pos_off && pos_len), "b200_naflex_assemble_bwd: dtable needs pos_off / pos_len");
  B200_CHECK_ARG(B > 0 && N > 0 && n_prefix >= 0 && D > 0 && D % 8 == 0, "b200_naflex_assemble_bwd: bad shape (D %% 8 == 0)");
  B200_CHECK_ARG(((reinterpret_cast<uintptr_t>(dprefix) | reinterpret_cast<uintptr_t>(dtable)) & 15) == 0,
                 "b200_naflex_assemble_bwd: dprefix / dtable must be 16-byte aligned");
  const long long total = static_cast<long long>(B) * (N + n_prefix) * (D / 8);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  naflex_assemble_bwd_kernel<<<static_cast<int>(blocks), threads, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(dtokens), reinterpret_cast<__nv_bfloat16*>(dproj), dprefix, dtable, pos_off,
      pos_len, B, N, n_prefix, D);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_scale_rows_cols(const void* x, const void* col_scale, const float* row_scale, void* out, int64_t rows,
                                    int32_t cols, int32_t rows_per_scale, int32_t col_scale_dtype, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(x && out, "b200_scale_rows_cols: null pointer");
  B200_CHECK_ARG(rows > 0 && cols > 0 && cols % 8 == 0, "b200_scale_rows_cols: bad shape (cols %% 8 == 0)");
  B200_CHECK_ARG(row_scale == nullptr || rows_per_scale > 0, "b200_scale_rows_cols: row_scale needs rows_per_scale > 0");
  const long long total = rows * (cols / 8);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  scale_rows_cols_kernel<<<static_cast<int>(blocks), threads, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), col_scale, row_scale, reinterpret_cast<__nv_bfloat16*>(out), rows, cols,
      rows_per_scale > 0 ? rows_per_scale : 1, col_scale_dtype == B200_DTYPE_F32);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_colsum_mul(const void* a, const void* b, const float* row_scale, float* out, int64_t rows,
                               int32_t cols, int32_t rows_per_scale, int32_t accumulate, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(a && b && out, "b200_colsum_mul: null pointer");
  B200_CHECK_ARG(rows > 0 && cols > 0 && cols % 8 == 0, "b200_colsum_mul: bad shape (cols %% 8 == 0)");
  B200_CHECK_ARG(row_scale == nullptr || rows_per_scale > 0, "b200_colsum_mul: row_scale needs rows_per_scale > 0");
  if (!accumulate) B200_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
  const int gx = (cols + 255) / 256;
  int slabs = (sm_count() * 4 + gx - 1) / gx;
  long long rps = (rows + slabs - 1) / slabs;
  if (rps < 8) rps = 8;
  slabs = static_cast<int>((rows + rps - 1) / rps);
  colsum_mul_kernel<<<dim3(gx, slabs), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(a),
                                                      reinterpret_cast<const __nv_bfloat16*>(b), row_scale, out, rows, cols,
                                                      rps, rows_per_scale > 0 ? rows_per_scale : 1);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_rope_qk(void* qkv, const float* emb, int32_t B, int32_t N, int32_t H, int32_t num_prefix,
                            int32_t half_layout, int32_t inverse, int32_t per_sample, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(qkv && emb, "b200_rope_qk: null pointer");
  B200_CHECK_ARG(B > 0 && N > 0 && H > 0 && num_prefix >= 0 && num_prefix < N, "b200_rope_qk: bad shape");
  B200_CHECK_ARG(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(emb)) & 15) == 0,
                 "b200_rope_qk: pointers must be 16-byte aligned");
  const long long rows = static_cast<long long>(B) * N;
  const long long total = rows * 2 * H * 4;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  rope_qk_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(qkv), emb, rows, N, H, num_prefix,
                                                          half_layout != 0, inverse != 0, per_sample != 0);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_glu_fwd(const void* h, void* out, int64_t M, int32_t H, int32_t gate_first, int32_t act, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(h && out && M > 0 && H > 0 && H % 8 == 0, "b200_glu_fwd: null pointer or bad shape (H %% 8 == 0)");
  B200_CHECK_ARG(act == 0 || act == 1, "b200_glu_fwd: act must be 0 (SiLU) or 1 (sigmoid)");
  const long long total = static_cast<long long>(M) * (H / 8);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  auto* hp = reinterpret_cast<const __nv_bfloat16*>(h);
  auto* op = reinterpret_cast<__nv_bfloat16*>(out);
  if (act == 0) glu_fwd_kernel<0><<<static_cast<int>(blocks), threads, 0, st>>>(hp, op, M, H, gate_first);
  else glu_fwd_kernel<1><<<static_cast<int>(blocks), threads, 0, st>>>(hp, op, M, H, gate_first);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_glu_bwd(const void* h, const void* dout, void* dh, int64_t M, int32_t H, int32_t gate_first, int32_t act,
                            void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(h && dout && dh && M > 0 && H > 0 && H % 8 == 0, "b200_glu_bwd: null pointer or bad shape (H %% 8 == 0)");
  B200_CHECK_ARG(act == 0 || act == 1, "b200_glu_bwd: act must be 0 (SiLU) or 1 (sigmoid)");
  const long long total = static_cast<long long>(M) * (H / 8);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  auto* hp = reinterpret_cast<const __nv_bfloat16*>(h);
  auto* dp = reinterpret_cast<const __nv_bfloat16*>(dout);
  auto* op = reinterpret_cast<__nv_bfloat16*>(dh);
  if (act == 0) glu_bwd_kernel<0><<<static_cast<int>(blocks), threads, 0, st>>>(hp, dp, op, M, H, gate_first);
  else glu_bwd_kernel<1><<<static_cast<int>(blocks), threads, 0, st>>>(hp, dp, op, M, H, gate_first);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}
