// Exact (erf) GELU and its derivative on packed fp32 pairs, shared by the GEMM epilogues (gemm.cu: BIAS_GELU, DGELU)
// and the GRN backward (grn.cu).  nn.GELU() of timm/layers/mlp.py:24,49.
#pragma once
#include "ptx.cuh"

namespace b200 {

// erf(x / sqrt2) for a pair, Abramowitz-Stegun 7.1.28:  erf(z) = 1 - (1 + a1 z + ... + a6 z^6)^-16, |err| < 3e-7.
// One MUFU (rcp) per element; for large |z| the power overflows to +inf and rcp gives 0, i.e. erf -> 1.
__device__ __forceinline__ f32x2 erf_pair(f32x2 x) {
  float x0, x1;
  upk(x, x0, x1);
  const f32x2 z = pk(fabsf(x0) * 0.70710678118654752f, fabsf(x1) * 0.70710678118654752f);
  f32x2 q = fma2(z, splat(0.0000430638f), splat(0.0002765672f));
  q = fma2(q, z, splat(0.0001520143f));
  q = fma2(q, z, splat(0.0092705272f));
  q = fma2(q, z, splat(0.0422820123f));
  q = fma2(q, z, splat(0.0705230784f));
  q = fma2(q, z, splat(1.0f));
  q = mul2(q, q);
  q = mul2(q, q);
  q = mul2(q, q);
  q = mul2(q, q);
  float d0, d1;
  upk(q, d0, d1);
  const float a0 = 1.0f - rcp_approx(d0), a1 = 1.0f - rcp_approx(d1);
  return pk(copysignf(a0, x0), copysignf(a1, x1));
}
__device__ __forceinline__ f32x2 gelu_erf2(f32x2 x) {
  return mul2(x, fma2(erf_pair(x), splat(0.5f), splat(0.5f)));     // 0.5 x (1 + erf(x / sqrt2))
}
__device__ __forceinline__ f32x2 dgelu_erf2(f32x2 x) {
  // d/dx [x Phi(x)] = Phi(x) + x phi(x),  phi(x) = exp(-x^2 / 2) / sqrt(2 pi).
  // Phi through Abramowitz-Stegun 7.1.26, which shares the exponential phi needs anyway:
  //   erfc(z) = (a1 t + ... + a5 t^5) e^{-z^2},  t = 1 / (1 + p z),  z = |x| / sqrt2,  |err| <= 1.5e-7
  //   q = erfc(z) / 2 = Phi(-|x|);   Phi(x) = 1/2 + copysign(1/2 - q, x)
  float x0, x1;
  upk(x, x0, x1);
  const f32x2 ax = pk(fabsf(x0), fabsf(x1));
  float d0, d1;
  upk(fma2(ax, splat(0.3275911f * 0.70710678118654752f), splat(1.0f)), d0, d1);
  const f32x2 t = pk(rcp_approx(d0), rcp_approx(d1));
  f32x2 q = fma2(t, splat(0.5f * 1.061405429f), splat(0.5f * -1.453152027f));
  q = fma2(q, t, splat(0.5f * 1.421413741f));
  q = fma2(q, t, splat(0.5f * -0.284496736f));
  q = fma2(q, t, splat(0.5f * 0.254829592f));
  q = mul2(q, t);
  float w0, w1;
  upk(mul2(mul2(x, x), splat(-0.72134752044448170f)), w0, w1);     // -x^2/2 * log2(e)
  const f32x2 e = pk(ex2_approx(w0), ex2_approx(w1));
  float s0, s1;
  upk(fma2(q, mul2(e, splat(-1.0f)), splat(0.5f)), s0, s1);        // 1/2 - q e  (>= 0)
  const f32x2 cdf = add2(pk(copysignf(s0, x0), copysignf(s1, x1)), splat(0.5f));
  return fma2(mul2(x, splat(0.3989422804014327f)), e, cdf);
}

// GELU'(x) on the FMA pipe only (no MUFU): GELU'(x) - 1/2 is odd and saturates at +-1/2, so with z = clamp(x / R, -1, 1)
// and t = 2 z^2 - 1:  GELU'(x) ~= 1/2 + z Q(t),  Q a degree-9 least-squares polynomial on R = 4.5 constrained to
// z Q(t) = 1/2 at z = 1 (the clamp then continues it with the exact limits 0 / 1).  Max abs error 6.9e-5 over all x
// (measured in fp32 Horner arithmetic; GELU'(4.5) - 1 = 6.9e-5 is the saturation step) -- a factor applied to a bf16
// gradient (relative rounding 3.9e-3).  The fc2-dgrad epilogue was bound by the 16-lane MUFU (2 per element: rcp + ex2)
// and by its FMA count; this is ~15 instructions per PAIR instead of ~26 + 4 MUFU.
__device__ __forceinline__ f32x2 dgelu_poly2(f32x2 x) {
  float x0, x1;
  upk(mul2(x, splat(1.0f / 4.5f)), x0, x1);
  const f32x2 z = pk(fminf(fmaxf(x0, -1.0f), 1.0f), fminf(fmaxf(x1, -1.0f), 1.0f));
  const f32x2 t = fma2(mul2(z, z), splat(2.0f), splat(-1.0f));
  f32x2 q = fma2(t, splat(-1.114164513e-01f), splat(2.526838988e-01f));
  q = fma2(q, t, splat(-2.053749153e-01f));
  q = fma2(q, t, splat(2.136556465e-01f));
  q = fma2(q, t, splat(-3.936981523e-01f));
  q = fma2(q, t, splat(4.730997042e-01f));
  q = fma2(q, t, splat(-4.296489081e-01f));
  q = fma2(q, t, splat(3.878010022e-01f));
  q = fma2(q, t, splat(-4.046076170e-01f));
  q = fma2(q, t, splat(7.175057925e-01f));
  return fma2(z, q, splat(0.5f));
}

// GELU(x) and GELU'(x) together from one evaluation of Phi (7.1.26) and phi: GELU = x Phi, GELU' = Phi + x phi.
__device__ __forceinline__ void gelu_and_dgelu2(f32x2 x, f32x2& g, f32x2& dg) {
  float x0, x1;
  upk(x, x0, x1);
  const f32x2 ax = pk(fabsf(x0), fabsf(x1));
  float d0, d1;
  upk(fma2(ax, splat(0.3275911f * 0.70710678118654752f), splat(1.0f)), d0, d1);
  const f32x2 t = pk(rcp_approx(d0), rcp_approx(d1));
  f32x2 q = fma2(t, splat(0.5f * 1.061405429f), splat(0.5f * -1.453152027f));
  q = fma2(q, t, splat(0.5f * 1.421413741f));
  q = fma2(q, t, splat(0.5f * -0.284496736f));
  q = fma2(q, t, splat(0.5f * 0.254829592f));
  q = mul2(q, t);
  float w0, w1;
  upk(mul2(mul2(x, x), splat(-0.72134752044448170f)), w0, w1);
  const f32x2 e = pk(ex2_approx(w0), ex2_approx(w1));
  float s0, s1;
  upk(fma2(q, mul2(e, splat(-1.0f)), splat(0.5f)), s0, s1);
  const f32x2 cdf = add2(pk(copysignf(s0, x0), copysignf(s1, x1)), splat(0.5f));
  g = mul2(x, cdf);
  dg = fma2(mul2(x, splat(0.3989422804014327f)), e, cdf);
}

// The other MLP activations timm configures on this path (create_act.py: 'gelu_tanh' = GELUTanh / nn.GELU('tanh'), used by
// the SigLIP ViTs and timm's test_convnext; 'silu'): selected at run time by b200_gemm_desc.act in the generic kernels.
//   tanh(u) = 1 - 2 / (1 + e^{2u})  (ex2 + rcp: two MUFU, abs. error ~1e-6, saturates correctly at +-inf)
__device__ __forceinline__ float tanh_fast(float u) {
  return 1.f - 2.f * rcp_approx(1.f + ex2_approx(u * 2.8853900817779268f));
}
__device__ __forceinline__ float gelu_tanh1(float x) {
  return 0.5f * x * (1.f + tanh_fast(0.7978845608028654f * fmaf(0.044715f * x * x, x, x)));
}
__device__ __forceinline__ float dgelu_tanh1(float x) {
  const float t = tanh_fast(0.7978845608028654f * fmaf(0.044715f * x * x, x, x));
  return fmaf(0.5f * x * (1.f - t * t), 0.7978845608028654f * fmaf(0.134145f * x, x, 1.f), 0.5f * (1.f + t));
}
__device__ __forceinline__ float silu1(float x) { return x * rcp_approx(1.f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float dsilu1(float x) {
  const float s = rcp_approx(1.f + ex2_approx(-1.4426950408889634f * x));
  return s * fmaf(x, 1.f - s, 1.f);
}
// act: 0 = exact GELU, 1 = tanh GELU, 2 = SiLU
__device__ __forceinline__ f32x2 act_fwd2(f32x2 v, int act) {
  if (act == 0) return gelu_erf2(v);
  float a, b;
  upk(v, a, b);
  return act == 1 ? pk(gelu_tanh1(a), gelu_tanh1(b)) : pk(silu1(a), silu1(b));
}
__device__ __forceinline__ f32x2 act_bwd2(f32x2 v, int act) {
  if (act == 0) return dgelu_erf2(v);
  float a, b;
  upk(v, a, b);
  return act == 1 ? pk(dgelu_tanh1(a), dgelu_tanh1(b)) : pk(dsilu1(a), dsilu1(b));
}

}  // namespace b200
