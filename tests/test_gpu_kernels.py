"""Kernel-level parity on the GPU: every libb200timm entry point against a plain PyTorch fp32 restatement of
the same op on the same seeded inputs (the op-level oracle; model-level parity lives in test_gpu_models.py).

Tolerances: outputs are bf16, accumulations fp32.  A bf16 result of magnitude |y| carries up to |y| * 2^-8
rounding error, so comparisons use atol + rtol * |ref| with rtol = 1e-2 unless stated otherwise.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from b200timm import ops
    return ops


def _close(got, ref, atol, rtol, what):
    got = got.float()
    ref = ref.float()
    assert got.shape == ref.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(got).all(), f"{what}: non-finite values"
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.numel()} elements off, max abs err {err.max().item():.4g} "
                           f"(ref max {ref.abs().max().item():.4g})")


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x * 0.7071067811865476))


def _dgelu(x):
    return 0.5 * (1.0 + torch.erf(x * 0.7071067811865476)) + x * torch.exp(-0.5 * x * x) * 0.3989422804014327


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (1000, 768, 768), (394, 200, 72), (77, 1000, 192),
                                   (640, 520, 320)])   # 5 m-blocks: the last CTA pair has an empty second half
def test_gemm_layouts(a_mn, b_mn, M, N, K):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    b = torch.randn(N, K, device=DEV, generator=g).bfloat16()
    # MN-major operands need M (resp. N) to be a multiple of 8 for the 16-byte TMA stride rule.
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("MN-major operand with a leading dimension that is not a multiple of 8")
    a_in = a.t().contiguous() if a_mn else a
    b_in = b.t().contiguous() if b_mn else b
    out = ops.gemm(a_in, b_in, M, N, K, a_mn=a_mn, b_mn=b_mn)
    ref = a.float() @ b.float().t()
    _close(out, ref, atol=2e-2 * math.sqrt(K / 64), rtol=1e-2, what=f"gemm a_mn={a_mn} b_mn={b_mn} {M}x{N}x{K}")


@pytest.mark.parametrize("param_dtype", [torch.bfloat16, torch.float32])
def test_gemm_epilogues(param_dtype):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(11)
    M, N, K = 1000, 768, 384
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g).to(param_dtype)
    gamma = torch.randn(N, device=DEV, generator=g).to(param_dtype)
    aux = torch.randn(M, N, device=DEV, generator=g).bfloat16()
    acc = a.float() @ w.float().t() + bias.float()

    _close(ops.gemm(a, w, M, N, K, bias=bias), acc, 2e-2, 1e-2, "bias")
    _close(ops.gemm(a, w, M, N, K, bias=bias, epilogue=ops.EPI_BIAS_GELU), _gelu(acc), 2e-2, 1e-2, "bias+gelu")
    _close(ops.gemm(a, w, M, N, K, bias=bias, epilogue=ops.EPI_BIAS_RESIDUAL, aux=aux), aux.float() + acc, 3e-2, 1e-2,
           "bias+residual")
    _close(ops.gemm(a, w, M, N, K, bias=bias, epilogue=ops.EPI_BIAS_RESIDUAL, aux=aux, gamma=gamma),
           aux.float() + gamma.float() * acc, 3e-2, 1e-2, "bias+layerscale+residual")
    _close(ops.gemm(a, w, M, N, K, bias=bias, epilogue=ops.EPI_DGELU, aux=aux), acc * _dgelu(aux.float()), 2e-2, 1e-2,
           "dgelu")


def test_gemm_fp32_out_accumulate_splitk():
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(5)
    M, N, K = 768, 384, 4096
    a = torch.randn(K, M, device=DEV, generator=g).bfloat16()  # stored [K, M]  (a_mn)
    b = torch.randn(K, N, device=DEV, generator=g).bfloat16()  # stored [K, N]  (b_mn)
    ref = a.float().t() @ b.float()
    out = ops.gemm(a, b, M, N, K, a_mn=True, b_mn=True, out_dtype=torch.float32)
    _close(out, ref, 1e-2, 1e-3, "fp32 out")
    base = torch.randn(M, N, device=DEV, generator=g)
    out2 = base.clone()
    ops.gemm(a, b, M, N, K, a_mn=True, b_mn=True, out=out2, accumulate=True)
    _close(out2, base + ref, 1e-2, 1e-3, "accumulate")
    out3 = base.clone()
    ops.gemm(a, b, M, N, K, a_mn=True, b_mn=True, out=out3, split_k=8)
    _close(out3, base + ref, 1e-2, 1e-3, "split_k=8 (atomic accumulate)")


def test_gemm_vit_base_shapes():
    """The four ViT-B/16 GEMMs at batch 16 (M = 16*197 = 3152, not a multiple of 128)."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(3)
    M = 16 * 197
    for N, K in ((2304, 768), (768, 768), (3072, 768), (768, 3072)):
        a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
        w = (torch.randn(N, K, device=DEV, generator=g) * 0.02).bfloat16()
        bias = torch.randn(N, device=DEV, generator=g).bfloat16()
        ref = a.float() @ w.float().t() + bias.float()
        _close(ops.linear_fwd(a, w, bias), ref, 2e-2, 1e-2, f"linear {N}x{K}")


def test_gemm_rejects_bad_arguments():
    ops = _ops()
    from b200timm import B200Error
    a = torch.zeros(16, 60, device=DEV, dtype=torch.bfloat16)  # K=60: row stride not a multiple of 16 bytes
    w = torch.zeros(8, 60, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(B200Error):
        ops.gemm(a, w, 16, 8, 60)
    with pytest.raises(B200Error):
        ops.gemm(torch.zeros(16, 64), torch.zeros(8, 64), 16, 8, 64)  # CPU tensors: no fallback


# ------------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("cols", [192, 768, 1024, 128, 1536])
@pytest.mark.parametrize("xdt,pdt", [(torch.bfloat16, torch.bfloat16), (torch.float32, torch.float32),
                                     (torch.bfloat16, torch.float32)])
def test_layernorm_fwd_bwd(cols, xdt, pdt):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(cols)
    rows = 1003
    x = (torch.randn(rows, cols, device=DEV, generator=g) * 2 + 0.5).to(xdt)
    w = (torch.randn(cols, device=DEV, generator=g) * 0.5 + 1).to(pdt)
    b = torch.randn(cols, device=DEV, generator=g).to(pdt)
    dy = torch.randn(rows, cols, device=DEV, generator=g).to(xdt)
    eps = 1e-6
    y, mean, rstd = ops.layernorm_fwd(x, w, b, eps, save_stats=True)
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True)
    br = b.float().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (cols,), wr, br, eps)
    atol = 3e-2 if xdt == torch.bfloat16 else 1e-4
    rtol = 1e-2 if xdt == torch.bfloat16 else 1e-4
    _close(y, yr.detach(), atol, rtol, "ln fwd")
    _close(mean, x.float().mean(-1), 1e-4, 1e-4, "ln mean")
    _close(rstd, 1.0 / torch.sqrt(x.float().var(-1, unbiased=False) + eps), 1e-3, 1e-3, "ln rstd")
    yr.backward(dy.float())
    dx, dw, db = ops.layernorm_bwd(dy, x, w, mean, rstd)
    _close(dx, xr.grad, atol, rtol, "ln dx")
    _close(dw, wr.grad, 2e-2 if xdt == torch.bfloat16 else 2e-3, 1e-3, "ln dweight")
    _close(db, br.grad, 2e-2 if xdt == torch.bfloat16 else 2e-3, 1e-3, "ln dbias")


# ------------------------------------------------------------------------------------------------ attention
def _attn_ref(qkv, B, N, H, hd, scale, kv_len=None):
    q, k, v = qkv.float().view(B, N, 3, H, hd).permute(2, 0, 3, 1, 4).unbind(0)
    s = (q @ k.transpose(-1, -2)) * scale
    if kv_len is not None:
        keys = torch.arange(N, device=qkv.device)[None, None, None, :]
        s = s.masked_fill(keys >= kv_len.view(B, 1, 1, 1), float("-inf"))
    p = s.softmax(-1)
    o = (p @ v).transpose(1, 2).reshape(B, N, H * hd)
    return o, torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,N,H", [(2, 197, 12), (3, 64, 3), (2, 256, 4), (1, 577, 16), (2, 208, 2), (2, 129, 2),
                                   (1, 1025, 2)])
def test_attention_fwd(B, N, H):
    ops = _ops()
    hd = 64
    g = torch.Generator(device=DEV).manual_seed(N + H)
    qkv = torch.randn(B, N, 3 * H * hd, device=DEV, generator=g).bfloat16()
    scale = hd ** -0.5
    out, lse = ops.attention_fwd(qkv, B, N, H, hd, scale, need_lse=True)
    ref, lse_ref = _attn_ref(qkv, B, N, H, hd, scale)
    _close(out, ref, 1e-2, 2e-2, f"attention B={B} N={N} H={H}")
    _close(lse, lse_ref, 1e-2, 1e-3, "lse")


@pytest.mark.parametrize("N", [197, 577])
def test_attention_fwd_key_padding(N):
    ops = _ops()
    B, H, hd = 4, 4, 64
    g = torch.Generator(device=DEV).manual_seed(N)
    qkv = (torch.randn(B, N, 3 * H * hd, device=DEV, generator=g) * 1.5).bfloat16()
    kv_len = torch.tensor([N, 1, N // 2, 130], device=DEV, dtype=torch.int32)
    out, lse = ops.attention_fwd(qkv, B, N, H, hd, 0.125, kv_len=kv_len, need_lse=True)
    ref, lse_ref = _attn_ref(qkv, B, N, H, hd, 0.125, kv_len)
    _close(out, ref, 1e-2, 2e-2, "masked attention")
    _close(lse, lse_ref, 1e-2, 1e-3, "masked lse")


# ------------------------------------------------------------------------------------------------ helpers
@pytest.mark.parametrize("p,size,dt", [(16, 224, torch.bfloat16), (14, 224, torch.bfloat16), (16, 64, torch.float32),
                                       (4, 32, torch.bfloat16)])
def test_patchify(p, size, dt):
    ops = _ops()
    B, C = 3, 3
    g = torch.Generator(device=DEV).manual_seed(p)
    img = torch.randn(B, C, size, size, device=DEV, generator=g).to(dt)
    out = ops.patchify(img, p)
    K = C * p * p
    ref = torch.nn.functional.unfold(img.float(), kernel_size=p, stride=p).transpose(1, 2).reshape(-1, K)
    assert out.shape[1] % 8 == 0 and out.shape[1] >= K
    assert torch.equal(out[:, :K].float(), ref.bfloat16().float())
    assert (out[:, K:] == 0).all()


def test_assemble_tokens_and_colsum():
    ops = _ops()
    B, G, D, npre = 3, 196, 768, 1
    g = torch.Generator(device=DEV).manual_seed(0)
    proj = torch.randn(B * G, D, device=DEV, generator=g).bfloat16()
    cls = torch.randn(1, 1, D, device=DEV, generator=g).bfloat16()
    pos = torch.randn(1, G + npre, D, device=DEV, generator=g).bfloat16()
    tok = ops.assemble_tokens(proj, cls, pos, B, G, npre, D)
    ref = torch.cat([cls.float().expand(B, -1, -1), proj.float().view(B, G, D)], 1) + pos.float()
    assert torch.equal(tok.float(), ref.bfloat16().float())
    x = torch.randn(5000, 776, device=DEV, generator=g).bfloat16()
    _close(ops.colsum(x), x.float().sum(0), 2e-2, 1e-4, "colsum")


# ------------------------------------------------------------------------------------------------ backward kernels
@pytest.mark.parametrize("B,N,H", [(2, 197, 12), (3, 64, 2), (2, 208, 3), (2, 129, 2), (1, 17, 1),
                                   (2, 577, 4), (1, 300, 2), (3, 224, 1), (1, 1025, 1),    # N > 208: attention_bwd_long.cu
                                   (1, 1370, 2), (1, 2240, 1)])                            # DINOv2 @518; the kernel's limit
def test_attention_bwd(B, N, H):
    """dq/dk/dv against PyTorch autograd of the fp32 restatement of attention.py:115-135."""
    ops = _ops()
    hd = 64
    g = torch.Generator(device=DEV).manual_seed(100 + N + H)
    qkv = torch.randn(B, N, 3 * H * hd, device=DEV, generator=g).bfloat16()
    dout = torch.randn(B, N, H * hd, device=DEV, generator=g).bfloat16()
    scale = hd ** -0.5
    out, lse = ops.attention_fwd(qkv, B, N, H, hd, scale, need_lse=True)
    dqkv = ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd, scale)
    ref_in = qkv.float().requires_grad_(True)
    ref_out, _ = _attn_ref(ref_in, B, N, H, hd, scale)
    ref_out.backward(dout.float())
    _close(dqkv, ref_in.grad, 2e-2, 2e-2, f"attention bwd B={B} N={N} H={H}")


@pytest.mark.parametrize("N", [197, 577])
def test_attention_bwd_key_padding(N):
    ops = _ops()
    B, H, hd = 4, 3, 64
    g = torch.Generator(device=DEV).manual_seed(7)
    qkv = torch.randn(B, N, 3 * H * hd, device=DEV, generator=g).bfloat16()
    dout = torch.randn(B, N, H * hd, device=DEV, generator=g).bfloat16()
    kv_len = torch.tensor([N, 1, 100, 130 if N == 197 else 400], device=DEV, dtype=torch.int32)
    out, lse = ops.attention_fwd(qkv, B, N, H, hd, 0.125, kv_len=kv_len, need_lse=True)
    dqkv = ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd, 0.125, kv_len=kv_len)
    ref_in = qkv.float().requires_grad_(True)
    ref_out, _ = _attn_ref(ref_in, B, N, H, hd, 0.125, kv_len)
    ref_out.backward(dout.float())
    _close(dqkv, ref_in.grad, 2e-2, 2e-2, "masked attention bwd")


def test_gemm_dual_output_and_dgrad_wgrad():
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(21)
    M, N, K = 1576, 3072, 768
    x = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) * 0.03).bfloat16()
    b = torch.randn(N, device=DEV, generator=g).bfloat16()
    pre = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    act = ops.linear_fwd(x, w, b, epilogue=ops.EPI_BIAS_GELU, out2=pre)
    ref_pre = x.float() @ w.float().t() + b.float()
    _close(pre, ref_pre, 2e-2, 1e-2, "pre-activation (out2)")
    assert torch.equal(act.float(), _gelu(pre.float()).bfloat16().float()) or \
        (act.float() - _gelu(pre.float())).abs().max() < 2e-2   # GELU applied to the stored bf16 pre-activation
    dy = torch.randn(M, N, device=DEV, generator=g).bfloat16()
    _close(ops.linear_dgrad(dy, w), dy.float() @ w.float(), 5e-2, 1e-2, "dgrad")
    da = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w2 = (torch.randn(K, N, device=DEV, generator=g) * 0.03).bfloat16()   # fc2 weight [K_out=768, N=3072]
    _close(ops.linear_dgrad(da, w2, epilogue=ops.EPI_DGELU, aux=pre), (da.float() @ w2.float()) * _dgelu(pre.float()),
           5e-2, 2e-2, "dgrad + GELU'")
    _close(ops.linear_wgrad(dy, x), dy.float().t() @ x.float(), 5e-2, 2e-3, "wgrad (split-K, fp32)")
    gbuf = torch.randn(N, K, device=DEV, generator=g).bfloat16()
    before = gbuf.float().clone()
    ws = ops.linear_wgrad(dy, x)
    ops.convert_f32(ws, gbuf, accumulate=True)
    _close(gbuf, before + ws, 0.5, 1e-2, "convert_f32 accumulate into bf16 grad")


def test_layernorm_bwd_with_residual_gradient():
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(4)
    rows, cols = 777, 768
    x = torch.randn(rows, cols, device=DEV, generator=g).bfloat16()
    w = torch.randn(cols, device=DEV, generator=g).bfloat16()
    b = torch.randn(cols, device=DEV, generator=g).bfloat16()
    dy = torch.randn(rows, cols, device=DEV, generator=g).bfloat16()
    dres = torch.randn(rows, cols, device=DEV, generator=g).bfloat16()
    _, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-6, save_stats=True)
    dx0, _, _ = ops.layernorm_bwd(dy, x, w, mean, rstd)
    dx1, _, _ = ops.layernorm_bwd(dy, x, w, mean, rstd, dres=dres)
    xr = x.float().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (cols,), w.float(), b.float(), 1e-6).backward(dy.float())
    _close(dx1, xr.grad + dres.float(), 4e-2, 1e-2, "ln bwd + residual grad")
    _close(dx0, xr.grad, 3e-2, 1e-2, "ln bwd")


def test_layerscale_droppath_kernels():
    """Residual epilogue with per-sample DropPath factor + LayerScale and its out2 (branch output); the two backward
    helper kernels."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(33)
    B, N, D, K = 6, 50, 256, 128
    M = B * N
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(D, K, device=DEV, generator=g) * 0.1).bfloat16()
    bias = torch.randn(D, device=DEV, generator=g).bfloat16()
    gamma = torch.randn(D, device=DEV, generator=g).bfloat16()
    aux = torch.randn(M, D, device=DEV, generator=g).bfloat16()
    rs = torch.tensor([0.0, 1.25, 1.25, 0.0, 1.25, 1.25], device=DEV)
    branch = a.float() @ w.float().t() + bias.float()
    rows = rs.repeat_interleave(N)[:, None]
    out = ops.linear_fwd(a, w, bias, epilogue=ops.EPI_BIAS_RESIDUAL, aux=aux, gamma=gamma, row_scale=rs, rows_per_scale=N)
    _close(out, aux.float() + rows * gamma.float() * branch, 3e-2, 1e-2, "residual + layerscale + droppath")
    assert torch.equal(out.view(B, N, D)[0], aux.view(B, N, D)[0])          # dropped sample: branch contributes exactly 0
    br = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    out2 = ops.linear_fwd(a, w, bias, epilogue=ops.EPI_BIAS_RESIDUAL, aux=aux, gamma=gamma, row_scale=rs, rows_per_scale=N,
                          out2=br)
    _close(br, branch, 2e-2, 1e-2, "branch output (out2)")
    _close(out2, aux.float() + rows * gamma.float() * br.float(), 2e-2, 1e-2, "residual computed from the stored branch")
    out3 = ops.linear_fwd(a, w, bias, epilogue=ops.EPI_BIAS_RESIDUAL, aux=aux, row_scale=rs, rows_per_scale=N)
    _close(out3, aux.float() + rows * branch, 3e-2, 1e-2, "residual + droppath (no layerscale)")
    dy = torch.randn(M, D, device=DEV, generator=g).bfloat16()
    _close(ops.scale_rows_cols(dy, gamma, rs, N), dy.float() * gamma.float() * rows, 2e-2, 1e-2, "scale_rows_cols")
    _close(ops.scale_rows_cols(dy, None, rs, N), dy.float() * rows, 1e-2, 1e-2, "scale rows only")
    _close(ops.colsum_mul(dy, br, rs, N), (dy.float() * br.float() * rows).sum(0), 5e-2, 2e-3, "colsum_mul")


@pytest.mark.parametrize("cols,pdt", [(768, torch.bfloat16), (192, torch.float32), (1024, torch.bfloat16)])
def test_rmsnorm_fwd_bwd(cols, pdt):
    """RMS mode of the norm kernels (timm RmsNorm, norm.py:202-248 / fast_norm.py:119-139) against fp32 autograd."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(cols)
    rows = 333
    x = torch.randn(rows, cols, device=DEV, generator=g).bfloat16()
    w = (1 + 0.2 * torch.randn(cols, device=DEV, generator=g)).to(pdt)
    dy = torch.randn(rows, cols, device=DEV, generator=g).bfloat16()
    dres = torch.randn(rows, cols, device=DEV, generator=g).bfloat16()
    eps = ops.RmsEps(1e-6)
    y, mean, rstd = ops.layernorm_fwd(x, w, None, eps, save_stats=True)
    assert mean is None and rstd is not None
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True)
    yr = xr * torch.rsqrt((xr * xr).mean(-1, keepdim=True) + 1e-6) * wr
    _close(y, yr.detach(), atol=2e-2, rtol=1e-2, what="rmsnorm fwd")
    yr.backward(dy.float())
    dx, dw, db = ops.layernorm_bwd(dy, x, w, None, rstd, dres=dres)
    assert db is None
    _close(dx, xr.grad + dres.float(), atol=4e-2, rtol=2e-2, what="rmsnorm dx (+ residual gradient)")
    _close(dw, wr.grad, atol=0.25, rtol=2e-2, what="rmsnorm dweight")


@pytest.mark.parametrize("M", [1576, 50432 // 8, 300])
def test_dgrad_epilogue_colsum_is_the_bias_gradient(M):
    """EPI_DGELU with `colsum`: the column sums of the bf16 output (the fc1 bias gradient) leave the GEMM epilogue; they
    must equal b200_colsum of the stored output (same bf16 values, fp32 sums: order differs, hence rtol 1e-4) and the
    output itself must not change.  Also pins the FMA-pipe GELU' polynomial (gelu.cuh dgelu_poly2, |err| <= 6.9e-5)."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(5)
    N, K = 768, 3072                       # dY [M, 768] @ W2 [768, 3072] -> dpre [M, 3072]
    dy = torch.randn(M, N, device=DEV, generator=g).bfloat16()
    w2 = (torch.randn(N, K, device=DEV, generator=g) * 0.03).bfloat16()
    pre = (torch.randn(M, K, device=DEV, generator=g) * 2.0).bfloat16()
    plain = ops.linear_dgrad(dy, w2, epilogue=ops.EPI_DGELU, aux=pre)
    assert ops.dgrad_colsum_fused(M, K)
    db = torch.zeros(K, device=DEV, dtype=torch.float32)
    fused = ops.linear_dgrad(dy, w2, epilogue=ops.EPI_DGELU, aux=pre, colsum=db)
    assert torch.equal(plain, fused)
    ref = ops.colsum(fused)
    torch.testing.assert_close(db, ref, atol=2e-3, rtol=1e-4)
    _close(fused, (dy.float() @ w2.float()) * _dgelu(pre.float()), 5e-2, 2e-2, "dgrad + GELU' (polynomial)")
    # accumulate semantics: a second launch adds
    ops.linear_dgrad(dy, w2, epilogue=ops.EPI_DGELU, aux=pre, colsum=db)
    torch.testing.assert_close(db, 2 * ref, atol=4e-3, rtol=1e-4)


def test_dgelu_polynomial_pointwise():
    """GELU'(x) of the DGELU epilogue over a dense grid incl. the saturated tails: identity weight, dY = 1."""
    ops = _ops()
    M, K = 256, 256
    x = torch.linspace(-9, 9, M * K, device=DEV).view(M, K).bfloat16()
    eye = torch.eye(K, device=DEV).bfloat16()
    ones = torch.ones(M, K, device=DEV).bfloat16()
    got = ops.linear_dgrad(ones, eye, epilogue=ops.EPI_DGELU, aux=x).float()
    ref = _dgelu(x.float())
    assert (got - ref).abs().max().item() <= 4e-3 + 7e-5        # bf16 rounding of the output (|GELU'| <= 1.13) + polynomial


@pytest.mark.parametrize("B,N,H", [(2, 197, 12), (3, 64, 2), (2, 208, 3), (1, 17, 1)])
def test_attention_bwd_bias_gradient(B, N, H):
    """b200_attn_bwd with dbias: the column sums of dqkv (the qkv bias gradient) leave the kernel's store epilogue."""
    ops = _ops()
    hd = 64
    g = torch.Generator(device=DEV).manual_seed(300 + N)
    qkv = torch.randn(B, N, 3 * H * hd, device=DEV, generator=g).bfloat16()
    dout = torch.randn(B, N, H * hd, device=DEV, generator=g).bfloat16()
    out, lse = ops.attention_fwd(qkv, B, N, H, hd, 0.125, need_lse=True)
    plain = ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd, 0.125)
    assert ops.attn_bwd_bias_fused(N)
    db = torch.zeros(3 * H * hd, device=DEV, dtype=torch.float32)
    fused = ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd, 0.125, dbias=db)
    assert torch.equal(plain, fused)
    torch.testing.assert_close(db, ops.colsum(fused.view(B * N, -1)), atol=2e-3, rtol=1e-4)
    kv_len = torch.tensor([N] + [max(1, N // 2)] * (B - 1), device=DEV, dtype=torch.int32)
    out, lse = ops.attention_fwd(qkv, B, N, H, hd, 0.125, kv_len=kv_len, need_lse=True)
    db.zero_()
    fused = ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd, 0.125, kv_len=kv_len, dbias=db)
    torch.testing.assert_close(db, ops.colsum(fused.view(B * N, -1)), atol=2e-3, rtol=1e-4)


@pytest.mark.parametrize("cols", [3072, 4096, 2056, 8192])
@pytest.mark.parametrize("pdt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("rms", [False, True])
def test_layernorm_wide_rows(cols, pdt, rms):
    """Rows wider than 2048 (the inner norm of an Mlp over the hidden width): the CTA-per-row kernels, LayerNorm and RMSNorm,
    forward, backward (with a residual gradient) and the parameter-gradient reduction, against torch fp32."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(cols)
    rows = 777
    x = (torch.randn(rows, cols, device=DEV, generator=g) * 2 + 0.5).bfloat16()
    w = (torch.randn(cols, device=DEV, generator=g) * 0.5 + 1).to(pdt)
    b = None if rms else torch.randn(cols, device=DEV, generator=g).to(pdt)
    dy = torch.randn(rows, cols, device=DEV, generator=g).bfloat16()
    dres = torch.randn(rows, cols, device=DEV, generator=g).bfloat16()
    eps = ops.RmsEps(1e-6) if rms else 1e-6
    y, mean, rstd = ops.layernorm_fwd(x, w, b, eps, save_stats=True)
    xr, wr = x.float().requires_grad_(True), w.float().requires_grad_(True)
    if rms:
        yr = xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-6) * wr
    else:
        br = b.float().requires_grad_(True)
        yr = torch.nn.functional.layer_norm(xr, (cols,), wr, br, 1e-6)
    _close(y, yr.detach(), 3e-2, 1e-2, "wide ln fwd")
    yr.backward(dy.float())
    dx, dw, db = ops.layernorm_bwd(dy, x, w, mean, rstd, dres=dres)
    _close(dx, xr.grad + dres.float(), 4e-2, 1e-2, "wide ln dx (+ residual gradient)")
    _close(dw, wr.grad, 3e-2, 2e-3, "wide ln dweight")
    if not rms:
        _close(db, br.grad, 3e-2, 2e-3, "wide ln dbias")


# ------------------------------------------------------------------------------------------------ causal / general masks
def _bias_cases(kind, B, N, H, g):
    """(reference additive bias broadcastable to [B,H,N,N], the reference's attn_mask / is_causal arguments)."""
    if kind == "causal":
        return torch.full((N, N), float("-inf"), device=DEV).triu_(1), None, True
    if kind == "bool_per_sample":       # random keep-mask per sample, the diagonal always kept (no empty rows)
        keep = torch.rand(B, 1, N, N, device=DEV, generator=g) > 0.4
        keep |= torch.eye(N, device=DEV, dtype=torch.bool)
        return torch.zeros(B, 1, N, N, device=DEV).masked_fill(~keep, float("-inf")), keep, False
    if kind == "float_per_head":        # additive bias per head (relative-position style), shared by the batch
        bias = torch.randn(1, H, N, N, device=DEV, generator=g)
        return bias, bias, False
    if kind == "float_full":
        bias = torch.randn(B, H, N, N, device=DEV, generator=g)
        return bias, bias.bfloat16(), False      # a bf16 mask is accepted (converted to fp32)
    raise AssertionError(kind)


@pytest.mark.parametrize("kind", ["causal", "bool_per_sample", "float_per_head", "float_full"])
@pytest.mark.parametrize("B,N,H", [(2, 197, 3), (2, 65, 2), (2, 208, 1), (1, 577, 2), (2, 300, 1)])
def test_attention_bias_fwd_bwd(kind, B, N, H):
    """softmax(q k^T * scale + bias) v and its gradient (b200_attn_fwd_bias / b200_attn_bwd_bias: is_causal and general
    attn_mask of timm/layers/attention.py:21-40,108-133) against fp32 autograd; all four production kernels
    (N <= 208: attention_fwd2 / bwd2, above: the long-sequence pair)."""
    import b200timm.layers as L
    ops = _ops()
    hd, scale = 64, 0.125
    g = torch.Generator(device=DEV).manual_seed(N + 7 * H + len(kind))
    qkv = torch.randn(B, N, 3 * H * hd, device=DEV, generator=g).bfloat16()
    dout = torch.randn(B, N, H * hd, device=DEV, generator=g).bfloat16()
    ref_bias, mask_arg, causal = _bias_cases(kind, B, N, H, g)
    if mask_arg is not None and mask_arg.dtype == torch.bfloat16:
        ref_bias = mask_arg.float()
    operand = L.attn_mask_operand(mask_arg, causal, N, B, DEV)
    assert operand.dtype == torch.float32 and operand.dim() == 4 and operand.shape[-1] % 4 == 0
    out, lse = ops.attention_fwd(qkv, B, N, H, hd, scale, kv_len=operand, need_lse=True)
    dqkv = ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd, scale, kv_len=operand)
    ref_in = qkv.float().requires_grad_(True)
    q, k, v = ref_in.view(B, N, 3, H, hd).permute(2, 0, 3, 1, 4).unbind(0)
    sc = (q @ k.transpose(-1, -2)) * scale + ref_bias
    ref = (sc.softmax(-1) @ v).transpose(1, 2).reshape(B, N, H * hd)
    ref.backward(dout.float())
    _close(out, ref, 1e-2, 2e-2, f"biased attention {kind} N={N}")
    _close(lse, torch.logsumexp(sc, -1), 1e-2, 1e-3, "biased lse")
    _close(dqkv, ref_in.grad, 2e-2, 2e-2, f"biased attention bwd {kind} N={N}")
    if kind == "causal":      # token 0 sees only itself: its output row is exactly v[0]
        v0 = qkv.view(B, N, 3, H, hd)[:, 0, 2].reshape(B, H * hd)
        assert torch.equal(out[:, 0], v0)


def test_attention_bias_with_fused_qkv_bias_gradient():
    """The dbias column sums of the N <= 208 backward also come out of the bias variant."""
    import b200timm.layers as L
    ops = _ops()
    B, N, H, hd = 2, 197, 2, 64
    g = torch.Generator(device=DEV).manual_seed(5)
    qkv = torch.randn(B, N, 3 * H * hd, device=DEV, generator=g).bfloat16()
    dout = torch.randn(B, N, H * hd, device=DEV, generator=g).bfloat16()
    operand = L.attn_mask_operand(None, True, N, B, DEV)
    out, lse = ops.attention_fwd(qkv, B, N, H, hd, 0.125, kv_len=operand, need_lse=True)
    db = torch.zeros(3 * H * hd, device=DEV)
    dqkv = ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd, 0.125, kv_len=operand, dbias=db)
    _close(db, dqkv.float().view(B * N, -1).sum(0), 5e-2, 1e-2, "dbias under a causal mask")


# ------------------------------------------------------------------------------------------------ folded LayerNorm
@pytest.mark.parametrize("M,D,N,gelu", [(1576, 768, 2304, False), (1576, 768, 3072, True), (300, 192, 576, False), (1000, 384, 1536, True),
                                       (1576, 768, 3072, "gelu_tanh"), (788, 384, 1536, "silu")])
def test_gemm_folded_layernorm(M, D, N, gelu):
    """b200_gemm_desc.ln_*: the residual GEMM leaves per-row partial sums of the bf16 values it stores; the next Linear runs on
    the un-normalised rows against weight * gamma and finishes LayerNorm in its epilogue.  Against fp32 LayerNorm + Linear
    (+ exact GELU) of timm/layers/norm.py:85-90, mlp.py:48-49 on the same bf16 inputs."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(M + N)
    a = torch.randn(M, D, device=DEV, generator=g).bfloat16()
    wp = (torch.randn(D, D, device=DEV, generator=g) * D ** -0.5).bfloat16()
    bp = torch.randn(D, device=DEV, generator=g).bfloat16()
    # a residual stream with a per-row offset and a few large channels (what a ViT's looks like)
    res = (torch.randn(M, D, device=DEV, generator=g) + torch.randn(M, 1, device=DEV, generator=g) * 0.5)
    res[:, :4] *= 12.0
    res = res.bfloat16()
    stats = ops.ln_stats_buffer(M, D, DEV)
    x = ops.linear_fwd(a, wp, bp, epilogue=ops.EPI_BIAS_RESIDUAL, aux=res, ln_stats_out=stats)
    x_plain = ops.linear_fwd(a, wp, bp, epilogue=ops.EPI_BIAS_RESIDUAL, aux=res)
    assert torch.equal(x, x_plain)                      # the producer stores the same values with and without statistics
    xf = x.float()
    s = stats[:-1].sum(0)
    _close(s[:, 0], xf.sum(1), 2e-2, 1e-4, "row sums")
    _close(s[:, 1], (xf * xf).sum(1), 1e-1, 1e-4, "row sums of squares")
    gamma = (1.0 + 0.2 * torch.randn(D, device=DEV, generator=g)).bfloat16()
    beta = (0.1 * torch.randn(D, device=DEV, generator=g)).bfloat16()
    w = (torch.randn(N, D, device=DEV, generator=g) * D ** -0.5).bfloat16()
    b = torch.randn(N, device=DEV, generator=g).bfloat16()
    eps = 1e-6
    wg = (w.float() * gamma.float()[None]).bfloat16().contiguous()
    c1 = wg.float().sum(1).contiguous()
    c2 = (w.float() @ beta.float() + b.float()).contiguous()
    epi = ops.EPI_BIAS_GELU if gelu else ops.EPI_BIAS
    act = ops.ACTS[gelu] if isinstance(gelu, str) else 0          # tanh GELU / SiLU: the activation-at-run-time consumer
    y = ops.linear_fwd(x, wg, c2, epilogue=epi, act=act, ln_in=(stats, c1, eps))
    ref = torch.nn.functional.layer_norm(xf, (D,), gamma.float(), beta.float(), eps) @ w.float().t() + b.float()
    if gelu == "gelu_tanh":
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    elif gelu == "silu":
        ref = torch.nn.functional.silu(ref)
    elif gelu:
        ref = torch.nn.functional.gelu(ref)
    _close(y, ref, 3e-2, 2e-2, f"folded LayerNorm -> Linear M={M} N={N} gelu={gelu}")
    # and against the unfolded kernel sequence (LayerNorm kernel, bf16 h, GEMM): the two differ by bf16 roundings only
    h = ops.layernorm_fwd(x, gamma, beta, eps)[0]
    y2 = ops.linear_fwd(h, w, b, epilogue=epi, act=act)
    _close(y, y2, 4e-2, 2e-2, "folded vs unfolded")


def test_gemm_folded_layernorm_rejects_unsupported_shapes():
    ops = _ops()
    a = torch.zeros(64, 256, device=DEV, dtype=torch.bfloat16)
    w = torch.zeros(256, 256, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(ops.B200Error):       # M < 256: no folded instantiation; callers keep the LayerNorm kernel
        ops.linear_fwd(a, w, None, ln_stats_out=ops.ln_stats_buffer(64, 256, DEV))


def test_vit_forward_with_and_without_folded_layernorm():
    """The whole inference forward with the LayerNorms folded into the GEMMs (default) and with LayerNorm kernels
    (functional._LN_FOLD = False): same logits within bf16 rounding, both within tolerance of the CPU oracle."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import vit_oracle as O
    from b200timm import functional as BF
    from b200timm.vision_transformer import VisionTransformer
    ops = _ops()
    cfg = dict(img_size=224, patch_size=16, embed_dim=256, depth=3, num_heads=4, init_values=0.3)
    sd = {k: v.bfloat16().float() for k, v in O.random_state_dict(cfg, seed=21, num_classes=40).items()}
    model = VisionTransformer(num_classes=40, **cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().to(torch.bfloat16).eval()
    img = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(1)).bfloat16()
    ref = O.vit_forward(img.float(), sd, dict(cfg, num_prefix_tokens=1))
    outs = {}
    for fold in (True, False):
        BF._LN_FOLD = fold
        try:
            ops.reset_counters()
            with torch.no_grad():
                outs[fold] = model(img.cuda()).float().cpu()
            launches = ops.launch_count()
        finally:
            BF._LN_FOLD = True
        outs[(fold, "launches")] = launches
        assert (outs[fold] - ref).abs().max().item() <= 2e-2 + 2e-2 * ref.abs().max().item()
    # 5 of the 6 LayerNorm launches are gone (the first block's norm1 stays); each became a statistics finalisation
    assert outs[(True, "launches")] == outs[(False, "launches")], outs
    assert (outs[True] - outs[False]).abs().max().item() <= 1e-2 + 1e-2 * ref.abs().max().item()


def test_wgrad_row_chunks_match_one_launch(monkeypatch):
    """ops.linear_wgrad splits a badly tiling weight gradient (ViT-L's fused qkv projection: 3072 x 1024, 48 tiles for 74 CTA
    pairs) into launches over row chunks; same result as the single split-K launch and as fp32 matmul."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(9)
    assert ops._wgrad_parts(3072, 1024, 36928) == [(0, 1536, 3), (1536, 1536, 3)]
    assert ops._wgrad_parts(2304, 768, 50432) is None and ops._wgrad_parts(768, 768, 50432) is None
    M, N, K = 1576, 3072, 1024
    dy = torch.randn(M, N, device=DEV, generator=g).bfloat16()
    x = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    monkeypatch.setattr(ops, "_wgrad_parts", lambda n, k, m: [(0, 1536, 3), (1536, 1536, 3)])     # the plan of the full-size call
    got = ops.linear_wgrad(dy, x)
    one = torch.zeros(N, K, device=DEV)
    ops.gemm(dy, x, N, K, M, a_mn=True, b_mn=True, out=one, split_k=2)
    ref = dy.float().t() @ x.float()
    _close(got, ref, 5e-2, 1e-3, "wgrad in row chunks")
    _close(got, one, 2e-3, 1e-4, "row chunks vs one launch")
    acc = torch.ones(N, K, device=DEV)               # accumulates into a non-zero buffer like the single launch
    ops.linear_wgrad(dy, x, out=acc)
    _close(acc, ref + 1.0, 5e-2, 1e-3, "wgrad accumulation")
