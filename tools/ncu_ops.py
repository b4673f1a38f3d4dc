"""One hot-path op at its ViT-B / ConvNeXt-B / ViT-L production shape, launched three times -- the workload
`tools/gpu_ncu_all.sh` profiles kernel by kernel with `ncu --set full` (one capture per kernel, third launch).
    python tools/ncu_ops.py ln_fwd | ln_bwd | attn_fwd | attn_bwd | attn_long | dwconv | dwconv_wgrad | rope | patchify |
                            assemble | colsum | gemm_dgelu | gemm_wgrad"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))
from b200timm import ops  # noqa: E402

M, D, B, N, H = 50432, 768, 256, 197, 12


def r(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).bfloat16()


def main(op):
    if op == "ln_fwd":
        x, w, b = r(M, D), r(D), r(D)
        f = lambda: ops.layernorm_fwd(x, w, b, 1e-6)                                     # noqa: E731
    elif op == "ln_bwd":
        x, w, b, dy, dres = r(M, D), r(D), r(D), r(M, D), r(M, D)
        _, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-6, save_stats=True)
        f = lambda: ops.layernorm_bwd(dy, x, w, mean, rstd, dres=dres)                   # noqa: E731
    elif op in ("attn_fwd", "attn_bwd"):
        qkv = r(B * N, 3 * H * 64)
        o, lse = ops.attention_fwd(qkv, B, N, H, 64, 0.125, need_lse=True)
        do = r(B, N, H * 64)
        f = (lambda: ops.attention_fwd(qkv, B, N, H, 64, 0.125)) if op == "attn_fwd" else \
            (lambda: ops.attention_bwd(qkv, o, do, lse, B, N, H, 64, 0.125))
    elif op == "attn_long":
        qkv = r(64 * 577, 3 * 16 * 64)
        f = lambda: ops.attention_fwd(qkv, 64, 577, 16, 64, 0.125)                       # noqa: E731
    elif op in ("dwconv", "dwconv_wgrad"):
        x, dy = r(256, 14, 14, 512), r(256, 14, 14, 512)
        wt = ops.dwconv_weight_t(r(512, 1, 7, 7, scale=0.1))
        bias = r(512)
        f = (lambda: ops.dwconv_nhwc(x, wt, bias)) if op == "dwconv" else (lambda: ops.dwconv_nhwc_wgrad(x, dy, 7))
    elif op == "rope":
        qkv = r(B * N, 3 * H * 64)
        emb = torch.rand(N - 1, 128, device="cuda")
        f = lambda: ops.rope_qk(qkv, emb, B, N, H, 1)                                    # noqa: E731
    elif op == "patchify":
        img = r(B, 3, 224, 224)
        f = lambda: ops.patchify(img, 16)                                                # noqa: E731
    elif op == "assemble":
        proj, prefix, pos = r(B * 196, D), r(1, D), r(N, D)
        f = lambda: ops.assemble_tokens(proj, prefix, pos, B, 196, 1, D)                 # noqa: E731
    elif op == "colsum":
        x = r(M, 3072)
        f = lambda: ops.colsum(x)                                                        # noqa: E731
    elif op in ("qk_norm", "qk_norm_bwd"):
        qkv, dq = r(B * N, 3 * H * 64), r(B * N, 3 * H * 64)
        qw, qb, kw, kb = r(64), r(64), r(64), r(64)
        raw, mean, rstd = ops.qk_norm_fwd(qkv.clone(), B * N, H, qw, qb, kw, kb, 1e-6, save=True)
        f = (lambda: ops.qk_norm_fwd(qkv, B * N, H, qw, qb, kw, kb, 1e-6, save=True)) if op == "qk_norm" else \
            (lambda: ops.qk_norm_bwd(dq, raw, mean, rstd, qw, kw, B * N, H))
    elif op == "glu":
        h = r(M, 2 * 2048)
        f = lambda: ops.glu_fwd(h, True, "silu")                                         # noqa: E731
    elif op in ("grn", "grn_bwd"):                     # ConvNeXt-V2-B stage 3: [256, 14*14, 2048]
        pre, w, b, d = r(256 * 196, 2048), r(2048, scale=0.3), r(2048, scale=0.1), r(256 * 196, 2048)
        act = torch.nn.functional.gelu(pre.float()).bfloat16()
        gsq, scale = ops.grn_fwd(act.clone(), w, b, 256, 196)
        f = (lambda: ops.grn_fwd(act, w, b, 256, 196)) if op == "grn" else \
            (lambda: ops.grn_bwd(d, pre, gsq, scale, w, 256, 196))
    elif op == "patchify_u8":
        img = torch.randint(0, 256, (B, 3, 224, 224), dtype=torch.uint8, device="cuda")
        ops.set_input_norm((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
        f = lambda: ops.patchify(img, 16)                                                # noqa: E731
    elif op == "ema":
        from b200timm.ema import lerp_
        es = [r(3072, 768) for _ in range(24)]
        ms = [r(3072, 768) for _ in range(24)]
        f = lambda: lerp_(es, ms, 1e-4)                                                  # noqa: E731
    elif op == "gemm_dgelu":
        dy, w, pre = r(M, D), r(D, 3072, scale=0.05), r(M, 3072)
        f = lambda: ops.linear_dgrad(dy, w, epilogue=ops.EPI_DGELU, aux=pre)             # noqa: E731
    elif op == "gemm_wgrad":
        dy, x = r(M, 3072), r(M, D)
        f = lambda: ops.linear_wgrad(dy, x)                                              # noqa: E731
    else:
        raise SystemExit(f"unknown op {op}")
    for _ in range(3):
        f()
    torch.cuda.synchronize()


if __name__ == "__main__":
    main(sys.argv[1])
