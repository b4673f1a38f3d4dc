"""Drop-in plugin for an installed timm: rebinds the ``forward`` of timm's own hot-path classes to the
libb200timm kernels, leaving constructors, parameters, ``state_dict()`` keys, the model registry,
``create_model()``, ``benchmark.py`` / ``train.py`` / ``validate.py`` untouched (seam 3).

    import b200timm.plugin; b200timm.plugin.install()         # or: python -m b200timm.run benchmark.py ...
    train.py --device-modules b200timm.plugin ...              # timm's own import hook (train.py:186,464-466)
    TORCH_BACKEND=b200timm.plugin pytest tests/test_models.py  # timm's test-suite hook (tests/test_models.py:36-39)

Dispatch rule per call: CUDA tensor + bf16 activations and weights + a module configuration the kernels cover
-> B200 kernels; everything else (CPU tensors such as BASELINE configs[0], fp32, scripting / export, other
model families) keeps timm's original forward, which *is* the reference path for those cases.  If
libb200timm.so cannot be loaded, install() raises -- it never silently leaves the PyTorch path in place.
Set B200TIMM_STRICT=1 to raise instead of deferring to timm whenever a CUDA bf16 call is not covered.

B200TIMM_FP32=bf16 (opt-in, the analogue of torch.backends.cuda.matmul.allow_tf32): fp32 CUDA calls are computed by the
bf16 kernels as well -- the activations are cast to bf16 at the rebound forward, fp32 parameters are read the way
torch.autocast(bfloat16) reads them (functional._w), the result is cast back.  This is how timm's own test-suite
(tests/test_models.py builds fp32 models) runs on the kernels through its TORCH_BACKEND hook.
"""
import os

import torch

from . import _lib
from . import functional as BF
from ._lib import B200Error

_ORIG = {}
_STRICT = os.environ.get("B200TIMM_STRICT", "0") == "1"
_FP32_AS_BF16 = os.environ.get("B200TIMM_FP32", "") == "bf16"

# Dispatch accounting: how many calls of each rebound forward ran on the kernels ("fast"), kept timm's forward because the
# input is not the plugin's business (CPU / fp32 / scripting: "passthrough") or were DEFERRED to timm although the input was
# CUDA bf16, because the kernels do not cover the module's configuration.  A deferral is reported once per (forward, reason)
# through `warnings` -- a benchmark that silently ran cuBLAS / SDPA would not be the kernels' number -- and raises under
# B200TIMM_STRICT=1.  stats() returns the table; B200TIMM_REPORT=<path> writes it as JSON at interpreter exit.
_STATS = {"fast": {}, "passthrough": {}, "deferred": {}}
_WARNED = set()


def _note_deferral(name, err):
    reason = str(err)
    d = _STATS["deferred"].setdefault(name, {})
    d[reason] = d.get(reason, 0) + 1
    if (name, reason) not in _WARNED:
        _WARNED.add((name, reason))
        import warnings
        warnings.warn(f"b200timm: {name} deferred to timm's PyTorch forward ({reason}); set B200TIMM_STRICT=1 to raise instead",
                      RuntimeWarning, stacklevel=3)


def stats():
    """{'fast': {forward: calls}, 'passthrough': {...}, 'deferred': {forward: {reason: calls}}, 'kernel_launches': n}"""
    from . import ops
    return {"fast": dict(_STATS["fast"]), "passthrough": dict(_STATS["passthrough"]),
            "deferred": {k: dict(v) for k, v in _STATS["deferred"].items()}, "kernel_launches": ops.launch_count()}


def reset_stats():
    for v in _STATS.values():
        v.clear()


def _write_report():
    path = os.environ.get("B200TIMM_REPORT")
    if path:
        import json
        try:
            with open(path, "w") as f:
                json.dump(stats(), f)
        except OSError:
            pass


def _autocast_bf16():
    return torch.is_autocast_enabled() and torch.get_autocast_gpu_dtype() == torch.bfloat16


def _fast(x):
    """CUDA bf16 activations -- or, under torch.autocast(bfloat16) (train.py --amp), fp32 IMAGES entering a rebound
    model / patch embedding: the kernels read fp32 pixels directly and fp32 master weights are cast as autocast casts
    them (functional._w)."""
    if not (isinstance(x, torch.Tensor) and x.is_cuda) or torch.jit.is_scripting() or torch.jit.is_tracing():
        return False
    if torch.is_grad_enabled() and _GRAD_UNSUPPORTED:
        return False
    if x.dtype == torch.uint8:     # raw loader batches: normalisation folded into the patch gather (b200timm.data)
        from . import ops
        return x.dim() == 4 and ops.input_norm() is not None
    return x.dtype == torch.bfloat16 or (x.dtype == torch.float32 and x.dim() == 4 and _autocast_bf16())


_GRAD_UNSUPPORTED = False  # the training path (b200timm.autograd) covers grad-enabled calls; uncovered configs raise B200Error -> timm


def _cast_tree(out, dtype):
    if isinstance(out, torch.Tensor):
        return out.to(dtype) if out.is_floating_point() else out
    if isinstance(out, (tuple, list)):
        return type(out)(_cast_tree(o, dtype) for o in out)
    return out


def _fp32_as_bf16(x):
    return (_FP32_AS_BF16 and isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32
            and not torch.is_autocast_enabled() and not (torch.jit.is_scripting() or torch.jit.is_tracing()))


def _dispatch(name, fast_fn):
    orig = _ORIG[name]

    def forward(self, x, *args, **kwargs):
        if _fp32_as_bf16(x):
            try:
                out = _cast_tree(fast_fn(self, x.to(torch.bfloat16), *args, **kwargs), torch.float32)
                _STATS["fast"][name] = _STATS["fast"].get(name, 0) + 1
                return out
            except B200Error as e:
                if _STRICT:
                    raise
                _note_deferral(name, e)
                return orig(self, x, *args, **kwargs)
        if _fast(x):
            try:
                out = fast_fn(self, x, *args, **kwargs)
                _STATS["fast"][name] = _STATS["fast"].get(name, 0) + 1
                return out
            except B200Error as e:
                if _STRICT:
                    raise
                _note_deferral(name, e)
        else:
            _STATS["passthrough"][name] = _STATS["passthrough"].get(name, 0) + 1
        return orig(self, x, *args, **kwargs)

    forward.__name__ = "forward"
    forward.__doc__ = f"b200timm dispatcher over {name}"
    forward._b200_original = orig
    return forward


def _mask(attn_mask, is_causal, x, n=None):
    """(attn_mask, is_causal) of the reference signatures -> the kernels' mask operand (layers.attn_mask_operand)."""
    if attn_mask is None and not is_causal:
        return None
    from .layers import attn_mask_operand
    return attn_mask_operand(attn_mask, is_causal, x.shape[1] if n is None else n, x.shape[0], x.device)


def _block(self, x, attn_mask=None, is_causal=False):
    return BF.block_forward(self, x, _mask(attn_mask, is_causal, x))


def _attention(self, x, attn_mask=None, is_causal=False):
    return BF.attention_forward(self, x, _mask(attn_mask, is_causal, x))


def _attention_rope(self, x, rope=None, attn_mask=None, is_causal=False):
    return BF.attention_rope_forward(self, x, rope, _mask(attn_mask, is_causal, x))


def _patch_embed(self, x):
    """PatchEmbed.forward (patch_embed.py:117-142) in every output format: the GEMM emits token-major rows [B, G, D], which IS
    the NLC result (flatten=True) and, viewed [B, gh, gw, D], the NHWC one; NCHW is a permuted view of that."""
    y = BF.patch_embed_forward(self, x)                      # [B, G, D]
    if self.flatten:
        return y
    ph, pw = self.patch_size
    H, W = x.shape[-2], x.shape[-1]
    if getattr(self, "dynamic_img_pad", False):
        H, W = H + (ph - H % ph) % ph, W + (pw - W % pw) % pw
    y = y.view(y.shape[0], H // ph, W // pw, y.shape[-1])
    fmt = getattr(self.output_fmt, "value", str(self.output_fmt))
    if fmt == "NHWC":
        return y
    if fmt == "NCHW":
        return y.permute(0, 3, 1, 2)
    raise B200Error(f"PatchEmbed output format {fmt} is not covered")


def _vit_tokens(self, x):
    """Sequence length the blocks will see for image batch `x` (dynamic_img_size models follow the input)."""
    ph, pw = self.patch_embed.patch_size
    H, W = x.shape[-2], x.shape[-1]
    if getattr(self.patch_embed, "dynamic_img_pad", False):
        H, W = H + (ph - H % ph) % ph, W + (pw - W % pw) % pw
    return (H // ph) * (W // pw) + self.num_prefix_tokens


def _vit_features(self, x, attn_mask=None, is_causal=False):
    return BF.vit_forward_features(self, x, _mask(attn_mask, is_causal, x, _vit_tokens(self, x)))


def _vit_forward(self, x, attn_mask=None, is_causal=False):
    return BF.vit_forward(self, x, _mask(attn_mask, is_causal, x, _vit_tokens(self, x)))


def _convnext_forward(self, x):
    from . import convnext as CN
    rows, B, H, W = CN.features_nhwc(self, x)
    return CN.head_forward_nhwc(self.head, rows, B, H * W)


def _convnext_features(self, x):
    from . import convnext as CN
    rows, B, H, W = CN.features_nhwc(self, x)
    return rows.view(B, H, W, -1).permute(0, 3, 1, 2)


def _convnext_head(self, x, pre_logits=False):
    from . import convnext as CN
    B, C, H, W = x.shape
    rows = x.permute(0, 2, 3, 1).contiguous().view(B * H * W, C)
    return CN.head_forward_nhwc(self.head, rows, B, H * W, pre_logits)


def _convnext_block(self, x):
    from . import convnext as CN
    B, C, H, W = x.shape
    rows = x.permute(0, 2, 3, 1).contiguous().view(B * H * W, C)
    return CN.block_forward_nhwc(self, rows, B, H, W).view(B, H, W, C).permute(0, 3, 1, 2)


def install():
    """Rebind timm's hot-path forwards.  Idempotent.  Raises if timm or libb200timm.so is unavailable."""
    _lib.load()  # fail loudly: no silent PyTorch path when the CUDA library is missing
    import timm  # noqa: F401
    if _ORIG:
        return
    table = _table()
    for name, cls, attr, fn in table:
        _ORIG[name] = getattr(cls, attr)
        setattr(cls, attr, _dispatch(name, fn))
        _ORIG[name + "#cls"] = (cls, attr)
    _install_training_tail()
    if os.environ.get("B200TIMM_REPORT") and not getattr(install, "_atexit", False):
        import atexit
        atexit.register(_write_report)
        install._atexit = True


def _ema_dispatch(orig, buffers):
    def apply(self, model, decay):
        first = next(iter(self.module.parameters()), None)
        if self.foreach and first is not None and first.is_cuda:
            try:
                from .ema import fast_update
                return fast_update(self, model, decay, buffers=buffers)     # validates every tensor before launching
            except B200Error as e:
                if _STRICT:
                    raise
                _note_deferral("timm.utils.ModelEmaV3.apply_update_", e)
        return orig(self, model, decay)
    apply._b200_original = orig
    return apply


def _install_training_tail():
    """The steps after the backward: the libb200timm optimizers join timm's registry as b200sgd / b200momentum /
    b200adamw / b200adam (an addition -- existing names keep their torch.optim classes), and ModelEmaV3's interpolation
    (timm/utils/model_ema.py:213-252) runs on b200_ema_update for CUDA models."""
    from timm.utils.model_ema import ModelEmaV3
    from .optim import register_with_timm
    register_with_timm()
    for attr, buffers in (("apply_update_", True), ("apply_update_no_buffers_", False)):
        name = f"timm.utils.ModelEmaV3.{attr}"
        _ORIG[name] = getattr(ModelEmaV3, attr)
        _ORIG[name + "#cls"] = (ModelEmaV3, attr)
        setattr(ModelEmaV3, attr, _ema_dispatch(_ORIG[name], buffers))


def _table():
    """(qualified name, class, attribute, fast function) of every forward the plugin covers."""
    from timm.layers.attention import Attention, AttentionRope
    from timm.layers.attention_pool import AttentionPoolLatent, AttentionPoolPrr
    from timm.layers.mlp import GluMlp, Mlp, SwiGLU
    from timm.layers.norm import LayerNorm, RmsNorm
    from timm.layers.patch_embed import PatchEmbed
    from timm.models.convnext import ConvNeXt, ConvNeXtBlock
    from timm.models.eva import EvaAttention
    from timm.models.vision_transformer import Block, ParallelScalingBlock, VisionTransformer
    return [
        ("timm.models.vision_transformer.ParallelScalingBlock.forward", ParallelScalingBlock, "forward", _block),
        ("timm.layers.LayerNorm.forward", LayerNorm, "forward", lambda self, x: BF.layernorm(self, x)),
        ("timm.layers.RmsNorm.forward", RmsNorm, "forward", lambda self, x: BF.layernorm(self, x)),
        ("timm.layers.Mlp.forward", Mlp, "forward", lambda self, x: BF.mlp_forward(self, x)),
        ("timm.layers.GluMlp.forward", GluMlp, "forward", lambda self, x: BF.glu_mlp_forward(self, x)),
        ("timm.layers.SwiGLU.forward", SwiGLU, "forward", lambda self, x: BF.glu_mlp_forward(self, x)),
        ("timm.layers.AttentionPoolLatent.forward", AttentionPoolLatent, "forward",
         lambda self, x, attn_mask=None: BF.attention_pool_forward(self, x) if attn_mask is None
         else (_ for _ in ()).throw(B200Error("masked attention pooling"))),
        ("timm.layers.Attention.forward", Attention, "forward", _attention),
        ("timm.layers.AttentionRope.forward", AttentionRope, "forward", _attention_rope),
        ("timm.models.eva.EvaAttention.forward", EvaAttention, "forward", _attention_rope),
        ("timm.layers.PatchEmbed.forward", PatchEmbed, "forward", _patch_embed),
        ("timm.layers.AttentionPoolPrr.forward", AttentionPoolPrr, "forward",
         lambda self, x: BF.attention_pool_prr_forward(self, x)),
        ("timm.models.vision_transformer.Block.forward", Block, "forward", _block),
        ("timm.models.vision_transformer.VisionTransformer.forward_features", VisionTransformer, "forward_features",
         _vit_features),
        ("timm.models.vision_transformer.VisionTransformer.forward_head", VisionTransformer, "forward_head",
         lambda self, x, pre_logits=False: BF.vit_forward_head(self, x, pre_logits)),
        ("timm.models.vision_transformer.VisionTransformer.forward", VisionTransformer, "forward", _vit_forward),
        ("timm.models.convnext.ConvNeXtBlock.forward", ConvNeXtBlock, "forward", _convnext_block),
        ("timm.models.convnext.ConvNeXt.forward_features", ConvNeXt, "forward_features", _convnext_features),
        ("timm.models.convnext.ConvNeXt.forward_head", ConvNeXt, "forward_head", _convnext_head),
        ("timm.models.convnext.ConvNeXt.forward", ConvNeXt, "forward", _convnext_forward),
    ]


def convert_to_b200(model):
    """Post-construction conversion of ONE model (SURVEY 8b seam 2; the precedent is timm's convert_sync_batchnorm,
    timm/layers/norm_act.py:167-208): every covered sub-module of `model` gets the B200 dispatcher bound on the
    INSTANCE -- same Parameter objects, same state_dict keys, nothing re-parented, other models in the process
    untouched.  Equivalent to install() for this model only; returns the model."""
    import types
    _lib.load()
    table = _table()
    for mod in model.modules():
        for name, cls, attr, fn in table:
            if type(mod) is cls:
                orig = _ORIG.get(name, getattr(cls, attr))

                def make(orig=orig, fn=fn, name=name):
                    def forward(self, x, *args, **kwargs):
                        if _fast(x):
                            try:
                                out = fn(self, x, *args, **kwargs)
                                _STATS["fast"][name] = _STATS["fast"].get(name, 0) + 1
                                return out
                            except B200Error as e:
                                if _STRICT:
                                    raise
                                _note_deferral(name, e)
                        else:
                            _STATS["passthrough"][name] = _STATS["passthrough"].get(name, 0) + 1
                        return orig(self, x, *args, **kwargs)
                    forward._b200_original = orig
                    return forward
                setattr(mod, attr, types.MethodType(make(), mod))
    return model


def uninstall():
    """Restore timm's original forwards."""
    for name in [k for k in _ORIG if not k.endswith("#cls")]:
        cls, attr = _ORIG[name + "#cls"]
        setattr(cls, attr, _ORIG[name])
    _ORIG.clear()


def installed():
    return bool(_ORIG)


if os.environ.get("B200TIMM_AUTOINSTALL", "1") == "1" and __name__ != "__main__":
    # `--device-modules b200timm.plugin` / TORCH_BACKEND=b200timm.plugin only import the module:
    # importing it is the installation (mirrors how out-of-tree torch backends register themselves).
    try:
        install()
    except ImportError:
        pass  # timm itself not importable: nothing to rebind (the mirror classes in b200timm.* still work)
