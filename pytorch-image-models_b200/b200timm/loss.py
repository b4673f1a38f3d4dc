"""Classification losses on libb200timm: mirrors of nn.CrossEntropyLoss (what train.py:855 / benchmark.py:380 build),
timm.loss.LabelSmoothingCrossEntropy and timm.loss.SoftTargetCrossEntropy (timm/loss/cross_entropy.py:11-36) -- same
constructor arguments and forward(x, target) -> scalar mean loss -- running b200_cross_entropy_fwd / _bwd: one
row-parallel pass forward (+ a one-block mean), one pass backward that recomputes softmax from the saved log-sum-exp.
Logits may be bf16 (pure-bf16 training, benchmark.py --precision bfloat16) or fp32; the loss is an fp32 scalar and the
logits gradient has the logits' dtype.  No CPU path."""
import torch
from torch import nn

from . import _lib
from ._lib import B200Error, DTYPE_BF16, DTYPE_F32
from .ops import _count, _stream

_DT = {torch.bfloat16: DTYPE_BF16, torch.float32: DTYPE_F32}


class _CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, smoothing, ignore_index=-100):
        if not logits.is_cuda or not target.is_cuda:
            raise B200Error("libb200timm kernels need CUDA tensors (there is no CPU path)")
        if logits.dim() != 2 or logits.dtype not in _DT:
            raise B200Error(f"cross entropy: logits must be [B, C] bf16 / fp32, got {tuple(logits.shape)} {logits.dtype}")
        B, C = logits.shape
        x = logits if logits.is_contiguous() else logits.contiguous()
        soft = target.dim() == 2
        if soft:
            if target.shape != logits.shape or target.dtype not in _DT:
                raise B200Error("cross entropy: soft targets must be [B, C] bf16 / fp32")
            if smoothing != 0.0:
                raise B200Error("cross entropy: label smoothing with soft targets is not supported")
        elif target.dim() != 1 or target.shape[0] != B or target.dtype != torch.int64:
            raise B200Error("cross entropy: class-index targets must be int64 [B]")
        t = target if target.is_contiguous() else target.contiguous()
        buf = torch.empty(3 * B + 2, dtype=torch.float32, device=logits.device)     # loss rows | lse | tsum | mean, denominator
        p = buf.data_ptr()
        rc = _lib.load().b200_cross_entropy_fwd(x.data_ptr(), _DT[x.dtype], None if soft else t.data_ptr(),
                                                t.data_ptr() if soft else None, _DT[t.dtype] if soft else DTYPE_F32,
                                                float(smoothing), int(ignore_index), p, p + 4 * B, p + 8 * B, p + 12 * B, B, C, _stream())
        _lib.check(rc, "b200_cross_entropy_fwd")
        _count(2)
        ctx.save_for_backward(x, t, buf)
        ctx.smoothing, ctx.soft = float(smoothing), soft
        return buf[3 * B]

    @staticmethod
    def backward(ctx, grad_out):
        x, t, buf = ctx.saved_tensors
        B, C = x.shape
        g = grad_out.to(torch.float32).contiguous()
        dx = torch.empty_like(x)
        p = buf.data_ptr()
        soft = ctx.soft
        rc = _lib.load().b200_cross_entropy_bwd(x.data_ptr(), _DT[x.dtype], None if soft else t.data_ptr(),
                                                t.data_ptr() if soft else None, _DT[t.dtype] if soft else DTYPE_F32,
                                                ctx.smoothing, p + 4 * B, p + 8 * B, g.data_ptr(), p + 12 * B + 4, dx.data_ptr(), B, C,
                                                _stream())
        _lib.check(rc, "b200_cross_entropy_bwd")
        _count()
        return dx, None, None, None


def cross_entropy(logits, target, label_smoothing: float = 0.0, ignore_index: int = -100):
    """F.cross_entropy(logits, target, ignore_index=..., label_smoothing=...) with 'mean' reduction; `target` int64 [B] or a
    [B, C] distribution."""
    return _CrossEntropyFn.apply(logits, target, label_smoothing, ignore_index)


class CrossEntropyLoss(nn.Module):
    """nn.CrossEntropyLoss() as train.py:855 / benchmark.py:380 use it (mean reduction, optional label_smoothing).
    Rows whose class index equals `ignore_index` contribute zero loss and zero gradient and the mean divides by the number
    of remaining rows, as torch does (aten/src/ATen/native/LossNLL.cpp; all rows ignored -> NaN, as in torch)."""

    def __init__(self, weight=None, size_average=None, ignore_index=-100, reduce=None, reduction="mean", label_smoothing=0.0):
        super().__init__()
        if weight is not None or reduction != "mean" or size_average is not None or reduce is not None:
            raise B200Error("CrossEntropyLoss: class weights / non-mean reductions are not supported")
        self.ignore_index, self.label_smoothing = ignore_index, label_smoothing

    def forward(self, x, target):
        return _CrossEntropyFn.apply(x, target, self.label_smoothing, self.ignore_index)


class LabelSmoothingCrossEntropy(nn.Module):
    """timm.loss.LabelSmoothingCrossEntropy (cross_entropy.py:11-26)."""

    def __init__(self, smoothing=0.1):
        super().__init__()
        assert smoothing < 1.0
        self.smoothing = smoothing
        self.confidence = 1. - smoothing

    def forward(self, x, target):
        return _CrossEntropyFn.apply(x, target, self.smoothing)


class SoftTargetCrossEntropy(nn.Module):
    """timm.loss.SoftTargetCrossEntropy (cross_entropy.py:29-36): targets are distributions (mixup / cutmix)."""

    def forward(self, x, target):
        if target.dim() != 2:
            raise B200Error("SoftTargetCrossEntropy: target must be a [B, C] distribution")
        return _CrossEntropyFn.apply(x, target, 0.0)
