"""Weight EMA on libb200timm: mirror of timm.utils.ModelEmaV3 (timm/utils/model_ema.py:118-262) -- same constructor
arguments, attributes (`module`, `decay`, ...), `get_decay` schedule, `update(model, step)`, `set`, `forward` -- with
the interpolation of every floating-point state_dict entry done by ONE b200_ema_update launch (torch runs
`_foreach_lerp_`, :225-226).  `ema_dtype=torch.float32` keeps the averaged weights in fp32 while the model trains in
bf16 (with decay 0.9999 a bf16 average cannot absorb the update: 1e-4 * (w - ema) is below its rounding step).
`fast_update(ema, model, decay)` is the function the plugin binds over ModelEmaV3.apply_update_ so train.py's
`--model-ema` runs on the kernel unchanged."""
from copy import deepcopy
from typing import Optional

import torch
from torch import nn

from . import _lib
from ._lib import B200Error, DTYPE_BF16, DTYPE_F32
from .ops import _count, _stream, note_params_written
from .optim import _table

_DT = {torch.bfloat16: DTYPE_BF16, torch.float32: DTYPE_F32}


def lerp_(ema_tensors, model_tensors, weight: float):
    """ema[i].lerp_(model[i], weight) for CUDA bf16 / fp32 tensors, grouped by dtype pair, one launch per pair."""
    groups = {}
    for e, m in zip(ema_tensors, model_tensors):
        if not e.is_cuda or not m.is_cuda:
            raise B200Error("libb200timm kernels need CUDA tensors (there is no CPU path)")
        if e.dtype not in _DT or m.dtype not in _DT or e.shape != m.shape or not e.is_contiguous():
            raise B200Error(f"ema lerp: unsupported tensors {e.dtype}{tuple(e.shape)} / {m.dtype}{tuple(m.shape)}")
        if e.numel():
            groups.setdefault((e.dtype, m.dtype), []).append((e, m if m.is_contiguous() else m.contiguous(), None, None))
    lib = _lib.load()
    for (ed, md), rows in groups.items():
        rc = lib.b200_ema_update(_table(rows), len(rows), _DT[ed], _DT[md], float(weight), _stream())
        _lib.check(rc, "b200_ema_update")
        _count()
    note_params_written([r[0] for rows in groups.values() for r in rows])


def fast_update(ema, model, decay: float, buffers: bool = True):
    """ModelEmaV3.apply_update_ (buffers=True, :213-235) / apply_update_no_buffers_ (:237-252)."""
    if buffers:
        pairs = list(zip(ema.module.state_dict().values(), model.state_dict().values()))
    else:
        pairs = list(zip(ema.module.parameters(), model.parameters()))
    lerp_([e for e, m in pairs if e.is_floating_point()], [m for e, m in pairs if e.is_floating_point()], 1.0 - decay)
    for e, m in pairs:
        if not e.is_floating_point():
            e.copy_(m)
    if not buffers:
        for e, m in zip(ema.module.buffers(), model.buffers()):
            e.copy_(m)


class ModelEma(nn.Module):
    def __init__(self, model, decay: float = 0.9999, min_decay: float = 0.0, update_after_step: int = 0,
                 use_warmup: bool = False, warmup_gamma: float = 1.0, warmup_power: float = 2 / 3,
                 device: Optional[torch.device] = None, foreach: bool = True, exclude_buffers: bool = False,
                 ema_dtype: Optional[torch.dtype] = None):
        super().__init__()
        if device is not None and device != next(model.parameters()).device:
            raise B200Error("ModelEma: the average lives on the model's device")
        self.module = deepcopy(model)
        self.module.eval()
        if ema_dtype is not None:
            self.module.to(ema_dtype)
        self.decay, self.min_decay, self.update_after_step = decay, min_decay, update_after_step
        self.use_warmup, self.warmup_gamma, self.warmup_power = use_warmup, warmup_gamma, warmup_power
        self.foreach, self.device, self.exclude_buffers = foreach, device, exclude_buffers

    def get_decay(self, step: Optional[int] = None) -> float:
        """The schedule of model_ema.py:186-203: 0 until `update_after_step`, then either the constant decay or the
        warm-up 1 - (1 + t / gamma)^-power clipped to [min_decay, decay]."""
        if step is None:
            return self.decay
        t = max(0, step - self.update_after_step - 1)
        if t <= 0:
            return 0.0
        if not self.use_warmup:
            return self.decay
        return max(min(1 - (1 + t / self.warmup_gamma) ** -self.warmup_power, self.decay), self.min_decay)

    @torch.no_grad()
    def update(self, model, step: Optional[int] = None):
        fast_update(self, model, self.get_decay(step), buffers=not self.exclude_buffers)

    @torch.no_grad()
    def set(self, model):
        for e, m in zip(self.module.state_dict().values(), model.state_dict().values()):
            e.copy_(m)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)
