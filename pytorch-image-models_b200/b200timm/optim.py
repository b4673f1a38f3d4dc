"""Optimizer step on libb200timm kernels: drop-in `torch.optim.Optimizer` subclasses for the optimizers timm's factory
builds for this path -- 'sgd' / 'momentum' (torch.optim.SGD, timm/optim/_optim_factory.py:355-369) and 'adamw' / 'adam'
(torch.optim.AdamW / Adam, :394-404) -- and the 2-norm gradient clipping of timm/utils/clip_grad.py:18
(torch.nn.utils.clip_grad_norm_).  Constructor signatures, param-group keys and update formulas follow torch.optim; the
differences are execution and state precision:

* one kernel launch updates a whole parameter group (b200_sgd_step / b200_adam_step: multi-tensor, pointer table in the
  kernel's parameter space) instead of torch's chain of foreach kernels;
* optimizer state (momentum buffer, exp_avg, exp_avg_sq) is fp32 even for bf16 parameters (torch keeps it in the
  parameter dtype); fp32 parameters reproduce torch.optim to rounding;
* `clip_grad_norm_(..., optimizer=opt)` hands the clip coefficient to the optimizer, whose next step multiplies it into
  the gradients as they are read (no extra pass over the gradients, no host synchronisation).

There is no CPU path: CPU parameters raise B200Error.
"""
from typing import Iterable, Optional

import torch

from . import _lib
from ._lib import B200Error, DTYPE_BF16, DTYPE_F32, OptimTensor
from .ops import _count, _stream, note_params_written

__all__ = ["SGD", "AdamW", "Adam", "clip_grad_norm_", "register_with_timm"]

_DT = {torch.bfloat16: DTYPE_BF16, torch.float32: DTYPE_F32}


def _table(rows):
    """rows: list of (param, grad, state1 or None, state2 or None) -> ctypes array of b200_optim_tensor."""
    arr = (OptimTensor * len(rows))()
    for i, (p, g, s1, s2) in enumerate(rows):
        e = arr[i]
        e.param = None if p is None else p.data_ptr()
        e.grad = g.data_ptr()
        e.state1 = None if s1 is None else s1.data_ptr()
        e.state2 = None if s2 is None else s2.data_ptr()
        e.numel = g.numel()
    return arr


def _grad_of(p):
    g = p.grad
    if g is None:
        return None
    if g.is_sparse:
        raise B200Error("sparse gradients are not supported")
    if not p.is_cuda or not g.is_cuda:
        raise B200Error("libb200timm kernels need CUDA tensors (there is no CPU path)")
    if p.dtype not in _DT or g.dtype != p.dtype:
        raise B200Error(f"parameter / gradient dtype must be bf16 or fp32 and equal, got {p.dtype} / {g.dtype}")
    if not p.is_contiguous():
        raise B200Error("parameters must be contiguous")
    return g if g.is_contiguous() else g.contiguous()


def _f32_state(state, key):
    """Optimizer.load_state_dict casts floating-point state to the parameter dtype; the kernels need fp32 state."""
    t = state[key]
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = state[key] = t.float().contiguous()
    return t


def _by_dtype(rows):
    out = {}
    for r in rows:
        out.setdefault(r[0].dtype if r[0] is not None else r[1].dtype, []).append(r)
    return out


class _B200Optimizer(torch.optim.Optimizer):
    _unsupported = ("maximize", "differentiable", "amsgrad")

    def _check_flags(self, kw):
        for k in self._unsupported:
            if kw.get(k):
                raise B200Error(f"{type(self).__name__}: {k}=True is not supported")

    def set_grad_scale(self, coef: Optional[torch.Tensor]):
        """fp32 device scalar multiplied into the gradients by the NEXT step() only (see clip_grad_norm_)."""
        self._grad_scale = coef

    def _take_scale(self):
        """The pending clip coefficient as a TENSOR (or None): step() keeps it in a local until every group is launched,
        so the caching allocator cannot hand its memory to an allocation made inside step() (momentum / exp_avg buffers
        of a first step, fp32 state casts after load_state_dict) before the kernels have read it."""
        c = getattr(self, "_grad_scale", None)
        self._grad_scale = None
        return c


class SGD(_B200Optimizer):
    """torch.optim.SGD (torch/optim/sgd.py) on b200_sgd_step."""

    def __init__(self, params, lr=1e-3, momentum=0., dampening=0., weight_decay=0., nesterov=False, *, maximize=False,
                 foreach=None, differentiable=False, fused=None):
        if lr < 0.0 or momentum < 0.0 or weight_decay < 0.0:
            raise ValueError("invalid lr / momentum / weight_decay")
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        self._check_flags(dict(maximize=maximize, differentiable=differentiable))
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                      nesterov=nesterov))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib, scale_t = _lib.load(), self._take_scale()
        scale = None if scale_t is None else scale_t.data_ptr()
        written = []
        for group in self.param_groups:
            mom = group["momentum"]
            fresh, seen = [], []
            for p in group["params"]:
                g = _grad_of(p)
                if g is None:
                    continue
                buf = None
                if mom != 0:
                    state = self.state[p]
                    buf = _f32_state(state, "momentum_buffer") if state.get("momentum_buffer") is not None else None
                    if buf is None:
                        buf = state["momentum_buffer"] = torch.empty(p.shape, dtype=torch.float32, device=p.device)
                        fresh.append((p, g, buf, None))
                        continue
                seen.append((p, g, buf, None))
            for first, rows in ((1, fresh), (0, seen)):
                for dt, rs in _by_dtype(rows).items():
                    st = _stream()
                    rc = lib.b200_sgd_step(_table(rs), len(rs), _DT[dt], group["lr"], mom, group["dampening"],
                                           group["weight_decay"], int(group["nesterov"]), first, scale, st)
                    _lib.check(rc, "b200_sgd_step")
                    _count()
                    written.extend(r[0] for r in rs)
        del scale_t            # alive until here: every launch that reads it is enqueued
        note_params_written(written)
        return loss


class AdamW(_B200Optimizer):
    """torch.optim.AdamW (torch/optim/adamw.py -> adam.py with decoupled weight decay) on b200_adam_step."""

    _decoupled = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, *,
                 maximize=False, foreach=None, capturable=False, differentiable=False, fused=None):
        if isinstance(lr, torch.Tensor):
            raise B200Error("tensor learning rates are not supported")
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid lr / eps / betas / weight_decay")
        self._check_flags(dict(maximize=maximize, differentiable=differentiable, amsgrad=amsgrad))
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib, scale_t = _lib.load(), self._take_scale()
        scale = None if scale_t is None else scale_t.data_ptr()
        written = []
        for group in self.param_groups:
            by_step = {}
            for p in group["params"]:
                g = _grad_of(p)
                if g is None:
                    continue
                state = self.state[p]
                if not state:
                    state["step"] = 0      # a Python int: the kernel needs it on the host (torch keeps a CPU tensor)
                    state["exp_avg"] = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
                    state["exp_avg_sq"] = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
                state["step"] = int(state["step"]) + 1
                by_step.setdefault((state["step"], p.dtype), []).append(
                    (p, g, _f32_state(state, "exp_avg"), _f32_state(state, "exp_avg_sq")))
            b1, b2 = group["betas"]
            for (step, dt), rs in by_step.items():
                st = _stream()
                rc = lib.b200_adam_step(_table(rs), len(rs), _DT[dt], group["lr"], b1, b2, group["eps"],
                                        group["weight_decay"], step, int(self._decoupled), scale, st)
                _lib.check(rc, "b200_adam_step")
                _count()
                written.extend(r[0] for r in rs)
        del scale_t            # alive until here: every launch that reads it is enqueued
        note_params_written(written)
        return loss


class Adam(AdamW):
    """torch.optim.Adam (L2 weight decay added to the gradient) on b200_adam_step."""

    _decoupled = False

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0., amsgrad=False, **kw):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad, **kw)


def clip_grad_norm_(parameters: Iterable[torch.Tensor], max_norm: float, norm_type: float = 2.0,
                    optimizer: Optional[_B200Optimizer] = None) -> torch.Tensor:
    """torch.nn.utils.clip_grad_norm_ for the 2-norm (what timm's dispatch_clip_grad(mode='norm') calls).  Returns the
    total gradient norm as a device scalar -- nothing synchronises with the host.  With `optimizer` given (one of this
    module's), the gradients are left untouched and the coefficient min(1, max_norm / (norm + 1e-6)) is applied inside
    that optimizer's next step(); without it they are scaled in place, as torch does."""
    if float(norm_type) != 2.0:
        raise B200Error("clip_grad_norm_: only the 2-norm is supported")
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    rows = []
    for p in parameters:
        g = p.grad
        if g is None:
            continue
        if not g.is_cuda or g.dtype not in _DT:
            raise B200Error("clip_grad_norm_: gradients must be CUDA bf16 / fp32 tensors")
        if not g.is_contiguous():
            raise B200Error("clip_grad_norm_: gradients must be contiguous")
        rows.append((None, g, None, None))
    if not rows:
        return torch.zeros((), dtype=torch.float32)
    dev = rows[0][1].device
    buf = torch.zeros(3, dtype=torch.float32, device=dev)        # sumsq, coef, total_norm
    lib, st = _lib.load(), _stream()
    groups = _by_dtype(rows)
    for dt, rs in groups.items():
        _lib.check(lib.b200_grad_sumsq(_table(rs), len(rs), _DT[dt], buf.data_ptr(), st), "b200_grad_sumsq")
        _count()
    _lib.check(lib.b200_clip_coef(buf.data_ptr(), float(max_norm), buf.data_ptr() + 4, buf.data_ptr() + 8, st),
               "b200_clip_coef")
    _count()
    if optimizer is not None:
        if not isinstance(optimizer, _B200Optimizer):
            raise B200Error("clip_grad_norm_(optimizer=...) needs a b200timm.optim optimizer")
        optimizer.set_grad_scale(buf[1:2])
    else:
        for dt, rs in groups.items():
            _lib.check(lib.b200_grad_scale(_table(rs), len(rs), _DT[dt], buf.data_ptr() + 4, st), "b200_grad_scale")
            _count()
    return buf[2]


def register_with_timm():
    """Add 'b200sgd', 'b200momentum', 'b200adamw', 'b200adam' to timm's optimizer registry (timm/optim/_optim_factory.py
    OptimInfo / default_registry), so `train.py --opt b200adamw` / create_optimizer_v2(model, 'b200sgd', ...) build these."""
    from timm.optim._optim_factory import OptimInfo, default_registry
    for name, cls, kw in (("b200sgd", SGD, dict(has_eps=False, has_momentum=True, defaults={"nesterov": True})),
                          ("b200momentum", SGD, dict(has_eps=False, has_momentum=True, defaults={"nesterov": False})),
                          ("b200adamw", AdamW, dict(has_betas=True)), ("b200adam", Adam, dict(has_betas=True))):
        default_registry.register(OptimInfo(name=name, opt_class=cls, description=f"libb200timm {cls.__name__}", **kw))
