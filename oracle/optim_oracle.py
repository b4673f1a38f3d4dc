"""ORACLE (test infrastructure, NOT product code): CPU fp32 restatement of the optimizer step on the reference's training
path.  timm builds torch.optim.SGD / AdamW / Adam (timm/optim/_optim_factory.py:355-404; 'sgd' = Nesterov) and clips with
torch.nn.utils.clip_grad_norm_ (timm/utils/clip_grad.py:18); the arithmetic lives in torch (third-party, the version in
this image is 2.11): torch/optim/sgd.py `_single_tensor_sgd`, torch/optim/adam.py `_single_tensor_adam`,
torch/nn/utils/clip_grad.py.  Pinned against tests/golden/optim.pt, produced by oracle/make_golden.py from the
unmodified reference (create_optimizer_v2 + dispatch_clip_grad) -- tests/test_oracle_golden.py.

Every function works on lists of fp32 tensors and updates them in place.  `round_to` (torch.bfloat16) models bf16
parameters with fp32 optimizer state, the storage the CUDA path uses: parameters are rounded once per step.
"""
import math

import torch


def clip_coef(grads, max_norm):
    """(total_norm, coef): torch/nn/utils/clip_grad.py -- 2-norm over all gradients, coef = min(1, max_norm / (norm + 1e-6))."""
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads))
    return total, min(1.0, max_norm / (total + 1e-6))


def _round(p, round_to):
    return p if round_to is None else p.to(round_to).float()


def sgd_step(params, grads, bufs, lr, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False, first=False,
             grad_scale=1.0, round_to=None):
    """torch/optim/sgd.py _single_tensor_sgd.  bufs: list of momentum buffers (ignored when momentum == 0)."""
    gs = torch.tensor(grad_scale, dtype=torch.float32)
    for i, (p, g) in enumerate(zip(params, grads)):
        g = g.float() * gs
        if weight_decay != 0:
            g = g + weight_decay * p
        if momentum != 0:
            if first:
                bufs[i].copy_(g)
            else:
                bufs[i].mul_(momentum).add_(g, alpha=1 - dampening)
            g = g + momentum * bufs[i] if nesterov else bufs[i]
        p.copy_(_round(p - lr * g, round_to))


def adam_step(params, grads, exp_avgs, exp_avg_sqs, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0,
              decoupled=True, grad_scale=1.0, round_to=None):
    """torch/optim/adam.py _single_tensor_adam (step counts from 1; decoupled=True is AdamW)."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    gs = torch.tensor(grad_scale, dtype=torch.float32)
    for p, g, m, v in zip(params, grads, exp_avgs, exp_avg_sqs):
        g = g.float() * gs
        q = p
        if weight_decay != 0:
            if decoupled:
                q = p * (1 - lr * weight_decay)
            else:
                g = g + weight_decay * p
        m.add_((g - m) * (1 - beta1))                      # lerp(m, g, 1 - beta1)
        v.mul_(beta2).add_(g * g, alpha=1 - beta2)
        denom = v.sqrt() / bc2_sqrt + eps
        p.copy_(_round(q - step_size * (m / denom), round_to))


def ema_decay(step, decay=0.9999, min_decay=0.0, update_after_step=0, use_warmup=False, warmup_gamma=1.0,
              warmup_power=2 / 3):
    """ModelEmaV3.get_decay, timm/utils/model_ema.py:186-203."""
    if step is None:
        return decay
    t = max(0, step - update_after_step - 1)
    if t <= 0:
        return 0.0
    if not use_warmup:
        return decay
    return max(min(1 - (1 + t / warmup_gamma) ** -warmup_power, decay), min_decay)


def ema_update(ema_sd, model_sd, decay, round_to=None):
    """ModelEmaV3.apply_update_, model_ema.py:213-235: floating-point entries ema.lerp_(model, 1 - decay) (torch.lerp:
    w < 0.5 -> a + w (b - a), else b - (b - a)(1 - w)), everything else copied.  In place on `ema_sd`."""
    w = torch.tensor(1.0 - decay, dtype=torch.float32)
    for k, e in ema_sd.items():
        m = model_sd[k]
        if not e.is_floating_point():
            e.copy_(m)
            continue
        ef, mf = e.float(), m.float()
        out = ef + w * (mf - ef) if float(w) < 0.5 else mf - (mf - ef) * (1 - w)
        e.copy_(out if round_to is None else out.to(round_to))


def cross_entropy(logits, target, smoothing=0.0):
    """Mean loss of nn.CrossEntropyLoss (smoothing 0), timm LabelSmoothingCrossEntropy (timm/loss/cross_entropy.py:20-26:
    (1 - s) * nll + s * mean_c(-logprobs)) for int64 class targets, or SoftTargetCrossEntropy (:34-36) for a [B, C]
    distribution.  Differentiable torch fp32 (autograd gives the reference gradient)."""
    logprobs = torch.log_softmax(logits.float(), dim=-1)
    if target.dim() == 2:
        return torch.sum(-target.float() * logprobs, dim=-1).mean()
    nll = -logprobs.gather(-1, target.unsqueeze(1)).squeeze(1)
    return ((1.0 - smoothing) * nll + smoothing * (-logprobs.mean(dim=-1))).mean()
