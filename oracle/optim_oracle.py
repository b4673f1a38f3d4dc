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
