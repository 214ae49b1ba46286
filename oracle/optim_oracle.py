"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the fused parameter update, SURVEY.md §8(f) row 1.

Restates, element-wise in numpy fp32, what the reference executes after ``l_total.backward()`` in
``SRModel.optimize_parameters`` (basicsr/models/sr_model.py:164-174):

  * ``torch.nn.utils.clip_grad_norm_(params, max_norm)`` (:166-167) — third-party (PyTorch; the reference pins
    pytorch=2.0.1, environment.yaml:85; torch 2.11 here, same algorithm): per-tensor 2-norms, the 2-norm of those,
    ``clip_coef = max_norm / (total_norm + 1e-6)`` clamped to 1, gradients scaled in place;
  * ``optimizer_g.step()`` (:169) with ``torch.optim.Adam`` / ``AdamW`` (BaseModel.get_optimizer, base_model.py:120-139) —
    PyTorch's ``_single_tensor_adam``: AdamW ``p *= 1 - lr*wd`` / Adam ``g += wd*p``; ``m.lerp_(g, 1-b1)``;
    ``v = v*b2 + (1-b2)*g*g``; ``denom = sqrt(v)/sqrt(1-b2^t) + eps``; ``p += -(lr/(1-b1^t)) * m/denom``;
  * ``BaseModel.model_ema`` (base_model.py:86-95): ``ema = ema*decay + (1-decay)*p``.

Pinned by tests/test_optim_oracle_cpu.py against the real torch.optim.Adam / AdamW + clip_grad_norm_ + the reference's
model_ema loop run on CPU, and by tests/golden/optim_step.npz (made by tests/golden/make_golden_optim.py from those).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""
import numpy as np

F32 = np.float32


def clip_grad_norm(grads, max_norm):
    """grads: list of fp32 arrays (modified in place).  Returns total_norm (fp32 scalar)."""
    norms = np.array([np.sqrt(np.sum(np.square(g.astype(np.float64)))) for g in grads], dtype=np.float64).astype(F32)
    total = F32(np.sqrt(np.sum(np.square(norms.astype(np.float64)))))
    coef = min(F32(max_norm) / (total + F32(1e-6)), F32(1.0))
    for g in grads:
        g *= F32(coef)
    return total


def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False):
    """One update of every tensor, in place; ``step`` is the 1-based count of this update (torch's ``step_t += 1`` first)."""
    b1, b2 = betas
    bc1, bc2 = 1.0 - b1 ** step, 1.0 - b2 ** step
    step_size = F32(lr / bc1)
    bc2_sqrt = F32(bc2 ** 0.5)
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        g = g.astype(F32, copy=True)
        if weight_decay != 0:
            if decoupled:
                p *= F32(1.0 - lr * weight_decay)
            else:
                g += F32(weight_decay) * p
        m += F32(1.0 - b1) * (g - m)
        v *= F32(b2)
        v += F32(1.0 - b2) * g * g
        denom = np.sqrt(v) / bc2_sqrt + F32(eps)
        p += -step_size * (m / denom)


def model_ema(ema, params, decay):
    for e, p in zip(ema, params):
        e *= F32(decay)
        e += F32(1.0 - decay) * p


def train_update(params, grads, exp_avg, exp_avg_sq, ema, step, grad_clip=None, ema_decay=0.0, **adam):
    """sr_model.py:164-174 after backward: clip -> optimizer step -> EMA.  Returns total_norm or None."""
    grads = [g.astype(F32, copy=True) for g in grads]
    total = clip_grad_norm(grads, grad_clip) if grad_clip else None
    adam_step(params, grads, exp_avg, exp_avg_sq, step, **adam)
    if ema is not None and ema_decay > 0:
        model_ema(ema, params, ema_decay)
    return total
