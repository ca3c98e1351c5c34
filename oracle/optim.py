"""TEST INFRASTRUCTURE: restatement of the reference's AdamW update (lib/helpers/optimizer_helper.py:88-127) with the
same torch operations in the same order (`mul_().add_()`, `addcmul_`, `sqrt().add_()`, `mul().addcdiv_()`, `add_`), so the
fused kernel can be compared with what the reference computes on the SAME device.  Deprecated positional-alpha overloads
of the reference are spelled with keywords (identical arithmetic).  Pinned against the unmodified reference class in
tests/test_oracle_optim.py."""
import math

import torch


def adamw_reference_step(params, grads, exp_avgs, exp_avg_sqs, step, lr, beta1, beta2, eps, weight_decays):
    """One step over lists of tensors; `step` is the 1-based step count AFTER the increment (:109)."""
    bias_correction1 = 1 - beta1 ** step
    bias_correction2 = 1 - beta2 ** step
    step_size = lr * math.sqrt(bias_correction2) / bias_correction1
    for p, g, m, v, wd in zip(params, grads, exp_avgs, exp_avg_sqs, weight_decays):
        if g is None:
            continue
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = v.sqrt().add_(eps)
        p.add_(torch.mul(p, wd).addcdiv_(m, denom, value=1), alpha=-step_size)
