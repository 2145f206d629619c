"""tests/kink.py (ReLU-kink-aware gradient comparison) checked on the oracle alone: forcing the oracle's own masks changes nothing, flipping
the unit closest to zero is reported as ONE flip within round-off of the kink and changes the gradient, a flip far from zero is reported as such
(the callers reject it)."""
import numpy as np
import torch

import kink


def _learner():
    from oracle.qmix import QmixConfig, QmixLearner, synth_batch, randomize_all
    cfg = QmixConfig(gain=1.0)
    L = QmixLearner(cfg, seed=2)
    randomize_all(L.agent, 1); randomize_all(L.mixer, 2)
    L.sync_targets()
    batch = synth_batch(cfg, 4, 6, seed=3, avail_p=0.7, var_len=True) + (None, None)
    return cfg, L, batch


def _own_masks(L, batch):
    pre = []
    hooks = [m.register_forward_hook(lambda _m, inp, _o: pre.append(inp[0].detach().clone()) if (torch.is_grad_enabled() and inp[0].requires_grad) else None)
             for m in kink._relu_modules(L.agent)]
    L.grads(batch)                       # raw gradients, no parameter update
    for h in hooks:
        h.remove()
    assert len(pre) == 2
    return pre, [(p > 0).float() for p in pre]


def test_forced_masks_reproduce_and_detect_flips():
    cfg, L, batch = _learner()
    pre, masks = _own_masks(L, batch)
    L0 = kink.snapshot(L)
    L.grads(batch)
    g_ref = {k: p.grad.clone() for k, p in L.agent.named_parameters() if p.grad is not None}
    # 1. the oracle's own masks: no flip, identical gradients
    La = kink.snapshot(L0)
    _, flips, _ = kink.redo_with_engine_masks(La, lambda LL: LL.grads(batch), masks)
    assert flips == 0
    for k, p in La.agent.named_parameters():
        if p.grad is not None:
            assert torch.equal(p.grad, g_ref[k]), k
    # 2. flip the unit closest to the kink in layer 2: one flip, tiny |pre|, different gradient
    m2 = [m.clone() for m in masks]
    flat = pre[1].abs().flatten()
    i = int(flat.argmin())
    m2[1].view(-1)[i] = 1.0 - m2[1].view(-1)[i]
    Lb = kink.snapshot(L0)
    _, flips, max_pre = kink.redo_with_engine_masks(Lb, lambda LL: LL.grads(batch), m2)
    assert flips == 1 and abs(max_pre - float(flat[i])) < 1e-12
    diff = max(float((p.grad - g_ref[k]).abs().max()) for k, p in Lb.agent.named_parameters() if p.grad is not None)
    assert diff > 0.0
    # 3. a flip far from zero is reported with its distance: callers require max_pre < KINK_TOL and fail otherwise
    m3 = [m.clone() for m in masks]
    j = int(pre[0].abs().flatten().argmax())
    m3[0].view(-1)[j] = 1.0 - m3[0].view(-1)[j]
    Lc = kink.snapshot(L0)
    _, flips, max_pre = kink.redo_with_engine_masks(Lc, lambda LL: LL.grads(batch), m3)
    assert flips == 1 and max_pre > kink.KINK_TOL
