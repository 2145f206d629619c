"""ReLU-kink-aware gradient comparison (test infrastructure).

A ReLU network's gradient is discontinuous where a pre-activation crosses zero.  With ~10^5..10^6 hidden units per learner step a
few pre-activations land within round-off of zero; the engine (3xTF32 tensor-core or FFMA accumulation order) and the fp32 oracle
then legitimately pick different sides, and that single (row, unit) changes a gradient tensor by ~1/rows of its magnitude -- far
above the 1e-4 budget although both are exact gradients of the same function at a valid sub-gradient.  When a plain comparison
fails, `redo_with_engine_masks` re-runs the oracle step from the saved pre-step state with the backward masks of the live agent's two
ReLU layers forced to the ENGINE's (u > 0 of the activations the engine saved), and reports how many units were flipped and how far
from the kink they were.  The caller then demands (a) every flipped unit had |pre-activation| below `KINK_TOL` in the oracle and
(b) the 1e-4 comparison passes against the re-run.  A real kernel error cannot pass: it either flips no unit or flips units far
from zero, or still disagrees afterwards.
"""
import copy

import torch

KINK_TOL = 3e-5          # ~30x the 3xTF32 absolute error of a K<=128 dot product of O(1) operands


class _ForcedReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask):
        ctx.save_for_backward(mask)
        return torch.relu(x)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return g * mask, None


def _relu_modules(agent):
    base = agent.rnn if hasattr(agent, "rnn") else agent.mlp        # oracle.qmix._RNNBase / oracle.mqmix._MLPBase
    m = base.mlp
    return [m.fc1[1], m.fc2[0][1]]


def engine_masks(tr, B, T, N, mlp):
    """(u1 > 0, u2 > 0) of the engine's saved post-ReLU activations, in the oracle's row order."""
    out = []
    M = B * (T + 1) * N
    for name in ("u1", "u2"):
        u = tr.ws_view(name)[:M * 64].view(B, T + 1, N, 64).cpu()
        if mlp:
            u = u[:, 0].permute(1, 0, 2).reshape(N * B, 64)                      # oracle rows n*B + b of the step-0 call
        else:
            u = u.permute(1, 2, 0, 3).reshape(T + 1, N * B, 64)                  # oracle (T+1, n*B + b, .)
        out.append((u > 0).float())
    return out


def redo_with_engine_masks(L0, step_fn, masks):
    """Run `step_fn(L0)` (one oracle step) with the live agent's ReLU backward masks forced; returns (result, n_flipped, max |pre| of a flip)."""
    stats = {"flips": 0, "max_pre": 0.0, "used": 0}
    hooks = []
    for mod, mask in zip(_relu_modules(L0.agent), masks):
        def hook(_m, inp, _out, mask=mask):
            x = inp[0]
            if torch.is_grad_enabled() and x.requires_grad and x.shape == mask.shape:
                diff = (x.detach() > 0).float() != mask
                stats["used"] += 1
                stats["flips"] += int(diff.sum())
                if diff.any():
                    stats["max_pre"] = max(stats["max_pre"], float(x.detach().abs()[diff].max()))
                return _ForcedReLU.apply(x, mask)
            return None
        hooks.append(mod.register_forward_hook(hook))
    try:
        res = step_fn(L0)
    finally:
        for h in hooks:
            h.remove()
    assert stats["used"] == 2, "forced-mask hooks did not see the live agent's forward (shape mismatch?)"
    return res, stats["flips"], stats["max_pre"]


def snapshot(L):
    return copy.deepcopy(L)


def adopt(L, L2):
    """Continue the lock-step run from the re-run's state."""
    L.__dict__.update(L2.__dict__)
