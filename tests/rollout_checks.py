"""Rollout-time policy surface (QMixPolicy.get_actions / get_q_values / get_random_actions) against the reference golden
`tests/golden/qmix_rollout.npz` (made by tests/golden/make_goldens.py rollout): shared by the emulated and the GPU tests."""
import ctypes as C

import numpy as np
import torch

from helpers import load_golden, sub
from oracle.qmix import QmixConfig
import qmix_checks as qc


def build_policy(g):
    n, o, a, hid, steps = [int(x) for x in g["meta.cfg"]]
    cfg = QmixConfig(n_agents=n, obs_dim=o, act_dim=a, state_dim=48)
    args, pol, tr = qc.build_trainer(cfg, 4, 8)
    pol.q_network.load_state_dict(sub(g, "init.agent."))
    return cfg, pol, steps


def check_rollout(tol=2e-5):
    g = load_golden("qmix_rollout")
    cfg, pol, steps = build_policy(g)
    obs, avail = g["in.obs"], g["in.avail"]
    R = obs.shape[1]
    # the runner's loop (smac_runner.py:73-98): the state returned by one call is passed to the next
    h = np.zeros((R, cfg.hidden), np.float32)
    for t in range(steps):
        a, h2, gq = pol.get_actions(obs[t], None, h, avail[t])
        assert np.array_equal(np.asarray(a, np.float32), g["greedy%d.actions" % t]), ("greedy actions", t)
        assert np.abs(h2.numpy() - g["greedy%d.h" % t]).max() <= tol, ("h", t, np.abs(h2.numpy() - g["greedy%d.h" % t]).max())
        assert tuple(gq.shape) == g["greedy%d.q" % t].shape
        assert np.abs(gq.numpy() - g["greedy%d.q" % t]).max() <= tol * max(1.0, np.abs(g["greedy%d.q" % t]).max()), ("greedy_Qs", t)
        h = h2 if t % 2 == 0 else h2.cpu().detach().numpy()        # both forms the runner can hand back (torch tensor / its ndarray)
    # a state that is NOT the one handed out must be uploaded, not taken from the device
    a, h2, _ = pol.get_actions(obs[1], None, g["greedy0.h"].copy(), avail[1])
    assert np.abs(h2.numpy() - g["greedy1.h"]).max() <= tol
    # sequence form of get_q_values
    q_seq, h_seq = pol.get_q_values(obs, None, torch.zeros(R, cfg.hidden))
    assert np.abs(q_seq.numpy() - g["seq.q"]).max() <= tol * max(1.0, np.abs(g["seq.q"]).max())
    assert np.abs(h_seq.numpy() - g["seq.h"]).max() <= tol
    # exploration: same generator calls in the same order as the reference -> identical actions under the same seeds
    for tag, av in (("explore", avail[0]), ("explore_noavail", None)):
        torch.manual_seed(5); np.random.seed(5)
        a, _, gq = pol.get_actions(obs[0], None, np.zeros((R, cfg.hidden), np.float32), av, t_env=20000, explore=True)
        assert np.array_equal(np.asarray(a, np.float32), g[tag + ".actions"]), tag
        assert tuple(gq.shape) == g[tag + ".q"].shape and np.abs(gq.numpy() - g[tag + ".q"]).max() <= tol * max(1.0, np.abs(g[tag + ".q"]).max())
    torch.manual_seed(6); np.random.seed(6)
    assert np.array_equal(np.asarray(pol.get_random_actions(obs[0], avail[0]), np.float32), g["random.actions"])
    torch.manual_seed(6); np.random.seed(6)
    assert np.array_equal(np.asarray(pol.get_random_actions(obs[0]), np.float32), g["random_noavail.actions"])


def check_errors():
    """argument validation of the C entry point (the reference would raise shape errors from torch)"""
    import ctypes as C
    from offpolicy._b200 import capi
    a = capi.PolicyStepArgs()
    assert capi.lib().mx_policy_step(C.byref(a), None) != 0
    assert b"null" in capi.lib().mx_last_error()


def check_in_place_state_edit_is_honoured():
    """The resident-state fast path keeps the recurrent state on the device when the runner hands back the array it was given.  A caller that
    edits that array IN PLACE (e.g. zeroes the rows of finished environments) must get the edited state, not the stale device copy."""
    from offpolicy._b200 import capi
    from offpolicy._b200.rollout import PolicyStepper
    I, A, R = 7, 4, 3
    cfg = capi.QmixCfg(n_agents=R, obs_dim=I, act_dim=A, state_dim=5, hidden=64, mixer_hidden=32, hyper_hidden=64, hyper_layers=2, episode_len=4, max_batch=2)
    total = C.c_int64()
    capi.lib().mx_qmix_param_layout(C.byref(cfg), None, 0, C.byref(total))
    theta = (0.1 * torch.randn(int(total.value), generator=torch.Generator().manual_seed(1))).to(capi.device())
    rs = np.random.RandomState(0)
    o1, o2 = rs.randn(R, I).astype(np.float32), rs.randn(R, I).astype(np.float32)
    a, b = PolicyStepper(I, A), PolicyStepper(I, A)
    _, h, _, _ = a.step(theta, o1, None)
    h[0] = 0.0                                         # in-place edit of the array the stepper handed out
    out_a, h_a, _, _ = a.step(theta, o2, h)
    out_b, h_b, _, _ = b.step(theta, o2, h.copy())     # a fresh stepper given the same (edited) state: always uploads it
    assert np.array_equal(out_a, out_b) and np.array_equal(h_a, h_b)
    out_c, _, _, _ = a.step(theta, o1, h_a)             # unedited hand-back: resident path, same result as an upload
    out_d, _, _, _ = b.step(theta, o1, h_b.copy())
    assert np.array_equal(out_c, out_d)
