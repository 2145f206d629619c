"""Rollout-time policy surface (QMixPolicy.get_actions / get_q_values / get_random_actions) against the reference golden
`tests/golden/qmix_rollout.npz` (made by tests/golden/make_goldens.py rollout): shared by the emulated and the GPU tests."""
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
