"""Shared test helpers: golden loading, oracle construction, tolerances."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FIELDS = ["obs", "share", "acts", "rew", "dones", "dones_env", "avail"]


def load_golden(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def golden_cfg(g):
    from oracle.qmix import QmixConfig
    n, o, a, s, h, me, hy, hl, B, T, steps = [int(v) for v in g["meta.cfg"]]
    fl = [bool(v) for v in g["meta.flags"]]
    dq, hub, per = fl[:3]
    prev = fl[3] if len(fl) > 3 else False
    nofn = fl[4] if len(fl) > 4 else False
    tanh = fl[5] if len(fl) > 5 else False
    gamma, lr, eps, mgn, tau, hd, nu, peps = [float(v) for v in g["meta.hparams"]]
    cfg = QmixConfig(n_agents=n, obs_dim=o, act_dim=a, state_dim=s, hidden=h, mixer_hidden=me, hyper_hidden=hy,
                     hyper_layers=hl, gamma=gamma, lr=lr, opti_eps=eps, max_grad_norm=mgn, tau=tau, double_q=dq,
                     huber=hub, huber_delta=hd, use_per=per, per_nu=nu, per_eps=peps, prev_act_inp=prev, feature_norm=not nofn, relu=not tanh)
    return cfg, B, T, steps


def sub(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def oracle_from_golden(g):
    from oracle.qmix import QmixLearner
    cfg, B, T, steps = golden_cfg(g)
    L = QmixLearner(cfg)
    L.agent.load_state_dict(sub(g, "init.agent."))
    L.mixer.load_state_dict(sub(g, "init.mixer."))
    L.tgt_agent.load_state_dict(sub(g, "init.tgt_agent."))
    L.tgt_mixer.load_state_dict(sub(g, "init.tgt_mixer."))
    return L, cfg, B, T, steps


def golden_batch(g, s):
    b = tuple(g["s%d.in.%s" % (s, k)] for k in FIELDS)
    w = g.get("s%d.in.weights" % s)
    idx = np.arange(b[0].shape[2]) if w is not None else None
    return b + (w, idx)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))
