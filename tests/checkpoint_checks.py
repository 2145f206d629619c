"""Checkpoint / resume: a run restored from `save_checkpoint` continues bit-identically (shared by emulated and GPU tests)."""
import os
import tempfile

import numpy as np
import torch

import qmix_checks as qc
import replay_checks as rc
from oracle.qmix import QmixConfig


def _episodes(cfg, T, n, rs):
    N, O, A, S = cfg.n_agents, cfg.obs_dim, cfg.act_dim, cfg.state_dim
    de = np.maximum.accumulate((rs.rand(T, n, 1) < 0.1).astype(np.float32), axis=0)
    return [rc.d(x) for x in (rs.randn(T + 1, n, N, O).astype(np.float32), np.repeat(rs.randn(T + 1, n, 1, S).astype(np.float32), N, 2),
                              np.eye(A, dtype=np.float32)[rs.randint(0, A, (T, n, N))], np.repeat(rs.randn(T, n, 1, 1).astype(np.float32), N, 2),
                              np.repeat(de[:, :, None], N, 2), de, (rs.rand(T + 1, n, N, A) < 0.7).astype(np.float32) + np.eye(A, dtype=np.float32)[0])]


def _build(cfg, B, T, E, per, device_rng):
    buf = rc.make_buffers(cfg.n_agents, cfg.obs_dim, cfg.act_dim, cfg.state_dim, T, E, per_alpha=0.6 if per else None,
                          rng="numpy", max_batch=max(B, 12))
    torch.manual_seed(2)
    args, pol, tr = qc.build_trainer(cfg, B, T)
    return buf, pol, tr


def _steps(buf, tr, cfg, B, n, rs, T):
    out = []
    for _ in range(n):
        buf.insert(1, *_episodes(cfg, T, 1, rs))
        smp = buf.sample(B, 0.5, "policy_0") if cfg.use_per else buf.sample(B)
        info, prio, idx = tr.train_policy_on_batch(smp)
        if cfg.use_per:
            buf.update_priorities(idx, prio, "policy_0")
        tr.soft_target_updates()
        out.append([float(info["loss"]), float(info["grad_norm"]), float(info["Q_tot"])])
    return out


def check_resume(per=False, device_rng=False, n=3):
    from offpolicy._b200.checkpoint import save_checkpoint, load_checkpoint
    cfg = QmixConfig(n_agents=3, obs_dim=9, act_dim=5, state_dim=11, use_per=per, gain=1.0)
    B, T, E = 4, 6, 12
    buf, pol, tr = _build(cfg, B, T, E, per, device_rng)
    rs = np.random.RandomState(5)
    buf.insert(12 - n, *_episodes(cfg, T, 12 - n, rs))      # so that the steps after the checkpoint wrap the 12-slot ring
    np.random.seed(11)
    if device_rng:
        buf.seed_device_rng(11)
    _steps(buf, tr, cfg, B, n, rs, T)                       # ring wraps during these + the next steps (12 slots)
    with tempfile.TemporaryDirectory() as d:
        path = save_checkpoint(os.path.join(d, "ck.pt"), tr, buf, extra={"episode": 3})
        rs_state = rs.get_state()
        want = _steps(buf, tr, cfg, B, n, rs, T)
        want_theta = tr.theta.cpu().clone()
        want_tgt = tr.theta_tgt.cpu().clone()
        want_len = len(buf)
        # a fresh process: new objects, different initial weights and generator states
        buf2, pol2, tr2 = _build(cfg, B, T, E, per, device_rng)
        np.random.seed(999); torch.manual_seed(999)
        extra = load_checkpoint(path, tr2, buf2)
        assert extra == {"episode": 3}
        rs2 = np.random.RandomState(0)
        rs2.set_state(rs_state)
        got = _steps(buf2, tr2, cfg, B, n, rs2, T)
    assert got == want, (got, want)                                   # bit-identical scalars
    assert torch.equal(tr2.theta.cpu(), want_theta) and torch.equal(tr2.theta_tgt.cpu(), want_tgt)
    assert len(buf2) == want_len
    # the per-network state_dicts keep the reference's key names (App. E), so its .pt files round-trip
    sd = pol2.q_network.state_dict()
    assert list(sd)[:2] == ["rnn.feature_norm.weight", "rnn.feature_norm.bias"] and "q.action_out.weight" in sd
    assert "hyper_w1.0.weight" in tr2.mixer.state_dict()


def check_rejects_wrong_shape():
    from offpolicy._b200.checkpoint import save_checkpoint, load_checkpoint
    cfg = QmixConfig(n_agents=3, obs_dim=9, act_dim=5, state_dim=11, gain=1.0)
    buf, pol, tr = _build(cfg, 4, 6, 12, False, False)
    cfg2 = QmixConfig(n_agents=3, obs_dim=10, act_dim=5, state_dim=11, gain=1.0)
    buf2, pol2, tr2 = _build(cfg2, 4, 6, 12, False, False)
    with tempfile.TemporaryDirectory() as d:
        path = save_checkpoint(os.path.join(d, "ck.pt"), tr, buf)
        for kw in (dict(trainer=tr2), dict(buffer=buf2)):
            try:
                load_checkpoint(path, **kw)
            except ValueError:
                continue
            raise AssertionError("a checkpoint of another configuration was accepted: %r" % (list(kw),))
