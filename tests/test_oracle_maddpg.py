"""Pin oracle/maddpg.py against outputs of the unmodified reference (tests/golden/ma*.npz)."""
import numpy as np
import pytest
import torch

from helpers import load_golden, sub, rel_err


def maddpg_from_golden(g):
    from oracle.maddpg import MaddpgConfig, MaddpgLearner
    meta = [int(v) for v in g["meta.cfg"]]
    n, o, a, s, h, B, T, steps, td3, per = meta[:10]
    disc = bool(meta[10]) if len(meta) > 10 else False
    nofn = bool(meta[11]) if len(meta) > 11 else False
    tanh = bool(meta[12]) if len(meta) > 12 else False
    gamma, lr, eps, mgn, tau, hd, nu, peps, tn, wd = [float(v) for v in g["meta.hparams"]]
    cfg = MaddpgConfig(n_agents=n, obs_dim=o, act_dim=a, state_dim=s, hidden=h, gamma=gamma, lr=lr, opti_eps=eps, max_grad_norm=mgn,
                       tau=tau, huber_delta=hd, per_nu=nu, per_eps=peps, td3=bool(td3), target_noise=tn, weight_decay=wd,
                       use_per=bool(per), actor_update_interval=2 if td3 else 1, discrete=disc, feature_norm=not nofn, relu=not tanh)
    L = MaddpgLearner(cfg)
    for tag, mod in (("actor", L.actor), ("critic", L.critic), ("tgt_actor", L.tgt_actor), ("tgt_critic", L.tgt_critic)):
        mod.load_state_dict(sub(g, "init.%s." % tag))
    return L, cfg, B, T, steps


def maddpg_batch(g, s):
    b = tuple(g["s%d.in.%s" % (s, k)] for k in ("obs", "share", "acts", "rew", "dones", "dones_env"))
    w = g.get("s%d.in.weights" % s)
    return b + (g.get("s%d.in.avail" % s), w, np.arange(b[0].shape[2]) if w is not None else None), g.get("s%d.in.noise" % s)


def actor_noise(g, s):
    return g.get("s%d.in.actor_noise" % s)


@pytest.mark.parametrize("name", ["maddpg_box", "matd3_box", "maddpg_box_per", "maddpg_disc", "matd3_disc", "matd3_disc_avail", "matd3_disc_nofn", "maddpg_box_tanh"])
def test_oracle_reproduces_reference_maddpg(name):
    torch.set_num_threads(1)
    g = load_golden(name)
    L, cfg, B, T, steps = maddpg_from_golden(g)
    for s in range(steps):
        batch, noise = maddpg_batch(g, s)
        info, prio = L.step(batch, noise, actor_noise(g, s))
        assert rel_err(info["critic_loss"], g["s%d.critic_loss" % s]) < 1e-6
        assert rel_err(info["critic_grad_norm"], g["s%d.critic_grad_norm" % s]) < 1e-5
        assert int(info["update_actor"]) == int(g["s%d.update_actor" % s])
        if info["update_actor"]:
            assert rel_err(info["actor_loss"], g["s%d.actor_loss" % s]) < 1e-5
            assert rel_err(info["actor_grad_norm"], g["s%d.actor_grad_norm" % s]) < 1e-5
            for k, gr in L.actor_grads.items():
                key = "s%d.grad.actor.%s" % (s, k)
                if key in g:
                    assert rel_err(gr, g[key]) < 2e-5, key
            L.soft_update()
        if cfg.use_per:
            assert rel_err(prio, g["s%d.prio" % s]) < 1e-5
    for tag, mod in (("actor", L.actor), ("critic", L.critic), ("tgt_actor", L.tgt_actor), ("tgt_critic", L.tgt_critic)):
        for k, v in mod.state_dict().items():
            assert rel_err(v, g["final.%s.%s" % (tag, k)]) < 5e-6, (tag, k)


def test_oracle_gumbel_draws_match_reference_stream():
    """`sample_gumbel` consumes torch's CPU generator like the reference (util.py:127-130): the goldens' noise was produced by
    the reference's own function after `torch.manual_seed(1000 + s)`."""
    from oracle.maddpg import sample_gumbel
    g = load_golden("matd3_disc")
    meta = [int(v) for v in g["meta.cfg"]]
    n, a, B, T = meta[0], meta[2], meta[5], meta[6]
    torch.manual_seed(1000)
    g1 = sample_gumbel((T + 1, n * B, a))
    g2 = sample_gumbel((T, n * B, a))
    assert np.array_equal(g1.numpy(), g["s0.in.noise"])
    assert np.array_equal(g2.numpy(), g["s0.in.actor_noise"])


def multi_from_golden(g):
    from oracle.maddpg import MaddpgConfig, MaddpgMultiLearner
    N, S, H, B, T, rounds, td3, disc = [int(v) for v in g["meta.cfg"]]
    specs = [tuple(int(x) for x in row) for row in g["meta.specs"]]
    gamma, lr, eps, mgn, tau, hd, nu, peps, tn, wd = [float(v) for v in g["meta.hparams"]]
    base = MaddpgConfig(hidden=H, gamma=gamma, lr=lr, opti_eps=eps, max_grad_norm=mgn, tau=tau, huber_delta=hd, per_nu=nu, per_eps=peps, td3=bool(td3),
                        target_noise=tn, weight_decay=wd, actor_update_interval=2 if td3 else 1, discrete=bool(disc), gain=1.0)
    L = MaddpgMultiLearner(specs, S, base)
    for i in range(N):
        for tag, mods in (("actor", L.actor), ("critic", L.critic), ("tgt_actor", L.tgt_actor), ("tgt_critic", L.tgt_critic)):
            mods[i].load_state_dict(sub(g, "init.p%d.%s." % (i, tag)))
    return L, specs, B, T, rounds, bool(td3), bool(disc)


def multi_round_inputs(g, r, N):
    obs = [g["r%d.in.p%d.obs" % (r, i)] for i in range(N)]
    acts = [g["r%d.in.p%d.acts" % (r, i)] for i in range(N)]
    dones = [g["r%d.in.p%d.dones" % (r, i)] for i in range(N)]
    return obs, g["r%d.in.share" % r], acts, g["r%d.in.rew" % r], dones, g["r%d.in.dones_env" % r]


@pytest.mark.parametrize("name", ["maddpg_multi_disc", "matd3_multi_box", "matd3_multi_disc"])
def test_oracle_reproduces_reference_per_agent_policies(name):
    """share_policy = False (scripts/train_mpe_rmaddpg.sh:14): one policy per agent; goldens = the unmodified reference's
    shared_train_policy_on_batch called for every policy on one sample, then soft updates (base_runner.py:225-256)."""
    torch.set_num_threads(1)
    g = load_golden(name)
    L, specs, B, T, rounds, td3, disc = multi_from_golden(g)
    N = len(specs)
    for r in range(rounds):
        obs, share, acts, rew, dones, de = multi_round_inputs(g, r, N)
        upd_any = False
        for i in range(N):
            noises = {q: g["r%d.u%d.noise.p%d" % (r, i, q)] for q in range(N)} if td3 else None
            info = L.step(i, obs, share, acts, rew, dones, de, noises, g.get("r%d.u%d.actor_noise" % (r, i)))
            assert rel_err(info["critic_loss"], g["r%d.u%d.critic_loss" % (r, i)]) < 1e-6
            assert rel_err(info["critic_grad_norm"], g["r%d.u%d.critic_grad_norm" % (r, i)]) < 1e-5
            assert int(info["update_actor"]) == int(g["r%d.u%d.update_actor" % (r, i)])
            if info["update_actor"]:
                upd_any = True
                assert rel_err(info["actor_loss"], g["r%d.u%d.actor_loss" % (r, i)]) < 1e-5
                assert rel_err(info["actor_grad_norm"], g["r%d.u%d.actor_grad_norm" % (r, i)]) < 1e-5
                for k, gr in L.actor_grads.items():
                    key = "r%d.u%d.grad.actor.%s" % (r, i, k)
                    if key in g:
                        assert rel_err(gr, g[key]) < 2e-5, key
        if upd_any:
            L.soft_update_all()
    for i in range(N):
        for tag, mods in (("actor", L.actor), ("critic", L.critic), ("tgt_actor", L.tgt_actor), ("tgt_critic", L.tgt_critic)):
            for k, v in mods[i].state_dict().items():
                assert rel_err(v, g["final.p%d.%s.%s" % (i, tag, k)]) < 5e-6, (i, tag, k)
