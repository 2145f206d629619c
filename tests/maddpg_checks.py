"""R-MADDPG / R-MATD3 learner parity checks (emulated CPU build and real GPU): reference goldens + the pinned oracle in lock-step."""
import types

import numpy as np
import torch

from helpers import load_golden, sub, rel_err
from qmix_checks import close
from test_oracle_maddpg import maddpg_from_golden, maddpg_batch, actor_noise


from offpolicy._b200.factory import Box, Discrete, maddpg_args as make_args, build_maddpg as build  # noqa: E402,F401


def ref_tuple(b):
    d = lambda x: {"policy_0": x}
    return tuple(d(x) for x in b[:7]) + (b[7], b[8])


def named_views(flat, entries):
    out = {}
    for name, off, rows, cols in entries:
        n = rows * (cols if cols else 1)
        out[name] = flat[off:off + n].view(rows, cols) if cols else flat[off:off + n]
    return out


def check_golden(name):
    g = load_golden(name)
    L, cfg, B, T, steps = maddpg_from_golden(g)
    args, pol, tr = build(cfg, B, T)
    for tag, mod in (("actor", pol.actor), ("critic", pol.critic), ("tgt_actor", pol.target_actor), ("tgt_critic", pol.target_critic)):
        mod.load_state_dict(sub(g, "init.%s." % tag))
    problems = []
    for s in range(steps):
        batch, noise = maddpg_batch(g, s)
        torch.manual_seed(1000 + s)                         # the trainer draws the MATD3 / Gumbel noise from torch's CPU RNG like the reference
        info, prio, _ = tr.shared_train_policy_on_batch("policy_0", ref_tuple(batch))
        ref, rprio = L.step(batch, noise, actor_noise(g, s))
        ga, gc = tr.grad_views()
        for key in ("critic_loss", "critic_grad_norm"):
            e = rel_err(info[key].cpu(), g["s%d.%s" % (s, key)])
            if e > 1e-4:
                problems.append("step %d %s rel err %.3e (got %r want %r)" % (s, key, e, float(info[key]), float(g["s%d.%s" % (s, key)])))
        coef = min(1.0, cfg.max_grad_norm / (float(ref["critic_grad_norm"]) + 1e-6))
        cviews = named_views(gc, pol._c_entries)
        for k, gr in L.critic_grads.items():
            ok, err, lim = close(cviews[k] / gc[pol.Pc] * coef, gr, 1e-4)
            if not ok:
                problems.append("step %d critic grad %s err %.3e > %.3e" % (s, k, err, lim))
        assert bool(info["update_actor"]) == bool(int(g["s%d.update_actor" % s]))
        if info["update_actor"]:
            for key in ("actor_loss", "actor_grad_norm"):
                e = rel_err(info[key].cpu(), g["s%d.%s" % (s, key)])
                if e > 1e-4:
                    problems.append("step %d %s rel err %.3e (got %r want %r)" % (s, key, e, float(info[key]), float(g["s%d.%s" % (s, key)])))
            coef = min(1.0, cfg.max_grad_norm / (float(g["s%d.actor_grad_norm" % s]) + 1e-6))
            aviews = named_views(ga, pol._a_entries)
            for k, v in aviews.items():
                key = "s%d.grad.actor.%s" % (s, k)
                if key in g:
                    ok, err, lim = close(v / ga[pol.Pa] * coef, g[key], 1e-4)
                    if not ok:
                        problems.append("step %d actor grad %s err %.3e > %.3e" % (s, k, err, lim))
            pol.soft_target_updates()
            L.soft_update()
        if cfg.use_per:
            ok, err, lim = close(np.asarray(prio), g["s%d.prio" % s], 1e-4)
            if not ok:
                problems.append("step %d priorities err %.3e" % (s, err))
    for tag, mod in (("actor", pol.actor), ("critic", pol.critic), ("tgt_actor", pol.target_actor), ("tgt_critic", pol.target_critic)):
        for k, v in mod.state_dict().items():
            want = g["final.%s.%s" % (tag, k)]
            err = np.abs(v.cpu().numpy() - want).max()
            if err > 5e-3 * cfg.lr * steps + 1e-7:
                problems.append("final %s.%s err %.3e" % (tag, k, err))
    assert not problems, "\n".join(problems[:40])


def check_get_actions(name):
    """Rollout-time `get_actions` / `get_random_actions` against the reference (seeded), rMADDPGPolicy.py:62-160."""
    g = load_golden(name)
    L, cfg, B, T, steps = maddpg_from_golden(g)
    args, pol, tr = build(cfg, B, T)
    pol.actor.load_state_dict(sub(g, "init.actor."))
    obs, h = g["act.in.obs"], torch.from_numpy(g["act.in.h"])
    a, h2, _ = pol.get_actions(obs, None, h, explore=False)
    a = a.cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    if cfg.discrete:
        assert np.array_equal(a, g["act.greedy"])
    else:
        assert np.abs(a - g["act.greedy"]).max() < 1e-5
    assert np.abs(h2.cpu().numpy() - g["act.new_h"]).max() < 1e-5
    torch.manual_seed(5); np.random.seed(5)
    a, _, eps = pol.get_actions(obs, None, h, t_env=20000, explore=True)
    a = a.cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    if cfg.discrete:
        assert abs(eps - float(g["act.eps"])) < 1e-12
        assert np.array_equal(a.astype(np.float32), g["act.explore"])
    else:
        assert np.abs(a - g["act.explore"]).max() < 1e-5
    torch.manual_seed(6); np.random.seed(6)
    r = np.asarray(pol.get_random_actions(obs), dtype=np.float32)
    assert np.array_equal(r, g["act.random"])


def check_graph_matches_eager(td3, disc, B=4, T=6, E=16, steps=4):
    """sample -> shared_train_policy_on_batch -> soft update through the drop-in classes, eager vs the captured whole-update
    CUDA graphs (one per update_actor variant): same device RNG stream, same torch CPU noise stream -> same parameters."""
    import replay_checks as rc
    from offpolicy.utils.rec_buffer import RecReplayBuffer
    from offpolicy._b200.graph import MaddpgStepGraph
    from oracle.maddpg import MaddpgConfig
    from oracle.qmix import randomize_all
    cfg = MaddpgConfig(act_dim=5 if disc else 2, discrete=disc, td3=td3, actor_update_interval=2 if td3 else 1, gain=1.0)
    n, o, a, sdim = cfg.n_agents, cfg.obs_dim, cfg.act_dim, cfg.state_dim
    results = []
    for mode in ("eager", "graph"):
        rs = np.random.RandomState(3)
        info = {"policy_0": dict(obs_space=[o], share_obs_space=[sdim], act_space=Discrete(a) if disc else Box(a))}
        buf = RecReplayBuffer(info, {"policy_0": list(range(n))}, E, T, True, False, rng="device", max_batch=max(B, E))
        acts = np.eye(a, dtype=np.float32)[rs.randint(0, a, (T, E, n))] if disc else rs.uniform(-1, 1, (T, E, n, a)).astype(np.float32)
        ep = [rs.randn(T + 1, E, n, o).astype(np.float32), np.repeat(rs.randn(T + 1, E, 1, sdim).astype(np.float32), n, 2), acts,
              np.repeat(rs.randn(T, E, 1, 1).astype(np.float32), n, 2), np.zeros((T, E, n, 1), np.float32), np.zeros((T, E, 1), np.float32)]
        buf.insert(E, *[rc.d(x) for x in ep], None)
        torch.manual_seed(1)
        args, pol, tr = build(cfg, B, T)
        init = torch.Generator().manual_seed(11)
        for vec in (pol.actor_vecs[0], pol.critic_vecs[0]):
            vec.add_((0.05 * torch.randn(vec.shape, generator=init)).to(vec.device))
        pol.hard_target_updates()
        buf.seed_device_rng(5)
        torch.manual_seed(99)
        upds = []
        if mode == "eager":
            for s in range(steps):
                smp = buf.sample(B)
                info_t, _, _ = tr.shared_train_policy_on_batch("policy_0", smp)
                upds.append(bool(info_t["update_actor"]))
                if info_t["update_actor"]:
                    pol.soft_target_updates()
        else:
            g = MaddpgStepGraph(buf, tr, B)
            for s in range(steps):
                upds.append(g.launch())
            g.synchronize()
            g.close()
        results.append(([v.clone().cpu() for v in pol.actor_vecs + pol.critic_vecs], upds))
    assert results[0][1] == results[1][1]
    for x, y in zip(results[0][0], results[1][0]):
        assert float((x - y).abs().max()) <= 1e-6 * float(x.abs().max()) + 1e-7


def check_replay_batch_equals_host_batch(T=5, B=4, E=9):
    """The replay's batch region pads every episode row of the per-step fields (rewards / dones / dones_env) to 16 bytes; the
    learner must index them with the reported episode strides.  T * N = 15 and T = 5 are not multiples of 4 here: training on the
    device-side sample and on the same sample handed over as NumPy arrays (dense host layout) must give the same update."""
    import replay_checks as rc
    from offpolicy.utils.rec_buffer import RecReplayBuffer
    from oracle.maddpg import MaddpgConfig
    cfg = MaddpgConfig(act_dim=2, discrete=False, td3=False, actor_update_interval=1, gain=1.0)
    n, o, a, sdim = cfg.n_agents, cfg.obs_dim, cfg.act_dim, cfg.state_dim
    rs = np.random.RandomState(3)
    info = {"policy_0": dict(obs_space=[o], share_obs_space=[sdim], act_space=Box(a))}
    buf = RecReplayBuffer(info, {"policy_0": list(range(n))}, E, T, True, False, rng="numpy", max_batch=max(B, E))
    de = np.maximum.accumulate((rs.rand(T, E, 1) < 0.2).astype(np.float32), axis=0)
    ep = [rs.randn(T + 1, E, n, o).astype(np.float32), np.repeat(rs.randn(T + 1, E, 1, sdim).astype(np.float32), n, 2),
          rs.uniform(-1, 1, (T, E, n, a)).astype(np.float32), np.repeat(rs.randn(T, E, 1, 1).astype(np.float32), n, 2),
          np.repeat(de[:, :, None], n, 2), de]
    buf.insert(E, *[rc.d(x) for x in ep], None)
    outs = []
    for mode in ("device", "host"):
        torch.manual_seed(1)
        args, pol, tr = build(cfg, B, T)
        init = torch.Generator().manual_seed(11)
        for vec in (pol.actor_vecs[0], pol.critic_vecs[0]):
            vec.add_((0.05 * torch.randn(vec.shape, generator=init)).to(vec.device))
        pol.hard_target_updates()
        np.random.seed(7)
        smp = buf.sample(B)
        if mode == "host":
            smp = tuple({"policy_0": smp[i]["policy_0"]} for i in range(6)) + (None, None, None)
        info_t, _, _ = tr.shared_train_policy_on_batch("policy_0", smp)
        outs.append(([float(info_t["critic_loss"]), float(info_t["actor_loss"])], [v.clone().cpu() for v in pol.actor_vecs[:1] + pol.critic_vecs[:1]]))
    assert outs[0][0] == outs[1][0], (outs[0][0], outs[1][0])
    for x, y in zip(outs[0][1], outs[1][1]):
        assert torch.equal(x, y)
