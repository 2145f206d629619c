"""MLP (transition-level) QMIX parity checks against goldens made by the reference's M_QMixPolicy / M_QMix (tests/golden/mqmix_*.npz,
make_goldens.py mqmix): shared by the emulated (CPU) and the GPU test modules.  Tolerances as in qmix_checks.py."""
import numpy as np
import torch

from helpers import load_golden, golden_cfg, sub, rel_err
import qmix_checks as qc
from replay_checks import Discrete


def build(cfg, B, debug=True):
    from offpolicy.algorithms.mqmix.algorithm.mQMixPolicy import M_QMixPolicy
    from offpolicy.algorithms.mqmix.mqmix import M_QMix
    from offpolicy._b200 import capi
    args = qc.make_args(cfg, B)
    info = dict(obs_space=[cfg.obs_dim], share_obs_space=[cfg.state_dim], act_space=Discrete(cfg.act_dim), cent_obs_dim=cfg.state_dim,
                cent_act_dim=cfg.act_dim * cfg.n_agents)
    pol = M_QMixPolicy({"args": args, "device": capi.device()}, info)
    tr = M_QMix(args, cfg.n_agents, {"policy_0": pol}, lambda a: "policy_0", device=capi.device())
    capi.lib().mx_qmix_set_debug(tr.handle, 1 if debug else 0)
    return args, pol, tr


def golden_transitions(g, s):
    d = lambda k: {"policy_0": g.get("s%d.in.%s" % (s, k))}
    w = g.get("s%d.in.weights" % s)
    B = g["s%d.in.obs" % s].shape[1]
    return (d("obs"), d("share"), d("acts"), d("rew"), d("nobs"), d("nshare"), d("dones"), d("dones_env"), d("valid"), d("avail"), d("navail"),
            w, np.arange(B) if w is not None else None)


def check_golden(name, debug=True):
    g = load_golden(name)
    cfg, B, T, steps = golden_cfg(g)
    assert T == 1
    args, pol, tr = build(cfg, B, debug)
    pol.q_network.load_state_dict(sub(g, "init.agent."))
    tr.target_q_network.load_state_dict(sub(g, "init.tgt_agent."))
    tr.mixer.load_state_dict(sub(g, "init.mixer."))
    tr.target_mixer.load_state_dict(sub(g, "init.tgt_mixer."))
    problems = []
    for s in range(steps):
        prev = {k: v.clone() for k, v in pol.q_network.state_dict().items()}
        prev_m = {k: v.clone() for k, v in tr.mixer.state_dict().items()}
        info, prio, idx = tr.train_policy_on_batch(golden_transitions(g, s), True)
        for key in ("loss", "grad_norm", "Q_tot"):
            e = rel_err(info[key].cpu(), g["s%d.%s" % (s, key)])
            if e > 1e-4:
                problems.append("step %d %s: rel err %.3e (got %r want %r)" % (s, key, e, float(info[key]), float(g["s%d.%s" % (s, key)])))
        if cfg.use_per:
            ok, err, lim = qc.close(np.asarray(prio), g["s%d.prio" % s], 1e-4)
            if not ok:
                problems.append("step %d priorities err %.3e" % (s, err))
        gn = float(g["s%d.grad_norm" % s])
        coef = min(1.0, cfg.max_grad_norm / (gn + 1e-6))
        for full, ours in tr.grad_views().items():
            role, pname = full.split(".", 1)
            key = "s%d.grad.%s.%s" % (s, role, pname)
            if key not in g:
                if float(ours.abs().max()) != 0.0:
                    problems.append("step %d grad %s should be zero (unused parameter)" % (s, full))
                continue
            ok, err, lim = qc.close(ours * coef, g[key], 1e-4)
            if not ok:
                problems.append("step %d grad %s: err %.3e > %.3e" % (s, full, err, lim))
        tr.soft_target_updates()
        for role, mod, prv in (("agent", pol.q_network, prev), ("mixer", tr.mixer, prev_m)):
            for k, v in mod.state_dict().items():
                want = g["s%d.%s.%s" % (s, role, k)]
                err = np.abs((v.cpu() - prv[k].cpu()).numpy() - (want - prv[k].cpu().numpy())).max()
                gkey = "s%d.grad.%s.%s" % (s, role, k)
                gmax = float(np.abs(g[gkey]).max()) if gkey in g else 0.0
                lim = cfg.lr * min(2.0, 5e-3 + 1e-4 * gmax / cfg.opti_eps)
                if err > lim + 1e-9:
                    problems.append("step %d param %s.%s: update err %.3e" % (s, role, k, err))
        for role, mod in (("tgt_agent", tr.target_q_network), ("tgt_mixer", tr.target_mixer)):
            for k, v in mod.state_dict().items():
                ok, err, lim = qc.close(v, g["s%d.%s.%s" % (s, role, k)], 1e-6, 1e-7)
                if not ok:
                    problems.append("step %d %s.%s: err %.3e" % (s, role, k, err))
    # rollout surface (the golden evaluates it after the last update; the weights agree to the Adam-step tolerance by then)
    a, q = pol.get_actions(g["roll.obs"], g["roll.avail"])
    if not np.array_equal(np.asarray(a, np.float32), g["roll.greedy"]):
        problems.append("greedy rollout actions differ")
    if np.abs(q.numpy() - g["roll.greedy_q"]).max() > 1e-4 * max(1.0, np.abs(g["roll.greedy_q"]).max()):
        problems.append("greedy_Qs differ")
    torch.manual_seed(5); np.random.seed(5)
    a, _ = pol.get_actions(g["roll.obs"], g["roll.avail"], t_env=20000, explore=True)
    if not np.array_equal(np.asarray(a, np.float32), g["roll.explore"]):
        problems.append("exploring rollout actions differ")
    qa = pol.get_q_values(g["roll.obs"]).numpy()
    if np.abs(qa - g["roll.q_all"]).max() > 1e-4 * max(1.0, np.abs(g["roll.q_all"]).max()):
        problems.append("get_q_values differs")
    # the unused recurrent slots of the flat vector never move
    used = torch.zeros(tr.P, dtype=torch.bool)
    for name_, off, rows, cols in tr.entries:
        used[off:off + rows * (cols if cols else 1)] = True
    if float(tr.theta.cpu()[~used].abs().max()) != 0.0:
        problems.append("unused slots of the parameter vector changed")
    assert not problems, "\n".join(problems[:40])


def check_buffer_vs_reference_layout(seed=0):
    """MlpReplayBuffer (length-1 episodes of the HBM replay) returns the reference's sample layout: inserts with ring wrap, uniform
    sampling from NumPy's stream, reward normalisation (mlp_buffer.py:203-240), checked against a NumPy restatement of the store."""
    import replay_checks as rc
    from offpolicy.utils.mlp_buffer import MlpReplayBuffer
    N, O, A, S, E, B = 3, 6, 4, 7, 20, 8
    info = {"policy_0": dict(obs_space=[O], share_obs_space=[S], act_space=Discrete(A))}
    rs = np.random.RandomState(seed)
    for norm, avail in ((False, True), (True, False)):
        buf = MlpReplayBuffer(info, {"policy_0": list(range(N))}, E, True, avail, use_reward_normalization=norm, max_batch=16)
        store = {k: np.zeros((E,) + sh, np.float32) for k, sh in dict(obs=(N, O), share=(S,), acts=(N, A), rew=(N, 1), nobs=(N, O), nshare=(S,),
                                                                        dones=(N, 1), dones_env=(1,), valid=(N, 1), avail=(N, A), navail=(N, A)).items()}
        cur = filled = 0
        d = lambda x: {"policy_0": x}
        for k in range(31):
            n = 1 if k % 5 else 3
            f = dict(obs=rs.randn(n, N, O), share=rs.randn(n, S), acts=np.eye(A)[rs.randint(0, A, (n, N))], rew=2.0 + rs.randn(n, N, 1), nobs=rs.randn(n, N, O),
                     nshare=rs.randn(n, S), dones=(rs.rand(n, N, 1) < 0.2) * 1.0, dones_env=(rs.rand(n, 1) < 0.2) * 1.0, valid=(rs.rand(n, N, 1) < 0.9) * 1.0,
                     avail=(rs.rand(n, N, A) < 0.5) * 1.0, navail=(rs.rand(n, N, A) < 0.5) * 1.0)
            f = {kk: v.astype(np.float32) for kk, v in f.items()}
            idx = buf.insert(n, d(f["obs"]), d(f["share"]), d(f["acts"]), d(f["rew"]), d(f["nobs"]), d(f["nshare"]), d(f["dones"]), d(f["dones_env"]),
                             d(f["valid"]), d(f["avail"] if avail else None), d(f["navail"] if avail else None))
            want_idx = (cur + np.arange(n)) % E
            assert np.array_equal(idx, want_idx), (k, idx, want_idx)
            for kk in store:
                store[kk][want_idx] = f[kk]
            cur, filled = int(want_idx[-1]) + 1, min(filled + n, E)
            assert len(buf) == filled
            if filled < B:
                continue
            st = np.random.get_state()
            smp = buf.sample(B)
            np.random.set_state(st)
            inds = np.random.choice(filled, B)
            rew = store["rew"][inds]
            if norm:
                allr = store["rew"][:filled]
                rew = (rew - allr.mean()) / allr.std()
            cast = lambda x: x.transpose(1, 0, 2)
            want = [cast(store["obs"][inds]), store["share"][inds], cast(store["acts"][inds]), cast(rew), cast(store["nobs"][inds]), store["nshare"][inds],
                    cast(store["dones"][inds]), store["dones_env"][inds], cast(store["valid"][inds]),
                    cast(store["avail"][inds]) if avail else None, cast(store["navail"][inds]) if avail else None]
            for i, w in enumerate(want):
                got = smp[i]["policy_0"]
                if w is None:
                    assert got is None
                elif i == 3 and norm:
                    assert np.abs(got - w).max() <= 2e-5 * max(1.0, np.abs(w).max()), (k, i)
                else:
                    assert np.array_equal(got, w), (k, i)
