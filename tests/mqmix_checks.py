"""MLP (transition-level) QMIX parity checks against goldens made by the reference's M_QMixPolicy / M_QMix (tests/golden/mqmix_*.npz,
make_goldens.py mqmix): shared by the emulated (CPU) and the GPU test modules.  Tolerances as in qmix_checks.py."""
import numpy as np
import torch

from helpers import load_golden, golden_cfg, sub, rel_err
import qmix_checks as qc
from replay_checks import Discrete


def build(cfg, B, debug=True):
    from offpolicy._b200 import factory
    return factory.build_mqmix(cfg, B, debug=debug)


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


def _to_dicts(b, w=None):
    d = lambda x: {"policy_0": x}
    B = b[0].shape[1]
    return tuple(d(x) for x in b) + (w, np.arange(B) if w is not None else None)


def check_vs_oracle(N=3, O=18, A=5, S=54, B=64, steps=2, avail=False, per=False, huber=False, double_q=True, vdn=False, hyper_layers=2, debug=False):
    """Any size: the MLP learner in lock-step with the pinned oracle (oracle/mqmix.py).  Defaults = scripts/train_mpe_mqmix.sh shapes
    (simple_spread: 3 agents, obs 18, Discrete(5), state 54, no availability masks)."""
    from oracle.qmix import QmixConfig, randomize_all
    from oracle.mqmix import MqmixLearner, synth_transitions
    cfg = QmixConfig(n_agents=N, obs_dim=O, act_dim=A, state_dim=S, gain=1.0, use_per=per, huber=huber, huber_delta=0.7, double_q=double_q, vdn=vdn,
                     hyper_layers=hyper_layers)
    L = MqmixLearner(cfg, seed=3)
    randomize_all(L.agent, 1)
    if not vdn:
        randomize_all(L.mixer, 2)
    L.sync_targets()
    randomize_all(L.tgt_agent, 3, 0.05)
    if not vdn:
        randomize_all(L.tgt_mixer, 4, 0.05)
    if vdn:
        from offpolicy.algorithms.mvdn.algorithm.mVDNPolicy import M_VDNPolicy as Pol
        from offpolicy.algorithms.mvdn.mvdn import M_VDN as Tr
    else:
        from offpolicy.algorithms.mqmix.algorithm.mQMixPolicy import M_QMixPolicy as Pol
        from offpolicy.algorithms.mqmix.mqmix import M_QMix as Tr
    from offpolicy._b200 import capi
    args = qc.make_args(cfg, B)
    info = dict(obs_space=[O], share_obs_space=[S], act_space=Discrete(A), cent_obs_dim=S, cent_act_dim=A * N)
    pol = Pol({"args": args, "device": capi.device()}, info)
    tr = Tr(args, N, {"policy_0": pol}, lambda a: "policy_0", device=capi.device())
    capi.lib().mx_qmix_set_debug(tr.handle, 1 if debug else 0)
    pol.q_network.load_state_dict(L.agent.state_dict())
    tr.target_q_network.load_state_dict(L.tgt_agent.state_dict())
    if not vdn:
        tr.mixer.load_state_dict(L.mixer.state_dict())
        tr.target_mixer.load_state_dict(L.tgt_mixer.state_dict())
    import kink
    for s in range(steps):
        b = synth_transitions(cfg, B, seed=50 + s, avail=avail)
        w = (np.random.RandomState(60 + s).rand(B) * 0.9 + 0.1) if per else None
        L0 = kink.snapshot(L)
        info_t, prio, _ = tr.train_policy_on_batch(_to_dicts(b, w), True)
        gv = {k: v.clone() for k, v in tr.grad_views().items()}
        ref, rprio, _ = L.step(b + (w, None))
        coef = min(1.0, cfg.max_grad_norm / (float(ref["grad_norm"]) + 1e-6))
        bad = qc.grad_failures(gv, coef, L, cfg, 1e-4)
        if bad:      # ReLU kink?  (tests/kink.py: B = 1000 transitions x 3 agents x 128 hidden units -- a pre-activation within round-off of zero is likely)
            masks = kink.engine_masks(tr, B, 1, N, mlp=True)
            (ref, rprio, _), flips, max_pre = kink.redo_with_engine_masks(L0, lambda LL: LL.step(b + (w, None)), masks)
            assert flips > 0 and max_pre < kink.KINK_TOL, (s, "gradient mismatch not explained by ReLU kinks", flips, max_pre, bad[:3])
            kink.adopt(L, L0)
            coef = min(1.0, cfg.max_grad_norm / (float(ref["grad_norm"]) + 1e-6))
            bad = qc.grad_failures(gv, coef, L, cfg, 1e-4)
            print("kink-aware comparison: %d ReLU unit(s) within %.1e of zero flipped" % (flips, max_pre))
        assert not bad, (s, bad[:4])
        tr.soft_target_updates()
        L.soft_update()
        for k in ("loss", "grad_norm", "Q_tot"):
            assert rel_err(info_t[k].cpu(), ref[k]) < 1e-4, (s, k, float(info_t[k]), float(ref[k]))
        if per:
            assert rel_err(np.asarray(prio), rprio) < 1e-4
        for k, v in pol.q_network.state_dict().items():
            assert float((v.cpu() - L.agent.state_dict()[k]).abs().max()) <= 5e-3 * cfg.lr * (s + 1) + 1e-7, (s, k)
        for k, v in tr.target_q_network.state_dict().items():
            assert float((v.cpu() - L.tgt_agent.state_dict()[k]).abs().max()) <= 1e-6, (s, k)


def fill_mlp_buffer(N, O, A, S, E, B, avail=False, per=False, norm=False, seed=0):
    from offpolicy.utils.mlp_buffer import MlpReplayBuffer, PrioritizedMlpReplayBuffer
    info = {"policy_0": dict(obs_space=[O], share_obs_space=[S], act_space=Discrete(A))}
    ag = {"policy_0": list(range(N))}
    buf = (PrioritizedMlpReplayBuffer(0.6, info, ag, E, True, avail, norm, max_batch=max(B, 256)) if per else
           MlpReplayBuffer(info, ag, E, True, avail, norm, max_batch=max(B, 256)))
    rs = np.random.RandomState(seed)
    d = lambda x: {"policy_0": None if x is None else x.astype(np.float32)}
    for c in range(0, E, 256):
        n = min(256, E - c)
        av = (rs.rand(n, N, A) < 0.6) * 1.0
        av[..., 0] = 1.0
        buf.insert(n, d(rs.randn(n, N, O)), d(rs.randn(n, S)), d(np.eye(A)[rs.randint(0, A, (n, N))]), d(np.repeat(rs.randn(n, 1, 1), N, 1)), d(rs.randn(n, N, O)),
                   d(rs.randn(n, S)), d(np.zeros((n, N, 1))), d((rs.rand(n, 1) < 0.2) * 1.0), d(np.ones((n, N, 1))), d(av if avail else None), d(av if avail else None))
    return buf


def check_step_graph_vs_eager(per=False, steps=3, B=32, E=300):
    """sample (device MT19937) -> M_QMix step [-> priority write-back] -> soft update: the captured whole-step sequence
    (StepGraph; on the emulator the same launch sequence re-run) leaves exactly the state the drop-in calls leave."""
    from oracle.qmix import QmixConfig
    from offpolicy._b200.graph import StepGraph
    N, O, A, S = 3, 18, 5, 54
    cfg = QmixConfig(n_agents=N, obs_dim=O, act_dim=A, state_dim=S, gain=1.0, use_per=per)
    outs = []
    for mode in ("eager", "graph"):
        torch.manual_seed(0); np.random.seed(0)
        buf = fill_mlp_buffer(N, O, A, S, E, B, avail=per, per=per)
        args, pol, tr = build(cfg, B, debug=False)
        buf.seed_device_rng(77)
        if mode == "eager":
            for s in range(steps):
                smp = buf.sample(B, 0.4 + 0.15 * s) if per else buf.sample(B)       # annealed exponent (base_runner.py:159-160)
                info, prio, idx = tr.train_policy_on_batch(smp, True)
                if per:
                    buf.update_priorities(idx, prio, "policy_0")
                tr.soft_target_updates()
        else:
            g = StepGraph(buf, tr, B, beta=0.4)
            for s in range(steps):
                g.launch(beta=0.4 + 0.15 * s)          # device-resident exponent: no re-capture (mx_replay_set_beta)
            g.synchronize()
            g.close()
        outs.append((tr.theta.clone().cpu(), tr.theta_tgt.clone().cpu(), tr.adam_m.clone().cpu()))
    # same kernels, same order, no float atomics: the results are expected to be bit-identical (and are on the emulator); the bound is kept
    # at round-off level like the recurrent twin of this test (tests/test_gpu_qmix.py)
    for a, b in zip(*outs):
        assert float((a - b).abs().max()) <= 1e-6 * float(a.abs().max()) + 1e-7
    assert float((outs[0][0] - outs[0][1]).abs().max()) > 0.0


def check_buffer_vs_reference_golden():
    """MlpReplayBuffer (HBM replay, length-1 episodes) driven with the schedule of tests/golden/mlp_replay_small.npz = the reference's own
    MlpReplayBuffer under the same inserts and the same NumPy seed: ring indices and every sampled array bit-equal; normalised rewards to
    2e-5 (running statistics kept in fp64 on the device instead of a rescan of all stored rewards)."""
    from oracle.mqmix import transition_replay_script
    from offpolicy.utils.mlp_buffer import MlpReplayBuffer
    g = load_golden("mlp_replay_small")
    N, O, A, S, E = [int(v) for v in g["meta.shape"]]
    names = ["obs", "share", "acts", "rew", "nobs", "nshare", "dones", "dones_env", "valid", "avail", "navail"]
    d = lambda x: {"policy_0": x}
    for tag, norm, avail in (("plain", False, True), ("norm", True, False)):
        info = {"policy_0": dict(obs_space=[O], share_obs_space=[S], act_space=Discrete(A))}
        buf = MlpReplayBuffer(info, {"policy_0": list(range(N))}, E, True, avail, use_reward_normalization=norm, max_batch=16)
        np.random.seed(11)
        ns = ni = 0
        for op, n, f in transition_replay_script():
            if op == "insert":
                idx = buf.insert(n, *[d(f[k]) for k in names])
                assert np.array_equal(idx, g["%s.idx%d" % (tag, ni)]), (tag, ni)
                ni += 1
                continue
            smp = buf.sample(n)
            for i, name in enumerate(names):
                key = "%s.s%d.%s" % (tag, ns, name)
                got = smp[i]["policy_0"]
                if key not in g:
                    assert got is None, key
                elif name == "rew" and norm:
                    assert np.abs(got - g[key]).max() <= 2e-5 * max(1.0, np.abs(g[key]).max()), key
                else:
                    assert np.array_equal(got, g[key]), key
            ns += 1
        assert ns == int(g["%s.n_samples" % tag])
