"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python tests/golden/make_goldens.py

The reference has no tests or golden vectors of its own; these fixtures are
"outputs of the reference itself run here".  Inputs are synthetic (oracle.qmix.synth_batch);
weights are the reference's own init, perturbed so every LayerNorm gain/bias and every
Linear bias is non-trivial and the target nets differ from the live nets.
Every array the CUDA path must reproduce is stored: sampled indices, loss, grad_norm, Q_tot,
clipped grads, post-Adam params, post-polyak targets, PER weights/priorities/tree leaves.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_harness as rh  # noqa: E402
from oracle.qmix import QmixConfig, synth_batch, randomize_all  # noqa: E402


def build_reference_qmix(cfg, extra_flags=(), T=8, seed=1):
    rh.import_reference()
    from offpolicy.algorithms.qmix.qmix import QMix
    from offpolicy.algorithms.qmix.algorithm.QMixPolicy import QMixPolicy
    sp = rh.gym_spaces()
    flags = ["--algorithm_name", "qmix", "--hidden_size", str(cfg.hidden), "--gain", str(cfg.gain),
             "--hypernet_layers", str(cfg.hyper_layers), "--lr", str(cfg.lr)] + list(extra_flags)
    args = rh.make_args(flags)
    torch.manual_seed(seed)
    np.random.seed(seed)
    info = dict(obs_space=[cfg.obs_dim], share_obs_space=[cfg.state_dim], act_space=sp.Discrete(cfg.act_dim),
                cent_obs_dim=cfg.state_dim, cent_act_dim=cfg.act_dim * cfg.n_agents)
    dev = torch.device("cpu")
    pol = QMixPolicy({"args": args, "device": dev}, info)
    tr = QMix(args, cfg.n_agents, {"policy_0": pol}, lambda a: "policy_0", device=dev, episode_length=T)
    # non-trivial weights everywhere; targets != live
    randomize_all(pol.q_network, 11)
    randomize_all(tr.mixer, 12)
    tr.hard_target_updates()
    randomize_all(tr.target_policies["policy_0"].q_network, 13, scale=0.05)
    randomize_all(tr.target_mixer, 14, scale=0.05)
    return args, pol, tr


def sd_np(prefix, module):
    return {prefix + k: v.detach().numpy().copy() for k, v in module.state_dict().items()}


def to_ref_batch(b, weights=None, idx=None):
    obs, share, acts, rew, dones, dones_env, avail = b
    d = lambda x: {"policy_0": x}
    return (d(obs), d(share), d(acts), d(rew), d(dones), d(dones_env), d(avail), weights, idx)


def gen_qmix(name, cfg, flags=(), B=4, T=8, steps=2, avail_p=0.7, var_len=True, per=False):
    args, pol, tr = build_reference_qmix(cfg, flags, T)
    out = {}
    out.update(sd_np("init.agent.", pol.q_network))
    out.update(sd_np("init.mixer.", tr.mixer))
    out.update(sd_np("init.tgt_agent.", tr.target_policies["policy_0"].q_network))
    out.update(sd_np("init.tgt_mixer.", tr.target_mixer))
    for s in range(steps):
        b = synth_batch(cfg, B, T, seed=100 + s, avail_p=avail_p, var_len=var_len)
        for k, v in zip(["obs", "share", "acts", "rew", "dones", "dones_env", "avail"], b):
            out["s%d.in.%s" % (s, k)] = v
        w = idx = None
        if per:
            w = np.random.RandomState(7 + s).rand(B).astype(np.float64) * 0.9 + 0.1
            idx = np.arange(B)
            out["s%d.in.weights" % s] = w
        info, prio, _ = tr.train_policy_on_batch(to_ref_batch(b, w, idx))
        out["s%d.loss" % s] = info["loss"].detach().numpy()
        out["s%d.grad_norm" % s] = np.asarray(float(info["grad_norm"]), np.float32)
        out["s%d.Q_tot" % s] = info["Q_tot"].detach().numpy()
        if prio is not None:
            out["s%d.prio" % s] = np.asarray(prio)
        names = [k for k, _ in pol.q_network.named_parameters()]
        for k, p in pol.q_network.named_parameters():
            if p.grad is not None:
                out["s%d.grad.agent.%s" % (s, k)] = p.grad.numpy().copy()
        for k, p in tr.mixer.named_parameters():
            out["s%d.grad.mixer.%s" % (s, k)] = p.grad.numpy().copy()
        tr.soft_target_updates()
        out.update(sd_np("s%d.agent." % s, pol.q_network))
        out.update(sd_np("s%d.mixer." % s, tr.mixer))
        out.update(sd_np("s%d.tgt_agent." % s, tr.target_policies["policy_0"].q_network))
        out.update(sd_np("s%d.tgt_mixer." % s, tr.target_mixer))
    out["meta.cfg"] = np.array([cfg.n_agents, cfg.obs_dim, cfg.act_dim, cfg.state_dim, cfg.hidden, cfg.mixer_hidden,
                                cfg.hyper_hidden, cfg.hyper_layers, B, T, steps])
    out["meta.flags"] = np.array([args.use_double_q, args.use_huber_loss, per, bool(args.prev_act_inp), not args.use_feature_normalization, not args.use_ReLU], dtype=np.int64)
    out["meta.hparams"] = np.array([args.gamma, args.lr, args.opti_eps, args.max_grad_norm, args.tau, args.huber_delta,
                                    args.per_nu, args.per_eps], dtype=np.float64)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024), "loss", out["s0.loss"])


def gen_replay(name="replay_small"):
    """Uniform + PER sampling goldens from the reference buffers (indices, gathered bytes, IS weights, leaves)."""
    rh.import_reference()
    from offpolicy.utils.rec_buffer import RecReplayBuffer, PrioritizedRecReplayBuffer
    sp = rh.gym_spaces()
    N, O, A, S, T, E = 3, 6, 5, 7, 4, 16
    info = {"policy_0": dict(obs_space=[O], share_obs_space=[S], act_space=sp.Discrete(A))}
    agents = {"policy_0": list(range(N))}
    out = {"meta": np.array([N, O, A, S, T, E])}
    rs = np.random.RandomState(3)

    def episodes(n):
        d = lambda x: {"policy_0": x.astype(np.float32)}
        de = (rs.rand(T, n, 1) < 0.2).astype(np.float32)
        de = np.maximum.accumulate(de, axis=0)
        return (d(rs.randn(T + 1, n, N, O)), d(np.repeat(rs.randn(T + 1, n, 1, S), N, axis=2)),
                d(np.eye(A)[rs.randint(0, A, (T, n, N))]), d(np.repeat(rs.randn(T, n, 1, 1), N, axis=2)),
                d(np.repeat(de[:, :, None], N, axis=2)), d(de), d((rs.rand(T + 1, n, N, A) < 0.6)))

    for norm in (False, True):
        tag = "norm" if norm else "plain"
        buf = RecReplayBuffer(info, agents, E, T, True, True, use_reward_normalization=norm)
        ins = []
        for n_ep in (5, 7, 9):          # third insert wraps the ring (5+7+9 = 21 > 16)
            ep = episodes(n_ep)
            ins.append(ep)
            r = buf.insert(n_ep, *ep)
            out["%s.idx_range%d" % (tag, len(ins))] = np.asarray(r)
        for j, ep in enumerate(ins):
            for k, f in zip(["obs", "share", "acts", "rew", "dones", "dones_env", "avail"], ep):
                out["%s.ins%d.%s" % (tag, j, k)] = f["policy_0"]
        np.random.seed(123)
        for d in range(3):
            state_pos = np.random.get_state()[2]
            smp = buf.sample(6)
            for k, f in zip(["obs", "share", "acts", "rew", "dones", "dones_env", "avail"], smp[:7]):
                out["%s.draw%d.%s" % (tag, d, k)] = np.ascontiguousarray(f["policy_0"])
        # the indices themselves (reference does not return them): replay the stream
        np.random.seed(123)
        out["%s.inds" % tag] = np.stack([np.random.choice(len(buf), 6) for _ in range(3)])

    # PER: prime leaves through update_priorities (insert priming is broken in the reference, App. D-2)
    per = PrioritizedRecReplayBuffer(0.6, info, agents, E, T, True, True)
    ep = episodes(6)
    per.insert(6, *ep)
    ep2 = episodes(6)
    per.insert(6, *ep2)
    pr = (rs.rand(12) * 3 + 0.05).astype(np.float32)
    per.update_priorities(np.arange(12), pr, "policy_0")
    out["per.prio0"] = pr
    out["per.leaves0"] = per._it_sums["policy_0"]._value.copy()
    out["per.minleaves0"] = per._it_mins["policy_0"]._value.copy()
    np.random.seed(77)
    smp = per.sample(5, 0.4, "policy_0")
    out["per.w0"], out["per.idx0"] = smp[7], smp[8]
    # duplicate-index write-back: last write wins
    idx = np.array([3, 5, 3, 7, 5, 3])
    pr2 = np.array([0.5, 1.5, 2.5, 0.25, 4.0, 0.125], np.float32)
    per.update_priorities(idx, pr2, "policy_0")
    out["per.upd_idx"], out["per.upd_prio"] = idx, pr2
    out["per.leaves1"] = per._it_sums["policy_0"]._value.copy()
    out["per.minleaves1"] = per._it_mins["policy_0"]._value.copy()
    out["per.maxprio1"] = np.asarray(per.max_priorities["policy_0"])
    smp = per.sample(8, 0.7, "policy_0")
    out["per.w1"], out["per.idx1"] = smp[7], smp[8]
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__" and "maddpg" not in sys.argv[1:] and "rollout" not in sys.argv[1:] and "prev_act" not in sys.argv[1:] and "mqmix" not in sys.argv[1:] and "mlp_replay" not in sys.argv[1:] and "nofn" not in sys.argv[1:] and "tanh" not in sys.argv[1:] and "multi" not in sys.argv[1:]:
    torch.set_num_threads(1)
    small = QmixConfig(n_agents=3, obs_dim=30, act_dim=9, state_dim=48)
    gen_qmix("qmix_small", small)
    gen_qmix("qmix_small_huber_nodq", small, flags=["--use_huber_loss", "--use_double_q", "--huber_delta", "0.5"], steps=1)
    gen_qmix("qmix_small_per", small, flags=["--use_per"], per=True, steps=1)
    gen_qmix("qmix_small_hyper1", QmixConfig(n_agents=3, obs_dim=30, act_dim=9, state_dim=48, hyper_layers=1), steps=1)
    gen_qmix("qmix_5ag", QmixConfig(n_agents=5, obs_dim=17, act_dim=11, state_dim=23), B=3, T=6, steps=1)
    gen_replay()


# ---------------------------------------------------------------------------------------------------------------
# recurrent MADDPG / MATD3 (Box actions; shared centralised observation)
# ---------------------------------------------------------------------------------------------------------------
def gen_maddpg(name, cfg, flags=(), B=4, T=6, steps=2, per=False, use_avail=False):
    from oracle.maddpg import synth_batch_cont, synth_batch_disc, synth_avail
    rh.import_reference()
    sp = rh.gym_spaces()
    algo = "rmatd3" if cfg.td3 else "rmaddpg"
    args = rh.make_args(["--algorithm_name", algo, "--hidden_size", str(cfg.hidden), "--gain", str(cfg.gain), "--lr", str(cfg.lr)] + list(flags))
    if cfg.td3:
        from offpolicy.algorithms.r_matd3.algorithm.rMATD3Policy import R_MATD3Policy as Policy
        from offpolicy.algorithms.r_matd3.r_matd3 import R_MATD3 as Trainer
    else:
        from offpolicy.algorithms.r_maddpg.algorithm.rMADDPGPolicy import R_MADDPGPolicy as Policy
        from offpolicy.algorithms.r_maddpg.r_maddpg import R_MADDPG as Trainer
    torch.manual_seed(1)
    np.random.seed(1)
    info = dict(obs_space=sp.Box(-np.inf, np.inf, (cfg.obs_dim,)), share_obs_space=sp.Box(-np.inf, np.inf, (cfg.state_dim,)),
                act_space=sp.Discrete(cfg.act_dim) if cfg.discrete else sp.Box(-1.0, 1.0, (cfg.act_dim,)), cent_obs_dim=cfg.state_dim,
                cent_act_dim=cfg.act_dim * cfg.n_agents)
    dev = torch.device("cpu")
    pol = Policy({"args": args, "device": dev}, info)
    tr = Trainer(args, cfg.n_agents, {"policy_0": pol}, lambda a: "policy_0", device=dev, episode_length=T)
    randomize_all(pol.actor, 21); randomize_all(pol.critic, 22)
    pol.hard_target_updates()
    randomize_all(pol.target_actor, 23, scale=0.05); randomize_all(pol.target_critic, 24, scale=0.05)
    out = {}
    for tag, mod in (("actor", pol.actor), ("critic", pol.critic), ("tgt_actor", pol.target_actor), ("tgt_critic", pol.target_critic)):
        out.update(sd_np("init.%s." % tag, mod))
    # rollout-time get_actions (one env step): greedy and exploring, seeded (rMADDPGPolicy.py:62-128)
    rs = np.random.RandomState(77)
    a_obs = rs.randn(cfg.n_agents * B, cfg.obs_dim).astype(np.float32)
    a_h = (0.3 * rs.randn(cfg.n_agents * B, cfg.hidden)).astype(np.float32)
    out["act.in.obs"], out["act.in.h"] = a_obs, a_h
    with torch.no_grad():
        a, h2, _ = pol.get_actions(a_obs, None, torch.from_numpy(a_h), explore=False)
        out["act.greedy"], out["act.new_h"] = np.asarray(a), h2.numpy()
        torch.manual_seed(5); np.random.seed(5)
        a, _, eps = pol.get_actions(a_obs, None, torch.from_numpy(a_h), t_env=20000, explore=True)
        out["act.explore"] = np.asarray(a, dtype=np.float32)
        if eps is not None:
            out["act.eps"] = np.asarray(eps, np.float64)
        torch.manual_seed(6); np.random.seed(6)
        out["act.random"] = np.asarray(pol.get_random_actions(a_obs), dtype=np.float32)
    for s in range(steps):
        b = (synth_batch_disc if cfg.discrete else synth_batch_cont)(cfg, B, T, seed=200 + s)
        for k, v in zip(["obs", "share", "acts", "rew", "dones", "dones_env"], b[:6]):
            out["s%d.in.%s" % (s, k)] = v
        w = idx = None
        if per:
            w = np.random.RandomState(9 + s).rand(B) * 0.9 + 0.1
            idx = np.arange(B)
            out["s%d.in.weights" % s] = w
        d = lambda x: {"policy_0": x}
        av = synth_avail(cfg, B, T, seed=300 + s) if use_avail else None
        if use_avail:
            out["s%d.in.avail" % s] = av
        batch = (d(b[0]), d(b[1]), d(b[2]), d(b[3]), d(b[4]), d(b[5]), d(av), w, idx)
        torch.manual_seed(1000 + s)
        if cfg.discrete:
            # replay the reference's own draws: target actor (MATD3 only), then the actor update's Gumbel-softmax
            from offpolicy.utils.util import sample_gumbel as ref_gumbel
            if cfg.td3:
                out["s%d.in.noise" % s] = ref_gumbel((T + 1, cfg.n_agents * B, cfg.act_dim)).numpy()
            if tr.num_updates["policy_0"] % tr.actor_update_interval == 0:
                out["s%d.in.actor_noise" % s] = ref_gumbel((T, cfg.n_agents * B, cfg.act_dim)).numpy()
            torch.manual_seed(1000 + s)
        elif cfg.td3:
            out["s%d.in.noise" % s] = torch.empty(T + 1, cfg.n_agents * B, cfg.act_dim).normal_(mean=0, std=float(args.target_action_noise_std)).numpy()
            torch.manual_seed(1000 + s)
        info_t, prio, _ = tr.shared_train_policy_on_batch("policy_0", batch)
        out["s%d.critic_loss" % s] = info_t["critic_loss"].detach().numpy()
        out["s%d.critic_grad_norm" % s] = np.asarray(float(info_t["critic_grad_norm"]), np.float32)
        out["s%d.update_actor" % s] = np.asarray(int(info_t["update_actor"]))
        if info_t["update_actor"]:
            out["s%d.actor_loss" % s] = info_t["actor_loss"].detach().numpy()
            out["s%d.actor_grad_norm" % s] = np.asarray(float(info_t["actor_grad_norm"]), np.float32)
            for k, p in pol.actor.named_parameters():
                if p.grad is not None:
                    out["s%d.grad.actor.%s" % (s, k)] = p.grad.numpy().copy()
        if prio is not None:
            out["s%d.prio" % s] = np.asarray(prio)
        if info_t["update_actor"]:
            pol.soft_target_updates()           # runner: base_runner.py:250-252
        if s == steps - 1:      # parameters only after the last step (keeps the fixture small)
            for tag, mod in (("actor", pol.actor), ("critic", pol.critic), ("tgt_actor", pol.target_actor), ("tgt_critic", pol.target_critic)):
                out.update(sd_np("final.%s." % tag, mod))
    out["meta.cfg"] = np.array([cfg.n_agents, cfg.obs_dim, cfg.act_dim, cfg.state_dim, cfg.hidden, B, T, steps, int(cfg.td3), int(per),
                                int(cfg.discrete), int(not args.use_feature_normalization), int(not args.use_ReLU)])
    out["meta.hparams"] = np.array([args.gamma, args.lr, args.opti_eps, args.max_grad_norm, args.tau, args.huber_delta, args.per_nu,
                                    args.per_eps, float(args.target_action_noise_std), args.weight_decay], dtype=np.float64)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024), "critic_loss", out["s0.critic_loss"])


def main_maddpg():
    from oracle.maddpg import MaddpgConfig
    gen_maddpg("maddpg_box", MaddpgConfig())
    gen_maddpg("matd3_box", MaddpgConfig(td3=True, actor_update_interval=2))
    gen_maddpg("maddpg_box_per", MaddpgConfig(use_per=True), flags=["--use_per"], per=True, steps=1)
    gen_maddpg("maddpg_disc", MaddpgConfig(act_dim=5, discrete=True))
    gen_maddpg("matd3_disc", MaddpgConfig(act_dim=5, discrete=True, td3=True, actor_update_interval=2), steps=3)
    gen_maddpg("matd3_disc_avail", MaddpgConfig(act_dim=5, discrete=True, td3=True, actor_update_interval=2), steps=3, use_avail=True)


if __name__ == "__main__" and "maddpg" in sys.argv[1:]:
    main_maddpg()


# ---------------------------------------------------------------------------------------------------------------
# rollout-time QMixPolicy surface (one env step per call): get_actions greedy / exploring, get_random_actions,
# get_q_values on a short sequence (QMixPolicy.py:42-191)
# ---------------------------------------------------------------------------------------------------------------
def gen_qmix_rollout(name="qmix_rollout"):
    cfg = QmixConfig(n_agents=3, obs_dim=30, act_dim=9, state_dim=48)
    args, pol, tr = build_reference_qmix(cfg, (), 8)
    out = {}
    out.update(sd_np("init.agent.", pol.q_network))
    rs = np.random.RandomState(91)
    R, steps = cfg.n_agents, 4
    obs = rs.randn(steps, R, cfg.obs_dim).astype(np.float32)
    avail = (rs.rand(steps, R, cfg.act_dim) < 0.6).astype(np.float32)
    avail[:, :, 0] = 1.0
    out["in.obs"], out["in.avail"] = obs, avail
    with torch.no_grad():
        h = np.zeros((R, cfg.hidden), np.float32)
        for t in range(steps):      # the runner's loop: the state returned by one call is handed to the next (smac_runner.py:73-98)
            a, h2, gq = pol.get_actions(obs[t], None, h, avail[t])
            out["greedy%d.actions" % t], out["greedy%d.h" % t], out["greedy%d.q" % t] = np.asarray(a, np.float32), h2.numpy().copy(), gq.numpy().copy()
            h = h2.numpy()
        q_seq, h_seq = pol.get_q_values(obs, None, torch.zeros(R, cfg.hidden))
        out["seq.q"], out["seq.h"] = q_seq.numpy().copy(), h_seq.numpy().copy()
        for tag, av in (("explore", avail[0]), ("explore_noavail", None)):
            torch.manual_seed(5); np.random.seed(5)
            a, h2, gq = pol.get_actions(obs[0], None, np.zeros((R, cfg.hidden), np.float32), av, t_env=20000, explore=True)
            out[tag + ".actions"], out[tag + ".q"] = np.asarray(a, np.float32), gq.numpy().copy()
        torch.manual_seed(6); np.random.seed(6)
        out["random.actions"] = np.asarray(pol.get_random_actions(obs[0], avail[0]), np.float32)
        torch.manual_seed(6); np.random.seed(6)
        out["random_noavail.actions"] = np.asarray(pol.get_random_actions(obs[0]), np.float32)
    out["meta.cfg"] = np.array([cfg.n_agents, cfg.obs_dim, cfg.act_dim, cfg.hidden, steps])
    out["meta.eps"] = np.array([args.epsilon_start, args.epsilon_finish, args.epsilon_anneal_time], np.float64)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__" and "rollout" in sys.argv[1:]:
    torch.set_num_threads(1)
    gen_qmix_rollout()


if __name__ == "__main__" and "prev_act" in sys.argv[1:]:
    torch.set_num_threads(1)
    gen_qmix("qmix_small_prev_act", QmixConfig(n_agents=3, obs_dim=30, act_dim=9, state_dim=48), flags=["--prev_act_inp"], steps=2)


# ---------------------------------------------------------------------------------------------------------------
# MLP (transition-level) QMIX: M_QMixPolicy + M_QMix (algorithms/mqmix/*), batches of single transitions
# ---------------------------------------------------------------------------------------------------------------
def synth_transitions(cfg, B, seed, avail=True):
    rs = np.random.RandomState(seed)
    N, O, A, S = cfg.n_agents, cfg.obs_dim, cfg.act_dim, cfg.state_dim
    av = (rs.rand(N, B, A) < 0.6).astype(np.float32); av[..., 0] = 1.0
    nav = (rs.rand(N, B, A) < 0.6).astype(np.float32); nav[..., 0] = 1.0
    logits = rs.rand(N, B, A) + 10.0 * av
    acts = np.eye(A, dtype=np.float32)[logits.argmax(-1)]
    rew = np.repeat(rs.randn(1, B, 1).astype(np.float32), N, 0)
    return dict(obs=rs.randn(N, B, O).astype(np.float32), share=rs.randn(B, S).astype(np.float32), acts=acts, rew=rew,
                nobs=rs.randn(N, B, O).astype(np.float32), nshare=rs.randn(B, S).astype(np.float32), dones=np.zeros((N, B, 1), np.float32),
                dones_env=(rs.rand(B, 1) < 0.3).astype(np.float32), valid=np.ones((N, B, 1), np.float32),
                avail=av if avail else None, navail=nav if avail else None)


def gen_mqmix(name, cfg, flags=(), B=8, steps=2, per=False, avail=True):
    rh.import_reference()
    from offpolicy.algorithms.mqmix.mqmix import M_QMix
    from offpolicy.algorithms.mqmix.algorithm.mQMixPolicy import M_QMixPolicy
    sp = rh.gym_spaces()
    args = rh.make_args(["--algorithm_name", "mqmix", "--hidden_size", str(cfg.hidden), "--gain", str(cfg.gain),
                         "--hypernet_layers", str(cfg.hyper_layers), "--lr", str(cfg.lr)] + list(flags))
    torch.manual_seed(1); np.random.seed(1)
    info = dict(obs_space=sp.Box(-np.inf, np.inf, (cfg.obs_dim,)), share_obs_space=sp.Box(-np.inf, np.inf, (cfg.state_dim,)),
                act_space=sp.Discrete(cfg.act_dim), cent_obs_dim=cfg.state_dim, cent_act_dim=cfg.act_dim * cfg.n_agents)
    dev = torch.device("cpu")
    pol = M_QMixPolicy({"args": args, "device": dev}, info)
    tr = M_QMix(args, cfg.n_agents, {"policy_0": pol}, lambda a: "policy_0", device=dev)
    randomize_all(pol.q_network, 31); randomize_all(tr.mixer, 32)
    tr.hard_target_updates()
    randomize_all(tr.target_policies["policy_0"].q_network, 33, scale=0.05); randomize_all(tr.target_mixer, 34, scale=0.05)
    out = {}
    out.update(sd_np("init.agent.", pol.q_network)); out.update(sd_np("init.mixer.", tr.mixer))
    out.update(sd_np("init.tgt_agent.", tr.target_policies["policy_0"].q_network)); out.update(sd_np("init.tgt_mixer.", tr.target_mixer))
    d = lambda x: {"policy_0": x}
    for s in range(steps):
        b = synth_transitions(cfg, B, 400 + s, avail)
        for k, v in b.items():
            if v is not None:
                out["s%d.in.%s" % (s, k)] = v
        w = idx = None
        if per:
            w = np.random.RandomState(17 + s).rand(B) * 0.9 + 0.1
            idx = np.arange(B)
            out["s%d.in.weights" % s] = w
        batch = (d(b["obs"]), d(b["share"]), d(b["acts"]), d(b["rew"]), d(b["nobs"]), d(b["nshare"]), d(b["dones"]), d(b["dones_env"]),
                 d(b["valid"]), d(b["avail"]), d(b["navail"]), w, idx)
        info_t, prio, _ = tr.train_policy_on_batch(batch, True)
        out["s%d.loss" % s] = info_t["loss"].detach().numpy()
        out["s%d.grad_norm" % s] = np.asarray(float(info_t["grad_norm"]), np.float32)
        out["s%d.Q_tot" % s] = info_t["Q_tot"].detach().numpy()
        if prio is not None:
            out["s%d.prio" % s] = np.asarray(prio)
        for k, p in pol.q_network.named_parameters():
            if p.grad is not None:
                out["s%d.grad.agent.%s" % (s, k)] = p.grad.numpy().copy()
        for k, p in tr.mixer.named_parameters():
            out["s%d.grad.mixer.%s" % (s, k)] = p.grad.numpy().copy()
        tr.soft_target_updates()
        out.update(sd_np("s%d.agent." % s, pol.q_network)); out.update(sd_np("s%d.mixer." % s, tr.mixer))
        out.update(sd_np("s%d.tgt_agent." % s, tr.target_policies["policy_0"].q_network)); out.update(sd_np("s%d.tgt_mixer." % s, tr.target_mixer))
    # rollout surface (one env step): greedy and exploring under fixed seeds
    rs = np.random.RandomState(5)
    r_obs = rs.randn(cfg.n_agents, cfg.obs_dim).astype(np.float32)
    r_av = (rs.rand(cfg.n_agents, cfg.act_dim) < 0.6).astype(np.float32); r_av[:, 0] = 1.0
    out["roll.obs"], out["roll.avail"] = r_obs, r_av
    with torch.no_grad():
        a, q = pol.get_actions(r_obs, r_av)
        out["roll.greedy"], out["roll.greedy_q"] = np.asarray(a, np.float32), q.numpy().copy()
        torch.manual_seed(5); np.random.seed(5)
        a, q = pol.get_actions(r_obs, r_av, t_env=20000, explore=True)
        out["roll.explore"] = np.asarray(a, np.float32)
        out["roll.q_all"] = pol.get_q_values(torch.from_numpy(r_obs)).numpy().copy()
    out["meta.cfg"] = np.array([cfg.n_agents, cfg.obs_dim, cfg.act_dim, cfg.state_dim, cfg.hidden, cfg.mixer_hidden, cfg.hyper_hidden,
                                cfg.hyper_layers, B, 1, steps])
    out["meta.flags"] = np.array([args.use_double_q, args.use_huber_loss, per, False, not args.use_feature_normalization, not args.use_ReLU], dtype=np.int64)
    out["meta.hparams"] = np.array([args.gamma, args.lr, args.opti_eps, args.max_grad_norm, args.tau, args.huber_delta, args.per_nu, args.per_eps],
                                   dtype=np.float64)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024), "loss", out["s0.loss"])


if __name__ == "__main__" and "mqmix" in sys.argv[1:]:
    torch.set_num_threads(1)
    small = QmixConfig(n_agents=3, obs_dim=18, act_dim=5, state_dim=54)
    gen_mqmix("mqmix_small", small)
    gen_mqmix("mqmix_small_per_huber_nodq", small, flags=["--use_per", "--use_huber_loss", "--use_double_q", "--huber_delta", "0.5"], per=True, steps=1)
    gen_mqmix("mqmix_small_noavail", small, avail=False, steps=1)


# ---------------------------------------------------------------------------------------------------------------
# transition replay: the reference's MlpReplayBuffer (utils/mlp_buffer.py) driven with seeded inserts and samples
# ---------------------------------------------------------------------------------------------------------------
from oracle.mqmix import transition_replay_script as mlp_replay_script  # noqa: E402


def gen_mlp_replay():
    rh.import_reference()
    from offpolicy.utils.mlp_buffer import MlpReplayBuffer
    N, O, A, S, E = 3, 6, 4, 7, 20
    out = {"meta.shape": np.array([N, O, A, S, E])}
    for tag, norm, avail in (("plain", False, True), ("norm", True, False)):
        info = {"policy_0": dict(obs_space=[O], share_obs_space=[S], act_space=rh.gym_spaces().Discrete(A))}
        buf = MlpReplayBuffer(info, {"policy_0": list(range(N))}, E, True, avail, use_reward_normalization=norm)
        np.random.seed(11)
        d = lambda x: {"policy_0": x}
        ns = ni = 0
        for op, n, f in mlp_replay_script():
            if op == "insert":
                idx = buf.insert(n, d(f["obs"]), d(f["share"]), d(f["acts"]), d(f["rew"]), d(f["nobs"]), d(f["nshare"]), d(f["dones"]), d(f["dones_env"]),
                                 d(f["valid"]), d(f["avail"]), d(f["navail"]))
                out["%s.idx%d" % (tag, ni)] = np.asarray(idx)
                ni += 1
            else:
                smp = buf.sample(n)
                for i, name in enumerate(["obs", "share", "acts", "rew", "nobs", "nshare", "dones", "dones_env", "valid", "avail", "navail"]):
                    if smp[i]["policy_0"] is not None:
                        out["%s.s%d.%s" % (tag, ns, name)] = np.asarray(smp[i]["policy_0"])
                ns += 1
        out["%s.n_samples" % tag] = np.array(ns)
    path = os.path.join(HERE, "mlp_replay_small.npz")
    np.savez_compressed(path, **out)
    print("mlp_replay_small ->", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__" and "mlp_replay" in sys.argv[1:]:
    gen_mlp_replay()


if __name__ == "__main__" and "tanh" in sys.argv[1:]:
    # --use_ReLU is a store_false flag (config.py): passing it makes the fc blocks Linear -> Tanh -> LayerNorm (mlp.py:12,19-22)
    torch.set_num_threads(1)
    from oracle.maddpg import MaddpgConfig
    gen_qmix("qmix_small_tanh", QmixConfig(n_agents=3, obs_dim=30, act_dim=9, state_dim=48, relu=False), flags=["--use_ReLU"], steps=2)
    gen_mqmix("mqmix_small_tanh", QmixConfig(n_agents=3, obs_dim=18, act_dim=5, state_dim=54, relu=False), flags=["--use_ReLU"], steps=1)
    gen_maddpg("maddpg_box_tanh", MaddpgConfig(relu=False), flags=["--use_ReLU"], steps=2)


if __name__ == "__main__" and "nofn" in sys.argv[1:]:
    # --use_feature_normalization is a store_false flag (config.py): passing it switches the input LayerNorm off
    torch.set_num_threads(1)
    gen_qmix("qmix_small_nofn", QmixConfig(n_agents=3, obs_dim=30, act_dim=9, state_dim=48, feature_norm=False), flags=["--use_feature_normalization"], steps=2)
    from oracle.maddpg import MaddpgConfig
    gen_maddpg("matd3_disc_nofn", MaddpgConfig(act_dim=5, discrete=True, td3=True, actor_update_interval=2, feature_norm=False), flags=["--use_feature_normalization"], steps=2)
    gen_mqmix("mqmix_small_nofn", QmixConfig(n_agents=3, obs_dim=18, act_dim=5, state_dim=54, feature_norm=False), flags=["--use_feature_normalization"], steps=1)


# ---------------------------------------------------------------------------------------------------------
# several policies (share_policy = False; scripts/train_mpe_rmaddpg.sh:14 -> train/train_mpe.py:139-150): one policy per agent,
# heterogeneous observation / action widths (simple_speaker_listener: speaker obs 3 / Discrete(3), listener obs 11 / Discrete(5))
# ---------------------------------------------------------------------------------------------------------
def gen_maddpg_multi(name, specs, state_dim, td3=False, discrete=True, B=4, T=6, rounds=2):
    """specs: [(obs_dim, act_dim)] per agent = per policy.  One round = the runner's batch_train (base_runner.py:225-256): every policy
    updated once from the SAME sample, then soft updates of all policies when the actor was updated."""
    rh.import_reference()
    sp = rh.gym_spaces()
    algo = "rmatd3" if td3 else "rmaddpg"
    args = rh.make_args(["--algorithm_name", algo, "--gain", "1.0", "--lr", "0.0005"])
    if td3:
        from offpolicy.algorithms.r_matd3.algorithm.rMATD3Policy import R_MATD3Policy as Policy
        from offpolicy.algorithms.r_matd3.r_matd3 import R_MATD3 as Trainer
    else:
        from offpolicy.algorithms.r_maddpg.algorithm.rMADDPGPolicy import R_MADDPGPolicy as Policy
        from offpolicy.algorithms.r_maddpg.r_maddpg import R_MADDPG as Trainer
    from offpolicy.utils.util import sample_gumbel as ref_gumbel
    torch.manual_seed(1)
    np.random.seed(1)
    N = len(specs)
    CA = sum(a for _, a in specs)
    dev = torch.device("cpu")
    pols = {}
    for i, (o, a) in enumerate(specs):
        info = dict(obs_space=sp.Box(-np.inf, np.inf, (o,)), share_obs_space=sp.Box(-np.inf, np.inf, (state_dim,)),
                    act_space=sp.Discrete(a) if discrete else sp.Box(-1.0, 1.0, (a,)), cent_obs_dim=state_dim, cent_act_dim=CA)
        pols["policy_%d" % i] = Policy({"args": args, "device": dev}, info)
    tr = Trainer(args, N, pols, lambda a: "policy_%d" % a, device=dev, episode_length=T)
    out = {}
    for i in range(N):
        pol = pols["policy_%d" % i]
        randomize_all(pol.actor, 21 + 10 * i); randomize_all(pol.critic, 22 + 10 * i)
        pol.hard_target_updates()
        randomize_all(pol.target_actor, 23 + 10 * i, scale=0.05); randomize_all(pol.target_critic, 24 + 10 * i, scale=0.05)
        for tag, mod in (("actor", pol.actor), ("critic", pol.critic), ("tgt_actor", pol.target_actor), ("tgt_critic", pol.target_critic)):
            out.update(sd_np("init.p%d.%s." % (i, tag), mod))
    for r in range(rounds):
        rs = np.random.RandomState(400 + r)
        share = rs.randn(T + 1, B, state_dim).astype(np.float32)
        length = rs.randint(T // 2, T + 1, size=B)
        de = (np.arange(T)[:, None] >= length[None, :] - 1).astype(np.float32)[..., None]            # env done from step len-1 on
        rew = rs.randn(T, B, 1).astype(np.float32)
        obs, acts, dones = {}, {}, {}
        for i, (o, a) in enumerate(specs):
            p = "policy_%d" % i
            obs[p] = rs.randn(1, T + 1, B, o).astype(np.float32)
            acts[p] = (np.eye(a, dtype=np.float32)[rs.randint(0, a, (1, T, B))] if discrete else rs.uniform(-1, 1, (1, T, B, a)).astype(np.float32))
            dones[p] = np.repeat(de[None], 1, 0).copy()
            out["r%d.in.p%d.obs" % (r, i)], out["r%d.in.p%d.acts" % (r, i)], out["r%d.in.p%d.dones" % (r, i)] = obs[p], acts[p], dones[p]
        out["r%d.in.share" % r], out["r%d.in.rew" % r], out["r%d.in.dones_env" % r] = share, rew, de
        pd = lambda v: {"policy_%d" % i: v for i in range(N)}
        batch = (obs, pd(share), acts, pd(np.repeat(rew[None], 1, 0)), dones, pd(de), pd(None), None, None)
        upd_any = False
        for i in range(N):
            p = "policy_%d" % i
            seed = 2000 + 10 * r + i
            # replay the reference's own draws for this update (r_maddpg.py:62-105 walks the policies in id order, then the actor update)
            torch.manual_seed(seed)
            if td3:
                for q, (oq, aq) in enumerate(specs):
                    shape = (T + 1, B, aq)
                    out["r%d.u%d.noise.p%d" % (r, i, q)] = (ref_gumbel(shape).numpy() if discrete else
                                                             torch.empty(*shape).normal_(mean=0, std=float(args.target_action_noise_std)).numpy())
            if discrete and tr.num_updates[p] % tr.actor_update_interval == 0:
                out["r%d.u%d.actor_noise" % (r, i)] = ref_gumbel((T, B, specs[i][1])).numpy()
            torch.manual_seed(seed)
            info_t, prio, _ = tr.shared_train_policy_on_batch(p, batch)
            pol = pols[p]
            out["r%d.u%d.critic_loss" % (r, i)] = info_t["critic_loss"].detach().numpy()
            out["r%d.u%d.critic_grad_norm" % (r, i)] = np.asarray(float(info_t["critic_grad_norm"]), np.float32)
            out["r%d.u%d.update_actor" % (r, i)] = np.asarray(int(info_t["update_actor"]))
            if info_t["update_actor"]:
                upd_any = True
                out["r%d.u%d.actor_loss" % (r, i)] = info_t["actor_loss"].detach().numpy()
                out["r%d.u%d.actor_grad_norm" % (r, i)] = np.asarray(float(info_t["actor_grad_norm"]), np.float32)
                for k, prm in pol.actor.named_parameters():
                    if prm.grad is not None:
                        out["r%d.u%d.grad.actor.%s" % (r, i, k)] = prm.grad.numpy().copy()
        if upd_any:
            for i in range(N):
                pols["policy_%d" % i].soft_target_updates()           # base_runner.py:250-252
    for i in range(N):
        pol = pols["policy_%d" % i]
        for tag, mod in (("actor", pol.actor), ("critic", pol.critic), ("tgt_actor", pol.target_actor), ("tgt_critic", pol.target_critic)):
            out.update(sd_np("final.p%d.%s." % (i, tag), mod))
    out["meta.cfg"] = np.array([N, state_dim, args.hidden_size, B, T, rounds, int(td3), int(discrete)])
    out["meta.specs"] = np.array(specs)
    out["meta.hparams"] = np.array([args.gamma, args.lr, args.opti_eps, args.max_grad_norm, args.tau, args.huber_delta, args.per_nu,
                                    args.per_eps, float(args.target_action_noise_std), args.weight_decay], dtype=np.float64)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024), "critic_loss", out["r0.u0.critic_loss"])


if __name__ == "__main__" and "multi" in sys.argv[1:]:
    gen_maddpg_multi("maddpg_multi_disc", [(3, 3), (11, 5)], state_dim=14)                       # simple_speaker_listener shapes (train_mpe_rmaddpg.sh)
    gen_maddpg_multi("matd3_multi_box", [(5, 2), (7, 3), (6, 2)], state_dim=18, td3=True, discrete=False, rounds=3)
    gen_maddpg_multi("matd3_multi_disc", [(4, 3), (6, 4)], state_dim=10, td3=True, discrete=True, rounds=2)
