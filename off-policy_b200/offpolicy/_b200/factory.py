"""Construction helpers for the drop-in classes outside a runner: what the reference's runner does in
runner/rnn/base_runner.py:110-178 (argparse namespace -> policy_info -> Policy -> Trainer -> Buffer), for callers that hold a plain
description of the workload instead of a parsed command line.  Used by bench.py and by the parity tests; no test or oracle code here.

`LearnerConfig` carries the reference's config.py defaults for the fields the learner path reads; any object with the same attribute
names works (the tests pass the oracle's own config dataclass)."""
import types
from dataclasses import dataclass

import numpy as np


@dataclass
class LearnerConfig:
    n_agents: int = 3
    obs_dim: int = 30
    act_dim: int = 9
    state_dim: int = 48
    hidden: int = 64                # config.py:63 hidden_size
    layer_n: int = 1                # config.py:67 layer_N
    mixer_hidden: int = 32          # config.py:101 mixer_hidden_dim
    hyper_hidden: int = 64          # config.py:103 hypernet_hidden_dim
    hyper_layers: int = 2           # config.py:105 hypernet_layers
    gamma: float = 0.99
    lr: float = 5e-4
    opti_eps: float = 1e-5
    max_grad_norm: float = 10.0
    tau: float = 0.005
    double_q: bool = True
    huber: bool = False
    huber_delta: float = 10.0
    use_per: bool = False
    per_nu: float = 0.9
    per_eps: float = 1e-6
    vdn: bool = False
    feature_norm: bool = True
    relu: bool = True
    prev_act_inp: bool = False
    gain: float = 0.01


@dataclass
class MaddpgLearnerConfig:
    """What a bench / example needs to describe an R-MADDPG / R-MATD3 learner (attribute names = the oracle's MaddpgConfig, so the CPU
    arm can be built from the same values; nothing here imports the oracle)."""
    n_agents: int = 3
    obs_dim: int = 18
    act_dim: int = 2
    state_dim: int = 54
    hidden: int = 64
    layer_n: int = 1
    feature_norm: bool = True
    relu: bool = True
    gamma: float = 0.99
    lr: float = 5e-4
    opti_eps: float = 1e-5
    weight_decay: float = 0.0
    max_grad_norm: float = 10.0
    tau: float = 0.005
    huber: bool = False
    huber_delta: float = 10.0
    use_per: bool = False
    per_nu: float = 0.9
    per_eps: float = 1e-6
    td3: bool = False
    target_noise: float = 0.2
    actor_update_interval: int = 1
    gain: float = 0.01
    discrete: bool = False


class Box(object):      # duck-typed gym.spaces.Box: what the policies read is .shape / .low / .high
    def __init__(self, d, low=-1.0, high=1.0):
        self.shape = (d,)
        self.low = np.full(d, low, np.float32)
        self.high = np.full(d, high, np.float32)


class Discrete(object):  # duck-typed gym.spaces.Discrete
    def __init__(self, n):
        self.n = n


def pd(x, p_id="policy_0"):
    """{policy_id: array}: the per-policy dict every buffer / trainer entry point of the reference takes."""
    return {p_id: x}


def qmix_args(cfg, B, **over):
    """The argparse namespace fields QMixPolicy / QMix / M_QMix read (config.py names), from a LearnerConfig-like object."""
    a = types.SimpleNamespace(
        hidden_size=cfg.hidden, layer_N=1, use_ReLU=bool(getattr(cfg, "relu", True)), use_feature_normalization=bool(cfg.feature_norm), use_orthogonal=True, gain=cfg.gain,
        use_conv1d=False, stacked_frames=1, use_rnn_layer=True, recurrent_N=1, prev_act_inp=bool(getattr(cfg, "prev_act_inp", False)), use_double_q=cfg.double_q,
        hypernet_layers=cfg.hyper_layers, mixer_hidden_dim=cfg.mixer_hidden, hypernet_hidden_dim=cfg.hyper_hidden, gamma=cfg.gamma,
        use_per=cfg.use_per, per_nu=cfg.per_nu, per_eps=cfg.per_eps, per_alpha=0.6, use_huber_loss=cfg.huber,
        huber_delta=cfg.huber_delta, max_grad_norm=cfg.max_grad_norm, lr=cfg.lr, opti_eps=cfg.opti_eps, weight_decay=0, tau=cfg.tau,
        use_popart=False, use_value_active_masks=False, use_same_share_obs=True, batch_size=B, episode_length=0,
        epsilon_start=1.0, epsilon_finish=0.05, epsilon_anneal_time=50000, use_available_actions=True)
    for k, v in over.items():
        setattr(a, k, v)
    return a


def _policy_info(cfg):
    return dict(obs_space=[cfg.obs_dim], share_obs_space=[cfg.state_dim], act_space=Discrete(cfg.act_dim),
                cent_obs_dim=cfg.state_dim, cent_act_dim=cfg.act_dim * cfg.n_agents)


def build_qmix(cfg, B, T, vdn=False, debug=False, **over):
    """(args, QMixPolicy, QMix) for recurrent QMIX / VDN.  debug=True also materialises the per-action Q values (parity tests) and keeps
    k_qhead / k_mix_core / k_qhead_bwd as separate launches; debug=False is the product configuration (fused k_mid)."""
    from offpolicy.algorithms.qmix.algorithm.QMixPolicy import QMixPolicy
    from offpolicy.algorithms.qmix.qmix import QMix
    from offpolicy._b200 import capi
    args = qmix_args(cfg, B, **over)
    pol = QMixPolicy({"args": args, "device": capi.device()}, _policy_info(cfg))
    tr = QMix(args, cfg.n_agents, {"policy_0": pol}, lambda a: "policy_0", device=capi.device(), episode_length=T, vdn=vdn)
    capi.lib().mx_qmix_set_debug(tr.handle, 1 if debug else 0)
    return args, pol, tr


def build_mqmix(cfg, B, debug=False):
    """(args, M_QMixPolicy, M_QMix): the transition-level (MLP) learner."""
    from offpolicy.algorithms.mqmix.algorithm.mQMixPolicy import M_QMixPolicy
    from offpolicy.algorithms.mqmix.mqmix import M_QMix
    from offpolicy._b200 import capi
    args = qmix_args(cfg, B)
    pol = M_QMixPolicy({"args": args, "device": capi.device()}, _policy_info(cfg))
    tr = M_QMix(args, cfg.n_agents, {"policy_0": pol}, lambda a: "policy_0", device=capi.device())
    capi.lib().mx_qmix_set_debug(tr.handle, 1 if debug else 0)
    return args, pol, tr


def make_rec_buffers(N, O, A, S, T, E, per_alpha=None, norm=False, rng="numpy", max_batch=32, avail=True):
    """RecReplayBuffer / PrioritizedRecReplayBuffer for one shared policy over N agents (rec_buffer.py:9-61, 243-270)."""
    from offpolicy.utils.rec_buffer import RecReplayBuffer, PrioritizedRecReplayBuffer
    info = {"policy_0": dict(obs_space=[O], share_obs_space=[S], act_space=Discrete(A))}
    agents = {"policy_0": list(range(N))}
    if per_alpha is None:
        return RecReplayBuffer(info, agents, E, T, True, avail, use_reward_normalization=norm, rng=rng, max_batch=max_batch)
    return PrioritizedRecReplayBuffer(per_alpha, info, agents, E, T, True, avail, use_reward_normalization=norm, rng=rng,
                                      max_batch=max_batch)


def maddpg_args(cfg, B):
    """Namespace fields R_MADDPGPolicy / R_MADDPG / R_MATD3 read, from a config object with the oracle's MaddpgConfig attribute names."""
    return types.SimpleNamespace(
        hidden_size=cfg.hidden, layer_N=1, use_ReLU=bool(cfg.relu), use_feature_normalization=bool(cfg.feature_norm), use_orthogonal=True, gain=cfg.gain,
        use_conv1d=False, stacked_frames=1, use_rnn_layer=True, recurrent_N=1, prev_act_inp=False, gamma=cfg.gamma, use_per=cfg.use_per,
        per_nu=cfg.per_nu, per_eps=cfg.per_eps, use_huber_loss=cfg.huber, huber_delta=cfg.huber_delta, max_grad_norm=cfg.max_grad_norm,
        lr=cfg.lr, opti_eps=cfg.opti_eps, weight_decay=cfg.weight_decay, tau=cfg.tau, use_popart=False, use_value_active_masks=False,
        use_same_share_obs=True, batch_size=B, episode_length=0, epsilon_start=1.0, epsilon_finish=0.05, epsilon_anneal_time=50000,
        act_noise_std=0.1, target_action_noise_std=cfg.target_noise)


def build_maddpg(cfg, B, T):
    """(args, policy, trainer) for R-MADDPG (cfg.td3 False) / R-MATD3 (True), one shared policy."""
    from offpolicy._b200 import capi
    if cfg.td3:
        from offpolicy.algorithms.r_matd3.algorithm.rMATD3Policy import R_MATD3Policy as Policy
        from offpolicy.algorithms.r_matd3.r_matd3 import R_MATD3 as Trainer
    else:
        from offpolicy.algorithms.r_maddpg.algorithm.rMADDPGPolicy import R_MADDPGPolicy as Policy
        from offpolicy.algorithms.r_maddpg.r_maddpg import R_MADDPG as Trainer
    args = maddpg_args(cfg, B)
    info = dict(obs_space=Box(cfg.obs_dim, -np.inf, np.inf), share_obs_space=Box(cfg.state_dim, -np.inf, np.inf),
                act_space=Discrete(cfg.act_dim) if cfg.discrete else Box(cfg.act_dim), cent_obs_dim=cfg.state_dim, cent_act_dim=cfg.act_dim * cfg.n_agents)
    pol = Policy({"args": args, "device": capi.device()}, info)
    tr = Trainer(args, cfg.n_agents, {"policy_0": pol}, lambda a: "policy_0", device=capi.device(), episode_length=T)
    return args, pol, tr
