"""Drop-in `QMixPolicy` (reference: offpolicy/algorithms/qmix/algorithm/QMixPolicy.py).

The agent Q-network's parameters are views into the flat device vector the CUDA learner trains in place
(`q_network.state_dict()` keeps the reference key names).  Update-time methods
(get_q_values over sequences, q_values_from_actions, greedy actions_from_q) are executed inside the fused
learner kernels (csrc/agent_fwd.cu).  The rollout-time surface the runner calls once per env step -- single-step
forward + masked arg-max (QMixPolicy.py:95-191) -- is ONE launch of k_policy_step (csrc/rollout.cu, SURVEY.md
section 8(f).1) with the recurrent state resident on the device; the epsilon-greedy draws stay on the host because
the reference takes them from the process-global NumPy / torch generators.
"""
import ctypes as C

import numpy as np
import torch

from offpolicy._b200 import capi
from offpolicy._b200.flat import FlatModule, reference_style_init
from offpolicy._b200.host_util import LinearDecay, space_dim, is_discrete, onehot


def qmix_cfg_struct(args, n_agents, obs_dim, act_dim, state_dim, episode_len, max_batch, vdn=False, use_avail=True, world_size=1, mlp=False):
    return capi.QmixCfg(
        n_agents=n_agents, obs_dim=obs_dim, act_dim=act_dim, state_dim=state_dim, hidden=args.hidden_size,
        mixer_hidden=args.mixer_hidden_dim, hyper_hidden=args.hypernet_hidden_dim, hyper_layers=args.hypernet_layers,
        episode_len=episode_len, max_batch=max_batch, vdn=int(vdn), double_q=int(args.use_double_q),
        use_huber=int(args.use_huber_loss), use_per=int(args.use_per), use_avail=int(use_avail), world_size=world_size,
        gamma=args.gamma, huber_delta=args.huber_delta, per_nu=args.per_nu, per_eps=args.per_eps, lr=args.lr,
        adam_beta1=0.9, adam_beta2=0.999, adam_eps=args.opti_eps, max_grad_norm=args.max_grad_norm, tau=args.tau,
        prev_act_inp=0 if mlp else int(bool(getattr(args, "prev_act_inp", False))), mlp=int(bool(mlp)),
        no_feature_norm=0 if getattr(args, "use_feature_normalization", True) else 1, use_tanh=0 if getattr(args, "use_ReLU", True) else 1)


def param_entries(cfg):
    lib = capi.lib()
    total = C.c_int64()
    n = lib.mx_qmix_param_layout(C.byref(cfg), None, 0, C.byref(total))
    if n < 0:
        raise capi.MxError(lib.mx_last_error().decode())
    arr = (capi.ParamEntry * n)()
    lib.mx_qmix_param_layout(C.byref(cfg), arr, n, C.byref(total))
    return [(e.name.decode(), int(e.offset), int(e.rows), int(e.cols)) for e in arr], int(total.value)


class QMixPolicy(object):
    def __init__(self, config, policy_config, train=True):
        self.args = config["args"]
        self.device = config["device"]
        self.obs_space = policy_config["obs_space"]
        self.obs_dim = space_dim(self.obs_space)
        self.act_space = policy_config["act_space"]
        self.act_dim = space_dim(self.act_space)
        self.output_dim = self.act_dim
        self.hidden_size = self.args.hidden_size
        self.central_obs_dim = policy_config["cent_obs_dim"]
        self.discrete = is_discrete(self.act_space)
        self.multidiscrete = False
        self.prev_act_inp = bool(getattr(self.args, "prev_act_inp", False))
        for flag, want in (("use_rnn_layer", True), ("use_conv1d", False)):
            if getattr(self.args, flag, want) != want:
                raise NotImplementedError("B200 QMIX path requires %s=%s" % (flag, want))
        if getattr(self.args, "layer_N", 1) != 1 or getattr(self.args, "recurrent_N", 1) != 1:
            raise NotImplementedError("B200 QMIX path requires layer_N=1, recurrent_N=1")
        self.q_network_input_dim = self.obs_dim + self.act_dim if self.prev_act_inp else self.obs_dim       # QMixPolicy.py:29-32

        capi.lib()
        self.dev = capi.device()
        # agent-only layout (vdn=1 -> no mixer block); the trainer re-binds these views into its full vector
        cfg = qmix_cfg_struct(self.args, 1, self.obs_dim, self.act_dim, 1, 1, 1, vdn=True)
        entries, total = param_entries(cfg)
        self._entries = entries
        flat = torch.zeros(total, dtype=torch.float32, device=self.dev)
        self.q_network = FlatModule(flat, entries, "agent.")
        init = reference_style_init(entries, dict(hidden=self.hidden_size, obs_dim=self.q_network_input_dim, act_dim=self.act_dim),
                                    gain=self.args.gain, use_orthogonal=self.args.use_orthogonal, use_relu=bool(getattr(self.args, "use_ReLU", True)))
        self.q_network.load_state_dict({k[len("agent."):]: v for k, v in init.items()})
        self._roll = None
        if train:
            self.exploration = LinearDecay(self.args.epsilon_start, self.args.epsilon_finish, self.args.epsilon_anneal_time)

    # -- rollout-time surface: one env step per call = ONE launch of k_policy_step (csrc/rollout.cu) ---------------------
    def _theta(self):
        """flat live parameter vector whose head is the agent block (re-bound to the trainer's vector once a QMix exists)"""
        return self.q_network.flat

    def _stepper(self):
        if self._roll is None:
            from offpolicy._b200.rollout import PolicyStepper
            self._roll = PolicyStepper(self.q_network_input_dim, self.act_dim, feature_norm=bool(getattr(self.args, "use_feature_normalization", True)),
                                       tanh=not getattr(self.args, "use_ReLU", True))
        return self._roll

    def _step(self, obs, rnn_states, available_actions=None, prev_actions=None):
        if self.prev_act_inp:                                                          # QMixPolicy.py:54-58
            obs = np.concatenate((obs, np.asarray(prev_actions, dtype=np.float32)), axis=-1)
        return self._stepper().step(self._theta(), obs, rnn_states, available_actions)

    def get_q_values(self, obs_batch, prev_action_batch, rnn_states, action_batch=None):
        """QMixPolicy.py:42-67.  (batch, dim) -> one step; (seq, batch, dim) -> the steps in turn, state resident on the device."""
        obs = np.asarray(obs_batch, dtype=np.float32)
        if obs.ndim == 3:
            qs, h = [], rnn_states
            pa = None if prev_action_batch is None else np.asarray(prev_action_batch, dtype=np.float32)
            for t in range(obs.shape[0]):
                q, h, _, _ = self._step(obs[t], h, prev_actions=None if pa is None else pa[t])
                qs.append(q)
            q = torch.from_numpy(np.stack(qs))
        else:
            q, h, _, _ = self._step(obs, rnn_states, prev_actions=prev_action_batch)
            q = torch.from_numpy(q)
        if action_batch is not None:
            q = self.q_values_from_actions(q, action_batch)
        return q, torch.from_numpy(h)

    def q_values_from_actions(self, q_batch, action_batch):
        a = torch.as_tensor(np.asarray(action_batch)).to(q_batch.device)
        return torch.gather(q_batch, q_batch.dim() - 1, a.max(dim=-1)[1].unsqueeze(-1))

    def get_actions(self, obs, prev_actions, rnn_states, available_actions=None, t_env=None, explore=False):
        """QMixPolicy.py:95-100: network step + masked arg-max on the device, epsilon-greedy mixing on the host with the
        reference's own generator calls (np.random.rand then Categorical.sample, QMixPolicy.py:157-165)."""
        obs = np.asarray(obs, dtype=np.float32)
        if obs.ndim != 2:
            q, h = self.get_q_values(obs, prev_actions, rnn_states)
            onehot_actions, greedy_Qs = self.actions_from_q(q, available_actions=available_actions, explore=explore, t_env=t_env)
            return onehot_actions, h, greedy_Qs
        _, h, greedy, greedy_q = self._step(obs, rnn_states, available_actions, prev_actions)
        greedy_Qs = torch.from_numpy(greedy_q)
        if explore:
            actions = self._eps_greedy(greedy, available_actions, t_env)
            return onehot(actions, self.act_dim), torch.from_numpy(h), greedy_Qs
        return onehot(greedy, self.act_dim), torch.from_numpy(h), greedy_Qs.unsqueeze(-1)

    def _eps_greedy(self, greedy, available_actions, t_env):
        batch = greedy.shape[0]
        eps = self.exploration.eval(t_env)
        rand = np.random.rand(batch)                                                  # QMixPolicy.py:160
        logits = torch.ones(batch, self.act_dim)
        if available_actions is not None:
            logits[torch.as_tensor(np.asarray(available_actions)) == 0] = -1e10       # avail_choose, util.py:297-302
        random_actions = torch.distributions.Categorical(logits=logits).sample().numpy()
        take = (rand < eps).astype(int)
        return (1 - take) * greedy + take * random_actions

    def actions_from_q(self, q_values, available_actions=None, explore=False, t_env=None):
        """QMixPolicy.py:102-174 for Q values that are already materialised (API compatibility; get_actions does not come here)."""
        q = torch.as_tensor(q_values).clone()
        if available_actions is not None:
            av = torch.as_tensor(np.asarray(available_actions)).to(q.device)
            q[av == 0] = -1e10
        greedy_Qs, greedy = q.max(dim=-1)
        if explore:
            assert q.dim() == 2, "Can only explore on non-sequences"
            return onehot(self._eps_greedy(greedy.cpu().numpy(), available_actions, t_env), self.act_dim), greedy_Qs
        return onehot(greedy.cpu().numpy(), self.act_dim), greedy_Qs.unsqueeze(-1)

    def get_random_actions(self, obs, available_actions=None):
        batch = obs.shape[0]
        logits = torch.ones(batch, self.act_dim)
        if available_actions is not None:
            logits[torch.as_tensor(np.asarray(available_actions)) == 0] = -1e10
        return torch.distributions.OneHotCategorical(logits=logits).sample().numpy()

    def init_hidden(self, num_agents, batch_size):
        if num_agents == -1:
            return torch.zeros(batch_size, self.hidden_size)
        return torch.zeros(num_agents, batch_size, self.hidden_size)

    def parameters(self):
        return self.q_network.parameters()

    def load_state(self, source_policy):
        self.q_network.load_state_dict(source_policy.q_network.state_dict())
