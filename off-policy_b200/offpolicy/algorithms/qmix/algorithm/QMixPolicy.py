"""Drop-in `QMixPolicy` (reference: offpolicy/algorithms/qmix/algorithm/QMixPolicy.py).

The agent Q-network's parameters are views into the flat device vector the CUDA learner trains in place
(`q_network.state_dict()` keeps the reference key names).  Update-time methods
(get_q_values over sequences, q_values_from_actions, greedy actions_from_q) are executed inside the fused
learner kernels (csrc/agent_fwd.cu); what remains here is the rollout-time surface the runner calls once per
env step -- single-step forward + epsilon-greedy (QMixPolicy.py:95-191) -- done with a handful of torch ops
on the same parameter views (SURVEY.md section 8(f).1 lists a dedicated rollout kernel as the next widening).
"""
import ctypes as C

import numpy as np
import torch
import torch.nn.functional as F

from offpolicy._b200 import capi
from offpolicy._b200.flat import FlatModule, reference_style_init
from offpolicy._b200.host_util import LinearDecay, space_dim, is_discrete, onehot


def qmix_cfg_struct(args, n_agents, obs_dim, act_dim, state_dim, episode_len, max_batch, vdn=False, use_avail=True, world_size=1):
    return capi.QmixCfg(
        n_agents=n_agents, obs_dim=obs_dim, act_dim=act_dim, state_dim=state_dim, hidden=args.hidden_size,
        mixer_hidden=args.mixer_hidden_dim, hyper_hidden=args.hypernet_hidden_dim, hyper_layers=args.hypernet_layers,
        episode_len=episode_len, max_batch=max_batch, vdn=int(vdn), double_q=int(args.use_double_q),
        use_huber=int(args.use_huber_loss), use_per=int(args.use_per), use_avail=int(use_avail), world_size=world_size,
        gamma=args.gamma, huber_delta=args.huber_delta, per_nu=args.per_nu, per_eps=args.per_eps, lr=args.lr,
        adam_beta1=0.9, adam_beta2=0.999, adam_eps=args.opti_eps, max_grad_norm=args.max_grad_norm, tau=args.tau)


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
        if getattr(self.args, "prev_act_inp", False):
            raise NotImplementedError("B200 QMIX path: --prev_act_inp is not implemented")
        for flag, want in (("use_rnn_layer", True), ("use_feature_normalization", True), ("use_ReLU", True), ("use_conv1d", False)):
            if getattr(self.args, flag, want) != want:
                raise NotImplementedError("B200 QMIX path requires %s=%s" % (flag, want))
        if getattr(self.args, "layer_N", 1) != 1 or getattr(self.args, "recurrent_N", 1) != 1:
            raise NotImplementedError("B200 QMIX path requires layer_N=1, recurrent_N=1")
        self.q_network_input_dim = self.obs_dim

        capi.lib()
        self.dev = capi.device()
        # agent-only layout (vdn=1 -> no mixer block); the trainer re-binds these views into its full vector
        cfg = qmix_cfg_struct(self.args, 1, self.obs_dim, self.act_dim, 1, 1, 1, vdn=True)
        entries, total = param_entries(cfg)
        self._entries = entries
        flat = torch.zeros(total, dtype=torch.float32, device=self.dev)
        self.q_network = FlatModule(flat, entries, "agent.")
        init = reference_style_init(entries, dict(hidden=self.hidden_size, obs_dim=self.obs_dim, act_dim=self.act_dim),
                                    gain=self.args.gain, use_orthogonal=self.args.use_orthogonal)
        self.q_network.load_state_dict({k[len("agent."):]: v for k, v in init.items()})
        if train:
            self.exploration = LinearDecay(self.args.epsilon_start, self.args.epsilon_finish, self.args.epsilon_anneal_time)

    # -- rollout-time forward (one env step, batch = agents) -------------------------------------------
    def _forward_step(self, obs, h):
        p = self.q_network.views
        H = self.hidden_size
        x = F.layer_norm(obs, (self.obs_dim,), p["rnn.feature_norm.weight"], p["rnn.feature_norm.bias"])
        x = F.layer_norm(F.relu(F.linear(x, p["rnn.mlp.fc1.0.weight"], p["rnn.mlp.fc1.0.bias"])), (H,),
                         p["rnn.mlp.fc1.2.weight"], p["rnn.mlp.fc1.2.bias"])
        x = F.layer_norm(F.relu(F.linear(x, p["rnn.mlp.fc2.0.0.weight"], p["rnn.mlp.fc2.0.0.bias"])), (H,),
                         p["rnn.mlp.fc2.0.2.weight"], p["rnn.mlp.fc2.0.2.bias"])
        gi = F.linear(x, p["rnn.rnn.rnn.weight_ih_l0"], p["rnn.rnn.rnn.bias_ih_l0"])
        gh = F.linear(h, p["rnn.rnn.rnn.weight_hh_l0"], p["rnn.rnn.rnn.bias_hh_l0"])
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h2 = (1 - z) * n + z * h
        y = F.layer_norm(h2, (H,), p["rnn.rnn.norm.weight"], p["rnn.rnn.norm.bias"])
        return F.linear(y, p["q.action_out.weight"], p["q.action_out.bias"]), h2

    def get_q_values(self, obs_batch, prev_action_batch, rnn_states, action_batch=None):
        obs = torch.as_tensor(np.asarray(obs_batch), dtype=torch.float32).to(self.dev)
        h = torch.as_tensor(rnn_states, dtype=torch.float32).to(self.dev)
        with torch.no_grad():
            if obs.dim() == 3:          # (seq, batch, dim): step through the sequence
                qs = []
                for t in range(obs.shape[0]):
                    q, h = self._forward_step(obs[t], h)
                    qs.append(q)
                q = torch.stack(qs)
            else:
                q, h = self._forward_step(obs, h)
        if action_batch is not None:
            q = self.q_values_from_actions(q, action_batch)
        return q, h

    def q_values_from_actions(self, q_batch, action_batch):
        a = torch.as_tensor(np.asarray(action_batch)).to(q_batch.device)
        return torch.gather(q_batch, q_batch.dim() - 1, a.max(dim=-1)[1].unsqueeze(-1))

    def get_actions(self, obs, prev_actions, rnn_states, available_actions=None, t_env=None, explore=False):
        q, h = self.get_q_values(obs, prev_actions, rnn_states)
        onehot_actions, greedy_Qs = self.actions_from_q(q, available_actions=available_actions, explore=explore, t_env=t_env)
        return onehot_actions, h, greedy_Qs

    def actions_from_q(self, q_values, available_actions=None, explore=False, t_env=None):
        q = q_values.clone()
        if available_actions is not None:
            av = torch.as_tensor(np.asarray(available_actions)).to(q.device)
            q[av == 0] = -1e10                                                         # util.py:297-302
        greedy_Qs, greedy = q.max(dim=-1)
        if explore:
            assert q.dim() == 2, "Can only explore on non-sequences"
            batch = q.shape[0]
            eps = self.exploration.eval(t_env)
            rand = np.random.rand(batch)                                              # QMixPolicy.py:160
            logits = torch.ones(batch, self.act_dim)
            if available_actions is not None:
                logits[torch.as_tensor(np.asarray(available_actions)) == 0] = -1e10
            random_actions = torch.distributions.Categorical(logits=logits).sample().numpy()
            take = (rand < eps).astype(int)
            actions = (1 - take) * greedy.cpu().numpy() + take * random_actions
            return onehot(actions, self.act_dim), greedy_Qs
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
