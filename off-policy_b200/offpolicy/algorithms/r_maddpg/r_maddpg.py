"""Drop-in `R_MADDPG` trainer (reference: offpolicy/algorithms/r_maddpg/r_maddpg.py) on the sm_100a learner.

`shared_train_policy_on_batch(p_id, batch)` = one `mx_maddpg_step`: target-actor next actions, critic sequence + target
branch steps, TD target, critic loss/backward/Adam, then (every `actor_update_interval`-th call) the actor update through
the updated critic -- all on the device.  MATD3's Gaussian target-action noise is drawn on the host with the reference's
own call (`torch.empty(shape).normal_`, utils/util.py:217-218) so a seeded run consumes torch's CPU RNG identically; for
Discrete actors the Gumbel draws of the target actions (MATD3) and of the actor update (`use_gumbel=True`, r_maddpg.py:277) are
drawn the same way (utils/util.py:127-130), in the reference's order.
`cent_train_policy_on_batch` (per-agent centralised observations) is unusable in the reference and is not built."""
import ctypes as C

import numpy as np
import torch

from offpolicy._b200 import capi
from offpolicy.algorithms.r_maddpg.algorithm.rMADDPGPolicy import maddpg_cfg_struct, sample_gumbel
from offpolicy.utils.rec_buffer import SampledBatch, DeviceArray


class _HostBatchC(object):
    """Device copy of a reference-layout NumPy batch."""

    def __init__(self, cfg, dev):
        B, T, N = cfg.max_batch, cfg.episode_len, cfg.n_agents
        r4 = lambda v: (v + 3) // 4 * 4
        self.cfg = cfg
        self.obs_ld, self.share_ld, self.act_ld = r4(cfg.obs_dim), r4(cfg.state_dim), r4(cfg.act_dim)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.obs, self.share, self.acts = z(B, T + 1, N, self.obs_ld), z(B, T + 1, self.share_ld), z(B, T, N, self.act_ld)
        self.rew, self.dones, self.dones_env, self.weights = z(B, T, N), z(B, T, N), z(B, T), z(B)
        self.avail = None
        self.dev = dev

    def pack(self, batch, p_id, use_per):
        obs, share, acts, rew, dones, dones_env, _avail, weights, _idx = batch
        c = self.cfg
        t = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float32).to(self.dev)
        o = t(obs[p_id])
        B = o.shape[2]
        self.obs[:B, :, :, :c.obs_dim] = o.permute(2, 1, 0, 3)
        self.share[:B, :, :c.state_dim] = t(share[p_id]).permute(1, 0, 2)
        self.acts[:B, :, :, :c.act_dim] = t(acts[p_id]).permute(2, 1, 0, 3)
        self.rew[:B] = t(rew[p_id])[..., 0].permute(2, 1, 0)
        self.dones[:B] = t(dones[p_id])[..., 0].permute(2, 1, 0)
        self.dones_env[:B] = t(dones_env[p_id])[..., 0].permute(1, 0)
        if use_per:
            self.weights[:B] = t(weights)
        has_avail = _avail is not None and _avail.get(p_id) is not None
        if has_avail:
            if self.avail is None:
                self.avail = torch.ones(c.max_batch, c.episode_len + 1, c.n_agents, self.act_ld, dtype=torch.float32, device=self.dev)
            self.avail[:B, :, :, :c.act_dim] = t(_avail[p_id]).permute(2, 1, 0, 3)
        b = capi.Batch()
        b.B, b.obs_ld, b.share_ld, b.act_ld = B, self.obs_ld, self.share_ld, self.act_ld
        b.obs, b.share, b.acts = self.obs.data_ptr(), self.share.data_ptr(), self.acts.data_ptr()
        b.rewards, b.dones, b.dones_env = self.rew.data_ptr(), self.dones.data_ptr(), self.dones_env.data_ptr()
        b.weights = self.weights.data_ptr() if use_per else None
        b.avail = self.avail.data_ptr() if has_avail else None
        return b


class R_MADDPG(object):
    def __init__(self, args, num_agents, policies, policy_mapping_fn, device=None, episode_length=None, actor_update_interval=1):
        self.args = args
        self.use_per = args.use_per
        self.num_agents = num_agents
        self.policies = policies
        self.policy_mapping_fn = policy_mapping_fn
        self.policy_ids = sorted(list(self.policies.keys()))
        if self.policy_ids != ["policy_0"]:
            raise NotImplementedError("B200 R-MADDPG path: only the shared-policy configuration ('policy_0') is implemented")
        self.policy_agents = {p: sorted(a for a in range(num_agents) if policy_mapping_fn(a) == p) for p in self.policies}
        self.episode_length = args.episode_length if episode_length is None else episode_length
        self.actor_update_interval = actor_update_interval
        self.num_updates = {p: 0 for p in self.policy_ids}
        self.use_same_share_obs = getattr(args, "use_same_share_obs", True)
        pol = self.policies["policy_0"]
        self.max_batch = int(getattr(args, "batch_size", 32))
        lib = capi.lib()
        self.dev = capi.device()
        self.cfg = maddpg_cfg_struct(args, num_agents, pol.obs_dim, pol.act_dim, pol.central_obs_dim, self.episode_length, self.max_batch,
                                     pol.td3, pol.target_noise if pol.td3 else 0.0, actor_update_interval, pol.discrete)
        nbytes = int(lib.mx_maddpg_workspace_bytes(C.byref(self.cfg)))
        if nbytes < 0:
            raise capi.MxError(lib.mx_last_error().decode())
        self.workspace = torch.zeros(nbytes, dtype=torch.uint8, device=self.dev)
        av = (C.c_void_p * 4)(*[v.data_ptr() for v in pol.actor_vecs])
        cv = (C.c_void_p * 4)(*[v.data_ptr() for v in pol.critic_vecs])
        h = C.c_void_p()
        capi.check(lib.mx_maddpg_create(C.byref(self.cfg), av, cv, capi.ptr(self.workspace), nbytes, C.byref(h)))
        self.handle = h
        pol._trainer = self
        ip = lib.mx_maddpg_info(self.handle) - self.workspace.data_ptr()
        self._info = self.workspace[ip:ip + 32].view(torch.float32)
        pp = lib.mx_maddpg_priorities(self.handle) - self.workspace.data_ptr()
        self._prio = self.workspace[pp:pp + 4 * self.max_batch].view(torch.float32)
        self._host_batch = None
        self._noise_dev = None
        self._actor_noise_dev = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                capi.lib().mx_maddpg_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def grad_views(self):
        """Numerator gradients (actor, critic) as flat views, for the parity tests."""
        a, c = C.c_int64(), C.c_int64()
        capi.lib().mx_maddpg_grad_views(self.handle, C.byref(a), C.byref(c))
        pol = self.policies["policy_0"]
        return (self.workspace[a.value:a.value + 4 * (pol.Pa + 4)].view(torch.float32),
                self.workspace[c.value:c.value + 4 * (pol.Pc + 4)].view(torch.float32))

    def _device_batch(self, batch):
        if isinstance(batch, SampledBatch):
            buf = batch.buffers["policy_0"]
            if buf.sample_serial != batch.serial["policy_0"]:
                raise RuntimeError("stale sample: the buffer has been sampled again since this batch was drawn")
            return buf.batch_struct(batch.B)
        if self._host_batch is None:
            self._host_batch = _HostBatchC(self.cfg, self.dev)
        return self._host_batch.pack(batch, "policy_0", self.use_per)

    def draw_target_noise(self, B):
        """The draw the reference makes for the target actions of one update: (T+1, N*B, Ac), agent-major rows, CPU RNG."""
        pol = self.policies["policy_0"]
        T, N, Ac = self.episode_length, self.num_agents, pol.act_dim
        if pol.discrete:
            return sample_gumbel((T + 1, N * B, Ac))                                               # util.py:137 via rMADDPGPolicy.py:105-106
        return torch.empty(T + 1, N * B, Ac).normal_(mean=0, std=float(pol.target_noise))          # util.py:217-218

    def draw_actor_noise(self, B):
        """Gumbel draws of the actor update's `get_actions(..., use_gumbel=True)` over obs[:-1] (r_maddpg.py:277): (T, N*B, Ac)."""
        pol = self.policies["policy_0"]
        return sample_gumbel((self.episode_length, self.num_agents * B, pol.act_dim))

    def _target_noise(self, B):
        """N(0, target_noise) / Gumbel draws for every target action, in batch row order on the device."""
        pol = self.policies["policy_0"]
        if not pol.td3:
            return None
        T, N, Ac = self.episode_length, self.num_agents, pol.act_dim
        noise = self.draw_target_noise(B)
        ours = noise.view(T + 1, N, B, Ac).permute(2, 0, 1, 3).contiguous()                      # -> [b][t][n][Ac]
        self._noise_dev = ours.to(self.dev, non_blocking=True)
        return self._noise_dev

    def _actor_noise(self, B):
        """Gumbel draws of the actor update's `get_actions(..., use_gumbel=True)` over obs[:-1] (r_maddpg.py:277), padded to T+1 steps."""
        pol = self.policies["policy_0"]
        T, N, Ac = self.episode_length, self.num_agents, pol.act_dim
        g = self.draw_actor_noise(B)
        ours = torch.zeros(B, T + 1, N, Ac)
        ours[:, :T] = g.view(T, N, B, Ac).permute(2, 0, 1, 3)
        self._actor_noise_dev = ours.to(self.dev, non_blocking=True)
        return self._actor_noise_dev

    def train_policy_on_batch(self, update_policy_id, batch):
        if self.use_same_share_obs:
            return self.shared_train_policy_on_batch(update_policy_id, batch)
        return self.cent_train_policy_on_batch(update_policy_id, batch)

    def cent_train_policy_on_batch(self, update_policy_id, batch):
        raise NotImplementedError("cent_train_policy_on_batch is unusable in the reference (missing train_info['update_actor']) and is not built")

    def shared_train_policy_on_batch(self, update_policy_id, batch):
        lib = capi.lib()
        b = self._device_batch(batch)
        noise = self._target_noise(b.B)
        pol = self.policies["policy_0"]
        will_update_actor = self.num_updates[update_policy_id] % self.actor_update_interval == 0
        actor_noise = self._actor_noise(b.B) if (pol.discrete and will_update_actor) else None
        upd = C.c_int32()
        capi.check(lib.mx_maddpg_step_ex(self.handle, C.byref(b), capi.ptr(noise), capi.ptr(actor_noise), C.byref(upd), capi.stream_ptr()))
        info = self._info
        train_info = {"critic_loss": info[0], "critic_grad_norm": info[1]}
        if upd.value:
            train_info["actor_loss"], train_info["actor_grad_norm"] = info[4], info[5]
        train_info["update_actor"] = bool(upd.value)
        self.num_updates[update_policy_id] += 1
        new_priorities = DeviceArray(self._prio[:b.B]) if self.use_per else None
        return train_info, new_priorities, batch[8]

    def prep_training(self):
        pass

    def prep_rollout(self):
        pass
