"""Drop-in `M_QMix` trainer (reference: offpolicy/algorithms/mqmix/mqmix.py): transition-level QMIX with MLP agent networks on the
fused sm_100a learner in its `mlp` mode (csrc/qmix.cu): a batch of B transitions = B episodes of length 1, the Q head is the first
act_dim rows of the weight_ih slot, there is no recurrence kernel; mixer, TD target, masked MSE/Huber (no steps are masked when
T = 1, so the loss is the reference's plain mean, mqmix.py:188-205), PER priorities |error| + eps, clip, Adam and the target updates
are the recurrent path's kernels.
"""
import ctypes as C

import numpy as np
import torch

from offpolicy._b200 import capi
from offpolicy.algorithms.qmix.algorithm.QMixPolicy import qmix_cfg_struct
from offpolicy.algorithms.qmix.qmix import QMix
from offpolicy.utils.mlp_buffer import MlpSampledBatch
from offpolicy.utils.rec_buffer import DeviceArray


class _HostTransitions(object):
    """Device copy of a batch handed over in the reference's NumPy layout (mlp_buffer.py:203-240): compatibility path."""

    def __init__(self, cfg, dev):
        self.cfg, self.dev = cfg, dev
        B, N = cfg.max_batch, cfg.n_agents
        r4 = lambda v: (v + 3) // 4 * 4
        self.obs_ld, self.share_ld, self.act_ld = r4(cfg.obs_dim), r4(cfg.state_dim), r4(cfg.act_dim)
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
        self.obs, self.share = z(B, 2, N, self.obs_ld), z(B, 2, self.share_ld)
        self.acts, self.act_idx, self.avail = z(B, 1, N, self.act_ld), z(B, 1, N, dt=torch.int32), z(B, 2, N, self.act_ld)
        self.rew, self.dones, self.dones_env, self.weights = z(B, 1, N), z(B, 1, N), z(B, 1), z(B)

    def pack(self, batch, p_id, use_per):
        obs, share, acts, rew, nobs, nshare, dones, dones_env, _valid, avail, navail = batch[:11]
        weights = batch[11] if len(batch) > 11 else None
        c, dev = self.cfg, self.dev
        t = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float32).to(dev)
        o = t(obs[p_id])                                          # (N, B, O)
        B = o.shape[1]
        self.obs[:B, 0, :, :c.obs_dim] = o.permute(1, 0, 2)
        self.obs[:B, 1, :, :c.obs_dim] = t(nobs[p_id]).permute(1, 0, 2)
        self.share[:B, 0, :c.state_dim] = t(share[p_id])
        self.share[:B, 1, :c.state_dim] = t(nshare[p_id])
        a = t(acts[p_id]).permute(1, 0, 2)
        self.acts[:B, 0, :, :c.act_dim] = a
        self.act_idx[:B, 0] = a.max(dim=-1)[1].to(torch.int32)
        have_avail = navail is not None and navail[p_id] is not None
        if have_avail:
            self.avail[:B, 1, :, :c.act_dim] = t(navail[p_id]).permute(1, 0, 2)
            if avail is not None and avail[p_id] is not None:
                self.avail[:B, 0, :, :c.act_dim] = t(avail[p_id]).permute(1, 0, 2)
        self.rew[:B, 0] = t(rew[p_id])[..., 0].permute(1, 0)
        self.dones[:B, 0] = t(dones[p_id])[..., 0].permute(1, 0)
        self.dones_env[:B, 0] = t(dones_env[p_id]).reshape(B)
        if use_per:
            self.weights[:B] = t(weights)
        b = capi.Batch()
        b.B = B
        b.obs_ld, b.share_ld, b.act_ld = self.obs_ld, self.share_ld, self.act_ld
        b.obs, b.share, b.acts, b.act_idx = self.obs.data_ptr(), self.share.data_ptr(), self.acts.data_ptr(), self.act_idx.data_ptr()
        b.avail = self.avail.data_ptr() if have_avail else None
        b.rewards, b.dones, b.dones_env = self.rew.data_ptr(), self.dones.data_ptr(), self.dones_env.data_ptr()
        b.weights = self.weights.data_ptr() if use_per else None
        b.idx = None
        return b


class M_QMix(QMix):
    def __init__(self, args, num_agents, policies, policy_mapping_fn, device=None, vdn=False):
        self._mlp = True
        QMix.__init__(self, args, num_agents, policies, policy_mapping_fn, device=device, episode_length=1, vdn=vdn)

    # hooks of QMix.__init__ ------------------------------------------------------------------------------------
    def _cfg_struct(self, args, num_agents, pol):
        return qmix_cfg_struct(args, num_agents, pol.obs_dim, pol.act_dim, pol.central_obs_dim, 1, self.max_batch, vdn=self.vdn,
                               use_avail=True, world_size=self.world_size, mlp=True)

    def train_policy_on_batch(self, batch, use_same_share_obs=True):
        """mqmix.py:67-216.  `batch` = what MlpReplayBuffer.sample / PrioritizedMlpReplayBuffer.sample returned."""
        lib, stream = capi.lib(), capi.stream_ptr()
        if isinstance(batch, MlpSampledBatch):
            rep = batch.buffers["policy_0"].rep
            if rep.sample_serial != batch.serial["policy_0"]:
                raise RuntimeError("stale sample: the buffer has been sampled again since this batch was drawn")
            b = rep.batch_struct(batch.B)
        else:
            if self._host_batch is None:
                self._host_batch = _HostTransitions(self.cfg, self.dev)
            b = self._host_batch.pack(batch, "policy_0", self.use_per)
        if self.world_size > 1 and not self._p2p:
            capi.check(lib.mx_qmix_backward_only(self.handle, C.byref(b), stream))
            torch.distributed.all_reduce(self._grad_buf)
            capi.check(lib.mx_qmix_apply(self.handle, stream))
        else:
            capi.check(lib.mx_qmix_step(self.handle, C.byref(b), stream))
        self._check_exchange()
        v = self._info_views
        train_info = {"loss": v[0], "grad_norm": v[1], "Q_tot": v[2]}
        new_priorities = DeviceArray(self._prio_view[:b.B]) if self.use_per else None
        return train_info, new_priorities, (batch[12] if len(batch) > 12 else None)
