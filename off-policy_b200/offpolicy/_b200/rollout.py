"""Rollout-time policy step through `mx_policy_step` (csrc/rollout.cu): one launch per env step.

Used by the drop-in policies' `get_actions` / `get_q_values` (reference: QMixPolicy.py:42-67, 95-174;
rMADDPGPolicy.py:77-137).  Inputs and outputs live in ONE block of mapped pinned host memory that the kernel addresses
directly (UVA): per call the host writes [obs | avail | (rnn_state)] into the block, launches the kernel and synchronises the
stream -- no memcpy calls; the kernel reads the observation over PCIe and writes [out | new rnn_state | greedy index | greedy
value] back the same way.  The recurrent state is ALSO kept in device memory: when the caller hands back the state it received
from the previous call (the runner does, smac_runner.py:86-98) the device copy is used and nothing is uploaded.
"""
import ctypes as C

import numpy as np
import torch

from offpolicy._b200 import capi

H = 64


def _host_ptr(x):
    if isinstance(x, torch.Tensor):
        return x.data_ptr() if x.device.type == "cpu" and x.dtype == torch.float32 and x.is_contiguous() else None
    if isinstance(x, np.ndarray) and x.dtype == np.float32 and x.flags["C_CONTIGUOUS"]:
        return x.__array_interface__["data"][0]
    return None


class PolicyStepper(object):
    def __init__(self, in_dim, out_dim, mlp=False, feature_norm=True, tanh=False):
        self.in_dim, self.out_dim = int(in_dim), int(out_dim)
        self.mlp = bool(mlp)      # non-recurrent net (M_QMixPolicy): no state is carried
        self.feature_norm = bool(feature_norm)      # False: --use_feature_normalization switched off
        self.tanh = bool(tanh)                      # True: --use_ReLU switched off
        self.dev = capi.device()
        self.rows = 0
        self._last_h = None      # (host array handed out, rows): its device copy is in self.d_h

    def _ensure(self, rows):
        if rows <= self.rows:
            return
        self.rows = rows
        cuda = self.dev.type == "cuda"
        I, A = self.in_dim, self.out_dim
        n = rows * (I + A + H) + rows * (A + H + 2)
        # pinned (page-locked) host memory is mapped into the device address space by the CUDA allocator torch uses
        self.block = torch.zeros(n, dtype=torch.float32, pin_memory=cuda)
        self.block_np = self.block.numpy()
        self.d_h = torch.zeros(rows * H, dtype=torch.float32, device=self.dev)     # recurrent state resident on the device
        self._last_h = None
        self._args = capi.PolicyStepArgs()

    def step(self, theta, obs, rnn_states, avail=None, want_greedy=True):
        """theta: flat device parameter vector; obs (R, in_dim); rnn_states (R, 64) array / tensor / None.
        Returns (out (R,out_dim), h_new (R,64), greedy_idx (R,) int64, greedy_val (R,)) as NumPy arrays -- one stream
        synchronisation per env step, which the CPU env loop needs anyway."""
        obs = np.asarray(obs, dtype=np.float32)
        R = obs.shape[0]
        self._ensure(R)
        I, A = self.in_dim, self.out_dim
        blk = self.block_np
        o_x, o_av, o_h = 0, R * I, R * (I + A)
        o_out = R * (I + A + H)
        o_hn, o_gi, o_gq = o_out + R * A, o_out + R * (A + H), o_out + R * (A + H + 1)
        blk[o_x:o_x + R * I] = obs.reshape(-1)
        resident = False
        if rnn_states is None:
            blk[o_h:o_h + R * H] = 0.0
        elif self._last_h is not None and self._last_h[1] == R and _host_ptr(rnn_states) == _host_ptr(self._last_h[0]) and \
                np.array_equal(self._last_h[0], self._last_h[2]):
            # the very buffer we handed out last step (holding a reference keeps its memory from being reused) AND nobody has edited it
            # in place since (a caller that zeroes the rows of finished envs must see its edit honoured): the device copy is current
            resident = True
        else:
            hs = rnn_states.detach().cpu().numpy() if isinstance(rnn_states, torch.Tensor) else np.asarray(rnn_states)
            blk[o_h:o_h + R * H] = hs.astype(np.float32, copy=False).reshape(-1)
        if avail is not None:
            av = avail.detach().cpu().numpy() if isinstance(avail, torch.Tensor) else np.asarray(avail)
            blk[o_av:o_av + R * A] = av.astype(np.float32, copy=False).reshape(-1)
        a = self._args
        base = self.block.data_ptr()
        a.theta = theta.data_ptr()
        a.in_dim, a.out_dim, a.rows, a.x_ld, a.avail_ld = I, A, R, I, A
        a.x = base + 4 * o_x
        a.mlp = int(self.mlp)
        a.no_feature_norm = 0 if self.feature_norm else 1
        a.use_tanh = 1 if self.tanh else 0
        a.h_in = None if self.mlp else (self.d_h.data_ptr() if resident else base + 4 * o_h)
        a.h_out = None if self.mlp else self.d_h.data_ptr()
        a.h_copy = None if self.mlp else base + 4 * o_hn
        a.out = base + 4 * o_out
        a.avail = base + 4 * o_av if avail is not None else None
        a.greedy = base + 4 * o_gi if want_greedy else None
        a.greedy_q = base + 4 * o_gq if want_greedy else None
        capi.check(capi.lib().mx_policy_step(C.byref(a), capi.stream_ptr()))
        if self.dev.type == "cuda":
            torch.cuda.current_stream().synchronize()
        out = blk[o_out:o_out + R * A].reshape(R, A).copy()
        h_new = blk[o_hn:o_hn + R * H].reshape(R, H).copy()
        gi = blk[o_gi:o_gi + R].view(np.int32).astype(np.int64) if want_greedy else None
        gq = blk[o_gq:o_gq + R].copy() if want_greedy else None
        self._last_h = (h_new, R, h_new.copy())
        return out, h_new, gi, gq
