"""Rollout-time policy step through `mx_policy_step` (csrc/rollout.cu): one launch per env step.

Used by the drop-in policies' `get_actions` / `get_q_values` (reference: QMixPolicy.py:42-67, 95-174;
rMADDPGPolicy.py:77-137).  Per call: ONE host->device copy of the packed inputs [obs | rnn_state | avail] from a pinned
staging buffer, ONE kernel, ONE device->host copy of the packed outputs [out | new rnn_state | greedy index | greedy value].
When the caller hands back the recurrent state it received from the previous call, the device copy is used directly
(`rnn_states` stay resident in HBM across the episode).
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
    def __init__(self, in_dim, out_dim):
        self.in_dim, self.out_dim = int(in_dim), int(out_dim)
        self.dev = capi.device()
        self.rows = 0
        self._last_h = None      # (host array handed out, device tensor it mirrors)

    def _ensure(self, rows):
        if rows <= self.rows:
            return
        self.rows = rows
        cuda = self.dev.type == "cuda"
        nin = rows * (self.in_dim + H + self.out_dim)
        nout = rows * (self.out_dim + H + 2)
        self.h_in_host = torch.zeros(nin, dtype=torch.float32, pin_memory=cuda)
        self.h_out_host = torch.zeros(nout, dtype=torch.float32, pin_memory=cuda)
        self.d_in = torch.zeros(nin, dtype=torch.float32, device=self.dev)
        self.d_out = torch.zeros(nout, dtype=torch.float32, device=self.dev)
        self._last_h = None      # the device-resident state lived in the old buffers

    def step(self, theta, obs, rnn_states, avail=None, want_greedy=True):
        """theta: flat device parameter vector; obs (R, in_dim); rnn_states (R, 64) array / tensor / None.
        Returns (out (R,out_dim), h_new (R,64), greedy_idx (R,) int64, greedy_val (R,)) as NumPy arrays (views of one pinned
        block, copied) -- one synchronisation per env step, which the CPU env loop needs anyway."""
        obs = np.asarray(obs, dtype=np.float32)
        R = obs.shape[0]
        self._ensure(R)
        I, A = self.in_dim, self.out_dim
        o_x, o_h, o_av = 0, R * I, R * (I + H)
        hin = self.h_in_host.numpy()
        hin[o_x:o_x + R * I] = obs.reshape(-1)
        resident = False
        if rnn_states is None:
            hin[o_h:o_h + R * H] = 0.0
        elif self._last_h is not None and self._last_h[1] == R and _host_ptr(rnn_states) == _host_ptr(self._last_h[0]):
            # the very buffer we handed out last step (the runner passes it straight back, smac_runner.py:86-98; holding a
            # reference keeps its memory from being reused): its device copy is still in the output block
            resident = True
        else:
            hs = rnn_states.detach().cpu().numpy() if isinstance(rnn_states, torch.Tensor) else np.asarray(rnn_states)
            hin[o_h:o_h + R * H] = hs.astype(np.float32, copy=False).reshape(-1)
        if avail is not None:
            av = avail.detach().cpu().numpy() if isinstance(avail, torch.Tensor) else np.asarray(avail)
            hin[o_av:o_av + R * A] = av.astype(np.float32, copy=False).reshape(-1)
        n_in = R * (I + H + A)
        self.d_in[:n_in].copy_(self.h_in_host[:n_in], non_blocking=True)
        a = capi.PolicyStepArgs()
        a.theta = theta.data_ptr()
        a.in_dim, a.out_dim, a.rows, a.x_ld, a.avail_ld = I, A, R, I, A
        base = self.d_in.data_ptr()
        a.x = base + 4 * o_x
        ob = self.d_out.data_ptr()
        # the new state is written into the packed output block; a resident state is read from that same place (the kernel
        # reads a row's state completely before it writes it)
        a.h_in = ob + 4 * R * A if resident else base + 4 * o_h
        a.h_out = ob + 4 * R * A
        a.out = ob
        a.avail = base + 4 * o_av if avail is not None else None
        a.greedy = ob + 4 * R * (A + H) if want_greedy else None
        a.greedy_q = ob + 4 * R * (A + H + 1) if want_greedy else None
        capi.check(capi.lib().mx_policy_step(C.byref(a), capi.stream_ptr()))
        n_out = R * (A + H + 2)
        self.h_out_host[:n_out].copy_(self.d_out[:n_out], non_blocking=True)
        if self.dev.type == "cuda":
            torch.cuda.current_stream().synchronize()
        res = self.h_out_host.numpy()
        out = res[:R * A].reshape(R, A).copy()
        h_new = res[R * A:R * (A + H)].reshape(R, H).copy()
        gi = res[R * (A + H):R * (A + H) + R].view(np.int32).astype(np.int64) if want_greedy else None
        gq = res[R * (A + H + 1):R * (A + H + 1) + R].copy() if want_greedy else None
        self._last_h = (h_new, R)
        return out, h_new, gi, gq
