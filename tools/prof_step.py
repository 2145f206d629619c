"""Minimal driver for ncu: a few eager QMIX learner steps at a BASELINE workload (no CPU baseline, no graphs)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "off-policy_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import bench
import qmix_checks as qc
import replay_checks as rc

workload = sys.argv[1] if len(sys.argv) > 1 else "qmix_3m"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg, T, B = bench.make_cfg(workload)
E = 256
rs = np.random.default_rng(0)
buf = rc.make_buffers(cfg.n_agents, cfg.obs_dim, cfg.act_dim, cfg.state_dim, T, E, per_alpha=0.6 if cfg.use_per else None, rng="device",
                      max_batch=max(B, 128))
for c in range(0, E, 128):
    buf.insert(128, *[rc.d(x) for x in bench.synth_episodes(cfg, T, 128, rs)])
args, pol, tr = qc.build_trainer(cfg, B, T)
from offpolicy._b200 import capi
capi.lib().mx_qmix_set_debug(tr.handle, 0)
buf.seed_device_rng(1)
for s in range(steps):
    smp = buf.sample(B, 0.4, "policy_0") if cfg.use_per else buf.sample(B)
    info, prio, idx = tr.train_policy_on_batch(smp)
    if cfg.use_per:
        buf.update_priorities(idx, prio, "policy_0")
    tr.soft_target_updates()
torch.cuda.synchronize()
print("done", float(info["loss"]))
