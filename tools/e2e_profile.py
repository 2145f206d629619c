"""Where the end-to-end step (host inputs through the drop-in API) spends its time: per-phase host enqueue cost (no sync)
and per-phase cost with a device sync after each phase.  Usage: python tools/e2e_profile.py [workload] [n]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "off-policy_b200"))
import numpy as np
import torch
import bench
from offpolicy._b200 import capi
from offpolicy._b200 import factory

w = sys.argv[1] if len(sys.argv) > 1 else "qmix_3m"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
torch.cuda.set_device(0)
torch.set_num_threads(1)
lib = capi.lib()
cfg, T, B = bench.make_cfg(w)
N, O, A, S = cfg.n_agents, cfg.obs_dim, cfg.act_dim, cfg.state_dim
E = 1024
rs = np.random.default_rng(0)
buf = factory.make_rec_buffers(N, O, A, S, T, E, per_alpha=0.6 if cfg.use_per else None, rng="numpy", max_batch=max(B, 128))
for c in range(0, E, 128):
    buf.insert(128, *[factory.pd(x) for x in bench.synth_episodes(cfg, T, 128, rs)])
torch.manual_seed(1); np.random.seed(1)
args_ns, pol, tr = factory.build_qmix(cfg, B, T, debug=False)
fresh = [bench.synth_episodes(cfg, T, 1, rs) for _ in range(8)]
phases = ["insert", "sample", "train", "prio", "soft", "loss"]


def run(sync):
    acc = dict.fromkeys(phases, 0.0)
    t_all = time.perf_counter()
    for i in range(n):
        t = time.perf_counter()
        buf.insert(1, *[factory.pd(x) for x in fresh[i % 8]])
        if sync: torch.cuda.synchronize()
        t2 = time.perf_counter(); acc["insert"] += t2 - t; t = t2
        smp = buf.sample(B, 0.4, "policy_0") if cfg.use_per else buf.sample(B)
        if sync: torch.cuda.synchronize()
        t2 = time.perf_counter(); acc["sample"] += t2 - t; t = t2
        info, prio, idx = tr.train_policy_on_batch(smp)
        if sync: torch.cuda.synchronize()
        t2 = time.perf_counter(); acc["train"] += t2 - t; t = t2
        if cfg.use_per:
            buf.update_priorities(idx, prio, "policy_0")
            if sync: torch.cuda.synchronize()
        t2 = time.perf_counter(); acc["prio"] += t2 - t; t = t2
        tr.soft_target_updates()
        if sync: torch.cuda.synchronize()
        t2 = time.perf_counter(); acc["soft"] += t2 - t; t = t2
        float(info["loss"])
        t2 = time.perf_counter(); acc["loss"] += t2 - t; t = t2
    tot = time.perf_counter() - t_all
    return {k: 1e6 * v / n for k, v in acc.items()}, 1e6 * tot / n


for _ in range(2):
    run(False)
for sync in (False, True):
    acc, tot = run(sync)
    print("sync_after_each_phase=%s  total %.1f us/step :: " % (sync, tot) + "  ".join("%s %.1f" % (k, acc[k]) for k in phases))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); run(False); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
