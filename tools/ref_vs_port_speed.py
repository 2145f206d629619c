"""Is the oracle port a fair stand-in for the reference in bench.py's CPU arm?  Times the UNMODIFIED reference learner (imported from
/root/reference: RecReplayBuffer.sample -> QMix.train_policy_on_batch -> soft_target_updates) and the oracle port on the same
QMIX 3m workload, same thread count, in this container.  (The GPU box has no reference checkout, so bench.py times the port there.)"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import numpy as np, torch
import bench


threads = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg, T, B = bench.make_cfg("qmix_3m")
E, steps, warm = 256, 12, 3

import ref_harness as rh
rh.import_reference()
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_goldens as mg
from offpolicy.utils.rec_buffer import RecReplayBuffer
torch.set_num_threads(threads)
args, pol, tr = mg.build_reference_qmix(cfg, ["--gain", "1"], T)
sp = rh.gym_spaces()
info = {"policy_0": dict(obs_space=[cfg.obs_dim], share_obs_space=[cfg.state_dim], act_space=sp.Discrete(cfg.act_dim))}
buf = RecReplayBuffer(info, {"policy_0": [0, 1, 2]}, E, T, True, True)
rs = np.random.default_rng(0)
d = lambda x: {"policy_0": x}
for c in range(0, E, 64):
    buf.insert(64, *[d(x) for x in bench.synth_episodes(cfg, T, 64, rs)])
def ref_run():
    times = []
    for s in range(warm + steps):
        t0 = time.perf_counter()
        smp = buf.sample(B)
        info_t, _, _ = tr.train_policy_on_batch(smp)
        tr.soft_target_updates()
        float(info_t["loss"])
        if s >= warm:
            times.append(time.perf_counter() - t0)
    return 1.0 / float(np.median(times))


# the container is shared and noisy: alternate the two arms three times and keep each arm's best median
ref_sps = port_sps = 0.0
for rep in range(3):
    port_sps = max(port_sps, bench.cpu_learner_steps_per_s(cfg, T, B, E, steps, warm, threads)[0])
    torch.set_num_threads(threads)
    ref_sps = max(ref_sps, ref_run())
print(json.dumps(dict(workload="qmix_3m", threads=threads, reference_steps_per_s=ref_sps, oracle_port_steps_per_s=port_sps, port_over_reference=port_sps / ref_sps)))
