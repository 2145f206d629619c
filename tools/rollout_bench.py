"""Rollout-time policy call: us per env step of QMixPolicy.get_actions (3m shapes: 3 agents, obs 30, 9 actions) through
k_policy_step vs the same network evaluated with ~25 eager torch CUDA ops (what the drop-in did before) and on the host CPU
(what the reference does with device=cpu).  Not a bench.py line: a §8(f).1 data point for DESIGN.md."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "off-policy_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import torch.nn.functional as F
import qmix_checks as qc
from oracle.qmix import QmixConfig

cfg = QmixConfig(n_agents=3, obs_dim=30, act_dim=9, state_dim=48, gain=1.0)
args, pol, tr = qc.build_trainer(cfg, 32, 60)
R = cfg.n_agents
obs = np.random.randn(R, cfg.obs_dim).astype(np.float32)
av = np.ones((R, cfg.act_dim), np.float32)


def torch_step(p, o, h, dev):
    H = 64
    x = F.layer_norm(o, (cfg.obs_dim,), p["rnn.feature_norm.weight"], p["rnn.feature_norm.bias"])
    x = F.layer_norm(F.relu(F.linear(x, p["rnn.mlp.fc1.0.weight"], p["rnn.mlp.fc1.0.bias"])), (H,), p["rnn.mlp.fc1.2.weight"], p["rnn.mlp.fc1.2.bias"])
    x = F.layer_norm(F.relu(F.linear(x, p["rnn.mlp.fc2.0.0.weight"], p["rnn.mlp.fc2.0.0.bias"])), (H,), p["rnn.mlp.fc2.0.2.weight"], p["rnn.mlp.fc2.0.2.bias"])
    gi = F.linear(x, p["rnn.rnn.rnn.weight_ih_l0"], p["rnn.rnn.rnn.bias_ih_l0"]); gh = F.linear(h, p["rnn.rnn.rnn.weight_hh_l0"], p["rnn.rnn.rnn.bias_hh_l0"])
    r = torch.sigmoid(gi[:, :H] + gh[:, :H]); z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H]); n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    h2 = (1 - z) * n + z * h
    q = F.linear(F.layer_norm(h2, (H,), p["rnn.rnn.norm.weight"], p["rnn.rnn.norm.bias"]), p["q.action_out.weight"], p["q.action_out.bias"])
    return q, h2


def timeit(fn, n=300, warm=30):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


state = {"h": np.zeros((R, 64), np.float32)}
def ours():
    a, h, _ = pol.get_actions(obs, None, state["h"], av)
    state["h"] = h
pg = dict(pol.q_network.views)
hg = {"h": torch.zeros(R, 64, device="cuda")}
def torch_gpu():
    with torch.no_grad():
        q, h = torch_step(pg, torch.as_tensor(obs).cuda(), hg["h"], "cuda")
        hg["h"] = h
        q.argmax(-1).cpu()
pc = {k: v.cpu() for k, v in pol.q_network.views.items()}
hc = {"h": torch.zeros(R, 64)}
torch.set_num_threads(1)
def torch_cpu():
    with torch.no_grad():
        q, h = torch_step(pc, torch.as_tensor(obs), hc["h"], "cpu")
        hc["h"] = h
        q.argmax(-1)
print(json.dumps(dict(what="us per QMixPolicy.get_actions call (3 agents, obs 30, 9 actions), includes H2D of obs + D2H of actions",
                      k_policy_step_us=timeit(ours), eager_torch_cuda_us=timeit(torch_gpu), eager_torch_cpu_1thread_us=timeit(torch_cpu))))
