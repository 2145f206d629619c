"""Diagnostic (GPU): M_QMix step vs the oracle, printing per-tensor gradient errors instead of asserting.
Used to decide whether test_mlp_learner_vs_oracle[avail_per_huber] fails on a kernel error or on lock-step drift."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "off-policy_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import qmix_checks as qc
import mqmix_checks as mc
from replay_checks import Discrete


def run(seed0, steps, front_tc=1, debug=False, B=1000, resync=False, N=3, O=18, A=5, S=54, avail=True, per=True, huber=True):
    from oracle.qmix import QmixConfig, randomize_all
    from oracle.mqmix import MqmixLearner, synth_transitions
    from offpolicy.algorithms.mqmix.algorithm.mQMixPolicy import M_QMixPolicy as Pol
    from offpolicy.algorithms.mqmix.mqmix import M_QMix as Tr
    from offpolicy._b200 import capi
    capi.lib().mx_set_option(b"front_tc", front_tc)
    cfg = QmixConfig(n_agents=N, obs_dim=O, act_dim=A, state_dim=S, gain=1.0, use_per=per, huber=huber, huber_delta=0.7)
    L = MqmixLearner(cfg, seed=3)
    randomize_all(L.agent, 1); randomize_all(L.mixer, 2); L.sync_targets()
    randomize_all(L.tgt_agent, 3, 0.05); randomize_all(L.tgt_mixer, 4, 0.05)
    args = qc.make_args(cfg, B)
    info = dict(obs_space=[O], share_obs_space=[S], act_space=Discrete(A), cent_obs_dim=S, cent_act_dim=A * N)
    pol = Pol({"args": args, "device": capi.device()}, info)
    tr = Tr(args, N, {"policy_0": pol}, lambda a: "policy_0", device=capi.device())
    capi.lib().mx_qmix_set_debug(tr.handle, 1 if debug else 0)
    pol.q_network.load_state_dict(L.agent.state_dict()); tr.target_q_network.load_state_dict(L.tgt_agent.state_dict())
    tr.mixer.load_state_dict(L.mixer.state_dict()); tr.target_mixer.load_state_dict(L.tgt_mixer.state_dict())
    for s in range(steps):
        b = synth_transitions(cfg, B, seed=seed0 + s, avail=avail)
        w = (np.random.RandomState(seed0 + 10 + s).rand(B) * 0.9 + 0.1) if per else None
        info_t, prio, _ = tr.train_policy_on_batch(mc._to_dicts(b, w), True)
        gv = {k: v.clone() for k, v in tr.grad_views().items()}
        tr.soft_target_updates()
        ref, rprio, _ = L.step(b + (w, None))
        coef = min(1.0, cfg.max_grad_norm / (float(ref["grad_norm"]) + 1e-6))
        L.soft_update()
        worst = []
        named = dict(("agent." + k, p) for k, p in L.agent.named_parameters())
        named.update(("mixer." + k, p) for k, p in L.mixer.named_parameters())
        for k, p in named.items():
            if p.grad is None:
                continue
            ok, err, lim = qc.close(gv[k] * coef, p.grad, 1e-4)
            worst.append((err / lim, k))
        worst.sort(reverse=True)
        pmax = max(float((v.cpu() - L.agent.state_dict()[k]).abs().max()) for k, v in pol.q_network.state_dict().items())
        print("seed0=%d step=%d front_tc=%d debug=%d resync=%d: loss %.3e gn %.3e qtot %.3e | worst grad err/lim %.2f (%s), #over %d | param diff %.2e (lim %.2e)" % (
            seed0, s, front_tc, debug, resync, mc.rel_err(info_t["loss"].cpu(), ref["loss"]), mc.rel_err(info_t["grad_norm"].cpu(), ref["grad_norm"]),
            mc.rel_err(info_t["Q_tot"].cpu(), ref["Q_tot"]), worst[0][0], worst[0][1], sum(1 for x in worst if x[0] > 1), pmax, 5e-3 * cfg.lr * (s + 1)), flush=True)
        if resync:
            # put the oracle's parameters into the engine so that step s+1 starts from identical weights (Adam moments still differ by round-off)
            pol.q_network.load_state_dict(L.agent.state_dict()); tr.target_q_network.load_state_dict(L.tgt_agent.state_dict())
            tr.mixer.load_state_dict(L.mixer.state_dict()); tr.target_mixer.load_state_dict(L.tgt_mixer.state_dict())


if __name__ == "__main__":
    from offpolicy._b200 import capi
    capi.lib()
    for ftc in (1, 0):
        run(50, 2, front_tc=ftc)
        run(50, 2, front_tc=ftc, resync=True)
        run(51, 1, front_tc=ftc)
    run(50, 2, debug=True)
    run(50, 2, B=64)
    for sd in (70, 80, 90):
        run(sd, 2)
        run(sd, 2, resync=True)
    capi.lib().mx_set_option(b"front_tc", 1)
