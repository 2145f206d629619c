"""Pin the MLP-path oracle (oracle/mqmix.py) against outputs of the unmodified reference's M_QMix (tests/golden/mqmix_*.npz,
make_goldens.py mqmix).  Same torch ops as the reference, so agreement is expected at float32 round-off."""
import numpy as np
import pytest
import torch

from helpers import load_golden, golden_cfg, sub, rel_err

CASES = ["mqmix_small", "mqmix_small_per_huber_nodq", "mqmix_small_noavail", "mqmix_small_nofn", "mqmix_small_tanh"]
TFIELDS = ["obs", "share", "acts", "rew", "nobs", "nshare", "dones", "dones_env", "valid", "avail", "navail"]


def oracle_from_golden(g):
    from oracle.mqmix import MqmixLearner
    cfg, B, T, steps = golden_cfg(g)
    assert T == 1
    L = MqmixLearner(cfg)
    L.agent.load_state_dict(sub(g, "init.agent."))
    L.mixer.load_state_dict(sub(g, "init.mixer."))
    L.tgt_agent.load_state_dict(sub(g, "init.tgt_agent."))
    L.tgt_mixer.load_state_dict(sub(g, "init.tgt_mixer."))
    return L, cfg, B, steps


def golden_transitions(g, s):
    b = tuple(g.get("s%d.in.%s" % (s, k)) for k in TFIELDS)
    w = g.get("s%d.in.weights" % s)
    return b + (w, np.arange(b[0].shape[1]) if w is not None else None)


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_mlp_step(name):
    torch.set_num_threads(1)
    g = load_golden(name)
    L, cfg, B, steps = oracle_from_golden(g)
    for s in range(steps):
        info, prio, _ = L.step(golden_transitions(g, s))
        assert rel_err(info["loss"], g["s%d.loss" % s]) < 1e-6
        assert rel_err(info["grad_norm"], g["s%d.grad_norm" % s]) < 1e-5
        assert rel_err(info["Q_tot"], g["s%d.Q_tot" % s]) < 1e-5
        if cfg.use_per:
            assert rel_err(prio, g["s%d.prio" % s]) < 1e-5
        for k, p in L.agent.named_parameters():
            key = "s%d.grad.agent.%s" % (s, k)
            if key in g:
                assert rel_err(p.grad, g[key]) < 2e-5, key
            else:
                assert p.grad is None and "fc_h" in k
        for k, p in L.mixer.named_parameters():
            assert rel_err(p.grad, g["s%d.grad.mixer.%s" % (s, k)]) < 2e-5, k
        L.soft_update()
        for tag, mod in (("agent", L.agent), ("mixer", L.mixer), ("tgt_agent", L.tgt_agent), ("tgt_mixer", L.tgt_mixer)):
            for k, v in mod.state_dict().items():
                assert rel_err(v, g["s%d.%s.%s" % (s, tag, k)]) < 2e-6, (tag, k)
    with torch.no_grad():
        assert rel_err(L.agent(torch.from_numpy(g["roll.obs"])), g["roll.q_all"]) < 1e-5


RFIELDS = ["obs", "share", "acts", "rew", "nobs", "nshare", "dones", "dones_env", "valid", "avail", "navail"]


@pytest.mark.parametrize("tag,norm,avail", [("plain", False, True), ("norm", True, False)])
def test_transition_replay_reproduces_reference_buffer(tag, norm, avail):
    """oracle.mqmix.TransitionReplay vs the reference's MlpReplayBuffer (golden mlp_replay_small): ring indices and sampled arrays bit-equal
    (normalised rewards: same float32 arithmetic, so also bit-equal)."""
    from oracle.mqmix import TransitionReplay, transition_replay_script
    g = load_golden("mlp_replay_small")
    N, O, A, S, E = [int(v) for v in g["meta.shape"]]
    buf = TransitionReplay(E, N, O, S, A, use_avail=avail, reward_norm=norm)
    np.random.seed(11)
    ns = ni = 0
    for op, n, f in transition_replay_script():
        if op == "insert":
            idx = buf.insert(n, *[f[k] for k in RFIELDS])
            assert np.array_equal(idx, g["%s.idx%d" % (tag, ni)])
            ni += 1
        else:
            out, _ = buf.sample(n)
            for i, name in enumerate(RFIELDS):
                key = "%s.s%d.%s" % (tag, ns, name)
                if out[i] is None:
                    assert key not in g
                else:
                    assert np.array_equal(out[i], g[key]), key
            ns += 1
    assert ns == int(g["%s.n_samples" % tag])
