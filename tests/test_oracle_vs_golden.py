"""Pin the oracle (oracle/qmix.py) against outputs of the unmodified reference (tests/golden/*.npz).

Tolerance: the oracle runs the same torch ops as the reference, so agreement is expected at
float32 round-off (1e-6 relative on loss, 2e-5 on tensors relative to their max-abs).
"""
import numpy as np
import pytest
import torch

from helpers import load_golden, oracle_from_golden, golden_batch, rel_err

CASES = ["qmix_small", "qmix_small_huber_nodq", "qmix_small_per", "qmix_small_hyper1", "qmix_5ag", "qmix_small_prev_act", "qmix_small_nofn", "qmix_small_tanh"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_step(name):
    torch.set_num_threads(1)
    g = load_golden(name)
    L, cfg, B, T, steps = oracle_from_golden(g)
    for s in range(steps):
        info, prio, _ = L.step(golden_batch(g, s))
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


def test_agent_trace_matches_module_forward():
    from oracle.qmix import agent_trace
    g = load_golden("qmix_small")
    L, cfg, B, T, _ = oracle_from_golden(g)
    x = L.stack_agents(g["s0.in.obs"])
    with torch.no_grad():
        q, hT = L.agent(x)
    tr = agent_trace(L.agent, x)
    assert rel_err(tr["q"], q) < 1e-5
    assert rel_err(tr["h"][-1], hT) < 1e-5
