"""Rollout-time policy step kernel (csrc/rollout.cu) on the B200 vs the reference golden, through the drop-in QMixPolicy."""
import numpy as np
import pytest
import torch

import rollout_checks as rc

pytestmark = pytest.mark.gpu


def test_qmix_rollout_matches_reference(gpu_engine):
    rc.check_rollout()


def test_policy_step_argument_errors(gpu_engine):
    rc.check_errors()


def test_many_rows_and_wide_obs_vs_torch(gpu_engine):
    """rows > grid cap (grid-stride over rows) and an observation wider than the thread block (SMAC 8m fork default: 204)."""
    import ctypes as C
    from offpolicy._b200 import capi
    from offpolicy._b200.rollout import PolicyStepper
    import torch.nn.functional as F
    from offpolicy.algorithms.qmix.algorithm.QMixPolicy import qmix_cfg_struct, param_entries
    import qmix_checks as qc
    from oracle.qmix import QmixConfig
    cfg = QmixConfig(n_agents=8, obs_dim=204, act_dim=14, state_dim=168)
    args, pol, tr = qc.build_trainer(cfg, 4, 8)
    g = torch.Generator().manual_seed(3)
    for k, v in pol.q_network.views.items():
        v.copy_(torch.randn(v.shape, generator=g).to(v.device) * (0.2 if v.dim() == 2 else 0.5) + (1.0 if "norm" in k or k.endswith("2.weight") else 0.0))
    R = 148 * 4 + 37
    obs = torch.randn(R, cfg.obs_dim, generator=g)
    h = torch.randn(R, 64, generator=g) * 0.5
    avail = (torch.rand(R, cfg.act_dim, generator=g) < 0.5).float()
    avail[:, 0] = 1
    st = PolicyStepper(cfg.obs_dim, cfg.act_dim)
    q, h2, gi, gq = st.step(pol.q_network.flat, obs.numpy(), h.numpy(), avail.numpy())
    p = {k: v.double().cpu() for k, v in pol.q_network.views.items()}
    x = F.layer_norm(obs.double(), (cfg.obs_dim,), p["rnn.feature_norm.weight"], p["rnn.feature_norm.bias"])
    x = F.layer_norm(F.relu(F.linear(x, p["rnn.mlp.fc1.0.weight"], p["rnn.mlp.fc1.0.bias"])), (64,), p["rnn.mlp.fc1.2.weight"], p["rnn.mlp.fc1.2.bias"])
    x = F.layer_norm(F.relu(F.linear(x, p["rnn.mlp.fc2.0.0.weight"], p["rnn.mlp.fc2.0.0.bias"])), (64,), p["rnn.mlp.fc2.0.2.weight"], p["rnn.mlp.fc2.0.2.bias"])
    gi_ = F.linear(x, p["rnn.rnn.rnn.weight_ih_l0"], p["rnn.rnn.rnn.bias_ih_l0"])
    gh_ = F.linear(h.double(), p["rnn.rnn.rnn.weight_hh_l0"], p["rnn.rnn.rnn.bias_hh_l0"])
    r = torch.sigmoid(gi_[:, :64] + gh_[:, :64]); z = torch.sigmoid(gi_[:, 64:128] + gh_[:, 64:128])
    n = torch.tanh(gi_[:, 128:] + r * gh_[:, 128:])
    href = (1 - z) * n + z * h.double()
    qref = F.linear(F.layer_norm(href, (64,), p["rnn.rnn.norm.weight"], p["rnn.rnn.norm.bias"]), p["q.action_out.weight"], p["q.action_out.bias"])
    assert np.abs(h2 - href.numpy()).max() <= 2e-5
    assert np.abs(q - qref.numpy()).max() <= 1e-4 * max(1.0, float(qref.abs().max()))
    masked = qref.clone(); masked[avail == 0] = -1e10
    top2 = masked.topk(2, dim=-1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-3            # rows whose arg-max is not a near-tie
    assert np.array_equal(gi[clear.numpy()], masked.argmax(-1).numpy()[clear.numpy()])
    assert np.abs(gq - masked.max(-1).values.numpy()).max() <= 1e-4 * max(1.0, float(qref.abs().max()))
