"""MLP (transition-level) path -- SURVEY.md section 8(f).4, first slice: M_QMix / M_VDN + MlpReplayBuffer on the CPU emulator."""
import pytest

import mqmix_checks as mc


@pytest.mark.parametrize("debug", [True, False])
@pytest.mark.parametrize("name", ["mqmix_small", "mqmix_small_per_huber_nodq", "mqmix_small_noavail", "mqmix_small_nofn", "mqmix_small_tanh"])
def test_mqmix_matches_reference_golden(emu_engine, name, debug):
    mc.check_golden(name, debug)


def test_mlp_buffer_sample_layout(emu_engine):
    mc.check_buffer_vs_reference_layout()


@pytest.mark.parametrize("kw", [dict(), dict(avail=True, per=True, huber=True), dict(avail=True, double_q=False), dict(vdn=True), dict(hyper_layers=1, N=5, O=11, A=7, S=23)],
                         ids=["mpe", "avail_per_huber", "avail_nodq", "vdn", "hyper1_5ag"])
def test_mlp_learner_vs_oracle(emu_engine, kw):
    mc.check_vs_oracle(B=48, steps=2, **kw)


def test_mlp_learner_vs_oracle_big_obs(emu_engine):
    """obs_dim > 64 (SMAC-like shapes): the FFMA front kernel in mlp mode."""
    mc.check_vs_oracle(B=40, steps=1, N=5, O=80, A=11, S=120, avail=True)


@pytest.mark.parametrize("per", [False, True], ids=["uniform", "per"])
def test_mlp_step_graph_sequence(emu_engine, per):
    mc.check_step_graph_vs_eager(per=per)


def test_mlp_buffer_vs_reference_golden(emu_engine):
    mc.check_buffer_vs_reference_golden()
