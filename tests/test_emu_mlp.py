"""MLP (transition-level) path -- SURVEY.md section 8(f).4, first slice: M_QMix / M_VDN + MlpReplayBuffer on the CPU emulator."""
import pytest

import mqmix_checks as mc


@pytest.mark.parametrize("debug", [True, False])
@pytest.mark.parametrize("name", ["mqmix_small", "mqmix_small_per_huber_nodq", "mqmix_small_noavail"])
def test_mqmix_matches_reference_golden(emu_engine, name, debug):
    mc.check_golden(name, debug)


def test_mlp_buffer_sample_layout(emu_engine):
    mc.check_buffer_vs_reference_layout()
