"""R-MADDPG / R-MATD3 kernels' logic on the CPU fiber emulator vs the reference goldens."""
import pytest

import maddpg_checks as mc


@pytest.mark.parametrize("name", ["maddpg_box", "matd3_box", "maddpg_box_per", "maddpg_disc", "matd3_disc", "matd3_disc_avail", "matd3_disc_nofn", "maddpg_box_tanh"])
def test_step_matches_reference_golden(emu_engine, name):
    mc.check_golden(name)


@pytest.mark.parametrize("name", ["maddpg_box", "maddpg_disc", "matd3_disc", "matd3_disc_nofn", "maddpg_box_tanh"])
def test_rollout_actions_match_reference(emu_engine, name):
    mc.check_get_actions(name)


@pytest.mark.parametrize("td3,disc", [(False, False), (True, True)])
def test_whole_update_graph_matches_eager(emu_engine, td3, disc):
    mc.check_graph_matches_eager(td3, disc)


def test_replay_batch_equals_host_batch_odd_episode_length(emu_engine):
    mc.check_replay_batch_equals_host_batch()


@pytest.mark.parametrize("name", ["maddpg_multi_disc", "matd3_multi_box", "matd3_multi_disc"])
def test_per_agent_policies_match_reference_golden(emu_engine, name):
    """share_policy = False (scripts/train_mpe_rmaddpg.sh:14): one policy per agent, heterogeneous observation / action widths."""
    mc.check_multi_golden(name)


def test_per_agent_policies_through_the_multi_policy_buffer(emu_engine):
    mc.check_multi_golden("maddpg_multi_disc", through_buffer=True)
