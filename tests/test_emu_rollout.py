"""Rollout-time policy step kernel (csrc/rollout.cu) on the CPU fiber emulator vs the reference golden."""
import rollout_checks as rc


def test_qmix_rollout_matches_reference(emu_engine):
    rc.check_rollout()


def test_policy_step_argument_errors(emu_engine):
    rc.check_errors()


def test_in_place_edit_of_the_returned_state_is_honoured(emu_engine):
    rc.check_in_place_state_edit_is_honoured()
