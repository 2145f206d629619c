"""Rollout-time policy step kernel (csrc/rollout.cu) on the CPU fiber emulator vs the reference golden."""
import rollout_checks as rc


def test_qmix_rollout_matches_reference(emu_engine):
    rc.check_rollout()


def test_policy_step_argument_errors(emu_engine):
    rc.check_errors()
