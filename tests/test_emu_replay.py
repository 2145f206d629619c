"""Replay kernels' logic on the CPU fiber emulator (tests/emu) vs reference goldens and the oracle."""
import pytest

import replay_checks as rc


def test_uniform_golden(emu_engine):
    rc.check_uniform_golden()


def test_device_rng_stream(emu_engine):
    rc.check_device_rng_stream()


def test_per_golden(emu_engine):
    rc.check_per_golden()


def test_per_vs_oracle_random(emu_engine):
    rc.check_per_vs_oracle_random()


@pytest.mark.parametrize("shape", [(3, 30, 9, 48, 5, 11, 4, True), (2, 5, 3, 7, 3, 6, 6, False), (1, 1, 2, 1, 1, 3, 2, True)])
def test_uniform_vs_oracle(emu_engine, shape):
    rc.check_uniform_vs_oracle_shapes(shape)


def test_reward_norm_running_statistics(emu_engine):
    rc.check_reward_norm_running_stats()
