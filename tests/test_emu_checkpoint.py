"""Checkpoint / resume on the CPU fiber emulator: the restored run continues bit-identically."""
import pytest

import checkpoint_checks as cc


@pytest.mark.parametrize("per,device_rng", [(False, False), (True, False), (False, True)])
def test_resume_is_bit_identical(emu_engine, per, device_rng):
    cc.check_resume(per, device_rng, n=2)


def test_checkpoint_of_another_configuration_is_rejected(emu_engine):
    cc.check_rejects_wrong_shape()
