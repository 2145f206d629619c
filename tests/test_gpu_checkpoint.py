"""Checkpoint / resume on the B200: the restored run continues bit-identically (graph-replayed learner step)."""
import pytest

import checkpoint_checks as cc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("per,device_rng", [(False, False), (True, False), (False, True)])
def test_resume_is_bit_identical(gpu_engine, per, device_rng):
    cc.check_resume(per, device_rng)


def test_checkpoint_of_another_configuration_is_rejected(gpu_engine):
    cc.check_rejects_wrong_shape()
