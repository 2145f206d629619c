"""QMIX learner kernels' logic on the CPU fiber emulator vs the reference goldens (and the oracle's trace)."""
import pytest

import qmix_checks as qc


@pytest.mark.parametrize("name", ["qmix_small", "qmix_small_huber_nodq", "qmix_small_per", "qmix_small_hyper1", "qmix_5ag"])
def test_step_matches_reference_golden(emu_engine, name):
    qc.check_step_against(None, name)
