"""R-MADDPG / R-MATD3 kernels' logic on the CPU fiber emulator vs the reference goldens."""
import pytest

import maddpg_checks as mc


@pytest.mark.parametrize("name", ["maddpg_box", "matd3_box", "maddpg_box_per"])
def test_step_matches_reference_golden(emu_engine, name):
    mc.check_golden(name)
