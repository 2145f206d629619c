"""GPU parity tests for features that were finished after the round's last GPU visit (emulator-verified only so far).  The file
name sorts last on purpose: `pytest -x` reaches these after every test that has already been seen green on a B200."""
import pytest

import qmix_checks as qc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("debug", [True, False])
def test_prev_act_inp_matches_reference_golden(gpu_engine, debug):
    """--prev_act_inp: input width 30 + 9 = 39 (tcgen05 front kernel with K padded to 40)."""
    qc.check_step_against(None, "qmix_small_prev_act", intermediates=False, debug=debug)
