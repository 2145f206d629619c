"""GPU parity tests for features that were finished after the round's last GPU visit (emulator-verified only so far).  The file
name sorts last on purpose: `pytest -x` reaches these after every test that has already been seen green on a B200."""
import pytest

import qmix_checks as qc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("debug", [True, False])
def test_prev_act_inp_matches_reference_golden(gpu_engine, debug):
    """--prev_act_inp: input width 30 + 9 = 39 (tcgen05 front kernel with K padded to 40)."""
    qc.check_step_against(None, "qmix_small_prev_act", intermediates=False, debug=debug)


@pytest.mark.parametrize("debug", [True, False])
@pytest.mark.parametrize("name", ["mqmix_small", "mqmix_small_per_huber_nodq", "mqmix_small_noavail"])
def test_mqmix_matches_reference_golden(gpu_engine, name, debug):
    """MLP (transition-level) QMIX, SURVEY.md section 8(f).4 first slice."""
    import mqmix_checks as mc
    mc.check_golden(name, debug)


def test_mlp_buffer_sample_layout(gpu_engine):
    import mqmix_checks as mc
    mc.check_buffer_vs_reference_layout()
