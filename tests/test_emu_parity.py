"""CPU tests: libpomgpu's kernel bodies + host orchestration (host-emulated build, see
tests/emu.py) against the oracle.  The GPU runs of the same checks are in test_gpu_parity.py."""
import numpy as np
import pytest

from extpom_b200 import synthetic as syn
from oracle.pomo import Oracle
from tests import parity_cases as pc
from tests.common import assert_close
from tests.emu import EmuPom


@pytest.mark.parametrize("case", pc.STEP_CASES, ids=pc.case_id)
def test_steps_match_oracle(case):
    pc.check_steps(EmuPom, case)


def test_bitwise_equal_over_40_steps_with_common_pow():
    assert pc.check_steps(EmuPom, ((60, 50, 20), 40, {}), pow_mode=1, tol=0.0) == 0.0
    pc.check_steps(EmuPom, ((60, 50, 20), 40, {}), pow_mode=0, tol=1e-9)


def test_stage_by_stage_is_bitwise_equal():
    pc.check_stages(EmuPom, (26, 21, 10), nstep=3)


@pytest.mark.parametrize("routine", pc.ROUTINES)
def test_routine_matches_oracle(routine):
    pc.check_routine(EmuPom, routine, (28, 22, 10))


@pytest.mark.parametrize("name", pc.GOLDEN)
def test_matches_golden(name):
    pc.check_golden(EmuPom, name)


@pytest.mark.parametrize("name", pc.REF_GOLDEN)
def test_matches_the_references_own_output(name):
    """Host build of the CUDA kernel bodies against what the reference's own source computes (tests/golden/ref_*)."""
    pc.check_ref_golden(EmuPom, name, tol=1e-11)


def test_unsupported_switches_set_error_status():
    st, g = syn.seamount(20, 17, 8, EmuPom, npg=3)
    with pytest.raises(Exception):
        g.step(1)
    assert g.getc("error_status") == 1


def test_restore_interior_matches_oracle():
    pc.check_restore(EmuPom)


def test_domain_stats_matches_oracle():
    pc.check_domain_stats(EmuPom)


def test_forcing_interpolation_matches_oracle():
    pc.check_forcing_interp(EmuPom)


def test_push_of_u_v_between_steps():
    pc.check_push_midrun(EmuPom)
