"""CPU tests of the oracle itself: committed golden fixtures and physical invariants.

The reference ships no tests or vectors, so the goldens are outputs of the
oracle (scripts/make_golden.py) and pin it against regressions only -- PARITY UNPINNED.
"""
import os

import numpy as np
import pytest

from extpom_b200 import synthetic as syn
from oracle.pomo import Oracle
from scripts.make_golden import CASES
from tests.common import F2, F3, digest, rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(name):
    c = CASES[name]
    z = np.load(os.path.join(GOLD, name + ".npz"))
    st, o = syn.seamount(*c["dims"], Oracle, **c["kw"])
    # the generated input must be the one the fixture was made from
    want = dict(s.split(":") for s in z["input_digest"])
    got = {k: digest(v) for k, v in st["fields"].items()}
    assert got == want
    for i in range(1, c["steps"] + 1):
        o.step(i)
    for n in F3 + F2:
        # same binary, same arithmetic: bitwise, except libm's pow/exp across glibc builds
        assert rel_err(z[n], o.get(n)) <= 1e-13, n
    assert abs(float(z["vamax"]) - o.check_velocity()) <= 1e-13


def test_state_of_rest_stays_at_rest():
    """Uniform T,S (so rho == rmean and the baroclinic PG is exactly zero), no wind, no
    inflow: the model must stay at rest over the seamount."""
    im, jm, kb = 30, 24, 10
    st = syn.make_state(im, jm, kb, wind=False, noise=False)
    f = st["fields"]
    for n in ("tb", "t", "tclim"): f[n][...] = 10.0
    for n in ("ub", "u", "uab", "ua"): f[n][...] = 0.0
    for n in ("uabe", "uabw"): f[n][...] = 0.0
    f["tsurf"][...] = 10.0
    for n in ("tbe", "tbw", "tbn", "tbs"): f[n][:, :kb - 1] = 10.0
    o = Oracle(im, jm, kb)
    o.load(st)
    syn.finish_init(st, o)
    for i in range(1, 11):
        o.step(i)
    for n in ("u", "v", "ua", "va", "el", "w"):
        assert np.abs(o.get(n)).max() <= 1e-10, n
    assert np.abs(o.get("t")[:, :, :kb - 1] - 10.0 * f["fsm"][:, :, None]).max() <= 1e-9


def test_uniform_salinity_stays_nearly_uniform():
    """S == sclim == 35 with S-flux 0: advt2 + proft must keep S uniform up to the leapfrog /
    Asselin inconsistency of the free surface (observed ~1e-5 over 20 steps)."""
    st, o = syn.seamount(33, 25, 11, Oracle)
    for i in range(1, 21):
        o.step(i)
    s = o.get("s")[:, :, :10]
    wet = st["fields"]["fsm"][:, :, None] * np.ones_like(s) > 0
    assert np.abs(s[wet] - 35.0).max() < 1e-4


def test_long_run_is_stable_and_finite():
    st, o = syn.seamount(40, 31, 12, Oracle)
    for i in range(1, 101):
        o.step(i)
    assert o.check_velocity() < 1.0 and o.getc("error_status") == 0
    for n in ("u", "v", "t", "s", "q2", "q2l", "km", "kh", "el"):
        assert np.isfinite(o.get(n)).all(), n


def test_single_precision_literal_quirks():
    """(15.8*cbcnst)**(2./3.) with promoted single-precision literals."""
    v = (float(np.float32(15.8)) * 100.0) ** float(np.float32(2.0) / np.float32(3.0))
    assert abs(v - 135.65572445446264) < 1e-12
    assert float(np.float32(0.1)) == 0.10000000149011612
