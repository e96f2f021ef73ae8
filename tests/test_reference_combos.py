"""Live differential runs against the reference's own source for namelist / forcing COMBINATIONS the committed
fixtures do not hold (tests/golden/ref_*.npz pin one switch at a time).

Runs wherever the reference can be executed: in the build container from /root/reference, elsewhere from the
translation `__graft_entry__.build()` leaves in oracle/_ref for this grid (oracle/f77ref.py).  Per combination:
the reference's Fortran (executed), the C oracle (must be BITWISE equal) and the host build of the CUDA kernel bodies
(<= 1e-11; the one non-identical operation is |S|**1.5) take three internal steps from the same state, records of
restore_interior handed over step by step as the reference holds them."""
import numpy as np
import pytest

from oracle.f77ref import Reference
from oracle.pomo import Oracle
from scripts import make_ref_golden as mrg
from tests.emu import EmuPom

DIMS = (16, 14, 7)
OPEN = {"walls": False, "fluxes": True, "obc": True}
COMBOS = [
    {"nbct": 1, "nbcs": 2},
    {"nbct": 4, "nbcs": 2, "ntp": 4},
    {"nbct": 3, "nbcs": 4, "fluxes": True},
    {"nbct": 2, "nbcs": 2, "ntp": 2, **OPEN},
    {"_set": {"ispadv": 1, "smoth": 0.0, "alpha": 0.0}},
    {"_set": {"ispadv": 30, "time0": 0.25}},
    {"isplit": 3, "dte": 6.0, "island": True},
    {"island": True, "mode": 4, **OPEN},
    {"nadv": 1, "nitera": 1, "npg": 2, "_set": {"tprni": 0.0, "horcon": 0.05, "ramp": 0.3}, **OPEN},
    {"nitera": 4, "sw": 0.75, "island": True, "fluxes": True},
    {"mode": 2, "island": True, "fluxes": True, "_set": {"ispadv": 2, "time0": 2.0}},
    {"wind": False, "noise": False, "aam_init": 0.0, **OPEN},
]


def _id(kw):
    return "-".join(f"{k}{v}" for k, v in kw.items() if k != "_set") + ("-" + "-".join(f"{k}{v}" for k, v in kw["_set"].items()) if "_set" in kw else "")


@pytest.mark.skipif(not Reference.available(*DIMS), reason="neither the reference source nor its translation for this grid is here")
@pytest.mark.parametrize("kw", COMBOS, ids=_id)
def test_reference_oracle_and_kernel_bodies_agree(kw):
    from oracle.f77ref import F77Ref
    res = {}
    for name, F in (("ref", F77Ref), ("oracle", Oracle), ("emu", EmuPom)):
        st, g = mrg.loaded(F, DIMS, kw)
        for i in range(1, 4):
            if name != "ref":
                mrg.ref_restore_records(g, st, i)
            g.step(i)
        res[name] = {n: g.get(n) for n in mrg.F3 + mrg.F2}
    assert all(np.isfinite(v).all() for v in res["ref"].values())
    assert np.abs(res["ref"]["u"]).max() > 0 and np.abs(res["ref"]["el"]).max() > 0
    for n, a in res["ref"].items():
        assert np.array_equal(a, res["oracle"][n]), ("oracle", n)
        if n in ("uf", "vf"):                 # work arrays: the CUDA path rotates pointers instead of copying
            continue
        b = res["emu"][n]
        if n in ("t", "tb", "s", "sb"):       # level kb is scratch (tests/common.py)
            a, b = a[:, :, :-1], b[:, :, :-1]
        assert np.abs(a - b).max() <= 1e-11 * (np.abs(a).max() + 1e-300), ("kernel bodies", n)
