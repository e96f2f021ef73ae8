"""The two CPU restatements of the reference (C: oracle/pomo_*.c, numpy: oracle/pomo_np.py) must
agree bitwise on a spun-up state -- the guard that stands in for the golden vectors the
reference does not have (parity unpinned)."""
import numpy as np
import pytest

from extpom_b200 import synthetic as syn
from oracle.pomo import Oracle
from oracle.pomo_np import NP

NEED = ("h fsm dum dvm dx dy art dt d et etb etf ua va uab vab aam2d vfluxb vfluxf z zz dz dzz "
        "u v w aam rmean rho t s kh q2 q2b q2l q2lb drhox drhoy uf vf wtsurf wssurf tsurf ssurf").split()


@pytest.fixture(scope="module")
def spun():
    st, o = syn.seamount(27, 22, 10, Oracle, island=True)
    for i in range(1, 5):
        o.step(i)
    o.set("iint", 5)
    f = {n: o.get(n) for n in NEED}
    c = {n: o.getc(n) for n in ("grav", "rhoref", "tbias", "sbias", "dti2", "umol", "ramp")}
    return o, NP(f, c), f


def test_dens(spun):
    o, n, f = spun
    want = n.dens(f["s"], f["t"])
    o.dens("s", "t", "rho")
    got = o.get("rho")
    # |S|**1.5: numpy's pow and glibc's pow may differ in the last bit
    assert np.abs(got - want).max() <= 2e-16 * np.abs(want).max()
    o.put("rho", f["rho"])


def test_baropg(spun):
    o, n, f = spun
    dx_, dy_, rho_ = n.baropg(f["rho"], f["drhox"], f["drhoy"])
    o.baropg()
    assert np.array_equal(o.get("drhox"), dx_)
    assert np.array_equal(o.get("drhoy"), dy_)
    assert np.array_equal(o.get("rho"), rho_)
    for k in ("rho", "drhox", "drhoy"):
        o.put(k, f[k])


def test_vertvl(spun):
    o, n, f = spun
    want = n.vertvl(f["w"])
    o.vertvl()
    assert np.array_equal(o.get("w"), want)
    o.put("w", f["w"])


def test_advq(spun):
    o, n, f = spun
    z3 = np.zeros_like(f["uf"])
    o.put("uf", z3); o.put("vf", z3)
    o.advq("q2b", "q2", "uf"); o.advq("q2lb", "q2l", "vf")
    assert np.array_equal(o.get("uf"), n.advq(f["q2b"], f["q2"], z3))
    assert np.array_equal(o.get("vf"), n.advq(f["q2lb"], f["q2l"], z3))
    o.put("uf", f["uf"]); o.put("vf", f["vf"])


def test_advave(spun):
    o, n, f = spun
    au, av = n.advave()
    o.advave()
    assert np.array_equal(o.get("advua"), au)
    assert np.array_equal(o.get("advva"), av)


@pytest.mark.parametrize("nbc", [1, 3])
def test_proft(spun, nbc):
    o, n, f = spun
    o.put("uf", f["t"])
    o.proft("uf", "wtsurf", "tsurf", nbc)
    assert np.array_equal(o.get("uf"), n.proft(f["t"], f["wtsurf"], f["tsurf"], nbc))
    o.put("uf", f["uf"])


def test_profq():
    """The heaviest routine (Mellor-Yamada 2.5, two tridiagonal solves): both restatements from
    the same state, after advq filled uf,vf."""
    from oracle.pomo_np import profq
    st, o = syn.seamount(27, 22, 10, Oracle, island=True)
    for i in range(1, 5):
        o.step(i)
    o.set("iint", 5)
    z3 = np.zeros((27, 22, 10), order="F")
    o.put("uf", z3); o.put("vf", z3)
    o.advq("q2b", "q2", "uf"); o.advq("q2lb", "q2l", "vf")
    names = ("h etf z zz dz dzz kq km kh t s rho q2b q2lb q2 u v wusurf wvsurf wubot wvbot l fsm").split()
    f = {n: o.get(n) for n in names}
    c = {n: o.getc(n) for n in ("grav", "rhoref", "tbias", "sbias", "dti2", "umol", "kappa", "small")}
    want = profq(NP(f, c), o.get("uf"), o.get("vf"))
    o.profq()
    for k, a in want.items():
        b = o.get(k)
        assert np.array_equal(a, b), (k, float(np.abs(a - b).max()))


def test_mode_external_all_substeps():
    """The 2-D external mode (advance.f:205-353) with bcond(1), bcond(2), advave, the etf
    accumulation of the last three substeps, the Asselin filter and the running means: both
    restatements through all isplit substeps of one internal step."""
    from oracle.pomo_np import mode_external
    st, o = syn.seamount(27, 22, 10, Oracle, island=True, isplit=8, dte=6.0)
    for i in range(1, 4):
        o.step(i)
    o.set("iint", 4); o.set("time", o.getc("dti") * 4 / 86400.0)
    o.lateral_viscosity(); o.mode_interaction()
    names = ("h fsm dum dvm dx dy art aru arv cor e_atmos d ua va uab vab el elb elf uaf vaf etf egf utf vtf "
             "advua advva adx2d ady2d drx2d dry2d wusurf wvsurf wubot wvbot vfluxf aam2d z "
             "uabw uabe vabw vabe elw ele vabs vabn uabs uabn els eln").split()
    f = {n: o.get(n) for n in names}
    c = {n: o.getc(n) for n in ("grav", "alpha", "dte", "dte2", "smoth", "ramp", "isplit", "ispadv", "ispi",
                                "isp2i", "rfw", "rfe", "rfs", "rfn")}
    for iext in range(1, 9):
        o.mode_external(iext)
        mode_external(f, c, iext)
        for k in ("el", "elb", "d", "ua", "uab", "va", "vab", "uaf", "vaf", "elf", "etf", "egf", "utf", "vtf",
                  "advua", "advva"):
            assert np.array_equal(o.get(k), f[k]), (iext, k, float(np.abs(o.get(k) - f[k]).max()))


@pytest.mark.parametrize("nitera,sw", [(1, 0.5), (3, 1.0)])
def test_advt2_smol_adif(nitera, sw):
    from oracle.pomo_np import advt2
    st, o = syn.seamount(27, 22, 10, Oracle, island=True, nitera=nitera, sw=sw)
    for i in range(1, 5):
        o.step(i)
    names = "h fsm dum dvm dx dy art aru arv dt etb etf u v w aam dz dzz tb t tclim uf".split()
    f = {n: o.get(n) for n in names}
    c = {n: o.getc(n) for n in ("dti2", "sw", "nitera", "tprni")}
    ff, fb = advt2(f, c, f["tb"], f["t"], f["tclim"], f["uf"])
    o.advt2("tb", "t", "tclim", "uf")
    assert np.array_equal(o.get("uf"), ff)
    assert np.array_equal(o.get("tb"), fb)


def test_baropg_mcc(spun):
    from oracle.pomo_np import baropg_mcc
    o, n, f = spun
    dx_, dy_, rho_ = baropg_mcc(n.f, n.c, f["rho"], f["drhox"], f["drhoy"])
    o.baropg_mcc()
    assert np.array_equal(o.get("drhox"), dx_)
    assert np.array_equal(o.get("drhoy"), dy_)
    assert np.array_equal(o.get("rho"), rho_)
    for k in ("rho", "drhox", "drhoy"):
        o.put(k, f[k])
