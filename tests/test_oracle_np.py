"""The two CPU restatements of the reference (C: oracle/pomo_*.c, numpy: oracle/pomo_np.py) must
agree bitwise on a spun-up state -- the round-1 guard from before the reference's own source could be
executed (oracle/f77ref.py, tests/golden/ref_*.npz pin the oracle now)."""
import numpy as np
import pytest

from extpom_b200 import synthetic as syn
from oracle.pomo import Oracle
from oracle.pomo_np import NP

NEED = ("h fsm dum dvm dx dy art dt d et etb etf ua va uab vab aam2d vfluxb vfluxf z zz dz dzz "
        "u v w aam rmean rho t s kh q2 q2b q2l q2lb drhox drhoy uf vf wtsurf wssurf tsurf ssurf swrad").split()


@pytest.fixture(scope="module")
def spun():
    st, o = syn.seamount(27, 22, 10, Oracle, island=True)
    for i in range(1, 5):
        o.step(i)
    o.set("iint", 5)
    f = {n: o.get(n) for n in NEED}
    c = {n: o.getc(n) for n in ("grav", "rhoref", "tbias", "sbias", "dti2", "umol", "ramp", "ntp")}
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


@pytest.mark.parametrize("nbc", [1, 2, 3, 4])
def test_proft(spun, nbc):
    o, n, f = spun
    o.put("uf", f["t"])
    o.proft("uf", "wtsurf", "tsurf", nbc)
    want = n.proft(f["t"], f["wtsurf"], f["tsurf"], nbc)
    if nbc in (1, 3):
        assert np.array_equal(o.get("uf"), want)
    else:   # short-wave penetration: quad-precision exp in the reference / C oracle, long double here
        assert np.abs(o.get("uf") - want).max() <= 4e-16 * np.abs(want).max()
        assert np.abs(want - n.proft(f["t"], f["wtsurf"], f["tsurf"], nbc - 1)).max() > 1e-9
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


MOM = ("h fsm dum dvm dx dy art aru arv cor dt d et etb etf egf egb e_atmos cbc wusurf wvsurf z zz dz dzz "
       "u v w ub vb aam km advx advy drhox drhoy uf vf").split()
MOMC = ("grav", "dti2", "umol", "horcon")


@pytest.fixture(scope="module")
def mom():
    """State in the middle of internal step 5, just before advu/advv (u, v adjusted, w from vertvl)."""
    st, o = syn.seamount(27, 22, 10, Oracle, island=True)
    for i in range(1, 5):
        o.step(i)
    o.set("iint", 5); o.set("time", o.getc("dti") * 5 / 86400.0)
    return o


def _snap(o):
    return {n: o.get(n) for n in MOM}, {n: o.getc(n) for n in MOMC}


def test_advct_and_smagorinsky(mom):
    from oracle.pomo_np import advct, smagorinsky
    o = mom
    f, c = _snap(o)
    ax, ay = advct(f, c)
    aam = smagorinsky(f, c)
    o.lateral_viscosity()          # advct; baropg; aam (advance.f:110-138)
    assert np.array_equal(o.get("advx"), ax)
    assert np.array_equal(o.get("advy"), ay)
    assert np.array_equal(o.get("aam"), aam)
    assert np.abs(ax).max() > 0 and np.abs(aam - f["aam"]).max() > 0


def test_advu_advv_profu_profv(mom):
    from oracle.pomo_np import advu, advv, profu, profv
    o = mom
    o.mode_interaction()
    for ie in range(1, int(o.getc("isplit")) + 1):
        o.mode_external(ie)
    for stg in range(0, 11):       # up to and including the t/s block: u, v adjusted, w, km updated
        o.internal_stage(5, stg)
    f, c = _snap(o)
    uf, vf = advu(f, c), advv(f, c)
    o.advu(); o.advv()
    assert np.array_equal(o.get("uf"), uf)
    assert np.array_equal(o.get("vf"), vf)
    f["wusurf"], f["wvsurf"] = o.get("wusurf"), o.get("wvsurf")
    uf2, wub = profu(f, c, uf)
    vf2, wvb = profv(f, c, vf)
    o.profu(); o.profv()
    assert np.array_equal(o.get("uf"), uf2)
    assert np.array_equal(o.get("vf"), vf2)
    assert np.array_equal(o.get("wubot")[1:-1, 1:-1], wub)
    assert np.array_equal(o.get("wvbot")[1:-1, 1:-1], wvb)
    assert np.abs(uf2 - uf).max() > 0


def test_realvertvl(mom):
    from oracle.pomo_np import realvertvl
    o = mom
    f, c = _snap(o)
    want = realvertvl(f, c)
    o.realvertvl()
    assert np.array_equal(o.get("wr"), want)
    assert np.abs(want).max() > 0


ALLC = ("alpha dte dti dti2 grav kappa ramp rfe rfn rfs rfw rhoref sbias small tbias time tprni umol vmaxl dte2 "
        "horcon ispi isp2i smoth sw time0 iint mode ntp ispadv isplit nadv nbct nbcs nitera npg").split()
STATE = ("aam advx advy drhox drhoy kh km kq l q2b q2 q2lb q2l rho sb s tb t ub uf u vb vf v w wr "
         "aam2d advua advva adx2d ady2d d drx2d dry2d dt egb egf el elb elf et etb etf ua uab uaf utb utf va vab "
         "vaf vtb vtf vfluxb wubot wvbot").split()


@pytest.mark.parametrize("kw", [{}, {"nitera": 2, "sw": 1.0}, {"nbct": 3, "nbcs": 3}, {"nadv": 1}],
                         ids=["default", "nitera2", "nbc3", "nadv1"])
def test_whole_step_second_restatement(kw):
    """One WHOLE internal step (advance.f:21-32: lateral_viscosity, mode_interaction, isplit x
    mode_external, mode_internal with every solver.f routine and bcond / bcondorl call) through the
    numpy restatement against the C oracle, from the same spun-up state, for three consecutive steps
    (re-synchronised each step).  Everything must agree bitwise; only rho (|S|**1.5 through two
    different pow implementations) gets one ulp."""
    from oracle import pomo_np
    st, o = syn.seamount(25, 20, 9, Oracle, island=True, isplit=6, dte=6.0, **kw)
    for i in range(1, 4):
        o.step(i)
    for i in range(4, 7):
        o.set("iint", i); o.set("time", o.getc("dti") * i / 86400.0)
        f = {n: o.get(n) for n in o.f}
        c = {n: o.getc(n) for n in ALLC}
        pomo_np.step(f, c, i)
        o.step(i)
        for n in STATE:
            a, b = o.get(n), f[n]
            if n == "rho":
                assert np.abs(a - b).max() <= 3e-16 * np.abs(a).max(), n
            elif n in ("uf", "vf"):
                # level kb of the work arrays keeps whatever the previous user left there
                assert np.array_equal(a[:, :, :-1], b[:, :, :-1]), (i, n, float(np.abs(a - b).max()))
            else:
                assert np.array_equal(a, b), (i, n, float(np.abs(a - b).max()))


def test_whole_step_with_restoring():
    """restore_interior's nudging (bounds_forcing.f:1083-1118) inside the whole step, both restatements."""
    from oracle import pomo_np
    st, o = syn.seamount(25, 20, 9, Oracle, island=True, isplit=6, dte=6.0)
    rng = np.random.default_rng(5)
    fl = st["fields"]
    for n, a in (("trstrb", fl["tclim"] + 0.5), ("trstrf", fl["tclim"] - 0.25), ("srstrb", fl["sclim"] + 0.1),
                 ("srstrf", fl["sclim"] - 0.05), ("taurstrb", 0.2 + 0.1 * rng.random(fl["tclim"].shape)),
                 ("taurstrf", 0.3 + 0.1 * rng.random(fl["tclim"].shape))):
        o.put(n, np.asfortranarray(a))
    o.set("lrestore", 1)
    for i in range(1, 4):
        o.step(i)
    i = 4
    o.set("iint", i); o.set("time", o.getc("dti") * i / 86400.0)
    f = {n: o.get(n) for n in o.f}
    c = {n: o.getc(n) for n in ALLC + ["lrestore"]}
    pomo_np.step(f, c, i)
    o.step(i)
    for n in ("t", "tb", "s", "sb", "u", "v", "q2", "el"):
        assert np.array_equal(o.get(n), f[n]), n


def test_mode2_external_substeps():
    """mode=2 (2-D only): advave's bottom-stress / curvature block (solver.f:123-195) inside the
    external substeps, both restatements."""
    from oracle.pomo_np import mode_external
    st, o = syn.seamount(27, 22, 10, Oracle, island=True, isplit=8, dte=6.0, mode=2)
    for i in range(1, 3):
        o.step(i)
    o.set("iint", 3); o.set("time", o.getc("dti") * 3 / 86400.0)
    o.lateral_viscosity(); o.mode_interaction()
    names = ("h fsm dum dvm dx dy art aru arv cor cbc e_atmos d ua va uab vab el elb elf uaf vaf etf egf utf vtf "
             "advua advva adx2d ady2d drx2d dry2d wusurf wvsurf wubot wvbot vfluxf aam2d z "
             "uabw uabe vabw vabe elw ele vabs vabn uabs uabn els eln").split()
    f = {n: o.get(n) for n in names}
    c = {n: o.getc(n) for n in ("grav", "alpha", "dte", "dte2", "smoth", "ramp", "isplit", "ispadv", "ispi",
                                "isp2i", "rfw", "rfe", "rfs", "rfn", "mode")}
    for iext in range(1, 9):
        o.mode_external(iext)
        mode_external(f, c, iext)
        for k in ("el", "ua", "va", "uaf", "vaf", "elf", "advua", "advva", "wubot", "wvbot", "utf", "vtf"):
            assert np.array_equal(o.get(k), f[k]), (iext, k, float(np.abs(o.get(k) - f[k]).max()))
    assert np.abs(f["wubot"]).max() > 0
