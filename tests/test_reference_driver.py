"""The reference's OWN driver loop over the drop-in library ("same names, same call sites").

`advance` (pom/advance.f:6-59) is executed from the reference source by oracle/f77ref.py twice, three... steps each,
from the same state and the same synthetic "files" (stand-ins for the four PnetCDF readers):

  A. the unmodified reference: get_time, surface_forcing (wind, heat, surface), lateral_bc, the four step routines,
     print_section, check_velocity -- all Fortran, executed;
  B. the GPU build of INTEGRATION.md 1.1: the glue file scripts/make_glue.py writes (advance.f without the four step
     routines + `restore_interior_records`) executed the same way, and the four routines it calls --
     lateral_viscosity, mode_interaction, mode_external, mode_internal -- bound to the symbols of libpomgpu_f
     (host-emulated kernel bodies), state in the COMMON blocks the library is linked to, the record half of
     restore_interior called back from inside mode_internal_.

B must reproduce A (<= 1e-11; the one non-identical operation is |S|**1.5): the Fortran glue computes the forcing
on the host every step, the library picks it up from COMMON, the glue's check_velocity reads the vaf the library
leaves.  Needs the reference tree (build container)."""
import os

import numpy as np
import pytest

from extpom_b200 import synthetic as syn
from scripts import make_ref_golden as mrg
from scripts.make_glue import CUT, cut, restore_records
from tests.fabi import RESTORE, FabiEmu, strips

REF = "/root/reference/pom"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "advance.f")), reason="reference tree not present")
DIMS = (16, 14, 7)


def _files(st, r):
    """Deterministic stand-ins for read_wind / read_heat / read_surface / read_boundary_conditions (record n)."""
    f = st["fields"]
    im, jm, kb = DIMS

    def wind(n, wu, wv):
        wu[...] = f["wusurf"] * (1.0 + 0.1 * n); wv[...] = f["wvsurf"] * (1.0 - 0.05 * n)

    def heat(n, shf, swr):
        shf[...] = f["wtsurf"] * (1.0 + 0.2 * n); swr[...] = f["swrad"] * (1.0 + 0.1 * n)

    def surface(n, sst, sss):
        sst[...] = f["tsurf"] + 0.01 * n; sss[...] = f["ssurf"]

    def bc(n, kb_, tbw, sbw, ubw, vbw, tbe, sbe, ube, vbe, tbn, sbn, vbn, ubn, tbs, sbs, vbs, ubs, elw, ele, eln, els):
        s = 1.0 + 0.02 * n
        for dst, src in ((tbw, "tbw"), (sbw, "sbw"), (tbe, "tbe"), (sbe, "sbe"), (tbn, "tbn"), (sbn, "sbn"), (tbs, "tbs"), (sbs, "sbs")):
            dst[...] = f[src] * (s if src[0] == "t" else 1.0)
        ube[...] = f["ube"] * s; ubw[...] = f["ubw"] * s; vbn[...] = f["vbn"] * s; vbs[...] = f["vbs"] * s
        vbw[...] = 0.; vbe[...] = 0.; ubn[...] = 0.; ubs[...] = 0.
        elw[...] = f["elw"] * s; ele[...] = f["ele"] * s; eln[...] = f["eln"] * s; els[...] = f["els"] * s

    r.ref.externals.update(read_wind_pnetcdf=wind, read_heat_pnetcdf=heat, read_surface_pnetcdf=surface,
                           read_boundary_conditions_pnetcdf=bc)


def _reference(kw):
    from oracle.f77ref import F77Ref
    st, r = mrg.loaded(F77Ref, DIMS, kw)
    _files(st, r)
    r.v["iprint"] = 2; r.v["irestart"] = 10 ** 6; r.v["iend"] = 10 ** 6     # print_section / domain_stats every 2nd step
    r.v["netcdf_file"] = "nonetcdf"; r.v["iswtch"] = 10 ** 6          # get_time keeps iprint (advance.f:67)
    return st, r


def _run_reference(kw, steps):
    st, r = _reference(kw)
    for i in range(1, steps + 1):
        r.v["iint"] = i
        r.ref.call("advance")
    return r


def _run_glue_over_library(kw, steps, tmp_path, factory):
    """B: the generated glue executed, the four step routines and the restore callback in the library."""
    from oracle import f77ref
    text, removed = cut(open(os.path.join(REF, "advance.f")).read())
    assert sorted(removed) == sorted(CUT)
    glue = tmp_path / "advance_glue.f"
    glue.write_text(text + "\n" + restore_records(open(os.path.join(REF, "bounds_forcing.f")).read()))
    st, r = _reference(kw)
    for n in CUT:                                   # the GPU build does not compile these bodies
        del r.ref.units[n]
    units = f77ref.split_units(str(glue))
    assert "advance" in units and "restore_interior_records" in units and not set(CUT) & set(units)
    r.ref.units.update(units)

    lib = factory(*DIMS)                            # COMMON blocks + libpomgpu_f
    names = [n for n, a in r.v.items() if isinstance(a, np.ndarray) and a.dtype == np.float64 and lib._view(n, a.shape) is not None]
    views = {n: lib._view(n, r.v[n].shape) for n in names}
    scal = [n for n, a in r.v.items() if not isinstance(a, np.ndarray) and lib._member(n)[0]
            and isinstance(a, (int, float, np.floating, np.integer)) and not isinstance(a, bool)]

    def to_common():
        for n in scal:
            lib.set(n, r.v[n])
        for n in names:
            views[n][...] = r.v[n]

    def from_common():
        for n in names:
            r.v[n][...] = views[n]
        r.v["error_status"] = int(lib.getc("error_status"))

    def bound(sym):
        def call():
            to_common()
            getattr(lib.L, sym + "_")()
            from_common()
        return call

    def records(_):                                 # the Fortran record half, executed, on the COMMON arrays
        for n in RESTORE:
            r.v[n][...] = views[n]
        r.ref.call("restore_interior_records")
        for n in RESTORE:
            views[n][...] = r.v[n]

    r.ref.externals.update({n: bound(n) for n in CUT})
    lib.set_records(records)
    try:
        for i in range(1, steps + 1):
            r.v["iint"] = i
            r.ref.call("advance")
        lib.L.pomgpu_f_pull_all_()                  # an output step (advance.f:35-49)
        from_common()
    finally:
        lib.set_records(None)
        lib.set_restore(0)
    return r


@pytest.mark.parametrize("factory", [FabiEmu, strips(FabiEmu, 2, ghost=2)], ids=["one_device", "two_strips"])
@pytest.mark.parametrize("kw", [{"obc": True, "fluxes": True, "island": True},
                                {"walls": False, "obc": True, "fluxes": True, "nbct": 2, "npg": 2}], ids=["channel", "open"])
def test_the_references_own_advance_over_the_drop_in_library(kw, factory, tmp_path):
    steps = 4
    a = _run_reference(kw, steps)
    b = _run_glue_over_library(kw, steps, tmp_path, factory)
    assert float(a.v["time"]) == float(b.v["time"]) and int(b.v["error_status"]) == 0
    assert np.abs(a.v["wusurf"]).max() > 0 and not np.array_equal(a.v["wusurf"], a.v["wusurff"])   # the glue interpolated
    for n in list(mrg.F3) + list(mrg.F2) + ["wusurf", "wtsurf", "swrad", "tsurf", "tbe", "ube", "uabe", "vabn", "ele"]:
        if n in ("uf", "vf"):
            continue
        x, y = a.v[n], b.v[n]
        if n in ("t", "tb", "s", "sb"):
            x, y = x[:, :, :-1], y[:, :, :-1]
        e = np.abs(x - y).max() / (np.abs(x).max() + 1e-300)
        assert e <= 1e-11, (n, e)


# ---- initialize.f: read_grid, initial_conditions, update_initial, bottom_friction executed from the source -------------
RAW = ("z zz dx dy h fsm dum dvm").split()                       # what read_grid_pnetcdf delivers (io_pnetcdf.F)
GIVEN = ("ub vb uab vab elb etb e_atmos vfluxb vfluxf wusurf wvsurf wtsurf wssurf swrad ele elw eln els vabe vabw vabn vabs "
         "uabn uabs uabe uabw ube ubw vbn vbs").split()           # the state the synthetic case starts from (no file in the reference)


def _initialize(kw, library=None):
    """initialize.f:19-37 without read_input (a namelist): initialize_arrays, read_grid, initial_conditions,
    update_initial, bottom_friction, executed from the reference source; the PnetCDF readers are played by the
    synthetic generator's raw inputs.  library = a tests/fabi.py driver: `dens`, `baropg`, `baropg_mcc` are then
    the symbols of libpomgpu_f (the GPU build drops solver.o), called with the reference's own argument lists
    -- dens(sclim,tclim,rmean), dens(sb,tb,rho) (initialize.f:416,425) -- by address."""
    from oracle import f77ref
    st = syn.make_state(*DIMS, **{k: v for k, v in kw.items() if k != "_set"})
    f, c = st["fields"], st["consts"]
    r = f77ref.F77Ref(*DIMS)
    r.ref.units.update(f77ref.split_units(os.path.join(REF, "initialize.f")))
    for k, v in c.items():                                       # read_input (initialize.f:67-191)
        if k in r.v and not isinstance(r.v[k], np.ndarray):
            r.set(k, v)
    r.v["pi"] = np.float64(np.arctan(np.float64(1.)) * 4.)      # initialize.f:179

    def read_grid():
        for n in RAW:
            r.v[n][...] = f[n]
        r.v["north_e"][...] = 43.3                               # cor is recomputed from the latitude (initialize.f:349)

    def read_ic(kb, tb, sb):
        tb[...] = f["tb"]; sb[...] = f["sb"]

    def read_clim(kb, n, tclim, sclim):
        tclim[...] = f["tclim"]; sclim[...] = f["sclim"]

    r.ref.externals.update(read_grid_pnetcdf=read_grid, check_cflmin_mpi=lambda: None,
                           read_initial_ts_pnetcdf=read_ic, read_clim_ts_pnetcdf=read_clim)
    if library is not None:
        lib = library
        names = [n for n, a in r.v.items() if isinstance(a, np.ndarray) and a.dtype == np.float64 and lib._view(n, a.shape) is not None]
        views = {n: lib._view(n, r.v[n].shape) for n in names}
        scal = [n for n, a in r.v.items() if not isinstance(a, np.ndarray) and lib._member(n)[0]
                and isinstance(a, (int, float, np.floating, np.integer)) and not isinstance(a, bool)]

        def to_common():
            for n in scal:
                lib.set(n, r.v[n])
            for n in names:
                views[n][...] = r.v[n]

        def from_common():
            for n in names:
                r.v[n][...] = views[n]
            r.v["error_status"] = int(lib.getc("error_status"))

        def name_of(a):
            return next(n for n in names if r.v[n] is a)

        def dens(si, ti, rhoo):
            to_common()
            lib.L.dens_(*[lib._addr(name_of(a)) for a in (si, ti, rhoo)])
            from_common()

        def bound(sym):
            def call():
                to_common(); getattr(lib.L, sym + "_")(); from_common()
            return call

        for n in ("dens", "baropg", "baropg_mcc"):
            del r.ref.units[n]
        r.ref.externals.update(dens=dens, baropg=bound("baropg"), baropg_mcc=bound("baropg_mcc"))
    r.ref.call("initialize_arrays")
    for n in GIVEN:
        if n in f and n in r.v:
            r.v[n][...] = f[n]
    for n in ("read_grid", "initial_conditions", "update_initial", "bottom_friction"):
        r.ref.call(n)
    return st, r


DERIVED = ("dz dzz art aru arv d dt rmean rho tsurf ssurf tbe tbw sbe sbw tbn tbs sbn sbs ua va el et etf w l q2b q2lb kh km kq aam "
           "q2 q2l t s u v drhox drhoy drx2d dry2d cbc").split()


@pytest.mark.parametrize("kw", [{"island": True, "fluxes": True}, {"walls": False, "obc": True, "npg": 2}], ids=["channel", "open_npg2"])
def test_the_generators_initialisation_is_the_references(kw):
    """extpom_b200/synthetic.py restates the derived-input formulas of initialize.f (:331-335, 363-384, 416-425,
    437-460, 472-518, 534-541); here the reference's own routines produce them from the same raw
    inputs: every derived array BITWISE (the oracle makes the dens / baropg calls for the generator)."""
    from oracle.pomo import Oracle
    st, r = _initialize(kw)
    st2, o = syn.seamount(*DIMS, Oracle, **kw)
    assert abs(float(r.v["cor"][3, 3]) - 1.0e-4) < 1e-6          # 2*7.29e-5*sin(43.3 deg): the reference's own formula
    for n in DERIVED:
        a = r.v[n]
        b = o.get(n) if n in o.f else st2["fields"][n]
        assert np.array_equal(a, np.asarray(b).reshape(a.shape, order="F")), n


@pytest.mark.parametrize("factory", [FabiEmu, strips(FabiEmu, 2, ghost=2)], ids=["one_device", "two_strips"])
def test_initialize_then_advance_with_solver_f_replaced_by_the_library(factory, tmp_path):
    """The whole program minus its file I/O: initialize.f's routines and then `advance`, executed from the reference
    source, (A) unmodified and (B) as the GPU build links it -- no solver.o, no step routines: dens, baropg and the
    four step routines are libpomgpu_f's."""
    from oracle import f77ref
    kw = {"walls": False, "obc": True, "fluxes": True, "island": True}
    out = []
    for lib in (None, factory(*DIMS)):
        st, r = _initialize(kw, lib)
        _files(st, r)
        r.v["iprint"] = 2; r.v["irestart"] = 10 ** 6; r.v["iend"] = 10 ** 6
        r.v["netcdf_file"] = "nonetcdf"; r.v["iswtch"] = 10 ** 6
        if lib is not None:
            text, _ = cut(open(os.path.join(REF, "advance.f")).read())
            glue = tmp_path / "advance_glue.f"
            glue.write_text(text + "\n" + restore_records(open(os.path.join(REF, "bounds_forcing.f")).read()))
            for n in CUT:
                del r.ref.units[n]
            r.ref.units.update(f77ref.split_units(str(glue)))
            names = [n for n, a in r.v.items() if isinstance(a, np.ndarray) and a.dtype == np.float64 and lib._view(n, a.shape) is not None]
            views = {n: lib._view(n, r.v[n].shape) for n in names}
            scal = [n for n, a in r.v.items() if not isinstance(a, np.ndarray) and lib._member(n)[0]
                    and isinstance(a, (int, float, np.floating, np.integer)) and not isinstance(a, bool)]

            def to_common():
                for n in scal:
                    lib.set(n, r.v[n])
                for n in names:
                    views[n][...] = r.v[n]

            def from_common():
                for n in names:
                    r.v[n][...] = views[n]
                r.v["error_status"] = int(lib.getc("error_status"))

            def bound(sym):
                def call():
                    to_common(); getattr(lib.L, sym + "_")(); from_common()
                return call

            def records(_):
                for n in RESTORE:
                    r.v[n][...] = views[n]
                r.ref.call("restore_interior_records")
                for n in RESTORE:
                    views[n][...] = r.v[n]

            r.ref.externals.update({n: bound(n) for n in CUT})
            lib.set_records(records)
        try:
            for i in range(1, 4):
                r.v["iint"] = i
                r.ref.call("advance")
            if lib is not None:
                lib.L.pomgpu_f_pull_all_()
                from_common()
        finally:
            if lib is not None:
                lib.set_records(None); lib.set_restore(0)
        out.append(r)
    a, b = out
    assert int(b.v["error_status"]) == 0
    for n in list(mrg.F3) + list(mrg.F2) + ["rmean", "cbc", "tsurf", "wusurf"]:
        if n in ("uf", "vf"):
            continue
        x, y = a.v[n], b.v[n]
        if n in ("t", "tb", "s", "sb"):
            x, y = x[:, :, :-1], y[:, :, :-1]
        e = np.abs(x - y).max() / (np.abs(x).max() + 1e-300)
        assert e <= 1e-11, (n, e)
