"""ctypes front end of the CPU parity ORACLE (oracle/libpomo.so).

TEST INFRASTRUCTURE ONLY -- parity pinned against the reference's own source run through
oracle/f77ref.py (see oracle/pomo.h, tests/golden/ref_*.npz).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (extpom_b200) never does.

The class mirrors extpom_b200.PomGpu's Python surface (load / get / step and the
reference's subroutine names) so that parity tests drive both the same way.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = {}

F3D = ("aam advx advy drhox drhoy dtef kh km kq l q2b q2 q2lb q2l rho rmean sb sclim s "
       "tb tclim t ub uf u vb vf v w wr zflux trstr trstrb trstrf srstr srstrb srstrf "
       "taurstr taurstrb taurstrf").split()
F2D = ("aam2d advua advva adx2d ady2d art aru arv cbc cor d drx2d dry2d dt dum dvm dx dy "
       "e_atmos egb egf el elb elf et etb etf fluxua fluxva fsm h swrad ssurf tsurf tps ua "
       "uab uaf utb utf va vab vaf vtb vtf vfluxb vfluxf wssurf wtsurf wubot wusurf wvbot "
       "wvsurf wusurfb wusurff wvsurfb wvsurff wtsurfb wtsurff swradb swradf").split()
BJ = "ele elw uabe uabw vabe vabw".split()
BI = "eln els vabn vabs uabn uabs".split()
BJK = ("tbe sbe tbw sbw ube ubw tbeb tbef sbeb sbef ubeb ubef tbwb tbwf sbwb sbwf ubwb ubwf").split()
BIK = ("tbn sbn tbs sbs vbn vbs tbnb tbnf sbnb sbnf vbnb vbnf tbsb tbsf sbsb sbsf vbsb vbsf").split()
F1D = "z zz dz dzz".split()


def build(force=False, variant=""):
    """Compile oracle/libpomo[_O0|_O3].so with the committed Makefile (gcc, -ffp-contract=off).
    variant "" = -O2 (the checker); "O0" / "O3" = the CPU-baseline variants of SURVEY.md 8(d)."""
    name = "libpomo.so" if not variant else f"libpomo_{variant}.so"
    so = os.path.join(_HERE, name)
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, name])
    return so


def set_threads(n):
    """OpenMP threads of the oracle from now on (bench.py: torchrun exports OMP_NUM_THREADS=1)."""
    os.environ["OMP_NUM_THREADS"] = str(int(n))
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass


def lib(variant=""):
    if variant not in _LIB:
        L = C.CDLL(build(variant=variant))
        L.pomo_create.restype = C.c_void_p
        L.pomo_create.argtypes = [C.c_int] * 3
        L.pomo_destroy.argtypes = [C.c_void_p]
        L.pomo_field.restype = C.POINTER(C.c_double)
        L.pomo_field.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_long)]
        L.pomo_set.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.pomo_get.restype = C.c_double
        L.pomo_get.argtypes = [C.c_void_p, C.c_char_p]
        L.pomo_check_velocity.restype = C.c_double
        L.pomo_domain_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        for n in ("step lateral_viscosity mode_interaction mode_external mode_internal advave "
                  "advct advu advv baropg baropg_mcc profq profu profv vertvl realvertvl "
                  "restore_interior check_velocity").split():
            getattr(L, "pomo_" + n).argtypes = [C.c_void_p]
        P = C.c_void_p
        L.pomo_advq.argtypes = [P] * 4
        L.pomo_advt1.argtypes = [P] * 5
        L.pomo_advt2.argtypes = [P] * 5
        L.pomo_dens.argtypes = [P] * 4
        L.pomo_proft.argtypes = [P, P, P, P, C.c_int]
        L.pomo_smol_adif.argtypes = [P] * 5
        L.pomo_bcond.argtypes = [P, C.c_int]
        L.pomo_bcondorl.argtypes = [P, C.c_int]
        L.pomo_internal_stage.argtypes = [P, C.c_int]
        for n in ("wind_interp", "heat_interp", "lateral_bc_interp"):
            getattr(L, "pomo_" + n).argtypes = [P, C.c_double]
        _LIB[variant] = L
    return _LIB[variant]


class Oracle:
    """One sub-domain of the reference model on the CPU (fp64, no FMA)."""

    def __init__(self, im, jm, kb, variant=""):
        self.L = lib(variant)
        self.im, self.jm, self.kb = im, jm, kb
        self.h = self.L.pomo_create(im, jm, kb)
        self.f = {}
        shapes = {}
        for n in F3D: shapes[n] = (im, jm, kb)
        for n in F2D: shapes[n] = (im, jm)
        for n in BJ: shapes[n] = (jm,)
        for n in BI: shapes[n] = (im,)
        for n in BJK: shapes[n] = (jm, kb)
        for n in BIK: shapes[n] = (im, kb)
        for n in F1D: shapes[n] = (kb,)
        self.shapes = shapes
        for n, shp in shapes.items():
            cnt = C.c_long(0)
            p = self.L.pomo_field(self.h, n.encode(), C.byref(cnt))
            assert cnt.value == int(np.prod(shp)), n
            a = np.ctypeslib.as_array(p, shape=(cnt.value,))
            self.f[n] = a.reshape(shp, order="F")  # column-major view, i fastest

    def close(self):
        if self.h:
            self.f.clear()
            self.L.pomo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state I/O -------------------------------------------------------
    def set(self, name, v):
        if self.L.pomo_set(self.h, name.encode(), float(v)) != 0:
            raise KeyError(name)

    def getc(self, name):
        return self.L.pomo_get(self.h, name.encode())

    def load(self, state):
        """state = {'consts': {...}, 'fields': {name: ndarray}} (see extpom_b200.synthetic)."""
        for k, v in state["consts"].items():
            self.L.pomo_set(self.h, k.encode(), float(v))  # unknown names ignored
        for k, v in state["fields"].items():
            if k in self.f:
                self.f[k][...] = v

    def get(self, name):
        return np.array(self.f[name], order="F", copy=True)

    def put(self, name, arr):
        self.f[name][...] = arr

    def _p(self, x):
        if isinstance(x, str):
            x = self.f[x]
        return x.ctypes.data_as(C.c_void_p)

    # -- the reference's subroutine surface --------------------------------
    def step(self, iint, time=None):
        """advance.f:21-32 for internal step `iint` (get_time restated: advance.f:62-75)."""
        self.set("iint", iint)
        dti, time0 = self.getc("dti"), self.getc("time0")
        self.set("time", dti * float(iint) / 86400.0 + time0 if time is None else time)
        self.L.pomo_step(self.h)

    def check_velocity(self):
        return self.L.pomo_check_velocity(self.h)

    def domain_stats(self):
        """advance.f:644-755 -> dict(vtot, atot, mtot, stot, tavg, savg, eavg, ekin)."""
        out = (C.c_double * 8)()
        self.L.pomo_domain_stats(self.h, out)
        return dict(zip("vtot atot mtot stot tavg savg eavg ekin".split(), list(out)))

    # forcing records (bounds_forcing.f:841-865,904-909,949-957): same surface as PomGpu
    def put_record(self, name, slot, arr):
        self.put(name + "bf"[slot], arr)

    def rotate_record(self, name):
        self.put(name + "b", self.get(name + "f"))   # `wusurfb = wusurff`

    def wind(self, fnew): self.L.pomo_wind_interp(self.h, fnew)
    def heat(self, fnew): self.L.pomo_heat_interp(self.h, fnew)
    def lateral_bc(self, fnew): self.L.pomo_lateral_bc_interp(self.h, fnew)
    def baropg_mcc(self): self.L.pomo_baropg_mcc(self.h)
    def lateral_viscosity(self): self.L.pomo_lateral_viscosity(self.h)
    def mode_interaction(self): self.L.pomo_mode_interaction(self.h)
    def mode_external(self, iext):
        self.set("iext", iext); self.L.pomo_mode_external(self.h)
    def mode_internal(self, iint):
        self.set("iint", iint); self.L.pomo_mode_internal(self.h)
    def internal_stage(self, iint, stage):
        self.set("iint", iint); self.L.pomo_internal_stage(self.h, stage)
    def advave(self): self.L.pomo_advave(self.h)
    def advct(self): self.L.pomo_advct(self.h)
    def advu(self): self.L.pomo_advu(self.h)
    def advv(self): self.L.pomo_advv(self.h)
    def baropg(self): self.L.pomo_baropg(self.h)
    def profq(self): self.L.pomo_profq(self.h)
    def profu(self): self.L.pomo_profu(self.h)
    def profv(self): self.L.pomo_profv(self.h)
    def vertvl(self): self.L.pomo_vertvl(self.h)
    def realvertvl(self): self.L.pomo_realvertvl(self.h)
    def restore_interior(self): self.L.pomo_restore_interior(self.h)
    def bcond(self, idx): self.L.pomo_bcond(self.h, idx)
    def bcondorl(self, idx): self.L.pomo_bcondorl(self.h, idx)
    def advq(self, qb, q, qf): self.L.pomo_advq(self.h, self._p(qb), self._p(q), self._p(qf))
    def advt1(self, fb, f, fclim, ff):
        self.L.pomo_advt1(self.h, self._p(fb), self._p(f), self._p(fclim), self._p(ff))
    def advt2(self, fb, f, fclim, ff):
        self.L.pomo_advt2(self.h, self._p(fb), self._p(f), self._p(fclim), self._p(ff))
    def dens(self, si, ti, rhoo): self.L.pomo_dens(self.h, self._p(si), self._p(ti), self._p(rhoo))
    def proft(self, f, wfsurf, fsurf, nbc):
        self.L.pomo_proft(self.h, self._p(f), self._p(wfsurf), self._p(fsurf), int(nbc))
