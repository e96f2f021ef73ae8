"""The reference's Fortran driver, played from Python against libpomgpu_f.so -- TEST INFRASTRUCTURE.

`FabiPom` has the interface of the solvers the parity cases drive (PomGpu, EmuPom, Oracle), but everything goes
through what a gfortran-built extPOM binds: the COMMON blocks `blksiz_ blkpar_ blkcon_ blk1d_ blk2d_ blk3d_ bdry_`
(defined by tests/c/common_blocks.c, loaded RTLD_GLOBAL like an executable's own) and the mangled entry points
`lateral_viscosity_ ... realvertvl_`, `advq_(qb,q,qf)`, `dens_(si,ti,rhoo)`, `bcond_(idx)` ... of include/pomgpu_f.h.
State is written into / read from COMMON memory exactly as the Fortran would; array arguments are passed as the
addresses of COMMON members (or of local arrays, for names that are no COMMON member), integers by reference.
`step()` spells advance.f:21-32.  One instance at a time (the COMMON blocks are process globals, as in the model).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from extpom_b200 import pomgpu as _pg

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_EMU_DIR = os.path.join(_ROOT, "tests", "_emu")
COMMON_SO = os.path.join(_EMU_DIR, "libpom_common.so")
FLIB_EMU = os.path.join(_EMU_DIR, "libpomgpu_f_emu.so")
FLIB = os.path.join(_ROOT, "extpom_b200", "libpomgpu_f.so")
RESTORE = ("trstrb", "trstrf", "srstrb", "srstrf", "taurstrb", "taurstrf")
_HOOK = C.CFUNCTYPE(None)
_common = None
_flibs = {}


def _load_common():
    """The COMMON blocks + the slot a Python function can be hung into as `restore_interior_records_`."""
    global _common
    if _common is None:
        os.makedirs(_EMU_DIR, exist_ok=True)
        src = os.path.join(_ROOT, "tests", "c", "common_blocks.c")
        if not os.path.exists(COMMON_SO) or os.path.getmtime(COMMON_SO) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", src, "-o", COMMON_SO])
        _common = C.CDLL(COMMON_SO, mode=C.RTLD_GLOBAL)
        for f in ("pom_common_max_n2", "pom_common_max_n3", "pom_common_max_edge", "pom_common_bytes"):
            getattr(_common, f).restype = C.c_long
    return _common


def _load_flib(path):
    if path not in _flibs:
        _load_common()                      # the blocks must be visible before the library's weak references bind
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        L.pomgpu_f_member.restype = C.c_void_p
        L.pomgpu_f_member.argtypes = [C.c_char_p, C.POINTER(C.c_long)]
        L.pomgpu_f_member_type.restype = C.c_char
        L.pomgpu_f_member_type.argtypes = [C.c_char_p]
        L.pomgpu_f_last_error.restype = C.c_char_p
        _flibs[path] = L
    return _flibs[path]


class FabiError(RuntimeError):
    pass


class FabiPom:
    """See the module docstring.  Subclasses choose the library: host emulation (CPU tests) or CUDA."""
    FLIB = None
    NDEV = None              # pomgpu_f_set_devices_: n strips on n devices; -n: n strips on one device
    GHOST = None             # POMGPU_F_GHOST
    _live = None

    def __init__(self, im, jm, kb):
        cm = _load_common()
        assert im * jm <= cm.pom_common_max_n2() and im * jm * kb <= cm.pom_common_max_n3() \
            and max(im, jm) * kb <= cm.pom_common_max_edge(), "grid larger than tests/c/common_blocks.c provides"
        self.L = _load_flib(self._flib())
        if FabiPom._live is not None:
            FabiPom._live.close()
        FabiPom._live = self
        self.im, self.jm, self.kb = im, jm, kb
        self.L.pomgpu_f_finalize_()
        self.L.pomgpu_f_set_dims_(C.byref(C.c_int(im)), C.byref(C.c_int(jm)), C.byref(C.c_int(kb)))
        if self.NDEV:
            self.L.pomgpu_f_set_devices_(C.byref(C.c_int(self.NDEV)))
        if self.GHOST:
            os.environ["POMGPU_F_GHOST"] = str(self.GHOST)
        else:
            os.environ.pop("POMGPU_F_GHOST", None)
        for n, blk in enumerate(("blksiz_", "blkpar_", "blkcon_", "blk1d_", "blk2d_", "blk3d_", "bdry_")):
            used = min(cm.pom_common_bytes(n), {4: 73 * im * jm * 8, 5: 40 * im * jm * kb * 8}.get(n, 1 << 62))
            C.memset(C.addressof(C.c_char.in_dll(cm, blk)), 0, used)
        # distribute_mpi on one rank (parallel_mpi.f:76-119)
        for n, v in (("im", im), ("imm1", im - 1), ("imm2", im - 2), ("jm", jm), ("jmm1", jm - 1), ("jmm2", jm - 2),
                     ("kbm1", kb - 1), ("kbm2", kb - 2), ("n_west", -1), ("n_east", -1), ("n_south", -1), ("n_north", -1),
                     ("iprint", 1000000), ("irestart", 1000000)):
            self.set(n, v)
        self.local = {}          # arrays that are no COMMON member (a Fortran caller's local arrays)
        self.stepped = False     # step-level calls ran since the host copies were last refreshed
        self.records = None      # Python stand-in for `subroutine restore_interior_records` (hung into COMMON_SO's slot)
        self._hook = None
        self.set_restore(0)      # the parity cases switch the nudging on explicitly (set("lrestore", 1))
        jl = jm
        self.shapes = {}
        for n in _pg.F3D + _pg.F3D_OPT + _pg.F3D_SCR: self.shapes[n] = (im, jl, kb)
        for n in _pg.F2D: self.shapes[n] = (im, jl)
        for n in _pg.BJ: self.shapes[n] = (jl,)
        for n in _pg.BI: self.shapes[n] = (im,)
        for n in _pg.BJK: self.shapes[n] = (jl, kb)
        for n in _pg.BIK: self.shapes[n] = (im, kb)
        for n in _pg.F1D: self.shapes[n] = (kb,)

    @classmethod
    def _flib(cls):
        return cls.FLIB

    def close(self):
        if FabiPom._live is self:
            self.L.pomgpu_f_finalize_()
            FabiPom._live = None

    # -- COMMON memory ---------------------------------------------------------------------------
    def _member(self, name):
        n = C.c_long(0)
        p = self.L.pomgpu_f_member(name.encode(), C.byref(n))
        return p, n.value

    def _view(self, name, shape=None):
        """numpy view of the COMMON member `name` (None if it is not one, or not of the given shape)."""
        p, n = self._member(name)
        if not p or self.L.pomgpu_f_member_type(name.encode()) != b"d" or n < 2:
            return None
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n,))
        shp = shape or self.shapes.get(name)
        if shape is not None and int(np.prod(shape)) != n:
            return None
        if shp is None:
            shp = (self.im, self.jm, self.kb) if n == self.im * self.jm * self.kb else (self.im, self.jm)
        assert int(np.prod(shp)) == n, (name, shp, n)
        return a.reshape(shp, order="F")

    def _addr(self, name):
        """What a Fortran caller passes for the array `name`: its address."""
        v = self._view(name)
        if v is None:
            if name not in self.local:
                self.local[name] = np.zeros(self.shapes.get(name, (self.im, self.jm, self.kb)), order="F")
            v = self.local[name]
        return C.c_void_p(v.ctypes.data)

    def set(self, name, v):
        if name == "lrestore":
            return self.set_restore(int(v))
        p, _ = self._member(name)
        if not p:
            return               # not a COMMON member (pow_mode ...): nothing a Fortran driver could set
        if self.L.pomgpu_f_member_type(name.encode()) == b"d":
            C.c_double.from_address(p).value = float(v)
        else:
            C.c_int.from_address(p).value = int(v)

    def getc(self, name):
        p, _ = self._member(name)
        if not p:
            raise FabiError(f"{name} is not a COMMON member")
        if self.L.pomgpu_f_member_type(name.encode()) == b"d":
            return C.c_double.from_address(p).value
        return float(C.c_int.from_address(p).value)

    def set_restore(self, on):
        self.L.pomgpu_f_set_restore_(C.byref(C.c_int(on)))

    def set_records(self, fn):
        """fn(self): what the driver's `restore_interior_records` does (bounds_forcing.f:1023-1081); None = no such
        routine in the executable.  Also returns the switch to the library's automatic rule."""
        cm = _load_common()
        self._hook = _HOOK((lambda: fn(self)) if fn else 0)
        C.c_void_p.in_dll(cm, "pom_records_hook").value = C.cast(self._hook, C.c_void_p).value if fn else None
        self.set_restore(-1)

    def _refresh_host(self):
        if self.stepped:
            self.L.pomgpu_f_pull_all_()          # what an output / restart step does (advance.f:35-49)
            self.stepped = False
            self._status("pull_all")

    def _status(self, what):
        if int(self.getc("error_status")) != 0:
            raise FabiError(f"{what}: error_status=1: {self.L.pomgpu_f_last_error().decode()}")

    def load(self, state):
        for k, v in state["consts"].items():
            self.set(k, v)
        for k, v in state["fields"].items():
            if k in _pg.F3D_OPT or k in _pg.F3D_SCR:
                continue
            dst = self._view(k)
            if dst is not None and dst.size == np.asarray(v).size:
                dst[...] = np.asarray(v, dtype=np.float64).reshape(dst.shape, order="F")

    def put(self, name, arr):
        """The driver changes a COMMON array on the host (and tells the library: pomgpu_f_push_)."""
        self._refresh_host()
        dst = self._view(name)
        a = np.asarray(arr, dtype=np.float64)
        if dst is None:
            self.local[name] = np.asfortranarray(a).copy(order="F")
            return
        dst[...] = a.reshape(dst.shape, order="F")
        self.L.pomgpu_f_push_(C.c_void_p(dst.ctypes.data))
        self._status(f"push({name})")

    def get(self, name):
        self._refresh_host()
        v = self._view(name)
        return np.array(v if v is not None else self.local[name], order="F", copy=True)

    # -- advance.f:21-32 ---------------------------------------------------------------------------
    def step(self, iint, time=None):
        self.set("iint", iint)
        self.set("time", self.getc("dti") * float(iint) / 86400.0 + self.getc("time0") if time is None else time)  # get_time
        self.L.lateral_viscosity_()
        self.L.mode_interaction_()
        isplit = int(self.getc("isplit"))
        for iext in range(1, isplit + 1):
            self.set("iext", iext)
            self.L.mode_external_()
        self.set("iext", isplit + 1)                 # a Fortran do loop leaves its variable one past the end
        self.L.mode_internal_()
        self.stepped = True
        self._status(f"step {iint}")

    def check_velocity(self):
        """advance.f:611-641 on the host copy of vaf, which mode_internal_ refreshes every step."""
        return float(np.abs(self._view("vaf")).max())

    # -- routine level (unit mode) -----------------------------------------------------------------
    def _call0(self, name):
        getattr(self.L, name + "_")()
        self.stepped = False       # a unit-mode call pulls the state first if the device was ahead
        self._status(name)

    def advct(self): self._call0("advct")
    def advave(self): self._call0("advave")
    def advu(self): self._call0("advu")
    def advv(self): self._call0("advv")
    def baropg(self): self._call0("baropg")
    def baropg_mcc(self): self._call0("baropg_mcc")
    def profq(self): self._call0("profq")
    def profu(self): self._call0("profu")
    def profv(self): self._call0("profv")
    def realvertvl(self): self._call0("realvertvl")

    def vertvl(self):
        """advance.f:396-398: vertvl, then bcondorl(5)."""
        self._call0("vertvl")
        self.bcondorl(5)

    def _calln(self, name, *args):
        getattr(self.L, name + "_")(*args)
        self.stepped = False
        self._status(name)

    def dens(self, si, ti, rhoo): self._calln("dens", self._addr(si), self._addr(ti), self._addr(rhoo))
    def advq_fields(self, qb, q, qf): self._calln("advq", self._addr(qb), self._addr(q), self._addr(qf))

    def advq(self):
        """advance.f:408-409: advq(q2b,q2,uf), advq(q2lb,q2l,vf)."""
        self.advq_fields("q2b", "q2", "uf")
        self.advq_fields("q2lb", "q2l", "vf")

    def advt1(self, fb, f, fclim, ff): self._calln("advt1", *[self._addr(n) for n in (fb, f, fclim, ff)])
    def advt2(self, fb, f, fclim, ff): self._calln("advt2", *[self._addr(n) for n in (fb, f, fclim, ff)])
    def smol_adif(self, x, y, z, ff): self._calln("smol_adif", *[self._addr(n) for n in (x, y, z, ff)])

    def proft(self, f, wfsurf, fsurf, nbc):
        self._calln("proft", self._addr(f), self._addr(wfsurf), self._addr(fsurf), C.byref(C.c_int(nbc)))

    def bcond(self, idx): self._calln("bcond", C.byref(C.c_int(idx)))
    def bcondorl(self, idx): self._calln("bcondorl", C.byref(C.c_int(idx)))


class FabiEmu(FabiPom):
    """Through the gfortran ABI into the host-emulated kernel bodies (CPU tests)."""
    @classmethod
    def _flib(cls):
        from tests import emu
        emu.build_emu()
        return FLIB_EMU


def strips(base, n, ghost=None):
    """`base` with the domain spread over n strips inside the library (pomgpu_f_set_devices_)."""
    return type(f"{base.__name__}x{abs(n)}", (base,), {"NDEV": n, "GHOST": ghost})


class FabiGpu(FabiPom):
    """Through the gfortran ABI into the CUDA library (`-m gpu`)."""
    FLIB = FLIB
