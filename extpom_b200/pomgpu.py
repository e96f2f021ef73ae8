"""Host-side mirror of the reference's time-stepping interface over libpomgpu's C ABI.

The reference (RinceWND/extPOM) is a Fortran program whose hot path is a set of
argument-less external subroutines working on COMMON blocks (pom/advance.f:21-32,
pom/solver.f).  `PomGpu` exposes the same subroutine names and the same array layout
(fp64, column-major, i fastest) over include/pomgpu.h; fields are addressed by their
COMMON member names.  All compute happens in the hand-written CUDA kernels of
extpom_b200/csrc (sm_100a); there is no CPU fallback: constructing a PomGpu without
the built library or without a CUDA device raises.
"""
import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(_HERE, "libpomgpu.so")

F3D = ("aam advx advy drhox drhoy kh km kq l q2b q2 q2lb q2l rho rmean sb sclim s tb tclim t "
       "ub uf u vb vf v w wr").split()
F3D_OPT = "trstrb trstrf srstrb srstrf taurstrb taurstrf".split()
F3D_SCR = "rho2 s3a s3b s3c s3d s3e".split()   # library-owned scratch, addressable for unit-mode calls (smol_adif's flux arrays)
F2D = ("aam2d advua advva adx2d ady2d art aru arv cbc cor d drx2d dry2d dt dum dvm dx dy "
       "e_atmos egb egf el elb elf et etb etf fsm h swrad ssurf tsurf ua uab uaf utb utf va "
       "vab vaf vtb vtf vfluxb vfluxf wssurf wtsurf wubot wusurf wvbot wvsurf").split()
BJ = "ele elw uabe uabw vabe vabw".split()
BI = "eln els vabn vabs uabn uabs".split()
BJK = "tbe sbe tbw sbw ube ubw".split()
BIK = "tbn sbn tbs sbs vbn vbs".split()
F1D = "z zz dz dzz".split()


class PomGpuError(RuntimeError):
    pass


def _bind(path):
    if not os.path.exists(path):
        raise PomGpuError(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(libpomgpu has no CPU fallback)")
    L = C.CDLL(path)
    P = C.c_void_p
    L.pomgpu_create.restype = P
    L.pomgpu_create.argtypes = [C.c_int] * 4
    L.pomgpu_create_strip.restype = P
    L.pomgpu_create_strip.argtypes = [C.c_int] * 7
    L.pomgpu_destroy.argtypes = [P]
    L.pomgpu_local_rows.argtypes = [P]
    L.pomgpu_row_offset.argtypes = [P]
    L.pomgpu_last_error.restype = C.c_char_p
    L.pomgpu_last_error.argtypes = [P]
    L.pomgpu_set_const.argtypes = [P, C.c_char_p, C.c_double]
    L.pomgpu_get_const.argtypes = [P, C.c_char_p, C.POINTER(C.c_double)]
    L.pomgpu_push.argtypes = [P, C.c_char_p, P]
    L.pomgpu_pull.argtypes = [P, C.c_char_p, P]
    L.pomgpu_push_rows.argtypes = [P, C.c_char_p, P, C.c_int, C.c_int]
    L.pomgpu_field_elems.restype = C.c_long
    L.pomgpu_field_elems.argtypes = [P, C.c_char_p]
    L.pomgpu_step.argtypes = [P, C.c_int, C.c_double, C.c_double]
    L.pomgpu_sync.argtypes = [P]
    L.pomgpu_check_velocity.restype = C.c_double
    L.pomgpu_check_velocity.argtypes = [P]
    L.pomgpu_push_async.argtypes = [P, C.c_char_p, P]
    L.pomgpu_domain_stats_rows.argtypes = [P, P]
    L.pomgpu_selftest_pdiv.restype = C.c_long
    L.pomgpu_selftest_pdiv.argtypes = [P, C.c_long, C.c_ulong, C.c_int]
    L.pomgpu_push_record.argtypes = [P, C.c_char_p, C.c_int, P]
    L.pomgpu_rotate_record.argtypes = [P, C.c_char_p]
    L.pomgpu_interp.argtypes = [P, C.c_char_p, C.c_double]
    for n in ("wind", "heat", "lateral_bc"):
        getattr(L, "pomgpu_" + n).argtypes = [P, C.c_double]
    L.pomgpu_check_velocity_lagged.restype = C.c_double
    L.pomgpu_check_velocity_lagged.argtypes = [P]
    L.pomgpu_field_absmax_lagged.restype = C.c_double
    L.pomgpu_field_absmax_lagged.argtypes = [P, C.c_char_p]
    L.pomgpu_pin_host.argtypes = [P, C.c_ulong]
    L.pomgpu_unpin_host.argtypes = [P]
    L.pomgpu_event_record.argtypes = [P, C.c_int]
    L.pomgpu_event_elapsed_ms.restype = C.c_double
    L.pomgpu_event_elapsed_ms.argtypes = [P, C.c_int, C.c_int]
    L.pomgpu_profile_begin.argtypes = [P]
    L.pomgpu_profile_end.argtypes = [P, C.c_char_p, C.c_int]
    L.pomgpu_launch_count.restype = C.c_long
    L.pomgpu_launch_count.argtypes = [P, C.c_int]
    for n in ("lateral_viscosity mode_interaction advave advct advq advu advv baropg baropg_mcc profq profu "
              "profv vertvl realvertvl").split():
        getattr(L, "pomgpu_" + n).argtypes = [P]
    L.pomgpu_mode_external.argtypes = [P, C.c_int]
    L.pomgpu_mode_internal.argtypes = [P, C.c_int]
    L.pomgpu_internal_stage.argtypes = [P, C.c_int, C.c_int]
    L.pomgpu_group_create.restype = P
    L.pomgpu_group_create.argtypes = [C.c_int, C.POINTER(P)]
    L.pomgpu_group_destroy.argtypes = [P]
    L.pomgpu_nccl_unique_id.argtypes = [P]
    L.pomgpu_group_connect_nccl.argtypes = [P, P, C.c_int, C.c_int]
    L.pomgpu_group_set_transport.argtypes = [P, P, P]
    L.pomgpu_group_step.argtypes = [P, C.c_int, C.c_double, C.c_double]
    L.pomgpu_group_dens.argtypes = [P] + [C.c_char_p] * 3
    L.pomgpu_group_baropg.argtypes = [P]
    L.pomgpu_group_check_velocity.restype = C.c_double
    L.pomgpu_group_check_velocity.argtypes = [P]
    L.pomgpu_group_transport.restype = C.c_int
    L.pomgpu_group_transport.argtypes = [P]
    L.pomgpu_group_exchanges.restype = C.c_long
    L.pomgpu_group_exchanges.argtypes = [P, C.POINTER(C.c_long), C.c_int]
    L.pomgpu_advt1.argtypes = [P] + [C.c_char_p] * 4
    L.pomgpu_advt2.argtypes = [P] + [C.c_char_p] * 4
    L.pomgpu_dens.argtypes = [P] + [C.c_char_p] * 3
    L.pomgpu_proft.argtypes = [P] + [C.c_char_p] * 3 + [C.c_int]
    L.pomgpu_advq_fields.argtypes = [P] + [C.c_char_p] * 3
    L.pomgpu_smol_adif.argtypes = [P] + [C.c_char_p] * 4
    L.pomgpu_bcond.argtypes = [P, C.c_int]
    L.pomgpu_bcondorl.argtypes = [P, C.c_int]
    return L


_LIBS = {}


def _lib(path):
    if path not in _LIBS:
        _LIBS[path] = _bind(path)
    return _LIBS[path]


class PomGpu:
    """One j-strip (default: the whole domain) of the model, resident in HBM on one B200."""

    def __init__(self, im, jm, kb, device=0, strip=None, ghost=0):
        self.L = self._library()
        self.im, self.jm, self.kb = im, jm, kb
        if strip is None:
            self.h = self.L.pomgpu_create(im, jm, kb, device)
        else:
            self.h = self.L.pomgpu_create_strip(im, jm, kb, strip[0], strip[1], ghost, device)
        if not self.h:
            raise PomGpuError("pomgpu_create failed (no CUDA device, bad extents or out of memory); "
                              "there is no CPU fallback")
        self.own = (1, jm) if strip is None else tuple(strip)
        self.jml = self.L.pomgpu_local_rows(self.h)
        self.joff = self.L.pomgpu_row_offset(self.h)
        jl = self.jml
        self.shapes = {}
        for n in F3D + F3D_OPT + F3D_SCR: self.shapes[n] = (im, jl, kb)
        for n in F2D: self.shapes[n] = (im, jl)
        for n in BJ: self.shapes[n] = (jl,)
        for n in BI: self.shapes[n] = (im,)
        for n in BJK: self.shapes[n] = (jl, kb)
        for n in BIK: self.shapes[n] = (im, kb)
        for n in F1D: self.shapes[n] = (kb,)

    @staticmethod
    def _library():
        """The CUDA library, and nothing else: the product class has no way to select another one
        (the host-emulated build used by the CPU tests is bound by a subclass in tests/emu.py)."""
        return _lib(LIBPATH)

    def close(self):
        if getattr(self, "h", None):
            self.L.pomgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise PomGpuError(f"{what} failed (rc={rc}): {self.L.pomgpu_last_error(self.h).decode()}")

    # -- state I/O -----------------------------------------------------------
    def set(self, name, v):
        self._ck(self.L.pomgpu_set_const(self.h, name.encode(), float(v)), f"set_const({name})")

    def getc(self, name):
        v = C.c_double(0)
        self._ck(self.L.pomgpu_get_const(self.h, name.encode(), C.byref(v)), f"get_const({name})")
        return v.value

    def _rows(self, name, arr):
        """Slice a global-sized host array down to this strip's rows."""
        shp = self.shapes[name]
        a = np.asarray(arr, dtype=np.float64)
        if a.shape == shp:
            return a
        j0, j1 = self.joff, self.joff + self.jml
        if name in F2D or name in F3D or name in F3D_OPT or name in F3D_SCR:
            return a[:, j0:j1]
        if name in BJ or name in BJK:
            return a[j0:j1]
        raise ValueError(f"{name}: shape {a.shape} != {shp}")

    def put(self, name, arr):
        a = np.asfortranarray(self._rows(name, arr), dtype=np.float64)
        assert a.shape == self.shapes[name], (name, a.shape, self.shapes[name])
        self._ck(self.L.pomgpu_push(self.h, name.encode(), a.ctypes.data_as(C.c_void_p)), f"push({name})")

    def put_rows(self, name, arr, row0):
        """Local rows row0.. of field `name` from an array holding just those rows."""
        a = np.asfortranarray(arr, dtype=np.float64)
        jdim = name in F2D or name in F3D or name in F3D_OPT
        nrows = a.shape[1] if jdim else (a.shape[0] if (name in BJ or name in BJK) else 0)
        self._ck(self.L.pomgpu_push_rows(self.h, name.encode(), a.ctypes.data_as(C.c_void_p), int(row0), int(max(nrows, 1))),
                 f"push_rows({name})")

    def load_rows(self, state, row0):
        """Like load() for a state generated for a band of this strip's rows (synthetic.make_state(rows=...))."""
        for k, v in state["consts"].items():
            self.L.pomgpu_set_const(self.h, k.encode(), float(v))
        for k, v in state["fields"].items():
            if k in self.shapes and (k not in F3D_OPT) and (k not in F3D_SCR):
                self.put_rows(k, v, row0)

    def get(self, name):
        out = np.empty(self.shapes[name], dtype=np.float64, order="F")
        self._ck(self.L.pomgpu_pull(self.h, name.encode(), out.ctypes.data_as(C.c_void_p)), f"pull({name})")
        return out

    def pinned(self, name):
        """A page-locked host array shaped like field `name` (per-step forcing buffers)."""
        a = np.zeros(self.shapes[name], dtype=np.float64, order="F")
        ptr = a.ctypes.data
        if self.L.pomgpu_pin_host(C.c_void_p(ptr), a.nbytes) != 0:
            raise PomGpuError("cudaHostRegister failed")
        # the registration must not outlive the buffer: a later allocation that overlaps a stale
        # page-locked range makes cudaMemcpy fail with "invalid argument"
        weakref.finalize(a, self.L.pomgpu_unpin_host, C.c_void_p(ptr))
        return a

    def put_async(self, name, a):
        """Enqueue host->HBM copy of a (pinned, Fortran-ordered) array; no host wait."""
        assert a.flags.f_contiguous and a.shape == self.shapes[name]
        self._ck(self.L.pomgpu_push_async(self.h, name.encode(), a.ctypes.data_as(C.c_void_p)), f"push_async({name})")

    # -- forcing records, interpolated in time on the device (bounds_forcing.f:841-865,904-909,949-957)
    def put_record(self, name, slot, arr):
        """Record `slot` (0 = older "b", 1 = newer "f") of forcing field `name`; asynchronous if `arr` is pinned."""
        a = np.asfortranarray(self._rows(name, arr), dtype=np.float64)
        assert a.shape == self.shapes[name], (name, a.shape, self.shapes[name])
        self._ck(self.L.pomgpu_push_record(self.h, name.encode(), int(slot), a.ctypes.data_as(C.c_void_p)), f"push_record({name})")

    def rotate_record(self, name):
        self._ck(self.L.pomgpu_rotate_record(self.h, name.encode()), f"rotate_record({name})")

    def interp(self, name, fnew):
        self._ck(self.L.pomgpu_interp(self.h, name.encode(), float(fnew)), f"interp({name})")

    def wind(self, fnew): self._ck(self.L.pomgpu_wind(self.h, float(fnew)), "wind")
    def heat(self, fnew): self._ck(self.L.pomgpu_heat(self.h, float(fnew)), "heat")
    def lateral_bc(self, fnew): self._ck(self.L.pomgpu_lateral_bc(self.h, float(fnew)), "lateral_bc")

    def event_record(self, slot):
        self._ck(self.L.pomgpu_event_record(self.h, slot), "event_record")

    def event_elapsed_ms(self, a, b):
        return self.L.pomgpu_event_elapsed_ms(self.h, a, b)

    def profile_begin(self):
        self.L.pomgpu_profile_begin(self.h)

    def profile_end(self):
        import json
        buf = C.create_string_buffer(1 << 16)
        self._ck(self.L.pomgpu_profile_end(self.h, buf, len(buf)), "profile_end")
        return json.loads(buf.value.decode())

    def load(self, state):
        """state = {'consts': {...}, 'fields': {name: ndarray}} (extpom_b200.synthetic)."""
        for k, v in state["consts"].items():
            self.L.pomgpu_set_const(self.h, k.encode(), float(v))  # names outside blkcon are ignored
        for k, v in state["fields"].items():
            if k in self.shapes and (k not in F3D_OPT) and (k not in F3D_SCR):
                self.put(k, v)

    # -- the reference's subroutine surface ------------------------------------
    def step(self, iint, time=None, ramp=None):
        """pom/advance.f:21-32 for internal step iint; time as get_time (advance.f:66); ramp: the value get_time
        would set (advance.f:67-72), default = the context's current `ramp` constant (1 unless the caller set it)."""
        if time is None:
            time = self.getc("dti") * float(iint) / 86400.0 + self.getc("time0")
        if ramp is None:
            ramp = self.getc("ramp")
        self._ck(self.L.pomgpu_step(self.h, int(iint), float(time), float(ramp)), "step")

    def sync(self):
        self._ck(self.L.pomgpu_sync(self.h), "sync")

    def check_velocity(self):
        return self.L.pomgpu_check_velocity(self.h)

    def domain_stats_rows(self):
        """Per-owned-row partial sums of domain_stats (advance.f:644-755), shape (rows, 7)."""
        n = self.own[1] - self.own[0] + 1
        out = np.empty((n, 7), dtype=np.float64)
        self._ck(self.L.pomgpu_domain_stats_rows(self.h, out.ctypes.data_as(C.c_void_p)), "domain_stats")
        return out

    def domain_stats(self):
        return finish_domain_stats(self.domain_stats_rows())

    def check_velocity_lagged(self):
        """max|vaf| of the PREVIOUS call's step (0.0 first); never waits for the step just enqueued."""
        return self.L.pomgpu_check_velocity_lagged(self.h)

    def field_absmax_lagged(self, name):
        """max|field| of the PREVIOUS call's state (0.0 first): a one-scalar read-back that never waits."""
        return self.L.pomgpu_field_absmax_lagged(self.h, name.encode())

    def launch_count(self, reset=False):
        return self.L.pomgpu_launch_count(self.h, int(reset))

    def lateral_viscosity(self): self._ck(self.L.pomgpu_lateral_viscosity(self.h), "lateral_viscosity")
    def mode_interaction(self): self._ck(self.L.pomgpu_mode_interaction(self.h), "mode_interaction")
    def mode_external(self, iext): self._ck(self.L.pomgpu_mode_external(self.h, iext), "mode_external")
    def mode_internal(self, iint): self._ck(self.L.pomgpu_mode_internal(self.h, iint), "mode_internal")
    def internal_stage(self, iint, stage):
        self._ck(self.L.pomgpu_internal_stage(self.h, int(iint), int(stage)), "internal_stage")
    def advave(self): self._ck(self.L.pomgpu_advave(self.h), "advave")
    def advct(self): self._ck(self.L.pomgpu_advct(self.h), "advct")
    def advq(self): self._ck(self.L.pomgpu_advq(self.h), "advq")
    def advu(self): self._ck(self.L.pomgpu_advu(self.h), "advu")
    def advv(self): self._ck(self.L.pomgpu_advv(self.h), "advv")
    def baropg(self): self._ck(self.L.pomgpu_baropg(self.h), "baropg")
    def baropg_mcc(self): self._ck(self.L.pomgpu_baropg_mcc(self.h), "baropg_mcc")
    def profq(self): self._ck(self.L.pomgpu_profq(self.h), "profq")
    def profu(self): self._ck(self.L.pomgpu_profu(self.h), "profu")
    def profv(self): self._ck(self.L.pomgpu_profv(self.h), "profv")
    def vertvl(self): self._ck(self.L.pomgpu_vertvl(self.h), "vertvl")
    def realvertvl(self): self._ck(self.L.pomgpu_realvertvl(self.h), "realvertvl")

    def advt1(self, fb, f, fclim, ff):
        self._ck(self.L.pomgpu_advt1(self.h, fb.encode(), f.encode(), fclim.encode(), ff.encode()), "advt1")

    def advt2(self, fb, f, fclim, ff):
        self._ck(self.L.pomgpu_advt2(self.h, fb.encode(), f.encode(), fclim.encode(), ff.encode()), "advt2")

    def dens(self, si, ti, rhoo):
        self._ck(self.L.pomgpu_dens(self.h, si.encode(), ti.encode(), rhoo.encode()), "dens")

    def advq_fields(self, qb, q, qf):
        self._ck(self.L.pomgpu_advq_fields(self.h, qb.encode(), q.encode(), qf.encode()), "advq")

    def smol_adif(self, xm, ym, zw, ff):
        self._ck(self.L.pomgpu_smol_adif(self.h, xm.encode(), ym.encode(), zw.encode(), ff.encode()), "smol_adif")

    def bcond(self, idx): self._ck(self.L.pomgpu_bcond(self.h, int(idx)), f"bcond({idx})")
    def bcondorl(self, idx): self._ck(self.L.pomgpu_bcondorl(self.h, int(idx)), f"bcondorl({idx})")

    def proft(self, f, wfsurf, fsurf, nbc):
        self._ck(self.L.pomgpu_proft(self.h, f.encode(), wfsurf.encode(), fsurf.encode(), int(nbc)), "proft")


def finish_domain_stats(rows):
    """Add the per-row partial sums in global row order (sequentially, so the result does not depend
    on the decomposition) and form vtot, atot, mtot, stot, tavg, savg, eavg, ekin (advance.f:685-739)."""
    acc = [0.0] * 7
    for r in rows:
        for q in range(7):
            acc[q] += float(r[q])
    atot, ea, vtot, mtot, ta, stot, ekin = acc
    return dict(vtot=vtot, atot=atot, mtot=mtot, stot=stot, tavg=ta / vtot if vtot else 0.0,
                savg=stot / vtot if vtot else 0.0, eavg=ea / atot if atot else 0.0, ekin=ekin)


HALO_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_long,
                      C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_long)


class PomGroup:
    """The ordered (south -> north) strips held by this process, stepped together; halo rows
    are exchanged inside the step only when a kernel would read a stale one (csrc/pom_halo.cu)."""

    def __init__(self, strips):
        self.strips = list(strips)
        self.L = self.strips[0].L
        arr = (C.c_void_p * len(self.strips))(*[s.h for s in self.strips])
        self.h = self.L.pomgpu_group_create(len(self.strips), arr)
        if not self.h:
            raise PomGpuError("pomgpu_group_create failed (strips not contiguous, different ghost widths, "
                              "or a strip thinner than ghost+4 rows)")
        self._cb = None

    def close(self):
        if getattr(self, "h", None):
            self.L.pomgpu_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def nccl_unique_id(lib):
        buf = C.create_string_buffer(128)
        if lib.pomgpu_nccl_unique_id(buf) != 0:
            raise PomGpuError("ncclGetUniqueId failed (libnccl.so.2 not found?)")
        return buf.raw

    def connect_nccl(self, uid, rank, world):
        if self.L.pomgpu_group_connect_nccl(self.h, uid, rank, world) != 0:
            raise PomGpuError("ncclCommInitRank failed: " + self.L.pomgpu_last_error(self.strips[0].h).decode())

    def set_transport(self, fn):
        """fn(send_s, recv_s, send_n, recv_n): numpy views of the staging buffers (None = no neighbour)."""
        def cb(user, ss, rs, ns, sn, rn, nn):
            try:
                a = lambda p, n: np.ctypeslib.as_array(p, shape=(n,)) if n else None
                fn(a(ss, ns), a(rs, ns), a(sn, nn), a(rn, nn))
                return 0
            except Exception as e:  # noqa: BLE001
                print("halo transport failed:", e)
                return 1
        self._cb = HALO_CB(cb)
        self.L.pomgpu_group_set_transport(self.h, C.cast(self._cb, C.c_void_p), None)

    def step(self, iint, time=None, ramp=None):
        s0 = self.strips[0]
        if time is None:
            time = s0.getc("dti") * float(iint) / 86400.0 + s0.getc("time0")
        if ramp is None:
            ramp = s0.getc("ramp")
        rc = self.L.pomgpu_group_step(self.h, int(iint), float(time), float(ramp))
        if rc != 0:
            raise PomGpuError(f"group_step failed (rc={rc}): {self.L.pomgpu_last_error(s0.h).decode()}")

    def dens(self, si, ti, rhoo):
        if self.L.pomgpu_group_dens(self.h, si.encode(), ti.encode(), rhoo.encode()) != 0:
            raise PomGpuError("group dens failed")

    def baropg(self):
        self.L.pomgpu_group_baropg(self.h)

    def check_velocity(self):
        return self.L.pomgpu_group_check_velocity(self.h)

    def domain_stats(self):
        return finish_domain_stats(np.concatenate([s.domain_stats_rows() for s in self.strips], axis=0))

    def transport(self):
        """How the seam rows travel (pomgpu_group_transport)."""
        return ("device copies", "NCCL send/recv", "CUDA IPC peer copies", "host callback")[self.L.pomgpu_group_transport(self.h)]

    def exchanges(self, reset=False):
        f = C.c_long(0)
        n = self.L.pomgpu_group_exchanges(self.h, C.byref(f), int(reset))
        return n, f.value

    def gather(self, name):
        """Owned rows of every strip of this process stacked in j (tests)."""
        parts = []
        for s in self.strips:
            a = s.get(name)
            lo = s.own[0] - 1 - s.joff
            hi = s.own[1] - s.joff
            if a.ndim >= 2 and a.shape[1] == s.jml and name not in BIK:
                parts.append(a[:, lo:hi])
            elif a.shape[0] == s.jml and (name in BJ or name in BJK):
                parts.append(a[lo:hi])
            else:
                return a
        return np.concatenate(parts, axis=1 if parts[0].ndim >= 2 and name not in BJ + BJK else 0)
