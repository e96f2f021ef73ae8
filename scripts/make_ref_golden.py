"""Golden vectors FROM THE REFERENCE ITSELF: runs /root/reference's own Fortran source of the hot path
(pom/solver.f, pom/advance.f, pom/bounds_forcing.f, pom.h_dist -- read where they lie, translated and executed
by oracle/f77ref.py, nothing restated by hand) on the synthetic seamount states of extpom_b200/synthetic.py and
stores the resulting fields under tests/golden/ref_<case>.npz.  The reference is not present on the GPU box, so
the fixtures are committed; tests compare the C oracle (bitwise) and the CUDA path with them.

    python scripts/make_ref_golden.py [--check] [case ...]     (--check: also compare with the C oracle now)

Restoring: once iint >= 2 the reference's restore_interior (bounds_forcing.f:1023-1120) reads its target
fields from a netCDF file and nudges T, S towards them with tau = 1/trst.  The harness plays the file with the
climatology (oracle/f77ref.py: F77Ref._read_restore); `ref_restore_setup` gives the same records to the other
solvers (lrestore=1)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from extpom_b200 import synthetic as syn  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name -> (dims, steps, generator / namelist overrides)
REF_CASES = {
    "default":        ((16, 14, 7), 4, {}),
    "island":         ((16, 14, 7), 4, {"island": True}),
    "nadv1":          ((16, 14, 7), 3, {"nadv": 1}),
    "nitera2":        ((16, 14, 7), 3, {"nitera": 2, "island": True}),
    "nitera3_sw1":    ((16, 14, 7), 3, {"nitera": 3, "sw": 1.0}),
    "nbct2":          ((16, 14, 7), 3, {"nbct": 2}),
    "nbct3_nbcs3":    ((16, 14, 7), 3, {"nbct": 3, "nbcs": 3}),
    "nbct4_ntp3":     ((16, 14, 7), 3, {"nbct": 4, "ntp": 3}),
    "mode4":          ((16, 14, 7), 3, {"mode": 4}),
    "mode2_island":   ((16, 14, 7), 3, {"mode": 2, "island": True}),
    "npg2_island":    ((16, 14, 7), 3, {"npg": 2, "island": True}),
    "isplit5":        ((16, 14, 7), 3, {"isplit": 5, "dte": 6.0}),
    "kb12":           ((12, 11, 12), 3, {}),
    "medium":         ((32, 26, 12), 6, {}),
    # constants the generator does not take as keywords: set on the loaded solver ("_set")
    "alpha_ramp_ispadv": ((16, 14, 7), 4, {"_set": {"alpha": 0.225, "ramp": 0.6, "ispadv": 3, "smoth": 0.05}}),
    "bias_rf_visc":   ((16, 14, 7), 3, {"island": True, "_set": {"tbias": 1.0, "sbias": 0.5, "rfe": 0.5, "rfw": 0.25,
                                                                   "rfn": 0.75, "rfs": 0.1, "horcon": 0.2, "tprni": 0.3,
                                                                   "umol": 2.e-5}}),
    "kb21":           ((14, 12, 21), 4, {}),
    # a restart (time0 != 0): the first internal step runs the 3-D block too (advance.f:362), restore_interior
    # interpolates records that have not been read yet (bounds_forcing.f:1038: only at iint=2)
    "hotstart":       ((16, 14, 7), 3, {"island": True, "_set": {"time0": 1.5}}),
    # the other Jerlov water types of proft's short-wave penetration (solver.f:1560-1567, 1604-1615)
    "nbct2_ntp1":     ((16, 14, 7), 3, {"nbct": 2, "ntp": 1}),
    "nbct4_ntp5":     ((16, 14, 7), 3, {"nbct": 4, "ntp": 5}),
    # three MPDATA iterations with the reference's default smoothing parameter (solver.f:625-687, 1915)
    "nitera3_sw05":   ((16, 14, 7), 3, {"nitera": 3, "sw": 0.5, "island": True}),
    # all four sides open (no channel walls: the north / south branches of bcond, bcondorl work on wet points),
    # non-zero e_atmos, vfluxb, vfluxf, wssurf and open-boundary elevations / velocities
    "open_fluxes_obc": ((16, 14, 7), 4, {"walls": False, "fluxes": True, "obc": True}),
    "open_nadv1_npg2": ((16, 14, 7), 3, {"walls": False, "fluxes": True, "obc": True, "nadv": 1, "npg": 2, "island": True}),
    "open_mode2":     ((16, 14, 7), 3, {"walls": False, "fluxes": True, "obc": True, "mode": 2}),
    "open_nbct3_it2": ((16, 14, 7), 3, {"walls": False, "fluxes": True, "obc": True, "nbct": 3, "nbcs": 3, "nitera": 2}),
}

# the fields compared (state + diagnostics of the step; COMMON member names)
F3 = ("u v ub vb t s tb sb q2 q2b q2l q2lb km kh kq l w wr rho aam advx advy drhox drhoy uf vf").split()
F2 = ("el elb et etb etf ua uab va vab d dt egf egb utf utb vtf vtb adx2d ady2d drx2d dry2d aam2d advua advva "
      "wubot wvbot").split()


def ref_restore_setup(solver, st):
    """What the reference's restore_interior holds from iint=2 on, for the solvers that take the records as
    inputs (C oracle, CUDA): both bracketing records = the climatology, tau = 1./trst (single-precision 1.)."""
    f = st["fields"]
    kb = st["dims"][2]
    solver.set("lrestore", 1)
    tau = float(np.float32(1.) / np.float64(30.))
    for n, src in (("trstrb", "tclim"), ("trstrf", "tclim"), ("srstrb", "sclim"), ("srstrf", "sclim")):
        solver.put(n, np.asfortranarray(f[src]))
    full = np.full(f["tclim"].shape, tau, order="F")
    solver.put("taurstrb", full)
    solver.put("taurstrf", full)
    assert kb == full.shape[2]


def ref_restore_records(solver, st, iint):
    """The records as the reference's restore_interior holds them WHEN STEP `iint` RUNS, for the solvers that take
    them as inputs: nothing has been read before iint=2 (bounds_forcing.f:1038), so a restart (time0 != 0), whose
    first step already runs the tracer block (advance.f:362), interpolates zeros there -- tau=0, no nudging, only the
    masks -- and gets the climatology from step 2 on.  Call before every step."""
    if iint == 1:
        solver.set("lrestore", 1)
        zero = np.zeros(st["fields"]["tclim"].shape, order="F")
        for n in ("trstrb", "trstrf", "srstrb", "srstrf", "taurstrb", "taurstrf"):
            solver.put(n, zero)
    elif iint == 2:
        ref_restore_setup(solver, st)


def loaded(factory, dims, kw):
    """state + solver: generate, load, apply the case's extra constants, then the initial dens / baropg calls."""
    kw = dict(kw)
    extra = kw.pop("_set", {})
    st = syn.make_state(*dims, **kw)
    sv = factory(*dims)
    sv.load(st)
    for k, v in extra.items():
        sv.set(k, v)
    syn.finish_init(st, sv)
    return st, sv


def run_reference(dims, steps, kw):
    from oracle.f77ref import F77Ref
    st, r = loaded(F77Ref, dims, kw)
    for i in range(1, steps + 1):
        r.step(i)
    return st, r


def main(argv):
    check = "--check" in argv
    names = [a for a in argv if not a.startswith("--")] or list(REF_CASES)
    os.makedirs(GOLD, exist_ok=True)
    for name in names:
        dims, steps, kw = REF_CASES[name]
        t0 = time.time()
        st, r = run_reference(dims, steps, kw)
        out = {n: r.get(n) for n in F3 + F2}
        out["vamax"] = np.array(r.check_velocity())
        np.savez_compressed(os.path.join(GOLD, f"ref_{name}.npz"), **out)
        msg = f"ref_{name}: {dims} {steps} steps {kw} in {time.time() - t0:.1f} s"
        if check:
            from oracle.pomo import Oracle
            st2, o = loaded(Oracle, dims, kw)
            for i in range(1, steps + 1):
                ref_restore_records(o, st2, i)
                o.step(i)
            bad = {}
            for n in F3 + F2:
                a, b = out[n], o.get(n)
                if n in ("trstrb",):
                    continue
                if not np.array_equal(a, b):
                    bad[n] = float(np.abs(a - b).max() / (np.abs(a).max() + 1e-300))
            msg += "  | C oracle: " + ("BITWISE EQUAL" if not bad else f"MISMATCH {bad}")
        print(msg, flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
