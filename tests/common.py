"""Shared helpers of the parity tests."""
import hashlib

import numpy as np

# prognostic + diagnostic fields compared after whole steps (restart list of
# io_pnetcdf.F:1661-2082 plus the mode-interaction integrals)
F3 = "u ub v vb t tb s sb q2 q2b q2l q2lb w rho km kh kq aam advx advy drhox drhoy l wr".split()
F2 = ("el elb et etb etf ua uab va vab d dt egf egb utf vtf utb vtb wubot wvbot adx2d ady2d "
      "drx2d dry2d aam2d advua advva").split()
# Level kb of t,s,tb,sb is scratch in the reference: `t=uf` (advance.f:447) copies whatever
# uf(:,:,kb) held from the q2 stage; nothing on the path reads it.  It is not compared.
KB_SCRATCH = ("t", "tb", "s", "sb")

# Tolerance of the GPU (and host-emulated) path against the oracle, per field, as
# max-abs error normalised by the field's max-abs.  Arithmetic is the same IEEE fp64
# sequence without FMA on both sides; the only differences are |S|**1.5 in dens
# (glibc pow vs an fma-based x*sqrt(x)) and exp() in proft's short-wave term (quad
# precision in the reference), both at the 1-ulp level.
RTOL = 1e-11


def rel_err(a, b):
    return float(np.abs(a - b).max() / (np.abs(a).max() + 1e-300))


def rel_l2(a, b):
    return float(np.sqrt(((a - b) ** 2).sum()) / (np.sqrt((a ** 2).sum()) + 1e-300))


def compare(o, g, names=None, kb=None, skip=()):
    """{field: (rel max-abs, rel L2)} between two solvers exposing .get(name)."""
    out = {}
    for n in names or (F3 + F2):
        if n in skip:
            continue
        a, b = o.get(n), g.get(n)
        if n in KB_SCRATCH and a.ndim == 3:
            a, b = a[:, :, :-1], b[:, :, :-1]
        out[n] = (rel_err(a, b), rel_l2(a, b))
    return out


def assert_close(o, g, tol=RTOL, **kw):
    errs = compare(o, g, **kw)
    bad = {k: v for k, v in errs.items() if not (v[0] <= tol)}
    assert not bad, f"fields beyond rel tol {tol}: {bad}"
    return max(v[0] for v in errs.values()) if errs else 0.0


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def spin_up(solver, nsteps, start=1):
    for iint in range(start, start + nsteps):
        solver.step(iint)
