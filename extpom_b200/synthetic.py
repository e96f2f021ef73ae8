"""Synthetic seamount-style input generator (HARNESS, not product).

The reference ships no idealised case and reads grid, initial conditions and forcing
from PnetCDF files that are not in the tree, so benchmarks and parity
tests need an author-defined input.  This module restates what the reference's
`initialize` derives before the first `advance` so that the hot path sees what the
Fortran driver would hand it:

  read_input constants ........ pom/initialize.f:80-168,178-191
  vertical grid dz,dzz ........ pom/initialize.f:331-335
  art, aru, arv ............... pom/initialize.f:363-384
  dum, dvm from fsm ........... pom/io_pnetcdf.F:2243-2254
  rmean, rho, tsurf, BC arrays  pom/initialize.f:416-460
  update_initial .............. pom/initialize.f:472-518   (calls baropg)
  bottom_friction ............. pom/initialize.f:534-541
  check_cflmin ................ pom/parallel_mpi.f:496-499

`dens` and `baropg` are part of the replaced path (pom/solver.f:848,1162): the
generator calls them through the `solver` object it is given (PomGpu for the product,
the CPU oracle in the CPU tests), exactly like the Fortran `initialize` would call the
drop-in's entry points.
"""
import numpy as np

SEED = 20161018


def default_consts(dte=6.0, isplit=30, nadv=2, nitera=1, sw=0.5, mode=3, npg=1,
                   nbct=1, nbcs=1, ntp=2, aam_init=500.0):
    """read_input (pom/initialize.f:67-191) with the pom.nml_dist namelist keys as arguments."""
    c = dict(
        rhoref=1025.0, tbias=0.0, sbias=0.0, grav=9.806, kappa=0.4, z0b=0.01,
        cbcmin=0.0025, cbcmax=1.0, horcon=0.1, tprni=0.1, umol=1.0e-6, vmaxl=100.0,
        slmax=2.0, ntp=ntp, nbct=nbct, nbcs=nbcs, ispadv=1, smoth=0.10, alpha=0.0,
        aam_init=aam_init, mode=mode, nadv=nadv, nitera=nitera, sw=sw, npg=npg,
        dte=float(dte), isplit=int(isplit), small=1.0e-9, time0=0.0, time=0.0, ramp=1.0,
        rfe=1.0, rfw=1.0, rfn=1.0, rfs=1.0, iint=0, iext=0, error_status=0, lrestore=0,
    )
    c["dti"] = c["dte"] * float(isplit)          # initialize.f:181
    c["dte2"] = c["dte"] * 2                     # :182
    c["dti2"] = c["dti"] * 2                     # :183
    c["ispi"] = 1.0 / float(isplit)              # :190
    c["isp2i"] = 1.0 / (2.0 * float(isplit))     # :191
    return c


def sigma_levels(kb):
    """Thin layers doubling in thickness over the top six, uniform below; z(1)=0, z(kb)=-1."""
    w = np.array([2.0 ** (min(k, 6) - 1) for k in range(1, kb)])
    z = np.zeros(kb)
    z[1:] = -np.cumsum(w) / w.sum()
    z[kb - 1] = -1.0
    zz = np.zeros(kb)
    zz[:kb - 1] = 0.5 * (z[:kb - 1] + z[1:])
    zz[kb - 1] = 2.0 * zz[kb - 2] - zz[kb - 3]
    dz = np.zeros(kb)
    dzz = np.zeros(kb)
    dz[:kb - 1] = z[:kb - 1] - z[1:]            # initialize.f:331-335
    dzz[:kb - 1] = zz[:kb - 1] - zz[1:]
    return z, zz, dz, dzz


def cflmin(f, grav, small):
    """check_cflmin_mpi (pom/parallel_mpi.f:496-499)."""
    cfl = 0.5 / np.sqrt(1.0 / f["dx"] ** 2 + 1.0 / f["dy"] ** 2) / np.sqrt(grav * (f["h"] + small)) * f["fsm"]
    return float(cfl[cfl > 0].min())


def _row_noise(kind, tag, im, rows, nk):
    """Symmetry-breaking noise that depends only on (seed, field tag, GLOBAL row j): a strip
    of the domain generates exactly the values the whole-domain state holds in its rows."""
    out = np.empty((im, len(rows), nk), order="F")
    for n, j in enumerate(rows):
        rng = np.random.default_rng([SEED, tag, int(j)])
        out[:, n, :] = rng.standard_normal((im, nk)) if kind == "normal" else rng.uniform(-1, 1, (im, nk))
    return out


def make_state(im, jm, kb, delta=8000.0, wind=True, noise=True, island=False, rows=None, walls=True, fluxes=False,
               obc=False, **nml):
    """Everything `initialize` sets that does not need dens/baropg.  Arrays are
    Fortran-ordered (i fastest) float64 with shapes (im,jm[,kb]).

    rows=(j_lo, j_hi) (global, 1-based, inclusive) generates only that band of rows of the
    (im, jm, kb) domain -- what one rank of distribute_mpi (pom/parallel_mpi.f:76-119) holds --
    without ever materialising the global arrays; every value equals the whole-domain one.

    Parity-test variants (the BASELINE workload keeps the defaults): walls=False opens the north and south sides too
    (no land rows at j=1, jm), so that the north / south branches of bcond / bcondorl work on wet points; fluxes=True
    makes e_atmos, vfluxb, vfluxf and wssurf non-zero (they multiply terms of vertvl, advu/advv, mode_external and
    proft that are otherwise never seen); obc=True gives the open-boundary elevations and velocities
    (ele..els, vabe..uabs, ube, ubw, vbn, vbs) non-zero values."""
    c = default_consts(**nml)
    F = lambda *shp: np.zeros(shp, order="F")
    f = {}
    z, zz, dz, dzz = sigma_levels(kb)
    f.update(z=z, zz=zz, dz=dz, dzz=dzz)
    jlo, jhi = (1, jm) if rows is None else (int(rows[0]), int(rows[1]))
    assert 1 <= jlo <= jhi <= jm
    jmg, jrows = jm, np.arange(jlo, jhi + 1)
    # one extra row to the south (dvm, arv look at j-1); dropped again before returning
    pad = 1 if jlo > 1 else 0
    jall = np.arange(jlo - pad, jhi + 1)
    jm = len(jall)                        # local extent from here on (global: jmg)

    i1 = np.arange(1, im + 1, dtype=np.float64)[:, None]
    j1 = jall.astype(np.float64)[None, :]
    dx = F(im, jm); dy = F(im, jm)
    dx[...] = delta - delta * np.sin(np.pi * i1 / im) / 2.0
    dy[...] = delta - delta * np.sin(np.pi * j1 / jmg) / 2.0
    dyg = delta - delta * np.sin(np.pi * np.arange(1, jmg + 1, dtype=np.float64) / jmg) / 2.0
    x = np.cumsum(dx[:, 0]) - 0.5 * dx[:, 0]
    yg = np.cumsum(dyg) - 0.5 * dyg
    y = yg[jall - 1]
    xc, yc = x[im // 2], yg[jmg // 2]
    ra = 25.0e3 * (im / 65.0)
    r2 = (x[:, None] - xc) ** 2 + (y[None, :] - yc) ** 2
    h = F(im, jm)
    h[...] = 4500.0 * (1.0 - 0.9 * np.exp(-r2 / ra ** 2))
    if walls:
        h[:, jall == 1] = 1.0
        h[:, jall == jmg] = 1.0   # closed channel walls
    if island:                     # a dry patch so that interior masks are exercised
        ic, jc = im // 3, (2 * jmg) // 3
        h[ic - 1:ic + 2, (jall >= jc) & (jall <= jc + 2)] = 1.0
    fsm = F(im, jm); fsm[...] = np.where(h > 1.0, 1.0, 0.0)
    dum = fsm.copy(order="F"); dvm = fsm.copy(order="F")
    # io_pnetcdf.F:2243-2254
    dvm[:, 1:][(fsm[:, :-1] == 0) & (fsm[:, 1:] != 0)] = 0.0
    dum[1:, :][(fsm[:-1, :] == 0) & (fsm[1:, :] != 0)] = 0.0
    cor = F(im, jm); cor[...] = 1.0e-4
    art = F(im, jm); art[...] = dx * dy                        # initialize.f:363
    aru = F(im, jm); arv = F(im, jm)
    aru[1:, 1:] = 0.25 * (dx[1:, 1:] + dx[:-1, 1:]) * (dy[1:, 1:] + dy[:-1, 1:])   # :366-371
    arv[1:, 1:] = 0.25 * (dx[1:, 1:] + dx[1:, :-1]) * (dy[1:, 1:] + dy[1:, :-1])
    aru[0, :] = aru[1, :]; arv[0, :] = arv[1, :]               # :375-378
    if jall[0] == 1:
        aru[:, 0] = aru[:, 1]; arv[:, 0] = arv[:, 1]           # :380-383
    f.update(dx=dx, dy=dy, h=h, fsm=fsm, dum=dum, dvm=dvm, cor=cor, art=art, aru=aru, arv=arv)

    tb = F(im, jm, kb); sb = F(im, jm, kb)
    tb[...] = 5.0 + 15.0 * np.exp(zz[None, None, :] * h[:, :, None] / 1000.0)
    sb[...] = 35.0
    tclim = tb.copy(order="F"); sclim = sb.copy(order="F")
    ub = F(im, jm, kb); vb = F(im, jm, kb)
    ub[:, :, :kb - 1] = 0.2 * dum[:, :, None]
    uab = F(im, jm); uab[...] = 0.2 * dum
    vab = F(im, jm)
    if noise:   # both branches of every upwind / abs() test must run
        tb[...] += 1.0e-2 * _row_noise("normal", 1, im, jall, kb) * fsm[:, :, None]
        ub[:, :, :kb - 1] += 1.0e-2 * _row_noise("uniform", 2, im, jall, kb - 1) * dum[:, :, None]
        vb[:, :, :kb - 1] += 1.0e-2 * _row_noise("uniform", 3, im, jall, kb - 1) * dvm[:, :, None]
    f.update(tb=tb, sb=sb, tclim=tclim, sclim=sclim, ub=ub, vb=vb, uab=uab, vab=vab)
    for n in "elb etb e_atmos vfluxb vfluxf wusurf wvsurf wtsurf wssurf swrad".split():
        f[n] = F(im, jm)                                         # initialize_arrays :270-294
    if wind:
        f["wusurf"][...] = -0.5e-4 * (1.0 + 0.5 * np.sin(2 * np.pi * j1 / jmg)) * fsm
        f["wvsurf"][...] = 0.2e-4 * np.cos(2 * np.pi * i1 / im) * fsm
        f["wtsurf"][...] = 2.0e-5 * np.sin(2 * np.pi * i1 / im) * np.cos(np.pi * j1 / jmg) * fsm
        f["swrad"][...] = -1.0e-5 * fsm
    if fluxes:
        f["e_atmos"][...] = 0.05 * np.sin(2 * np.pi * i1 / im) * np.cos(2 * np.pi * j1 / jmg) * fsm
        f["vfluxf"][...] = 2.0e-7 * (1.0 + 0.5 * np.cos(np.pi * i1 / im) * np.sin(np.pi * j1 / jmg)) * fsm
        f["vfluxb"][...] = 0.9 * f["vfluxf"]
        f["wssurf"][...] = 1.0e-6 * np.sin(np.pi * j1 / jmg) * np.cos(np.pi * i1 / im) * fsm
    f["tsurf"] = tb[:, :, 0].copy(order="F")                     # initialize.f:441-442
    f["ssurf"] = sb[:, :, 0].copy(order="F")
    km1 = kb - 1
    for nm, src in (("t", tb), ("s", sb)):                       # :449-460
        e = np.zeros((jm, kb), order="F"); w_ = np.zeros((jm, kb), order="F")
        n_ = np.zeros((im, kb), order="F"); s_ = np.zeros((im, kb), order="F")
        e[:, :km1] = src[im - 1, :, :km1]; w_[:, :km1] = src[0, :, :km1]
        if jall[-1] == jmg: n_[:, :km1] = src[:, jm - 1, :km1]      # only the strips that hold the
        if jall[0] == 1: s_[:, :km1] = src[:, 0, :km1]             # north / south edge use these
        f[nm + "be"], f[nm + "bw"], f[nm + "bn"], f[nm + "bs"] = e, w_, n_, s_
    f["uabw"] = uab[1, :].copy(); f["uabe"] = uab[im - 2, :].copy()
    for n in ("ele", "elw", "vabe", "vabw"): f[n] = np.zeros(jm)
    for n in ("eln", "els", "vabn", "vabs", "uabn", "uabs"): f[n] = np.zeros(im)
    if obc:
        jj, ii = jall.astype(np.float64), np.arange(1, im + 1, dtype=np.float64)
        sk = np.linspace(1.0, 0.6, kb)[None, :] * (np.arange(kb) < km1)[None, :]
        f["ele"] = 0.02 * np.sin(2 * np.pi * jj / jmg); f["elw"] = -0.015 * np.cos(2 * np.pi * jj / jmg)
        f["eln"] = 0.01 * np.sin(2 * np.pi * ii / im); f["els"] = -0.012 * np.cos(2 * np.pi * ii / im)
        f["vabe"] = 0.01 * np.cos(np.pi * jj / jmg); f["vabw"] = -0.008 * np.sin(np.pi * jj / jmg)
        f["vabn"] = 0.03 * np.sin(np.pi * ii / im); f["vabs"] = 0.025 * np.sin(np.pi * ii / im)
        f["uabn"] = 0.004 * np.cos(np.pi * ii / im); f["uabs"] = -0.003 * np.cos(np.pi * ii / im)
        f["ube"] = np.asfortranarray(0.18 * (1.0 + 0.1 * np.sin(np.pi * jj / jmg))[:, None] * sk)
        f["ubw"] = np.asfortranarray(0.21 * (1.0 - 0.1 * np.sin(np.pi * jj / jmg))[:, None] * sk)
        f["vbn"] = np.asfortranarray(0.03 * np.sin(np.pi * ii / im)[:, None] * sk)
        f["vbs"] = np.asfortranarray(0.025 * np.sin(np.pi * ii / im)[:, None] * sk)

    # update_initial (initialize.f:472-495)
    f["ua"] = uab.copy(order="F"); f["va"] = vab.copy(order="F")
    f["el"] = f["elb"].copy(order="F"); f["et"] = f["etb"].copy(order="F")
    f["etf"] = f["et"].copy(order="F")
    f["d"] = h + f["el"]; f["dt"] = h + f["et"]
    w3 = F(im, jm, kb); w3[:, :, 0] = f["vfluxf"]; f["w"] = w3
    lfac = float(np.float32(0.1))                                # `0.1*dt`: single literal (:484)
    l3 = F(im, jm, kb); l3[...] = lfac * f["dt"][:, :, None]
    q2b = F(im, jm, kb); q2b[...] = c["small"]
    q2lb = l3 * q2b
    kh = l3 * np.sqrt(q2b)
    aam = F(im, jm, kb); aam[...] = c["aam_init"]
    f.update(l=l3, q2b=q2b, q2lb=q2lb, kh=kh, km=kh.copy(order="F"), kq=kh.copy(order="F"), aam=aam,
             q2=q2b.copy(order="F"), q2l=q2lb.copy(order="F"), t=tb.copy(order="F"),
             s=sb.copy(order="F"), u=ub.copy(order="F"), v=vb.copy(order="F"))
    # bottom_friction (initialize.f:534-541)
    cbc = (c["kappa"] / np.log((1.0 + zz[kb - 2]) * h / c["z0b"])) ** 2
    f["cbc"] = np.asfortranarray(np.minimum(c["cbcmax"], np.maximum(c["cbcmin"], cbc)))
    for n in ("drx2d", "dry2d", "wubot", "wvbot", "egb", "egf", "utb", "vtb", "utf", "vtf",
              "adx2d", "ady2d", "advua", "advva", "aam2d", "elf", "uaf", "vaf"):
        f[n] = F(im, jm)
    for n in ("drhox", "drhoy", "advx", "advy", "rho", "rmean", "uf", "vf", "wr", "dtef"):
        f[n] = F(im, jm, kb)
    cmin = cflmin(f, c["grav"], c["small"])
    assert cmin >= c["dte"], f"dte={c['dte']} violates CFL ({cmin:.3f})"
    if pad:                               # drop the helper row south of the band
        for n, a in f.items():
            if n in ("z", "zz", "dz", "dzz") or n[1:] in ("bn", "bs") or n in ("eln", "els", "vabn", "vabs", "uabn", "uabs"):
                continue
            if a.ndim >= 2 and a.shape[0] == im and a.shape[1] == jm:
                f[n] = np.asfortranarray(a[:, pad:])
            elif a.shape[0] == jm:
                f[n] = np.asfortranarray(a[pad:])
            else:
                raise AssertionError((n, a.shape))
    return {"dims": (im, len(jrows), kb), "global_jm": jmg, "rows": (jlo, jhi), "consts": c, "fields": f,
            "cflmin": cmin}


def finish_init(state, solver):
    """The part of initial_conditions/update_initial that calls the replaced path:
    rmean=dens(sclim,tclim), rho=dens(sb,tb) (initialize.f:416,425), baropg (:502) and the
    drx2d/dry2d sums (:510-517).  `solver` already holds the state (solver.load(state))."""
    f = state["fields"]
    solver.dens("sclim", "tclim", "rmean")
    solver.dens("sb", "tb", "rho")
    if int(state["consts"].get("npg", 1)) == 2:      # initialize.f:502-505
        solver.baropg_mcc()
    else:
        solver.baropg()
    dz = f["dz"]
    for n in ("rmean", "rho", "drhox", "drhoy"):
        f[n] = solver.get(n)
    drx = np.zeros_like(f["drx2d"]); dry = np.zeros_like(f["dry2d"])
    for k in range(state["dims"][2] - 1):
        drx = drx + f["drhox"][:, :, k] * dz[k]
        dry = dry + f["drhoy"][:, :, k] * dz[k]
    f["drx2d"] = np.asfortranarray(drx); f["dry2d"] = np.asfortranarray(dry)
    solver.put("drx2d", f["drx2d"]); solver.put("dry2d", f["dry2d"])
    return state


def seamount(im, jm, kb, solver_factory, **kw):
    """Full initial state + a solver loaded with it.  solver_factory(im,jm,kb) -> solver."""
    st = make_state(im, jm, kb, **kw)
    sv = solver_factory(im, jm, kb)
    sv.load(st)
    finish_init(st, sv)
    return st, sv
