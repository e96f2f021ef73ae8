"""A SECOND, independent restatement (numpy) of the WHOLE hot path.

TEST INFRASTRUCTURE ONLY (see oracle/pomo.h: the C oracle is pinned bitwise against the reference's own
source executed by oracle/f77ref.py; this older guard stays).  The reference cannot be
compiled here, so the C oracle is itself a restatement; this module restates
    dens    pom/solver.f:1162-1209      baropg  pom/solver.f:848-940
    vertvl  pom/solver.f:1970-2021      advq    pom/solver.f:411-477
    advave  pom/solver.f:6-121          proft   pom/solver.f:1541-1683
    profq   pom/solver.f:1212-1538      advt2 + smol_adif  pom/solver.f:577-731,1880-1967
    baropg_mcc  pom/solver.f:943-1159   advct   pom/solver.f:201-409
    advu / advv   pom/solver.f:734-845  profu / profv  pom/solver.f:1686-1877
    realvertvl    pom/solver.f:2024-2066   advt1  pom/solver.f:480-574
    mode_external + bcond(1), bcond(2)  pom/advance.f:205-353, pom/bounds_forcing.f:18-83
    lateral_viscosity, mode_interaction, mode_internal  pom/advance.f:96-202,356-537
    bcond(4), bcond(6), bcondorl(3), bcondorl(5)  pom/bounds_forcing.f:151-324,418-487,550-561
and `step` = pom/advance.f:21-32,
a second time, written from the Fortran text with whole-array slices instead of loops, and
tests/test_oracle_np.py requires the two restatements to agree BITWISE -- routine by routine and
for whole internal steps (same IEEE operations in the same order; numpy does not contract to
FMA).  A transcription slip would have to be made twice, in two different notations, to go
unnoticed.

Arrays are (im,jm[,kb]) Fortran-ordered; `sl(a,b)` is the Fortran index range a:b (1-based,
inclusive), optionally shifted: x[sl(2,imm1,-1), ...] is x(i-1,...) for i=2..imm1.
"""
import numpy as np


def sl(a, b, off=0):
    return slice(a - 1 + off, b + off)


class NP:
    def __init__(self, f, c):
        """f: dict of field arrays (not modified), c: dict of constants."""
        self.f, self.c = f, c
        self.im, self.jm = f["h"].shape
        self.kb = f["z"].shape[0]

    # ------------------------------------------------------------------ dens
    def dens(self, si, ti):
        f, c = self.f, self.c
        kbm1 = self.kb - 1
        rhoo = np.zeros_like(si)
        for k in range(1, kbm1 + 1):
            tr = ti[:, :, k - 1] + c["tbias"]
            sr = si[:, :, k - 1] + c["sbias"]
            tr2 = tr * tr
            tr3 = tr2 * tr
            tr4 = tr3 * tr
            p = c["grav"] * c["rhoref"] * (-f["zz"][k - 1] * f["h"]) * 1.e-5
            rhor = (-0.157406 + 6.793952e-2 * tr - 9.095290e-3 * tr2 + 1.001685e-4 * tr3
                    - 1.120083e-6 * tr4 + 6.536332e-9 * tr4 * tr)
            rhor = (rhor + (0.824493 - 4.0899e-3 * tr + 7.6438e-5 * tr2 - 8.2467e-7 * tr3 + 5.3875e-9 * tr4) * sr
                    + (-5.72466e-3 + 1.0227e-4 * tr - 1.6546e-6 * tr2) * np.abs(sr) ** 1.5
                    + 4.8314e-4 * sr * sr)
            cr = 1449.1 + .0821 * p + 4.55 * tr - .045 * tr2 + 1.34 * (sr - 35.)
            rhor = rhor + 1.e5 * p / (cr * cr) * (1. - 2. * p / (cr * cr))
            rhoo[:, :, k - 1] = rhor / c["rhoref"] * f["fsm"]
        return rhoo

    # ---------------------------------------------------------------- baropg
    def baropg(self, rho_in, drhox0, drhoy0):
        """Returns (drhox, drhoy, rho) like the in-place Fortran; drhox0/drhoy0 = previous content
        (edge cells are never assigned)."""
        f, c = self.f, self.c
        im, jm, kb = self.im, self.jm, self.kb
        imm1, jmm1, kbm1 = im - 1, jm - 1, kb - 1
        zz, dt, grav = f["zz"], f["dt"], c["grav"]
        rho = rho_in - f["rmean"]
        I, J = sl(2, imm1), sl(2, jmm1)
        Iw, Js = sl(2, imm1, -1), sl(2, jmm1, -1)
        out = []
        for (A, B, mask, met) in (((Iw, J), None, f["dum"], f["dy"]), ((I, Js), None, f["dvm"], f["dx"])):
            d = np.zeros((im, jm, kb), order="F")
            dsum = dt[I, J] + dt[A]
            ddif = dt[I, J] - dt[A]
            d[I, J, 0] = .5 * grav * (-zz[0]) * dsum * (rho[I, J, 0] - rho[A + (0,)])
            for k in range(2, kbm1 + 1):
                d[I, J, k - 1] = (d[I, J, k - 2]
                                  + grav * .25 * (zz[k - 2] - zz[k - 1]) * dsum
                                  * (rho[I, J, k - 1] - rho[A + (k - 1,)] + rho[I, J, k - 2] - rho[A + (k - 2,)])
                                  + grav * .25 * (zz[k - 2] + zz[k - 1]) * ddif
                                  * (rho[I, J, k - 1] + rho[A + (k - 1,)] - rho[I, J, k - 2] - rho[A + (k - 2,)]))
            for k in range(1, kbm1 + 1):
                d[I, J, k - 1] = .25 * dsum * d[I, J, k - 1] * mask[I, J] * (met[I, J] + met[A])
            out.append(d)
        drhox, drhoy = drhox0.copy(order="F"), drhoy0.copy(order="F")
        drhox[I, J, :kbm1] = out[0][I, J, :kbm1]
        drhoy[I, J, :kbm1] = out[1][I, J, :kbm1]
        drhox[I, J, :] = c["ramp"] * drhox[I, J, :]
        drhoy[I, J, :] = c["ramp"] * drhoy[I, J, :]
        return drhox, drhoy, rho + f["rmean"]

    # ---------------------------------------------------------------- vertvl
    def vertvl(self, w_in):
        f, c = self.f, self.c
        im, jm, kb = self.im, self.jm, self.kb
        imm1, jmm1, kbm1 = im - 1, jm - 1, kb - 1
        dx, dy, dt, u, v = f["dx"], f["dy"], f["dt"], f["u"], f["v"]
        xflux = np.zeros((im, jm, kb), order="F")
        yflux = np.zeros((im, jm, kb), order="F")
        I, J = sl(2, im), sl(2, jm)
        cx = .25 * (dy[I, J] + dy[sl(2, im, -1), J]) * (dt[I, J] + dt[sl(2, im, -1), J])
        cy = .25 * (dx[I, J] + dx[I, sl(2, jm, -1)]) * (dt[I, J] + dt[I, sl(2, jm, -1)])
        for k in range(1, kbm1 + 1):
            xflux[I, J, k - 1] = cx * u[I, J, k - 1]
            yflux[I, J, k - 1] = cy * v[I, J, k - 1]
        w = w_in.copy(order="F")
        I, J = sl(2, imm1), sl(2, jmm1)
        w[I, J, 0] = 0.5 * (f["vfluxb"][I, J] + f["vfluxf"][I, J])
        for k in range(1, kbm1 + 1):
            w[I, J, k] = w[I, J, k - 1] + f["dz"][k - 1] * (
                (xflux[sl(2, imm1, 1), J, k - 1] - xflux[I, J, k - 1]
                 + yflux[I, sl(2, jmm1, 1), k - 1] - yflux[I, J, k - 1]) / (dx[I, J] * dy[I, J])
                + (f["etf"][I, J] - f["etb"][I, J]) / c["dti2"])
        return w

    # ------------------------------------------------------------------ advq
    def advq(self, qb, q, qf_in):
        f, c = self.f, self.c
        im, jm, kb = self.im, self.jm, self.kb
        imm1, jmm1, kbm1 = im - 1, jm - 1, kb - 1
        dx, dy, dt, h, u, v, w, aam = f["dx"], f["dy"], f["dt"], f["h"], f["u"], f["v"], f["w"], f["aam"]
        xflux = np.zeros((im, jm, kb), order="F")
        yflux = np.zeros((im, jm, kb), order="F")
        I, J, Iw, Js = sl(2, im), sl(2, jm), sl(2, im, -1), sl(2, jm, -1)
        for k in range(2, kbm1 + 1):
            K, Km = k - 1, k - 2
            xflux[I, J, K] = .125 * (q[I, J, K] + q[Iw, J, K]) * (dt[I, J] + dt[Iw, J]) * (u[I, J, K] + u[I, J, Km])
            yflux[I, J, K] = .125 * (q[I, J, K] + q[I, Js, K]) * (dt[I, J] + dt[I, Js]) * (v[I, J, K] + v[I, J, Km])
        for k in range(2, kbm1 + 1):
            K, Km = k - 1, k - 2
            xflux[I, J, K] = (xflux[I, J, K]
                              - .25 * (aam[I, J, K] + aam[Iw, J, K] + aam[I, J, Km] + aam[Iw, J, Km])
                              * (h[I, J] + h[Iw, J]) * (qb[I, J, K] - qb[Iw, J, K]) * f["dum"][I, J]
                              / (dx[I, J] + dx[Iw, J]))
            yflux[I, J, K] = (yflux[I, J, K]
                              - .25 * (aam[I, J, K] + aam[I, Js, K] + aam[I, J, Km] + aam[I, Js, Km])
                              * (h[I, J] + h[I, Js]) * (qb[I, J, K] - qb[I, Js, K]) * f["dvm"][I, J]
                              / (dy[I, J] + dy[I, Js]))
            xflux[I, J, K] = .5 * (dy[I, J] + dy[Iw, J]) * xflux[I, J, K]
            yflux[I, J, K] = .5 * (dx[I, J] + dx[I, Js]) * yflux[I, J, K]
        qf = qf_in.copy(order="F")
        I, J = sl(2, imm1), sl(2, jmm1)
        art = f["art"]
        for k in range(2, kbm1 + 1):
            K = k - 1
            t = ((w[I, J, K - 1] * q[I, J, K - 1] - w[I, J, K + 1] * q[I, J, K + 1]) * art[I, J]
                 / (f["dz"][K] + f["dz"][K - 1])
                 + xflux[sl(2, imm1, 1), J, K] - xflux[I, J, K]
                 + yflux[I, sl(2, jmm1, 1), K] - yflux[I, J, K])
            qf[I, J, K] = (((h[I, J] + f["etb"][I, J]) * art[I, J] * qb[I, J, K] - c["dti2"] * t)
                           / ((h[I, J] + f["etf"][I, J]) * art[I, J]))
        return qf

    # ---------------------------------------------------------------- advave
    def advave(self):
        f = self.f
        im, jm = self.im, self.jm
        imm1, jmm1 = im - 1, jm - 1
        d, ua, va, uab, vab, am, dx, dy = (f[n] for n in ("d", "ua", "va", "uab", "vab", "aam2d", "dx", "dy"))
        Z = lambda: np.zeros((im, jm), order="F")
        fluxua, fluxva, tps = Z(), Z(), Z()
        J, Js = sl(2, jm), sl(2, jm, -1)
        I, Ie, Iw = sl(2, imm1), sl(2, imm1, 1), sl(2, imm1, -1)
        fluxua[I, J] = .125 * ((d[Ie, J] + d[I, J]) * ua[Ie, J] + (d[I, J] + d[Iw, J]) * ua[I, J]) * (ua[Ie, J] + ua[I, J])
        I2, I2w = sl(2, im), sl(2, im, -1)
        fluxva[I2, J] = (.125 * ((d[I2, J] + d[I2, Js]) * va[I2, J] + (d[I2w, J] + d[I2w, Js]) * va[I2w, J])
                         * (ua[I2, J] + ua[I2, Js]))
        fluxua[I, J] = fluxua[I, J] - d[I, J] * 2. * am[I, J] * (uab[Ie, J] - uab[I, J]) / dx[I, J]
        dy4 = dy[I2, J] + dy[I2w, J] + dy[I2, Js] + dy[I2w, Js]
        dx4 = dx[I2, J] + dx[I2w, J] + dx[I2, Js] + dx[I2w, Js]
        tps[I2, J] = (.25 * (d[I2, J] + d[I2w, J] + d[I2, Js] + d[I2w, Js])
                      * (am[I2, J] + am[I2, Js] + am[I2w, J] + am[I2w, Js])
                      * ((uab[I2, J] - uab[I2, Js]) / dy4 + (vab[I2, J] - vab[I2w, J]) / dx4))
        fluxua[I2, J] = fluxua[I2, J] * dy[I2, J]
        fluxva[I2, J] = (fluxva[I2, J] - tps[I2, J]) * .25 * dx4
        advua = Z()
        Ji, Jn = sl(2, jmm1), sl(2, jmm1, 1)
        advua[I, Ji] = fluxua[I, Ji] - fluxua[Iw, Ji] + fluxva[I, Jn] - fluxva[I, Ji]
        # v half
        fluxua, fluxva = Z(), Z()
        fluxua[I2, J] = (.125 * ((d[I2, J] + d[I2w, J]) * ua[I2, J] + (d[I2, Js] + d[I2w, Js]) * ua[I2, Js])
                         * (va[I2w, J] + va[I2, J]))
        Jss = sl(2, jmm1, -1)
        fluxva[I2, Ji] = (.125 * ((d[I2, Jn] + d[I2, Ji]) * va[I2, Jn] + (d[I2, Ji] + d[I2, Jss]) * va[I2, Ji])
                          * (va[I2, Jn] + va[I2, Ji]))
        fluxva[I2, Ji] = fluxva[I2, Ji] - d[I2, Ji] * 2. * am[I2, Ji] * (vab[I2, Jn] - vab[I2, Ji]) / dy[I2, Ji]
        fluxva[I2, J] = fluxva[I2, J] * dx[I2, J]
        fluxua[I2, J] = (fluxua[I2, J] - tps[I2, J]) * .25 * dy4
        advva = Z()
        advva[I, Ji] = fluxua[Ie, Ji] - fluxua[I, Ji] + fluxva[I, Ji] - fluxva[I, Jss]
        return advua, advva

    # ----------------------------------------------------------------- proft
    def proft(self, fin, wfsurf, fsurf, nbc):
        """pom/solver.f:1541-1683.  nbc 2 and 4 add the short-wave penetration `rad` (:1602-1615), which
        the reference evaluates in quad precision; numpy's long double (80-bit here) stands in for it, so
        those two cases are compared with a 1-ulp-scale tolerance instead of bitwise."""
        assert nbc in (1, 2, 3, 4)
        f, c = self.f, self.c
        im, jm, kb = self.im, self.jm, self.kb
        kbm1, kbm2 = kb - 1, kb - 2
        dz, dzz, kh, dti2, umol = f["dz"], f["dzz"], f["kh"], c["dti2"], c["umol"]
        dh = f["h"] + f["etf"]
        a = np.zeros((im, jm, kb), order="F")
        cc = np.zeros((im, jm, kb), order="F")
        ee = np.zeros((im, jm, kb), order="F")
        gg = np.zeros((im, jm, kb), order="F")
        for k in range(2, kbm1 + 1):
            a[:, :, k - 2] = -dti2 * (kh[:, :, k - 1] + umol) / (dz[k - 2] * dzz[k - 2] * dh * dh)
            cc[:, :, k - 1] = -dti2 * (kh[:, :, k - 1] + umol) / (dz[k - 1] * dzz[k - 2] * dh * dh)
        ff = fin.copy(order="F")
        rad = np.zeros((im, jm, kb), order="F")
        if nbc in (2, 4):
            ntp = int(c["ntp"])
            r_ = (.58, .62, .67, .77, .78)[ntp - 1]
            ad1 = (.35, .60, 1.0, 1.5, 1.4)[ntp - 1]
            ad2 = (23., 20., 17., 14., 7.9)[ntp - 1]
            L = np.longdouble
            for k in range(1, kbm1 + 1):
                x1 = (f["z"][k - 1] * dh / ad1).astype(L)
                x2 = (f["z"][k - 1] * dh / ad2).astype(L)
                rad[:, :, k - 1] = (f["swrad"].astype(L) * (L(r_) * np.exp(x1) + L(1. - r_) * np.exp(x2))).astype(np.float64)
        if nbc == 1:
            ee[:, :, 0] = a[:, :, 0] / (a[:, :, 0] - 1.)
            gg[:, :, 0] = dti2 * wfsurf / (dz[0] * dh) - ff[:, :, 0]
            gg[:, :, 0] = gg[:, :, 0] / (a[:, :, 0] - 1.)
        elif nbc == 2:
            ee[:, :, 0] = a[:, :, 0] / (a[:, :, 0] - 1.)
            gg[:, :, 0] = dti2 * (wfsurf + rad[:, :, 0] - rad[:, :, 1]) / (dz[0] * dh) - ff[:, :, 0]
            gg[:, :, 0] = gg[:, :, 0] / (a[:, :, 0] - 1.)
        else:
            ee[:, :, 0] = 0.
            gg[:, :, 0] = fsurf
        for k in range(2, kbm2 + 1):
            K = k - 1
            gg[:, :, K] = 1. / (a[:, :, K] + cc[:, :, K] * (1. - ee[:, :, K - 1]) - 1.)
            ee[:, :, K] = a[:, :, K] * gg[:, :, K]
            gg[:, :, K] = ((cc[:, :, K] * gg[:, :, K - 1] - ff[:, :, K]
                            + dti2 * (rad[:, :, K] - rad[:, :, K + 1]) / (dh * dz[K])) * gg[:, :, K])
        K = kbm1 - 1
        ff[:, :, K] = ((cc[:, :, K] * gg[:, :, K - 1] - ff[:, :, K]
                        + dti2 * (rad[:, :, K] - rad[:, :, K + 1]) / (dh * dz[K]))
                       / (cc[:, :, K] * (1. - ee[:, :, K - 1]) - 1.))
        for k in range(2, kbm1 + 1):
            ki = kb - k
            ff[:, :, ki - 1] = ee[:, :, ki - 1] * ff[:, :, ki] + gg[:, :, ki - 1]
        return ff


# ------------------------------------------------------------------------- profq
def profq(n, uf_in, vf_in):
    """pom/solver.f:1212-1538.  n: NP with fields kq km kh t s rho q2b q2lb q2 u v wusurf wvsurf
    wubot wvbot l (previous content) and constants kappa small; uf_in, vf_in = the advq results.
    Returns dict(uf, vf, kq, km, kh, l, q2b, q2lb)."""
    f, c = n.f, n.c
    im, jm, kb = n.im, n.jm, n.kb
    imm1, jmm1, kbm1, kbm2 = im - 1, jm - 1, kb - 1, kb - 2
    a1, b1, a2, b2, c1 = 0.92, 16.6, 0.74, 10.1, 0.08
    e1, e2, sef, cbcnst, surfl, shiw = 1.8, 1.33, 1., 100., 2.e5, 0.
    dti2, umol, grav, kappa, small = c["dti2"], c["umol"], c["grav"], c["kappa"], c["small"]
    z, zz, dz, dzz, h = f["z"], f["zz"], f["dz"], f["dzz"], f["h"]
    Z3 = lambda: np.zeros((im, jm, kb), order="F")
    kq, km, kh = (f[x].copy(order="F") for x in ("kq", "km", "kh"))
    uf, vf = uf_in.copy(order="F"), vf_in.copy(order="F")
    q2b, q2lb, l = f["q2b"].copy(order="F"), f["q2lb"].copy(order="F"), f["l"].copy(order="F")
    q2, u, v, rho, t, s = f["q2"], f["u"], f["v"], f["rho"], f["t"], f["s"]
    dh = h + f["etf"]
    a, cc_, ee, gg = Z3(), Z3(), Z3(), Z3()
    for k in range(2, kbm1 + 1):
        K = k - 1
        a[:, :, K] = -dti2 * (kq[:, :, K + 1] + kq[:, :, K] + 2. * umol) * .5 / (dzz[K - 1] * dz[K] * dh * dh)
        cc_[:, :, K] = -dti2 * (kq[:, :, K - 1] + kq[:, :, K] + 2. * umol) * .5 / (dzz[K - 1] * dz[K - 1] * dh * dh)
    const1 = (16.6 ** (2. / 3.)) * sef
    utau2 = np.zeros((im, jm), order="F")
    I, J, Ie, Jn = sl(1, imm1), sl(1, jmm1), sl(1, imm1, 1), sl(1, jmm1, 1)
    wus, wvs, wub, wvb = f["wusurf"], f["wvsurf"], f["wubot"], f["wvbot"]
    utau2[I, J] = np.sqrt((.5 * (wus[I, J] + wus[Ie, J])) ** 2 + (.5 * (wvs[I, J] + wvs[I, Jn])) ** 2)
    uf[I, J, kb - 1] = np.sqrt((.5 * (wub[I, J] + wub[Ie, J])) ** 2 + (.5 * (wvb[I, J] + wvb[I, Jn])) ** 2) * const1
    sp_lit = float(np.float32(15.8)) * cbcnst                      # single-precision literals, promoted
    ee[:, :, 0] = 0.
    gg[:, :, 0] = sp_lit ** float(np.float32(2.) / np.float32(3.)) * utau2
    l0 = surfl * utau2 / grav
    cc = Z3()
    for k in range(1, kbm1 + 1):
        K = k - 1
        tp = t[:, :, K] + c["tbias"]
        sp = s[:, :, K] + c["sbias"]
        p = grav * c["rhoref"] * (-zz[K] * h) * 1.e-4
        cv = 1449.1 + .00821 * p + 4.55 * tp - .045 * (tp * tp) + 1.34 * (sp - 35.0)
        cc[:, :, K] = cv / np.sqrt((1. - .01642 * p / cv) * (1. - 0.40 * p / (cv * cv)))
    boygr, gh, prod = Z3(), Z3(), Z3()
    for k in range(2, kbm1 + 1):
        K = k - 1
        q2b[:, :, K] = np.abs(q2b[:, :, K])
        q2lb[:, :, K] = np.abs(q2lb[:, :, K])
        boygr[:, :, K] = (grav * (rho[:, :, K - 1] - rho[:, :, K]) / (dzz[K - 1] * h)
                          + (grav * grav) * 2. / (cc[:, :, K - 1] * cc[:, :, K - 1] + cc[:, :, K] * cc[:, :, K]))
    for k in range(2, kbm1 + 1):
        K = k - 1
        l[:, :, K] = np.abs(q2lb[:, :, K] / q2b[:, :, K])
        if z[K] > -0.5:
            l[:, :, K] = np.maximum(l[:, :, K], kappa * l0)
        gh[:, :, K] = (l[:, :, K] * l[:, :, K]) * boygr[:, :, K] / q2b[:, :, K]
        gh[:, :, K] = np.minimum(gh[:, :, K], .028)
    l[:, :, 0] = kappa * l0
    l[:, :, kb - 1] = 0.
    I, J = sl(2, imm1), sl(2, jmm1)
    Ie, Jn = sl(2, imm1, 1), sl(2, jmm1, 1)
    for k in range(2, kbm1 + 1):
        K = k - 1
        su = u[I, J, K] - u[I, J, K - 1] + u[Ie, J, K] - u[Ie, J, K - 1]
        sv = v[I, J, K] - v[I, J, K - 1] + v[I, Jn, K] - v[I, Jn, K - 1]
        dd = dzz[K - 1] * dh[I, J]
        prod[I, J, K] = (km[I, J, K] * .25 * sef * (su * su + sv * sv) / (dd * dd)
                         - shiw * km[I, J, K] * boygr[I, J, K])
        prod[I, J, K] = prod[I, J, K] + kh[I, J, K] * boygr[I, J, K]
    dtef = np.sqrt(np.abs(q2b)) * 1. / (b1 * l + small)
    for k in range(2, kbm1 + 1):
        K = k - 1
        gg[:, :, K] = 1. / (a[:, :, K] + cc_[:, :, K] * (1. - ee[:, :, K - 1]) - (2. * dti2 * dtef[:, :, K] + 1.))
        ee[:, :, K] = a[:, :, K] * gg[:, :, K]
        gg[:, :, K] = (-2. * dti2 * prod[:, :, K] + cc_[:, :, K] * gg[:, :, K - 1] - uf[:, :, K]) * gg[:, :, K]
    for k in range(1, kbm1 + 1):
        ki = kb - k
        uf[:, :, ki - 1] = ee[:, :, ki - 1] * uf[:, :, ki] + gg[:, :, ki - 1]
    vf[:, :, 0] = 0.
    vf[:, :, kb - 1] = 0.
    ee[:, :, 1] = 0.
    gg[:, :, 1] = -kappa * z[1] * dh * q2[:, :, 1]
    vf[:, :, kb - 2] = kappa * (1 + z[kbm1 - 1]) * dh * q2[:, :, kbm1 - 1]
    for k in range(2, kbm1 + 1):
        K = k - 1
        r = (1. / abs(z[K] - z[0]) + 1. / abs(z[K] - z[kb - 1])) * l[:, :, K] / (dh * kappa)
        dtef[:, :, K] = dtef[:, :, K] * (1. + e2 * (r * r))
    for k in range(3, kbm1 + 1):
        K = k - 1
        gg[:, :, K] = 1. / (a[:, :, K] + cc_[:, :, K] * (1. - ee[:, :, K - 1]) - (dti2 * dtef[:, :, K] + 1.))
        ee[:, :, K] = a[:, :, K] * gg[:, :, K]
        gg[:, :, K] = (dti2 * (-prod[:, :, K] * l[:, :, K] * e1) + cc_[:, :, K] * gg[:, :, K - 1] - vf[:, :, K]) * gg[:, :, K]
    for k in range(1, kb - 2 + 1):
        ki = kb - k
        vf[:, :, ki - 1] = ee[:, :, ki - 1] * vf[:, :, ki] + gg[:, :, ki - 1]
    uf[:, :, 1:kbm1] = np.abs(uf[:, :, 1:kbm1])
    vf[:, :, 1:kbm1] = np.abs(vf[:, :, 1:kbm1])
    coef4 = 18. * a1 * a1 + 9. * a1 * a2
    coef5 = 9. * a1 * a2
    coef1 = a2 * (1. - 6. * a1 / b1 * 1.)
    coef2 = 3. * a2 * b2 / 1. + 18. * a1 * a2
    coef3 = a1 * (1. - 3. * c1 - 6. * a1 / b1 * 1.)
    sh = coef1 / (1. - coef2 * gh)
    sm = coef3 + sh * coef4 * gh
    sm = sm / (1. - coef5 * gh)
    prod = l * np.sqrt(np.abs(q2))
    kq = (prod * .41 * sh + kq) * .5
    km = (prod * sm + km) * .5
    kh = (prod * sh + kh) * .5
    for x in (km, kh, kq):                       # N, S, E, W cosmetics, then the mask
        x[:, jm - 1, :] = x[:, jmm1 - 1, :]
        x[:, 0, :] = x[:, 1, :]
        x[im - 1, :, :] = x[imm1 - 1, :, :]
        x[0, :, :] = x[1, :, :]
        x *= f["fsm"][:, :, None]
    return dict(uf=uf, vf=vf, kq=kq, km=km, kh=kh, l=l, q2b=q2b, q2lb=q2lb)


# ----------------------------------------------------------------- mode_external
def mode_external(f, c, iext):
    """pom/advance.f:205-353 with bcond(1), bcond(2) (pom/bounds_forcing.f:18-83) and advave, one
    sub-domain.  f: dict of 2-D fields and edge arrays, UPDATED IN PLACE like the Fortran COMMON
    members (ua, va, uab, vab, el, elb, d, elf, uaf, vaf, etf, egf, utf, vtf, advua, advva)."""
    im, jm = f["h"].shape
    imm1, jmm1 = im - 1, jm - 1
    d, ua, va, dx, dy, h = f["d"], f["ua"], f["va"], f["dx"], f["dy"], f["h"]
    grav, alpha, dte, dte2, smoth, ramp = c["grav"], c["alpha"], c["dte"], c["dte2"], c["smoth"], c["ramp"]
    isplit = int(c["isplit"])
    fluxua = np.zeros((im, jm), order="F")
    fluxva = np.zeros((im, jm), order="F")
    I, J, Iw, Js = sl(2, im), sl(2, jm), sl(2, im, -1), sl(2, jm, -1)
    fluxua[I, J] = .25 * (d[I, J] + d[Iw, J]) * (dy[I, J] + dy[Iw, J]) * ua[I, J]
    fluxva[I, J] = .25 * (d[I, J] + d[I, Js]) * (dx[I, J] + dx[I, Js]) * va[I, J]
    elf = f["elf"]
    I, J = sl(2, imm1), sl(2, jmm1)
    elf[I, J] = f["elb"][I, J] + dte2 * (-(fluxua[sl(2, imm1, 1), J] - fluxua[I, J]
                                           + fluxva[I, sl(2, jmm1, 1)] - fluxva[I, J]) / f["art"][I, J]
                                         - f["vfluxf"][I, J])
    # bcond(1)
    elf[0, :] = elf[1, :]
    elf[im - 1, :] = elf[imm1 - 1, :]
    elf[:, 0] = elf[:, 1]
    elf[:, jm - 1] = elf[:, jmm1 - 1]
    elf[...] = elf * f["fsm"]
    if iext % int(c["ispadv"]) == 0:
        f["advua"][...], f["advva"][...] = NP(f, c).advave()
        if int(c.get("mode", 3)) == 2:
            advave_mode2(f)
    el, elb, uaf, vaf, cor, ea = f["el"], f["elb"], f["uaf"], f["vaf"], f["cor"], f["e_atmos"]
    aru, arv = f["aru"], f["arv"]
    I, J, Iw, Jn = sl(2, im), sl(2, jmm1), sl(2, im, -1), sl(2, jmm1, 1)
    r = (f["adx2d"][I, J] + f["advua"][I, J]
         - aru[I, J] * .25 * (cor[I, J] * d[I, J] * (va[I, Jn] + va[I, J])
                              + cor[Iw, J] * d[Iw, J] * (va[Iw, Jn] + va[Iw, J]))
         + .25 * grav * (dy[I, J] + dy[Iw, J]) * (d[I, J] + d[Iw, J])
         * ((1. - 2. * alpha) * (el[I, J] - el[Iw, J])
            + alpha * (elb[I, J] - elb[Iw, J] + elf[I, J] - elf[Iw, J])
            + ea[I, J] - ea[Iw, J])
         + f["drx2d"][I, J] + aru[I, J] * (f["wusurf"][I, J] - f["wubot"][I, J]))
    uaf[I, J] = (((h[I, J] + elb[I, J] + h[Iw, J] + elb[Iw, J]) * aru[I, J] * f["uab"][I, J] - 4. * dte * r)
                 / ((h[I, J] + elf[I, J] + h[Iw, J] + elf[Iw, J]) * aru[I, J]))
    I, J, Ie, Js = sl(2, imm1), sl(2, jm), sl(2, imm1, 1), sl(2, jm, -1)
    r = (f["ady2d"][I, J] + f["advva"][I, J]
         + arv[I, J] * .25 * (cor[I, J] * d[I, J] * (ua[Ie, J] + ua[I, J])
                              + cor[I, Js] * d[I, Js] * (ua[Ie, Js] + ua[I, Js]))
         + .25 * grav * (dx[I, J] + dx[I, Js]) * (d[I, J] + d[I, Js])
         * ((1. - 2. * alpha) * (el[I, J] - el[I, Js])
            + alpha * (elb[I, J] - elb[I, Js] + elf[I, J] - elf[I, Js])
            + ea[I, J] - ea[I, Js])
         + f["dry2d"][I, J] + arv[I, J] * (f["wvsurf"][I, J] - f["wvbot"][I, J]))
    vaf[I, J] = (((h[I, J] + elb[I, J] + h[I, Js] + elb[I, Js]) * arv[I, J] * f["vab"][I, J] - 4. * dte * r)
                 / ((h[I, J] + elf[I, J] + h[I, Js] + elf[I, Js]) * arv[I, J]))
    # bcond(2)
    J = sl(2, jmm1)
    uaf[1, J] = f["uabw"][J] - c["rfw"] * np.sqrt(grav / d[1, J]) * (el[1, J] - f["elw"][J])
    uaf[1, J] = ramp * uaf[1, J]
    uaf[0, J] = uaf[1, J]
    vaf[0, J] = f["vabw"][J]
    uaf[im - 1, J] = f["uabe"][J] + c["rfe"] * np.sqrt(grav / d[imm1 - 1, J]) * (el[imm1 - 1, J] - f["ele"][J])
    uaf[im - 1, J] = ramp * uaf[im - 1, J]
    vaf[im - 1, J] = f["vabe"][J]
    I = sl(2, imm1)
    vaf[I, 1] = f["vabs"][I] - c["rfs"] * np.sqrt(grav / d[I, 1]) * (el[I, 1] - f["els"][I])
    vaf[I, 1] = ramp * vaf[I, 1]
    vaf[I, 0] = vaf[I, 1]
    uaf[I, 0] = f["uabs"][I]
    vaf[I, jm - 1] = f["vabn"][I] + c["rfn"] * np.sqrt(grav / d[I, jmm1 - 1]) * (el[I, jmm1 - 1] - f["eln"][I])
    vaf[I, jm - 1] = ramp * vaf[I, jm - 1]
    uaf[I, jm - 1] = f["uabn"][I]
    uaf[...] = uaf * f["dum"]
    vaf[...] = vaf * f["dvm"]
    etf = f["etf"]
    if iext == isplit - 2:
        etf[...] = .25 * smoth * elf
    elif iext == isplit - 1:
        etf[...] = etf + .5 * (1. - .5 * smoth) * elf
    elif iext == isplit:
        etf[...] = (etf + .5 * elf) * f["fsm"]
    f["ua"][...] = ua + .5 * smoth * (f["uab"] - 2. * ua + uaf)
    f["va"][...] = va + .5 * smoth * (f["vab"] - 2. * va + vaf)
    el[...] = el + .5 * smoth * (elb - 2. * el + elf)
    elb[...] = el
    el[...] = elf
    d[...] = h + el
    f["uab"][...] = f["ua"]
    f["ua"][...] = uaf
    f["vab"][...] = f["va"]
    f["va"][...] = vaf
    if iext != isplit:
        f["egf"][...] = f["egf"] + el * c["ispi"]
        I, Iw = sl(2, im), sl(2, im, -1)
        f["utf"][I, :] = f["utf"][I, :] + f["ua"][I, :] * (d[I, :] + d[Iw, :]) * c["isp2i"]
        J, Js = sl(2, jm), sl(2, jm, -1)
        f["vtf"][:, J] = f["vtf"][:, J] + f["va"][:, J] * (d[:, J] + d[:, Js]) * c["isp2i"]


# -------------------------------------------------------------- advt2 / smol_adif
def smol_adif(f, c, xm, ym, zw, ff):
    """pom/solver.f:1880-1967, in place on xm, ym, zw, ff."""
    im, jm, kb = ff.shape
    imm1, jmm1, kbm1 = im - 1, jm - 1, kb - 1
    vmin, eps, dti2, sw, dt = 1.e-9, 1.0e-14, c["dti2"], c["sw"], f["dt"]
    ff *= f["fsm"][:, :, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        for k in range(1, kbm1 + 1):
            K = k - 1
            I, J, Iw = sl(2, im), sl(2, jmm1), sl(2, im, -1)
            x = xm[I, J, K]
            udx = np.abs(x)
            u2dt = dti2 * x * x * 2. / (f["aru"][I, J] * (dt[Iw, J] + dt[I, J]))
            mol = (ff[I, J, K] - ff[Iw, J, K]) / (ff[Iw, J, K] + ff[I, J, K] + eps)
            new = (udx - u2dt) * mol * sw
            new = np.where(np.abs(udx) < np.abs(u2dt), 0., new)
            xm[I, J, K] = np.where((ff[I, J, K] < vmin) | (ff[Iw, J, K] < vmin), 0., new)
            I, J, Js = sl(2, imm1), sl(2, jm), sl(2, jm, -1)
            y = ym[I, J, K]
            vdy = np.abs(y)
            v2dt = dti2 * y * y * 2. / (f["arv"][I, J] * (dt[I, Js] + dt[I, J]))
            mol = (ff[I, J, K] - ff[I, Js, K]) / (ff[I, Js, K] + ff[I, J, K] + eps)
            new = (vdy - v2dt) * mol * sw
            new = np.where(np.abs(vdy) < np.abs(v2dt), 0., new)
            ym[I, J, K] = np.where((ff[I, J, K] < vmin) | (ff[I, Js, K] < vmin), 0., new)
        I, J = sl(2, imm1), sl(2, jmm1)
        for k in range(2, kbm1 + 1):
            K = k - 1
            zf = zw[I, J, K]
            wdz = np.abs(zf)
            w2dt = dti2 * zf * zf / (f["dzz"][K - 1] * dt[I, J])
            mol = (ff[I, J, K - 1] - ff[I, J, K]) / (ff[I, J, K] + ff[I, J, K - 1] + eps)
            new = (wdz - w2dt) * mol * sw
            new = np.where(np.abs(wdz) < np.abs(w2dt), 0., new)
            zw[I, J, K] = np.where((ff[I, J, K] < vmin) | (ff[I, J, K - 1] < vmin), 0., new)


def advt2(f, c, fb, fq, fclim, ff_in):
    """pom/solver.f:577-731.  Returns (ff, fb) -- fb with the reference's side effects
    (level kb copy, (fb-fclim)+fclim round trip)."""
    im, jm, kb = fb.shape
    imm1, jmm1, kbm1 = im - 1, jm - 1, kb - 1
    dx, dy, dt, h, art = f["dx"], f["dy"], f["dt"], f["h"], f["art"]
    dti2 = c["dti2"]
    Z3 = lambda: np.zeros((im, jm, kb), order="F")
    xflux, yflux, xm, ym = Z3(), Z3(), Z3(), Z3()
    fb = fb.copy(order="F")
    ff = ff_in.copy(order="F")
    for k in range(1, kbm1 + 1):
        K = k - 1
        I, J, Iw = sl(2, im), sl(2, jmm1), sl(2, im, -1)
        xm[I, J, K] = 0.25 * (dy[Iw, J] + dy[I, J]) * (dt[Iw, J] + dt[I, J]) * f["u"][I, J, K]
        I, J, Js = sl(2, imm1), sl(2, jm), sl(2, jm, -1)
        ym[I, J, K] = 0.25 * (dx[I, Js] + dx[I, J]) * (dt[I, Js] + dt[I, J]) * f["v"][I, J, K]
    fb[:, :, kb - 1] = fb[:, :, kbm1 - 1]
    eta = f["etb"].copy(order="F")
    zw = f["w"].copy(order="F")
    fbmem = fb.copy(order="F")
    zflux = Z3()
    Ii, Ji = sl(2, imm1), sl(2, jmm1)
    for itera in range(1, int(c["nitera"]) + 1):
        I, J, Iw, Js = sl(2, im), sl(2, jm), sl(2, im, -1), sl(2, jm, -1)
        for k in range(1, kbm1 + 1):
            K = k - 1
            xflux[I, J, K] = 0.5 * ((xm[I, J, K] + np.abs(xm[I, J, K])) * fbmem[Iw, J, K]
                                    + (xm[I, J, K] - np.abs(xm[I, J, K])) * fbmem[I, J, K])
            yflux[I, J, K] = 0.5 * ((ym[I, J, K] + np.abs(ym[I, J, K])) * fbmem[I, Js, K]
                                    + (ym[I, J, K] - np.abs(ym[I, J, K])) * fbmem[I, J, K])
        zflux[Ii, Ji, 0] = 0.
        if itera == 1:
            zflux[Ii, Ji, 0] = f["w"][Ii, Ji, 0] * fq[Ii, Ji, 0] * art[Ii, Ji]
        zflux[Ii, Ji, kb - 1] = 0.
        for k in range(2, kbm1 + 1):
            K = k - 1
            zflux[Ii, Ji, K] = 0.5 * ((zw[Ii, Ji, K] + np.abs(zw[Ii, Ji, K])) * fbmem[Ii, Ji, K]
                                      + (zw[Ii, Ji, K] - np.abs(zw[Ii, Ji, K])) * fbmem[Ii, Ji, K - 1])
            zflux[Ii, Ji, K] = zflux[Ii, Ji, K] * art[Ii, Ji]
        for k in range(1, kbm1 + 1):
            K = k - 1
            t = (xflux[sl(2, imm1, 1), Ji, K] - xflux[Ii, Ji, K] + yflux[Ii, sl(2, jmm1, 1), K] - yflux[Ii, Ji, K]
                 + (zflux[Ii, Ji, K] - zflux[Ii, Ji, K + 1]) / f["dz"][K])
            ff[Ii, Ji, K] = ((fbmem[Ii, Ji, K] * ((h[Ii, Ji] + eta[Ii, Ji]) * art[Ii, Ji]) - dti2 * t)
                             / ((h[Ii, Ji] + f["etf"][Ii, Ji]) * art[Ii, Ji]))
        smol_adif(f, c, xm, ym, zw, ff)
        eta = f["etf"].copy(order="F")
        fbmem = ff.copy(order="F")
    fb = fb - fclim
    I, J, Iw, Js = sl(2, im), sl(2, jm), sl(2, im, -1), sl(2, jm, -1)
    aam = f["aam"]
    for k in range(1, kbm1 + 1):
        K = k - 1
        xm[I, J, K] = 0.5 * (aam[I, J, K] + aam[Iw, J, K])
        ym[I, J, K] = 0.5 * (aam[I, J, K] + aam[I, Js, K])
    for k in range(1, kbm1 + 1):
        K = k - 1
        xflux[I, J, K] = (-xm[I, J, K] * (h[I, J] + h[Iw, J]) * c["tprni"] * (fb[I, J, K] - fb[Iw, J, K]) * f["dum"][I, J]
                          * (dy[I, J] + dy[Iw, J]) * 0.5 / (dx[I, J] + dx[Iw, J]))
        yflux[I, J, K] = (-ym[I, J, K] * (h[I, J] + h[I, Js]) * c["tprni"] * (fb[I, J, K] - fb[I, Js, K]) * f["dvm"][I, J]
                          * (dx[I, J] + dx[I, Js]) * 0.5 / (dy[I, J] + dy[I, Js]))
    fb = fb + fclim
    for k in range(1, kbm1 + 1):
        K = k - 1
        ff[Ii, Ji, K] = ff[Ii, Ji, K] - dti2 * (xflux[sl(2, imm1, 1), Ji, K] - xflux[Ii, Ji, K]
                                                + yflux[Ii, sl(2, jmm1, 1), K] - yflux[Ii, Ji, K]) / (
            (h[Ii, Ji] + f["etf"][Ii, Ji]) * art[Ii, Ji])
    return ff, fb


# ------------------------------------------------------------------- baropg_mcc
def baropg_mcc(f, c, rho_in, drhox0, drhoy0):
    """pom/solver.f:943-1159 (npg=2), one sub-domain.  Returns (drhox, drhoy, rho)."""
    im, jm, kb = rho_in.shape
    imm1, jmm1, kbm1 = im - 1, jm - 1, kb - 1
    c24, c16 = float(np.float32(1.) / np.float32(24.)), float(np.float32(1.) / np.float32(16.))
    grav, zz, dzz, d, dt = c["grav"], f["zz"], f["dzz"], f["d"], f["dt"]
    rho = rho_in - f["rmean"]
    res = []
    for comp in (0, 1):
        m = f["dum"] if comp == 0 else f["dvm"]
        met = f["dy"] if comp == 0 else f["dx"]
        sh = (lambda a, n: np.roll(a, -n, axis=comp))          # sh(a, n)[i] = a[i+n] along the component axis
        lo_all = (sl(2, im), slice(None)) if comp == 0 else (slice(None), sl(2, jm))
        lo_cor = (sl(3, imm1), slice(None)) if comp == 0 else (slice(None), sl(3, jmm1))
        drho = np.zeros((im, jm, kb), order="F")
        rhou = np.zeros((im, jm, kb), order="F")
        ddx = np.zeros((im, jm), order="F")
        d4 = np.zeros((im, jm), order="F")
        for k in range(1, kbm1 + 1):
            r = rho[:, :, k - 1]
            drho[:, :, k - 1][lo_all] = ((r - sh(r, -1)) * m)[lo_all]
            rhou[:, :, k - 1][lo_all] = (0.5 * (r + sh(r, -1)) * m)[lo_all]
        ddx[lo_all] = ((d - sh(d, -1)) * m)[lo_all]
        d4[lo_all] = (.5 * (d + sh(d, -1)) * m)[lo_all]
        for k in range(1, kbm1 + 1):
            r = rho[:, :, k - 1]
            drho[:, :, k - 1][lo_cor] = (drho[:, :, k - 1] - c24 * (sh(m, 1) * (sh(r, 1) - r) - 2 * (r - sh(r, -1))
                                                                    + sh(m, -1) * (sh(r, -1) - sh(r, -2))))[lo_cor]
            rhou[:, :, k - 1][lo_cor] = (rhou[:, :, k - 1] + c16 * (sh(m, 1) * (r - sh(r, 1))
                                                                    + sh(m, -1) * (sh(r, -1) - sh(r, -2))))[lo_cor]
        ddx[lo_cor] = (ddx - c24 * (sh(m, 1) * (sh(d, 1) - d) - 2 * (d - sh(d, -1)) + sh(m, -1) * (sh(d, -1) - sh(d, -2))))[lo_cor]
        d4[lo_cor] = (d4 + c16 * (sh(m, 1) * (d - sh(d, 1)) + sh(m, -1) * (sh(d, -1) - sh(d, -2))))[lo_cor]
        I, J = sl(2, imm1), sl(2, jmm1)
        g3 = np.zeros((im, jm, kb), order="F")
        g3[I, J, 0] = grav * (-zz[0]) * d4[I, J] * drho[I, J, 0]
        for k in range(2, kbm1 + 1):
            g3[I, J, k - 1] = (g3[I, J, k - 2]
                               + grav * 0.5 * dzz[k - 2] * d4[I, J] * (drho[I, J, k - 2] + drho[I, J, k - 1])
                               + grav * 0.5 * (zz[k - 2] + zz[k - 1]) * ddx[I, J] * (rhou[I, J, k - 1] - rhou[I, J, k - 2]))
        dsum = (dt + sh(dt, -1))
        msum = (met + sh(met, -1))
        for k in range(1, kbm1 + 1):
            g3[I, J, k - 1] = .25 * dsum[I, J] * g3[I, J, k - 1] * m[I, J] * msum[I, J]
        res.append(g3)
    drhox, drhoy = drhox0.copy(order="F"), drhoy0.copy(order="F")
    I, J = sl(2, imm1), sl(2, jmm1)
    drhox[I, J, :kbm1] = res[0][I, J, :kbm1]
    drhoy[I, J, :kbm1] = res[1][I, J, :kbm1]
    drhox[I, J, :] = c["ramp"] * drhox[I, J, :]
    drhoy[I, J, :] = c["ramp"] * drhoy[I, J, :]
    return drhox, drhoy, rho + f["rmean"]


# ---------------------------------------------------------------------------------------------
# Second batch: the momentum path.  R(x, I, J, di, dj) is x(i+di, j+dj, 1:kbm1) for i in I=(a,b),
# j in J=(c,d) (Fortran 1-based inclusive ranges); 2-D operands get a trailing axis for broadcast.
def _R(x, I, J, di=0, dj=0, kbm1=None):
    a = x[I[0] - 1 + di:I[1] + di, J[0] - 1 + dj:J[1] + dj]
    if x.ndim == 2:
        return a[:, :, None]
    return a[:, :, :kbm1]


def advct(f, c):
    """pom/solver.f:201-409 -> (advx, advy); one sub-domain (n_west = n_south = -1)."""
    u, v, ub, vb, aam, dt, dx, dy, aru, arv = (f[n] for n in "u v ub vb aam dt dx dy aru arv".split())
    im, jm, kb = u.shape
    imm1, jmm1, kbm1 = im - 1, jm - 1, kb - 1
    Z = lambda: np.zeros((im, jm, kb), order="F")

    def R(x, I, J, di=0, dj=0):
        return _R(x, I, J, di, dj, kbm1)

    def W(x, I, J):       # assignable view x(I, J, 1:kbm1)
        return x[I[0] - 1:I[1], J[0] - 1:J[1], :kbm1]

    curv, advx, xflux, yflux = Z(), Z(), Z(), Z()
    I, J = (2, imm1), (2, jmm1)
    W(curv, I, J)[...] = (.25 * ((R(v, I, J, 0, 1) + R(v, I, J)) * (R(dy, I, J, 1, 0) - R(dy, I, J, -1, 0))
                                 - (R(u, I, J, 1, 0) + R(u, I, J)) * (R(dx, I, J, 0, 1) - R(dx, I, J, 0, -1)))
                          / (R(dx, I, J) * R(dy, I, J)))
    # x-component: advective fluxes
    I, J = (2, imm1), (1, jm)
    W(xflux, I, J)[...] = (.125 * ((R(dt, I, J, 1, 0) + R(dt, I, J)) * R(u, I, J, 1, 0)
                                   + (R(dt, I, J) + R(dt, I, J, -1, 0)) * R(u, I, J))
                           * (R(u, I, J, 1, 0) + R(u, I, J)))
    I, J = (2, im), (2, jm)
    W(yflux, I, J)[...] = (.125 * ((R(dt, I, J) + R(dt, I, J, 0, -1)) * R(v, I, J)
                                   + (R(dt, I, J, -1, 0) + R(dt, I, J, -1, -1)) * R(v, I, J, -1, 0))
                           * (R(u, I, J) + R(u, I, J, 0, -1)))
    # diffusive fluxes
    I, J = (2, imm1), (2, jm)
    W(xflux, I, J)[...] = (R(xflux, I, J)
                           - R(dt, I, J) * R(aam, I, J) * 2. * (R(ub, I, J, 1, 0) - R(ub, I, J)) / R(dx, I, J))
    dtaam = (.25 * (R(dt, I, J) + R(dt, I, J, -1, 0) + R(dt, I, J, 0, -1) + R(dt, I, J, -1, -1))
             * (R(aam, I, J) + R(aam, I, J, -1, 0) + R(aam, I, J, 0, -1) + R(aam, I, J, -1, -1)))
    W(yflux, I, J)[...] = (R(yflux, I, J)
                           - dtaam * ((R(ub, I, J) - R(ub, I, J, 0, -1))
                                      / (R(dy, I, J) + R(dy, I, J, -1, 0) + R(dy, I, J, 0, -1) + R(dy, I, J, -1, -1))
                                      + (R(vb, I, J) - R(vb, I, J, -1, 0))
                                      / (R(dx, I, J) + R(dx, I, J, -1, 0) + R(dx, I, J, 0, -1) + R(dx, I, J, -1, -1))))
    W(xflux, I, J)[...] = R(dy, I, J) * R(xflux, I, J)
    W(yflux, I, J)[...] = (.25 * (R(dx, I, J) + R(dx, I, J, -1, 0) + R(dx, I, J, 0, -1) + R(dx, I, J, -1, -1))
                           * R(yflux, I, J))
    I, J = (2, imm1), (2, jmm1)
    W(advx, I, J)[...] = R(xflux, I, J) - R(xflux, I, J, -1, 0) + R(yflux, I, J, 0, 1) - R(yflux, I, J)
    I = (3, imm1)                                   # n_west == -1
    W(advx, I, J)[...] = (R(advx, I, J)
                          - R(aru, I, J) * .25
                          * (R(curv, I, J) * R(dt, I, J) * (R(v, I, J, 0, 1) + R(v, I, J))
                             + R(curv, I, J, -1, 0) * R(dt, I, J, -1, 0) * (R(v, I, J, -1, 1) + R(v, I, J, -1, 0))))
    # y-component
    advy, xflux, yflux = Z(), Z(), Z()
    I, J = (2, im), (2, jm)
    W(xflux, I, J)[...] = (.125 * ((R(dt, I, J) + R(dt, I, J, -1, 0)) * R(u, I, J)
                                   + (R(dt, I, J, 0, -1) + R(dt, I, J, -1, -1)) * R(u, I, J, 0, -1))
                           * (R(v, I, J) + R(v, I, J, -1, 0)))
    I, J = (1, im), (2, jmm1)
    W(yflux, I, J)[...] = (.125 * ((R(dt, I, J, 0, 1) + R(dt, I, J)) * R(v, I, J, 0, 1)
                                   + (R(dt, I, J) + R(dt, I, J, 0, -1)) * R(v, I, J))
                           * (R(v, I, J, 0, 1) + R(v, I, J)))
    I, J = (2, im), (2, jmm1)
    dtaam = (.25 * (R(dt, I, J) + R(dt, I, J, -1, 0) + R(dt, I, J, 0, -1) + R(dt, I, J, -1, -1))
             * (R(aam, I, J) + R(aam, I, J, -1, 0) + R(aam, I, J, 0, -1) + R(aam, I, J, -1, -1)))
    W(xflux, I, J)[...] = (R(xflux, I, J)
                           - dtaam * ((R(ub, I, J) - R(ub, I, J, 0, -1))
                                      / (R(dy, I, J) + R(dy, I, J, -1, 0) + R(dy, I, J, 0, -1) + R(dy, I, J, -1, -1))
                                      + (R(vb, I, J) - R(vb, I, J, -1, 0))
                                      / (R(dx, I, J) + R(dx, I, J, -1, 0) + R(dx, I, J, 0, -1) + R(dx, I, J, -1, -1))))
    W(yflux, I, J)[...] = (R(yflux, I, J)
                           - R(dt, I, J) * R(aam, I, J) * 2. * (R(vb, I, J, 0, 1) - R(vb, I, J)) / R(dy, I, J))
    W(xflux, I, J)[...] = (.25 * (R(dy, I, J) + R(dy, I, J, -1, 0) + R(dy, I, J, 0, -1) + R(dy, I, J, -1, -1))
                           * R(xflux, I, J))
    W(yflux, I, J)[...] = R(dx, I, J) * R(yflux, I, J)
    I, J = (2, imm1), (2, jmm1)
    W(advy, I, J)[...] = R(xflux, I, J, 1, 0) - R(xflux, I, J) + R(yflux, I, J) - R(yflux, I, J, 0, -1)
    J = (3, jmm1)                                   # n_south == -1
    W(advy, I, J)[...] = (R(advy, I, J)
                          + R(arv, I, J) * .25
                          * (R(curv, I, J) * R(dt, I, J) * (R(u, I, J, 1, 0) + R(u, I, J))
                             + R(curv, I, J, 0, -1) * R(dt, I, J, 0, -1) * (R(u, I, J, 1, -1) + R(u, I, J, 0, -1))))
    return advx, advy


def smagorinsky(f, c):
    """lateral_viscosity's aam (pom/advance.f:122-136); boundary cells keep f['aam']."""
    u, v, dx, dy = f["u"], f["v"], f["dx"], f["dy"]
    im, jm, kb = u.shape
    kbm1 = kb - 1
    I, J = (2, im - 1), (2, jm - 1)
    R = lambda x, di=0, dj=0: _R(x, I, J, di, dj, kbm1)
    aam = f["aam"].copy(order="F")
    aam[1:im - 1, 1:jm - 1, :kbm1] = (
        c["horcon"] * R(dx) * R(dy)
        * np.sqrt(((R(u, 1, 0) - R(u)) / R(dx)) ** 2
                  + ((R(v, 0, 1) - R(v)) / R(dy)) ** 2
                  + .5 * (.25 * (R(u, 0, 1) + R(u, 1, 1) - R(u, 0, -1) - R(u, 1, -1)) / R(dy)
                          + .25 * (R(v, 1, 0) + R(v, 1, 1) - R(v, -1, 0) - R(v, -1, 1)) / R(dx)) ** 2))
    return aam


def advu(f, c):
    """pom/solver.f:734-788 -> uf."""
    u, v, w, ub, advx, drhox = (f[n] for n in "u v w ub advx drhox".split())
    aru, cor, dt, dy, egf, egb, ea, h, etb, etf, dz = (f[n] for n in "aru cor dt dy egf egb e_atmos h etb etf dz".split())
    im, jm, kb = u.shape
    imm1, jmm1, kbm1 = im - 1, jm - 1, kb - 1
    uf = np.zeros((im, jm, kb), order="F")
    # k=2..kbm1, j=1..jm, i=2..im
    uf[1:im, :, 1:kbm1] = .25 * (w[1:im, :, 1:kbm1] + w[0:im - 1, :, 1:kbm1]) * (u[1:im, :, 1:kbm1] + u[1:im, :, 0:kbm1 - 1])
    I, J = (2, imm1), (2, jmm1)
    R = lambda x, di=0, dj=0: _R(x, I, J, di, dj, kbm1)
    ufk1 = uf[1:imm1, 1:jmm1, 1:kb]                 # uf(i,j,k+1), k=1..kbm1
    new = (R(advx)
           + (R(uf) - ufk1) * R(aru) / dz[None, None, :kbm1]
           - R(aru) * .25 * (R(cor) * R(dt) * (R(v, 0, 1) + R(v)) + R(cor, -1, 0) * R(dt, -1, 0) * (R(v, -1, 1) + R(v, -1, 0)))
           + c["grav"] * .125 * (R(dt) + R(dt, -1, 0))
           * (R(egf) - R(egf, -1, 0) + R(egb) - R(egb, -1, 0) + (R(ea) - R(ea, -1, 0)) * 2.)
           * (R(dy) + R(dy, -1, 0))
           + R(drhox))
    uf[1:imm1, 1:jmm1, :kbm1] = new
    uf[1:imm1, 1:jmm1, :kbm1] = (((R(h) + R(etb) + R(h, -1, 0) + R(etb, -1, 0)) * R(aru) * R(ub) - 2. * c["dti2"] * R(uf))
                                 / ((R(h) + R(etf) + R(h, -1, 0) + R(etf, -1, 0)) * R(aru)))
    return uf


def advv(f, c):
    """pom/solver.f:791-845 -> vf."""
    u, v, w, vb, advy, drhoy = (f[n] for n in "u v w vb advy drhoy".split())
    arv, cor, dt, dx, egf, egb, ea, h, etb, etf, dz = (f[n] for n in "arv cor dt dx egf egb e_atmos h etb etf dz".split())
    im, jm, kb = u.shape
    imm1, jmm1, kbm1 = im - 1, jm - 1, kb - 1
    vf = np.zeros((im, jm, kb), order="F")
    vf[:, 1:jm, 1:kbm1] = .25 * (w[:, 1:jm, 1:kbm1] + w[:, 0:jm - 1, 1:kbm1]) * (v[:, 1:jm, 1:kbm1] + v[:, 1:jm, 0:kbm1 - 1])
    I, J = (2, imm1), (2, jmm1)
    R = lambda x, di=0, dj=0: _R(x, I, J, di, dj, kbm1)
    vfk1 = vf[1:imm1, 1:jmm1, 1:kb]
    new = (R(advy)
           + (R(vf) - vfk1) * R(arv) / dz[None, None, :kbm1]
           + R(arv) * .25 * (R(cor) * R(dt) * (R(u, 1, 0) + R(u)) + R(cor, 0, -1) * R(dt, 0, -1) * (R(u, 1, -1) + R(u, 0, -1)))
           + c["grav"] * .125 * (R(dt) + R(dt, 0, -1))
           * (R(egf) - R(egf, 0, -1) + R(egb) - R(egb, 0, -1) + (R(ea) - R(ea, 0, -1)) * 2.)
           * (R(dx) + R(dx, 0, -1))
           + R(drhoy))
    vf[1:imm1, 1:jmm1, :kbm1] = new
    vf[1:imm1, 1:jmm1, :kbm1] = (((R(h) + R(etb) + R(h, 0, -1) + R(etb, 0, -1)) * R(arv) * R(vb) - 2. * c["dti2"] * R(vf))
                                 / ((R(h) + R(etf) + R(h, 0, -1) + R(etf, 0, -1)) * R(arv)))
    return vf


def _prof_uv(f, c, xf_in, km, wsurf, mask, tps, di, dj):
    """Shared skeleton of profu / profv (pom/solver.f:1686-1780, 1783-1877): (i-di, j-dj) is the
    neighbour that is averaged with; tps (the drag coefficient * speed) is computed by the caller."""
    h, etf, dz, dzz = f["h"], f["etf"], f["dz"], f["dzz"]
    im, jm, kb = xf_in.shape
    imm1, jmm1, kbm1, kbm2 = im - 1, jm - 1, kb - 1, kb - 2
    dti2, umol = c["dti2"], c["umol"]
    xf = xf_in.copy(order="F")
    dh = np.ones((im, jm), order="F")
    if di:
        dh[1:, 1:] = (h[1:, 1:] + etf[1:, 1:] + h[:-1, 1:] + etf[:-1, 1:]) * .5
    else:
        dh[1:, 1:] = .5 * (h[1:, 1:] + etf[1:, 1:] + h[1:, :-1] + etf[1:, :-1])
    Z = lambda: np.zeros((im, jm, kb), order="F")
    a, cc, ee, gg = Z(), Z(), Z(), Z()
    if di:
        cc[1:, 1:, :] = (km[1:, 1:, :] + km[:-1, 1:, :]) * .5
    else:
        cc[1:, 1:, :] = (km[1:, 1:, :] + km[1:, :-1, :]) * .5
    for k in range(2, kbm1 + 1):
        a[:, :, k - 2] = -dti2 * (cc[:, :, k - 1] + umol) / (dz[k - 2] * dzz[k - 2] * dh * dh)
        cc[:, :, k - 1] = -dti2 * (cc[:, :, k - 1] + umol) / (dz[k - 1] * dzz[k - 2] * dh * dh)
    ee[:, :, 0] = a[:, :, 0] / (a[:, :, 0] - 1.)
    gg[:, :, 0] = (-dti2 * wsurf / (-dz[0] * dh) - xf[:, :, 0]) / (a[:, :, 0] - 1.)
    for k in range(2, kbm2 + 1):
        gg[:, :, k - 1] = 1. / (a[:, :, k - 1] + cc[:, :, k - 1] * (1. - ee[:, :, k - 2]) - 1.)
        ee[:, :, k - 1] = a[:, :, k - 1] * gg[:, :, k - 1]
        gg[:, :, k - 1] = (cc[:, :, k - 1] * gg[:, :, k - 2] - xf[:, :, k - 1]) * gg[:, :, k - 1]
    In = (slice(1, imm1), slice(1, jmm1))
    xf[In + (kbm1 - 1,)] = ((cc[In + (kbm1 - 1,)] * gg[In + (kbm2 - 1,)] - xf[In + (kbm1 - 1,)])
                            / (tps * dti2 / (-dz[kbm1 - 1] * dh[In]) - 1. - (ee[In + (kbm2 - 1,)] - 1.) * cc[In + (kbm1 - 1,)]))
    xf[In + (kbm1 - 1,)] = xf[In + (kbm1 - 1,)] * mask[In]
    for k in range(2, kbm1 + 1):
        ki = kb - k
        xf[In + (ki - 1,)] = (ee[In + (ki - 1,)] * xf[In + (ki,)] + gg[In + (ki - 1,)]) * mask[In]
    wbot = -tps * xf[In + (kbm1 - 1,)]
    return xf, wbot


def profu(f, c, uf_in):
    """pom/solver.f:1686-1780 -> (uf, wubot interior (2:imm1,2:jmm1))."""
    cbc, ub, vb = f["cbc"], f["ub"], f["vb"]
    im, jm, kb = uf_in.shape
    k = kb - 2          # level kbm1, 0-based
    I, J = (2, im - 1), (2, jm - 1)
    r = lambda x, di=0, dj=0: x[I[0] - 1 + di:I[1] + di, J[0] - 1 + dj:J[1] + dj]
    ubk, vbk = ub[:, :, k], vb[:, :, k]
    tps = (0.5 * (r(cbc) + r(cbc, -1, 0))
           * np.sqrt(r(ubk) ** 2 + (.25 * (r(vbk) + r(vbk, 0, 1) + r(vbk, -1, 0) + r(vbk, -1, 1))) ** 2))
    return _prof_uv(f, c, uf_in, f["km"], f["wusurf"], f["dum"], tps, 1, 0)


def profv(f, c, vf_in):
    """pom/solver.f:1783-1877 -> (vf, wvbot interior)."""
    cbc, ub, vb = f["cbc"], f["ub"], f["vb"]
    im, jm, kb = vf_in.shape
    k = kb - 2
    I, J = (2, im - 1), (2, jm - 1)
    r = lambda x, di=0, dj=0: x[I[0] - 1 + di:I[1] + di, J[0] - 1 + dj:J[1] + dj]
    ubk, vbk = ub[:, :, k], vb[:, :, k]
    tps = (0.5 * (r(cbc) + r(cbc, 0, -1))
           * np.sqrt((.25 * (r(ubk) + r(ubk, 1, 0) + r(ubk, 0, -1) + r(ubk, 1, -1))) ** 2 + r(vbk) ** 2))
    return _prof_uv(f, c, vf_in, f["km"], f["wvsurf"], f["dvm"], tps, 0, 1)


def realvertvl(f, c):
    """pom/solver.f:2024-2066 -> wr (one sub-domain: all four edge copies)."""
    w, u, v, dt, et, etf, etb, dx, dy, fsm, zz = (f[n] for n in "w u v dt et etf etb dx dy fsm zz".split())
    im, jm, kb = w.shape
    imm1, jmm1, kbm1 = im - 1, jm - 1, kb - 1
    wr = np.zeros((im, jm, kb), order="F")
    I, J = (2, imm1), (2, jmm1)
    r = lambda x, di=0, dj=0: x[I[0] - 1 + di:I[1] + di, J[0] - 1 + dj:J[1] + dj]
    two = np.float64(np.float32(2.0)); half = np.float64(np.float32(0.5)); one = np.float64(np.float32(1.0))
    for k in range(1, kbm1 + 1):
        tps = zz[k - 1] * dt + et
        dxr = two / (r(dx, 1, 0) + r(dx)); dxl = two / (r(dx) + r(dx, -1, 0))
        dyt = two / (r(dy, 0, 1) + r(dy)); dyb = two / (r(dy) + r(dy, 0, -1))
        uk, vk = u[:, :, k - 1], v[:, :, k - 1]
        wr[1:imm1, 1:jmm1, k - 1] = (half * (r(w[:, :, k - 1]) + r(w[:, :, k])) + half
                                     * (r(uk, 1, 0) * (r(tps, 1, 0) - r(tps)) * dxr
                                        + r(uk) * (r(tps) - r(tps, -1, 0)) * dxl
                                        + r(vk, 0, 1) * (r(tps, 0, 1) - r(tps)) * dyt
                                        + r(vk) * (r(tps) - r(tps, 0, -1)) * dyb)
                                     + (one + zz[k - 1]) * (r(etf) - r(etb)) / c["dti2"])
    wr[:, 0, :] = wr[:, 1, :]
    wr[:, jm - 1, :] = wr[:, jmm1 - 1, :]
    wr[0, :, :] = wr[1, :, :]
    wr[im - 1, :, :] = wr[imm1 - 1, :, :]
    for k in range(1, kbm1 + 1):
        wr[:, :, k - 1] = fsm * wr[:, :, k - 1]
    return wr


# ---------------------------------------------------------------------------------------------
# Third batch: the glue of advance.f and the open-boundary routines, so that one whole internal
# step (pom/advance.f:21-32; mode=3, nadv=2, npg=1, nbct=nbcs=1 or 3, no restoring) exists a second
# time.  f is a dict of ALL fields, updated in place like the COMMON blocks.
def _upstream_edges(f, c, A, B, uf, vf, ea, eb, K, vertical):
    """The four-edge upstream advection shared in FORM (not in code) by bcond(4) and bcond(6) of the
    reference; written once here with the edge data passed in.  A, B: the two advected fields;
    ea(side, k) / eb(side, k): the prescribed outside values; K: number of levels."""
    u, v, w, dx, dy, dt, zz, dti = f["u"], f["v"], f["w"], f["dx"], f["dy"], f["dt"], f["zz"], c["dti"]
    im, jm, kb = u.shape
    imm1, jmm1, kbm1 = im - 1, jm - 1, kb - 1
    for k in range(1, K + 1):
        z = k - 1
        mid = vertical and k != 1 and k != kbm1
        # east (i = im)
        u1 = 2. * u[im - 1, :, z] * dti / (dx[im - 1, :] + dx[imm1 - 1, :])
        for X, xf, e in ((A, uf, ea), (B, vf, eb)):
            inflow = X[im - 1, :, z] - u1 * (e("e", z) - X[im - 1, :, z])
            out = X[im - 1, :, z] - u1 * (X[im - 1, :, z] - X[imm1 - 1, :, z])
            if mid:
                wm = .5 * (w[imm1 - 1, :, z] + w[imm1 - 1, :, z + 1]) * dti / ((zz[z - 1] - zz[z + 1]) * dt[imm1 - 1, :])
                out = out - wm * (X[imm1 - 1, :, z - 1] - X[imm1 - 1, :, z + 1])
            xf[im - 1, :, z] = np.where(u1 <= 0., inflow, out)
        # west (i = 1, velocity at i = 2)
        u1 = 2. * u[1, :, z] * dti / (dx[0, :] + dx[1, :])
        for X, xf, e in ((A, uf, ea), (B, vf, eb)):
            inflow = X[0, :, z] - u1 * (X[0, :, z] - e("w", z))
            out = X[0, :, z] - u1 * (X[1, :, z] - X[0, :, z])
            if mid:
                wm = .5 * (w[1, :, z] + w[1, :, z + 1]) * dti / ((zz[z - 1] - zz[z + 1]) * dt[1, :])
                out = out - wm * (X[1, :, z - 1] - X[1, :, z + 1])
            xf[0, :, z] = np.where(u1 >= 0., inflow, out)
        # south (j = 1, velocity at j = 2)
        u1 = 2. * v[:, 1, z] * dti / (dy[:, 0] + dy[:, 1])
        for X, xf, e in ((A, uf, ea), (B, vf, eb)):
            inflow = X[:, 0, z] - u1 * (X[:, 0, z] - e("s", z))
            out = X[:, 0, z] - u1 * (X[:, 1, z] - X[:, 0, z])
            if mid:
                wm = .5 * (w[:, 1, z] + w[:, 1, z + 1]) * dti / ((zz[z - 1] - zz[z + 1]) * dt[:, 1])
                out = out - wm * (X[:, 1, z - 1] - X[:, 1, z + 1])
            xf[:, 0, z] = np.where(u1 >= 0., inflow, out)
        # north (j = jm)
        u1 = 2. * v[:, jm - 1, z] * dti / (dy[:, jm - 1] + dy[:, jmm1 - 1])
        for X, xf, e in ((A, uf, ea), (B, vf, eb)):
            inflow = X[:, jm - 1, z] - u1 * (e("n", z) - X[:, jm - 1, z])
            out = X[:, jm - 1, z] - u1 * (X[:, jm - 1, z] - X[:, jmm1 - 1, z])
            if mid:
                wm = .5 * (w[:, jmm1 - 1, z] + w[:, jmm1 - 1, z + 1]) * dti / ((zz[z - 1] - zz[z + 1]) * dt[:, jmm1 - 1])
                out = out - wm * (X[:, jmm1 - 1, z - 1] - X[:, jmm1 - 1, z + 1])
            xf[:, jm - 1, z] = np.where(u1 <= 0., inflow, out)


def bcond4(f, c):
    """pom/bounds_forcing.f:151-242: T, S open boundaries on uf, vf, then the fsm mask (k=1..kbm1)."""
    kbm1 = f["u"].shape[2] - 1
    te = {"e": f["tbe"], "w": f["tbw"], "s": f["tbs"], "n": f["tbn"]}
    se = {"e": f["sbe"], "w": f["sbw"], "s": f["sbs"], "n": f["sbn"]}
    _upstream_edges(f, c, f["t"], f["s"], f["uf"], f["vf"], lambda s, z: te[s][:, z], lambda s, z: se[s][:, z], kbm1, True)
    f["uf"][:, :, :kbm1] = f["uf"][:, :, :kbm1] * f["fsm"][:, :, None]
    f["vf"][:, :, :kbm1] = f["vf"][:, :, :kbm1] * f["fsm"][:, :, None]


def bcond6(f, c):
    """pom/bounds_forcing.f:257-324: q2, q2l open boundaries on uf, vf (k=1..kb; the reference does
    west before east here, which only matters for im=1), then `*fsm + 1e-10`."""
    kb = f["u"].shape[2]
    small = c["small"]
    _upstream_edges(f, c, f["q2"], f["q2l"], f["uf"], f["vf"], lambda s, z: small, lambda s, z: small, kb, False)
    f["uf"][...] = f["uf"] * f["fsm"][:, :, None] + 1.e-10
    f["vf"][...] = f["vf"] * f["fsm"][:, :, None] + 1.e-10


def bcondorl3(f, c):
    """pom/bounds_forcing.f:418-487: Orlanski radiation on uf, vf, then the dum / dvm masks."""
    u, v, ub, vb, uf, vf = (f[n] for n in "u v ub vb uf vf".split())
    im, jm, kb = u.shape
    kbm1 = kb - 1
    J = slice(1, jm - 1)      # j = 2..jmm1
    I = slice(1, im - 1)

    def rad(xf1, xb1, x2, xb0, x1):
        denom = xf1 + xb1 - 2. * x2
        denom = np.where(denom == 0., 0.01, denom)
        cl = (xb1 - xf1) / denom
        cl = np.where(cl > 1., 1., cl)
        cl = np.where(cl < 0., 0., cl)
        return (xb0 * (1. - cl) + 2. * cl * x1) / (1. + cl)

    K = slice(0, kbm1)
    # east
    uf[im - 1, J, K] = rad(uf[im - 2, J, K], ub[im - 2, J, K], u[im - 3, J, K], ub[im - 1, J, K], u[im - 2, J, K])
    vf[im - 1, J, K] = 0.
    # west
    uf[1, J, K] = rad(uf[2, J, K], ub[2, J, K], u[3, J, K], ub[1, J, K], u[2, J, K])
    uf[0, J, K] = uf[1, J, K]
    vf[0, J, K] = 0.
    # south
    vf[I, 1, K] = rad(vf[I, 2, K], vb[I, 2, K], v[I, 3, K], vb[I, 1, K], v[I, 2, K])
    vf[I, 0, K] = vf[I, 1, K]
    uf[I, 0, K] = 0.
    # north
    vf[I, jm - 1, K] = rad(vf[I, jm - 2, K], vb[I, jm - 2, K], v[I, jm - 3, K], vb[I, jm - 1, K], v[I, jm - 2, K])
    uf[I, jm - 1, K] = 0.
    uf[:, :, K] = uf[:, :, K] * f["dum"][:, :, None]
    vf[:, :, K] = vf[:, :, K] * f["dvm"][:, :, None]


def lateral_viscosity(f, c):
    """pom/advance.f:96-141 (mode /= 2, npg = 1)."""
    f["advx"], f["advy"] = advct(f, c)
    f["drhox"], f["drhoy"], f["rho"] = NP(f, c).baropg(f["rho"], f["drhox"], f["drhoy"])
    f["aam"] = smagorinsky(f, c)


def mode_interaction(f, c):
    """pom/advance.f:144-202."""
    im, jm, kb = f["u"].shape
    dz = f["dz"]
    for n2, n3 in (("adx2d", "advx"), ("ady2d", "advy"), ("drx2d", "drhox"), ("dry2d", "drhoy"), ("aam2d", "aam")):
        acc = np.zeros((im, jm), order="F")
        for k in range(kb - 1):
            acc = acc + f[n3][:, :, k] * dz[k]
        f[n2] = acc
    f["advua"], f["advva"] = NP(f, c).advave()
    f["adx2d"] = f["adx2d"] - f["advua"]
    f["ady2d"] = f["ady2d"] - f["advva"]
    f["egf"] = f["el"] * c["ispi"]
    d, ua, va = f["d"], f["ua"], f["va"]
    f["utf"][1:, :] = ua[1:, :] * (d[1:, :] + d[:-1, :]) * c["isp2i"]
    f["vtf"][:, 1:] = va[:, 1:] * (d[:, 1:] + d[:, :-1]) * c["isp2i"]


def mode_internal(f, c, first_cold_step=False):
    """pom/advance.f:356-537 for mode=3, nadv in {1,2}, nitera from c, nbct/nbcs in {1,3}, lrestore off."""
    im, jm, kb = f["u"].shape
    kbm1 = kb - 1
    dz, dt, smoth = f["dz"], f["dt"], c["smoth"]
    if not first_cold_step:
        u, v = f["u"], f["v"]
        tps = np.zeros((im, jm), order="F")
        for k in range(kbm1):
            tps = tps + u[:, :, k] * dz[k]
        for k in range(kbm1):
            u[1:, :, k] = (u[1:, :, k] - tps[1:, :]) + (f["utb"][1:, :] + f["utf"][1:, :]) / (dt[1:, :] + dt[:-1, :])
        tps = np.zeros((im, jm), order="F")
        for k in range(kbm1):
            tps = tps + v[:, :, k] * dz[k]
        for k in range(kbm1):
            v[:, 1:, k] = (v[:, 1:, k] - tps[:, 1:]) + (f["vtb"][:, 1:] + f["vtf"][:, 1:]) / (dt[:, 1:] + dt[:, :-1])
        n = NP(f, c)
        w = n.vertvl(f["w"])
        w[:, :, :kbm1] = w[:, :, :kbm1] * f["fsm"][:, :, None]          # bcondorl(5)
        f["w"] = w
        z3 = np.zeros((im, jm, kb), order="F")
        uf = n.advq(f["q2b"], f["q2"], z3)
        vf = n.advq(f["q2lb"], f["q2l"], z3)
        out = profq(n, uf, vf)
        for k_, a in out.items():
            f[k_] = a
        bcond6(f, c)
        q2 = f["q2"] + .5 * smoth * (f["uf"] + f["q2b"] - 2. * f["q2"])
        q2l = f["q2l"] + .5 * smoth * (f["vf"] + f["q2lb"] - 2. * f["q2l"])
        f["q2b"], f["q2"] = q2, f["uf"].copy(order="F")
        f["q2lb"], f["q2l"] = q2l, f["vf"].copy(order="F")
        # tracers
        if int(c["nadv"]) == 1:
            f["uf"], f["tb"], f["t"] = advt1(f, c, f["tb"], f["t"], f["tclim"], f["uf"])
            f["vf"], f["sb"], f["s"] = advt1(f, c, f["sb"], f["s"], f["sclim"], f["vf"])
        else:
            f["uf"], f["tb"] = advt2(f, c, f["tb"], f["t"], f["tclim"], f["uf"])
            f["vf"], f["sb"] = advt2(f, c, f["sb"], f["s"], f["sclim"], f["vf"])
        n = NP(f, c)
        f["uf"] = n.proft(f["uf"], f["wtsurf"], f["tsurf"], int(c["nbct"]))
        f["vf"] = n.proft(f["vf"], f["wssurf"], f["ssurf"], int(c["nbcs"]))
        bcond4(f, c)
        t = f["t"] + .5 * smoth * (f["uf"] + f["tb"] - 2. * f["t"])
        s = f["s"] + .5 * smoth * (f["vf"] + f["sb"] - 2. * f["s"])
        f["tb"], f["t"] = t, f["uf"].copy(order="F")
        f["sb"], f["s"] = s, f["vf"].copy(order="F")
        if int(c.get("lrestore", 0)):                                    # bounds_forcing.f:1083-1109
            trst = np.float64(np.float32(30.))
            ntime = int(c["time"] / trst)
            fnew = c["time"] / trst - ntime
            fold = 1. - fnew
            K = slice(0, kbm1)
            trstr = fold * f["trstrb"][:, :, K] + fnew * f["trstrf"][:, :, K]
            srstr = fold * f["srstrb"][:, :, K] + fnew * f["srstrf"][:, :, K]
            tau = fold * f["taurstrb"][:, :, K] + fnew * f["taurstrf"][:, :, K]
            g2 = 2. * c["dti"] / 86400.
            for nme, tgt in (("t", trstr), ("tb", trstr), ("s", srstr), ("sb", srstr)):
                x = f[nme][:, :, K]
                f[nme][:, :, K] = x + g2 * tau * (tgt - x)
        for nme in ("t", "tb", "s", "sb"):                               # restore_interior's mask (:1113-1118)
            f[nme][:, :, :kbm1] = f[nme][:, :, :kbm1] * f["fsm"][:, :, None]
        f["rho"] = NP(f, c).dens(f["s"], f["t"])
        # momentum
        uf_, vf_ = advu(f, c), advv(f, c)
        f["uf"], wub = profu(f, c, uf_)
        f["wubot"][1:-1, 1:-1] = wub
        f["vf"], wvb = profv(f, c, vf_)
        f["wvbot"][1:-1, 1:-1] = wvb
        bcondorl3(f, c)
        for xn, xb, xf in (("u", "ub", "uf"), ("v", "vb", "vf")):
            x, b, ff = f[xn], f[xb], f[xf]
            tps = np.zeros((im, jm), order="F")
            for k in range(kbm1):
                tps = tps + (ff[:, :, k] + b[:, :, k] - 2. * x[:, :, k]) * dz[k]
            for k in range(kbm1):
                x[:, :, k] = x[:, :, k] + .5 * smoth * (ff[:, :, k] + b[:, :, k] - 2. * x[:, :, k] - tps)
            f[xb], f[xn] = x, ff.copy(order="F")
    f["egb"] = f["egf"].copy(order="F")
    f["etb"] = f["et"].copy(order="F")
    f["et"] = f["etf"].copy(order="F")
    f["dt"] = f["h"] + f["et"]
    f["utb"] = f["utf"].copy(order="F")
    f["vtb"] = f["vtf"].copy(order="F")
    f["vfluxb"] = f["vfluxf"].copy(order="F")
    f["wr"] = realvertvl(f, c)


def step(f, c, iint):
    """One internal step, pom/advance.f:21-32 (time/ramp handling left to the caller)."""
    lateral_viscosity(f, c)
    mode_interaction(f, c)
    for iext in range(1, int(c["isplit"]) + 1):
        mode_external(f, c, iext)
    mode_internal(f, c, first_cold_step=(iint == 1 and c["time0"] == 0.))


def advt1(f, c, fb_in, fq_in, fclim, ff_in):
    """pom/solver.f:480-574 (nadv=1).  Returns (ff, fb, f) -- fb and f with the reference's side
    effects (level kb copies, the (fb-fclim)+fclim round trip)."""
    u, v, w, aam, dt, h, dx, dy, dum, dvm, art, etb, etf, dz = (
        f[n] for n in "u v w aam dt h dx dy dum dvm art etb etf dz".split())
    im, jm, kb = u.shape
    imm1, jmm1, kbm1 = im - 1, jm - 1, kb - 1
    fq = fq_in.copy(order="F"); fb = fb_in.copy(order="F"); ff = ff_in.copy(order="F")
    fq[:, :, kb - 1] = fq[:, :, kbm1 - 1]
    fb[:, :, kb - 1] = fb[:, :, kbm1 - 1]
    xflux = np.zeros((im, jm, kb), order="F"); yflux = np.zeros((im, jm, kb), order="F")
    I, J = (2, im), (2, jm)
    R = lambda x, di=0, dj=0: _R(x, I, J, di, dj, kbm1)
    xflux[1:, 1:, :kbm1] = .25 * ((R(dt) + R(dt, -1, 0)) * (R(fq) + R(fq, -1, 0)) * R(u))
    yflux[1:, 1:, :kbm1] = .25 * ((R(dt) + R(dt, 0, -1)) * (R(fq) + R(fq, 0, -1)) * R(v))
    fb = fb - fclim
    xflux[1:, 1:, :kbm1] = (R(xflux) - .5 * (R(aam) + R(aam, -1, 0)) * (R(h) + R(h, -1, 0)) * c["tprni"]
                            * (R(fb) - R(fb, -1, 0)) * R(dum) / (R(dx) + R(dx, -1, 0)))
    yflux[1:, 1:, :kbm1] = (R(yflux) - .5 * (R(aam) + R(aam, 0, -1)) * (R(h) + R(h, 0, -1)) * c["tprni"]
                            * (R(fb) - R(fb, 0, -1)) * R(dvm) / (R(dy) + R(dy, 0, -1)))
    xflux[1:, 1:, :kbm1] = .5 * (R(dy) + R(dy, -1, 0)) * R(xflux)
    yflux[1:, 1:, :kbm1] = .5 * (R(dx) + R(dx, 0, -1)) * R(yflux)
    fb = fb + fclim
    zflux = np.zeros((im, jm, kb + 1), order="F")          # zflux(:,:,kb)=0, and k+1 up to kb
    In = (slice(1, imm1), slice(1, jmm1))
    zflux[In + (0,)] = fq[In + (0,)] * w[In + (0,)] * art[In]
    for k in range(2, kbm1 + 1):
        zflux[In + (k - 1,)] = .5 * (fq[In + (k - 2,)] + fq[In + (k - 1,)]) * w[In + (k - 1,)] * art[In]
    for k in range(1, kbm1 + 1):
        z = k - 1
        ff[In + (z,)] = (xflux[2:im, 1:jmm1, z] - xflux[1:imm1, 1:jmm1, z]
                         + yflux[1:imm1, 2:jm, z] - yflux[1:imm1, 1:jmm1, z]
                         + (zflux[In + (z,)] - zflux[In + (z + 1,)]) / dz[z])
        ff[In + (z,)] = ((fb[In + (z,)] * (h[In] + etb[In]) * art[In] - c["dti2"] * ff[In + (z,)])
                         / ((h[In] + etf[In]) * art[In]))
    return ff, fb, fq


def advave_mode2(f):
    """advave's mode=2 block (pom/solver.f:123-195): bottom stress from the depth-averaged velocity
    and the curvature terms, in place on wubot, wvbot, advua, advva (one sub-domain)."""
    cbc, uab, vab, ua, va, dx, dy, d, aru, arv = (f[n] for n in "cbc uab vab ua va dx dy d aru arv".split())
    im, jm = d.shape
    imm1, jmm1 = im - 1, jm - 1
    I, J = (2, imm1), (2, jmm1)
    r = lambda x, di=0, dj=0, I=I, J=J: x[I[0] - 1 + di:I[1] + di, J[0] - 1 + dj:J[1] + dj]
    f["wubot"][1:imm1, 1:jmm1] = (-0.5 * (r(cbc) + r(cbc, -1, 0))
                                  * np.sqrt(r(uab) ** 2 + (.25 * (r(vab) + r(vab, 0, 1) + r(vab, -1, 0) + r(vab, -1, 1))) ** 2)
                                  * r(uab))
    f["wvbot"][1:imm1, 1:jmm1] = (-0.5 * (r(cbc) + r(cbc, 0, -1))
                                  * np.sqrt(r(vab) ** 2 + (.25 * (r(uab) + r(uab, 1, 0) + r(uab, 0, -1) + r(uab, 1, -1))) ** 2)
                                  * r(vab))
    curv = np.zeros((im, jm), order="F")
    curv[1:imm1, 1:jmm1] = (.25 * ((r(va, 0, 1) + r(va)) * (r(dy, 1, 0) - r(dy, -1, 0))
                                   - (r(ua, 1, 0) + r(ua)) * (r(dx, 0, 1) - r(dx, 0, -1)))
                            / (r(dx) * r(dy)))
    Iu = (3, imm1)                                   # n_west == -1
    ru = lambda x, di=0, dj=0: r(x, di, dj, Iu, J)
    f["advua"][2:imm1, 1:jmm1] = (ru(f["advua"]) - ru(aru) * .25
                                  * (ru(curv) * ru(d) * (ru(va, 0, 1) + ru(va))
                                     + ru(curv, -1, 0) * ru(d, -1, 0) * (ru(va, -1, 1) + ru(va, -1, 0))))
    Jv = (3, jmm1)                                   # n_south == -1
    rv = lambda x, di=0, dj=0: r(x, di, dj, I, Jv)
    f["advva"][1:imm1, 2:jmm1] = (rv(f["advva"]) + rv(arv) * .25
                                  * (rv(curv) * rv(d) * (rv(ua, 1, 0) + rv(ua))
                                     + rv(curv, 0, -1) * rv(d, 0, -1) * (rv(ua, 1, -1) + rv(ua, 0, -1))))
