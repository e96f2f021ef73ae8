"""A SECOND, independent restatement (numpy) of a subset of the reference routines.

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see oracle/pomo.h).  The reference cannot be
compiled here, so the C oracle is itself a restatement; this module restates
    dens    pom/solver.f:1162-1209      baropg  pom/solver.f:848-940
    vertvl  pom/solver.f:1970-2021      advq    pom/solver.f:411-477
    advave  pom/solver.f:6-121          proft   pom/solver.f:1541-1683
    profq   pom/solver.f:1212-1538      advt2 + smol_adif  pom/solver.f:577-731,1880-1967
    baropg_mcc  pom/solver.f:943-1159
    mode_external + bcond(1), bcond(2)  pom/advance.f:205-353, pom/bounds_forcing.f:18-83
a second time, written from the Fortran text with whole-array slices instead of loops, and
tests/test_oracle_np.py requires the two restatements to agree BITWISE (same IEEE operations in
the same order; numpy does not contract to FMA).  A transcription slip would have to be made
twice, in two different notations, to go unnoticed.

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
        """nbc 1 or 3 (no short-wave penetration: rad == 0)."""
        assert nbc in (1, 3)
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
        if nbc == 1:
            ee[:, :, 0] = a[:, :, 0] / (a[:, :, 0] - 1.)
            gg[:, :, 0] = dti2 * wfsurf / (dz[0] * dh) - ff[:, :, 0]
            gg[:, :, 0] = gg[:, :, 0] / (a[:, :, 0] - 1.)
        else:
            ee[:, :, 0] = 0.
            gg[:, :, 0] = fsurf
        for k in range(2, kbm2 + 1):
            K = k - 1
            gg[:, :, K] = 1. / (a[:, :, K] + cc[:, :, K] * (1. - ee[:, :, K - 1]) - 1.)
            ee[:, :, K] = a[:, :, K] * gg[:, :, K]
            gg[:, :, K] = (cc[:, :, K] * gg[:, :, K - 1] - ff[:, :, K]) * gg[:, :, K]
        K = kbm1 - 1
        ff[:, :, K] = ((cc[:, :, K] * gg[:, :, K - 1] - ff[:, :, K])
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
