/* pomo_solver.c -- CPU ORACLE restatement of pom/solver.f (reference).
 * TEST INFRASTRUCTURE ONLY; parity pinned against the reference's own source (see pomo.h).
 * Each function cites the solver.f lines it follows; loop bounds, zero
 * fills, evaluation order and single-precision-literal quirks are kept.
 * Build: gcc -O2 -ffp-contract=off (no FMA, like the reference's -O0). */
#define POMO_IMPL
#include "pomo.h"
#include <math.h>
#include <quadmath.h>
#include <string.h>

static void zero3(pomo_t *S, double *a) { memset(a, 0, sizeof(double) * (size_t)S->im * S->jm * S->kb); }
static void zero2(pomo_t *S, double *a) { memset(a, 0, sizeof(double) * (size_t)S->im * S->jm); }

/* ------------------------------------------------------------------ */
/* solver.f:6-198 advave, including the mode=2 block :123-195 (bottom stress from the 2-D
 * velocities and the curvature terms; curv2d lives in the scratch array scr2[1]) */
void pomo_advave(pomo_t *S) {
  DIMS;
  /* :16-18 */
  zero2(S, S->advua); zero2(S, S->fluxua); zero2(S, S->fluxva);
  /* :20-26 */
  OMP_FOR
  DO(j, 2, jm) DO(i, 2, imm1)
    fluxua(i,j)=.125*((d(i+1,j)+d(i,j))*ua(i+1,j)
                      +(d(i,j)+d(i-1,j))*ua(i,j))
                     *(ua(i+1,j)+ua(i,j));
  /* :28-34 */
  OMP_FOR
  DO(j, 2, jm) DO(i, 2, im)
    fluxva(i,j)=.125*((d(i,j)+d(i,j-1))*va(i,j)
                      +(d(i-1,j)+d(i-1,j-1))*va(i-1,j))
                     *(ua(i,j)+ua(i,j-1));
  /* :37-43 */
  OMP_FOR
  DO(j, 2, jm) DO(i, 2, imm1)
    fluxua(i,j)=fluxua(i,j)
                -d(i,j)*2.*aam2d(i,j)*(uab(i+1,j)-uab(i,j))
                  /dx(i,j);
  /* :45-58 */
  OMP_FOR
  DO(j, 2, jm) DO(i, 2, im) {
    tps(i,j)=.25*(d(i,j)+d(i-1,j)+d(i,j-1)+d(i-1,j-1))
             *(aam2d(i,j)+aam2d(i,j-1)
               +aam2d(i-1,j)+aam2d(i-1,j-1))
             *((uab(i,j)-uab(i,j-1))
                /(dy(i,j)+dy(i-1,j)+dy(i,j-1)+dy(i-1,j-1))
              +(vab(i,j)-vab(i-1,j))
                /(dx(i,j)+dx(i-1,j)+dx(i,j-1)+dx(i-1,j-1)));
    fluxua(i,j)=fluxua(i,j)*dy(i,j);
    fluxva(i,j)=(fluxva(i,j)-tps(i,j))*.25
                *(dx(i,j)+dx(i-1,j)+dx(i,j-1)+dx(i-1,j-1));
  }
  /* :60-61 exchange2d_mpi no-op */
  /* :63-68 */
  OMP_FOR
  DO(j, 2, jmm1) DO(i, 2, imm1)
    advua(i,j)=fluxua(i,j)-fluxua(i-1,j)
               +fluxva(i,j+1)-fluxva(i,j);
  /* :73-75 */
  zero2(S, S->advva); zero2(S, S->fluxua); zero2(S, S->fluxva);
  /* :78-84 */
  OMP_FOR
  DO(j, 2, jm) DO(i, 2, im)
    fluxua(i,j)=.125*((d(i,j)+d(i-1,j))*ua(i,j)
                      +(d(i,j-1)+d(i-1,j-1))*ua(i,j-1))
                     *(va(i-1,j)+va(i,j));
  /* :86-92 */
  OMP_FOR
  DO(j, 2, jmm1) DO(i, 2, im)
    fluxva(i,j)=.125*((d(i,j+1)+d(i,j))*va(i,j+1)
                      +(d(i,j)+d(i,j-1))*va(i,j))
                     *(va(i,j+1)+va(i,j));
  /* :95-101 */
  OMP_FOR
  DO(j, 2, jmm1) DO(i, 2, im)
    fluxva(i,j)=fluxva(i,j)
                -d(i,j)*2.*aam2d(i,j)*(vab(i,j+1)-vab(i,j))
                  /dy(i,j);
  /* :103-109 (tps reused from the u half, :47) */
  OMP_FOR
  DO(j, 2, jm) DO(i, 2, im) {
    fluxva(i,j)=fluxva(i,j)*dx(i,j);
    fluxua(i,j)=(fluxua(i,j)-tps(i,j))*.25
                *(dy(i,j)+dy(i-1,j)+dy(i,j-1)+dy(i-1,j-1));
  }
  /* :114-119 */
  OMP_FOR
  DO(j, 2, jmm1) DO(i, 2, imm1)
    advva(i,j)=fluxua(i+1,j)-fluxua(i,j)
               +fluxva(i,j)-fluxva(i,j-1);
  if (S->mode == 2) {
    double *curv2dp = S->scr2[1];
#define curv2d(i, j) (curv2dp[I2(i, j)])
    zero2(S, curv2dp);                    /* COMMON curv2d: zero outside the interior */
    /* :125-133 */
    OMP_FOR
    DO(j, 2, jmm1) DO(i, 2, imm1) {
      double q=.25*(vab(i,j)+vab(i,j+1)+vab(i-1,j)+vab(i-1,j+1));
      wubot(i,j)=-0.5*(cbc(i,j)+cbc(i-1,j))
                 *sqrt(uab(i,j)*uab(i,j)+q*q)
                 *uab(i,j);
    }
    /* :135-143 */
    OMP_FOR
    DO(j, 2, jmm1) DO(i, 2, imm1) {
      double q=.25*(uab(i,j)+uab(i+1,j)+uab(i,j-1)+uab(i+1,j-1));
      wvbot(i,j)=-0.5*(cbc(i,j)+cbc(i,j-1))
                 *sqrt(vab(i,j)*vab(i,j)+q*q)
                 *vab(i,j);
    }
    /* :145-152 */
    OMP_FOR
    DO(j, 2, jmm1) DO(i, 2, imm1)
      curv2d(i,j)=.25
                  *((va(i,j+1)+va(i,j))*(dy(i+1,j)-dy(i-1,j))
                   -(ua(i+1,j)+ua(i,j))*(dx(i,j+1)-dx(i,j-1)))
                  /(dx(i,j)*dy(i,j));
    /* :155-172 (n_west == -1) */
    OMP_FOR
    DO(j, 2, jmm1) DO(i, 3, imm1)
      advua(i,j)=advua(i,j)-aru(i,j)*.25
                 *(curv2d(i,j)*d(i,j)
                   *(va(i,j+1)+va(i,j))
                   +curv2d(i-1,j)*d(i-1,j)
                   *(va(i-1,j+1)+va(i-1,j)));
    /* :174-191 (n_south == -1) */
    DO(i, 2, imm1) DO(j, 3, jmm1)
      advva(i,j)=advva(i,j)+arv(i,j)*.25
                 *(curv2d(i,j)*d(i,j)
                   *(ua(i+1,j)+ua(i,j))
                   +curv2d(i,j-1)*d(i,j-1)
                   *(ua(i+1,j-1)+ua(i,j-1)));
#undef curv2d
  }
}

/* ------------------------------------------------------------------ */
/* solver.f:201-409 advct */
void pomo_advct(pomo_t *S) {
  DIMS;
  double *xfluxp = S->scr3[0], *yfluxp = S->scr3[1], *curvp = S->scr3[2];
#define xflux(i, j, k) (xfluxp[I3(i, j, k)])
#define yflux(i, j, k) (yfluxp[I3(i, j, k)])
#define curv(i, j, k) (curvp[I3(i, j, k)])
  double dtaam;
  /* :213-216 */
  zero3(S, curvp); zero3(S, S->advx); zero3(S, xfluxp); zero3(S, yfluxp);
  /* :218-228 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1)
    curv(i,j,k)=.25*((v(i,j+1,k)+v(i,j,k))
                      *(dy(i+1,j)-dy(i-1,j))
                     -(u(i+1,j,k)+u(i,j,k))
                      *(dx(i,j+1)-dx(i,j-1)))
                    /(dx(i,j)*dy(i,j));
  /* :234-242 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 1, jm) DO(i, 2, imm1)
    xflux(i,j,k)=.125*((dt(i+1,j)+dt(i,j))*u(i+1,j,k)
                        +(dt(i,j)+dt(i-1,j))*u(i,j,k))
                       *(u(i+1,j,k)+u(i,j,k));
  /* :244-252 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jm) DO(i, 2, im)
    yflux(i,j,k)=.125*((dt(i,j)+dt(i,j-1))*v(i,j,k)
                        +(dt(i-1,j)+dt(i-1,j-1))*v(i-1,j,k))
                       *(u(i,j,k)+u(i,j-1,k));
  /* :255-277 */
#pragma omp parallel for schedule(static) private(dtaam)
  DO(k, 1, kbm1) DO(j, 2, jm) DO(i, 2, imm1) {
    xflux(i,j,k)=xflux(i,j,k)
                 -dt(i,j)*aam(i,j,k)*2.
                 *(ub(i+1,j,k)-ub(i,j,k))/dx(i,j);
    dtaam=.25*(dt(i,j)+dt(i-1,j)+dt(i,j-1)+dt(i-1,j-1))
          *(aam(i,j,k)+aam(i-1,j,k)
            +aam(i,j-1,k)+aam(i-1,j-1,k));
    yflux(i,j,k)=yflux(i,j,k)
                 -dtaam*((ub(i,j,k)-ub(i,j-1,k))
                         /(dy(i,j)+dy(i-1,j)
                           +dy(i,j-1)+dy(i-1,j-1))
                         +(vb(i,j,k)-vb(i-1,j,k))
                         /(dx(i,j)+dx(i-1,j)
                           +dx(i,j-1)+dx(i-1,j-1)));
    xflux(i,j,k)=dy(i,j)*xflux(i,j,k);
    yflux(i,j,k)=.25*(dx(i,j)+dx(i-1,j)
                       +dx(i,j-1)+dx(i-1,j-1))*yflux(i,j,k);
  }
  /* :282-289 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1)
    advx(i,j,k)=xflux(i,j,k)-xflux(i-1,j,k)
                +yflux(i,j+1,k)-yflux(i,j,k);
  /* :291-313 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jmm1) {
    int i0 = (n_west == -1) ? 3 : 2;
    DO(i, i0, imm1)
      advx(i,j,k)=advx(i,j,k)
                  -aru(i,j)*.25
                    *(curv(i,j,k)*dt(i,j)
                       *(v(i,j+1,k)+v(i,j,k))
                      +curv(i-1,j,k)*dt(i-1,j)
                       *(v(i-1,j+1,k)+v(i-1,j,k)));
  }
  /* :319-321 */
  zero3(S, S->advy); zero3(S, xfluxp); zero3(S, yfluxp);
  /* :324-332 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jm) DO(i, 2, im)
    xflux(i,j,k)=.125*((dt(i,j)+dt(i-1,j))*u(i,j,k)
                        +(dt(i,j-1)+dt(i-1,j-1))*u(i,j-1,k))
                       *(v(i,j,k)+v(i-1,j,k));
  /* :334-342 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jmm1) DO(i, 1, im)
    yflux(i,j,k)=.125*((dt(i,j+1)+dt(i,j))*v(i,j+1,k)
                        +(dt(i,j)+dt(i,j-1))*v(i,j,k))
                       *(v(i,j+1,k)+v(i,j,k));
  /* :345-367 */
#pragma omp parallel for schedule(static) private(dtaam)
  DO(k, 1, kbm1) DO(j, 2, jmm1) DO(i, 2, im) {
    dtaam=.25*(dt(i,j)+dt(i-1,j)+dt(i,j-1)+dt(i-1,j-1))
          *(aam(i,j,k)+aam(i-1,j,k)
            +aam(i,j-1,k)+aam(i-1,j-1,k));
    xflux(i,j,k)=xflux(i,j,k)
                 -dtaam*((ub(i,j,k)-ub(i,j-1,k))
                         /(dy(i,j)+dy(i-1,j)
                           +dy(i,j-1)+dy(i-1,j-1))
                         +(vb(i,j,k)-vb(i-1,j,k))
                         /(dx(i,j)+dx(i-1,j)
                           +dx(i,j-1)+dx(i-1,j-1)));
    yflux(i,j,k)=yflux(i,j,k)
                 -dt(i,j)*aam(i,j,k)*2.
                 *(vb(i,j+1,k)-vb(i,j,k))/dy(i,j);
    xflux(i,j,k)=.25*(dy(i,j)+dy(i-1,j)
                       +dy(i,j-1)+dy(i-1,j-1))*xflux(i,j,k);
    yflux(i,j,k)=dx(i,j)*yflux(i,j,k);
  }
  /* :372-379 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1)
    advy(i,j,k)=xflux(i+1,j,k)-xflux(i,j,k)
                +yflux(i,j,k)-yflux(i,j-1,k);
  /* :381-403 */
  OMP_FOR
  DO(k, 1, kbm1) DO(i, 2, imm1) {
    int j0 = (n_south == -1) ? 3 : 2;
    DO(j, j0, jmm1)
      advy(i,j,k)=advy(i,j,k)
                  +arv(i,j)*.25
                    *(curv(i,j,k)*dt(i,j)
                       *(u(i+1,j,k)+u(i,j,k))
                      +curv(i,j-1,k)*dt(i,j-1)
                       *(u(i+1,j-1,k)+u(i,j-1,k)));
  }
#undef xflux
#undef yflux
#undef curv
}

/* ------------------------------------------------------------------ */
/* solver.f:411-477 advq */
void pomo_advq(pomo_t *S, double *qbp, double *qp, double *qfp) {
  DIMS;
  double *xfluxp = S->scr3[0], *yfluxp = S->scr3[1];
#define xflux(i, j, k) (xfluxp[I3(i, j, k)])
#define yflux(i, j, k) (yfluxp[I3(i, j, k)])
#define qb(i, j, k) (qbp[I3(i, j, k)])
#define q(i, j, k) (qp[I3(i, j, k)])
#define qf(i, j, k) (qfp[I3(i, j, k)])
  zero3(S, xfluxp); zero3(S, yfluxp);
  /* :425-434 */
  OMP_FOR
  DO(k, 2, kbm1) DO(j, 2, jm) DO(i, 2, im) {
    xflux(i,j,k)=.125*(q(i,j,k)+q(i-1,j,k))
                 *(dt(i,j)+dt(i-1,j))*(u(i,j,k)+u(i,j,k-1));
    yflux(i,j,k)=.125*(q(i,j,k)+q(i,j-1,k))
                 *(dt(i,j)+dt(i,j-1))*(v(i,j,k)+v(i,j,k-1));
  }
  /* :437-456 */
  OMP_FOR
  DO(k, 2, kbm1) DO(j, 2, jm) DO(i, 2, im) {
    xflux(i,j,k)=xflux(i,j,k)
                 -.25*(aam(i,j,k)+aam(i-1,j,k)
                       +aam(i,j,k-1)+aam(i-1,j,k-1))
                     *(h(i,j)+h(i-1,j))
                     *(qb(i,j,k)-qb(i-1,j,k))*dum(i,j)
                     /(dx(i,j)+dx(i-1,j));
    yflux(i,j,k)=yflux(i,j,k)
                 -.25*(aam(i,j,k)+aam(i,j-1,k)
                       +aam(i,j,k-1)+aam(i,j-1,k-1))
                     *(h(i,j)+h(i,j-1))
                     *(qb(i,j,k)-qb(i,j-1,k))*dvm(i,j)
                     /(dy(i,j)+dy(i,j-1));
    xflux(i,j,k)=.5*(dy(i,j)+dy(i-1,j))*xflux(i,j,k);
    yflux(i,j,k)=.5*(dx(i,j)+dx(i,j-1))*yflux(i,j,k);
  }
  /* :462-474 */
  OMP_FOR
  DO(k, 2, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1) {
    qf(i,j,k)=(w(i,j,k-1)*q(i,j,k-1)-w(i,j,k+1)*q(i,j,k+1))
              *art(i,j)/(dz(k)+dz(k-1))
              +xflux(i+1,j,k)-xflux(i,j,k)
              +yflux(i,j+1,k)-yflux(i,j,k);
    qf(i,j,k)=((h(i,j)+etb(i,j))*art(i,j)
               *qb(i,j,k)-dti2*qf(i,j,k))
              /((h(i,j)+etf(i,j))*art(i,j));
  }
#undef xflux
#undef yflux
#undef qb
#undef q
#undef qf
}

/* ------------------------------------------------------------------ */
/* solver.f:480-574 advt1 */
void pomo_advt1(pomo_t *S, double *fbp, double *fp, double *fclimp, double *ffp) {
  DIMS;
  double *xfluxp = S->scr3[0], *yfluxp = S->scr3[1];
#define xflux(i, j, k) (xfluxp[I3(i, j, k)])
#define yflux(i, j, k) (yfluxp[I3(i, j, k)])
#define fb(i, j, k) (fbp[I3(i, j, k)])
#define f(i, j, k) (fp[I3(i, j, k)])
#define ff(i, j, k) (ffp[I3(i, j, k)])
  zero3(S, xfluxp); zero3(S, yfluxp);
  /* :495-496 */
  DO(j, 1, jm) DO(i, 1, im) { f(i,j,kb)=f(i,j,kbm1); fb(i,j,kb)=fb(i,j,kbm1); }
  /* :499-508 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jm) DO(i, 2, im) {
    xflux(i,j,k)=.25*((dt(i,j)+dt(i-1,j))
                       *(f(i,j,k)+f(i-1,j,k))*u(i,j,k));
    yflux(i,j,k)=.25*((dt(i,j)+dt(i,j-1))
                       *(f(i,j,k)+f(i,j-1,k))*v(i,j,k));
  }
  /* :511 */
  for (size_t n = 0; n < N3; ++n) fbp[n] = fbp[n] - fclimp[n];
  /* :513-530 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jm) DO(i, 2, im) {
    xflux(i,j,k)=xflux(i,j,k)
                 -.5*(aam(i,j,k)+aam(i-1,j,k))
                    *(h(i,j)+h(i-1,j))*tprni
                    *(fb(i,j,k)-fb(i-1,j,k))*dum(i,j)
                    /(dx(i,j)+dx(i-1,j));
    yflux(i,j,k)=yflux(i,j,k)
                 -.5*(aam(i,j,k)+aam(i,j-1,k))
                    *(h(i,j)+h(i,j-1))*tprni
                    *(fb(i,j,k)-fb(i,j-1,k))*dvm(i,j)
                    /(dy(i,j)+dy(i,j-1));
    xflux(i,j,k)=.5*(dy(i,j)+dy(i-1,j))*xflux(i,j,k);
    yflux(i,j,k)=.5*(dx(i,j)+dx(i,j-1))*yflux(i,j,k);
  }
  /* :532 */
  for (size_t n = 0; n < N3; ++n) fbp[n] = fbp[n] + fclimp[n];
  /* :535-540 */
  DO(j, 2, jmm1) DO(i, 2, imm1) {
    zflux(i,j,1)=f(i,j,1)*w(i,j,1)*art(i,j);
    zflux(i,j,kb)=0.;
  }
  /* :542-548 */
  OMP_FOR
  DO(k, 2, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1)
    zflux(i,j,k)=.5*(f(i,j,k-1)+f(i,j,k))*w(i,j,k)*art(i,j);
  /* :562-571 */
  OMP_FOR
  DO(k, 1, kbm1) {
    DO(j, 2, jmm1) DO(i, 2, imm1)
      ff(i,j,k)=xflux(i+1,j,k)-xflux(i,j,k)
                +yflux(i,j+1,k)-yflux(i,j,k)
                +(zflux(i,j,k)-zflux(i,j,k+1))/dz(k);
    DO(j, 2, jmm1) DO(i, 2, imm1)
      ff(i,j,k)=(fb(i,j,k)
                 *(h(i,j)+etb(i,j))*art(i,j)
                 -dti2*ff(i,j,k))
                /((h(i,j)+etf(i,j))
                  *art(i,j));
  }
#undef xflux
#undef yflux
#undef fb
#undef f
#undef ff
}

/* ------------------------------------------------------------------ */
/* solver.f:1880-1967 smol_adif */
void pomo_smol_adif(pomo_t *S, double *xmp, double *ymp, double *zwp, double *ffp) {
  DIMS;
#define xmassflux(i, j, k) (xmp[I3(i, j, k)])
#define ymassflux(i, j, k) (ymp[I3(i, j, k)])
#define zwflux(i, j, k) (zwp[I3(i, j, k)])
#define ff(i, j, k) (ffp[I3(i, j, k)])
  const double value_min = 1.e-9, epsilon = 1.0e-14;
  /* :1898-1900 */
  DO(k, 1, kb) DO(j, 1, jm) DO(i, 1, im) ff(i,j,k)=ff(i,j,k)*fsm(i,j);
  /* :1903-1922 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jmm1) DO(i, 2, im) {
    if (ff(i,j,k) < value_min || ff(i-1,j,k) < value_min) {
      xmassflux(i,j,k)=0.;
    } else {
      double udx=fabs(xmassflux(i,j,k));
      double u2dt=dti2*xmassflux(i,j,k)*xmassflux(i,j,k)*2.
                  /(aru(i,j)*(dt(i-1,j)+dt(i,j)));
      double mol=(ff(i,j,k)-ff(i-1,j,k))
                 /(ff(i-1,j,k)+ff(i,j,k)+epsilon);
      xmassflux(i,j,k)=(udx-u2dt)*mol*sw;
      double abs_1=fabs(udx), abs_2=fabs(u2dt);
      if (abs_1 < abs_2) xmassflux(i,j,k)=0.;
    }
  }
  /* :1924-1943 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jm) DO(i, 2, imm1) {
    if (ff(i,j,k) < value_min || ff(i,j-1,k) < value_min) {
      ymassflux(i,j,k)=0.;
    } else {
      double vdy=fabs(ymassflux(i,j,k));
      double v2dt=dti2*ymassflux(i,j,k)*ymassflux(i,j,k)*2.
                  /(arv(i,j)*(dt(i,j-1)+dt(i,j)));
      double mol=(ff(i,j,k)-ff(i,j-1,k))
                 /(ff(i,j-1,k)+ff(i,j,k)+epsilon);
      ymassflux(i,j,k)=(vdy-v2dt)*mol*sw;
      double abs_1=fabs(vdy), abs_2=fabs(v2dt);
      if (abs_1 < abs_2) ymassflux(i,j,k)=0.;
    }
  }
  /* :1945-1964 */
  OMP_FOR
  DO(k, 2, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1) {
    if (ff(i,j,k) < value_min || ff(i,j,k-1) < value_min) {
      zwflux(i,j,k)=0.;
    } else {
      double wdz=fabs(zwflux(i,j,k));
      double w2dt=dti2*zwflux(i,j,k)*zwflux(i,j,k)/
                  (dzz(k-1)*dt(i,j));
      double mol=(ff(i,j,k-1)-ff(i,j,k))
                 /(ff(i,j,k)+ff(i,j,k-1)+epsilon);
      zwflux(i,j,k)=(wdz-w2dt)*mol*sw;
      double abs_1=fabs(wdz), abs_2=fabs(w2dt);
      if (abs_1 < abs_2) zwflux(i,j,k)=0.;
    }
  }
#undef xmassflux
#undef ymassflux
#undef zwflux
#undef ff
}

/* ------------------------------------------------------------------ */
/* solver.f:577-731 advt2 */
void pomo_advt2(pomo_t *S, double *fbp, double *fp, double *fclimp, double *ffp) {
  DIMS;
  double *xfluxp = S->scr3[0], *yfluxp = S->scr3[1], *fbmemp = S->scr3[2];
  double *xmp = S->scr3[3], *ymp = S->scr3[4], *zwp = S->scr3[5];
  double *etap = S->scr2[0];
#define xflux(i, j, k) (xfluxp[I3(i, j, k)])
#define yflux(i, j, k) (yfluxp[I3(i, j, k)])
#define fbmem(i, j, k) (fbmemp[I3(i, j, k)])
#define xmassflux(i, j, k) (xmp[I3(i, j, k)])
#define ymassflux(i, j, k) (ymp[I3(i, j, k)])
#define zwflux(i, j, k) (zwp[I3(i, j, k)])
#define eta(i, j) (etap[I2(i, j)])
#define fb(i, j, k) (fbp[I3(i, j, k)])
#define f(i, j, k) (fp[I3(i, j, k)])
#define ff(i, j, k) (ffp[I3(i, j, k)])
  /* :597-600 */
  zero3(S, xfluxp); zero3(S, yfluxp); zero3(S, xmp); zero3(S, ymp);
  /* :602-616 */
  OMP_FOR
  DO(k, 1, kbm1) {
    DO(j, 2, jmm1) DO(i, 2, im)
      xmassflux(i,j,k)=0.25*(dy(i-1,j)+dy(i,j))
                           *(dt(i-1,j)+dt(i,j))*u(i,j,k);
    DO(j, 2, jm) DO(i, 2, imm1)
      ymassflux(i,j,k)=0.25*(dx(i,j-1)+dx(i,j))
                           *(dt(i,j-1)+dt(i,j))*v(i,j,k);
  }
  /* :618-622 */
  DO(j, 1, jm) DO(i, 1, im) fb(i,j,kb)=fb(i,j,kbm1);
  DO(j, 1, jm) DO(i, 1, im) eta(i,j)=etb(i,j);
  memcpy(zwp, S->w, sizeof(double) * N3);
  memcpy(fbmemp, fbp, sizeof(double) * N3);
  /* :625 start Smolarkiewicz scheme */
  DO(itera, 1, S->nitera) {
    /* :628-644 */
    OMP_FOR
    DO(k, 1, kbm1) DO(j, 2, jm) DO(i, 2, im) {
      xflux(i,j,k)=0.5
                   *((xmassflux(i,j,k)+fabs(xmassflux(i,j,k)))
                     *fbmem(i-1,j,k)+
                     (xmassflux(i,j,k)-fabs(xmassflux(i,j,k)))
                     *fbmem(i,j,k));
      yflux(i,j,k)=0.5
                   *((ymassflux(i,j,k)+fabs(ymassflux(i,j,k)))
                     *fbmem(i,j-1,k)+
                     (ymassflux(i,j,k)-fabs(ymassflux(i,j,k)))
                     *fbmem(i,j,k));
    }
    /* :646-651 */
    DO(j, 2, jmm1) DO(i, 2, imm1) zflux(i,j,1)=0.;
    if (itera == 1)
      DO(j, 2, jmm1) DO(i, 2, imm1)
        zflux(i,j,1)=w(i,j,1)*f(i,j,1)*art(i,j);
    DO(j, 2, jmm1) DO(i, 2, imm1) zflux(i,j,kb)=0.;
    /* :653-664 */
    OMP_FOR
    DO(k, 2, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1) {
      zflux(i,j,k)=0.5
                   *((zwflux(i,j,k)+fabs(zwflux(i,j,k)))
                    *fbmem(i,j,k)+
                     (zwflux(i,j,k)-fabs(zwflux(i,j,k)))
                    *fbmem(i,j,k-1));
      zflux(i,j,k)=zflux(i,j,k)*art(i,j);
    }
    /* :667-677 */
    OMP_FOR
    DO(j, 2, jmm1) DO(i, 2, imm1) DO(k, 1, kbm1) {
      ff(i,j,k)=xflux(i+1,j,k)-xflux(i,j,k)
                +yflux(i,j+1,k)-yflux(i,j,k)
                +(zflux(i,j,k)-zflux(i,j,k+1))/dz(k);
      ff(i,j,k)=(fbmem(i,j,k)*((h(i,j)+eta(i,j))*art(i,j))
                 -dti2*ff(i,j,k))/((h(i,j)+etf(i,j))*art(i,j));
    }
    /* :679 exchange no-op; :682 */
    pomo_smol_adif(S, xmp, ymp, zwp, ffp);
    /* :684-685 */
    DO(j, 1, jm) DO(i, 1, im) eta(i,j)=etf(i,j);
    memcpy(fbmemp, ffp, sizeof(double) * N3);
  }
  /* :691 */
  for (size_t n = 0; n < N3; ++n) fbp[n] = fbp[n] - fclimp[n];
  /* :693-700 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jm) DO(i, 2, im) {
    xmassflux(i,j,k)=0.5*(aam(i,j,k)+aam(i-1,j,k));
    ymassflux(i,j,k)=0.5*(aam(i,j,k)+aam(i,j-1,k));
  }
  /* :702-713 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jm) DO(i, 2, im) {
    xflux(i,j,k)=-xmassflux(i,j,k)*(h(i,j)+h(i-1,j))*tprni
                 *(fb(i,j,k)-fb(i-1,j,k))*dum(i,j)
                 *(dy(i,j)+dy(i-1,j))*0.5/(dx(i,j)+dx(i-1,j));
    yflux(i,j,k)=-ymassflux(i,j,k)*(h(i,j)+h(i,j-1))*tprni
                 *(fb(i,j,k)-fb(i,j-1,k))*dvm(i,j)
                 *(dx(i,j)+dx(i,j-1))*0.5/(dy(i,j)+dy(i,j-1));
  }
  /* :715 */
  for (size_t n = 0; n < N3; ++n) fbp[n] = fbp[n] + fclimp[n];
  /* :718-726 */
  OMP_FOR
  DO(j, 2, jmm1) DO(i, 2, imm1) DO(k, 1, kbm1)
    ff(i,j,k)=ff(i,j,k)-dti2*(xflux(i+1,j,k)-xflux(i,j,k)
                              +yflux(i,j+1,k)-yflux(i,j,k))
                        /((h(i,j)+etf(i,j))*art(i,j));
#undef xflux
#undef yflux
#undef fbmem
#undef xmassflux
#undef ymassflux
#undef zwflux
#undef eta
#undef fb
#undef f
#undef ff
}

/* ------------------------------------------------------------------ */
/* solver.f:734-788 advu */
void pomo_advu(pomo_t *S) {
  DIMS;
  zero3(S, S->uf); /* :742 */
  /* :744-751 */
  OMP_FOR
  DO(k, 2, kbm1) DO(j, 1, jm) DO(i, 2, im)
    uf(i,j,k)=.25*(w(i,j,k)+w(i-1,j,k))
                  *(u(i,j,k)+u(i,j,k-1));
  /* :755-772 (ascending k: reads uf(k+1) before it is overwritten) */
  DO(k, 1, kbm1) {
    OMP_FOR
    DO(j, 2, jmm1) DO(i, 2, imm1)
      uf(i,j,k)=advx(i,j,k)
                +(uf(i,j,k)-uf(i,j,k+1))*aru(i,j)/dz(k)
                -aru(i,j)*.25
                  *(cor(i,j)*dt(i,j)
                     *(v(i,j+1,k)+v(i,j,k))
                    +cor(i-1,j)*dt(i-1,j)
                      *(v(i-1,j+1,k)+v(i-1,j,k)))
                +grav*.125*(dt(i,j)+dt(i-1,j))
                  *(egf(i,j)-egf(i-1,j)+egb(i,j)-egb(i-1,j)
                    +(e_atmos(i,j)-e_atmos(i-1,j))*2.)
                  *(dy(i,j)+dy(i-1,j))
                +drhox(i,j,k);
  }
  /* :775-785 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1)
    uf(i,j,k)=((h(i,j)+etb(i,j)+h(i-1,j)+etb(i-1,j))
               *aru(i,j)*ub(i,j,k)
               -2.*dti2*uf(i,j,k))
              /((h(i,j)+etf(i,j)+h(i-1,j)+etf(i-1,j))
                *aru(i,j));
}

/* solver.f:791-845 advv */
void pomo_advv(pomo_t *S) {
  DIMS;
  zero3(S, S->vf); /* :799 */
  /* :801-808 */
  OMP_FOR
  DO(k, 2, kbm1) DO(j, 2, jm) DO(i, 1, im)
    vf(i,j,k)=.25*(w(i,j,k)+w(i,j-1,k))
                  *(v(i,j,k)+v(i,j,k-1));
  /* :812-829 */
  DO(k, 1, kbm1) {
    OMP_FOR
    DO(j, 2, jmm1) DO(i, 2, imm1)
      vf(i,j,k)=advy(i,j,k)
                +(vf(i,j,k)-vf(i,j,k+1))*arv(i,j)/dz(k)
                +arv(i,j)*.25
                  *(cor(i,j)*dt(i,j)
                     *(u(i+1,j,k)+u(i,j,k))
                    +cor(i,j-1)*dt(i,j-1)
                      *(u(i+1,j-1,k)+u(i,j-1,k)))
                +grav*.125*(dt(i,j)+dt(i,j-1))
                  *(egf(i,j)-egf(i,j-1)+egb(i,j)-egb(i,j-1)
                    +(e_atmos(i,j)-e_atmos(i,j-1))*2.)
                  *(dx(i,j)+dx(i,j-1))
                +drhoy(i,j,k);
  }
  /* :832-842 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1)
    vf(i,j,k)=((h(i,j)+etb(i,j)+h(i,j-1)+etb(i,j-1))
               *arv(i,j)*vb(i,j,k)
               -2.*dti2*vf(i,j,k))
              /((h(i,j)+etf(i,j)+h(i,j-1)+etf(i,j-1))
                *arv(i,j));
}

/* ------------------------------------------------------------------ */
/* solver.f:848-940 baropg */
void pomo_baropg(pomo_t *S) {
  DIMS;
  /* :854 */
  for (size_t n = 0; n < N3; ++n) S->rho[n] = S->rho[n] - S->rmean[n];
  /* :857-862 */
  DO(j, 2, jmm1) DO(i, 2, imm1)
    drhox(i,j,1)=.5*grav*(-zz(1))*(dt(i,j)+dt(i-1,j))
                 *(rho(i,j,1)-rho(i-1,j,1));
  /* :864-878 */
  DO(k, 2, kbm1) {
    OMP_FOR
    DO(j, 2, jmm1) DO(i, 2, imm1)
      drhox(i,j,k)=drhox(i,j,k-1)
                   +grav*.25*(zz(k-1)-zz(k))
                     *(dt(i,j)+dt(i-1,j))
                     *(rho(i,j,k)-rho(i-1,j,k)
                       +rho(i,j,k-1)-rho(i-1,j,k-1))
                   +grav*.25*(zz(k-1)+zz(k))
                     *(dt(i,j)-dt(i-1,j))
                     *(rho(i,j,k)+rho(i-1,j,k)
                       -rho(i,j,k-1)-rho(i-1,j,k-1));
  }
  /* :880-888 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1)
    drhox(i,j,k)=.25*(dt(i,j)+dt(i-1,j))
                     *drhox(i,j,k)*dum(i,j)
                     *(dy(i,j)+dy(i-1,j));
  /* :893-898 */
  DO(j, 2, jmm1) DO(i, 2, imm1)
    drhoy(i,j,1)=.5*grav*(-zz(1))*(dt(i,j)+dt(i,j-1))
                 *(rho(i,j,1)-rho(i,j-1,1));
  /* :900-914 */
  DO(k, 2, kbm1) {
    OMP_FOR
    DO(j, 2, jmm1) DO(i, 2, imm1)
      drhoy(i,j,k)=drhoy(i,j,k-1)
                   +grav*.25*(zz(k-1)-zz(k))
                     *(dt(i,j)+dt(i,j-1))
                     *(rho(i,j,k)-rho(i,j-1,k)
                       +rho(i,j,k-1)-rho(i,j-1,k-1))
                   +grav*.25*(zz(k-1)+zz(k))
                     *(dt(i,j)-dt(i,j-1))
                     *(rho(i,j,k)+rho(i,j-1,k)
                       -rho(i,j,k-1)-rho(i,j-1,k-1));
  }
  /* :916-924 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1)
    drhoy(i,j,k)=.25*(dt(i,j)+dt(i,j-1))
                     *drhoy(i,j,k)*dvm(i,j)
                     *(dx(i,j)+dx(i,j-1));
  /* :928-935 */
  OMP_FOR
  DO(k, 1, kb) DO(j, 2, jmm1) DO(i, 2, imm1) {
    drhox(i,j,k)=ramp*drhox(i,j,k);
    drhoy(i,j,k)=ramp*drhoy(i,j,k);
  }
  /* :937 */
  for (size_t n = 0; n < N3; ++n) S->rho[n] = S->rho[n] + S->rmean[n];
}

/* ------------------------------------------------------------------ */
/* solver.f:943-1159 baropg_mcc: 4th-order (McCalpin) pressure gradient, one sub-domain
 * (n_west == n_south == -1, so rho4th/d4th of order2d/3d_mpi are never read).  Note the
 * single-precision literals (1./24.), (1./24), (1./16.), 0.5 and that d (not dt) is used. */
void pomo_baropg_mcc(pomo_t *S) {
  DIMS;
  double *d4p = S->scr2[1], *ddxp = S->scr2[2], *drhop = S->scr3[0], *rhoup = S->scr3[1];
#define d4(i, j) (d4p[I2(i, j)])
#define ddx(i, j) (ddxp[I2(i, j)])
#define drho(i, j, k) (drhop[I3(i, j, k)])
#define rhou(i, j, k) (rhoup[I3(i, j, k)])
  const double c24 = (double)(1.f / 24.f), c16 = (double)(1.f / 16.f);
  for (size_t n = 0; n < N3; ++n) S->rho[n] = S->rho[n] - S->rmean[n];
  zero2(S, ddxp); zero2(S, d4p); zero3(S, rhoup); zero3(S, drhop);
  /* :970-980 */
  DO(j, 1, jm) DO(i, 2, im) {
    DO(k, 1, kbm1) {
      drho(i,j,k)=(rho(i,j,k)-rho(i-1,j,k))*dum(i,j);
      rhou(i,j,k)=0.5*(rho(i,j,k)+rho(i-1,j,k))*dum(i,j);
    }
    ddx(i,j)=(d(i,j)-d(i-1,j))*dum(i,j);
    d4(i,j)=.5*(d(i,j)+d(i-1,j))*dum(i,j);
  }
  /* :982-1003 (n_west == -1) */
  DO(j, 1, jm) DO(i, 3, imm1) {
    DO(k, 1, kbm1) {
      drho(i,j,k)=drho(i,j,k) - c24*
                  (dum(i+1,j)*(rho(i+1,j,k)-rho(i,j,k))-
                  2*(rho(i,j,k)-rho(i-1,j,k))+
                  dum(i-1,j)*(rho(i-1,j,k)-rho(i-2,j,k)));
      rhou(i,j,k)=rhou(i,j,k) + c16*
                  (dum(i+1,j)*(rho(i,j,k)-rho(i+1,j,k))+
                  dum(i-1,j)*(rho(i-1,j,k)-rho(i-2,j,k)));
    }
    ddx(i,j)=ddx(i,j)-c24*
             (dum(i+1,j)*(d(i+1,j)-d(i,j))-
             2*(d(i,j)-d(i-1,j))+
             dum(i-1,j)*(d(i-1,j)-d(i-2,j)));
    d4(i,j)=d4(i,j)+c16*
            (dum(i+1,j)*(d(i,j)-d(i+1,j))+
            dum(i-1,j)*(d(i-1,j)-d(i-2,j)));
  }
  /* :1029-1057 */
  DO(j, 2, jmm1) DO(i, 2, imm1)
    drhox(i,j,1)=grav*(-zz(1))*d4(i,j)*drho(i,j,1);
  DO(k, 2, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1)
    drhox(i,j,k)=drhox(i,j,k-1)
                 +grav*0.5*dzz(k-1)*d4(i,j)
                 *(drho(i,j,k-1)+drho(i,j,k))
                 +grav*0.5*(zz(k-1)+zz(k))*ddx(i,j)
                 *(rhou(i,j,k)-rhou(i,j,k-1));
  DO(k, 1, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1)
    drhox(i,j,k)=.25*(dt(i,j)+dt(i-1,j))
                      *drhox(i,j,k)*dum(i,j)
                      *(dy(i,j)+dy(i-1,j));
  /* :1062-1065 */
  zero2(S, ddxp); zero2(S, d4p); zero3(S, rhoup); zero3(S, drhop);
  /* :1068-1078 */
  DO(j, 2, jm) DO(i, 1, im) {
    DO(k, 1, kbm1) {
      drho(i,j,k)=(rho(i,j,k)-rho(i,j-1,k))*dvm(i,j);
      rhou(i,j,k)=.5*(rho(i,j,k)+rho(i,j-1,k))*dvm(i,j);
    }
    ddx(i,j)=(d(i,j)-d(i,j-1))*dvm(i,j);
    d4(i,j)=.5*(d(i,j)+d(i,j-1))*dvm(i,j);
  }
  /* :1080-1101 (n_south == -1) */
  DO(j, 3, jmm1) DO(i, 1, im) {
    DO(k, 1, kbm1) {
      drho(i,j,k)=drho(i,j,k)-c24*
                  (dvm(i,j+1)*(rho(i,j+1,k)-rho(i,j,k))-
                  2*(rho(i,j,k)-rho(i,j-1,k))+
                  dvm(i,j-1)*(rho(i,j-1,k)-rho(i,j-2,k)));
      rhou(i,j,k)=rhou(i,j,k)+c16*
                  (dvm(i,j+1)*(rho(i,j,k)-rho(i,j+1,k))+
                  dvm(i,j-1)*(rho(i,j-1,k)-rho(i,j-2,k)));
    }
    ddx(i,j)=ddx(i,j)-c24*
             (dvm(i,j+1)*(d(i,j+1)-d(i,j))-
             2*(d(i,j)-d(i,j-1))+
             dvm(i,j-1)*(d(i,j-1)-d(i,j-2)));
    d4(i,j)=d4(i,j)+c16*
            (dvm(i,j+1)*(d(i,j)-d(i,j+1))+
            dvm(i,j-1)*(d(i,j-1)-d(i,j-2)));
  }
  /* :1127-1153 */
  DO(j, 2, jmm1) DO(i, 2, imm1)
    drhoy(i,j,1)=grav*(-zz(1))*d4(i,j)*drho(i,j,1);
  DO(k, 2, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1)
    drhoy(i,j,k)=drhoy(i,j,k-1)
                 +grav*0.5*dzz(k-1)*d4(i,j)
                 *(drho(i,j,k-1)+drho(i,j,k))
                 +grav*0.5*(zz(k-1)+zz(k))*ddx(i,j)
                 *(rhou(i,j,k)-rhou(i,j,k-1));
  DO(k, 1, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1)
    drhoy(i,j,k)=.25*(dt(i,j)+dt(i,j-1))
                      *drhoy(i,j,k)*dvm(i,j)
                      *(dx(i,j)+dx(i,j-1));
  /* :1157-1164 */
  DO(k, 1, kb) DO(j, 2, jmm1) DO(i, 2, imm1) {
    drhox(i,j,k)=ramp*drhox(i,j,k);
    drhoy(i,j,k)=ramp*drhoy(i,j,k);
  }
  for (size_t n = 0; n < N3; ++n) S->rho[n] = S->rho[n] + S->rmean[n];
#undef d4
#undef ddx
#undef drho
#undef rhou
}

/* ------------------------------------------------------------------ */
/* |S|**1.5 (solver.f:1195).  pow_mode 0 (default): libm pow, what gfortran emits for a real
 * exponent.  pow_mode 1: x*sqrt(x) with the rounding errors of sqrt and of the product
 * recovered by fma -- the routine the CUDA path uses; tests switch to it to show that this
 * libm call is the ONLY source of GPU/oracle differences (everything else is bitwise). */
static double pow15(const pomo_t *S, double x) {
  if (S->pow_mode == 0) return pow(x, 1.5);
  if (!(x > 0.)) return 0.;
  double sq = sqrt(x);
  double res = fma(-sq, sq, x);
  double ds = res / (2. * sq);
  double pr = x * sq;
  double er = fma(x, sq, -pr);
  return pr + (er + x * ds);
}

/* solver.f:1162-1209 dens */
void pomo_dens(pomo_t *S, double *sip, double *tip, double *rhoop) {
  DIMS;
#define si(i, j, k) (sip[I3(i, j, k)])
#define ti(i, j, k) (tip[I3(i, j, k)])
#define rhoo(i, j, k) (rhoop[I3(i, j, k)])
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 1, jm) DO(i, 1, im) {
    double tr=ti(i,j,k)+tbias;
    double sr=si(i,j,k)+sbias;
    double tr2=tr*tr;
    double tr3=tr2*tr;
    double tr4=tr3*tr;
    double p=grav*rhoref*(-zz(k)* h(i,j))*1.e-5;
    double rhor=-0.157406+6.793952e-2*tr
                -9.095290e-3*tr2+1.001685e-4*tr3
                -1.120083e-6*tr4+6.536332e-9*tr4*tr;
    rhor=rhor+(0.824493-4.0899e-3*tr
               +7.6438e-5*tr2-8.2467e-7*tr3
               +5.3875e-9*tr4)*sr
             +(-5.72466e-3+1.0227e-4*tr
               -1.6546e-6*tr2)*pow15(S,fabs(sr))
             +4.8314e-4*sr*sr;
    double cr=1449.1+.0821*p+4.55*tr-.045*tr2
              +1.34*(sr-35.);
    rhor=rhor+1.e5*p/(cr*cr)*(1.-2.*p/(cr*cr));
    rhoo(i,j,k)=rhor/rhoref*fsm(i,j);
  }
#undef si
#undef ti
#undef rhoo
}

/* ------------------------------------------------------------------ */
/* solver.f:1212-1538 profq */
void pomo_profq(pomo_t *S) {
  DIMS;
  double *ap = S->scr3[0], *cp = S->scr3[1], *eep = S->scr3[2], *ggp = S->scr3[3];
  double *smp = S->scr3[4], *shp = S->scr3[5], *ccp = S->scr3[6], *ghp = S->scr3[7];
  double *boygrp = S->scr3[8], *stfp = S->scr3[9], *prodp = S->scr3[10];
  double *dhp = S->scr2[0], *l0p = S->scr2[1], *utau2p = S->scr2[2];
#define a(i, j, k) (ap[I3(i, j, k)])
#define c(i, j, k) (cp[I3(i, j, k)])
#define ee(i, j, k) (eep[I3(i, j, k)])
#define gg(i, j, k) (ggp[I3(i, j, k)])
#define sm(i, j, k) (smp[I3(i, j, k)])
#define sh(i, j, k) (shp[I3(i, j, k)])
#define cc(i, j, k) (ccp[I3(i, j, k)])
#define gh(i, j, k) (ghp[I3(i, j, k)])
#define boygr(i, j, k) (boygrp[I3(i, j, k)])
#define stf(i, j, k) (stfp[I3(i, j, k)])
#define prod(i, j, k) (prodp[I3(i, j, k)])
#define dh(i, j) (dhp[I2(i, j)])
#define l0(i, j) (l0p[I2(i, j)])
#define utau2(i, j) (utau2p[I2(i, j)])
  /* :1241-1244 */
  const double a1 = 0.92, b1 = 16.6, a2 = 0.74, b2 = 10.1, c1 = 0.08;
  const double e1 = 1.8, e2 = 1.33;
  const double sef = 1.;
  const double cbcnst = 100., surfl = 2.e5, shiw = 0.;
  double coef1, coef2, coef3, coef4, coef5, const1, ghc;
  /* :1246-1250 */
  DO(j, 1, jm) DO(i, 1, im) dh(i,j)=h(i,j)+etf(i,j);
  /* :1252-1256 */
  zero3(S, ap); zero3(S, cp); zero3(S, eep); zero3(S, ggp); zero2(S, utau2p);
  /* :1258-1267 */
  OMP_FOR
  DO(k, 2, kbm1) DO(j, 1, jm) DO(i, 1, im) {
    a(i,j,k)=-dti2*(kq(i,j,k+1)+kq(i,j,k)+2.*umol)*.5
             /(dzz(k-1)*dz(k)*dh(i,j)*dh(i,j));
    c(i,j,k)=-dti2*(kq(i,j,k-1)+kq(i,j,k)+2.*umol)*.5
             /(dzz(k-1)*dz(k-1)*dh(i,j)*dh(i,j));
  }
  /* :1273 */
  const1=(pow(16.6,2./3.))*sef;
  /* :1277-1279 */
  zero2(S, l0p); zero3(S, boygrp); zero3(S, prodp);
  /* :1281-1288 */
  DO(j, 1, jmm1) DO(i, 1, imm1) {
    utau2(i,j)=sqrt((.5*(wusurf(i,j)+wusurf(i+1,j)))*(.5*(wusurf(i,j)+wusurf(i+1,j)))
                   +(.5*(wvsurf(i,j)+wvsurf(i,j+1)))*(.5*(wvsurf(i,j)+wvsurf(i,j+1))));
    uf(i,j,kb)=sqrt((.5*(wubot(i,j)+wubot(i+1,j)))*(.5*(wubot(i,j)+wubot(i+1,j)))
                   +(.5*(wvbot(i,j)+wvbot(i,j+1)))*(.5*(wvbot(i,j)+wvbot(i,j+1))))*const1;
  }
  /* :1292-1301; NB single-precision literals 15.8 and 2./3. */
  {
    const double c158 = (double)15.8f * cbcnst;
    const double e23 = (double)(2.f / 3.f);
    const double cgg = pow(c158, e23);
    DO(j, 1, jm) DO(i, 1, im) {
      ee(i,j,1)=0.;
      gg(i,j,1)=cgg*utau2(i,j);
      l0(i,j)=surfl*utau2(i,j)/grav;
    }
  }
  /* :1304-1319 */
  zero3(S, ccp);
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 1, jm) DO(i, 1, im) {
    double tp=t(i,j,k)+tbias;
    double sp=s(i,j,k)+sbias;
    double p=grav*rhoref*(-zz(k)*h(i,j))*1.e-4;
    cc(i,j,k)=1449.1+.00821*p+4.55*tp-.045*(tp*tp)
              +1.34*(sp-35.0);
    cc(i,j,k)=cc(i,j,k)
              /sqrt((1.-.01642*p/cc(i,j,k))
                *(1.-0.40*p/(cc(i,j,k)*cc(i,j,k))));
  }
  /* :1322-1333 */
  OMP_FOR
  DO(k, 2, kbm1) DO(j, 1, jm) DO(i, 1, im) {
    q2b(i,j,k)=fabs(q2b(i,j,k));
    q2lb(i,j,k)=fabs(q2lb(i,j,k));
    boygr(i,j,k)=grav*(rho(i,j,k-1)-rho(i,j,k))
                 /(dzz(k-1)*h(i,j))
         +(grav*grav)*2./(cc(i,j,k-1)*cc(i,j,k-1)+cc(i,j,k)*cc(i,j,k));
  }
  /* :1335-1347 */
  OMP_FOR
  DO(k, 2, kbm1) DO(j, 1, jm) DO(i, 1, im) {
    l(i,j,k)=fabs(q2lb(i,j,k)/q2b(i,j,k));
    if (z(k) > -0.5) l(i,j,k)=fmax(l(i,j,k),kappa*l0(i,j));
    gh(i,j,k)=(l(i,j,k)*l(i,j,k))*boygr(i,j,k)/q2b(i,j,k);
    gh(i,j,k)=fmin(gh(i,j,k),.028);
  }
  /* :1349-1356 */
  DO(j, 1, jm) DO(i, 1, im) {
    l(i,j,1)=kappa*l0(i,j);
    l(i,j,kb)=0.;
    gh(i,j,1)=0.;
    gh(i,j,kb)=0.;
  }
  /* :1359-1373 */
  OMP_FOR
  DO(k, 2, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1) {
    double su=u(i,j,k)-u(i,j,k-1)+u(i+1,j,k)-u(i+1,j,k-1);
    double sv=v(i,j,k)-v(i,j,k-1)+v(i,j+1,k)-v(i,j+1,k-1);
    double dd=dzz(k-1)*dh(i,j);
    prod(i,j,k)=km(i,j,k)*.25*sef
                *(su*su+sv*sv)
                /(dd*dd)
                -shiw*km(i,j,k)*boygr(i,j,k);
    prod(i,j,k)=prod(i,j,k)+kh(i,j,k)*boygr(i,j,k);
  }
  /* :1379-1392 */
  ghc=-6.0; (void)ghc;
  OMP_FOR
  DO(k, 1, kb) DO(j, 1, jm) DO(i, 1, im) {
    stf(i,j,k)=1.;
    dtef(i,j,k)=sqrt(fabs(q2b(i,j,k)))*stf(i,j,k)
                /(b1*l(i,j,k)+small);
  }
  /* :1394-1404 */
  DO(k, 2, kbm1) {
    OMP_FOR
    DO(j, 1, jm) DO(i, 1, im) {
      gg(i,j,k)=1./(a(i,j,k)+c(i,j,k)*(1.-ee(i,j,k-1))
                    -(2.*dti2*dtef(i,j,k)+1.));
      ee(i,j,k)=a(i,j,k)*gg(i,j,k);
      gg(i,j,k)=(-2.*dti2*prod(i,j,k)+c(i,j,k)*gg(i,j,k-1)
                 -uf(i,j,k))*gg(i,j,k);
    }
  }
  /* :1406-1413 */
  DO(k, 1, kbm1) {
    int ki=kb-k;
    OMP_FOR
    DO(j, 1, jm) DO(i, 1, im)
      uf(i,j,ki)=ee(i,j,ki)*uf(i,j,ki+1)+gg(i,j,ki);
  }
  /* :1417-1425 */
  DO(j, 1, jm) DO(i, 1, im) {
    vf(i,j,1)=0.;
    vf(i,j,kb)=0.;
    ee(i,j,2)=0.;
    gg(i,j,2)=-kappa*z(2)*dh(i,j)*q2(i,j,2);
    vf(i,j,kb-1)=kappa*(1+z(kbm1))*dh(i,j)*q2(i,j,kbm1);
  }
  /* :1426-1435 */
  OMP_FOR
  DO(k, 2, kbm1) DO(j, 1, jm) DO(i, 1, im) {
    double r=(1./fabs(z(k)-z(1))
              +1./fabs(z(k)-z(kb)))
             *l(i,j,k)/(dh(i,j)*kappa);
    dtef(i,j,k)=dtef(i,j,k)
                *(1.+e2*(r*r));
  }
  /* :1436-1446 */
  DO(k, 3, kbm1) {
    OMP_FOR
    DO(j, 1, jm) DO(i, 1, im) {
      gg(i,j,k)=1./(a(i,j,k)+c(i,j,k)*(1.-ee(i,j,k-1))
                    -(dti2*dtef(i,j,k)+1.));
      ee(i,j,k)=a(i,j,k)*gg(i,j,k);
      gg(i,j,k)=(dti2*(-prod(i,j,k)*l(i,j,k)*e1)
                 +c(i,j,k)*gg(i,j,k-1)-vf(i,j,k))*gg(i,j,k);
    }
  }
  /* :1448-1455 */
  DO(k, 1, kb-2) {
    int ki=kb-k;
    OMP_FOR
    DO(j, 1, jm) DO(i, 1, im)
      vf(i,j,ki)=ee(i,j,ki)*vf(i,j,ki+1)+gg(i,j,ki);
  }
  /* :1460-1471 */
  OMP_FOR
  DO(k, 2, kbm1) DO(j, 1, jm) DO(i, 1, im) {
    uf(i,j,k)=fabs(uf(i,j,k));
    vf(i,j,k)=fabs(vf(i,j,k));
  }
  /* :1474-1475 */
  coef4=18.*a1*a1+9.*a1*a2;
  coef5=9.*a1*a2;
  /* :1478-1489 */
#pragma omp parallel for schedule(static) private(coef1, coef2, coef3)
  DO(k, 1, kb) DO(j, 1, jm) DO(i, 1, im) {
    coef1=a2*(1.-6.*a1/b1*stf(i,j,k));
    coef2=3.*a2*b2/stf(i,j,k)+18.*a1*a2;
    coef3=a1*(1.-3.*c1-6.*a1/b1*stf(i,j,k));
    sh(i,j,k)=coef1/(1.-coef2*gh(i,j,k));
    sm(i,j,k)=coef3+sh(i,j,k)*coef4*gh(i,j,k);
    sm(i,j,k)=sm(i,j,k)/(1.-coef5*gh(i,j,k));
  }
  /* :1496-1506 */
  OMP_FOR
  DO(k, 1, kb) DO(j, 1, jm) DO(i, 1, im) {
    prod(i,j,k)=l(i,j,k)*sqrt(fabs(q2(i,j,k)));
    kq(i,j,k)=(prod(i,j,k)*.41*sh(i,j,k)+kq(i,j,k))*.5;
    km(i,j,k)=(prod(i,j,k)*sm(i,j,k)+km(i,j,k))*.5;
    kh(i,j,k)=(prod(i,j,k)*sh(i,j,k)+kh(i,j,k))*.5;
  }
  /* :1510-1529 */
  if (n_north == -1) DO(k, 1, kb) DO(i, 1, im) {
    km(i,jm,k)=km(i,jmm1,k); kh(i,jm,k)=kh(i,jmm1,k); kq(i,jm,k)=kq(i,jmm1,k); }
  if (n_south == -1) DO(k, 1, kb) DO(i, 1, im) {
    km(i,1,k)=km(i,2,k); kh(i,1,k)=kh(i,2,k); kq(i,1,k)=kq(i,2,k); }
  if (n_east == -1) DO(k, 1, kb) DO(j, 1, jm) {
    km(im,j,k)=km(imm1,j,k); kh(im,j,k)=kh(imm1,j,k); kq(im,j,k)=kq(imm1,j,k); }
  if (n_west == -1) DO(k, 1, kb) DO(j, 1, jm) {
    km(1,j,k)=km(2,j,k); kh(1,j,k)=kh(2,j,k); kq(1,j,k)=kq(2,j,k); }
  /* :1531-1535 */
  OMP_FOR
  DO(k, 1, kb) DO(j, 1, jm) DO(i, 1, im) {
    km(i,j,k)=km(i,j,k)*fsm(i,j);
    kh(i,j,k)=kh(i,j,k)*fsm(i,j);
    kq(i,j,k)=kq(i,j,k)*fsm(i,j);
  }
#undef a
#undef c
#undef ee
#undef gg
#undef sm
#undef sh
#undef cc
#undef gh
#undef boygr
#undef stf
#undef prod
#undef dh
#undef l0
#undef utau2
}

/* ------------------------------------------------------------------ */
/* solver.f:1541-1683 proft */
void pomo_proft(pomo_t *S, double *fp, double *wfsurfp, double *fsurfp, int nbc) {
  DIMS;
  double *ap = S->scr3[0], *cp = S->scr3[1], *eep = S->scr3[2], *ggp = S->scr3[3];
  double *radp = S->scr3[4];
  double *dhp = S->scr2[0];
#define a(i, j, k) (ap[I3(i, j, k)])
#define c(i, j, k) (cp[I3(i, j, k)])
#define ee(i, j, k) (eep[I3(i, j, k)])
#define gg(i, j, k) (ggp[I3(i, j, k)])
#define rad(i, j, k) (radp[I3(i, j, k)])
#define dh(i, j) (dhp[I2(i, j)])
#define f(i, j, k) (fp[I3(i, j, k)])
#define wfsurf(i, j) (wfsurfp[I2(i, j)])
#define fsurf(i, j) (fsurfp[I2(i, j)])
  /* :1561-1563 */
  static const double r[5] = {.58, .62, .67, .77, .78};
  static const double ad1[5] = {.35, .60, 1.0, 1.5, 1.4};
  static const double ad2[5] = {23., 20., 17., 14., 7.9};
  const int ntp = S->ntp;
  /* :1578-1582 */
  DO(j, 1, jm) DO(i, 1, im) dh(i,j)=h(i,j)+etf(i,j);
  /* :1584-1587 */
  zero3(S, ap); zero3(S, cp); zero3(S, eep); zero3(S, ggp);
  /* :1589-1598 */
  OMP_FOR
  DO(k, 2, kbm1) DO(j, 1, jm) DO(i, 1, im) {
    a(i,j,k-1)=-dti2*(kh(i,j,k)+umol)
               /(dz(k-1)*dzz(k-1)*dh(i,j)*dh(i,j));
    c(i,j,k)=-dti2*(kh(i,j,k)+umol)
             /(dz(k)*dzz(k-1)*dh(i,j)*dh(i,j));
  }
  /* :1602-1615; quad-precision exp (real(..,16)) */
  zero3(S, radp);
  if (nbc == 2 || nbc == 4) {
    OMP_FOR
    DO(k, 1, kbm1) DO(j, 1, jm) DO(i, 1, im) {
      __float128 e1q = (__float128)(z(k)*dh(i,j)/ad1[ntp-1]);
      __float128 e2q = (__float128)(z(k)*dh(i,j)/ad2[ntp-1]);
      rad(i,j,k)=(double)((__float128)swrad(i,j)
                 *((__float128)r[ntp-1]*expq(e1q)
                  +(__float128)(1.-r[ntp-1])*expq(e2q)));
    }
  }
  /* :1617-1648 */
  if (nbc == 1) {
    DO(j, 1, jm) DO(i, 1, im) {
      ee(i,j,1)=a(i,j,1)/(a(i,j,1)-1.);
      gg(i,j,1)=dti2*wfsurf(i,j)/(dz(1)*dh(i,j))-f(i,j,1);
      gg(i,j,1)=gg(i,j,1)/(a(i,j,1)-1.);
    }
  } else if (nbc == 2) {
    DO(j, 1, jm) DO(i, 1, im) {
      ee(i,j,1)=a(i,j,1)/(a(i,j,1)-1.);
      gg(i,j,1)=dti2*(wfsurf(i,j)+rad(i,j,1)-rad(i,j,2))
                /(dz(1)*dh(i,j))
                  -f(i,j,1);
      gg(i,j,1)=gg(i,j,1)/(a(i,j,1)-1.);
    }
  } else if (nbc == 3 || nbc == 4) {
    DO(j, 1, jm) DO(i, 1, im) {
      ee(i,j,1)=0.;
      gg(i,j,1)=fsurf(i,j);
    }
  }
  /* :1650-1661 */
  DO(k, 2, kbm2) {
    OMP_FOR
    DO(j, 1, jm) DO(i, 1, im) {
      gg(i,j,k)=1./(a(i,j,k)+c(i,j,k)*(1.-ee(i,j,k-1))-1.);
      ee(i,j,k)=a(i,j,k)*gg(i,j,k);
      gg(i,j,k)=(c(i,j,k)*gg(i,j,k-1)-f(i,j,k)
                 +dti2*(rad(i,j,k)-rad(i,j,k+1))
                   /(dh(i,j)*dz(k)))
                *gg(i,j,k);
    }
  }
  /* :1664-1671 */
  DO(j, 1, jm) DO(i, 1, im)
    f(i,j,kbm1)=(c(i,j,kbm1)*gg(i,j,kbm2)-f(i,j,kbm1)
                 +dti2*(rad(i,j,kbm1)-rad(i,j,kb))
                   /(dh(i,j)*dz(kbm1)))
                /(c(i,j,kbm1)*(1.-ee(i,j,kbm2))-1.);
  /* :1673-1680 */
  DO(k, 2, kbm1) {
    int ki=kb-k;
    OMP_FOR
    DO(j, 1, jm) DO(i, 1, im)
      f(i,j,ki)=(ee(i,j,ki)*f(i,j,ki+1)+gg(i,j,ki));
  }
#undef a
#undef c
#undef ee
#undef gg
#undef rad
#undef dh
#undef f
#undef wfsurf
#undef fsurf
}

/* ------------------------------------------------------------------ */
/* solver.f:1686-1780 profu */
void pomo_profu(pomo_t *S) {
  DIMS;
  double *ap = S->scr3[0], *cp = S->scr3[1], *eep = S->scr3[2], *ggp = S->scr3[3];
  double *dhp = S->scr2[0];
#define a(i, j, k) (ap[I3(i, j, k)])
#define c(i, j, k) (cp[I3(i, j, k)])
#define ee(i, j, k) (eep[I3(i, j, k)])
#define gg(i, j, k) (ggp[I3(i, j, k)])
#define dh(i, j) (dhp[I2(i, j)])
  /* :1699-1705 */
  DO(j, 1, jm) DO(i, 1, im) dh(i,j)=1.;
  DO(j, 2, jm) DO(i, 2, im)
    dh(i,j)=(h(i,j)+etf(i,j)+h(i-1,j)+etf(i-1,j))*.5;
  /* :1707-1710 */
  zero3(S, ap); zero3(S, cp); zero3(S, eep); zero3(S, ggp);
  /* :1712-1718 */
  OMP_FOR
  DO(k, 1, kb) DO(j, 2, jm) DO(i, 2, im)
    c(i,j,k)=(km(i,j,k)+km(i-1,j,k))*.5;
  /* :1720-1729 */
  OMP_FOR
  DO(k, 2, kbm1) DO(j, 1, jm) DO(i, 1, im) {
    a(i,j,k-1)=-dti2*(c(i,j,k)+umol)
               /(dz(k-1)*dzz(k-1)*dh(i,j)*dh(i,j));
    c(i,j,k)=-dti2*(c(i,j,k)+umol)
             /(dz(k)*dzz(k-1)*dh(i,j)*dh(i,j));
  }
  /* :1731-1738 */
  DO(j, 1, jm) DO(i, 1, im) {
    ee(i,j,1)=a(i,j,1)/(a(i,j,1)-1.);
    gg(i,j,1)=(-dti2*wusurf(i,j)/(-dz(1)*dh(i,j))
               -uf(i,j,1))
              /(a(i,j,1)-1.);
  }
  /* :1740-1748 */
  DO(k, 2, kbm2) {
    OMP_FOR
    DO(j, 1, jm) DO(i, 1, im) {
      gg(i,j,k)=1./(a(i,j,k)+c(i,j,k)*(1.-ee(i,j,k-1))-1.);
      ee(i,j,k)=a(i,j,k)*gg(i,j,k);
      gg(i,j,k)=(c(i,j,k)*gg(i,j,k-1)-uf(i,j,k))*gg(i,j,k);
    }
  }
  /* :1750-1761 */
  DO(j, 2, jmm1) DO(i, 2, imm1) {
    double vbar=.25*(vb(i,j,kbm1)+vb(i,j+1,kbm1)
                     +vb(i-1,j,kbm1)+vb(i-1,j+1,kbm1));
    tps(i,j)=0.5*(cbc(i,j)+cbc(i-1,j))
             *sqrt(ub(i,j,kbm1)*ub(i,j,kbm1)
               +vbar*vbar);
    uf(i,j,kbm1)=(c(i,j,kbm1)*gg(i,j,kbm2)-uf(i,j,kbm1))
                 /(tps(i,j)*dti2/(-dz(kbm1)*dh(i,j))-1.
                   -(ee(i,j,kbm2)-1.)*c(i,j,kbm1));
    uf(i,j,kbm1)=uf(i,j,kbm1)*dum(i,j);
  }
  /* :1763-1770 */
  DO(k, 2, kbm1) {
    int ki=kb-k;
    OMP_FOR
    DO(j, 2, jmm1) DO(i, 2, imm1)
      uf(i,j,ki)=(ee(i,j,ki)*uf(i,j,ki+1)+gg(i,j,ki))*dum(i,j);
  }
  /* :1772-1776 */
  DO(j, 2, jmm1) DO(i, 2, imm1)
    wubot(i,j)=-tps(i,j)*uf(i,j,kbm1);
#undef a
#undef c
#undef ee
#undef gg
#undef dh
}

/* solver.f:1783-1877 profv */
void pomo_profv(pomo_t *S) {
  DIMS;
  double *ap = S->scr3[0], *cp = S->scr3[1], *eep = S->scr3[2], *ggp = S->scr3[3];
  double *dhp = S->scr2[0];
#define a(i, j, k) (ap[I3(i, j, k)])
#define c(i, j, k) (cp[I3(i, j, k)])
#define ee(i, j, k) (eep[I3(i, j, k)])
#define gg(i, j, k) (ggp[I3(i, j, k)])
#define dh(i, j) (dhp[I2(i, j)])
  /* :1797-1803 */
  DO(j, 1, jm) DO(i, 1, im) dh(i,j)=1.;
  DO(j, 2, jm) DO(i, 2, im)
    dh(i,j)=.5*(h(i,j)+etf(i,j)+h(i,j-1)+etf(i,j-1));
  /* :1805-1808 */
  zero3(S, ap); zero3(S, cp); zero3(S, eep); zero3(S, ggp);
  /* :1810-1816 */
  OMP_FOR
  DO(k, 1, kb) DO(j, 2, jm) DO(i, 2, im)
    c(i,j,k)=(km(i,j,k)+km(i,j-1,k))*.5;
  /* :1818-1827 */
  OMP_FOR
  DO(k, 2, kbm1) DO(j, 1, jm) DO(i, 1, im) {
    a(i,j,k-1)=-dti2*(c(i,j,k)+umol)
               /(dz(k-1)*dzz(k-1)*dh(i,j)*dh(i,j));
    c(i,j,k)=-dti2*(c(i,j,k)+umol)
             /(dz(k)*dzz(k-1)*dh(i,j)*dh(i,j));
  }
  /* :1829-1835 */
  DO(j, 1, jm) DO(i, 1, im) {
    ee(i,j,1)=a(i,j,1)/(a(i,j,1)-1.);
    gg(i,j,1)=(-dti2*wvsurf(i,j)/(-dz(1)*dh(i,j))-vf(i,j,1))
              /(a(i,j,1)-1.);
  }
  /* :1837-1845 */
  DO(k, 2, kbm2) {
    OMP_FOR
    DO(j, 1, jm) DO(i, 1, im) {
      gg(i,j,k)=1./(a(i,j,k)+c(i,j,k)*(1.-ee(i,j,k-1))-1.);
      ee(i,j,k)=a(i,j,k)*gg(i,j,k);
      gg(i,j,k)=(c(i,j,k)*gg(i,j,k-1)-vf(i,j,k))*gg(i,j,k);
    }
  }
  /* :1847-1858 */
  DO(j, 2, jmm1) DO(i, 2, imm1) {
    double ubar=.25*(ub(i,j,kbm1)+ub(i+1,j,kbm1)
                     +ub(i,j-1,kbm1)+ub(i+1,j-1,kbm1));
    tps(i,j)=0.5*(cbc(i,j)+cbc(i,j-1))
             *sqrt(ubar*ubar
                   +vb(i,j,kbm1)*vb(i,j,kbm1));
    vf(i,j,kbm1)=(c(i,j,kbm1)*gg(i,j,kbm2)-vf(i,j,kbm1))
                 /(tps(i,j)*dti2/(-dz(kbm1)*dh(i,j))-1.
                   -(ee(i,j,kbm2)-1.)*c(i,j,kbm1));
    vf(i,j,kbm1)=vf(i,j,kbm1)*dvm(i,j);
  }
  /* :1860-1867 */
  DO(k, 2, kbm1) {
    int ki=kb-k;
    OMP_FOR
    DO(j, 2, jmm1) DO(i, 2, imm1)
      vf(i,j,ki)=(ee(i,j,ki)*vf(i,j,ki+1)+gg(i,j,ki))*dvm(i,j);
  }
  /* :1869-1873 */
  DO(j, 2, jmm1) DO(i, 2, imm1)
    wvbot(i,j)=-tps(i,j)*vf(i,j,kbm1);
#undef a
#undef c
#undef ee
#undef gg
#undef dh
}

/* ------------------------------------------------------------------ */
/* solver.f:1970-2021 vertvl */
void pomo_vertvl(pomo_t *S) {
  DIMS;
  double *xfluxp = S->scr3[0], *yfluxp = S->scr3[1];
#define xflux(i, j, k) (xfluxp[I3(i, j, k)])
#define yflux(i, j, k) (yfluxp[I3(i, j, k)])
  zero3(S, xfluxp); zero3(S, yfluxp);
  /* :1981-1997 */
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jm) DO(i, 2, im)
    xflux(i,j,k)=.25*(dy(i,j)+dy(i-1,j))
                 *(dt(i,j)+dt(i-1,j))*u(i,j,k);
  OMP_FOR
  DO(k, 1, kbm1) DO(j, 2, jm) DO(i, 2, im)
    yflux(i,j,k)=.25*(dx(i,j)+dx(i,j-1))
                 *(dt(i,j)+dt(i,j-1))*v(i,j,k);
  /* :2002-2006 */
  DO(j, 2, jmm1) DO(i, 2, imm1)
    w(i,j,1)=0.5*(vfluxb(i,j)+vfluxf(i,j));
  /* :2008-2018 */
  DO(k, 1, kbm1) {
    OMP_FOR
    DO(j, 2, jmm1) DO(i, 2, imm1)
      w(i,j,k+1)=w(i,j,k)
                 +dz(k)*((xflux(i+1,j,k)-xflux(i,j,k)
                         +yflux(i,j+1,k)-yflux(i,j,k))
                         /(dx(i,j)*dy(i,j))
                         +(etf(i,j)-etb(i,j))/dti2);
  }
#undef xflux
#undef yflux
}

/* solver.f:2024-2066 realvertvl */
void pomo_realvertvl(pomo_t *S) {
  DIMS;
  zero3(S, S->wr); /* :2031 */
  /* :2033-2053 */
  DO(k, 1, kbm1) {
    DO(j, 1, jm) DO(i, 1, im)
      tps(i,j)=zz(k)*dt(i,j) + et(i,j);
    OMP_FOR
    DO(j, 2, jmm1) DO(i, 2, imm1) {
      double dxr=2.0/(dx(i+1,j)+dx(i,j));
      double dxl=2.0/(dx(i,j)+dx(i-1,j));
      double dyt=2.0/(dy(i,j+1)+dy(i,j));
      double dyb=2.0/(dy(i,j)+dy(i,j-1));
      wr(i,j,k)=0.5*(w(i,j,k)+w(i,j,k+1))+0.5*
                (u(i+1,j,k)*(tps(i+1,j)-tps(i,j))*dxr+
                 u(i,j,k)*(tps(i,j)-tps(i-1,j))*dxl+
                 v(i,j+1,k)*(tps(i,j+1)-tps(i,j))*dyt+
                 v(i,j,k)*(tps(i,j)-tps(i,j-1))*dyb)
                +(1.0+zz(k))*(etf(i,j)-etb(i,j))/dti2;
    }
  }
  /* :2057-2060 (all k=1..kb) */
  if (n_south == -1) DO(k, 1, kb) DO(i, 1, im) wr(i,1,k)=wr(i,2,k);
  if (n_north == -1) DO(k, 1, kb) DO(i, 1, im) wr(i,jm,k)=wr(i,jmm1,k);
  if (n_west == -1) DO(k, 1, kb) DO(j, 1, jm) wr(1,j,k)=wr(2,j,k);
  if (n_east == -1) DO(k, 1, kb) DO(j, 1, jm) wr(im,j,k)=wr(imm1,j,k);
  /* :2062-2064 */
  DO(k, 1, kbm1) DO(j, 1, jm) DO(i, 1, im) wr(i,j,k)=fsm(i,j)*wr(i,j,k);
}
