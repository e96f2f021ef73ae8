/* pomo_advance.c -- CPU ORACLE restatement of pom/advance.f:96-537
 * (lateral_viscosity, mode_interaction, mode_external, mode_internal) and
 * the hot-path part of advance (advance.f:21-32).
 * TEST INFRASTRUCTURE ONLY; parity pinned against the reference's own source (see pomo.h). */
#define POMO_IMPL
#include "pomo.h"
#include <math.h>
#include <stdio.h>
#include <string.h>

/* advance.f:96-141 */
void pomo_lateral_viscosity(pomo_t *S) {
  DIMS;
  if (S->mode != 2) {
    pomo_advct(S);
    if (S->npg == 1) {
      pomo_baropg(S);
    } else if (S->npg == 2) {
      pomo_baropg_mcc(S);
    } else {
      S->error_status = 1;
      fprintf(stderr, "\nError: invalid value for npg\n");
    }
    /* :122-136 */
    OMP_FOR
    DO(k, 1, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1) {
      double a1=(u(i+1,j,k)-u(i,j,k))/dx(i,j);
      double a2=(v(i,j+1,k)-v(i,j,k))/dy(i,j);
      double a3=.25*(u(i,j+1,k)+u(i+1,j+1,k)
                     -u(i,j-1,k)-u(i+1,j-1,k))
                /dy(i,j)
               +.25*(v(i+1,j,k)+v(i+1,j+1,k)
                     -v(i-1,j,k)-v(i-1,j+1,k))
                /dx(i,j);
      aam(i,j,k)=horcon*dx(i,j)*dy(i,j)
                 *sqrt( a1*a1
                       +a2*a2
                 +.5*(a3*a3));
    }
  }
}

/* advance.f:144-202 */
void pomo_mode_interaction(pomo_t *S) {
  DIMS;
  if (S->mode != 2) {
    memset(S->adx2d, 0, sizeof(double) * N2);
    memset(S->ady2d, 0, sizeof(double) * N2);
    memset(S->drx2d, 0, sizeof(double) * N2);
    memset(S->dry2d, 0, sizeof(double) * N2);
    memset(S->aam2d, 0, sizeof(double) * N2);
    /* :158-168 */
    DO(k, 1, kbm1) {
      OMP_FOR
      DO(j, 1, jm) DO(i, 1, im) {
        adx2d(i,j)=adx2d(i,j)+advx(i,j,k)*dz(k);
        ady2d(i,j)=ady2d(i,j)+advy(i,j,k)*dz(k);
        drx2d(i,j)=drx2d(i,j)+drhox(i,j,k)*dz(k);
        dry2d(i,j)=dry2d(i,j)+drhoy(i,j,k)*dz(k);
        aam2d(i,j)=aam2d(i,j)+aam(i,j,k)*dz(k);
      }
    }
    pomo_advave(S);
    /* :172-177 */
    OMP_FOR
    DO(j, 1, jm) DO(i, 1, im) {
      adx2d(i,j)=adx2d(i,j)-advua(i,j);
      ady2d(i,j)=ady2d(i,j)-advva(i,j);
    }
  }
  /* :181-196 */
  OMP_FOR
  DO(j, 1, jm) DO(i, 1, im) egf(i,j)=el(i,j)*ispi;
  OMP_FOR
  DO(j, 1, jm) DO(i, 2, im) utf(i,j)=ua(i,j)*(d(i,j)+d(i-1,j))*isp2i;
  OMP_FOR
  DO(j, 2, jm) DO(i, 1, im) vtf(i,j)=va(i,j)*(d(i,j)+d(i,j-1))*isp2i;
}

/* advance.f:205-353 */
void pomo_mode_external(pomo_t *S) {
  DIMS;
  const int iext = S->iext, isplit = S->isplit;
  /* :211-218 */
  OMP_FOR
  DO(j, 2, jm) DO(i, 2, im) {
    fluxua(i,j)=.25*(d(i,j)+d(i-1,j))
                *(dy(i,j)+dy(i-1,j))*ua(i,j);
    fluxva(i,j)=.25*(d(i,j)+d(i,j-1))
                *(dx(i,j)+dx(i,j-1))*va(i,j);
  }
  /* :222-229 */
  OMP_FOR
  DO(j, 2, jmm1) DO(i, 2, imm1)
    elf(i,j)=elb(i,j)
             +dte2*(-(fluxua(i+1,j)-fluxua(i,j)
                     +fluxva(i,j+1)-fluxva(i,j))/art(i,j)
                     -vfluxf(i,j));
  pomo_bcond(S, 1); /* :231 */
  if (iext % S->ispadv == 0) pomo_advave(S); /* :235 */
  /* :237-252 */
  OMP_FOR
  DO(j, 2, jmm1) DO(i, 2, im)
    uaf(i,j)=adx2d(i,j)+advua(i,j)
             -aru(i,j)*.25
               *(cor(i,j)*d(i,j)*(va(i,j+1)+va(i,j))
                +cor(i-1,j)*d(i-1,j)*(va(i-1,j+1)+va(i-1,j)))
             +.25*grav*(dy(i,j)+dy(i-1,j))
               *(d(i,j)+d(i-1,j))
               *((1.-2.*alpha)
                  *(el(i,j)-el(i-1,j))
                 +alpha*(elb(i,j)-elb(i-1,j)
                        +elf(i,j)-elf(i-1,j))
                 +e_atmos(i,j)-e_atmos(i-1,j))
             +drx2d(i,j)+aru(i,j)*(wusurf(i,j)-wubot(i,j));
  /* :254-262 */
  OMP_FOR
  DO(j, 2, jmm1) DO(i, 2, im)
    uaf(i,j)=((h(i,j)+elb(i,j)+h(i-1,j)+elb(i-1,j))
               *aru(i,j)*uab(i,j)
             -4.*dte*uaf(i,j))
            /((h(i,j)+elf(i,j)+h(i-1,j)+elf(i-1,j))
                *aru(i,j));
  /* :264-278 */
  OMP_FOR
  DO(j, 2, jm) DO(i, 2, imm1)
    vaf(i,j)=ady2d(i,j)+advva(i,j)
             +arv(i,j)*.25
               *(cor(i,j)*d(i,j)*(ua(i+1,j)+ua(i,j))
              +cor(i,j-1)*d(i,j-1)*(ua(i+1,j-1)+ua(i,j-1)))
             +.25*grav*(dx(i,j)+dx(i,j-1))
               *(d(i,j)+d(i,j-1))
               *((1.-2.*alpha)*(el(i,j)-el(i,j-1))
                 +alpha*(elb(i,j)-elb(i,j-1)
                        +elf(i,j)-elf(i,j-1))
                 +e_atmos(i,j)-e_atmos(i,j-1))
             +dry2d(i,j)+arv(i,j)*(wvsurf(i,j)-wvbot(i,j));
  /* :280-288 */
  OMP_FOR
  DO(j, 2, jm) DO(i, 2, imm1)
    vaf(i,j)=((h(i,j)+elb(i,j)+h(i,j-1)+elb(i,j-1))
               *arv(i,j)*vab(i,j)
             -4.*dte*vaf(i,j))
            /((h(i,j)+elf(i,j)+h(i,j-1)+elf(i,j-1))
                *arv(i,j));
  pomo_bcond(S, 2); /* :290 */
  /* :295-318 */
  if (iext == (isplit-2)) {
    OMP_FOR
    DO(j, 1, jm) DO(i, 1, im) S->etf[I2(i,j)]=.25*smoth*elf(i,j);
  } else if (iext == (isplit-1)) {
    OMP_FOR
    DO(j, 1, jm) DO(i, 1, im) S->etf[I2(i,j)]=S->etf[I2(i,j)]+.5*(1.-.5*smoth)*elf(i,j);
  } else if (iext == isplit) {
    OMP_FOR
    DO(j, 1, jm) DO(i, 1, im) S->etf[I2(i,j)]=(S->etf[I2(i,j)]+.5*elf(i,j))*fsm(i,j);
  }
  /* :321-330 whole-array filter and time rotation */
  OMP_FOR
  for (size_t n = 0; n < N2; ++n) {
    S->ua[n] = S->ua[n]+.5*smoth*(S->uab[n]-2.*S->ua[n]+S->uaf[n]);
    S->va[n] = S->va[n]+.5*smoth*(S->vab[n]-2.*S->va[n]+S->vaf[n]);
    S->el[n] = S->el[n]+.5*smoth*(S->elb[n]-2.*S->el[n]+S->elf[n]);
    S->elb[n] = S->el[n];
    S->el[n] = S->elf[n];
    S->d[n] = S->h[n]+S->el[n];
    S->uab[n] = S->ua[n];
    S->ua[n] = S->uaf[n];
    S->vab[n] = S->va[n];
    S->va[n] = S->vaf[n];
  }
  /* :332-350 */
  if (iext != isplit) {
    OMP_FOR
    DO(j, 1, jm) DO(i, 1, im) egf(i,j)=egf(i,j)+el(i,j)*ispi;
    OMP_FOR
    DO(j, 1, jm) DO(i, 2, im) utf(i,j)=utf(i,j)+ua(i,j)*(d(i,j)+d(i-1,j))*isp2i;
    OMP_FOR
    DO(j, 2, jm) DO(i, 1, im) vtf(i,j)=vtf(i,j)+va(i,j)*(d(i,j)+d(i,j-1))*isp2i;
  }
}

/* advance.f:356-537, split into stages so that tests can compare the GPU
 * path block by block (pomo_internal_stage); pomo_mode_internal runs them all. */
static int internal_active(pomo_t *S) {
  return (S->iint != 1 || S->time0 != 0.) && S->mode != 2; /* :362 */
}

void pomo_internal_stage(pomo_t *S, int stage) {
  DIMS;
  switch (stage) {
  case 0: /* :365-393 */
    memset(S->tps, 0, sizeof(double) * N2);
    DO(k, 1, kbm1) OMP_FOR DO(j, 1, jm) DO(i, 1, im) tps(i,j)=tps(i,j)+u(i,j,k)*dz(k);
    OMP_FOR
    DO(k, 1, kbm1) DO(j, 1, jm) DO(i, 2, im)
      u(i,j,k)=(u(i,j,k)-tps(i,j))+
               (utb(i,j)+utf(i,j))/(dt(i,j)+dt(i-1,j));
    memset(S->tps, 0, sizeof(double) * N2);
    DO(k, 1, kbm1) OMP_FOR DO(j, 1, jm) DO(i, 1, im) tps(i,j)=tps(i,j)+v(i,j,k)*dz(k);
    OMP_FOR
    DO(k, 1, kbm1) DO(j, 2, jm) DO(i, 1, im)
      v(i,j,k)=(v(i,j,k)-tps(i,j))+
               (vtb(i,j)+vtf(i,j))/(dt(i,j)+dt(i,j-1));
    break;
  case 1: /* :396-400 */
    pomo_vertvl(S);
    pomo_bcondorl(S, 5);
    break;
  case 2: /* :403-408 */
    memset(S->uf, 0, sizeof(double) * N3);
    memset(S->vf, 0, sizeof(double) * N3);
    pomo_advq(S, S->q2b, S->q2, S->uf);
    pomo_advq(S, S->q2lb, S->q2l, S->vf);
    break;
  case 3: pomo_profq(S); break; /* :409 */
  case 4: /* :414-421 */
    pomo_bcond(S, 6);
    OMP_FOR
    for (size_t n = 0; n < N3; ++n) {
      S->q2[n] = S->q2[n]+.5*smoth*(S->uf[n]+S->q2b[n]-2.*S->q2[n]);
      S->q2l[n] = S->q2l[n]+.5*smoth*(S->vf[n]+S->q2lb[n]-2.*S->q2l[n]);
      S->q2b[n] = S->q2[n];
      S->q2[n] = S->uf[n];
      S->q2lb[n] = S->q2l[n];
      S->q2l[n] = S->vf[n];
    }
    break;
  case 5: /* :425-434 (T) */
    if (S->mode == 4) break;
    if (S->nadv == 1) pomo_advt1(S, S->tb, S->t, S->tclim, S->uf);
    else if (S->nadv == 2) pomo_advt2(S, S->tb, S->t, S->tclim, S->uf);
    else { S->error_status = 1; fprintf(stderr, "\nError: invalid value for nadv\n"); }
    break;
  case 6: /* :425-434 (S) */
    if (S->mode == 4) break;
    if (S->nadv == 1) pomo_advt1(S, S->sb, S->s, S->sclim, S->vf);
    else if (S->nadv == 2) pomo_advt2(S, S->sb, S->s, S->sclim, S->vf);
    break;
  case 7: if (S->mode != 4) pomo_proft(S, S->uf, S->wtsurf, S->tsurf, S->nbct); break; /* :439 */
  case 8: if (S->mode != 4) pomo_proft(S, S->vf, S->wssurf, S->ssurf, S->nbcs); break; /* :440 */
  case 9: /* :442-452 */
    if (S->mode == 4) break;
    pomo_bcond(S, 4);
    OMP_FOR
    for (size_t n = 0; n < N3; ++n) {
      S->t[n] = S->t[n]+.5*smoth*(S->uf[n]+S->tb[n]-2.*S->t[n]);
      S->s[n] = S->s[n]+.5*smoth*(S->vf[n]+S->sb[n]-2.*S->s[n]);
      S->tb[n] = S->t[n];
      S->t[n] = S->uf[n];
      S->sb[n] = S->s[n];
      S->s[n] = S->vf[n];
    }
    pomo_restore_interior(S);
    break;
  case 10: if (S->mode != 4) pomo_dens(S, S->s, S->t, S->rho); break; /* :454 */
  case 11: pomo_advu(S); break; /* :459 */
  case 12: pomo_advv(S); break;
  case 13: pomo_profu(S); break;
  case 14: pomo_profv(S); break;
  case 15: /* :464-514 */
    pomo_bcondorl(S, 3);
    memset(S->tps, 0, sizeof(double) * N2);
    DO(k, 1, kbm1) OMP_FOR DO(j, 1, jm) DO(i, 1, im)
      tps(i,j)=tps(i,j)
               +(uf(i,j,k)+ub(i,j,k)-2.*u(i,j,k))*dz(k);
    OMP_FOR
    DO(k, 1, kbm1) DO(j, 1, jm) DO(i, 1, im)
      u(i,j,k)=u(i,j,k)
               +.5*smoth*(uf(i,j,k)+ub(i,j,k)
                          -2.*u(i,j,k)-tps(i,j));
    memset(S->tps, 0, sizeof(double) * N2);
    DO(k, 1, kbm1) OMP_FOR DO(j, 1, jm) DO(i, 1, im)
      tps(i,j)=tps(i,j)
               +(vf(i,j,k)+vb(i,j,k)-2.*v(i,j,k))*dz(k);
    OMP_FOR
    DO(k, 1, kbm1) DO(j, 1, jm) DO(i, 1, im)
      v(i,j,k)=v(i,j,k)
               +.5*smoth*(vf(i,j,k)+vb(i,j,k)
                          -2.*v(i,j,k)-tps(i,j));
    memcpy(S->ub, S->u, sizeof(double) * N3);
    memcpy(S->u, S->uf, sizeof(double) * N3);
    memcpy(S->vb, S->v, sizeof(double) * N3);
    memcpy(S->v, S->vf, sizeof(double) * N3);
    break;
  case 16: /* :525-531 */
    memcpy(S->egb, S->egf, sizeof(double) * N2);
    memcpy(S->etb, S->et, sizeof(double) * N2);
    memcpy(S->et, S->etf, sizeof(double) * N2);
    for (size_t n = 0; n < N2; ++n) S->dt[n] = S->h[n]+S->et[n];
    memcpy(S->utb, S->utf, sizeof(double) * N2);
    memcpy(S->vtb, S->vtf, sizeof(double) * N2);
    memcpy(S->vfluxb, S->vfluxf, sizeof(double) * N2);
    break;
  case 17: pomo_realvertvl(S); break; /* :534 */
  }
}

void pomo_mode_internal(pomo_t *S) {
  if (internal_active(S))
    for (int st = 0; st <= 15; ++st) pomo_internal_stage(S, st);
  pomo_internal_stage(S, 16);
  pomo_internal_stage(S, 17);
}

/* advance.f:21-32: the hot path of one internal step.  The caller sets
 * iint (and time/ramp, advance.f:62-75) like the Fortran driver does. */
void pomo_step(pomo_t *S) {
  pomo_lateral_viscosity(S);
  pomo_mode_interaction(S);
  for (S->iext = 1; S->iext <= S->isplit; ++S->iext) pomo_mode_external(S);
  S->iext = S->isplit + 1; /* Fortran do-loop exit value */
  pomo_mode_internal(S);
}

/* advance.f:611-641 */
double pomo_check_velocity(pomo_t *S) {
  DIMS;
  double vamax = 0.;
  DO(j, 1, jm) DO(i, 1, im)
    if (fabs(vaf(i,j)) >= vamax) vamax = fabs(vaf(i,j));
  if (vamax > S->vmaxl) S->error_status = 1;
  return vamax;
}

/* advance.f:644-755 domain_stats for one sub-domain (all neighbours -1).  Fortran's sum() is
 * restated as a plain sequential sum in array element order (i fastest). out[8] =
 * vtot, atot, mtot, stot, tavg, savg, eavg, ekin. */
void pomo_domain_stats(pomo_t *S, double *out) {
  DIMS;
  double vtot = 0., atot = 0., mtot = 0., stot = 0., tavg = 0., savg = 0., eavg = 0., ekin = 0.;
#define DAREA(i, j) (dx(i,j)*dy(i,j)*fsm(i,j))
#define INREG(i, j) (((i) >= 2 && (i) <= imm1 && (j) >= 2 && (j) <= jmm1) || (((i) == 1 || (i) == im) && (j) >= 2 && (j) <= jmm1) || (((j) == 1 || (j) == jm) && (i) >= 2 && (i) <= imm1))
  /* :669-680 interior first, then W, E, S, N edges */
  /* every `sum(section)` of the reference is a SEPARATE accumulation from zero that is then added to the total
   * (atot = atot+sum(darea(1,2:jmm1))): found by running the reference source itself (oracle/f77ref.py) -- folding
   * the edge values into the running total changes the last bit */
  { double p;
    DO(j, 2, jmm1) DO(i, 2, imm1) atot += DAREA(i,j);
    p = 0.; DO(j, 2, jmm1) p += DAREA(1,j);  atot = atot + p;
    p = 0.; DO(j, 2, jmm1) p += DAREA(im,j); atot = atot + p;
    p = 0.; DO(i, 2, imm1) p += DAREA(i,1);  atot = atot + p;
    p = 0.; DO(i, 2, imm1) p += DAREA(i,jm); atot = atot + p;
    DO(j, 2, jmm1) DO(i, 2, imm1) eavg += et(i,j)*DAREA(i,j);
    p = 0.; DO(j, 2, jmm1) p += et(1,j)*DAREA(1,j);   eavg = eavg + p;
    p = 0.; DO(j, 2, jmm1) p += et(im,j)*DAREA(im,j); eavg = eavg + p;
    p = 0.; DO(i, 2, imm1) p += et(i,1)*DAREA(i,1);   eavg = eavg + p;
    p = 0.; DO(i, 2, imm1) p += et(i,jm)*DAREA(i,jm); eavg = eavg + p;
  }
  eavg = (atot != 0.) ? eavg / atot : 0.;                     /* :685-691 */
  /* :693-703: dvol is only assigned on the interior (2:imm1,2:jmm1); the edge sums add zeros */
#define DVOL(i, j, k) (DAREA(i,j)*dt(i,j)*dz(k))
  DO(k, 1, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1) vtot += DVOL(i,j,k);
  DO(k, 1, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1) mtot += DVOL(i,j,k)*(rho(i,j,k)*rhoref+1000.);   /* :704-707 */
  DO(k, 1, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1) tavg += tb(i,j,k)*DVOL(i,j,k);                     /* :709-727 */
  DO(k, 1, kbm1) DO(j, 2, jmm1) DO(i, 2, imm1) stot += sb(i,j,k)*DVOL(i,j,k);
  if (vtot != 0.) { tavg = tavg / vtot; savg = stot / vtot; } else { tavg = 0.; savg = 0.; }     /* :731-739 */
  /* :742-747 (dmass is zero outside the interior, so the E and N edge terms vanish) */
  DO(k, 1, kbm1) {
    double sk = 0.;
    DO(j, 2, jmm1) DO(i, 2, imm1)
      sk += DVOL(i,j,k)*(rho(i,j,k)*rhoref+1000.)*(u(i,j,k)*u(i,j,k)+v(i,j,k)*v(i,j,k));
    ekin = ekin + .5*sk;
  }
  out[0] = vtot; out[1] = atot; out[2] = mtot; out[3] = stot; out[4] = tavg; out[5] = savg; out[6] = eavg; out[7] = ekin;
#undef DAREA
#undef DVOL
#undef INREG
}
