/* pomo.h -- CPU parity ORACLE for the extPOM time-stepping hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, link or call it.
 *
 * PARITY PINNED AGAINST THE REFERENCE'S OWN SOURCE.  The reference (RinceWND/extPOM,
 * fixed-form Fortran) ships no tests, no golden vectors and no input data, and cannot be
 * COMPILED in this image (no Fortran compiler, MPI or PnetCDF) -- but its source can be
 * EXECUTED: oracle/f77ref.py reads pom/solver.f, pom/advance.f, pom/bounds_forcing.f and
 * pom.h_dist where they lie under /root/reference, translates them statement by statement
 * (gfortran -O0 semantics: binary64, single-precision literals, integer division, powi) and
 * runs them; scripts/make_ref_golden.py stores its outputs for 14 namelist variants under
 * tests/golden/ref_*.npz, and this restatement equals every one of them BIT FOR BIT
 * (tests/test_oracle.py; plus a live run of the reference source where it is present).
 * This file set is a line-by-line restatement in C of
 *     pom/advance.f:96-537, pom/solver.f:6-940,1162-2067,
 *     pom/bounds_forcing.f:6-328,331-590,1083-1118
 * with 1-based column-major accessor macros named after the Fortran arrays
 * so that every loop bound and expression can be compared with the source
 * side by side.  Arithmetic is IEEE binary64, evaluated left to right with
 * no FMA contraction (build with -ffp-contract=off), mirroring the
 * reference's `mpif90 -O0` build (makefile_dist:17).
 *
 * One sub-domain only (all neighbours -1, parallel_mpi.f:109-119), so every
 * exchange2d_mpi/exchange3d_mpi is a no-op (parallel_mpi.f:171-236) and
 * im_local==im, jm_local==jm.
 */
#ifndef POMO_H
#define POMO_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* X-lists of the COMMON-block members this path touches (pom.h_dist) */
#define POMO_F3D(X) \
  X(aam) X(advx) X(advy) X(drhox) X(drhoy) X(dtef) X(kh) X(km) X(kq) X(l) \
  X(q2b) X(q2) X(q2lb) X(q2l) X(rho) X(rmean) X(sb) X(sclim) X(s) \
  X(tb) X(tclim) X(t) X(ub) X(uf) X(u) X(vb) X(vf) X(v) X(w) X(wr) X(zflux) \
  X(trstr) X(trstrb) X(trstrf) X(srstr) X(srstrb) X(srstrf) \
  X(taurstr) X(taurstrb) X(taurstrf)

#define POMO_F2D(X) \
  X(aam2d) X(advua) X(advva) X(adx2d) X(ady2d) X(art) X(aru) X(arv) X(cbc) \
  X(cor) X(d) X(drx2d) X(dry2d) X(dt) X(dum) X(dvm) X(dx) X(dy) X(e_atmos) \
  X(egb) X(egf) X(el) X(elb) X(elf) X(et) X(etb) X(etf) X(fluxua) X(fluxva) \
  X(fsm) X(h) X(swrad) X(ssurf) X(tsurf) X(tps) X(ua) X(uab) X(uaf) X(utb) \
  X(utf) X(va) X(vab) X(vaf) X(vtb) X(vtf) X(vfluxb) X(vfluxf) X(wssurf) \
  X(wtsurf) X(wubot) X(wusurf) X(wvbot) X(wvsurf) \
  X(wusurfb) X(wusurff) X(wvsurfb) X(wvsurff) X(wtsurfb) X(wtsurff) X(swradb) X(swradf)

/* boundary arrays: (jm) */
#define POMO_BJ(X) X(ele) X(elw) X(uabe) X(uabw) X(vabe) X(vabw)
/* (im) */
#define POMO_BI(X) X(eln) X(els) X(vabn) X(vabs) X(uabn) X(uabs)
/* (jm,kb) */
#define POMO_BJK(X) X(tbe) X(sbe) X(tbw) X(sbw) X(ube) X(ubw) \
  X(tbeb) X(tbef) X(sbeb) X(sbef) X(ubeb) X(ubef) X(tbwb) X(tbwf) X(sbwb) X(sbwf) X(ubwb) X(ubwf)
/* (im,kb) */
#define POMO_BIK(X) X(tbn) X(sbn) X(tbs) X(sbs) X(vbn) X(vbs) \
  X(tbnb) X(tbnf) X(sbnb) X(sbnf) X(vbnb) X(vbnf) X(tbsb) X(tbsf) X(sbsb) X(sbsf) X(vbsb) X(vbsf)
/* (kb) */
#define POMO_F1D(X) X(z) X(zz) X(dz) X(dzz)

/* blkcon scalars used on the path (pom.h_dist:69-198) */
#define POMO_SCAL_D(X) \
  X(alpha) X(dte) X(dti) X(dti2) X(grav) X(kappa) X(ramp) X(rfe) X(rfn) \
  X(rfs) X(rfw) X(rhoref) X(sbias) X(small) X(tbias) X(time) X(tprni) \
  X(umol) X(vmaxl) X(dte2) X(horcon) X(ispi) X(isp2i) X(smoth) X(sw) X(time0)
#define POMO_SCAL_I(X) \
  X(iint) X(mode) X(ntp) X(iext) X(ispadv) X(isplit) X(nadv) X(nbct) X(nbcs) \
  X(nitera) X(npg) X(error_status) X(n_west) X(n_east) X(n_south) X(n_north) \
  X(lrestore) X(pow_mode)

#define POMO_NSCR3 12
#define POMO_NSCR2 6

typedef struct pomo {
  int im, jm, kb, imm1, imm2, jmm1, jmm2, kbm1, kbm2;
#define X(n) double *n;
  POMO_F3D(X) POMO_F2D(X) POMO_BJ(X) POMO_BI(X) POMO_BJK(X) POMO_BIK(X) POMO_F1D(X)
#undef X
#define X(n) double n;
  POMO_SCAL_D(X)
#undef X
#define X(n) int n;
  POMO_SCAL_I(X)
#undef X
  double *scr3[POMO_NSCR3]; /* automatic 3-D temporaries of solver.f */
  double *scr2[POMO_NSCR2];
} pomo_t;

pomo_t *pomo_create(int im, int jm, int kb);
void pomo_destroy(pomo_t *S);
double *pomo_field(pomo_t *S, const char *name, long *n);
int pomo_set(pomo_t *S, const char *name, double v);
double pomo_get(pomo_t *S, const char *name);

/* advance.f */
void pomo_step(pomo_t *S); /* advance.f:21-32 */
void pomo_lateral_viscosity(pomo_t *S);
void pomo_mode_interaction(pomo_t *S);
void pomo_mode_external(pomo_t *S);
void pomo_mode_internal(pomo_t *S);
void pomo_internal_stage(pomo_t *S, int stage); /* blocks of advance.f:356-537 */
double pomo_check_velocity(pomo_t *S); /* advance.f:611-641, returns vamax */
void pomo_baropg_mcc(pomo_t *S);   /* solver.f:943-1159 (npg=2) */
/* the per-step time interpolation of the forcing records (the file reads around it are out of scope) */
void pomo_wind_interp(pomo_t *S, double fnew);        /* bounds_forcing.f:904-909 */
void pomo_heat_interp(pomo_t *S, double fnew);        /* bounds_forcing.f:949-957 */
void pomo_lateral_bc_interp(pomo_t *S, double fnew);  /* bounds_forcing.f:841-865 */
void pomo_domain_stats(pomo_t *S, double *out8); /* advance.f:644-755: vtot atot mtot stot tavg savg eavg ekin */
/* solver.f */
void pomo_advave(pomo_t *S);
void pomo_advct(pomo_t *S);
void pomo_advq(pomo_t *S, double *qb, double *q, double *qf);
void pomo_advt1(pomo_t *S, double *fb, double *f, double *fclim, double *ff);
void pomo_advt2(pomo_t *S, double *fb, double *f, double *fclim, double *ff);
void pomo_advu(pomo_t *S);
void pomo_advv(pomo_t *S);
void pomo_baropg(pomo_t *S);
void pomo_dens(pomo_t *S, double *si, double *ti, double *rhoo);
void pomo_profq(pomo_t *S);
void pomo_proft(pomo_t *S, double *f, double *wfsurf, double *fsurf, int nbc);
void pomo_profu(pomo_t *S);
void pomo_profv(pomo_t *S);
void pomo_smol_adif(pomo_t *S, double *xmassflux, double *ymassflux,
                    double *zwflux, double *ff);
void pomo_vertvl(pomo_t *S);
void pomo_realvertvl(pomo_t *S);
/* bounds_forcing.f */
void pomo_bcond(pomo_t *S, int idx);
void pomo_bcondorl(pomo_t *S, int idx);
void pomo_restore_interior(pomo_t *S);

#ifdef __cplusplus
}
#endif

/* ---- accessor macros (only for the oracle's own .c files) ---- */
#ifdef POMO_IMPL
#define I3(i, j, k) \
  ((size_t)((i)-1) + (size_t)im * ((size_t)((j)-1) + (size_t)jm * (size_t)((k)-1)))
#define I2(i, j) ((size_t)((i)-1) + (size_t)im * (size_t)((j)-1))
#define N3 ((size_t)im * jm * kb)
#define N2 ((size_t)im * jm)
#define DIMS                                                         \
  const int im = S->im, jm = S->jm, kb = S->kb, imm1 = S->imm1,       \
            jmm1 = S->jmm1, kbm1 = S->kbm1, kbm2 = S->kbm2;           \
  (void)imm1; (void)jmm1; (void)kbm1; (void)kbm2; (void)kb; (void)jm; (void)im
/* 3-D COMMON arrays */
#define aam(i, j, k) (S->aam[I3(i, j, k)])
#define advx(i, j, k) (S->advx[I3(i, j, k)])
#define advy(i, j, k) (S->advy[I3(i, j, k)])
#define drhox(i, j, k) (S->drhox[I3(i, j, k)])
#define drhoy(i, j, k) (S->drhoy[I3(i, j, k)])
#define dtef(i, j, k) (S->dtef[I3(i, j, k)])
#define kh(i, j, k) (S->kh[I3(i, j, k)])
#define km(i, j, k) (S->km[I3(i, j, k)])
#define kq(i, j, k) (S->kq[I3(i, j, k)])
#define l(i, j, k) (S->l[I3(i, j, k)])
#define q2b(i, j, k) (S->q2b[I3(i, j, k)])
#define q2(i, j, k) (S->q2[I3(i, j, k)])
#define q2lb(i, j, k) (S->q2lb[I3(i, j, k)])
#define q2l(i, j, k) (S->q2l[I3(i, j, k)])
#define rho(i, j, k) (S->rho[I3(i, j, k)])
#define rmean(i, j, k) (S->rmean[I3(i, j, k)])
#define sb(i, j, k) (S->sb[I3(i, j, k)])
#define s(i, j, k) (S->s[I3(i, j, k)])
#define tb(i, j, k) (S->tb[I3(i, j, k)])
#define t(i, j, k) (S->t[I3(i, j, k)])
#define ub(i, j, k) (S->ub[I3(i, j, k)])
#define uf(i, j, k) (S->uf[I3(i, j, k)])
#define u(i, j, k) (S->u[I3(i, j, k)])
#define vb(i, j, k) (S->vb[I3(i, j, k)])
#define vf(i, j, k) (S->vf[I3(i, j, k)])
#define v(i, j, k) (S->v[I3(i, j, k)])
#define w(i, j, k) (S->w[I3(i, j, k)])
#define wr(i, j, k) (S->wr[I3(i, j, k)])
#define zflux(i, j, k) (S->zflux[I3(i, j, k)])
#define trstr(i, j, k) (S->trstr[I3(i, j, k)])
#define trstrb(i, j, k) (S->trstrb[I3(i, j, k)])
#define trstrf(i, j, k) (S->trstrf[I3(i, j, k)])
#define srstr(i, j, k) (S->srstr[I3(i, j, k)])
#define srstrb(i, j, k) (S->srstrb[I3(i, j, k)])
#define srstrf(i, j, k) (S->srstrf[I3(i, j, k)])
#define taurstr(i, j, k) (S->taurstr[I3(i, j, k)])
#define taurstrb(i, j, k) (S->taurstrb[I3(i, j, k)])
#define taurstrf(i, j, k) (S->taurstrf[I3(i, j, k)])
/* 2-D COMMON arrays */
#define aam2d(i, j) (S->aam2d[I2(i, j)])
#define advua(i, j) (S->advua[I2(i, j)])
#define advva(i, j) (S->advva[I2(i, j)])
#define adx2d(i, j) (S->adx2d[I2(i, j)])
#define ady2d(i, j) (S->ady2d[I2(i, j)])
#define art(i, j) (S->art[I2(i, j)])
#define aru(i, j) (S->aru[I2(i, j)])
#define arv(i, j) (S->arv[I2(i, j)])
#define cbc(i, j) (S->cbc[I2(i, j)])
#define cor(i, j) (S->cor[I2(i, j)])
#define d(i, j) (S->d[I2(i, j)])
#define drx2d(i, j) (S->drx2d[I2(i, j)])
#define dry2d(i, j) (S->dry2d[I2(i, j)])
#define dt(i, j) (S->dt[I2(i, j)])
#define dum(i, j) (S->dum[I2(i, j)])
#define dvm(i, j) (S->dvm[I2(i, j)])
#define dx(i, j) (S->dx[I2(i, j)])
#define dy(i, j) (S->dy[I2(i, j)])
#define e_atmos(i, j) (S->e_atmos[I2(i, j)])
#define egb(i, j) (S->egb[I2(i, j)])
#define egf(i, j) (S->egf[I2(i, j)])
#define el(i, j) (S->el[I2(i, j)])
#define elb(i, j) (S->elb[I2(i, j)])
#define elf(i, j) (S->elf[I2(i, j)])
#define et(i, j) (S->et[I2(i, j)])
#define etb(i, j) (S->etb[I2(i, j)])
#define etf(i, j) (S->etf[I2(i, j)])
#define fluxua(i, j) (S->fluxua[I2(i, j)])
#define fluxva(i, j) (S->fluxva[I2(i, j)])
#define fsm(i, j) (S->fsm[I2(i, j)])
#define h(i, j) (S->h[I2(i, j)])
#define swrad(i, j) (S->swrad[I2(i, j)])
#define tps(i, j) (S->tps[I2(i, j)])
#define ua(i, j) (S->ua[I2(i, j)])
#define uab(i, j) (S->uab[I2(i, j)])
#define uaf(i, j) (S->uaf[I2(i, j)])
#define utb(i, j) (S->utb[I2(i, j)])
#define utf(i, j) (S->utf[I2(i, j)])
#define va(i, j) (S->va[I2(i, j)])
#define vab(i, j) (S->vab[I2(i, j)])
#define vaf(i, j) (S->vaf[I2(i, j)])
#define vtb(i, j) (S->vtb[I2(i, j)])
#define vtf(i, j) (S->vtf[I2(i, j)])
#define vfluxb(i, j) (S->vfluxb[I2(i, j)])
#define vfluxf(i, j) (S->vfluxf[I2(i, j)])
#define wtsurf(i, j) (S->wtsurf[I2(i, j)])
#define wubot(i, j) (S->wubot[I2(i, j)])
#define wusurf(i, j) (S->wusurf[I2(i, j)])
#define wvbot(i, j) (S->wvbot[I2(i, j)])
#define wvsurf(i, j) (S->wvsurf[I2(i, j)])
/* boundary arrays */
#define ele(j) (S->ele[(j)-1])
#define elw(j) (S->elw[(j)-1])
#define eln(i) (S->eln[(i)-1])
#define els(i) (S->els[(i)-1])
#define uabe(j) (S->uabe[(j)-1])
#define uabw(j) (S->uabw[(j)-1])
#define vabe(j) (S->vabe[(j)-1])
#define vabw(j) (S->vabw[(j)-1])
#define vabn(i) (S->vabn[(i)-1])
#define vabs(i) (S->vabs[(i)-1])
#define uabn(i) (S->uabn[(i)-1])
#define uabs(i) (S->uabs[(i)-1])
#define tbe(j, k) (S->tbe[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define sbe(j, k) (S->sbe[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define tbw(j, k) (S->tbw[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define sbw(j, k) (S->sbw[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define tbn(i, k) (S->tbn[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define sbn(i, k) (S->sbn[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define tbs(i, k) (S->tbs[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define sbs(i, k) (S->sbs[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define tbeb(j, k) (S->tbeb[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define tbef(j, k) (S->tbef[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define sbeb(j, k) (S->sbeb[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define sbef(j, k) (S->sbef[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define ubeb(j, k) (S->ubeb[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define ubef(j, k) (S->ubef[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define tbwb(j, k) (S->tbwb[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define tbwf(j, k) (S->tbwf[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define sbwb(j, k) (S->sbwb[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define sbwf(j, k) (S->sbwf[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define ubwb(j, k) (S->ubwb[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define ubwf(j, k) (S->ubwf[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define ube(j, k) (S->ube[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define ubw(j, k) (S->ubw[(size_t)((j)-1) + (size_t)jm * ((k)-1)])
#define tbnb(i, k) (S->tbnb[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define tbnf(i, k) (S->tbnf[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define sbnb(i, k) (S->sbnb[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define sbnf(i, k) (S->sbnf[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define vbnb(i, k) (S->vbnb[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define vbnf(i, k) (S->vbnf[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define tbsb(i, k) (S->tbsb[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define tbsf(i, k) (S->tbsf[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define sbsb(i, k) (S->sbsb[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define sbsf(i, k) (S->sbsf[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define vbsb(i, k) (S->vbsb[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define vbsf(i, k) (S->vbsf[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define vbn(i, k) (S->vbn[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define vbs(i, k) (S->vbs[(size_t)((i)-1) + (size_t)im * ((k)-1)])
#define wusurfb(i, j) (S->wusurfb[I2(i, j)])
#define wusurff(i, j) (S->wusurff[I2(i, j)])
#define wvsurfb(i, j) (S->wvsurfb[I2(i, j)])
#define wvsurff(i, j) (S->wvsurff[I2(i, j)])
#define wtsurfb(i, j) (S->wtsurfb[I2(i, j)])
#define wtsurff(i, j) (S->wtsurff[I2(i, j)])
#define swradb(i, j) (S->swradb[I2(i, j)])
#define swradf(i, j) (S->swradf[I2(i, j)])
/* 1-D */
#define z(k) (S->z[(k)-1])
#define zz(k) (S->zz[(k)-1])
#define dz(k) (S->dz[(k)-1])
#define dzz(k) (S->dzz[(k)-1])
/* blkcon scalars */
#define alpha (S->alpha)
#define dte (S->dte)
#define dti (S->dti)
#define dti2 (S->dti2)
#define grav (S->grav)
#define kappa (S->kappa)
#define ramp (S->ramp)
#define rfe (S->rfe)
#define rfn (S->rfn)
#define rfs (S->rfs)
#define rfw (S->rfw)
#define rhoref (S->rhoref)
#define sbias (S->sbias)
#define small (S->small)
#define tbias (S->tbias)
#define tprni (S->tprni)
#define umol (S->umol)
#define dte2 (S->dte2)
#define horcon (S->horcon)
#define ispi (S->ispi)
#define isp2i (S->isp2i)
#define smoth (S->smoth)
#define sw (S->sw)
#define n_west (S->n_west)
#define n_east (S->n_east)
#define n_south (S->n_south)
#define n_north (S->n_north)
/* loop helper: Fortran `do v=a,b` */
#define DO(v, a, b) for (int v = (a); v <= (b); ++v)
#define OMP_FOR _Pragma("omp parallel for schedule(static)")
#endif /* POMO_IMPL */
#endif
