/* fortran_abi_driver.c -- plays the reference's Fortran driver against libpomgpu_f.so, in C.
 *
 * It DEFINES the COMMON blocks of pom.h (blksiz_, blkcon_, blk1d_, blk2d_, blk3d_, bdry_: what
 * gfortran would emit for `include 'pom.h'`), fills them from a state file written by
 * scripts/dump_state.py (the stand-in for the netCDF inputs), makes the solver.f calls `initialize`
 * makes (dens x2, baropg: initialize.f:416,425,502-517) and then runs `advance`'s hot path exactly
 * as advance.f:21-32 spells it -- argument-less calls, iint / iext / time / ramp passed through
 * blkcon -- followed by check_velocity's host-side read of vaf (advance.f:619-629).
 *
 *   fortran_abi_driver STATE.bin NSTEPS OUT.bin [unit]
 * With `unit` it then calls a few routine-level entries (unit mode, host arrays incl. local ones).
 * Build: gcc -DIML=.. -DJML=.. -DKB=.. fortran_abi_driver.c -lpomgpu_f -lpomgpu
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pomgpu_f.h"

#ifndef IML
#error "compile with -DIML=<im_local> -DJML=<jm_local> -DKB=<kb> (the parameter statement of pom.h)"
#endif
#define N2 ((size_t)IML * JML)
#define N3 ((size_t)IML * JML * KB)
/* the COMMON blocks, sized as pom.h_dist declares them */
int blksiz_[8];
int blkpar_[7 + IML + JML];
double blkcon_[22 + 2 + 16 + 7];                    /* 22 dp, 4 int, 16 dp, 14 int = 58 eight-byte words */
double blk1d_[4 * KB];
double blk2d_[73 * N2];
double blk3d_[40 * N3];
double bdry_[2 * JML + 2 * IML + 12 * JML * KB + 12 * IML * KB + 12 * JML + 12 * JML * KB + 12 * IML + 12 * IML * KB];

static double* D(const char* n) { return (double*)pomgpu_f_member(n, NULL); }
static void setc(const char* n, double v) {
  void* p = pomgpu_f_member(n, NULL);
  if (!p) return;
  if (pomgpu_f_member_type(n) == 'd') *(double*)p = v; else *(int*)p = (int)v;
}
static double getc_(const char* n) {
  void* p = pomgpu_f_member(n, NULL);
  if (!p) return 0.;
  return pomgpu_f_member_type(n) == 'd' ? *(double*)p : (double)*(int*)p;
}
static int rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }
static void put(FILE* f, const char* name, const double* a, long cnt) {
  char nm[32]; memset(nm, 0, 32); strncpy(nm, name, 31);
  fwrite(nm, 1, 32, f); fwrite(&cnt, 8, 1, f); fwrite(a, 8, (size_t)cnt, f);
}

#ifdef WITH_RESTORE
/* `subroutine restore_interior_records`: the record half of restore_interior (bounds_forcing.f:1023-1081) that
 * scripts/make_glue.py leaves in the Fortran glue.  libpomgpu_f's mode_internal_ calls it back at the place of
 * advance.f:452.  The netCDF reader (read_restore_ts_interior_pnetcdf) is played by the climatology. */
static void read_records(void) {
  memcpy(D("trstrf"), D("tclim"), N3 * 8);                                   /* :1040-1041, 1066-1067 */
  memcpy(D("srstrf"), D("sclim"), N3 * 8);
  double* tau = D("taurstrf");
  for (size_t q = 0; q < N3; ++q) tau[q] = 1.f / 30.;                        /* taurstrf = 1./trst, :1042 */
}
void restore_interior_records_(void) {
  const double trst = 30.;
  const int iint = (int)getc_("iint"), irst = (int)(trst * 86400. / getc_("dti")), iend = (int)getc_("iend");
  if (iint == 2) read_records();                                             /* :1038 */
  if (iint == 2 || iint % irst == 0) {                                       /* :1053 */
    memcpy(D("trstrb"), D("trstrf"), N2 * (KB - 1) * 8);                     /* k = 1..kbm1, :1054-1062 */
    memcpy(D("srstrb"), D("srstrf"), N2 * (KB - 1) * 8);
    memcpy(D("taurstrb"), D("taurstrf"), N2 * (KB - 1) * 8);
    if (iint != iend) read_records();                                        /* :1063-1067 */
  }
}
#endif

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: fortran_abi_driver STATE.bin NSTEPS OUT.bin [unit]\n"); return 2; }
  const int unit = argc > 4 && !strcmp(argv[4], "unit");
  int iml = IML, jml = JML, kb = KB;
  pomgpu_f_set_dims_(&iml, &jml, &kb);
  FILE* f = fopen(argv[1], "rb");
  char magic[8]; int dims[3], nconst = 0, nfield = 0;
  if (!f || !rd(f, magic, 8) || memcmp(magic, "POMSTAT1", 8) || !rd(f, dims, 12) || !rd(f, &nconst, 4)) { fprintf(stderr, "bad state file\n"); return 2; }
  if (dims[0] != IML || dims[1] != JML || dims[2] != KB) { fprintf(stderr, "state file is %dx%dx%d, compiled for %dx%dx%d\n", dims[0], dims[1], dims[2], IML, JML, KB); return 2; }
  if (!pomgpu_f_member("u", NULL)) { fprintf(stderr, "COMMON blocks not bound: %s\n", pomgpu_f_last_error()); return 4; }
  /* distribute_mpi on one rank (parallel_mpi.f:76-119) */
  setc("im", IML); setc("imm1", IML - 1); setc("imm2", IML - 2); setc("jm", JML); setc("jmm1", JML - 1);
  setc("jmm2", JML - 2); setc("kbm1", KB - 1); setc("kbm2", KB - 2);
  setc("n_west", -1); setc("n_east", -1); setc("n_south", -1); setc("n_north", -1);
  for (int n = 0; n < nconst; ++n) {   /* read_input (initialize.f:67-191) */
    char name[32]; double v;
    if (!rd(f, name, 32) || !rd(f, &v, 8)) return 2;
    setc(name, v);
  }
  setc("iprint", 1000000); setc("irestart", 1000000);
  if (!rd(f, &nfield, 4)) return 2;
  for (int n = 0; n < nfield; ++n) {   /* read_grid / initial_conditions */
    char name[32]; long cnt, have = 0;
    if (!rd(f, name, 32) || !rd(f, &cnt, 8)) return 2;
    double* dst = (double*)pomgpu_f_member(name, &have);
    if (dst && have == cnt && pomgpu_f_member_type(name) == 'd') { if (!rd(f, dst, (size_t)cnt * 8)) return 2; }
    else fseek(f, cnt * 8, SEEK_CUR);
  }
  fclose(f);
  /* initialize.f:416,425: rmean=dens(sclim,tclim), rho=dens(sb,tb); :502-505 the first baropg; :510-517 */
  dens_(D("sclim"), D("tclim"), D("rmean"));
  dens_(D("sb"), D("tb"), D("rho"));
  if ((int)getc_("npg") == 2) baropg_mcc_(); else baropg_();
  {
    double *drx = D("drx2d"), *dry = D("dry2d"), *dx3 = D("drhox"), *dy3 = D("drhoy"), *dz = D("dz");
    for (size_t q = 0; q < N2; ++q) { drx[q] = 0.; dry[q] = 0.; }
    for (int k = 0; k < KB - 1; ++k)
      for (size_t q = 0; q < N2; ++q) { drx[q] = drx[q] + dx3[k * N2 + q] * dz[k]; dry[q] = dry[q] + dy3[k * N2 + q] * dz[k]; }
  }
  const int iend = atoi(argv[2]), isplit = (int)getc_("isplit");
  const double dti = getc_("dti"), time0 = getc_("time0"), vmaxl = getc_("vmaxl");
  double vamax = 0.;
  for (int iint = 1; iint <= iend; ++iint) {   /* pom.f:17-19 */
    setc("iint", iint);
    setc("time", dti * (double)iint / 86400. + time0);   /* get_time, advance.f:66 */
    setc("ramp", 1.);
    lateral_viscosity_();                                /* advance.f:21 */
    mode_interaction_();                                 /* advance.f:24 */
    int iext;
    for (iext = 1; iext <= isplit; ++iext) { setc("iext", iext); mode_external_(); }   /* advance.f:27-29 */
    setc("iext", iext);
    mode_internal_();                                    /* advance.f:32 */
    /* check_velocity (advance.f:611-641) on the host copy of vaf, like the Fortran glue */
    const double* vaf = D("vaf");
    vamax = 0.;
    for (size_t q = 0; q < N2; ++q) { double v = fabs(vaf[q]); if (!(v <= vamax)) vamax = v; }
    if (!(vamax <= vmaxl)) setc("error_status", 1);
    if ((int)getc_("error_status")) { fprintf(stderr, "error_status=1 at iint=%d: %s\n", iint, pomgpu_f_last_error()); return 3; }
  }
  printf("vamax %.17g\n", vamax);
  FILE* o = fopen(argv[3], "wb");
  if (!o) return 2;
  if (unit) {
    /* routine level, with the device ahead of the host: the library pulls the state first */
    advct_();                                            /* COMMON in, COMMON out */
    put(o, "advx", D("advx"), (long)N3); put(o, "advy", D("advy"), (long)N3);
    double *sl = malloc(N3 * 8), *tl = malloc(N3 * 8), *rl = calloc(N3, 8);
    memcpy(sl, D("s"), N3 * 8); memcpy(tl, D("t"), N3 * 8);
    dens_(sl, tl, rl);                                   /* three arrays that are NOT COMMON members */
    put(o, "rho_local", rl, (long)N3);
    advq_(D("q2b"), D("q2"), rl);                        /* COMMON in, local out */
    put(o, "qf_local", rl, (long)N3);
    int nbc = 1;
    memcpy(rl, D("t"), N3 * 8);
    proft_(rl, D("wtsurf"), D("tsurf"), &nbc);           /* local in/out, COMMON 2-D arguments */
    put(o, "proft_local", rl, (long)N3);
    int six = 6, three = 3;
    bcond_(&six);
    put(o, "uf_bcond6", D("uf"), (long)N3);
    bcondorl_(&three);
    put(o, "vf_bcondorl3", D("vf"), (long)N3);
    free(sl); free(tl); free(rl);
  } else {
    pomgpu_f_pull_all_();                                /* what a restart step would do (advance.f:49) */
    static const char* names[] = {"u", "v", "ub", "vb", "t", "s", "tb", "sb", "q2", "q2b", "q2l", "q2lb", "w", "rho", "km", "kh",
                                  "kq", "aam", "l", "wr", "el", "elb", "et", "etb", "ua", "uab", "va", "vab", "uaf", "vaf", "elf", "uf", "vf",
                                  "egb", "utb", "vtb", "wubot", "wvbot", "adx2d", "ady2d", "advua", "advva", "aam2d", "d", "dt", NULL};
    for (int n = 0; names[n]; ++n) {
      long cnt = 0;
      double* a = (double*)pomgpu_f_member(names[n], &cnt);
      if (a) put(o, names[n], a, cnt);
    }
  }
  fclose(o);
  pomgpu_f_finalize_();
  return 0;
}
