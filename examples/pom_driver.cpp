// pom_driver.cpp -- a COMPILED host for libpomgpu, using nothing but the C ABI of include/pomgpu.h.
//
// It has the shape of the reference's driver: `program pom` (pom/pom.f:10-25: initialize, then
// `do iint=1,iend; call advance; end do`) with `advance` reduced to what stays on the host once the
// four step routines live in the library (pom/advance.f:6-59: get_time, the step, check_velocity and
// the error_status test).  The reference reads its initial state from netCDF files (initialize.f,
// io_pnetcdf.F); this driver reads the same arrays from a flat state file written by
// scripts/dump_state.py, pushes them by their COMMON-block names, makes the solver.f calls that
// `initialize` makes (dens x2, baropg / baropg_mcc: initialize.f:416,425,502-505), steps, and writes
// the prognostic fields back out -- the restart write of advance.f:43-49.
//
// The Fortran maintainer's version of the same calls is INTEGRATION.md; tests/test_driver.py runs this
// program against the Python-driven library and requires bitwise identical output.
//
//   pom_driver STATE.bin NSTEPS OUT.bin [NSTRIPS [same]]
// NSTRIPS > 1 spreads the domain over that many GPUs of the box, one j-strip each (`same`: all strips on
// device 0), in THIS process: what `distribute_mpi` + `exchange2d/3d_mpi` do for the reference
// (parallel_mpi.f:34-122,154-351) is pomgpu_create_strip + pomgpu_group_create here, the state file's global
// arrays are scattered / gathered with pomgpu_push_global / pomgpu_pull_global, and the result is bitwise the
// one-GPU result.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "pomgpu.h"

struct Field { std::string name; std::vector<double> data; };

static bool read_exact(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: pom_driver STATE.bin NSTEPS OUT.bin [NSTRIPS [same]]\n"); return 2; }
  const int nstrips = argc > 4 ? atoi(argv[4]) : 1;
  const bool same_device = argc > 5;
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  const int iend = atoi(argv[2]);
  char magic[8];
  int dims[3], nconst = 0, nfield = 0;
  if (!read_exact(f, magic, 8) || memcmp(magic, "POMSTAT1", 8) || !read_exact(f, dims, sizeof dims) ||
      !read_exact(f, &nconst, 4)) { fprintf(stderr, "bad state file\n"); return 2; }
  // one context (the whole domain) or NSTRIPS strips of (almost) equal height, south to north, 8 ghost rows per seam
  std::vector<pomgpu_t*> strips;
  pomgpu_group_t* grp = nullptr;
  if (nstrips > 1) {
    int j = 1;
    for (int r = 0; r < nstrips; ++r) {
      const int n = dims[1] / nstrips + (r < dims[1] % nstrips ? 1 : 0);
      pomgpu_t* c = pomgpu_create_strip(dims[0], dims[1], dims[2], j, j + n - 1, argc > 6 ? atoi(argv[6]) : 8, same_device ? 0 : r);
      if (!c) { fprintf(stderr, "pomgpu_create_strip %d failed\n", r); return 3; }
      strips.push_back(c);
      j += n;
    }
    grp = pomgpu_group_create(nstrips, strips.data());
    if (!grp) { fprintf(stderr, "pomgpu_group_create failed (each strip needs ghost+4 rows)\n"); return 3; }
  }
  pomgpu_t* ctx = grp ? strips[0] : pomgpu_create(dims[0], dims[1], dims[2], 0);
  if (!ctx) { fprintf(stderr, "pomgpu_create failed (no CUDA device? libpomgpu has no CPU fallback)\n"); return 3; }
  // blkcon scalars (pom.h_dist:69-198), by name
  for (int n = 0; n < nconst; ++n) {
    char name[32]; double v;
    if (!read_exact(f, name, 32) || !read_exact(f, &v, 8)) return 2;
    pomgpu_set_const(ctx, name, v);           // names outside blkcon are ignored
    for (size_t r = 1; r < strips.size(); ++r) pomgpu_set_const(strips[r], name, v);
  }
  // COMMON arrays, by name, column-major exactly as the Fortran holds them
  if (!read_exact(f, &nfield, 4)) return 2;
  for (int n = 0; n < nfield; ++n) {
    char name[32]; long cnt;
    if (!read_exact(f, name, 32) || !read_exact(f, &cnt, 8)) return 2;
    std::vector<double> a((size_t)cnt);
    if (!read_exact(f, a.data(), (size_t)cnt * 8)) return 2;
    if (pomgpu_field_global_elems(ctx, name) != cnt) continue;     // not a field of the hot path
    if (!grp) {
      if (pomgpu_push(ctx, name, a.data())) { fprintf(stderr, "push(%s): %s\n", name, pomgpu_last_error(ctx)); return 3; }
    } else {
      for (pomgpu_t* c : strips)               // every strip takes the rows it holds (owned + ghost)
        if (pomgpu_push_global(c, name, a.data())) { fprintf(stderr, "push(%s): %s\n", name, pomgpu_last_error(c)); return 3; }
    }
  }
  fclose(f);
  // initialize.f:416,425: rmean = dens(sclim,tclim), rho = dens(sb,tb); :502-505: the first baropg
  double npg = 1., dti = 0., time0 = 0., vmaxl = 100.;
  pomgpu_get_const(ctx, "npg", &npg);
  pomgpu_get_const(ctx, "dti", &dti);
  pomgpu_get_const(ctx, "time0", &time0);
  pomgpu_get_const(ctx, "vmaxl", &vmaxl);
  if (!grp) {
    pomgpu_dens(ctx, "sclim", "tclim", "rmean");
    pomgpu_dens(ctx, "sb", "tb", "rho");
    if ((int)npg == 2) pomgpu_baropg_mcc(ctx); else pomgpu_baropg(ctx);
  } else {
    pomgpu_group_dens(grp, "sclim", "tclim", "rmean");
    pomgpu_group_dens(grp, "sb", "tb", "rho");
    pomgpu_group_baropg(grp);
  }
  // pom.f:16-20
  for (int iint = 1; iint <= iend; ++iint) {
    const double time = dti * (double)iint / 86400. + time0;     // get_time, advance.f:66
    const double ramp = 1.;                                       // lramp = .false. (advance.f:68-73)
    if (grp ? pomgpu_group_step(grp, iint, time, ramp) : pomgpu_step(ctx, iint, time, ramp)) { fprintf(stderr, "step %d: %s\n", iint, pomgpu_last_error(ctx)); return 3; }
    const double vamax = grp ? pomgpu_group_check_velocity(grp) : pomgpu_check_velocity(ctx);   // advance.f:52,611-641
    double err = 0.;
    if (grp) err = (double)pomgpu_group_error_status(grp); else pomgpu_get_const(ctx, "error_status", &err);
    if (vamax > vmaxl || err != 0.) {                             // advance.f:623-638, 556-563
      fprintf(stderr, "stopped at iint=%d: vamax=%g error_status=%g\n", iint, vamax, err);
      return 4;
    }
    if (iint == iend) printf("iint %d time %.6f vamax %.17g\n", iint, time, vamax);
  }
  // advance.f:43-49 (restart write): pull what the next run needs
  static const char* const out[] = {"el", "elb", "ua", "uab", "va", "vab", "u", "ub", "v", "vb", "w", "t", "tb", "s", "sb",
                                    "rho", "km", "kh", "kq", "l", "q2", "q2b", "q2l", "q2lb", "aam", "wubot", "wvbot"};
  FILE* o = fopen(argv[3], "wb");
  if (!o) { perror(argv[3]); return 2; }
  for (const char* n : out) {
    const long cnt = pomgpu_field_global_elems(ctx, n);
    std::vector<double> a((size_t)cnt);
    if (!grp) { if (pomgpu_pull(ctx, n, a.data())) return 3; }
    else for (pomgpu_t* c : strips) if (pomgpu_pull_global(c, n, a.data())) return 3;   // every strip returns the rows it owns
    char name[32] = {0};
    strncpy(name, n, 31);
    fwrite(name, 1, 32, o); fwrite(&cnt, 8, 1, o); fwrite(a.data(), 8, (size_t)cnt, o);
  }
  fclose(o);
  if (grp) { pomgpu_group_destroy(grp); for (pomgpu_t* c : strips) pomgpu_destroy(c); }
  else pomgpu_destroy(ctx);
  return 0;
}
