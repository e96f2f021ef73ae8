// pomgpu_f.cc -- libpomgpu_f.so: the reference's Fortran entry points (gfortran ABI) and COMMON-block
// mirrors over the C ABI of libpomgpu.so.  See include/pomgpu_f.h for the contract; this file is
// plain host C++ (no CUDA, no Fortran): symbol names, by-reference arguments and the COMMON layout
// generated from the model's pom.h (include/pom_common_layout.h) are all the "Fortran" there is.
#include "pomgpu_f.h"
#include "pomgpu.h"
#include "pom_common_layout.h"
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

// The driver's COMMON blocks (pom.h_dist:46-54,69-78,142-198,208-212,291-364,410-450,532-608).  Weak:
// the library must load (and list its symbols) in a process that has no such blocks; an entry point
// called there reports the missing binding through error_status-style messages instead of crashing.
extern "C" {
extern char blksiz_[] __attribute__((weak));
extern char blkpar_[] __attribute__((weak));
extern char blkcon_[] __attribute__((weak));
extern char blk1d_[] __attribute__((weak));
extern char blk2d_[] __attribute__((weak));
extern char blk3d_[] __attribute__((weak));
extern char bdry_[] __attribute__((weak));
// The record half of restore_interior (bounds_forcing.f:1023-1081: the netCDF reads and the `b = f` copies that
// precede "linear interpolation in time"), kept in Fortran: scripts/make_glue.py writes it into the glue file as
// `subroutine restore_interior_records`.  The reference calls restore_interior from INSIDE mode_internal
// (advance.f:452), which is now this library's, so mode_internal_ calls the record half back at the same place.
void restore_interior_records_(void) __attribute__((weak));
}

namespace {

struct Member { std::string name; char type; int block; size_t off, elems; };

int g_iml = POMF_IM_LOCAL, g_jml = POMF_JM_LOCAL, g_kb = POMF_KB, g_device = 0;
bool g_resolved = false;
std::vector<Member> g_mem;
char* g_base[POMF_NBLOCKS];
pomgpu_t* g_ctx = nullptr;      // the whole domain on one device: the resident state (one device) and the unit-mode target
// more than one device (pomgpu_f_set_devices_ / POMGPU_F_DEVICES): the single-rank driver's domain is cut into
// j-strips, one per GPU of the box, inside this process; the COMMON arrays keep their global extents on the host
int g_ndev = 0;                 // 0 = not set: POMGPU_F_DEVICES or 1;  < 0: |n| strips, all on g_device (tests on one GPU)
int g_ghost = 8;
std::vector<pomgpu_t*> g_strips;
pomgpu_group_t* g_grp = nullptr;
bool g_t_main = true;           // target of the t_* helpers: the resident state (true) or the unit-mode context g_ctx
bool multi() { return !g_strips.empty(); }
pomgpu_t* any_ctx() { return multi() ? g_strips[0] : g_ctx; }
std::vector<int> g_dev;        // COMMON members that are device fields (index into g_mem)
bool g_full_pushed = false;    // the whole state has been pushed once (start of resident mode)
bool g_device_ahead = false;   // step-level calls ran since the last full pull: host copies are stale
int g_restore = -1;            // restore_interior's nudging (bounds_forcing.f:1083-1118): -1 = on iff the driver supplies
                               // restore_interior_records or pushes a restoring record, 0 / 1 = pomgpu_f_set_restore_
bool g_restore_pushed = false; // the driver pushed a restoring record by hand (pomgpu_f_push_)
char g_err[256] = "";

void fail(const char* what) {
  snprintf(g_err, sizeof(g_err), "%s", what);
  fprintf(stderr, "\npomgpu_f: %s\n", what);
  // the reference's convention (advance.f:118-119,556-563): error_status=1, the driver stops at the
  // next print step
  for (auto& m : g_mem)
    if (m.name == "error_status" && g_base[m.block]) *(int*)(g_base[m.block] + m.off) = 1;
}

char* block_symbol(const char* blk) {
  if (!strcmp(blk, "blksiz") && blksiz_) return blksiz_;
  if (!strcmp(blk, "blkpar") && blkpar_) return blkpar_;
  if (!strcmp(blk, "blkcon") && blkcon_) return blkcon_;
  if (!strcmp(blk, "blk1d") && blk1d_) return blk1d_;
  if (!strcmp(blk, "blk2d") && blk2d_) return blk2d_;
  if (!strcmp(blk, "blk3d") && blk3d_) return blk3d_;
  if (!strcmp(blk, "bdry") && bdry_) return bdry_;
  std::string s = std::string(blk) + "_";     // an executable linked with -rdynamic, or a block we have no weak reference for
  return (char*)dlsym(RTLD_DEFAULT, s.c_str());
}

// byte offsets of every member: declaration order, each aligned to its element size (gfortran pads
// a COMMON block only for alignment; pom.h needs none)
void resolve() {
  if (g_resolved) return;
  g_mem.clear();
  for (int b = 0; b < POMF_NBLOCKS; ++b) {
    const pomf_block_t& B = POMF_BLOCKS[b];
    g_base[b] = block_symbol(B.block);
    size_t off = 0;
    for (int n = 0; n < B.n; ++n) {
      const pomf_member_t& m = B.m[n];
      size_t es = (m.type == 'd') ? 8 : (m.type == 'c') ? (size_t)m.len : 4, cnt = 1;
      for (const char* d = m.dims; *d; ++d) cnt *= (*d == 'I') ? g_iml : (*d == 'J') ? g_jml : g_kb;
      const size_t al = (m.type == 'c') ? 1 : es;
      off = (off + al - 1) / al * al;
      g_mem.push_back({m.name, m.type, b, off, cnt});
      off += es * cnt;
    }
  }
  g_resolved = true;
}

const Member* find(const char* name) {
  resolve();
  for (auto& m : g_mem)
    if (m.name == name) return &m;
  return nullptr;
}
char* addr(const Member* m) { return (m && g_base[m->block]) ? g_base[m->block] + m->off : nullptr; }
int iget(const char* name, int dflt) { const Member* m = find(name); char* a = addr(m); return (a && m->type != 'd') ? *(int*)a : dflt; }

// the COMMON member a host pointer is the start of (array arguments of advq, advt1/2, dens, proft)
const Member* member_at(const void* p) {
  resolve();
  for (auto& m : g_mem)
    if (m.type == 'd' && m.elems > 1 && addr(&m) == (const char*)p) return &m;
  return nullptr;
}

long dev_elems(const char* name) { pomgpu_t* c = any_ctx(); return c ? pomgpu_field_global_elems(c, name) : 0; }
int t_push(const char* name, const double* host) {
  if (!(g_t_main && multi())) return pomgpu_push(g_ctx, name, host);
  for (pomgpu_t* c : g_strips)
    if (int rc = pomgpu_push_global(c, name, host)) return rc;
  return 0;
}
int t_pull(const char* name, double* host) {
  if (!(g_t_main && multi())) return pomgpu_pull(g_ctx, name, host);
  for (pomgpu_t* c : g_strips)
    if (int rc = pomgpu_pull_global(c, name, host)) return rc;
  return 0;
}
void t_set_const(const char* name, double v) {
  for (pomgpu_t* c : g_strips) pomgpu_set_const(c, name, v);
  if (g_ctx) pomgpu_set_const(g_ctx, name, v);
}

bool ensure_ctx() {
  if (g_ctx || multi()) return true;
  resolve();
  if (!g_base[0] || !find("im")) { fail("the COMMON blocks of pom.h (blksiz_, blkcon_, blk1d_, blk2d_, blk3d_, bdry_) are not visible to libpomgpu_f: link the driver against it (or with -rdynamic)"); return false; }
  const int im = iget("im", 0), jm = iget("jm", 0);
  if (im != g_iml || jm != g_jml) {
    // exchange*_mpi is called with im_local / jm_local in the reference, i.e. uniform blocks are
    // assumed there too (advance.f:137 vs parallel_mpi.f:173); the strides of the COMMON arrays are
    // im_local, jm_local, so the active extents must fill them
    char m[200];
    snprintf(m, sizeof(m), "blksiz says im=%d jm=%d but the COMMON arrays are (%d,%d,%d): call pomgpu_f_set_dims_ or regenerate pom_common_layout.h from this build's pom.h", im, jm, g_iml, g_jml, g_kb);
    fail(m);
    return false;
  }
  if (g_ndev == 0) { const char* e = getenv("POMGPU_F_DEVICES"); g_ndev = e ? atoi(e) : 1; if (g_ndev == 0) g_ndev = 1; }
  g_ghost = 8;
  if (const char* e = getenv("POMGPU_F_GHOST")) g_ghost = atoi(e);
  const int ns = g_ndev < 0 ? -g_ndev : g_ndev;
  if (ns == 1) {
    g_ctx = pomgpu_create(im, jm, g_kb, g_device);
    if (!g_ctx) { fail("pomgpu_create failed (no CUDA device, bad extents or out of memory); libpomgpu has no CPU fallback"); return false; }
  } else {
    // j-strips of (almost) equal height, south to north; strip r on device g_device + r
    int j = 1;
    for (int r = 0; r < ns; ++r) {
      const int n = jm / ns + (r < jm % ns ? 1 : 0);
      pomgpu_t* c = (n >= 1) ? pomgpu_create_strip(im, jm, g_kb, j, j + n - 1, g_ghost, g_ndev < 0 ? g_device : g_device + r) : nullptr;
      if (!c) {
        for (pomgpu_t* q : g_strips) pomgpu_destroy(q);
        g_strips.clear();
        fail("pomgpu_create_strip failed (fewer CUDA devices than pomgpu_f_set_devices_ asked for, too few rows per strip, or out of memory)");
        return false;
      }
      g_strips.push_back(c);
      j += n;
    }
    g_grp = pomgpu_group_create(ns, g_strips.data());
    if (!g_grp) {
      for (pomgpu_t* q : g_strips) pomgpu_destroy(q);
      g_strips.clear();
      fail("pomgpu_group_create failed: every strip needs at least ghost+4 rows (POMGPU_F_GHOST, default 8) and there are at most 16 strips");
      return false;
    }
  }
  g_dev.clear();
  for (size_t n = 0; n < g_mem.size(); ++n) {
    const Member& m = g_mem[n];
    if (m.type != 'd' || m.elems < 2 || !g_base[m.block]) continue;
    if (dev_elems(m.name.c_str()) == (long)m.elems) g_dev.push_back((int)n);
  }
  return true;
}
// the unit-mode target of the routines that have no group entry point: the whole domain on one device
bool ensure_unit() {
  if (g_ctx) return true;
  g_ctx = pomgpu_create(iget("im", 0), iget("jm", 0), g_kb, g_device);
  if (!g_ctx) fail("unit mode with several devices runs the routine on ONE device and the whole domain does not fit there (dens, baropg, baropg_mcc run on the strips)");
  return g_ctx != nullptr;
}

bool ck(int rc, const char* what) {
  if (rc == 0) return true;
  char m[300];
  snprintf(m, sizeof(m), "%s failed (rc=%d): %s", what, rc, any_ctx() ? pomgpu_last_error(g_t_main ? any_ctx() : g_ctx) : "");
  fail(m);
  return false;
}

void push_consts() {   // blkcon scalars by name (names the device does not know are ignored)
  for (auto& m : g_mem) {
    if (strcmp(POMF_BLOCKS[m.block].block, "blkcon") || !g_base[m.block]) continue;
    const double v = (m.type == 'd') ? *(double*)addr(&m) : (double)*(int*)addr(&m);
    if (m.name == "error_status") continue;           // the device's own flag is not overwritten
    t_set_const(m.name.c_str(), v);
  }
}
bool push_names(const char* list) {
  std::string s(list);
  size_t a = 0;
  while (a < s.size()) {
    size_t b = s.find(' ', a);
    if (b == std::string::npos) b = s.size();
    if (b > a) {
      const std::string n = s.substr(a, b - a);
      const Member* m = find(n.c_str());
      if (m && addr(m) && dev_elems(n.c_str()) == (long)m->elems)
        if (!ck(t_push(n.c_str(), (double*)addr(m)), ("push " + n).c_str())) return false;
    }
    a = b + 1;
  }
  return true;
}
bool pull_names(const char* list) {
  std::string s(list);
  size_t a = 0;
  while (a < s.size()) {
    size_t b = s.find(' ', a);
    if (b == std::string::npos) b = s.size();
    if (b > a) {
      const std::string n = s.substr(a, b - a);
      const Member* m = find(n.c_str());
      if (m && addr(m) && dev_elems(n.c_str()) == (long)m->elems)
        if (!ck(t_pull(n.c_str(), (double*)addr(m)), ("pull " + n).c_str())) return false;
    }
    a = b + 1;
  }
  return true;
}

// what the Fortran driver refreshes on the host before every hot-path call: surface forcing
// (bounds_forcing.f:908-909 wind, :954-955 heat, :978 surface, water), the open-boundary arrays
// lateral_bc interpolates (:844-865) and reads (:613-616)
const char* FORCING = "wusurf wvsurf wtsurf wssurf swrad tsurf ssurf e_atmos vfluxf "
                      "ele eln els elw tbe sbe tbw sbw tbn sbn tbs sbs ube ubw vbn vbs "
                      "uabe uabw vabe vabw vabn vabs uabn uabs";
const char* BDRY = "ele eln els elw tbe sbe tbw sbw tbn sbn tbs sbs ube ubw vbn vbs uabe uabw vabe vabw vabn vabs uabn uabs";
// geometry, masks and the vertical grid: constant after `initialize`, but unit-mode calls happen
// INSIDE initialize (dens at initialize.f:416, before bottom_friction fills cbc), so they are pushed
// with every unit-mode call
const char* STATIC = "z zz dz dzz dx dy art aru arv cor h fsm dum dvm cbc";

bool push_all() {
  if (!ensure_ctx()) return false;
  g_t_main = true;
  push_consts();
  for (int n : g_dev)
    if (!ck(t_push(g_mem[n].name.c_str(), (double*)addr(&g_mem[n])), g_mem[n].name.c_str())) return false;
  g_full_pushed = true;
  g_device_ahead = false;
  return true;
}

// after the time rotations the device keeps the newest level under the `n` name only
// (ua<->uaf, va<->vaf, el<->elf, u<->uf, v<->vf are pointer swaps there); in the reference the copies
// leave both names equal (advance.f:324-330: ua=uaf ..., :511-514: u=uf, v=vf)
void mirror_rotated() {
  static const char* pairs[][2] = {{"uaf", "ua"}, {"vaf", "va"}, {"elf", "el"}, {"uf", "u"}, {"vf", "v"}};
  for (auto& pr : pairs) {
    const Member *a = find(pr[0]), *b = find(pr[1]);
    if (a && b && addr(a) && addr(b)) memcpy(addr(a), addr(b), a->elems * 8);
  }
}

bool pull_all() {
  if (!any_ctx()) return true;
  g_t_main = true;
  for (int n : g_dev)
    if (!ck(t_pull(g_mem[n].name.c_str(), (double*)addr(&g_mem[n])), g_mem[n].name.c_str())) return false;
  if (g_device_ahead) mirror_rotated();
  g_device_ahead = false;
  return true;
}

void pull_error_status() {
  double es = 0.;
  if (g_t_main && multi()) es = (double)pomgpu_group_error_status(g_grp);
  else if (g_ctx) pomgpu_get_const(g_ctx, "error_status", &es);
  if (es != 0.) {
    const Member* m = find("error_status");
    if (addr(m)) *(int*)addr(m) = 1;
  }
}

// ---- restore_interior (bounds_forcing.f:1023-1118), called from inside mode_internal (advance.f:452) -------
typedef void (*records_fn)(void);
records_fn records_hook() {
  if (restore_interior_records_) return restore_interior_records_;
  return (records_fn)dlsym(RTLD_DEFAULT, "restore_interior_records_");
}
const char* RESTORE = "trstrb trstrf srstrb srstrf taurstrb taurstrf";
bool is_restore_member(const std::string& n) { return (" " + std::string(RESTORE) + " ").find(" " + n + " ") != std::string::npos; }
// Runs where the reference runs the first half of restore_interior: only when mode_internal's tracer block does
// (advance.f:362,424: not on the skipped first step of a cold start, not for mode 2 or 4).  The Fortran routine
// re-reads its records when `iint.eq.2 .or. mod(iint,irst).eq.0` (:1038,1053); the six arrays travel to HBM on
// exactly those steps.  The time interpolation and the nudging (:1083-1118) are part of the device step.
bool restore_begin(int iint) {
  records_fn hook = records_hook();
  const int on = (g_restore >= 0) ? g_restore : ((hook || g_restore_pushed) ? 1 : 0);
  t_set_const("lrestore", (double)on);
  if (!on || !hook) return true;
  const Member *t0 = find("time0"), *dti = find("dti");
  const double time0 = addr(t0) ? *(double*)addr(t0) : 0., dt_i = addr(dti) ? *(double*)addr(dti) : 0.;
  const int mode = iget("mode", 3);
  if (!((iint != 1 || time0 != 0.) && mode != 2 && mode != 4)) return true;
  hook();
  const double trst = 30.;                                   // bounds_forcing.f:1033
  const int irst = (int)(trst * 86400. / dt_i);              // :1034
  if (iint == 2 || (irst > 0 && iint % irst == 0)) return push_names(RESTORE);
  return true;
}

// ---- step level: resident ------------------------------------------------------------------
bool step_begin() {
  if (!ensure_ctx()) return false;
  g_t_main = true;
  if (!g_full_pushed && !push_all()) return false;
  return true;
}

// ---- routine level: unit mode ----------------------------------------------------------------
struct Arg { const double* host; const Member* mem; std::string dev; };

// a host array argument -> the device field it is bound to for this call
bool bind_args(Arg* a, int n, const char* const* scratch) {
  for (int q = 0; q < n; ++q) {
    a[q].mem = member_at(a[q].host);
    if (a[q].mem && dev_elems(a[q].mem->name.c_str()) == (long)a[q].mem->elems) a[q].dev = a[q].mem->name;
    else { a[q].mem = nullptr; a[q].dev = scratch[q]; }    // a caller's local array: staged through a scratch field
  }
  return true;
}
// on_strips: the routine has a group entry point (dens, baropg, baropg_mcc: what `initialize` calls) and runs on the
// strips when there are several devices; every other routine runs on ONE device holding the whole domain
bool unit_begin(const char* inputs, bool on_strips = false) {
  if (!ensure_ctx()) return false;
  if (g_device_ahead && !pull_all()) return false;
  g_t_main = !multi() || on_strips;
  if (!g_t_main && !ensure_unit()) return false;
  push_consts();
  return push_names(STATIC) && push_names(BDRY) && push_names(inputs);
}
bool push_arg(const Arg& a) { return ck(t_push(a.dev.c_str(), a.host), ("push argument -> " + a.dev).c_str()); }
bool pull_arg(const Arg& a, double* host) { return ck(t_pull(a.dev.c_str(), host), ("pull argument <- " + a.dev).c_str()); }
void unit_end(const char* outputs) {
  pull_names(outputs);
  pull_error_status();
}
const char* S3[] = {"s3a", "s3b", "s3c", "s3d"};

}  // namespace

extern "C" {

// ================================ step level (advance.f) =====================================
void lateral_viscosity_(void) {   // advance.f:96-141
  if (!step_begin()) return;
  // the start of the hot path of one internal step (advance.f:21): time, ramp, iint and the forcing
  // the driver has just computed on the host (advance.f:12-18)
  push_consts();
  if (!push_names(FORCING)) return;
  g_device_ahead = true;
  ck(multi() ? pomgpu_group_lateral_viscosity(g_grp) : pomgpu_lateral_viscosity(g_ctx), "lateral_viscosity");
}
void mode_interaction_(void) {    // advance.f:144-202
  if (!step_begin()) return;
  g_device_ahead = true;
  ck(multi() ? pomgpu_group_mode_interaction(g_grp) : pomgpu_mode_interaction(g_ctx), "mode_interaction");
}
void mode_external_(void) {       // advance.f:205-353; iext is the driver's loop variable in blkcon (advance.f:27)
  if (!step_begin()) return;
  g_device_ahead = true;
  ck(multi() ? pomgpu_group_mode_external(g_grp, iget("iext", 1)) : pomgpu_mode_external(g_ctx, iget("iext", 1)), "mode_external");
}
void mode_internal_(void) {       // advance.f:356-537
  if (!step_begin()) return;
  g_device_ahead = true;
  const int iint = iget("iint", 1);
  t_set_const("iext", (double)iget("iext", 0));
  if (!restore_begin(iint)) return;
  if (!ck(multi() ? pomgpu_group_mode_internal(g_grp, iint) : pomgpu_mode_internal(g_ctx, iint), "mode_internal")) return;
  // what the Fortran glue reads on the host after EVERY step: vaf (check_velocity, advance.f:52,
  // 619-629).  On the device the rotation left the new vaf under the name va.
  const Member* vaf = find("vaf");
  if (addr(vaf)) ck(t_pull("va", (double*)addr(vaf)), "pull vaf");
  pull_error_status();
  // print / output and restart steps read the whole state on the host (advance.f:35-49)
  const int iprint = iget("iprint", 0), irestart = iget("irestart", 0);
  if ((iprint > 0 && iint % iprint == 0) || (irestart > 0 && iint % irestart == 0)) pull_all();
}

// ================================ routine level (solver.f) ===================================
#define UNIT0(fname, call, ins, outs) \
  void fname(void) { if (!unit_begin(ins)) return; if (ck(call(g_ctx), #call)) unit_end(outs); }
UNIT0(advct_, pomgpu_advct, "u v ub vb aam dt", "advx advy")                                                     // solver.f:201
UNIT0(advu_, pomgpu_advu, "w u v advx drhox ub dt egf egb e_atmos etb etf", "uf")                               // solver.f:734
UNIT0(advv_, pomgpu_advv, "w u v advy drhoy vb dt egf egb e_atmos etb etf", "vf")                               // solver.f:791
void baropg_(void) {              // solver.f:848
  if (!unit_begin("rho rmean dt", true)) return;
  if (ck(multi() ? pomgpu_group_baropg_kind(g_grp, 1) : pomgpu_baropg(g_ctx), "baropg")) unit_end("drhox drhoy rho");
}
void baropg_mcc_(void) {          // solver.f:943
  if (!unit_begin("rho rmean d dt", true)) return;
  if (ck(multi() ? pomgpu_group_baropg_kind(g_grp, 2) : pomgpu_baropg_mcc(g_ctx), "baropg_mcc")) unit_end("drhox drhoy rho");
}
UNIT0(profq_, pomgpu_profq, "t s rho q2b q2lb q2 q2l u v km kh kq uf vf etf wusurf wvsurf wubot wvbot l",
      "uf vf km kh kq l q2b q2lb")                                                                              // solver.f:1212
UNIT0(profu_, pomgpu_profu, "km uf ub vb etf wusurf wubot", "uf wubot")                                         // solver.f:1686
UNIT0(profv_, pomgpu_profv, "km vf vb ub etf wvsurf wvbot", "vf wvbot")                                         // solver.f:1783
UNIT0(vertvl_, pomgpu_vertvl, "u v dt etf etb vfluxb vfluxf w", "w")                                            // solver.f:1970
UNIT0(realvertvl_, pomgpu_realvertvl, "w u v dt et etf etb", "wr")                                              // solver.f:2024

void advave_(void) {              // solver.f:6-198 (mode=2: also the bottom stress, :123-195)
  if (!unit_begin("d ua va uab vab aam2d wubot wvbot")) return;
  if (ck(pomgpu_advave(g_ctx), "advave")) unit_end(iget("mode", 3) == 2 ? "advua advva wubot wvbot" : "advua advva");
}

void advq_(double* qb, double* q, double* qf) {   // solver.f:411-477
  if (!unit_begin("u v w aam dt etb etf")) return;
  Arg a[3] = {{qb}, {q}, {qf}};
  bind_args(a, 3, S3);
  if (!push_arg(a[0]) || !push_arg(a[1]) || !push_arg(a[2])) return;
  if (!ck(pomgpu_advq_fields(g_ctx, a[0].dev.c_str(), a[1].dev.c_str(), a[2].dev.c_str()), "advq")) return;
  pull_arg(a[2], qf);
  unit_end("");
}

static void advt(int nadv, double* fb, double* f, double* fclim, double* ff) {
  if (!unit_begin("u v w aam dt etb etf")) return;
  Arg a[4] = {{fb}, {f}, {fclim}, {ff}};
  bind_args(a, 4, S3);
  for (int q = 0; q < 4; ++q)
    if (!push_arg(a[q])) return;
  const int rc = (nadv == 1) ? pomgpu_advt1(g_ctx, a[0].dev.c_str(), a[1].dev.c_str(), a[2].dev.c_str(), a[3].dev.c_str())
                             : pomgpu_advt2(g_ctx, a[0].dev.c_str(), a[1].dev.c_str(), a[2].dev.c_str(), a[3].dev.c_str());
  if (!ck(rc, nadv == 1 ? "advt1" : "advt2")) return;
  pull_arg(a[3], ff);
  pull_arg(a[0], fb);                 // side effects on fb: level kb and the (fb-fclim)+fclim round trip (solver.f:496,511,532 / 618,691,715)
  if (nadv == 1) pull_arg(a[1], f);   // advt1 also sets f(:,:,kb) (solver.f:495)
  unit_end("");
}
void advt1_(double* fb, double* f, double* fclim, double* ff) { advt(1, fb, f, fclim, ff); }   // solver.f:480
void advt2_(double* fb, double* f, double* fclim, double* ff) { advt(2, fb, f, fclim, ff); }   // solver.f:577

void dens_(double* si, double* ti, double* rhoo) {   // solver.f:1162-1209; initialize.f:416,425: dens(sclim,tclim,rmean), dens(sb,tb,rho)
  if (!unit_begin("", true)) return;
  Arg a[3] = {{si}, {ti}, {rhoo}};
  bind_args(a, 3, S3);
  if (!push_arg(a[0]) || !push_arg(a[1]) || !push_arg(a[2])) return;   // rhoo too: level kb is not assigned (:1175)
  if (!ck(multi() ? pomgpu_group_dens(g_grp, a[0].dev.c_str(), a[1].dev.c_str(), a[2].dev.c_str())
                  : pomgpu_dens(g_ctx, a[0].dev.c_str(), a[1].dev.c_str(), a[2].dev.c_str()), "dens")) return;
  pull_arg(a[2], rhoo);
  unit_end("");
}

void proft_(double* f, double* wfsurf, double* fsurf, int* nbc) {   // solver.f:1541-1683
  if (!unit_begin("kh etf swrad")) return;
  static const char* scr[] = {"s3a", "s2a", "s2b"};
  Arg a[3] = {{f}, {wfsurf}, {fsurf}};
  bind_args(a, 3, scr);
  if (!push_arg(a[0]) || !push_arg(a[1]) || !push_arg(a[2])) return;
  if (!ck(pomgpu_proft(g_ctx, a[0].dev.c_str(), a[1].dev.c_str(), a[2].dev.c_str(), *nbc), "proft")) return;
  pull_arg(a[0], f);
  unit_end("");
}

void smol_adif_(double* xmassflux, double* ymassflux, double* zwflux, double* ff) {   // solver.f:1880-1967
  if (!unit_begin("dt")) return;
  Arg a[4] = {{xmassflux}, {ymassflux}, {zwflux}, {ff}};
  bind_args(a, 4, S3);
  for (int q = 0; q < 4; ++q)
    if (!push_arg(a[q])) return;
  if (!ck(pomgpu_smol_adif(g_ctx, a[0].dev.c_str(), a[1].dev.c_str(), a[2].dev.c_str(), a[3].dev.c_str()), "smol_adif")) return;
  pull_arg(a[0], xmassflux); pull_arg(a[1], ymassflux); pull_arg(a[2], zwflux); pull_arg(a[3], ff);
  unit_end("");
}

// ================================ bounds_forcing.f ===========================================
void bcond_(int* idx) {           // bounds_forcing.f:6-324
  static const char* in[] = {"", "elf", "uaf vaf d el", "", "uf vf t s u v w dt", "w", "uf vf q2 q2l u v"};
  static const char* out[] = {"", "elf", "uaf vaf", "", "uf vf", "w", "uf vf"};
  const int i = *idx;
  if (i < 1 || i > 6 || i == 3) { fail("bcond: this branch is not on the hot path (advance.f calls bcond(1), (2), (4), (6) only)"); return; }
  if (!unit_begin(in[i])) return;
  if (ck(pomgpu_bcond(g_ctx, i), "bcond")) unit_end(out[i]);
}
void bcondorl_(int* idx) {        // bounds_forcing.f:331-590
  const int i = *idx;
  if (i != 3 && i != 5) { fail("bcondorl: this branch is not on the hot path (advance.f calls bcondorl(3) and (5) only)"); return; }
  if (!unit_begin(i == 3 ? "uf vf u v ub vb" : "w")) return;
  if (ck(pomgpu_bcondorl(g_ctx, i), "bcondorl")) unit_end(i == 3 ? "uf vf" : "w");
}

// ================================ parallel_mpi.f =============================================
// single rank: every neighbour is -1 and the reference's exchange does nothing (parallel_mpi.f:171-237)
void exchange2d_mpi_(double* work, int* nx, int* ny) { (void)work; (void)nx; (void)ny; }
void exchange3d_mpi_(double* work, int* nx, int* ny, int* nz) { (void)work; (void)nx; (void)ny; (void)nz; }

// ================================ control ==========================================================
void pomgpu_f_set_devices_(const int* n) {
  if (any_ctx()) { fail("pomgpu_f_set_devices_ must be called before the first entry point"); return; }
  g_ndev = *n;
}
void pomgpu_f_set_dims_(const int* im_local, const int* jm_local, const int* kb) {
  if (any_ctx()) { fail("pomgpu_f_set_dims_ must be called before the first entry point"); return; }
  g_iml = *im_local; g_jml = *jm_local; g_kb = *kb;
  g_resolved = false;
}
void pomgpu_f_set_device_(const int* device) { g_device = *device; }
void pomgpu_f_push_all_(void) { push_all(); }
void pomgpu_f_pull_all_(void) { pull_all(); }
void pomgpu_f_push_(const double* member) {
  if (!ensure_ctx()) return;
  g_t_main = true;
  const Member* m = member_at(member);
  if (!m) { fail("pomgpu_f_push_: not the start of a COMMON array"); return; }
  if (is_restore_member(m->name)) g_restore_pushed = true;
  ck(t_push(m->name.c_str(), member), m->name.c_str());
}
void pomgpu_f_set_restore_(const int* on) { g_restore = (*on < 0) ? -1 : (*on != 0); }
void pomgpu_f_pull_(double* member) {
  if (!ensure_ctx()) return;
  g_t_main = true;
  const Member* m = member_at(member);
  if (!m) { fail("pomgpu_f_pull_: not the start of a COMMON array"); return; }
  ck(t_pull(m->name.c_str(), member), m->name.c_str());
}
void pomgpu_f_finalize_(void) {
  if (g_grp) pomgpu_group_destroy(g_grp);
  for (pomgpu_t* c : g_strips) pomgpu_destroy(c);
  g_strips.clear(); g_grp = nullptr; g_ndev = 0; g_t_main = true;
  if (g_ctx) pomgpu_destroy(g_ctx);
  g_ctx = nullptr; g_full_pushed = false; g_device_ahead = false; g_restore_pushed = false; g_restore = -1;
}
void* pomgpu_f_member(const char* name, long* elems) {
  const Member* m = find(name);
  if (elems) *elems = m ? (long)m->elems : 0;
  return addr(m);
}
char pomgpu_f_member_type(const char* name) { const Member* m = find(name); return m ? m->type : 0; }
const char* pomgpu_f_last_error(void) { return g_err; }

}  // extern "C"
