// pom_step.cu -- orchestration of one internal step (advance.f:21-32) on the
// HBM-resident state, the check_velocity reduction (advance.f:611-641) and the
// C ABI declared in include/pomgpu.h.
#include "pom_core.h"
#include "../../include/pomgpu.h"
#include <cstdlib>

namespace pom {
// pom_state.cu
Ctx* ctx_create(int im, int jm_global, int kb, int j_first, int j_last, int ghost, int device);
void ctx_destroy(Ctx* c);
double** ctx_slot(Ctx* c, const char* name, const FieldInfo** fi);
int ctx_push(Ctx* c, const char* name, const double* host);
int ctx_pull(Ctx* c, const char* name, double* host);
int ctx_set_const(Ctx* c, const char* name, double v);
int ctx_get_const(Ctx* c, const char* name, double* v);
int prof_report(Ctx* c, char* buf, int n);
// kernels
void run_advct(Ctx*, int, int);
void run_baropg(Ctx*, int, int);
void run_smag(Ctx*, int, int);
void run_advave(Ctx*, int, int);
void run_mode_inter_tail(Ctx*, int, int);
void run_ext_elf(Ctx*, int, int);
void run_ext_uv(Ctx*, int iext, int, int);
void run_uvadjust(Ctx*, int, int);
void run_vertvl(Ctx*, int, int);
void run_advq(Ctx*, int, int);
void run_profq(Ctx*, int, int);
void run_qfilter(Ctx*, int, int);
void run_advt(Ctx*, int nadv, const double* fb, const double* f, const double* fc, double* ff, int, int);
void run_fb_roundtrip(Ctx*, double* fb, const double* fc, double* f, int, int);
void run_proft(Ctx*, double* f, const double* wf, const double* fs, int nbc, int, int);
void run_tsfilter(Ctx*, int, int);
void run_dens(Ctx*, const double* si, const double* ti, double* ro, int, int);
void run_advu(Ctx*, int, int);
void run_advv(Ctx*, int, int);
void run_profu(Ctx*, int, int);
void run_profv(Ctx*, int, int);
void run_uvfilter(Ctx*, int, int);
void run_endstep2d(Ctx*, int, int);
void run_realvertvl(Ctx*, int, int);

static inline int J0(Ctx* c) { return c->g.joff + 1; }
static inline int J1(Ctx* c) { return c->g.joff + c->g.jml; }

static int check_switches(Ctx* c) {
  // run-time switches honoured; others follow the reference's error
  // convention: error_status=1 and a message (advance.f:118-119,432-433)
  const Consts& k = c->c;
  const char* bad = nullptr;
  if (k.mode != 3 && k.mode != 4) bad = "mode (3 or 4 supported)";
  else if (k.npg != 1) bad = "npg";
  else if (k.nadv != 1 && k.nadv != 2) bad = "nadv";
  else if (k.nadv == 2 && k.nitera != 1) bad = "nitera (1 supported with nadv=2)";
  else if (k.isplit < 3) bad = "isplit";
  else if (k.ispadv < 1) bad = "ispadv";
  if (bad) {
    snprintf(c->err, sizeof(c->err), "Error: invalid value for %s", bad);
    fprintf(stderr, "\npomgpu: %s\n", c->err);
    c->c.error_status = 1;
    return 2;
  }
  if (k.lrestore && !(c->p.trstrb && c->p.trstrf && c->p.srstrb && c->p.srstrf && c->p.taurstrb && c->p.taurstrf)) {
    snprintf(c->err, sizeof(c->err), "lrestore=1 but the restoring fields were not pushed");
    c->c.error_status = 1;
    return 2;
  }
  return 0;
}

// advance.f:96-141
int lateral_viscosity(Ctx* c) {
  if (c->c.mode != 2) {
    run_advct(c, J0(c), J1(c));
    run_baropg(c, J0(c), J1(c));
    run_smag(c, J0(c), J1(c));
  }
  return 0;
}
// advance.f:144-202 (the vertical integrals were accumulated by the producers)
int mode_interaction(Ctx* c) {
  if (c->c.mode != 2) run_advave(c, J0(c), J1(c));
  run_mode_inter_tail(c, J0(c), J1(c));
  return 0;
}
// advance.f:205-353
int mode_external(Ctx* c, int iext) {
  c->c.iext = iext;
  run_ext_elf(c, J0(c), J1(c));
  if (iext % c->c.ispadv == 0) run_advave(c, J0(c), J1(c));
  run_ext_uv(c, iext, J0(c), J1(c));
  return 0;
}
// advance.f:356-537
int mode_internal(Ctx* c, int iint) {
  c->c.iint = iint;
  Ptrs& p = c->p;
  if ((iint != 1 || c->c.time0 != 0.) && c->c.mode != 2) {
    run_uvadjust(c, J0(c), J1(c));
    run_vertvl(c, J0(c), J1(c));
    run_advq(c, J0(c), J1(c));
    run_profq(c, J0(c), J1(c));
    run_qfilter(c, J0(c), J1(c));
    if (c->c.mode != 4) {
      run_advt(c, c->c.nadv, p.tb, p.t, p.tclim, p.uf, J0(c), J1(c));
      run_advt(c, c->c.nadv, p.sb, p.s, p.sclim, p.vf, J0(c), J1(c));
      run_proft(c, p.uf, p.wtsurf, p.tsurf, c->c.nbct, J0(c), J1(c));
      run_proft(c, p.vf, p.wssurf, p.ssurf, c->c.nbcs, J0(c), J1(c));
      run_tsfilter(c, J0(c), J1(c));
      run_dens(c, p.s, p.t, p.rho, J0(c), J1(c));
    }
    run_advu(c, J0(c), J1(c));
    run_advv(c, J0(c), J1(c));
    run_profu(c, J0(c), J1(c));
    run_profv(c, J0(c), J1(c));
    run_uvfilter(c, J0(c), J1(c));
  }
  run_endstep2d(c, J0(c), J1(c));
  run_realvertvl(c, J0(c), J1(c));
  return 0;
}

// one block of mode_internal (same numbering as the oracle's pomo_internal_stage): lets
// tests compare block by block
int internal_stage(Ctx* c, int iint, int st) {
  c->c.iint = iint;
  Ptrs& p = c->p;
  const bool ts = (c->c.mode != 4);
  switch (st) {
    case 0: run_uvadjust(c, J0(c), J1(c)); break;
    case 1: run_vertvl(c, J0(c), J1(c)); break;
    case 2: run_advq(c, J0(c), J1(c)); break;
    case 3: run_profq(c, J0(c), J1(c)); break;
    case 4: run_qfilter(c, J0(c), J1(c)); break;
    case 5: if (ts) run_advt(c, c->c.nadv, p.tb, p.t, p.tclim, p.uf, J0(c), J1(c)); break;
    case 6: if (ts) run_advt(c, c->c.nadv, p.sb, p.s, p.sclim, p.vf, J0(c), J1(c)); break;
    case 7: if (ts) run_proft(c, p.uf, p.wtsurf, p.tsurf, c->c.nbct, J0(c), J1(c)); break;
    case 8: if (ts) run_proft(c, p.vf, p.wssurf, p.ssurf, c->c.nbcs, J0(c), J1(c)); break;
    case 9: if (ts) run_tsfilter(c, J0(c), J1(c)); break;
    case 10: if (ts) run_dens(c, p.s, p.t, p.rho, J0(c), J1(c)); break;
    case 11: run_advu(c, J0(c), J1(c)); break;
    case 12: run_advv(c, J0(c), J1(c)); break;
    case 13: run_profu(c, J0(c), J1(c)); break;
    case 14: run_profv(c, J0(c), J1(c)); break;
    case 15: run_uvfilter(c, J0(c), J1(c)); break;
    case 16: run_endstep2d(c, J0(c), J1(c)); break;
    case 17: run_realvertvl(c, J0(c), J1(c)); break;
    default: return 2;
  }
  return 0;
}

int step(Ctx* c, int iint, double time, double ramp) {
  c->c.iint = iint; c->c.time = time; c->c.ramp = ramp;
  if (int r = check_switches(c)) return r;
  lateral_viscosity(c);
  mode_interaction(c);
  for (int iext = 1; iext <= c->c.isplit; ++iext) mode_external(c, iext);
  c->c.iext = c->c.isplit + 1;
  mode_internal(c, iint);
  return 0;
}

// ---- check_velocity: max|vaf| (after the substep rotation vaf's values live in va)
#ifndef POMGPU_EMU
__global__ void absmax_kernel(const double* __restrict__ a, size_t n, unsigned long long* out) {
  double m = 0.;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double v = fabs(a[i]);
    if (!(v <= m)) m = v;   // NaN propagates as "larger"
  }
  for (int o = 16; o > 0; o >>= 1) {
    double v = __shfl_xor_sync(0xffffffffu, m, o);
    if (!(v <= m)) m = v;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}
#endif

double check_velocity(Ctx* c) {
  const double* a = c->p.va;
  size_t n = c->g.n2;
  double vamax = 0.;
#ifdef POMGPU_EMU
  for (size_t i = 0; i < n; ++i) { double v = fabs(a[i]); if (!(v <= vamax)) vamax = v; }
#else
  cudaSetDevice(c->device);
  cudaStream_t s = (cudaStream_t)c->stream;
  cudaMemsetAsync(c->d_red, 0, 8, s);
  int blocks = (int)((n + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
  c->launches++;
  absmax_kernel<<<blocks, 256, 0, s>>>(a, n, (unsigned long long*)c->d_red);
  cudaMemcpyAsync(c->h_red, c->d_red, 8, cudaMemcpyDeviceToHost, s);
  if (dev_sync(c)) return 1.0 / 0.0;
  vamax = c->h_red[0];
#endif
  if (!(vamax <= c->c.vmaxl)) c->c.error_status = 1;   // advance.f:631-638
  return vamax;
}

}  // namespace pom

// =============================== C ABI =======================================
using namespace pom;
struct pomgpu { Ctx c; };
static inline Ctx* X(pomgpu_t* p) { return (Ctx*)p; }

extern "C" {

pomgpu_t* pomgpu_create(int im, int jm, int kb, int device) {
  return (pomgpu_t*)ctx_create(im, jm, kb, 1, jm, 0, device);
}
pomgpu_t* pomgpu_create_strip(int im, int jm_global, int kb, int j_first, int j_last, int ghost, int device) {
  if (j_first < 1 || j_last > jm_global || j_last < j_first || ghost < 0) return nullptr;
  return (pomgpu_t*)ctx_create(im, jm_global, kb, j_first, j_last, ghost, device);
}
void pomgpu_destroy(pomgpu_t* p) { ctx_destroy(X(p)); }
int pomgpu_local_rows(const pomgpu_t* p) { return ((const Ctx*)p)->g.jml; }
int pomgpu_row_offset(const pomgpu_t* p) { return ((const Ctx*)p)->g.joff; }
const char* pomgpu_last_error(const pomgpu_t* p) { return ((const Ctx*)p)->err; }
int pomgpu_set_const(pomgpu_t* p, const char* name, double v) { return ctx_set_const(X(p), name, v); }
int pomgpu_get_const(pomgpu_t* p, const char* name, double* v) { return ctx_get_const(X(p), name, v); }
int pomgpu_push(pomgpu_t* p, const char* name, const double* host) { return ctx_push(X(p), name, host); }
int pomgpu_pull(pomgpu_t* p, const char* name, double* host) { return ctx_pull(X(p), name, host); }
long pomgpu_field_elems(pomgpu_t* p, const char* name) {
  const FieldInfo* f = find_field(name);
  return f ? (long)field_elems(X(p), f) : 0;
}
int pomgpu_step(pomgpu_t* p, int iint, double time, double ramp) { return step(X(p), iint, time, ramp); }
int pomgpu_sync(pomgpu_t* p) { return dev_sync(X(p)); }
double pomgpu_check_velocity(pomgpu_t* p) { return check_velocity(X(p)); }
int pomgpu_push_async(pomgpu_t* p, const char* name, const double* host) {
  Ctx* c = X(p);
  const FieldInfo* f;
  double** slot = ctx_slot(c, name, &f);
  if (!slot || !*slot) return 2;
#ifdef POMGPU_EMU
  return dev_h2d(c, *slot, host, field_elems(c, f));
#else
  cudaSetDevice(c->device);
  cudaError_t e = cudaMemcpyAsync(*slot, host, field_elems(c, f) * 8, cudaMemcpyHostToDevice, (cudaStream_t)c->stream);
  if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "push_async(%s): %s", name, cudaGetErrorString(e)); c->c.error_status = 1; return 1; }
  return 0;
#endif
}
int pomgpu_pin_host(void* ptr, size_t bytes) {
#ifdef POMGPU_EMU
  (void)ptr; (void)bytes; return 0;
#else
  return cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) == cudaSuccess ? 0 : 1;
#endif
}
int pomgpu_unpin_host(void* ptr) {
#ifdef POMGPU_EMU
  (void)ptr; return 0;
#else
  return cudaHostUnregister(ptr) == cudaSuccess ? 0 : 1;
#endif
}
// CUDA events on the library's launch stream (bench.py times the step loop with these)
#ifndef POMGPU_EMU
static cudaEvent_t g_ev[8];
static bool g_ev_init = false;
#endif
int pomgpu_event_record(pomgpu_t* p, int slot) {
#ifdef POMGPU_EMU
  (void)p; (void)slot; return 0;
#else
  if (slot < 0 || slot >= 8) return 2;
  cudaSetDevice(X(p)->device);
  if (!g_ev_init) { for (int i = 0; i < 8; ++i) cudaEventCreate(&g_ev[i]); g_ev_init = true; }
  return cudaEventRecord(g_ev[slot], (cudaStream_t)X(p)->stream) == cudaSuccess ? 0 : 1;
#endif
}
double pomgpu_event_elapsed_ms(pomgpu_t* p, int a, int b) {
#ifdef POMGPU_EMU
  (void)p; (void)a; (void)b; return 0.;
#else
  float ms = -1.f;
  cudaSetDevice(X(p)->device);
  cudaEventSynchronize(g_ev[b]);
  cudaEventElapsedTime(&ms, g_ev[a], g_ev[b]);
  return ms;
#endif
}
int pomgpu_profile_begin(pomgpu_t* p) { dev_sync(X(p)); X(p)->prof_on = 1; return 0; }
int pomgpu_profile_end(pomgpu_t* p, char* json, int len) { X(p)->prof_on = 0; return prof_report(X(p), json, len); }
long pomgpu_launch_count(pomgpu_t* p, int reset) {
  long n = X(p)->launches;
  if (reset) X(p)->launches = 0;
  return n;
}

int pomgpu_lateral_viscosity(pomgpu_t* p) { if (int r = check_switches(X(p))) return r; return lateral_viscosity(X(p)); }
int pomgpu_mode_interaction(pomgpu_t* p) { return mode_interaction(X(p)); }
int pomgpu_mode_external(pomgpu_t* p, int iext) { if (int r = check_switches(X(p))) return r; return mode_external(X(p), iext); }
int pomgpu_internal_stage(pomgpu_t* p, int iint, int stage) { return internal_stage(X(p), iint, stage); }
int pomgpu_mode_internal(pomgpu_t* p, int iint) { if (int r = check_switches(X(p))) return r; return mode_internal(X(p), iint); }

#define W0 J0(c), J1(c)
int pomgpu_advave(pomgpu_t* p) { Ctx* c = X(p); run_advave(c, W0); return 0; }
int pomgpu_advct(pomgpu_t* p) { Ctx* c = X(p); run_advct(c, W0); return 0; }
int pomgpu_advq(pomgpu_t* p) { Ctx* c = X(p); run_advq(c, W0); return 0; }
int pomgpu_advu(pomgpu_t* p) { Ctx* c = X(p); run_advu(c, W0); return 0; }
int pomgpu_advv(pomgpu_t* p) { Ctx* c = X(p); run_advv(c, W0); return 0; }
int pomgpu_baropg(pomgpu_t* p) { Ctx* c = X(p); run_baropg(c, W0); return 0; }
int pomgpu_profq(pomgpu_t* p) { Ctx* c = X(p); run_profq(c, W0); return 0; }
int pomgpu_profu(pomgpu_t* p) { Ctx* c = X(p); run_profu(c, W0); return 0; }
int pomgpu_profv(pomgpu_t* p) { Ctx* c = X(p); run_profv(c, W0); return 0; }
int pomgpu_vertvl(pomgpu_t* p) { Ctx* c = X(p); run_vertvl(c, W0); return 0; }
int pomgpu_realvertvl(pomgpu_t* p) { Ctx* c = X(p); run_realvertvl(c, W0); return 0; }

static double* fld(Ctx* c, const char* name) {
  double** s = ctx_slot(c, name, nullptr);
  return s ? *s : nullptr;
}
static int advt(pomgpu_t* p, int nadv, const char* fb, const char* f, const char* fclim, const char* ff) {
  Ctx* c = X(p);
  double *a = fld(c, fb), *b = fld(c, f), *cl = fld(c, fclim), *o = fld(c, ff);
  if (!a || !b || !cl || !o) return 2;
  if (nadv == 2 && c->c.nitera != 1) return 2;
  run_advt(c, nadv, a, b, cl, o, W0);
  run_fb_roundtrip(c, a, cl, nadv == 1 ? b : nullptr, W0);   // side effects on fb (and f for advt1)
  return 0;
}
int pomgpu_advt1(pomgpu_t* p, const char* fb, const char* f, const char* fclim, const char* ff) { return advt(p, 1, fb, f, fclim, ff); }
int pomgpu_advt2(pomgpu_t* p, const char* fb, const char* f, const char* fclim, const char* ff) { return advt(p, 2, fb, f, fclim, ff); }
int pomgpu_dens(pomgpu_t* p, const char* si, const char* ti, const char* rhoo) {
  Ctx* c = X(p);
  double *a = fld(c, si), *b = fld(c, ti), *o = fld(c, rhoo);
  if (!a || !b || !o) return 2;
  run_dens(c, a, b, o, W0);
  return 0;
}
int pomgpu_proft(pomgpu_t* p, const char* f, const char* wfsurf, const char* fsurf, int nbc) {
  Ctx* c = X(p);
  double *a = fld(c, f), *b = fld(c, wfsurf), *s = fld(c, fsurf);
  if (!a || !b || !s) return 2;
  run_proft(c, a, b, s, nbc, W0);
  return 0;
}

}  // extern "C"
