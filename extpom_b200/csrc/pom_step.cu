// pom_step.cu -- orchestration of one internal step (advance.f:21-32) on the HBM-resident
// state of a group of j-strips, the check_velocity reduction (advance.f:611-641) and the
// C ABI declared in include/pomgpu.h.
//
// Every kernel launch is described by the fields it reads (with the j-radius of the read)
// and the fields it writes; pom_halo.cu turns that into launch windows and halo exchanges.
#include "pom_halo.h"
#include "../../include/pomgpu.h"
#include <cstdlib>

namespace pom {
// pom_state.cu
Ctx* ctx_create(int im, int jm_global, int kb, int j_first, int j_last, int ghost, int device);
void ctx_destroy(Ctx* c);
double** ctx_slot(Ctx* c, const char* name, const FieldInfo** fi);
int ctx_push(Ctx* c, const char* name, const double* host);
int ctx_pull(Ctx* c, const char* name, double* host);
int ctx_push_rows(Ctx* c, const char* name, const double* host, int row0, int nrows);
int ctx_set_const(Ctx* c, const char* name, double v);
int ctx_get_const(Ctx* c, const char* name, double* v);
int prof_report(Ctx* c, char* buf, int n);
// kernels
void run_advct(Ctx*, int, int);
void run_baropg(Ctx*, int, int);
void run_baropg_mcc(Ctx*, int, int);
void run_smag(Ctx*, int, int);
void run_advave(Ctx*, int, int);
void run_mode_inter_tail(Ctx*, int, int);
void run_ext_step(Ctx*, int iext, int do_adv, int, int);
void run_uvadjust(Ctx*, int, int);
void run_vertvl(Ctx*, int, int);
void run_uvadj_vertvl(Ctx*, int, int);
void run_uvsum(Ctx*, int, int);
void run_advq(Ctx*, int, int);
void run_profq(Ctx*, int fuse_filter, int, int);
void run_qfilter(Ctx*, int, int);
void run_advt(Ctx*, int nadv, const double* fb, const double* f, const double* fc, double* ff, int, int);
void run_fb_roundtrip(Ctx*, double* fb, const double* fc, double* f, int, int);
void run_advt2_up(Ctx*, const double* fbm, const double* f, const double* xm, const double* ym, const double* zw,
                  const double* stale, double* ff, int first, int, int);
void run_smol_adif(Ctx*, const double* ff, double* xm, double* ym, double* zw, int first, int, int);
void run_advt2_diff(Ctx*, const double* fb, const double* fc, double* ff, int, int);
void run_proft(Ctx*, double* f, const double* wf, const double* fs, int nbc, int, int);
void run_tsfilter(Ctx*, int with_dens, int, int);
void run_proft_ts(Ctx*, int fuse, int, int);
void run_advt2_ts(Ctx*, int, int);
void run_advprof_u(Ctx*, int, int);
void run_advprof_v(Ctx*, int, int);
void run_dens(Ctx*, const double* si, const double* ti, double* ro, int, int);
void run_advu(Ctx*, int, int);
void run_advv(Ctx*, int, int);
void run_profu(Ctx*, int, int);
void run_profv(Ctx*, int, int);
void run_uvfilter(Ctx*, int, int);
void run_endstep2d(Ctx*, int, int);
void run_realvertvl(Ctx*, int, int);
int domain_stats_rows(Ctx* c, double* rows);
int run_bcond(Ctx*, int idx, int orl, int, int);
void run_mask_fsm(Ctx*, double* ff, int, int);

// ---- launch windows ---------------------------------------------------------------------
static inline int WLO(const Ctx* c, int e) { return c->jown0 > 1 ? c->jown0 - e : 1; }
static inline int WHI(const Ctx* c, int e) { return c->jown1 < c->g.jmg ? c->jown1 + e : c->g.jmg; }
static inline double* FP(Ctx* c, int f) {
  int n;
  const FieldInfo* t = field_table(&n);
  return *(double**)((char*)&c->p + t[f].offset);
}
#define NEED(...) ([&]() { const Req rq[] = {__VA_ARGS__}; return group_need(G, rq, (int)(sizeof(rq) / sizeof(rq[0]))); }())
// the fields the kernel about to be launched WRITES, declared before NEED: they are left out of the
// exchange it may trigger (the kernel recomputes their ghost rows anyway), which is what allows the
// packs to run on the communication stream while the kernel's interior part already writes
#define WILL(...) do { const int wl[] = {__VA_ARGS__}; group_will(G, wl, (int)(sizeof(wl) / sizeof(wl[0]))); } while (0)
#define MADE(e, ...) do { const int ou[] = {__VA_ARGS__}; group_produced(G, e, ou, (int)(sizeof(ou) / sizeof(ou[0]))); } while (0)
// (after a failed halo exchange nothing is launched any more: the kernels would read stale ghost rows)
// One launch per strip over its window [j0, j1] = owned rows + e ghost rows -- or, while a halo
// exchange is in flight on the communication stream, three: the interior first, the seam bands once
// the rows have arrived (pom_halo.cu: group_window).
#define EACH(stmt) do {                                                                         \
    const int np_ = group_parts(G);                                                             \
    for (int part_ = 0; part_ < np_ && !G->failed; ++part_) {                                   \
      if (part_ == 1) group_wait_halo(G);                                                       \
      for (int r_ = 0; r_ < G->n; ++r_) {                                                       \
        Ctx* c = G->c[r_];                                                                      \
        int j0, j1;                                                                             \
        if (!group_window(G, c, e, part_, np_, &j0, &j1)) continue;                             \
        stmt;                                                                                   \
      }                                                                                         \
    }                                                                                           \
    group_launched(G);                                                                          \
  } while (0)
// the scalars of strip 0 are the group's; error_status stays per strip (a CUDA failure or a blow-up
// on strip r must survive until the driver reads it, advance.f:556-563)
#define CSYNC() for (int r_ = 1; r_ < G->n; ++r_) { const int es_ = G->c[r_]->c.error_status; G->c[r_]->c = G->c[0]->c; G->c[r_]->c.error_status = es_; }
static int group_status(Group* G) {
  int es = G->failed ? 1 : 0;
  for (int r = 0; r < G->n; ++r) es |= G->c[r]->c.error_status;
  return es;
}

// pushes made with pomgpu_push_async since the last step: make the compute stream wait for the
// copies and swap the shadow buffers in (the old buffers become the next shadows)
static void apply_pending(Ctx* c) {
  if (!c->npending) return;
#ifndef POMGPU_EMU
  cudaSetDevice(c->device);
  cudaStreamWaitEvent((cudaStream_t)c->stream, (cudaEvent_t)c->ev_copied, 0);
#endif
  int ntab;
  const FieldInfo* tab = field_table(&ntab);
  for (int f = 0; f < ntab && f < 256; ++f)
    if (c->pending[f]) {
      double** slot = (double**)((char*)&c->p + tab[f].offset);
      double* t = *slot; *slot = c->shadow[f]; c->shadow[f] = t;
      c->pending[f] = 0;
    }
  c->npending = 0;
#ifndef POMGPU_EMU
  // copies into the (new) shadows must wait until everything enqueued so far has read them
  cudaEventRecord((cudaEvent_t)c->ev_swapped, (cudaStream_t)c->stream);
#endif
}
static void apply_pending(Group* G) { for (int r = 0; r < G->n; ++r) apply_pending(G->c[r]); }

static int check_switches(Group* G) {
  // run-time switches honoured; others follow the reference's error
  // convention: error_status=1 and a message (advance.f:118-119,432-433)
  Ctx* c = G->c[0];
  const Consts& k = c->c;
  const char* bad = nullptr;
  if (k.mode != 2 && k.mode != 3 && k.mode != 4) bad = "mode (2, 3 or 4)";
  else if (k.npg != 1 && k.npg != 2) bad = "npg";
  else if (k.nadv != 1 && k.nadv != 2) bad = "nadv";
  else if (k.nadv == 2 && k.nitera < 1) bad = "nitera";
  else if (k.isplit < 3) bad = "isplit";
  else if (k.ispadv < 1) bad = "ispadv";
  if (bad) {
    snprintf(c->err, sizeof(c->err), "Error: invalid value for %s", bad);
    fprintf(stderr, "\npomgpu: %s\n", c->err);
    for (int r = 0; r < G->n; ++r) G->c[r]->c.error_status = 1;
    return 2;
  }
  if (k.lrestore && !(c->p.trstrb && c->p.trstrf && c->p.srstrb && c->p.srstrf && c->p.taurstrb && c->p.taurstrf)) {
    snprintf(c->err, sizeof(c->err), "lrestore=1 but the restoring fields were not pushed");
    c->c.error_status = 1;
    return 2;
  }
  return 0;
}

// ---- the kernels of the path, with what they read (field, j-radius) and write ------------
static void k_advct(Group* G) {
  WILL(F_advx, F_advy, F_adx2d, F_ady2d);
  int e = NEED({F_u, 1}, {F_v, 1}, {F_ub, 1}, {F_vb, 1}, {F_aam, 1}, {F_dt, 2});
  EACH(run_advct(c, j0, j1));
  MADE(e, F_advx, F_advy, F_adx2d, F_ady2d);
}
static void k_baropg(Group* G, int npg) {
  // (drhox/drhoy/aam/w are also read in place on the i=1,im columns, whose values never change)
  // npg=2: baropg_mcc reads rho-rmean and d two rows away (order2d/3d_mpi in the reference)
  WILL(F_drhox, F_drhoy, F_drx2d, F_dry2d, F_rho2);
  int e = (npg == 2) ? NEED({F_rho, 2}, {F_d, 2}, {F_dt, 1}) : NEED({F_rho, 1}, {F_dt, 1});
  if (npg == 2) { EACH(run_baropg_mcc(c, j0, j1)); } else { EACH(run_baropg(c, j0, j1)); }
  MADE(e, F_drhox, F_drhoy, F_drx2d, F_dry2d, F_rho2);
  group_swap(G, F_rho, F_rho2);   // rho <- (rho-rmean)+rmean (solver.f:854,937)
}
static void k_smag(Group* G) {
  WILL(F_aam, F_aam2d);
  int e = NEED({F_u, 1}, {F_v, 1});
  EACH(run_smag(c, j0, j1));
  MADE(e, F_aam, F_aam2d);
}
static void k_advave(Group* G) {
  WILL(F_advua, F_advva);
  int e = NEED({F_d, 2}, {F_ua, 1}, {F_va, 1}, {F_uab, 1}, {F_vab, 1}, {F_aam2d, 1});
  EACH(run_advave(c, j0, j1));
  MADE(e, F_advua, F_advva);
}
static void k_mode_inter_tail(Group* G) {
  WILL(F_adx2d, F_ady2d, F_egf, F_utf, F_vtf);
  int e = NEED({F_adx2d, 0}, {F_ady2d, 0}, {F_advua, 0}, {F_advva, 0}, {F_el, 0}, {F_ua, 0}, {F_va, 0}, {F_d, 1});
  EACH(run_mode_inter_tail(c, j0, j1));
  MADE(e, F_adx2d, F_ady2d, F_egf, F_utf, F_vtf);
}
// one external substep (advance.f:205-353) in one kernel; output row j depends on the staged
// operands of rows j-2..j+1 (elf(j-1) <- transports(j-1) <- d,va(j-2))
static void k_ext_step(Group* G, int iext, int do_adv) {
  const int isplit = G->c[0]->c.isplit;
  // (ua, va are read with radius 2; the kernel only seeds their four physical corner cells)
  WILL(F_elf, F_uaf, F_vaf, F_s2a, F_s2b, F_el2, F_d2, F_etf, F_egf, F_utf, F_vtf);
  int e = NEED({F_d, 2}, {F_ua, 2}, {F_va, 2}, {F_uab, 2}, {F_vab, 2}, {F_aam2d, 2}, {F_el, 1}, {F_elb, 1},
               {F_adx2d, 0}, {F_ady2d, 0}, {F_drx2d, 0}, {F_dry2d, 0}, {F_wubot, 0}, {F_wvbot, 0},
               {F_egf, 0}, {F_utf, 0}, {F_vtf, 0});
  if (!do_adv) { const Req q[] = {{F_advua, 0}, {F_advva, 0}}; int e2 = group_need(G, q, 2); if (e2 < e) e = e2; }
  if (iext > isplit - 2) { const Req q[] = {{F_etf, 0}}; int e2 = group_need(G, q, 1); if (e2 < e) e = e2; }
  EACH(run_ext_step(c, iext, do_adv, j0, j1));
  MADE(e, F_elf, F_uaf, F_vaf, F_s2a, F_s2b, F_el2, F_d2, F_ua, F_va);
  if (do_adv) MADE(e, F_advua, F_advva);
  if (do_adv && G->c[0]->c.mode == 2) MADE(e, F_wubot, F_wvbot);
  if (iext >= isplit - 2) MADE(e, F_etf);
  if (iext != isplit) MADE(e, F_egf, F_utf, F_vtf);
  // time rotation (advance.f:324-330)
  group_swap(G, F_ua, F_uaf); group_swap(G, F_va, F_vaf);
  group_swap(G, F_uab, F_s2a); group_swap(G, F_vab, F_s2b);
  group_swap(G, F_elb, F_el2); group_swap(G, F_el, F_elf);
  group_swap(G, F_d, F_d2);
}
static void k_uvadjust(Group* G) {
  WILL(F_u, F_v);
  int e = NEED({F_u, 0}, {F_v, 0}, {F_utb, 0}, {F_utf, 0}, {F_vtb, 0}, {F_vtf, 0}, {F_dt, 1});
  EACH(run_uvadjust(c, j0, j1));
  MADE(e, F_u, F_v);
  for (int r = 0; r < G->n; ++r) G->c[r]->uvsum_ok = 0;   // u, v changed in place
}
// stages 0+1 in one sweep (what the step runs); the depth sums come from the last uv_filter
static void k_uvadjust_vertvl(Group* G) {
  bool ok = true;
  for (int r = 0; r < G->n; ++r) ok = ok && G->c[r]->uvsum_ok;
  if (!ok) {
    // (a process whose driver pushed u or v; the others may not take this branch, so the ghost-row
    // validity it records must be the one uv_filter left -- exchange decisions stay identical on
    // every rank: u, v are at least as valid as the sums uv_filter produced with them)
    int e = NEED({F_u, 0}, {F_v, 0});
    EACH(run_uvsum(c, j0, j1));
    int eo = e;
    if (G->valid[F_s2c] < eo) eo = G->valid[F_s2c];
    if (G->valid[F_s2d] < eo) eo = G->valid[F_s2d];
    MADE(eo, F_s2c, F_s2d);
  }
  int e = NEED({F_u, 0}, {F_v, 1}, {F_s2c, 0}, {F_s2d, 1}, {F_utb, 0}, {F_utf, 0}, {F_vtb, 1}, {F_vtf, 1}, {F_dt, 1},
               {F_etf, 0}, {F_etb, 0}, {F_vfluxb, 0}, {F_w, 0});
  EACH(run_uvadj_vertvl(c, j0, j1));
  MADE(e, F_s3a, F_s3b, F_w);
  group_swap(G, F_u, F_s3a); group_swap(G, F_v, F_s3b);
  for (int r = 0; r < G->n; ++r) G->c[r]->uvsum_ok = 0;
}
static void k_vertvl(Group* G) {
  WILL(F_w);
  int e = NEED({F_u, 0}, {F_v, 1}, {F_dt, 1}, {F_etf, 0}, {F_etb, 0}, {F_vfluxb, 0});
  EACH(run_vertvl(c, j0, j1));
  MADE(e, F_w);
}
static void k_advq(Group* G) {
  WILL(F_uf, F_vf);
  int e = NEED({F_q2, 1}, {F_q2b, 1}, {F_q2l, 1}, {F_q2lb, 1}, {F_u, 0}, {F_v, 1}, {F_aam, 1}, {F_dt, 1},
               {F_w, 0}, {F_etb, 0}, {F_etf, 0});
  EACH(run_advq(c, j0, j1));
  MADE(e, F_uf, F_vf);
}
static void k_profq(Group* G, int fuse_filter) {
  WILL(F_uf, F_vf, F_km, F_kh, F_kq, F_l, F_q2b, F_q2lb);
  int e = NEED({F_t, 0}, {F_s, 0}, {F_rho, 0}, {F_q2b, 0}, {F_q2lb, 0}, {F_q2, 0}, {F_u, 0}, {F_v, 1},
               {F_km, 0}, {F_kh, 0}, {F_kq, 0}, {F_uf, 0}, {F_vf, 0}, {F_etf, 0}, {F_wubot, 0}, {F_wvbot, 1});
  if (fuse_filter) { const Req q[] = {{F_q2l, 0}}; int e2 = group_need(G, q, 1); if (e2 < e) e = e2; }
  EACH(run_profq(c, fuse_filter, j0, j1));
  MADE(e, F_uf, F_vf, F_km, F_kh, F_kq, F_l, F_q2b, F_q2lb);
  if (fuse_filter) { group_swap(G, F_q2, F_uf); group_swap(G, F_q2l, F_vf); }   // advance.f:418-421
}
static void k_qfilter(Group* G) {
  WILL(F_uf, F_vf, F_q2b, F_q2lb);
  int e = NEED({F_uf, 0}, {F_vf, 0}, {F_q2, 0}, {F_q2b, 0}, {F_q2l, 0}, {F_q2lb, 0}, {F_u, 0}, {F_v, 0});
  EACH(run_qfilter(c, j0, j1));
  MADE(e, F_uf, F_vf, F_q2b, F_q2lb);
  group_swap(G, F_q2, F_uf); group_swap(G, F_q2l, F_vf);   // advance.f:418-421
}
// advt2 with nitera > 1 (solver.f:625-687): { upwind step + mask, smol_adif }
// per iteration on scratch fields, then the diffusion.  `stale` = the field whose values the
// reference's ff array holds where the scheme never assigns it (boundary columns, level kb):
// in the step that is the new q2 / q2l, which `q2=uf` left in uf (advance.f:419-421).
static void k_advt2_iter(Group* G, int fb, int f, int fc, int ff, int stale) {
  const int XM = F_s3c, YM = F_s3d, ZW = F_s3e, PING = F_s3a;
  const int nitera = G->c[0]->c.nitera;
  // (first iteration: the mass fluxes of :602-621 are formed inside the kernels from u, v, w)
  int src = fb;
  for (int it = 1; it <= nitera; ++it) {
    const int dst = ((nitera - it) % 2 == 0) ? ff : PING;   // the last iterate lands in ff
    const int first = (it == 1);
    const int fx = first ? F_u : XM, fy = first ? F_v : YM, fz = first ? F_w : ZW;
    {
      int e = NEED({src, 1}, {f, 0}, {fx, 0}, {fy, 1}, {fz, 0}, {stale, 0}, {F_w, 0}, {F_dt, 1}, {F_etb, 0}, {F_etf, 0});
      EACH(run_advt2_up(c, FP(c, src), FP(c, f), FP(c, XM), FP(c, YM), FP(c, ZW), FP(c, stale), FP(c, dst), first, j0, j1));
      MADE(e, dst);
    }
    if (it < nitera) {   // the fluxes after the last iteration are never used
      int e = NEED({dst, 1}, {fx, 0}, {fy, 0}, {fz, 0}, {F_dt, 1});
      EACH(run_smol_adif(c, FP(c, dst), FP(c, XM), FP(c, YM), FP(c, ZW), first, j0, j1));
      MADE(e, XM, YM, ZW);
    }
    src = dst;
  }
  int e = NEED({fb, 1}, {fc, 1}, {F_aam, 1}, {ff, 0}, {F_etf, 0});
  EACH(run_advt2_diff(c, FP(c, fb), FP(c, fc), FP(c, ff), j0, j1));
  MADE(e, ff);
}
static void k_advt(Group* G, int fb, int f, int fc, int ff, int stale) {
  const Consts& k = G->c[0]->c;
  if (k.nadv == 2 && k.nitera > 1) { k_advt2_iter(G, fb, f, fc, ff, stale); return; }
  int e = NEED({fb, 1}, {f, 0}, {F_u, 0}, {F_v, 1}, {F_w, 0}, {F_aam, 1}, {F_dt, 1}, {F_etb, 0}, {F_etf, 0});
  EACH(run_advt(c, c->c.nadv, FP(c, fb), FP(c, f), FP(c, fc), FP(c, ff), j0, j1));
  MADE(e, ff);
}
static void k_proft(Group* G, int f, int wf, int fs, int nbc) {
  WILL(f);
  int e = NEED({f, 0}, {F_kh, 0}, {F_etf, 0});
  EACH(run_proft(c, FP(c, f), FP(c, wf), FP(c, fs), nbc, j0, j1));
  MADE(e, f);
}
static void k_advt2_ts(Group* G) {   // advt2(tb,t,tclim,uf) and advt2(sb,s,sclim,vf) in one pass (nitera=1)
  WILL(F_uf, F_vf);
  int e = NEED({F_tb, 1}, {F_t, 0}, {F_sb, 1}, {F_s, 0}, {F_u, 0}, {F_v, 1}, {F_w, 0}, {F_aam, 1}, {F_dt, 1},
               {F_etb, 0}, {F_etf, 0});
  EACH(run_advt2_ts(c, j0, j1));
  MADE(e, F_uf, F_vf);
}
static void k_proft_ts(Group* G) {   // proft(uf,wtsurf,tsurf,nbct) and proft(vf,wssurf,ssurf,nbcs) in one pass
  WILL(F_uf, F_vf);
  int e = NEED({F_uf, 0}, {F_vf, 0}, {F_kh, 0}, {F_etf, 0});
  EACH(run_proft_ts(c, 0, j0, j1));
  MADE(e, F_uf, F_vf);
}
// proft of T and S + bcond(4) + t/s filter + restore_interior + dens in one kernel (advance.f:439-454)
static void k_proft_tsfilter(Group* G) {
  WILL(F_uf, F_vf, F_tb, F_sb, F_rho);
  int e = NEED({F_uf, 0}, {F_vf, 0}, {F_kh, 0}, {F_etf, 0}, {F_t, 0}, {F_s, 0}, {F_tb, 0}, {F_sb, 0}, {F_u, 0},
               {F_v, 0}, {F_w, 0}, {F_dt, 0});
  EACH(run_proft_ts(c, 1, j0, j1));
  MADE(e, F_uf, F_vf, F_tb, F_sb, F_rho);
  group_swap(G, F_t, F_uf); group_swap(G, F_s, F_vf);      // advance.f:446-449
}
static void k_tsfilter(Group* G, int with_dens) {
  WILL(F_uf, F_vf, F_tb, F_sb, F_rho);
  int e = NEED({F_uf, 0}, {F_vf, 0}, {F_t, 0}, {F_s, 0}, {F_tb, 0}, {F_sb, 0}, {F_u, 0}, {F_v, 0}, {F_w, 0}, {F_dt, 0});
  EACH(run_tsfilter(c, with_dens, j0, j1));
  MADE(e, F_uf, F_vf, F_tb, F_sb);
  if (with_dens) MADE(e, F_rho);
  group_swap(G, F_t, F_uf); group_swap(G, F_s, F_vf);      // advance.f:446-449
}
static void k_dens(Group* G, int si, int ti, int ro) {
  WILL(ro);
  int e = NEED({si, 0}, {ti, 0});
  EACH(run_dens(c, FP(c, si), FP(c, ti), FP(c, ro), j0, j1));
  MADE(e, ro);
}
static void k_advu(Group* G) {
  WILL(F_uf);
  int e = NEED({F_w, 0}, {F_u, 0}, {F_v, 1}, {F_advx, 0}, {F_drhox, 0}, {F_ub, 0}, {F_dt, 0}, {F_egf, 0},
               {F_egb, 0}, {F_etb, 0}, {F_etf, 0});
  EACH(run_advu(c, j0, j1));
  MADE(e, F_uf);
}
static void k_advv(Group* G) {
  WILL(F_vf);
  int e = NEED({F_w, 1}, {F_v, 0}, {F_u, 1}, {F_advy, 0}, {F_drhoy, 0}, {F_vb, 0}, {F_dt, 1}, {F_egf, 1},
               {F_egb, 1}, {F_etb, 1}, {F_etf, 1});
  EACH(run_advv(c, j0, j1));
  MADE(e, F_vf);
}
static void k_profu(Group* G) {
  WILL(F_uf, F_wubot);
  int e = NEED({F_km, 0}, {F_uf, 0}, {F_ub, 0}, {F_vb, 1}, {F_etf, 0}, {F_wubot, 0});
  EACH(run_profu(c, j0, j1));
  MADE(e, F_uf, F_wubot);
}
static void k_profv(Group* G) {
  WILL(F_vf, F_wvbot);
  int e = NEED({F_km, 1}, {F_vf, 0}, {F_vb, 0}, {F_ub, 1}, {F_etf, 1}, {F_wvbot, 0});
  EACH(run_profv(c, j0, j1));
  MADE(e, F_vf, F_wvbot);
}
// advu+profu and advv+profv fused (the order advu, advv, profu, profv of advance.f:459-462 does
// not matter: each pair only reads u, v, ub, vb, w and writes its own uf / vf, wubot / wvbot)
static void k_advprof_u(Group* G) {
  WILL(F_uf, F_wubot);
  int e = NEED({F_w, 0}, {F_u, 0}, {F_v, 1}, {F_advx, 0}, {F_drhox, 0}, {F_ub, 0}, {F_vb, 1}, {F_km, 0}, {F_dt, 0},
               {F_egf, 0}, {F_egb, 0}, {F_etb, 0}, {F_etf, 0}, {F_wubot, 0});
  EACH(run_advprof_u(c, j0, j1));
  MADE(e, F_uf, F_wubot);
}
static void k_advprof_v(Group* G) {
  WILL(F_vf, F_wvbot);
  int e = NEED({F_w, 1}, {F_v, 0}, {F_u, 1}, {F_advy, 0}, {F_drhoy, 0}, {F_vb, 0}, {F_ub, 1}, {F_km, 1}, {F_dt, 1},
               {F_egf, 1}, {F_egb, 1}, {F_etb, 1}, {F_etf, 1}, {F_wvbot, 0});
  EACH(run_advprof_v(c, j0, j1));
  MADE(e, F_vf, F_wvbot);
}
static void k_uvfilter(Group* G) {
  WILL(F_uf, F_vf, F_s3a, F_s3b, F_s2c, F_s2d);
  int e = NEED({F_uf, 0}, {F_vf, 0}, {F_u, 0}, {F_v, 0}, {F_ub, 0}, {F_vb, 0});
  EACH(run_uvfilter(c, j0, j1));
  MADE(e, F_uf, F_vf, F_s3a, F_s3b, F_s2c, F_s2d);
  for (int r = 0; r < G->n; ++r) G->c[r]->uvsum_ok = 1;   // s2c, s2d = depth sums of what becomes u, v
  group_swap(G, F_u, F_uf); group_swap(G, F_v, F_vf);      // advance.f:511-514
  group_swap(G, F_ub, F_s3a); group_swap(G, F_vb, F_s3b);
}
static void k_endstep2d(Group* G) {
  WILL(F_egb, F_etb, F_et, F_dt, F_utb, F_vtb, F_vfluxb);
  int e = NEED({F_egf, 0}, {F_et, 0}, {F_etf, 0}, {F_utf, 0}, {F_vtf, 0});
  EACH(run_endstep2d(c, j0, j1));
  MADE(e, F_egb, F_etb, F_et, F_dt, F_utb, F_vtb, F_vfluxb);
}
static void k_realvertvl(Group* G) {
  WILL(F_wr);
  int e = NEED({F_w, 0}, {F_u, 0}, {F_v, 1}, {F_dt, 1}, {F_et, 1}, {F_etf, 0}, {F_etb, 0});
  EACH(run_realvertvl(c, j0, j1));
  MADE(e, F_wr);
}

// ---- advance.f:96-141 ---------------------------------------------------------------------
static int lateral_viscosity(Group* G) {
  if (G->c[0]->c.mode != 2) { k_advct(G); k_baropg(G, G->c[0]->c.npg); k_smag(G); }
  return 0;
}
// advance.f:144-202 (the vertical integrals were accumulated by the producers)
static int mode_interaction(Group* G) {
  if (G->c[0]->c.mode != 2) k_advave(G);
  k_mode_inter_tail(G);
  return 0;
}
// advance.f:205-353
static int mode_external(Group* G, int iext) {
  G->c[0]->c.iext = iext; CSYNC();
  k_ext_step(G, iext, iext % G->c[0]->c.ispadv == 0);
  return 0;
}
// one block of mode_internal (same numbering as the oracle's pomo_internal_stage)
static int internal_stage(Group* G, int iint, int st) {
  G->c[0]->c.iint = iint; CSYNC();
  const Consts& k = G->c[0]->c;
  const bool ts = (k.mode != 4);
  switch (st) {
    case 0: k_uvadjust(G); break;
    case 1: k_vertvl(G); break;
    case 100: k_uvadjust_vertvl(G); break;          // stages 0-1 in one sweep (what the step runs)
    case 2: k_advq(G); break;
    case 3: k_profq(G, 0); break;
    case 103: k_profq(G, 1); break;                   // profq + bcond(6) + q filter fused (what the step runs)
    case 4: k_qfilter(G); break;
    case 5: if (ts) k_advt(G, F_tb, F_t, F_tclim, F_uf, F_q2); break;
    case 6: if (ts) k_advt(G, F_sb, F_s, F_sclim, F_vf, F_q2l); break;
    case 7: if (ts) k_proft(G, F_uf, F_wtsurf, F_tsurf, k.nbct); break;
    case 8: if (ts) k_proft(G, F_vf, F_wssurf, F_ssurf, k.nbcs); break;
    case 9: if (ts) k_tsfilter(G, 0); break;
    case 105: if (ts) { if (k.nadv == 2 && k.nitera == 1) k_advt2_ts(G); else { internal_stage(G, iint, 5); internal_stage(G, iint, 6); } } break;
    case 107: if (ts) k_proft_ts(G); break;
    case 207: if (ts) k_proft_tsfilter(G); break;     // stages 7-10 in one kernel (what the step runs)
    case 111: k_advprof_u(G); k_advprof_v(G); break;   // stages 11-14 as two fused kernels (what the step runs)           // proft of T and S fused (what the step runs)
    case 109: if (ts) k_tsfilter(G, 1); break;        // + dens fused (what the step runs)
    case 10: if (ts) k_dens(G, F_s, F_t, F_rho); break;
    case 11: k_advu(G); break;
    case 12: k_advv(G); break;
    case 13: k_profu(G); break;
    case 14: k_profv(G); break;
    case 15: k_uvfilter(G); break;
    case 16: k_endstep2d(G); break;
    case 17: k_realvertvl(G); break;
    default: return 2;
  }
  return 0;
}
// advance.f:356-537
static int mode_internal(Group* G, int iint) {
  const Consts& k = G->c[0]->c;
  if ((iint != 1 || k.time0 != 0.) && k.mode != 2)
    for (int st = 0; st <= 15; ++st) {
      if (st == 0) { internal_stage(G, iint, 100); ++st; continue; }   // u,v adjustment + vertvl in one sweep
      if (st == 3) { internal_stage(G, iint, 103); ++st; continue; }   // profq with the q2/q2l filter fused
      if (st == 5) { internal_stage(G, iint, 105); ++st; continue; }   // advt2 of T and S in one kernel
      if (st == 7) { internal_stage(G, iint, 207); st = 10; continue; }  // proft T,S + t/s filter + dens in one kernel
      if (st == 11) { internal_stage(G, iint, 111); st = 14; continue; }  // advu+profu, advv+profv
      internal_stage(G, iint, st);
    }
  internal_stage(G, iint, 16);
  internal_stage(G, iint, 17);
  return 0;
}

static int step(Group* G, int iint, double time, double ramp) {
  Ctx* c0 = G->c[0];
  apply_pending(G);
  c0->c.iint = iint; c0->c.time = time; c0->c.ramp = ramp; CSYNC();
  if (int r = check_switches(G)) return r;
  lateral_viscosity(G);
  mode_interaction(G);
  for (int iext = 1; iext <= c0->c.isplit; ++iext) mode_external(G, iext);
  c0->c.iext = c0->c.isplit + 1; CSYNC();
  mode_internal(G, iint);
  return group_status(G) ? 1 : 0;
}

// ---- check_velocity: max|vaf| over the owned rows (after the rotation vaf lives in va) ----
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

static double check_velocity(Ctx* c) {
  const double* a = c->p.va + (size_t)(c->jown0 - 1 - c->g.joff) * c->g.im;
  size_t n = (size_t)(c->jown1 - c->jown0 + 1) * c->g.im;
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
static inline Ctx* X(pomgpu_t* p) { return (Ctx*)p; }
static inline Group* GG(pomgpu_group_t* g) { return (Group*)g; }
// every context owns a group of one, so the single-strip entry points share the code path
static Group* self_group(Ctx* c) {
  if (!c->self) { Ctx* a[1] = {c}; c->self = group_create(1, a); }
  return (Group*)c->self;
}

extern "C" {

pomgpu_t* pomgpu_create(int im, int jm, int kb, int device) {
  return (pomgpu_t*)ctx_create(im, jm, kb, 1, jm, 0, device);
}
pomgpu_t* pomgpu_create_strip(int im, int jm_global, int kb, int j_first, int j_last, int ghost, int device) {
  if (j_first < 1 || j_last > jm_global || j_last < j_first || ghost < 0) return nullptr;
  if ((j_first > 1 || j_last < jm_global) && ghost < 2) return nullptr;   // advct reads dt two rows away
  return (pomgpu_t*)ctx_create(im, jm_global, kb, j_first, j_last, ghost, device);
}
void pomgpu_destroy(pomgpu_t* p) {
  if (!p) return;
  if (X(p)->self) group_destroy((Group*)X(p)->self);
  ctx_destroy(X(p));
}
int pomgpu_local_rows(const pomgpu_t* p) { return ((const Ctx*)p)->g.jml; }
int pomgpu_row_offset(const pomgpu_t* p) { return ((const Ctx*)p)->g.joff; }
const char* pomgpu_last_error(const pomgpu_t* p) { return ((const Ctx*)p)->err; }
int pomgpu_set_const(pomgpu_t* p, const char* name, double v) { return ctx_set_const(X(p), name, v); }
int pomgpu_get_const(pomgpu_t* p, const char* name, double* v) { return ctx_get_const(X(p), name, v); }
static void touched(Ctx* c, const char* name) {   // a host push of u or v makes uv_filter's depth sums stale
  if (!strcmp(name, "u") || !strcmp(name, "v") || !strcmp(name, "s2c") || !strcmp(name, "s2d")) c->uvsum_ok = 0;
}
int pomgpu_push(pomgpu_t* p, const char* name, const double* host) { apply_pending(X(p)); touched(X(p), name); return ctx_push(X(p), name, host); }
int pomgpu_pull(pomgpu_t* p, const char* name, double* host) { apply_pending(X(p)); return ctx_pull(X(p), name, host); }
int pomgpu_push_rows(pomgpu_t* p, const char* name, const double* host, int row0, int nrows) {
  apply_pending(X(p));
  touched(X(p), name);
  return ctx_push_rows(X(p), name, host, row0, nrows);
}
long pomgpu_field_global_elems(pomgpu_t* p, const char* name) {
  const FieldInfo* f = find_field(name);
  return f ? (long)field_global_elems(X(p), f) : 0;
}
int pomgpu_push_global(pomgpu_t* p, const char* name, const double* host) { apply_pending(X(p)); touched(X(p), name); return ctx_push_global(X(p), name, host); }
int pomgpu_pull_global(pomgpu_t* p, const char* name, double* host) { apply_pending(X(p)); return ctx_pull_global(X(p), name, host); }
long pomgpu_field_elems(pomgpu_t* p, const char* name) {
  const FieldInfo* f = find_field(name);
  return f ? (long)field_elems(X(p), f) : 0;
}
int pomgpu_push_async(pomgpu_t* p, const char* name, const double* host) {
  Ctx* c = X(p);
  touched(c, name);
  const FieldInfo* f;
  double** slot = ctx_slot(c, name, &f);
  if (!slot || !*slot) return 2;
#ifdef POMGPU_EMU
  return dev_h2d(c, *slot, host, field_elems(c, f));
#else
  int ntab;
  const int id = (int)(f - field_table(&ntab));
  if (id < 0 || id >= 256) return 2;
  cudaSetDevice(c->device);
  if (!c->copy_stream) {
    cudaStream_t s; cudaEvent_t e;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return 1;
    c->copy_stream = (void*)s;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming); c->ev_copied = (void*)e;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming); c->ev_swapped = (void*)e;
    cudaEventRecord((cudaEvent_t)c->ev_swapped, (cudaStream_t)c->stream);
  }
  if (!c->shadow[id]) {
    // plain allocation: dev_alloc's zero fill runs on the COMPUTE stream and could land after the
    // copy below (which runs on the copy stream); the copy overwrites the whole buffer anyway
    if (cudaMalloc((void**)&c->shadow[id], field_elems(c, f) * sizeof(double)) != cudaSuccess) {
      snprintf(c->err, sizeof(c->err), "push_async(%s): out of device memory", name);
      c->c.error_status = 1;
      return 1;
    }
  }
  cudaStream_t cs = (cudaStream_t)c->copy_stream;
  cudaStreamWaitEvent(cs, (cudaEvent_t)c->ev_swapped, 0);   // the shadow was live until the last swap
  cudaError_t e = cudaMemcpyAsync(c->shadow[id], host, field_elems(c, f) * 8, cudaMemcpyHostToDevice, cs);
  if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "push_async(%s): %s", name, cudaGetErrorString(e)); c->c.error_status = 1; return 1; }
  cudaEventRecord((cudaEvent_t)c->ev_copied, cs);
  if (!c->pending[id]) { c->pending[id] = 1; c->npending++; }
  return 0;
#endif
}
int pomgpu_pin_host(void* ptr, unsigned long bytes) {
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
int pomgpu_step(pomgpu_t* p, int iint, double time, double ramp) { return step(self_group(X(p)), iint, time, ramp); }
int pomgpu_sync(pomgpu_t* p) { return dev_sync(X(p)); }
double pomgpu_check_velocity(pomgpu_t* p) { return check_velocity(X(p)); }
// enqueue the reduction for the step just enqueued and return the result of the PREVIOUS call
// (0 on the first): the host never waits for the step it has just launched
// max|a[0..n)| of the PREVIOUS call (0 on the first): the reduction of this call is only enqueued
static double absmax_lagged(Ctx* c, const double* a, size_t n, bool is_velocity) {
#ifdef POMGPU_EMU
  double m = 0.;
  for (size_t i = 0; i < n; ++i) { double v = fabs(a[i]); if (!(v <= m)) m = v; }
  if (is_velocity && !(m <= c->c.vmaxl)) c->c.error_status = 1;
  return m;
#else
  double prev = 0.;
  cudaSetDevice(c->device);
  cudaStream_t s = (cudaStream_t)c->stream;
  if (!c->ev_vel) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); c->ev_vel = (void*)e; }
  if (c->vel_lag) {
    cudaEventSynchronize((cudaEvent_t)c->ev_vel);
    prev = c->h_red[1];
    if (c->vel_lag == 1 && !(prev <= c->c.vmaxl)) c->c.error_status = 1;   // advance.f:631-638
  }
  cudaMemsetAsync(c->d_red + 1, 0, 8, s);
  int blocks = (int)((n + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
  c->launches++;
  absmax_kernel<<<blocks, 256, 0, s>>>(a, n, (unsigned long long*)(c->d_red + 1));
  cudaMemcpyAsync(c->h_red + 1, c->d_red + 1, 8, cudaMemcpyDeviceToHost, s);
  cudaEventRecord((cudaEvent_t)c->ev_vel, s);
  c->vel_lag = is_velocity ? 1 : 2;
  return prev;
#endif
}
double pomgpu_check_velocity_lagged(pomgpu_t* p) {
  Ctx* c = X(p);
#ifdef POMGPU_EMU
  return check_velocity(c);
#else
  const double* a = c->p.va + (size_t)(c->jown0 - 1 - c->g.joff) * c->g.im;
  return absmax_lagged(c, a, (size_t)(c->jown1 - c->jown0 + 1) * c->g.im, true);
#endif
}
// the same one-scalar-per-step read-back for any field (max |x| over the rows this strip holds):
// what a driver that runs single routines (the tracer-only bench) uses as its blow-up check
double pomgpu_field_absmax_lagged(pomgpu_t* p, const char* name) {
  Ctx* c = X(p);
  const FieldInfo* f;
  double** slot = ctx_slot(c, name, &f);
  if (!slot || !*slot) return -1.;
  return absmax_lagged(c, *slot, field_elems(c, f), false);
}
long pomgpu_selftest_pdiv(pomgpu_t* p, long n, unsigned long seed, int emax) { return selftest_pdiv(X(p), n, seed, emax); }
// ---- on-device time interpolation of forcing / boundary records (pom_forcing.cu) ----
int pomgpu_push_record(pomgpu_t* p, const char* name, int slot, const double* host) { return record_push(X(p), name, slot, host); }
int pomgpu_rotate_record(pomgpu_t* p, const char* name) { return record_rotate(X(p), name); }
int pomgpu_interp(pomgpu_t* p, const char* name, double fnew) { apply_pending(X(p)); return record_interp(X(p), &name, 1, fnew); }
int pomgpu_wind(pomgpu_t* p, double fnew) {   // bounds_forcing.f:904-909
  static const char* const n[] = {"wusurf", "wvsurf"};
  apply_pending(X(p));
  return record_interp(X(p), n, 2, fnew);
}
int pomgpu_heat(pomgpu_t* p, double fnew) {   // bounds_forcing.f:949-957
  static const char* const n[] = {"wtsurf", "swrad"};
  apply_pending(X(p));
  return record_interp(X(p), n, 2, fnew);
}
int pomgpu_lateral_bc(pomgpu_t* p, double fnew) { apply_pending(X(p)); return record_lateral_bc(X(p), fnew); }   // :841-865
int pomgpu_domain_stats_rows(pomgpu_t* p, double* rows) { apply_pending(X(p)); return domain_stats_rows(X(p), rows); }
long pomgpu_launch_count(pomgpu_t* p, int reset) {
  long n = X(p)->launches;
  if (reset) X(p)->launches = 0;
  return n;
}

// CUDA events on the library's launch stream (bench.py times the step loop with these); they
// belong to the context, i.e. to its device
int pomgpu_event_record(pomgpu_t* p, int slot) {
#ifdef POMGPU_EMU
  (void)p; (void)slot; return 0;
#else
  if (slot < 0 || slot >= 8) return 2;
  Ctx* c = X(p);
  cudaSetDevice(c->device);
  if (!c->ev[slot]) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return 1; c->ev[slot] = (void*)e; }
  return cudaEventRecord((cudaEvent_t)c->ev[slot], (cudaStream_t)c->stream) == cudaSuccess ? 0 : 1;
#endif
}
double pomgpu_event_elapsed_ms(pomgpu_t* p, int a, int b) {
#ifdef POMGPU_EMU
  (void)p; (void)a; (void)b; return 0.;
#else
  float ms = -1.f;
  Ctx* c = X(p);
  if (a < 0 || a >= 8 || b < 0 || b >= 8 || !c->ev[a] || !c->ev[b]) return -1.;
  cudaSetDevice(c->device);
  cudaEventSynchronize((cudaEvent_t)c->ev[b]);
  cudaEventElapsedTime(&ms, (cudaEvent_t)c->ev[a], (cudaEvent_t)c->ev[b]);
  return ms;
#endif
}
int pomgpu_profile_begin(pomgpu_t* p) { dev_sync(X(p)); X(p)->prof_on = 1; return 0; }
int pomgpu_profile_end(pomgpu_t* p, char* json, int len) { X(p)->prof_on = 0; return prof_report(X(p), json, len); }

// ---- strip groups (multi-GPU) ----------------------------------------------------------------
pomgpu_group_t* pomgpu_group_create(int n, pomgpu_t** ctxs) { return (pomgpu_group_t*)group_create(n, (Ctx**)ctxs); }
void pomgpu_group_destroy(pomgpu_group_t* g) { group_destroy(GG(g)); }
int pomgpu_nccl_unique_id(void* out128) { return nccl_unique_id(out128); }
int pomgpu_group_connect_nccl(pomgpu_group_t* g, const void* id128, int rank, int world) {
  return group_connect_nccl(GG(g), id128, rank, world);
}
void pomgpu_group_set_transport(pomgpu_group_t* g, pomgpu_halo_cb cb, void* user) { group_set_callback(GG(g), (halo_cb)cb, user); }
int pomgpu_group_step(pomgpu_group_t* g, int iint, double time, double ramp) { return step(GG(g), iint, time, ramp); }
double pomgpu_group_check_velocity(pomgpu_group_t* g) {
  double m = 0.;
  for (int r = 0; r < GG(g)->n; ++r) { double v = check_velocity(GG(g)->c[r]); if (!(v <= m)) m = v; }
  return m;
}
int pomgpu_group_halo_trace(pomgpu_group_t* g, char* buf, int len) { return group_trace_report(GG(g), buf, len); }
int pomgpu_group_transport(pomgpu_group_t* g) {
  Group* G = GG(g);
  if (G->cb) return 3;
  if (G->ipc) return 2;
  return G->nccl ? 1 : 0;
}
long pomgpu_group_exchanges(pomgpu_group_t* g, long* fields, int reset) {
  long n = GG(g)->n_exchanges;
  if (fields) *fields = GG(g)->n_fields_exchanged;
  if (reset) { GG(g)->n_exchanges = 0; GG(g)->n_fields_exchanged = 0; }
  return n;
}

// dens / baropg on a group: what the Fortran `initialize` calls before the first step
// (initialize.f:416,425,502)
static int fid(const char* name);
int pomgpu_group_dens(pomgpu_group_t* g, const char* si, const char* ti, const char* rhoo) {
  int a = fid(si), b = fid(ti), o = fid(rhoo);
  if (a < 0 || b < 0 || o < 0) return 2;
  apply_pending(GG(g));
  k_dens(GG(g), a, b, o);
  return 0;
}
int pomgpu_group_baropg(pomgpu_group_t* g) { apply_pending(GG(g)); k_baropg(GG(g), GG(g)->c[0]->c.npg); return 0; }
// the four step routines of advance.f:21-32 one by one on a group (a driver that keeps the reference's own
// `advance`, e.g. libpomgpu_f over several strips); the constants are those of strip 0
int pomgpu_group_lateral_viscosity(pomgpu_group_t* g) {
  Group* G = GG(g); apply_pending(G); CSYNC();
  if (int r = check_switches(G)) return r;
  return lateral_viscosity(G);
}
int pomgpu_group_mode_interaction(pomgpu_group_t* g) { Group* G = GG(g); apply_pending(G); CSYNC(); return mode_interaction(G); }
int pomgpu_group_mode_external(pomgpu_group_t* g, int iext) {
  Group* G = GG(g); apply_pending(G); CSYNC();
  if (int r = check_switches(G)) return r;
  return mode_external(G, iext);
}
int pomgpu_group_mode_internal(pomgpu_group_t* g, int iint) {
  Group* G = GG(g); apply_pending(G); CSYNC();
  if (int r = check_switches(G)) return r;
  return mode_internal(G, iint);
}
int pomgpu_group_error_status(pomgpu_group_t* g) { return group_status(GG(g)); }
int pomgpu_group_baropg_kind(pomgpu_group_t* g, int npg) { Group* G = GG(g); apply_pending(G); CSYNC(); k_baropg(G, npg == 2 ? 2 : 1); return 0; }

// ---- the reference's subroutines on the resident state ----------------------------------------
#define SG Group* G = self_group(X(p)); apply_pending(G)
int pomgpu_lateral_viscosity(pomgpu_t* p) { SG; if (int r = check_switches(G)) return r; return lateral_viscosity(G); }
int pomgpu_mode_interaction(pomgpu_t* p) { SG; return mode_interaction(G); }
int pomgpu_mode_external(pomgpu_t* p, int iext) { SG; if (int r = check_switches(G)) return r; return mode_external(G, iext); }
int pomgpu_internal_stage(pomgpu_t* p, int iint, int stage) { SG; return internal_stage(G, iint, stage); }
int pomgpu_mode_internal(pomgpu_t* p, int iint) { SG; if (int r = check_switches(G)) return r; return mode_internal(G, iint); }
int pomgpu_advave(pomgpu_t* p) { SG; k_advave(G); return 0; }
int pomgpu_advct(pomgpu_t* p) { SG; k_advct(G); return 0; }
int pomgpu_advq(pomgpu_t* p) { SG; k_advq(G); return 0; }
int pomgpu_advu(pomgpu_t* p) { SG; k_advu(G); return 0; }
int pomgpu_advv(pomgpu_t* p) { SG; k_advv(G); return 0; }
int pomgpu_baropg(pomgpu_t* p) { SG; k_baropg(G, 1); return 0; }
int pomgpu_baropg_mcc(pomgpu_t* p) { SG; k_baropg(G, 2); return 0; }
int pomgpu_profq(pomgpu_t* p) { SG; k_profq(G, 0); return 0; }
int pomgpu_profu(pomgpu_t* p) { SG; k_profu(G); return 0; }
int pomgpu_profv(pomgpu_t* p) { SG; k_profv(G); return 0; }
int pomgpu_vertvl(pomgpu_t* p) { SG; k_vertvl(G); return 0; }
int pomgpu_realvertvl(pomgpu_t* p) { SG; k_realvertvl(G); return 0; }

// ---- bcond / bcondorl / smol_adif / single-quantity advq as stand-alone entries (unit mode) ----
static int k_bcond(Group* G, int idx, int orl) {
  const int e = 0;
  int rc = 0;
  EACH(rc |= run_bcond(c, idx, orl, j0, j1));
  return rc;
}
int pomgpu_bcond(pomgpu_t* p, int idx) { SG; return k_bcond(G, idx, 0); }
int pomgpu_bcondorl(pomgpu_t* p, int idx) { SG; return k_bcond(G, idx, 1); }

static int fid(const char* name) {
  int n;
  const FieldInfo* t = field_table(&n);
  for (int i = 0; i < n; ++i)
    if (!strcmp(t[i].name, name)) return i;
  return -1;
}
static int advt(pomgpu_t* p, int nadv, const char* fb, const char* f, const char* fclim, const char* ff) {
  SG;
  Ctx* c = X(p);
  int a = fid(fb), b = fid(f), cl = fid(fclim), o = fid(ff);
  if (a < 0 || b < 0 || cl < 0 || o < 0) return 2;
  const int keep = c->c.nadv;
  c->c.nadv = nadv;
  k_advt(G, a, b, cl, o, o);
  c->c.nadv = keep;
  // side effects the reference leaves on fb (and f for advt1): solver.f:496,511,532 / 618,691,715
  run_fb_roundtrip(c, FP(c, a), FP(c, cl), nadv == 1 ? FP(c, b) : nullptr, 1, c->g.jmg);
  return 0;
}
int pomgpu_advt1(pomgpu_t* p, const char* fb, const char* f, const char* fclim, const char* ff) { return advt(p, 1, fb, f, fclim, ff); }
int pomgpu_advt2(pomgpu_t* p, const char* fb, const char* f, const char* fclim, const char* ff) { return advt(p, 2, fb, f, fclim, ff); }
int pomgpu_dens(pomgpu_t* p, const char* si, const char* ti, const char* rhoo) {
  SG;
  int a = fid(si), b = fid(ti), o = fid(rhoo);
  if (a < 0 || b < 0 || o < 0) return 2;
  k_dens(G, a, b, o);
  return 0;
}
// smol_adif(xmassflux,ymassflux,zwflux,ff) (solver.f:1880-1967): ff*fsm, then the anti-diffusive
// mass fluxes in place
int pomgpu_smol_adif(pomgpu_t* p, const char* xm, const char* ym, const char* zw, const char* ff) {
  SG;
  Ctx* c = X(p);
  int a = fid(xm), b = fid(ym), w = fid(zw), o = fid(ff);
  if (a < 0 || b < 0 || w < 0 || o < 0) return 2;
  run_mask_fsm(c, FP(c, o), 1, c->g.jmg);
  run_smol_adif(c, FP(c, o), FP(c, a), FP(c, b), FP(c, w), 0, 1, c->g.jmg);
  return 0;
}
// advq(qb,q,qf) (solver.f:411-477) for ONE quantity: the device kernel advances q2 and q2l together
// (pomgpu_advq), so the same quantity is bound to both of its slots and the second result discarded
int pomgpu_advq_fields(pomgpu_t* p, const char* qb, const char* q, const char* qf) {
  SG;
  Ctx* c = X(p);
  int a = fid(qb), b = fid(q), o = fid(qf);
  if (a < 0 || b < 0 || o < 0) return 2;
  const Ptrs keep = c->p;
  c->p.q2b = FP(c, a); c->p.q2lb = FP(c, a); c->p.q2 = FP(c, b); c->p.q2l = FP(c, b);
  c->p.uf = FP(c, o); c->p.vf = keep.s3e;
  run_advq(c, 1, c->g.jmg);
  c->p = keep;
  return 0;
}
int pomgpu_proft(pomgpu_t* p, const char* f, const char* wfsurf, const char* fsurf, int nbc) {
  SG;
  int a = fid(f), b = fid(wfsurf), s = fid(fsurf);
  if (a < 0 || b < 0 || s < 0) return 2;
  k_proft(G, a, b, s, nbc);
  return 0;
}

}  // extern "C"
