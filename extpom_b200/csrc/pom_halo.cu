// pom_halo.cu -- j-strip domain decomposition: ghost-row validity tracking and halo
// exchange.  Replaces distribute_mpi / exchange2d_mpi / exchange3d_mpi
// (pom/parallel_mpi.f:34-122,154-351).
//
// Memory is i-contiguous, so the domain is cut in j only; each strip holds `ghost` extra
// rows on its interior seams.  Every kernel computes with global-index semantics, so a
// cell gets bit-identical arithmetic on whichever strip computes it, and instead of the
// reference's 29 3-D + 340 2-D single-row exchanges per step the strips
// recompute the rows next to a seam redundantly: each field carries the number of ghost
// rows that are still valid, every kernel consumes `radius` rows of validity of its inputs,
// and a batched exchange of the depleted fields is issued only when a kernel would
// otherwise read a stale row.  An N-strip run is therefore bitwise equal to the 1-strip run.
//
// Transports: (a) strips held by one process (tests; one GPU or peers) -- device copies;
// (b) one process per GPU -- NCCL send/recv over NVLink, libnccl resolved with dlopen so the
// library links without it; (c) a host callback (gloo in the CPU tests).
#include "pom_halo.h"
#include <cstdlib>
#ifndef POMGPU_EMU
#include <dlfcn.h>
#endif

namespace pom {

// ---- pack / unpack ----------------------------------------------------------------
struct PackJob {
  double* f[HALO_MAXF];
  int nk[HALO_MAXF];   // levels of field n (1 for 2-D)
  int nf;
  int im, n2, rows, row0;   // rows to move, first local row (0-based)
  long total;               // sum_n nk[n]*rows*im
};

POM_HD void pack_one(const PackJob& J, double* buf, long e, bool unpack) {
  // element e of the packed buffer -> (field n, level k, row r, column i)
  long per = (long)J.rows * J.im;
  int n = 0;
  long base = 0;
  while (n < J.nf - 1 && e >= base + per * J.nk[n]) { base += per * J.nk[n]; ++n; }
  long q = e - base;
  int k = (int)(q / per);
  long rr = q - (long)k * per;   // r*im + i, rows are contiguous in memory
  double* a = J.f[n] + (long)k * J.n2 + (long)J.row0 * J.im + rr;
  if (unpack) *a = buf[e]; else buf[e] = *a;
}

#ifndef POMGPU_EMU
__global__ void pack_kernel(PackJob J, double* buf, bool unpack) {
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < J.total; e += (long)gridDim.x * blockDim.x)
    pack_one(J, buf, e, unpack);
}
#endif

static void run_pack(Ctx* c, const PackJob& J, double* buf, bool unpack, void* stream) {
  if (J.total == 0) return;
  c->launches++;
#ifdef POMGPU_EMU
  (void)stream;
  for (long e = 0; e < J.total; ++e) pack_one(J, buf, e, unpack);
#else
  cudaSetDevice(c->device);
  int blocks = (int)((J.total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(J, buf, unpack);
#endif
}

#ifndef POMGPU_EMU
// the communication stream of a strip (high priority: its few blocks are scheduled ahead of the
// compute kernels' when SM slots free up) and the two events that order it against the compute stream
static bool comm_ready(Ctx* c) {
  if (c->comm_stream) return true;
  cudaSetDevice(c->device);
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  cudaStream_t s; cudaEvent_t e;
  if (cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi) != cudaSuccess) return false;
  c->comm_stream = (void*)s;
  cudaEventCreateWithFlags(&e, cudaEventDisableTiming); c->ev_packed = (void*)e;
  cudaEventCreateWithFlags(&e, cudaEventDisableTiming); c->ev_halo = (void*)e;
  return true;
}
#endif

// ---- NCCL through dlopen -----------------------------------------------------------------
#ifndef POMGPU_EMU
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(void*);
  int (*CommInitRank)(void**, int, NcclId, int);
  int (*CommDestroy)(void*);
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t);
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char* (*GetErrorString)(int);
};
static NcclApi g_nccl;
static int nccl_load() {
  if (g_nccl.h) return 0;
  // RTLD_NOLOAD first: reuse the libnccl the host process (e.g. torch) already mapped
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return 1;
#define SYM(n) *(void**)(&g_nccl.n) = dlsym(h, "nccl" #n); if (!g_nccl.n) return 1;
  SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(Send) SYM(Recv) SYM(GroupStart) SYM(GroupEnd)
  SYM(GetErrorString)
#undef SYM
  g_nccl.h = h;
  return 0;
}
#endif

int nccl_unique_id(void* out128) {
#ifdef POMGPU_EMU
  memset(out128, 0, 128); return 1;
#else
  if (nccl_load()) return 1;
  return g_nccl.GetUniqueId(out128) == 0 ? 0 : 1;
#endif
}

// ---- group --------------------------------------------------------------------------------
Group* group_create(int n, Ctx** ctxs) {
  if (n < 1 || n > 16) return nullptr;
  Group* G = (Group*)calloc(1, sizeof(Group));
  G->n = n;
  G->ghost = ctxs[0]->ghost;
  for (int r = 0; r < n; ++r) {
    G->c[r] = ctxs[r];
    if (ctxs[r]->ghost != G->ghost) { free(G); return nullptr; }
    if (r > 0 && ctxs[r]->jown0 != ctxs[r - 1]->jown1 + 1) { free(G); return nullptr; }
#ifndef POMGPU_EMU
    // strips of one process run in program order on one stream (same device in the tests)
    if (r > 0 && ctxs[r]->device == ctxs[0]->device) ctxs[r]->stream = ctxs[0]->stream;
#endif
  }
  G->seams = false;
  for (int r = 0; r < n; ++r) {
    Ctx* c = G->c[r];
    if (c->jown0 > 1 || c->jown1 < c->g.jmg) G->seams = true;
    // a strip must be deep enough that the open-boundary kernels (rows 1..4, jm-3..jm) and the
    // ghost rows of the neighbours never overlap
    if ((c->jown0 > 1 || c->jown1 < c->g.jmg) && c->jown1 - c->jown0 + 1 < G->ghost + 4) { free(G); return nullptr; }
  }
  for (int f = 0; f < F_COUNT; ++f) G->valid[f] = G->ghost;
  // overlap needs one compute stream per process side of a seam: strips of one device (shared stream)
  // or a single strip per process; in-process seams between devices stay on the ordered path
  G->overlap = (getenv("POMGPU_NO_OVERLAP") == nullptr);
#ifndef POMGPU_EMU
  for (int r = 1; r < n; ++r) if (ctxs[r]->device != ctxs[0]->device) G->overlap = 0;
#endif
  return G;
}

void group_destroy(Group* G) {
  if (!G) return;
  for (int r = 0; r < G->n; ++r) {
    dev_sync(G->c[r]);
    for (int b = 0; b < 4; ++b)
      if (G->buf[r][b]) dev_free(G->c[r], G->buf[r][b]);
    G->c[r]->stream = G->c[r]->own_stream;
  }
#ifndef POMGPU_EMU
  if (G->nccl) g_nccl.CommDestroy(G->nccl);
  for (int b = 0; b < 4; ++b) if (G->hbuf[b]) cudaFreeHost(G->hbuf[b]);
  for (int r = 0; r < 16; ++r) {
    if (G->ev_pack[r]) cudaEventDestroy((cudaEvent_t)G->ev_pack[r]);
    if (G->ev_copy[r]) cudaEventDestroy((cudaEvent_t)G->ev_copy[r]);
  }
#endif
  free(G);
}

int group_connect_nccl(Group* G, const void* id128, int rank, int world) {
#ifdef POMGPU_EMU
  (void)G; (void)id128; (void)rank; (void)world; return 1;
#else
  if (nccl_load()) { snprintf(G->c[0]->err, 256, "libnccl.so.2 not found"); return 1; }
  NcclId id;
  memcpy(&id, id128, 128);
  cudaSetDevice(G->c[0]->device);
  int rc = g_nccl.CommInitRank(&G->nccl, world, id, rank);
  if (rc) { snprintf(G->c[0]->err, 256, "ncclCommInitRank: %s", g_nccl.GetErrorString(rc)); return 1; }
  G->rank = rank; G->world = world;
  return 0;
#endif
}

void group_set_callback(Group* G, halo_cb cb, void* user) { G->cb = cb; G->cb_user = user; }

static bool has_s(const Ctx* c) { return c->jown0 > 1; }
static bool has_n(const Ctx* c) { return c->jown1 < c->g.jmg; }

// a transport failure: the ghost rows are stale from here on -- mark the group (nothing is
// launched any more, pom_step.cu) and every strip (the reference's error convention)
static int fail(Group* G, const char* msg) {
  G->failed = 1;
  for (int r = 0; r < G->n; ++r) {
    snprintf(G->c[r]->err, sizeof(G->c[r]->err), "%s", msg);
    G->c[r]->c.error_status = 1;
  }
  fprintf(stderr, "pomgpu: %s\n", msg);
  return 1;
}

// copy between the staging buffers of two strips of this process, on the DESTINATION's stream
static int dev_copy_between(Ctx* dc, double* dst, Ctx* sc, const double* src, size_t n, void* stream) {
#ifdef POMGPU_EMU
  (void)sc; (void)stream; return dev_d2d(dc, dst, src, n);
#else
  if (dc->device == sc->device) {
    cudaSetDevice(dc->device);
    if (cudaMemcpyAsync(dst, src, n * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) == cudaSuccess) return 0;
    snprintf(dc->err, sizeof(dc->err), "halo copy failed");
    dc->c.error_status = 1;
    return 1;
  }
  cudaSetDevice(dc->device);
  if (cudaMemcpyPeerAsync(dst, dc->device, src, sc->device, n * 8, (cudaStream_t)stream) != cudaSuccess) {
    snprintf(dc->err, sizeof(dc->err), "peer copy between devices %d and %d failed", sc->device, dc->device);
    dc->c.error_status = 1;
    return 1;
  }
  return 0;
#endif
}

// Exchange `ghost` rows of the listed fields across every seam; afterwards they are valid to
// full depth.  One pack kernel, one transfer and one unpack kernel per direction and strip,
// whatever the number of fields.
int group_exchange(Group* G, const int* fields, int nf) {
  if (!G->seams || nf == 0) return 0;
  const FieldInfo* tab;
  int ntab;
  tab = field_table(&ntab);
  const int gh = G->ghost;
  G->n_exchanges++;
  G->n_fields_exchanged += nf;
  for (int done = 0; done < nf; done += HALO_MAXF) {
    const int m = (nf - done < HALO_MAXF) ? nf - done : HALO_MAXF;
    PackJob J[16][4];   // per strip: 0 send south, 1 send north, 2 recv south, 3 recv north
    for (int r = 0; r < G->n; ++r) {
      Ctx* c = G->c[r];
      PackJob P;
      P.nf = m; P.im = c->g.im; P.n2 = c->g.n2; P.rows = gh; P.total = 0;
      for (int n = 0; n < m; ++n) {
        const FieldInfo& fi = tab[fields[done + n]];
        P.f[n] = *(double**)((char*)&c->p + fi.offset);
        P.nk[n] = (fi.kind == K3D) ? c->g.kb : 1;
        P.total += (long)P.nk[n] * gh * c->g.im;
      }
      const int l0 = c->jown0 - 1 - c->g.joff, l1 = c->jown1 - 1 - c->g.joff;   // owned local rows
      J[r][0] = P; J[r][0].row0 = l0;            // my first owned rows -> south neighbour's north ghosts
      J[r][1] = P; J[r][1].row0 = l1 - gh + 1;   // my last owned rows  -> north neighbour's south ghosts
      J[r][2] = P; J[r][2].row0 = l0 - gh;       // my south ghost rows
      J[r][3] = P; J[r][3].row0 = l1 + 1;        // my north ghost rows
      for (int b = 0; b < 4; ++b) {
        if ((size_t)P.total > G->bufcap[r][b]) {
          if (G->buf[r][b]) { dev_sync(c); dev_free(c, G->buf[r][b]); }
          G->bufcap[r][b] = (size_t)P.total * 2;
          if (dev_alloc(c, &G->buf[r][b], G->bufcap[r][b])) return fail(G, "halo exchange: out of device memory");
        }
      }
      if (has_s(c)) run_pack(c, J[r][0], G->buf[r][0], false, c->stream);
      if (has_n(c)) run_pack(c, J[r][1], G->buf[r][1], false, c->stream);
    }
    // from here on (transfer, unpack) on the communication stream when the transport can overlap:
    // xs[r] = the stream the rest of strip r's exchange is enqueued on
    void* xs_[16];
    bool ov = false;
#ifndef POMGPU_EMU
    ov = G->overlap && !G->cb;
    for (int r = 0; r < G->n && ov; ++r) ov = comm_ready(G->c[r]);
#endif
    for (int r = 0; r < G->n; ++r) {
      Ctx* c = G->c[r];
      xs_[r] = c->stream;
#ifndef POMGPU_EMU
      if (ov) {
        // strips of one device share the compute stream and (strip 0's) communication stream
        Ctx* c0 = G->c[0];
        xs_[r] = c0->comm_stream;
        if (r == 0) {
          cudaSetDevice(c0->device);
          cudaEventRecord((cudaEvent_t)c0->ev_packed, (cudaStream_t)c0->stream);
          cudaStreamWaitEvent((cudaStream_t)c0->comm_stream, (cudaEvent_t)c0->ev_packed, 0);
        }
      }
#endif
    }
    // seams inside this process.  Strips on one device share a stream (program order); strips on
    // different devices have their own streams, so the copy into b's receive buffer must wait for
    // a's pack (and vice versa), and nobody may re-pack a send buffer the other side still reads.
    for (int r = 0; r + 1 < G->n; ++r) {
      Ctx *a = G->c[r], *b = G->c[r + 1];
#ifndef POMGPU_EMU
      const bool cross = (a->stream != b->stream);
      if (cross) {
        for (int q = r; q <= r + 1; ++q)
          if (!G->ev_pack[q]) {
            cudaEvent_t e;
            cudaSetDevice(G->c[q]->device);
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming); G->ev_pack[q] = (void*)e;
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming); G->ev_copy[q] = (void*)e;
          }
        cudaSetDevice(a->device); cudaEventRecord((cudaEvent_t)G->ev_pack[r], (cudaStream_t)a->stream);
        cudaSetDevice(b->device); cudaEventRecord((cudaEvent_t)G->ev_pack[r + 1], (cudaStream_t)b->stream);
        cudaStreamWaitEvent((cudaStream_t)b->stream, (cudaEvent_t)G->ev_pack[r], 0);
        cudaSetDevice(a->device); cudaStreamWaitEvent((cudaStream_t)a->stream, (cudaEvent_t)G->ev_pack[r + 1], 0);
      }
#endif
      int rc = dev_copy_between(b, G->buf[r + 1][2], a, G->buf[r][1], (size_t)J[r][1].total, xs_[r + 1]);   // a's north rows -> b's south ghosts
      rc |= dev_copy_between(a, G->buf[r][3], b, G->buf[r + 1][0], (size_t)J[r + 1][0].total, xs_[r]);
      if (rc) return 1;
#ifndef POMGPU_EMU
      if (cross) {   // the next pack into a send buffer waits until the neighbour's copy has read it
        cudaSetDevice(b->device); cudaEventRecord((cudaEvent_t)G->ev_copy[r + 1], (cudaStream_t)b->stream);
        cudaSetDevice(a->device); cudaEventRecord((cudaEvent_t)G->ev_copy[r], (cudaStream_t)a->stream);
        cudaStreamWaitEvent((cudaStream_t)a->stream, (cudaEvent_t)G->ev_copy[r + 1], 0);
        cudaSetDevice(b->device); cudaStreamWaitEvent((cudaStream_t)b->stream, (cudaEvent_t)G->ev_copy[r], 0);
      }
#endif
    }
    // seams to other processes
    Ctx* cs = G->c[0];
    Ctx* cn = G->c[G->n - 1];
    const bool xs = has_s(cs), xn = has_n(cn);
    if (xs || xn) {
      if (G->cb) {
        const long ns = xs ? J[0][0].total : 0, nn = xn ? J[G->n - 1][1].total : 0;
        double *ss = xs ? G->buf[0][0] : nullptr, *rs = xs ? G->buf[0][2] : nullptr;
        double *sn = xn ? G->buf[G->n - 1][1] : nullptr, *rn = xn ? G->buf[G->n - 1][3] : nullptr;
#ifndef POMGPU_EMU
        // the callback works on host memory: mirror the device staging buffers in pinned memory
        const long want[4] = {ns, nn, ns, nn};
        for (int b = 0; b < 4; ++b)
          if ((size_t)want[b] > G->hbufcap[b]) {
            if (G->hbuf[b]) cudaFreeHost(G->hbuf[b]);
            G->hbufcap[b] = (size_t)want[b] * 2;
            if (cudaMallocHost((void**)&G->hbuf[b], G->hbufcap[b] * sizeof(double)) != cudaSuccess) { G->hbuf[b] = nullptr; G->hbufcap[b] = 0; return fail(G, "halo transport: out of pinned host memory"); }
          }
        if (xs && dev_d2h(cs, G->hbuf[0], ss, (size_t)ns)) return fail(G, "halo transport: D2H");
        if (xn && dev_d2h(cn, G->hbuf[1], sn, (size_t)nn)) return fail(G, "halo transport: D2H");
        ss = xs ? G->hbuf[0] : nullptr; sn = xn ? G->hbuf[1] : nullptr;
        double *drs = rs, *drn = rn;
        rs = xs ? G->hbuf[2] : nullptr; rn = xn ? G->hbuf[3] : nullptr;
#else
        dev_sync(cs); if (cn != cs) dev_sync(cn);
#endif
        if (G->cb(G->cb_user, ss, rs, ns, sn, rn, nn)) return fail(G, "halo transport callback failed");
#ifndef POMGPU_EMU
        if (xs && dev_h2d(cs, drs, rs, (size_t)ns)) return fail(G, "halo transport: H2D");
        if (xn && dev_h2d(cn, drn, rn, (size_t)nn)) return fail(G, "halo transport: H2D");
#endif
      }
#ifndef POMGPU_EMU
      else if (G->nccl) {
        const int ncclDouble = 8;
        cudaSetDevice(cs->device);
        g_nccl.GroupStart();
        if (xs) {
          g_nccl.Send(G->buf[0][0], (size_t)J[0][0].total, ncclDouble, G->rank - 1, G->nccl, (cudaStream_t)xs_[0]);
          g_nccl.Recv(G->buf[0][2], (size_t)J[0][0].total, ncclDouble, G->rank - 1, G->nccl, (cudaStream_t)xs_[0]);
        }
        if (xn) {
          g_nccl.Send(G->buf[G->n - 1][1], (size_t)J[G->n - 1][1].total, ncclDouble, G->rank + 1, G->nccl, (cudaStream_t)xs_[G->n - 1]);
          g_nccl.Recv(G->buf[G->n - 1][3], (size_t)J[G->n - 1][1].total, ncclDouble, G->rank + 1, G->nccl, (cudaStream_t)xs_[G->n - 1]);
        }
        int rc = g_nccl.GroupEnd();
        if (rc) { char m[200]; snprintf(m, sizeof(m), "nccl exchange: %s", g_nccl.GetErrorString(rc)); return fail(G, m); }
      }
#endif
      else return fail(G, "strip has a seam but no transport is connected");
    }
    for (int r = 0; r < G->n; ++r) {
      Ctx* c = G->c[r];
      if (has_s(c)) run_pack(c, J[r][2], G->buf[r][2], true, xs_[r]);
      if (has_n(c)) run_pack(c, J[r][3], G->buf[r][3], true, xs_[r]);
    }
#ifndef POMGPU_EMU
    if (ov) {   // the kernel that asked for these rows waits for this event before it touches a seam band
      Ctx* c0 = G->c[0];
      cudaSetDevice(c0->device);
      cudaEventRecord((cudaEvent_t)c0->ev_halo, (cudaStream_t)c0->comm_stream);
      G->ov_active = 1;
    }
#else
    if (G->overlap && !G->cb) G->ov_active = 1;   // the emulation runs everything in order, but splits the windows the same way
#endif
  }
  for (int n = 0; n < nf; ++n) G->valid[fields[n]] = gh;
  return 0;
}

// A kernel is about to read fields `in` (id, j-radius).  Exchange what is too stale, then
// return how many ghost rows the kernel can also compute (its window extends that far).
int group_need(Group* G, const Req* in, int n) {
  if (!G->seams) return 0;
  for (int q = 0; q < n; ++q)
    if (in[q].r > G->ov_r) G->ov_r = in[q].r;
  int stale[64], ns = 0;
  bool must = false;
  for (int q = 0; q < n; ++q)
    if (G->valid[in[q].f] < in[q].r) must = true;
  if (must) {
    // A message costs latency, not bandwidth (a 3-D halo is ~1 MB), so batch: everything this
    // kernel reads that is not at full depth, plus every live 2-D field, plus -- when a 3-D
    // field triggered the exchange -- every live 3-D field.  Owned rows are always valid, so
    // refreshing more ghosts than strictly needed is always correct.
    static const int live2d[] = {F_ua, F_va, F_uab, F_vab, F_el, F_elb, F_d, F_dt, F_et, F_etb, F_etf, F_egf,
                                 F_egb, F_utf, F_vtf, F_utb, F_vtb, F_wubot, F_wvbot, F_aam2d, F_adx2d,
                                 F_ady2d, F_drx2d, F_dry2d, F_advua, F_advva, F_vfluxb};
    static const int live3d[] = {F_u, F_v, F_ub, F_vb, F_t, F_s, F_tb, F_sb, F_q2, F_q2b, F_q2l, F_q2lb, F_w,
                                 F_aam, F_km, F_kh, F_kq, F_rho, F_advx, F_advy, F_drhox, F_drhoy};
    int ntab;
    const FieldInfo* tab = field_table(&ntab);
    bool any3d = false;
    auto add = [&](int f) {
      for (int s = 0; s < ns; ++s) if (stale[s] == f) return;
      if (G->valid[f] < G->ghost && ns < 64) stale[ns++] = f;
    };
    for (int q = 0; q < n; ++q) {
      if (G->valid[in[q].f] < in[q].r && tab[in[q].f].kind == K3D) any3d = true;
      add(in[q].f);
    }
    for (size_t q = 0; q < sizeof(live2d) / sizeof(int); ++q) add(live2d[q]);
    if (any3d) for (size_t q = 0; q < sizeof(live3d) / sizeof(int); ++q) add(live3d[q]);
    if (G->ov_active) group_wait_halo(G);         // (a second exchange before the launch: keep the streams in order)
    if (group_exchange(G, stale, ns)) return 0;   // G->failed is set: the caller launches nothing
  }
  int e = G->ghost;
  for (int q = 0; q < n; ++q) {
    int v = G->valid[in[q].f] - in[q].r;
    if (v < e) e = v;
  }
  return e < 0 ? 0 : e;
}

// ---- launch windows of the kernel that follows a group_need (pom_step.cu: EACH) ----------------
// No exchange in flight: one part, the owned rows plus e ghost rows.  Exchange in flight: part 0 =
// the rows that read no ghost row (everything at least ov_r rows away from a seam), then -- after
// ev_halo -- part 1 = the band at the south seam, part 2 = the band at the north seam.  A cell gets
// the same arithmetic whichever launch computes it, so the split does not change a bit.
int group_parts(Group* G) { return (G->seams && G->ov_active) ? 3 : 1; }
bool group_window(Group* G, const Ctx* c, int e, int part, int nparts, int* j0, int* j1) {
  const int lo = c->jown0 > 1 ? c->jown0 - e : 1, hi = c->jown1 < c->g.jmg ? c->jown1 + e : c->g.jmg;
  if (nparts == 1) { *j0 = lo; *j1 = hi; return true; }
  const int a0 = has_s(c) ? c->jown0 + G->ov_r : lo, a1 = has_n(c) ? c->jown1 - G->ov_r : hi;
  if (part == 0) { *j0 = a0; *j1 = a1; return a1 >= a0; }
  if (part == 1) { *j0 = lo; *j1 = a0 - 1; return has_s(c) && a0 - 1 >= lo; }
  *j0 = a1 + 1; *j1 = hi;
  return has_n(c) && hi >= a1 + 1;
}
void group_wait_halo(Group* G) {
#ifndef POMGPU_EMU
  if (G->ov_active && G->c[0]->ev_halo)
    for (int r = 0; r < G->n; ++r) {
      if (r > 0 && G->c[r]->stream == G->c[0]->stream) continue;
      cudaSetDevice(G->c[r]->device);
      cudaStreamWaitEvent((cudaStream_t)G->c[r]->stream, (cudaEvent_t)G->c[0]->ev_halo, 0);
    }
#endif
  G->ov_active = 0;
}
void group_launched(Group* G) {
  if (G->ov_active) group_wait_halo(G);   // (nothing was launched on the bands: still order the streams)
  G->ov_r = 0;
}

void group_produced(Group* G, int e, const int* out, int n) {
  if (!G->seams) return;
  for (int q = 0; q < n; ++q) G->valid[out[q]] = e;
}

void group_swap(Group* G, int fa, int fb) {
  int ntab;
  const FieldInfo* tab = field_table(&ntab);
  for (int r = 0; r < G->n; ++r) {
    double** a = (double**)((char*)&G->c[r]->p + tab[fa].offset);
    double** b = (double**)((char*)&G->c[r]->p + tab[fb].offset);
    double* t = *a; *a = *b; *b = t;
  }
  int v = G->valid[fa]; G->valid[fa] = G->valid[fb]; G->valid[fb] = v;
}

}  // namespace pom
