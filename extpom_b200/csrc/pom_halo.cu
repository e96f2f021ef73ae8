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
#include <chrono>
#ifndef POMGPU_EMU
#include <cuda.h>
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

static bool has_s(const Ctx* c) { return c->jown0 > 1; }
static bool has_n(const Ctx* c) { return c->jown1 < c->g.jmg; }

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
  G->trace = (getenv("POMGPU_HALO_TRACE") != nullptr);
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
  if (G->ipc) {
    for (int q = 0; q < 2; ++q) {
      if (G->peer_rbuf[q]) cudaIpcCloseMemHandle(G->peer_rbuf[q]);
      if (G->peer_flags[q]) cudaIpcCloseMemHandle(G->peer_flags[q]);
    }
  }
  if (G->flags) cudaFree(G->flags);
  if (G->nccl) g_nccl.CommDestroy(G->nccl);
  for (int b = 0; b < 4; ++b) if (G->hbuf[b]) cudaFreeHost(G->hbuf[b]);
  for (int r = 0; r < 16; ++r) {
    if (G->ev_pack[r]) cudaEventDestroy((cudaEvent_t)G->ev_pack[r]);
    if (G->ev_copy[r]) cudaEventDestroy((cudaEvent_t)G->ev_copy[r]);
  }
#endif
  free(G);
}

#ifndef POMGPU_EMU
// ---- direct peer transport over CUDA IPC ---------------------------------------------------
// stream memory operations (driver API, resolved through the runtime like cuTensorMapEncodeTiled)
typedef CUresult (*StreamWait32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*StreamWrite32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static StreamWait32Fn p_wait32 = nullptr;
static StreamWrite32Fn p_write32 = nullptr;
static bool memops_load() {
  if (p_wait32 && p_write32) return true;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
  p_wait32 = (StreamWait32Fn)p;
  if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
  p_write32 = (StreamWrite32Fn)p;
  return true;
}

// Map the neighbours' receive buffers and flag words.  The staging buffers get their final size
// here (every live field at `ghost` rows), because an IPC handle names one allocation.  The handles
// travel through the NCCL communicator that was just created.  Returns 0 on success.
static int ipc_connect(Group* G) {
  if (G->n != 1 || !G->nccl || !memops_load()) return 1;
  Ctx* c = G->c[0];
  cudaSetDevice(c->device);
  const bool xs = has_s(c), xn = has_n(c);
  if (!xs && !xn) return 1;
  const size_t cap = (size_t)(32 * c->g.kb + 48) * G->ghost * c->g.im;
  for (int b = 0; b < 4; ++b) {
    if (G->buf[0][b]) dev_free(c, G->buf[0][b]);
    if (dev_alloc(c, &G->buf[0][b], cap)) return 1;
    G->bufcap[0][b] = cap;
  }
  if (cudaMalloc((void**)&G->flags, 64) != cudaSuccess) return 1;
  cudaMemset(G->flags, 0, 64);
  // what a neighbour needs from me: my receive buffers (2: from south, from north) and my flags
  struct Pack { cudaIpcMemHandle_t rs, rn, fl; } mine, south, north;
  if (cudaIpcGetMemHandle(&mine.rs, G->buf[0][2]) != cudaSuccess || cudaIpcGetMemHandle(&mine.rn, G->buf[0][3]) != cudaSuccess ||
      cudaIpcGetMemHandle(&mine.fl, G->flags) != cudaSuccess) { (void)cudaGetLastError(); return 1; }
  char *dsend = nullptr, *drecv = nullptr;
  if (cudaMalloc((void**)&dsend, sizeof(Pack)) != cudaSuccess || cudaMalloc((void**)&drecv, 2 * sizeof(Pack)) != cudaSuccess) return 1;
  cudaMemcpy(dsend, &mine, sizeof(Pack), cudaMemcpyHostToDevice);
  cudaStream_t st = (cudaStream_t)c->stream;
  const int ncclChar = 0;
  g_nccl.GroupStart();
  if (xs) { g_nccl.Send(dsend, sizeof(Pack), ncclChar, G->rank - 1, G->nccl, st); g_nccl.Recv(drecv, sizeof(Pack), ncclChar, G->rank - 1, G->nccl, st); }
  if (xn) { g_nccl.Send(dsend, sizeof(Pack), ncclChar, G->rank + 1, G->nccl, st); g_nccl.Recv(drecv + sizeof(Pack), sizeof(Pack), ncclChar, G->rank + 1, G->nccl, st); }
  int rc = g_nccl.GroupEnd();
  cudaStreamSynchronize(st);
  if (rc) return 1;
  if (xs) cudaMemcpy(&south, drecv, sizeof(Pack), cudaMemcpyDeviceToHost);
  if (xn) cudaMemcpy(&north, drecv + sizeof(Pack), sizeof(Pack), cudaMemcpyDeviceToHost);
  cudaFree(dsend); cudaFree(drecv);
  int ok = 1;
  if (xs) {   // I write into the south neighbour's receive-from-north buffer
    ok &= cudaIpcOpenMemHandle((void**)&G->peer_rbuf[0], south.rn, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    ok &= cudaIpcOpenMemHandle((void**)&G->peer_flags[0], south.fl, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
  }
  if (xn) {
    ok &= cudaIpcOpenMemHandle((void**)&G->peer_rbuf[1], north.rs, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    ok &= cudaIpcOpenMemHandle((void**)&G->peer_flags[1], north.fl, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
  }
  // every rank must take the same decision: agree through one more (tiny) exchange
  int *dok = nullptr, hok[3] = {ok, 1, 1};
  cudaMalloc((void**)&dok, 3 * sizeof(int));
  cudaMemcpy(dok, hok, 3 * sizeof(int), cudaMemcpyHostToDevice);
  const int ncclInt = 2;
  g_nccl.GroupStart();
  if (xs) { g_nccl.Send(dok, 1, ncclInt, G->rank - 1, G->nccl, st); g_nccl.Recv(dok + 1, 1, ncclInt, G->rank - 1, G->nccl, st); }
  if (xn) { g_nccl.Send(dok, 1, ncclInt, G->rank + 1, G->nccl, st); g_nccl.Recv(dok + 2, 1, ncclInt, G->rank + 1, G->nccl, st); }
  g_nccl.GroupEnd();
  cudaStreamSynchronize(st);
  cudaMemcpy(hok, dok, 3 * sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(dok);
  (void)cudaGetLastError();
  // (a chain: a failure anywhere must reach every rank; neighbours agree pairwise, and a rank whose
  // neighbour failed fails too, which is enough because the fallback is decided per seam below)
  G->ipc = ok && hok[1] && hok[2];
  G->ipc_cap = cap;
  G->seq = 0;
  return G->ipc ? 0 : 1;
}

// one exchange over the mapped peer buffers, enqueued on stream `st` (after the packs): ns / nn doubles
static int ipc_transfer(Group* G, cudaStream_t st, long ns, long nn) {
  Ctx* c = G->c[0];
  const bool xs = has_s(c), xn = has_n(c);
  const unsigned seq = ++G->seq;
  CUstream s = (CUstream)st;
  CUdeviceptr mine = (CUdeviceptr)G->flags;
  int bad = 0;
  // the neighbour must have unpacked what I sent last time before I overwrite its buffer
  if (xs) bad |= p_wait32(s, mine + 2 * 4, seq - 1, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS;
  if (xn) bad |= p_wait32(s, mine + 3 * 4, seq - 1, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS;
  if (xs) bad |= cudaMemcpyAsync(G->peer_rbuf[0], G->buf[0][0], (size_t)ns * 8, cudaMemcpyDeviceToDevice, st) != cudaSuccess;
  if (xn) bad |= cudaMemcpyAsync(G->peer_rbuf[1], G->buf[0][1], (size_t)nn * 8, cudaMemcpyDeviceToDevice, st) != cudaSuccess;
  // "arrived": into the south neighbour's from-north word, the north neighbour's from-south word
  if (xs) bad |= p_write32(s, (CUdeviceptr)G->peer_flags[0] + 1 * 4, seq, CU_STREAM_WRITE_VALUE_DEFAULT) != CUDA_SUCCESS;
  if (xn) bad |= p_write32(s, (CUdeviceptr)G->peer_flags[1] + 0 * 4, seq, CU_STREAM_WRITE_VALUE_DEFAULT) != CUDA_SUCCESS;
  // my own rows from the neighbours
  if (xs) bad |= p_wait32(s, mine + 0 * 4, seq, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS;
  if (xn) bad |= p_wait32(s, mine + 1 * 4, seq, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS;
  return bad;
}
// after the unpacks: tell the neighbours that their rows have been consumed
static int ipc_ack(Group* G, cudaStream_t st) {
  Ctx* c = G->c[0];
  CUstream s = (CUstream)st;
  int bad = 0;
  if (has_s(c)) bad |= p_write32(s, (CUdeviceptr)G->peer_flags[0] + 3 * 4, G->seq, CU_STREAM_WRITE_VALUE_DEFAULT) != CUDA_SUCCESS;
  if (has_n(c)) bad |= p_write32(s, (CUdeviceptr)G->peer_flags[1] + 2 * 4, G->seq, CU_STREAM_WRITE_VALUE_DEFAULT) != CUDA_SUCCESS;
  return bad;
}
#endif

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
  if (!getenv("POMGPU_HALO_NCCL") && ipc_connect(G))   // peer copies when the neighbours' memory can be mapped
    fprintf(stderr, "pomgpu: CUDA IPC peer mapping not available, halo rows go through ncclSend/ncclRecv\n");
  return 0;
#endif
}

void group_set_callback(Group* G, halo_cb cb, void* user) { G->cb = cb; G->cb_user = user; }


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
int group_exchange(Group* G, const int* fields, int nf, bool pack_on_comm) {
  if (!G->seams || nf == 0) return 0;
  // (timing experiments only: 1 = skip the transfer, 2 = skip pack / unpack as well; results are wrong)
  static const int dbg = getenv("POMGPU_HALO_DEBUG") ? atoi(getenv("POMGPU_HALO_DEBUG")) : 0;
  const FieldInfo* tab;
  int ntab;
  tab = field_table(&ntab);
  const int gh = G->ghost;
  if (G->ipc && nf > 1) {   // the mapped staging buffers cannot grow: an oversized batch goes in two halves
    size_t tot = 0;
    for (int n = 0; n < nf; ++n) tot += (size_t)(tab[fields[n]].kind == K3D ? G->c[0]->g.kb : 1) * gh * G->c[0]->g.im;
    if (tot > G->ipc_cap) return group_exchange(G, fields, nf / 2, pack_on_comm) || group_exchange(G, fields + nf / 2, nf - nf / 2, pack_on_comm);
  }
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double th0 = now();
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
          if (G->buf[r][b]) {
            dev_sync(c);
#ifndef POMGPU_EMU
            if (G->c[0]->comm_stream) cudaStreamSynchronize((cudaStream_t)G->c[0]->comm_stream);
#endif
            dev_free(c, G->buf[r][b]);
          }
          G->bufcap[r][b] = (size_t)P.total * 2;
          if (dev_alloc(c, &G->buf[r][b], G->bufcap[r][b])) return fail(G, "halo exchange: out of device memory");
        }
      }
#ifndef POMGPU_EMU
      if (G->trace && r == 0 && G->ntrace < 256) {
        for (int q = 0; q < 4; ++q) if (!G->tev[G->ntrace][q]) { cudaEvent_t e; cudaEventCreate(&e); G->tev[G->ntrace][q] = (void*)e; }
        G->tbytes[G->ntrace] = P.total * 8;
        cudaEventRecord((cudaEvent_t)G->tev[G->ntrace][0], (cudaStream_t)c->stream);
      }
#endif
    }
    // Transfer and unpack run on the communication stream when the transport can overlap
    // (xs_[r] = the stream the rest of strip r's exchange is enqueued on); the packs too when no
    // exchanged field is written by the kernel about to be launched (pack_on_comm), otherwise they
    // stay on the compute stream, ahead of that kernel.
    void* xs_[16];
    bool ov = false;
#ifndef POMGPU_EMU
    ov = G->overlap && !G->cb;
    for (int r = 0; r < G->n && ov; ++r) ov = comm_ready(G->c[r]);
#endif
    const bool early = ov && pack_on_comm;
    for (int r = 0; r < G->n; ++r) xs_[r] = G->c[r]->stream;
#ifndef POMGPU_EMU
    if (ov) for (int r = 0; r < G->n; ++r) xs_[r] = G->c[0]->comm_stream;   // strips of one device share both streams
    auto fork = [&]() {   // the communication stream continues from this point of the compute stream
      Ctx* c0 = G->c[0];
      cudaSetDevice(c0->device);
      cudaEventRecord((cudaEvent_t)c0->ev_packed, (cudaStream_t)c0->stream);
      cudaStreamWaitEvent((cudaStream_t)c0->comm_stream, (cudaEvent_t)c0->ev_packed, 0);
    };
    if (early) fork();
#endif
    for (int r = 0; r < G->n && dbg < 2; ++r) {
      Ctx* c = G->c[r];
      void* ps = early ? xs_[r] : c->stream;
      if (has_s(c)) run_pack(c, J[r][0], G->buf[r][0], false, ps);
      if (has_n(c)) run_pack(c, J[r][1], G->buf[r][1], false, ps);
    }
#ifndef POMGPU_EMU
    if (ov && !early) fork();
    if (ov && G->trace && G->ntrace < 256) cudaEventRecord((cudaEvent_t)G->tev[G->ntrace][1], (cudaStream_t)xs_[0]);
#endif
    G->host_ms[0] += now() - th0; th0 = now();
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
    if ((xs || xn) && dbg == 0) {
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
      else if (G->ipc && (size_t)J[0][0].total <= G->ipc_cap) {
        if (ipc_transfer(G, (cudaStream_t)xs_[0], xs ? J[0][0].total : 0, xn ? J[0][1].total : 0)) return fail(G, "halo exchange: peer copy / stream memory operation failed");
      }
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
    G->host_ms[1] += now() - th0; th0 = now();
#ifndef POMGPU_EMU
    if (G->trace && ov && G->ntrace < 256) cudaEventRecord((cudaEvent_t)G->tev[G->ntrace][2], (cudaStream_t)xs_[0]);
#endif
    for (int r = 0; r < G->n; ++r) {
      Ctx* c = G->c[r];
      if (dbg >= 2) continue;
      if (has_s(c)) run_pack(c, J[r][2], G->buf[r][2], true, xs_[r]);
      if (has_n(c)) run_pack(c, J[r][3], G->buf[r][3], true, xs_[r]);
    }
#ifndef POMGPU_EMU
    if (G->ipc && (xs || xn) && dbg == 0 && !G->cb && (size_t)J[0][0].total <= G->ipc_cap && ipc_ack(G, (cudaStream_t)xs_[0]))
      return fail(G, "halo exchange: stream memory operation failed");
    if (ov) {   // the kernel that asked for these rows waits for this event before it touches a seam band
      Ctx* c0 = G->c[0];
      cudaSetDevice(c0->device);
      cudaEventRecord((cudaEvent_t)c0->ev_halo, (cudaStream_t)c0->comm_stream);
      if (G->trace && G->ntrace < 256) { cudaEventRecord((cudaEvent_t)G->tev[G->ntrace][3], (cudaStream_t)c0->comm_stream); G->ntrace++; }
      G->ov_active = 1;
    }
#else
    if (G->overlap && !G->cb) G->ov_active = 1;   // the emulation runs everything in order, but splits the windows the same way
#endif
  }
  G->host_ms[2] += now() - th0; G->host_n++;
  for (int n = 0; n < nf; ++n) G->valid[fields[n]] = gh;
  return 0;
}

// A kernel is about to read fields `in` (id, j-radius).  Exchange what is too stale, then
// return how many ghost rows the kernel can also compute (its window extends that far).
//
// Policy.  A message costs latency, not bandwidth, for the 2-D fields (0.8 MB per exchange at
// im=1024) but real time for the 3-D ones (60 MB for all of them), so:
//  * whenever anything is exchanged, every live 2-D field that is not at full depth goes along;
//  * a kernel with 3-D inputs first gets its 2-D inputs refreshed if THEY would limit how many ghost
//    rows it can compute: otherwise the few rows of validity the external mode leaves on dt, etf, ...
//    would be inherited by the kernel's 3-D outputs and force the big exchange every step (measured:
//    two 60 MB exchanges per step before, one every two to three steps after);
//  * when a 3-D field triggers the exchange, every live 3-D field goes along (they age together);
//  * fields the kernel is about to overwrite (WILL) are not refreshed for nothing -- and when none of
//    the exchanged fields is one of them, the packs can run on the communication stream too.
void group_will(Group* G, const int* out, int n) {
  G->nwill = 0;
  for (int q = 0; q < n && q < 32; ++q) G->will[G->nwill++] = out[q];
}

int group_need(Group* G, const Req* in, int n) {
  if (!G->seams) return 0;
  for (int q = 0; q < n; ++q)
    if (in[q].r > G->ov_r) G->ov_r = in[q].r;
  static const int live2d[] = {F_ua, F_va, F_uab, F_vab, F_el, F_elb, F_d, F_dt, F_et, F_etb, F_etf, F_egf,
                               F_egb, F_utf, F_vtf, F_utb, F_vtb, F_wubot, F_wvbot, F_aam2d, F_adx2d,
                               F_ady2d, F_drx2d, F_dry2d, F_advua, F_advva, F_vfluxb};
  static const int live3d[] = {F_u, F_v, F_ub, F_vb, F_t, F_s, F_tb, F_sb, F_q2, F_q2b, F_q2l, F_q2lb, F_w,
                               F_aam, F_km, F_kh, F_kq, F_rho, F_advx, F_advy, F_drhox, F_drhoy};
  int ntab;
  const FieldInfo* tab = field_table(&ntab);
  auto willed = [&](int f) { for (int q = 0; q < G->nwill; ++q) if (G->will[q] == f) return true; return false; };
  // what the inputs allow, separately for the 3-D and the 2-D (and 1-D edge) ones
  int e3 = G->ghost, e2 = G->ghost;
  bool has3 = false, must = false, any3d = false;
  for (int q = 0; q < n; ++q) {
    const int v = G->valid[in[q].f] - in[q].r;
    if (tab[in[q].f].kind == K3D) { has3 = true; if (v < e3) e3 = v; if (v < 0) any3d = true; }
    else if (v < e2) e2 = v;
    if (v < 0) must = true;
  }
  const bool two_d_limits = has3 && e2 < e3 && e2 < G->ghost - G->ov_r;
  if (must || two_d_limits) {
    int stale[HALO_MAXF], ns = 0;
    bool hazard = false;   // an exchanged field is written by the kernel about to run
    // (a field the kernel updates IN PLACE is both an input -- its ghost rows limit e -- and written:
    // it is refreshed, but then the packs must stay ahead of the kernel on the compute stream)
    auto add = [&](int f, bool input) {
      for (int s = 0; s < ns; ++s) if (stale[s] == f) return;
      if (G->valid[f] >= G->ghost || ns >= HALO_MAXF) return;
      if (willed(f)) { if (!input) return; hazard = true; }
      stale[ns++] = f;
    };
    for (int q = 0; q < n; ++q)
      if (G->valid[in[q].f] < in[q].r || tab[in[q].f].kind != K3D || any3d) add(in[q].f, true);
    for (size_t q = 0; q < sizeof(live2d) / sizeof(int); ++q) add(live2d[q], false);
    if (any3d) for (size_t q = 0; q < sizeof(live3d) / sizeof(int); ++q) add(live3d[q], false);
    if (ns) {
      if (getenv("POMGPU_HALO_WHY")) { int n3 = 0; for (int q = 0; q < ns; ++q) n3 += tab[stale[q]].kind == K3D; fprintf(stderr, "exchange: must=%d 2dlimits=%d e2=%d e3=%d first-input=%s nfields=%d (3-D %d)\n", must, two_d_limits, e2, e3, tab[in[0].f].name, ns, n3); }
      if (G->ov_active) group_wait_halo(G);         // (a second exchange before the launch: keep the streams in order)
      if (group_exchange(G, stale, ns, !hazard)) return 0;   // G->failed is set: the caller launches nothing
    }
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
  G->nwill = 0;
}

// per traced exchange: bytes per direction, ms from the start of the pack to the start of the
// transfer (pack), to the end of the transfer, to the end of the unpack, and the start relative to the first
int group_trace_report(Group* G, char* buf, int n) {
  int o = 0;
#ifndef POMGPU_EMU
  for (int r = 0; r < G->n; ++r) dev_sync(G->c[r]);
  if (G->c[0]->comm_stream) cudaStreamSynchronize((cudaStream_t)G->c[0]->comm_stream);
  for (int q = 0; q < G->ntrace && o < n - 120; ++q) {
    float t0 = 0, a = 0, b = 0, c3 = 0;
    cudaEventElapsedTime(&t0, (cudaEvent_t)G->tev[0][0], (cudaEvent_t)G->tev[q][0]);
    cudaEventElapsedTime(&a, (cudaEvent_t)G->tev[q][0], (cudaEvent_t)G->tev[q][1]);
    cudaEventElapsedTime(&b, (cudaEvent_t)G->tev[q][0], (cudaEvent_t)G->tev[q][2]);
    cudaEventElapsedTime(&c3, (cudaEvent_t)G->tev[q][0], (cudaEvent_t)G->tev[q][3]);
    o += snprintf(buf + o, n - o, "t=%8.3f ms  %8.2f MB  pack %.3f  +transfer %.3f  +unpack %.3f\n", t0, G->tbytes[q] / 1e6, a, b, c3);
  }
#endif
  if (o < n - 160)
    o += snprintf(buf + o, n - o, "host ms over %ld exchanges: packs %.3f  transport %.3f  unpacks %.3f\n", G->host_n, G->host_ms[0], G->host_ms[1], G->host_ms[2]);
  G->host_ms[0] = G->host_ms[1] = G->host_ms[2] = 0.; G->host_n = 0;
  G->ntrace = 0;
  if (o < n) buf[o] = 0;
  return o;
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
