// pom_forcing.cu -- time interpolation of the surface forcing and the open-boundary records on the
// device.  The reference re-reads a forcing record every `iwind`/`iheat`/`ibc`
// internal steps but interpolates between the two bracketing records EVERY step on the host
// (bounds_forcing.f:841-865 lateral_bc, :904-909 wind, :949-957 heat); a driver of the resident
// library would have to push ~25 arrays per step.  Here the driver pushes a record only when the
// reference reads one (pomgpu_push_record, asynchronous on the copy stream), rotates f -> b by a
// pointer swap where the reference copies (`wusurfb=wusurff`, :888-893), and the per-step
// interpolation `x = fold*xb + fnew*xf`, fold = 1.-fnew, runs as one launch on the compute stream.
#include "pom_core.h"
#include <cstdlib>

namespace pom {

#define INTERP_MAXF 12
struct InterpJob {
  double* dst[INTERP_MAXF];
  const double* b[INTERP_MAXF];
  const double* f[INTERP_MAXF];
  long end[INTERP_MAXF];   // running end offset of segment n in the flattened index space
  int nf;
  double fold, fnew;
};

POM_HD void interp_one(const InterpJob& J, long e) {
  int n = 0;
  while (n < J.nf - 1 && e >= J.end[n]) ++n;
  const long q = e - (n ? J.end[n - 1] : 0);
  J.dst[n][q] = J.fold * J.b[n][q] + J.fnew * J.f[n][q];   // bounds_forcing.f:844,908,954
}

// lateral_bc's tail (bounds_forcing.f:844-865): one thread per edge point; level loop interpolates
// T, S and the normal velocity and accumulates the depth integral of the latter in k order.
struct EdgeJob {
  // [0]=west, [1]=east (pitch jml), [2]=north, [3]=south (pitch im)
  double* t[4]; double* s[4]; double* u[4]; double* ua[4];
  const double* tb[4]; const double* tf[4];
  const double* sb[4]; const double* sf[4];
  const double* ub[4]; const double* uf[4];
  const double* dz;
  int n[4];     // points per edge (jml, jml, im, im)
  int kb;
  double fold, fnew;
};

POM_HD void edge_one(const EdgeJob& J, int e, int q) {
  const int n = J.n[e];
  double acc = 0.;   // uabe = 0. (:856-859)
  for (int k = 0; k < J.kb; ++k) {
    const long a = (long)k * n + q;
    J.t[e][a] = J.fold * J.tb[e][a] + J.fnew * J.tf[e][a];
    J.s[e][a] = J.fold * J.sb[e][a] + J.fnew * J.sf[e][a];
    const double un = J.fold * J.ub[e][a] + J.fnew * J.uf[e][a];
    J.u[e][a] = un;
    acc = acc + un * J.dz[k];   // :860-865, k = 1..kb
  }
  J.ua[e][q] = acc;
}

#ifndef POMGPU_EMU
__global__ void interp_kernel(InterpJob J) {
  const long total = J.end[J.nf - 1];
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x)
    interp_one(J, e);
}
__global__ void edge_kernel(EdgeJob J) {
  const int e = blockIdx.y;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < J.n[e]; q += gridDim.x * blockDim.x) edge_one(J, e, q);
}
#endif

static int rec_id(Ctx* c, const char* name, const FieldInfo** fi) {
  const FieldInfo* f = find_field(name);
  int ntab;
  const FieldInfo* tab = field_table(&ntab);
  if (!f || f->scratch || f->kind == K1D) { snprintf(c->err, sizeof(c->err), "record: unknown field '%s'", name); return -1; }
  if (fi) *fi = f;
  const int id = (int)(f - tab);
  return id < 256 ? id : -1;
}

#ifndef POMGPU_EMU
static int rec_streams(Ctx* c) {
  cudaSetDevice(c->device);
  if (!c->rec_stream) {
    cudaStream_t s; cudaEvent_t e;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return 1;
    c->rec_stream = (void*)s;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming); c->ev_rec_copied = (void*)e;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming); c->ev_rec_read = (void*)e;
    cudaEventRecord((cudaEvent_t)c->ev_rec_read, (cudaStream_t)c->stream);
  }
  return 0;
}
#endif

// slot 0 = the older record ("b"), 1 = the newer ("f").  The copy is asynchronous w.r.t. the compute
// stream when `host` is page-locked; it is ordered after every interpolation enqueued so far (which
// may still read the buffer) and before the next one.
int record_push(Ctx* c, const char* name, int slot, const double* host) {
  const FieldInfo* f;
  const int id = rec_id(c, name, &f);
  if (id < 0 || slot < 0 || slot > 1) return 2;
  const size_t n = field_elems(c, f);
  double** r = &c->rec[id][slot];
#ifdef POMGPU_EMU
  if (!*r && dev_alloc(c, r, n)) return 1;
  return dev_h2d(c, *r, host, n);
#else
  if (rec_streams(c)) return 1;
  if (!*r && cudaMalloc((void**)r, (n ? n : 1) * sizeof(double)) != cudaSuccess) {
    snprintf(c->err, sizeof(c->err), "push_record(%s): out of device memory", name);
    c->c.error_status = 1;
    return 1;
  }
  cudaStream_t rs = (cudaStream_t)c->rec_stream;
  cudaStreamWaitEvent(rs, (cudaEvent_t)c->ev_rec_read, 0);
  cudaError_t e = cudaMemcpyAsync(*r, host, n * 8, cudaMemcpyHostToDevice, rs);
  if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "push_record(%s): %s", name, cudaGetErrorString(e)); c->c.error_status = 1; return 1; }
  cudaEventRecord((cudaEvent_t)c->ev_rec_copied, rs);
  return 0;
#endif
}

// `xb = xf` of the reference (bounds_forcing.f:888-893,933-938): the buffers change roles; the next
// record goes into slot 1
int record_rotate(Ctx* c, const char* name) {
  const int id = rec_id(c, name, nullptr);
  if (id < 0) return 2;
  double* t = c->rec[id][0]; c->rec[id][0] = c->rec[id][1]; c->rec[id][1] = t;
  return 0;
}

static void rec_before_launch(Ctx* c) {
#ifndef POMGPU_EMU
  cudaSetDevice(c->device);
  if (c->rec_stream) cudaStreamWaitEvent((cudaStream_t)c->stream, (cudaEvent_t)c->ev_rec_copied, 0);
#endif
}
static void rec_after_launch(Ctx* c) {
  c->launches++;
#ifndef POMGPU_EMU
  if (c->rec_stream) cudaEventRecord((cudaEvent_t)c->ev_rec_read, (cudaStream_t)c->stream);
#endif
}

// names: up to INTERP_MAXF field names; every one needs both records
int record_interp(Ctx* c, const char* const* names, int nn, double fnew) {
  if (nn < 1 || nn > INTERP_MAXF) return 2;
  InterpJob J;
  J.nf = nn; J.fnew = fnew; J.fold = 1. - fnew;   // fold=1.-fnew (:843,906,951)
  long tot = 0;
  for (int n = 0; n < nn; ++n) {
    const FieldInfo* f;
    const int id = rec_id(c, names[n], &f);
    if (id < 0) return 2;
    double** slot = (double**)((char*)&c->p + f->offset);
    if (!c->rec[id][0] || !c->rec[id][1] || !*slot) {
      snprintf(c->err, sizeof(c->err), "interp(%s): both records must be pushed first", names[n]);
      return 2;
    }
    J.dst[n] = *slot; J.b[n] = c->rec[id][0]; J.f[n] = c->rec[id][1];
    tot += (long)field_elems(c, f);
    J.end[n] = tot;
  }
  rec_before_launch(c);
#ifdef POMGPU_EMU
  for (long e = 0; e < tot; ++e) interp_one(J, e);
#else
  int blocks = (int)((tot + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  interp_kernel<<<blocks, 256, 0, (cudaStream_t)c->stream>>>(J);
#endif
  rec_after_launch(c);
  return 0;
}

// lateral_bc's interpolation (bounds_forcing.f:841-865): tbw,sbw,ubw,tbe,sbe,ube,tbn,sbn,vbn,tbs,sbs,vbs
// and the depth-integrated normal velocities uabw,uabe,vabn,vabs
int record_lateral_bc(Ctx* c, double fnew) {
  static const char* const T[4] = {"tbw", "tbe", "tbn", "tbs"};
  static const char* const S[4] = {"sbw", "sbe", "sbn", "sbs"};
  static const char* const U[4] = {"ubw", "ube", "vbn", "vbs"};
  static const char* const A[4] = {"uabw", "uabe", "vabn", "vabs"};
  EdgeJob J;
  J.fnew = fnew; J.fold = 1. - fnew; J.kb = c->g.kb; J.dz = c->p.dz;
  for (int e = 0; e < 4; ++e) {
    const FieldInfo *ft, *fs, *fu, *fa;
    const int it = rec_id(c, T[e], &ft), is = rec_id(c, S[e], &fs), iu = rec_id(c, U[e], &fu);
    if (it < 0 || is < 0 || iu < 0 || rec_id(c, A[e], &fa) < 0) return 2;
    for (int s = 0; s < 2; ++s)
      if (!c->rec[it][s] || !c->rec[is][s] || !c->rec[iu][s]) {
        snprintf(c->err, sizeof(c->err), "lateral_bc: records of %s/%s/%s must be pushed first", T[e], S[e], U[e]);
        return 2;
      }
    J.t[e] = *(double**)((char*)&c->p + ft->offset);
    J.s[e] = *(double**)((char*)&c->p + fs->offset);
    J.u[e] = *(double**)((char*)&c->p + fu->offset);
    J.ua[e] = *(double**)((char*)&c->p + fa->offset);
    J.tb[e] = c->rec[it][0]; J.tf[e] = c->rec[it][1];
    J.sb[e] = c->rec[is][0]; J.sf[e] = c->rec[is][1];
    J.ub[e] = c->rec[iu][0]; J.uf[e] = c->rec[iu][1];
    J.n[e] = e < 2 ? c->g.jml : c->g.im;
  }
  rec_before_launch(c);
#ifdef POMGPU_EMU
  for (int e = 0; e < 4; ++e)
    for (int q = 0; q < J.n[e]; ++q) edge_one(J, e, q);
#else
  const int nmax = c->g.jml > c->g.im ? c->g.jml : c->g.im;
  edge_kernel<<<dim3((nmax + 127) / 128, 4), 128, 0, (cudaStream_t)c->stream>>>(J);
#endif
  rec_after_launch(c);
  return 0;
}

void record_free(Ctx* c) {
  for (int f = 0; f < 256; ++f)
    for (int s = 0; s < 2; ++s)
      if (c->rec[f][s]) {
#ifdef POMGPU_EMU
        free(c->rec[f][s]);
#else
        cudaFree(c->rec[f][s]);
#endif
        c->rec[f][s] = nullptr;
      }
#ifndef POMGPU_EMU
  if (c->rec_stream) cudaStreamDestroy((cudaStream_t)c->rec_stream);
  if (c->ev_rec_copied) cudaEventDestroy((cudaEvent_t)c->ev_rec_copied);
  if (c->ev_rec_read) cudaEventDestroy((cudaEvent_t)c->ev_rec_read);
#endif
}

}  // namespace pom
