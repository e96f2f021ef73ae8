// pom_state.cu -- context lifecycle, field registry and host<->HBM transfers.
// Device copies of the reference's COMMON blocks (pom.h_dist:142-198,208-212,
// 291-364,410-450,532-608) with run-time extents; the driver's host arrays stay
// authoritative only at init, after a pull, and for per-step forcing pushes
// (SURVEY.md 8(b)).
#include "pom_core.h"
#include "pom_tma.h"
#include <cstdlib>

namespace pom {

static const FieldInfo g_fields[] = {
#define X(n) {#n, K3D, offsetof(Ptrs, n), false, false},
    POM_F3D(X)
#undef X
#define X(n) {#n, K3D, offsetof(Ptrs, n), true, false},
    POM_F3D_OPT(X)
#undef X
#define X(n) {#n, K3D, offsetof(Ptrs, n), false, true},
    POM_F3D_SCR(X)
#undef X
#define X(n) {#n, K2D, offsetof(Ptrs, n), false, false},
    POM_F2D(X)
#undef X
#define X(n) {#n, K2D, offsetof(Ptrs, n), false, true},
    POM_F2D_SCR(X)
#undef X
#define X(n) {#n, KBJ, offsetof(Ptrs, n), false, false},
    POM_BJ(X)
#undef X
#define X(n) {#n, KBI, offsetof(Ptrs, n), false, false},
    POM_BI(X)
#undef X
#define X(n) {#n, KBJK, offsetof(Ptrs, n), false, false},
    POM_BJK(X)
#undef X
#define X(n) {#n, KBIK, offsetof(Ptrs, n), false, false},
    POM_BIK(X)
#undef X
#define X(n) {#n, K1D, offsetof(Ptrs, n), false, false},
    POM_F1D(X)
#undef X
};

const FieldInfo* field_table(int* n) {
  *n = (int)(sizeof(g_fields) / sizeof(g_fields[0]));
  return g_fields;
}

const FieldInfo* find_field(const char* name) {
  int n;
  const FieldInfo* t = field_table(&n);
  for (int i = 0; i < n; ++i)
    if (!strcmp(t[i].name, name)) return &t[i];
  return nullptr;
}

size_t field_elems(const Ctx* c, const FieldInfo* f) {
  const Geo& g = c->g;
  switch (f->kind) {
    case K3D: return (size_t)g.n2 * g.kb;
    case K2D: return (size_t)g.n2;
    case KBJ: return g.jml;
    case KBI: return g.im;
    case KBJK: return (size_t)g.jml * g.kb;
    case KBIK: return (size_t)g.im * g.kb;
    case K1D: return g.kb;
  }
  return 0;
}

// ---- backend -----------------------------------------------------------------
#ifdef POMGPU_EMU
int dev_init(Ctx* c) { c->stream = nullptr; c->own_stream = nullptr; c->d_red = (double*)calloc(4096, 8); c->h_red = (double*)calloc(4096, 8); return 0; }
int dev_alloc(Ctx*, double** p, size_t n) { *p = (double*)calloc(n ? n : 1, sizeof(double)); return *p ? 0 : 1; }
void dev_free(Ctx*, double* p) { free(p); }
int dev_h2d(Ctx*, double* d, const double* s, size_t n) { memcpy(d, s, n * 8); return 0; }
int dev_d2h(Ctx*, double* d, const double* s, size_t n) { memcpy(d, s, n * 8); return 0; }
int dev_d2d(Ctx*, double* d, const double* s, size_t n) { memmove(d, s, n * 8); return 0; }
int dev_zero(Ctx*, double* p, size_t n) { memset(p, 0, n * 8); return 0; }
int dev_sync(Ctx*) { return 0; }
#else
static int cuda_fail(Ctx* c, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  snprintf(c->err, sizeof(c->err), "CUDA error in %s: %s", what, cudaGetErrorString(e));
  fprintf(stderr, "pomgpu: %s\n", c->err);
  c->c.error_status = 1;  // reference error convention (advance.f:118,556-563)
  return 1;
}
int dev_init(Ctx* c) {
  int n = 0;
  if (cuda_fail(c, cudaGetDeviceCount(&n), "cudaGetDeviceCount") || n == 0) {
    snprintf(c->err, sizeof(c->err), "no CUDA device: libpomgpu has no CPU fallback");
    fprintf(stderr, "pomgpu: %s\n", c->err);
    return 1;
  }
  if (cuda_fail(c, cudaSetDevice(c->device), "cudaSetDevice")) return 1;
  if (cuda_fail(c, cudaDeviceGetAttribute(&c->nsm, cudaDevAttrMultiProcessorCount, c->device), "device attribute")) return 1;
  cudaStream_t s;
  if (cuda_fail(c, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate")) return 1;
  c->stream = (void*)s;
  c->own_stream = (void*)s;
  if (cuda_fail(c, cudaMalloc((void**)&c->d_red, 4096 * 8), "cudaMalloc")) return 1;
  if (cuda_fail(c, cudaMallocHost((void**)&c->h_red, 4096 * 8), "cudaMallocHost")) return 1;
  return 0;
}
int dev_alloc(Ctx* c, double** p, size_t n) {
  cudaSetDevice(c->device);
  if (cuda_fail(c, cudaMalloc((void**)p, (n ? n : 1) * sizeof(double)), "cudaMalloc")) return 1;
  return cuda_fail(c, cudaMemsetAsync(*p, 0, (n ? n : 1) * sizeof(double), (cudaStream_t)c->stream), "cudaMemset");
}
void dev_free(Ctx* c, double* p) { cudaSetDevice(c->device); cudaFree(p); }
int dev_h2d(Ctx* c, double* d, const double* s, size_t n) {
  cudaSetDevice(c->device);
  if (cuda_fail(c, cudaMemcpyAsync(d, s, n * 8, cudaMemcpyHostToDevice, (cudaStream_t)c->stream), "H2D")) return 1;
  return cuda_fail(c, cudaStreamSynchronize((cudaStream_t)c->stream), "H2D sync");
}
int dev_d2h(Ctx* c, double* d, const double* s, size_t n) {
  cudaSetDevice(c->device);
  if (cuda_fail(c, cudaMemcpyAsync(d, s, n * 8, cudaMemcpyDeviceToHost, (cudaStream_t)c->stream), "D2H")) return 1;
  return cuda_fail(c, cudaStreamSynchronize((cudaStream_t)c->stream), "D2H sync");
}
int dev_d2d(Ctx* c, double* d, const double* s, size_t n) {
  return cuda_fail(c, cudaMemcpyAsync(d, s, n * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)c->stream), "D2D");
}
int dev_zero(Ctx* c, double* p, size_t n) {
  return cuda_fail(c, cudaMemsetAsync(p, 0, n * 8, (cudaStream_t)c->stream), "memset");
}
int dev_sync(Ctx* c) {
  cudaSetDevice(c->device);
  if (cuda_fail(c, cudaStreamSynchronize((cudaStream_t)c->stream), "stream sync")) return 1;
  return cuda_fail(c, cudaGetLastError(), "kernel launch");
}
#endif

#ifndef POMGPU_EMU
// ---- TMA tensor maps ------------------------------------------------------------------
// A field is described to the TMA as a 3-D fp64 tensor (im, jml, nk), i fastest, box
// (bw, bh, 1), no swizzle, out-of-range elements filled with zeros.  cuTensorMapEncodeTiled
// is a pure host function; it is resolved through the runtime so that libpomgpu does not link
// libcuda directly.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// descriptors are cached per (buffer, levels, box): a step re-uses ~100 distinct ones ~600 times
struct TmaCacheEnt { const double* base; int nk, bw, bh; CUtensorMap m; };
static TmaCacheEnt* tma_cache_of(Ctx* c) {
  if (!c->tma_cache) c->tma_cache = calloc(512, sizeof(TmaCacheEnt));
  return (TmaCacheEnt*)c->tma_cache;
}
int tma_encode(Ctx* c, CUtensorMap* m, const double* base, int nk, int bw, int bh, int bk) {
  TmaCacheEnt* tc = tma_cache_of(c);
  bh += 4096 * (bk - 1);   // (cache key: the box depth folded into the height)
  const size_t hsh = (((size_t)base >> 8) * 2654435761u + (size_t)nk * 97 + (size_t)bw * 7 + (size_t)bh) & 511;
  TmaCacheEnt& ce = tc[hsh];
  if (ce.base == base && ce.nk == nk && ce.bw == bw && ce.bh == bh) { *m = ce.m; return 0; }
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return 1;
    fn = (EncodeTiledFn)p;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)c->g.im, (cuuint64_t)c->g.jml, (cuuint64_t)nk};
  const cuuint64_t strides[2] = {(cuuint64_t)c->g.im * 8, (cuuint64_t)c->g.n2 * 8};
  const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)(bh % 4096), (cuuint32_t)bk};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return 1;
  ce.base = base; ce.nk = nk; ce.bw = bw; ce.bh = bh; ce.m = *m;
  return 0;
}
#endif

// ---- per-launch profiling with CUDA events on the launch stream ---------------------
#ifdef POMGPU_EMU
void prof_before(Ctx*, const KInfo*, double) {}
void prof_after(Ctx*) {}
int prof_report(Ctx*, char* buf, int n) { if (n > 2) strcpy(buf, "[]"); return 0; }
#else
void prof_before(Ctx* c, const KInfo* info, double bytes) {
  if (c->nprof == c->capprof) {
    c->capprof = c->capprof ? 2 * c->capprof : 1024;
    c->prof = (ProfRec*)realloc(c->prof, sizeof(ProfRec) * c->capprof);
  }
  ProfRec& r = c->prof[c->nprof];
  r.info = info; r.bytes = bytes;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  r.e0 = a; r.e1 = b;
  cudaEventRecord(a, (cudaStream_t)c->stream);
}
void prof_after(Ctx* c) {
  cudaEventRecord((cudaEvent_t)c->prof[c->nprof].e1, (cudaStream_t)c->stream);
  c->nprof++;
}
// JSON array: [{"name":..,"launches":n,"ms":total,"bytes":total algorithmic}, ...]
int prof_report(Ctx* c, char* buf, int n) {
  dev_sync(c);
  struct Agg { const KInfo* k; long cnt; double ms, bytes; } agg[64];
  int na = 0;
  for (int i = 0; i < c->nprof; ++i) {
    ProfRec& r = c->prof[i];
    float ms = 0.f;
    cudaEventElapsedTime(&ms, (cudaEvent_t)r.e0, (cudaEvent_t)r.e1);
    cudaEventDestroy((cudaEvent_t)r.e0); cudaEventDestroy((cudaEvent_t)r.e1);
    int a = 0;
    while (a < na && agg[a].k != r.info) ++a;
    if (a == na) { if (na == 64) continue; agg[na++] = {r.info, 0, 0., 0.}; }
    agg[a].cnt++; agg[a].ms += ms; agg[a].bytes += r.bytes;
  }
  c->nprof = 0;
  int o = snprintf(buf, n, "[");
  for (int a = 0; a < na && o < n; ++a)
    o += snprintf(buf + o, n - o, "%s{\"name\":\"%s\",\"launches\":%ld,\"ms\":%.6f,\"bytes\":%.0f}",
                  a ? "," : "", agg[a].k->name, agg[a].cnt, agg[a].ms, agg[a].bytes);
  if (o < n) o += snprintf(buf + o, n - o, "]");
  return o < n ? 0 : 2;
}
#endif

// ---- lifecycle ------------------------------------------------------------------
Ctx* ctx_create(int im, int jm_global, int kb, int j_first, int j_last, int ghost, int device) {
  if (im < 6 || jm_global < 6 || kb < 4 || kb > KMAX) return nullptr;   // most levels the column solvers hold
  Ctx* c = (Ctx*)calloc(1, sizeof(Ctx));
  c->hz = c->hk[0];
  c->no_tma = (getenv("POMGPU_NO_TMA") != nullptr);
  c->device = device;
  c->jown0 = j_first; c->jown1 = j_last; c->ghost = ghost;
  int r0 = j_first - ghost; if (r0 < 1) r0 = 1;
  int r1 = j_last + ghost; if (r1 > jm_global) r1 = jm_global;
  c->g.im = im; c->g.kb = kb; c->g.jmg = jm_global;
  c->g.joff = r0 - 1; c->g.jml = r1 - r0 + 1;
  if ((double)im * c->g.jml * kb >= 2147483647.) { free(c); return nullptr; }
  c->g.n2 = im * c->g.jml;
  if (dev_init(c)) { free(c); return nullptr; }
  int n;
  const FieldInfo* t = field_table(&n);
  for (int i = 0; i < n; ++i) {
    double** slot = (double**)((char*)&c->p + t[i].offset);
    *slot = nullptr;
    if (t[i].optional) continue;
    if (dev_alloc(c, slot, field_elems(c, &t[i]))) { return nullptr; }
  }
  dev_sync(c);
  return c;
}

void ctx_destroy(Ctx* c) {
  if (!c) return;
  int n;
  const FieldInfo* t = field_table(&n);
  dev_sync(c);
  record_free(c);
  for (int i = 0; i < n; ++i) {
    double** slot = (double**)((char*)&c->p + t[i].offset);
    if (*slot) dev_free(c, *slot);
  }
#ifdef POMGPU_EMU
  free(c->d_red); free(c->h_red);
#else
  cudaFree(c->d_red); cudaFreeHost(c->h_red);
  for (int f = 0; f < 256; ++f) if (c->shadow[f]) cudaFree(c->shadow[f]);
  free(c->tma_cache);
  if (c->copy_stream) cudaStreamDestroy((cudaStream_t)c->copy_stream);
  if (c->ev_copied) cudaEventDestroy((cudaEvent_t)c->ev_copied);
  if (c->ev_swapped) cudaEventDestroy((cudaEvent_t)c->ev_swapped);
  if (c->ev_vel) cudaEventDestroy((cudaEvent_t)c->ev_vel);
  for (int i = 0; i < 8; ++i) if (c->ev[i]) cudaEventDestroy((cudaEvent_t)c->ev[i]);
  if (c->comm_stream) cudaStreamDestroy((cudaStream_t)c->comm_stream);
  if (c->ev_packed) cudaEventDestroy((cudaEvent_t)c->ev_packed);
  if (c->ev_halo) cudaEventDestroy((cudaEvent_t)c->ev_halo);
  cudaStreamDestroy((cudaStream_t)c->own_stream);
#endif
  free(c);
}

double** ctx_slot(Ctx* c, const char* name, const FieldInfo** fi) {
  const FieldInfo* f = find_field(name);
  if (fi) *fi = f;
  if (!f) return nullptr;
  return (double**)((char*)&c->p + f->offset);
}

int ctx_push(Ctx* c, const char* name, const double* host) {
  const FieldInfo* f;
  double** slot = ctx_slot(c, name, &f);
  if (!slot) { snprintf(c->err, sizeof(c->err), "unknown field '%s'", name); return 2; }
  if (!*slot && dev_alloc(c, slot, field_elems(c, f))) return 1;
  for (int t = 0; t < 4; ++t)
    if (!strcmp(name, t == 0 ? "z" : t == 1 ? "zz" : t == 2 ? "dz" : "dzz"))
      for (int k = 0; k < c->g.kb && k < KMAX; ++k) c->hk[t][k] = host[k];
  return dev_h2d(c, *slot, host, field_elems(c, f));
}

// rows [row0, row0+nrows) (local, 0-based) of a j-dimensioned field from a host array shaped like
// the field but with nrows rows: lets a driver fill a large strip band by band
int ctx_push_rows(Ctx* c, const char* name, const double* host, int row0, int nrows) {
  const FieldInfo* f;
  double** slot = ctx_slot(c, name, &f);
  if (!slot) { snprintf(c->err, sizeof(c->err), "unknown field '%s'", name); return 2; }
  if (row0 < 0 || nrows < 1 || row0 + nrows > c->g.jml) return 2;
  if (!*slot && dev_alloc(c, slot, field_elems(c, f))) return 1;
  const int im = c->g.im, jml = c->g.jml, kb = c->g.kb;
  size_t w, sp, dp, off; int nk;     // row run (doubles), source / destination level pitch, offset, levels
  switch (f->kind) {
    case K3D: w = (size_t)im * nrows; sp = w; dp = (size_t)im * jml; off = (size_t)im * row0; nk = kb; break;
    case K2D: w = (size_t)im * nrows; sp = w; dp = (size_t)im * jml; off = (size_t)im * row0; nk = 1; break;
    case KBJ: w = nrows; sp = w; dp = jml; off = row0; nk = 1; break;
    case KBJK: w = nrows; sp = w; dp = jml; off = row0; nk = kb; break;
    default: return ctx_push(c, name, host);   // no j dimension: the whole array
  }
  for (int k = 0; k < nk; ++k)
    if (dev_h2d(c, *slot + off + (size_t)k * dp, host + (size_t)k * sp, w)) return 1;
  return 0;
}

// The same for a host array with GLOBAL extents (im, jm_global[, kb]) -- a COMMON member of a single-rank driver whose
// domain is spread over several strips (libpomgpu_f with more than one device): push takes the rows this strip
// holds (owned + ghost), pull returns the rows it OWNS, so the strips of a group assemble the array between them.
int ctx_pull(Ctx* c, const char* name, double* host);
size_t field_global_elems(const Ctx* c, const FieldInfo* f) {
  const Geo& g = c->g;
  switch (f->kind) {
    case K3D: return (size_t)g.im * g.jmg * g.kb;
    case K2D: return (size_t)g.im * g.jmg;
    case KBJ: return g.jmg;
    case KBJK: return (size_t)g.jmg * g.kb;
    default: return field_elems(c, f);
  }
}
int ctx_push_global(Ctx* c, const char* name, const double* host) {
  const FieldInfo* f;
  double** slot = ctx_slot(c, name, &f);
  if (!slot) { snprintf(c->err, sizeof(c->err), "unknown field '%s'", name); return 2; }
  if (!*slot && dev_alloc(c, slot, field_elems(c, f))) return 1;
  const size_t im = c->g.im, jml = c->g.jml, jmg = c->g.jmg, joff = c->g.joff;
  size_t w, sp, dp, off; int nk;     // run (doubles), source / destination level pitch, source offset, levels
  switch (f->kind) {
    case K3D: w = im * jml; sp = im * jmg; dp = w; off = im * joff; nk = c->g.kb; break;
    case K2D: w = im * jml; sp = im * jmg; dp = w; off = im * joff; nk = 1; break;
    case KBJ: w = jml; sp = jmg; dp = w; off = joff; nk = 1; break;
    case KBJK: w = jml; sp = jmg; dp = w; off = joff; nk = c->g.kb; break;
    default: return ctx_push(c, name, host);   // no j dimension: the whole array
  }
  for (int k = 0; k < nk; ++k)
    if (dev_h2d(c, *slot + (size_t)k * dp, host + off + (size_t)k * sp, w)) return 1;
  return 0;
}
int ctx_pull_global(Ctx* c, const char* name, double* host) {
  const FieldInfo* f;
  double** slot = ctx_slot(c, name, &f);
  if (!slot || !*slot) { snprintf(c->err, sizeof(c->err), "unknown or unallocated field '%s'", name); return 2; }
  const size_t im = c->g.im, jml = c->g.jml, jmg = c->g.jmg;
  const size_t g0 = (size_t)(c->jown0 - 1), l0 = g0 - (size_t)c->g.joff, nown = (size_t)(c->jown1 - c->jown0 + 1);
  size_t w, hp, dp, hoff, doff; int nk;   // run, host / device level pitch, host / device offset of the first owned row
  switch (f->kind) {
    case K3D: w = im * nown; hp = im * jmg; dp = im * jml; hoff = im * g0; doff = im * l0; nk = c->g.kb; break;
    case K2D: w = im * nown; hp = im * jmg; dp = im * jml; hoff = im * g0; doff = im * l0; nk = 1; break;
    case KBJ: w = nown; hp = jmg; dp = jml; hoff = g0; doff = l0; nk = 1; break;
    case KBJK: w = nown; hp = jmg; dp = jml; hoff = g0; doff = l0; nk = c->g.kb; break;
    default: return ctx_pull(c, name, host);
  }
  for (int k = 0; k < nk; ++k)
    if (dev_d2h(c, host + hoff + (size_t)k * hp, *slot + doff + (size_t)k * dp, w)) return 1;
  return 0;
}

int ctx_pull(Ctx* c, const char* name, double* host) {
  const FieldInfo* f;
  double** slot = ctx_slot(c, name, &f);
  if (!slot || !*slot) { snprintf(c->err, sizeof(c->err), "unknown or unallocated field '%s'", name); return 2; }
  return dev_d2h(c, host, *slot, field_elems(c, f));
}

int ctx_set_const(Ctx* c, const char* name, double v) {
#define X(n) if (!strcmp(name, #n)) { c->c.n = v; return 0; }
  POM_SCAL_D(X)
#undef X
#define X(n) if (!strcmp(name, #n)) { c->c.n = (int)v; return 0; }
  POM_SCAL_I(X)
#undef X
  return 2;
}

int ctx_get_const(Ctx* c, const char* name, double* v) {
#define X(n) if (!strcmp(name, #n)) { *v = c->c.n; return 0; }
  POM_SCAL_D(X)
#undef X
#define X(n) if (!strcmp(name, #n)) { *v = (double)c->c.n; return 0; }
  POM_SCAL_I(X)
#undef X
  return 2;
}

}  // namespace pom
