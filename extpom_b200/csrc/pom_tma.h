// pom_tma.h -- TMA-staged shared-memory tile kernels for the horizontal stencils.
//
// A block of 32 x TY threads owns a tile of points and marches in k.  The operand planes of
// every level (tile + halo, NF fields) are brought from HBM into a ring of NS shared-memory
// stages by the TMA (cp.async.bulk.tensor.3d, one elected thread, mbarrier complete_tx), two
// or more levels AHEAD of the level being computed: the loads in flight per SM are
// (NS-2) x NF x 4.6 kB whatever the register pressure / occupancy of the fp64 body is, which
// is what these latency-bound kernels were missing.  Out-of-range box coordinates (domain
// edge, strip edge) are zero-filled by the TMA.  Per level every thread evaluates the
// flux-like values of ITS OWN point once from the staged operands into a double-buffered
// flux tile (one __syncthreads per level); the tile interior differences its neighbours'.
//
// A functor F provides
//   static constexpr int NF, NV, HL, HR, HB, HT, TY, NS;  flux halo of the thread tile
//   static constexpr int OHL, OHR, OHB, OHT, BW, BH;       operand halo (W,E,S,N) and TMA box
//   static constexpr bool UP;                              reads level k+1 at its own point
//   void fields(const double* b[NF]);  k0(), k1(), kl1() (last level staged)
//   pre(i,j,inside,out,State&);  stage(i,j,k,State&,op,v[NV]);  combine(i,j,k,State&,op,Tile);
//   post(i,j,State&)
// with op(F,di,dj) = field F at (i+di,j+dj,k) and op.up(F) = field F at (i,j,k+1).
// The same functor runs on direct global loads (GlobalOp) in the host-emulated test build and
// when the layout rules out a tensor map (odd im: row pitch not a multiple of 16 bytes).
#pragma once
#include "pom_core.h"
#include <cstdlib>
#ifndef POMGPU_EMU
#include <cuda.h>
#endif

namespace pom {

POM_HD constexpr int tma_plane(int bw, int bh) { return ((bw * bh + 15) / 16) * 16; }   // 128-byte aligned planes

template <class F>
struct GlobalOp {
  const double* const* fld;
  Geo g;
  int i, j, k;
  POM_HD double operator()(int f, int di, int dj) const {
    const int ii = i + di, jj = j + dj;
    if (ii < 1 || ii > g.im || jj < g.joff + 1 || jj > g.joff + g.jml) return 0.;
    return fld[f][POM_I3(ii, jj, k)];
  }
  POM_HD double up(int f) const { return (k + 1 <= g.kb) ? fld[f][POM_I3(i, j, k + 1)] : 0.; }
  POM_HD double up(int f, int di, int dj) const {
    const int ii = i + di, jj = j + dj;
    if (k + 1 > g.kb || ii < 1 || ii > g.im || jj < g.joff + 1 || jj > g.joff + g.jml) return 0.;
    return fld[f][POM_I3(ii, jj, k + 1)];
  }
};

struct Tile2 {
  const double* s; int tx, ty, ny;   // ny = rows of the thread tile (F::TY)
  POM_HD double operator()(int v, int di, int dj) const { return s[(v * ny + (ty + dj)) * TILE_X + (tx + di)]; }
  POM_HD double operator()(int di, int dj) const { return s[(ty + dj) * TILE_X + (tx + di)]; }
};

#ifndef POMGPU_EMU
template <int NF> struct TmaMaps { CUtensorMap m[NF]; };
int tma_encode(Ctx* c, CUtensorMap* m, const double* base, int nk, int bw, int bh, int bk = 1);   // pom_state.cu (box bw x bh x bk)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

template <class F>
struct SmemOp {
  const double* cur;   // stage of level k, at this thread's own point
  const double* nxt;   // stage of level k+1
  static constexpr int PL = tma_plane(F::BW, F::BH);
  __device__ __forceinline__ double operator()(int f, int di, int dj) const { return cur[f * PL + dj * F::BW + di]; }
  __device__ __forceinline__ double up(int f) const { return nxt[f * PL]; }
  __device__ __forceinline__ double up(int f, int di, int dj) const { return nxt[f * PL + dj * F::BW + di]; }
};

// POM_TILE_PRODUCER: the TMA loads of the tile kernels are issued by one lane of an EXTRA warp (thread row
// ty == F::TY) instead of by thread (0,0) of the computing warps: it takes part in the per-level block barrier
// (which tells it that the stage of level k-1 is free) and then issues level k-1+NS while the computing warps
// are already in combine(k) -- the NF serial TMA issues leave the critical path of every level.
#ifndef POM_TILE_PRODUCER
#define POM_TILE_PRODUCER 1
#endif
template <class F>
__global__ void __launch_bounds__(TILE_X * (F::TY + POM_TILE_PRODUCER), F::MINB)
tmakernel(const __grid_constant__ TmaMaps<F::NF> maps, const F f, int i0, int i1, int j0, int j1) {
  constexpr int NF = F::NF, NS = F::NS, NV = F::NV, PL = tma_plane(F::BW, F::BH);
  constexpr int OX = TILE_X - F::HL - F::HR, OY = F::TY - F::HB - F::HT;
  extern __shared__ __align__(128) double pom_tsm[];
  double* ring = pom_tsm;                                  // [NS][NF][PL]
  double* S = ring + NS * NF * PL;                         // [2][NV][TY][TILE_X]
  uint64_t* bar = (uint64_t*)(S + 2 * NV * F::TY * TILE_X);   // [NS]
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ti0 = i0 + blockIdx.x * OX - F::HL, tj0 = j0 + blockIdx.y * OY - F::HB;
  const int i = ti0 + tx, j = tj0 + ty;
  // 0-based box origin in the arrays.  The TMA needs a 16-byte aligned start address, i.e. an
  // even i-origin (measured: an odd one raises "illegal instruction"); negative origins and
  // boxes that stick out of the array are fine (zero fill).  BW leaves room for the shift.
  static_assert(F::BW % 2 == 0 && F::BW >= TILE_X + F::OHL + F::OHR + 1, "TMA box too narrow");
  static_assert(F::BH >= F::TY + F::OHB + F::OHT, "TMA box too short");
  const int n0 = ti0 - 1 - F::OHL, shift = n0 & 1;
  const int c0 = n0 - shift, c1 = tj0 - 1 - f.g.joff - F::OHB;
  const int k0 = f.k0(), k1 = f.k1(), kl1 = f.kl1();
  const bool leader = POM_TILE_PRODUCER ? (tx == 0 && ty == F::TY) : (tx == 0 && ty == 0);
  if (leader) {
    for (int s = 0; s < NS; ++s) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int L) {
    const int s = (L - k0) % NS;
    mbar_expect_tx(&bar[s], (uint32_t)(NF * F::BW * F::BH * sizeof(double)));
#pragma unroll
    for (int n = 0; n < NF; ++n) tma_load_3d(ring + (s * NF + n) * PL, &maps.m[n], &bar[s], c0, c1, L - 1);
  };
  if (leader)
    for (int L = k0; L < k0 + NS && L <= kl1; ++L) issue(L);
  if (POM_TILE_PRODUCER && ty == F::TY) {                  // the producer warp: one barrier per level, like the others
    for (int k = k0; k <= k1; ++k) {
      __syncthreads();
      if (leader && k - 1 >= k0 && k - 1 + NS <= kl1) issue(k - 1 + NS);
    }
    return;
  }
  const bool inside = (i >= 1 && i <= f.g.im && j >= f.g.joff + 1 && j <= f.g.joff + f.g.jml);
  const bool out = (tx >= F::HL && tx < TILE_X - F::HR && ty >= F::HB && ty < F::TY - F::HT && i <= i1 && j <= j1);
  typename F::State st{};
  f.pre(i, j, inside, out, st);
  const int own = (ty + F::OHB) * F::BW + tx + F::OHL + shift;
  int buf = 0;
  for (int k = k0; k <= k1; ++k) {
    const int q0 = k - k0, s0 = q0 % NS;
    mbar_wait(&bar[s0], (q0 / NS) & 1);
    int s1 = s0;
    if (F::UP && k + 1 <= kl1) {
      s1 = (q0 + 1) % NS;
      mbar_wait(&bar[s1], ((q0 + 1) / NS) & 1);
    }
    const SmemOp<F> op{ring + s0 * NF * PL + own, ring + s1 * NF * PL + own};
    double v[NV];
    if (!F::FULL) {
#pragma unroll
      for (int n = 0; n < NV; ++n) v[n] = 0.;
    }
    if (F::FULL || inside) f.stage(i, j, k, st, op, v);
    double* Sb = S + buf * NV * F::TY * TILE_X;
#pragma unroll
    for (int n = 0; n < NV; ++n) Sb[(n * F::TY + ty) * TILE_X + tx] = v[n];
    __syncthreads();
    // every thread is past combine(k-1): the stage of level k-1 is free for level k-1+NS
    if (!POM_TILE_PRODUCER && leader && k - 1 >= k0 && k - 1 + NS <= kl1) issue(k - 1 + NS);
    if (out) f.combine(i, j, k, st, op, Tile2{Sb, tx, ty, F::TY});
    buf ^= 1;
  }
  if (out) f.post(i, j, st);
}

#endif   // !POMGPU_EMU

// ---- per-thread column vectors (the eliminated Thomas coefficients ee/gg) ---------------------
// A column functor declares NVEC vectors of kb doubles per thread and uses them only through
// put(v,k,x) / get(v,k).  They live in per-thread local memory (dynamically indexed), so that the
// scalars of the functor's State stay in registers.  (Parking them in an L2-resident global scratch
// instead -- evict_last stores, discard.global.L2 after the upward sweep, a persisting-L2 set-aside --
// does remove their HBM round trip, ncu: advu_profu 4.0 -> 2.9 GB, but not a microsecond of run time;
// measured in round 2, DESIGN.md section 3, profiles/r2_colscratch_experiments.txt.)
template <int NVEC>
struct LocalCols {
  double a[NVEC][KMAX];
  POM_HD void put(int v, int k, double x) { a[v][k] = x; }
  POM_HD double get(int v, int k) const { return a[v][k]; }
};

#ifndef POMGPU_EMU
// Column kernel on the TMA ring: no flux exchange between threads (thread tile = output tile),
// only the operand staging.  F provides NF, NS, TY, OHL/OHR/OHB/OHT, BW, BH, UP, NK, NVEC,
// fields(), k0(), k1(), kl1(), pre(i,j,State&,CM&), level(i,j,k,State&,CM&,op),
// post(i,j,State&,CM&) with CM = the column vectors above.
// `post` runs the upward sweeps (Thomas back-substitution) on the thread's own column.
template <class F>
__global__ void __launch_bounds__(TILE_X * F::TY, F::MINB)
tmacolkernel(const __grid_constant__ TmaMaps<F::NF> maps, const F f, int i0, int i1, int j0, int j1) {
  constexpr int NF = F::NF, NS = F::NS, PL = tma_plane(F::BW, F::BH);
  extern __shared__ __align__(128) double pom_tsm[];
  double* ring = pom_tsm;
  uint64_t* bar = (uint64_t*)(ring + NS * NF * PL);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ti0 = i0 + blockIdx.x * TILE_X, tj0 = j0 + blockIdx.y * F::TY;
  const int i = ti0 + tx, j = tj0 + ty;
  static_assert(F::BW % 2 == 0 && F::BW >= TILE_X + F::OHL + F::OHR + 1, "TMA box too narrow");
  static_assert(F::BH >= F::TY + F::OHB + F::OHT, "TMA box too short");
  const int n0 = ti0 - 1 - F::OHL, shift = n0 & 1;
  const int c0 = n0 - shift, c1 = tj0 - 1 - f.g.joff - F::OHB;
  const int k0 = f.k0(), k1 = f.k1(), kl1 = f.kl1();
  const bool leader = (tx == 0 && ty == 0);
  if (leader) {
    for (int s = 0; s < NS; ++s) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int L) {
    const int s = (L - k0) % NS;
    mbar_expect_tx(&bar[s], (uint32_t)(NF * F::BW * F::BH * sizeof(double)));
#pragma unroll
    for (int n = 0; n < NF; ++n)
      tma_load_3d(ring + (s * NF + n) * PL, &maps.m[n], &bar[s], c0, c1, L - 1);
  };
  if (leader)
    for (int L = k0; L < k0 + NS && L <= kl1; ++L) issue(L);
  const bool active = (i <= i1 && j <= j1);
  typename F::State st;        // scalars: registers
  LocalCols<F::NVEC> cm;       // per-thread column vectors (dynamically indexed: local memory)
  if (active) f.pre(i, j, st, cm);
  const int own = (ty + F::OHB) * F::BW + tx + F::OHL + shift;
  for (int k = k0; k <= k1; ++k) {
    const int q0 = k - k0, s0 = q0 % NS;
    mbar_wait(&bar[s0], (q0 / NS) & 1);
    int s1 = s0;
    if (F::UP && k + 1 <= kl1) {
      s1 = (q0 + 1) % NS;
      mbar_wait(&bar[s1], ((q0 + 1) / NS) & 1);
    }
    const SmemOp<F> op{ring + s0 * NF * PL + own, ring + s1 * NF * PL + own};
    if (active) f.level(i, j, k, st, cm, op);
    __syncthreads();
    if (leader && k + NS <= kl1) issue(k + NS);   // everyone is done with the stage of level k
  }
  if (active) f.post(i, j, st, cm);
}

template <class F>
__global__ void __launch_bounds__(TILE_X * F::TY, F::MINB)
colkernel_g(const F f, int i0, int i1, int j0, int j1) {
  const int i = i0 + blockIdx.x * TILE_X + threadIdx.x, j = j0 + blockIdx.y * F::TY + threadIdx.y;
  if (i > i1 || j > j1) return;
  const double* fld[F::NF];
  f.fields(fld);
  typename F::State st;
  LocalCols<F::NVEC> cm;
  f.pre(i, j, st, cm);
  const int k1 = f.k1();
  for (int k = f.k0(); k <= k1; ++k) f.level(i, j, k, st, cm, GlobalOp<F>{fld, f.g, i, j, k});
  f.post(i, j, st, cm);
}

// the same functor on direct global loads (layouts the TMA cannot address)
template <class F>
__global__ void __launch_bounds__(TILE_X * F::TY, F::MINB)
tilekernel_g(const F f, int i0, int i1, int j0, int j1) {
  constexpr int OX = TILE_X - F::HL - F::HR, OY = F::TY - F::HB - F::HT, NV = F::NV;
  __shared__ double S[NV * F::TY * TILE_X];   // single buffer (static shared memory is capped at 48 kB)
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = i0 + blockIdx.x * OX - F::HL + tx, j = j0 + blockIdx.y * OY - F::HB + ty;
  const bool inside = (i >= 1 && i <= f.g.im && j >= f.g.joff + 1 && j <= f.g.joff + f.g.jml);
  const bool out = (tx >= F::HL && tx < TILE_X - F::HR && ty >= F::HB && ty < F::TY - F::HT && i <= i1 && j <= j1);
  const double* fld[F::NF];
  f.fields(fld);
  typename F::State st{};
  f.pre(i, j, inside, out, st);
  const int k1 = f.k1();
  for (int k = f.k0(); k <= k1; ++k) {
    const GlobalOp<F> op{fld, f.g, i, j, k};
    double v[NV];
#pragma unroll
    for (int n = 0; n < NV; ++n) v[n] = 0.;
    if (inside) f.stage(i, j, k, st, op, v);
#pragma unroll
    for (int n = 0; n < NV; ++n) S[(n * F::TY + ty) * TILE_X + tx] = v[n];
    __syncthreads();
    if (out) f.combine(i, j, k, st, op, Tile2{S, tx, ty, F::TY});
    __syncthreads();
  }
  if (out) f.post(i, j, st);
}
#endif

template <class F>
inline void launch_tma_tiles(Ctx* c, const F& f, int i0, int i1, int j0, int j1) {
  if (i1 < i0 || j1 < j0) return;
  c->launches++;
  constexpr int OX = TILE_X - F::HL - F::HR, OY = F::TY - F::HB - F::HT;
  const int nbx = (i1 - i0 + OX) / OX, nby = (j1 - j0 + OY) / OY;
#ifdef POMGPU_EMU
  static typename F::State st[TILE_Y][TILE_X];
  static double S[F::NV * TILE_Y * TILE_X];
  static bool ins[TILE_Y][TILE_X], outm[TILE_Y][TILE_X];
  const double* fld[F::NF];
  f.fields(fld);
  for (int by = 0; by < nby; ++by)
    for (int bx = 0; bx < nbx; ++bx) {
      POM_TILE_LOOP {
        POM_TILE_IJ;
        ins[ty][tx] = (i >= 1 && i <= f.g.im && j >= f.g.joff + 1 && j <= f.g.joff + f.g.jml);
        outm[ty][tx] = (tx >= F::HL && tx < TILE_X - F::HR && ty >= F::HB && ty < F::TY - F::HT && i <= i1 && j <= j1);
        f.pre(i, j, ins[ty][tx], outm[ty][tx], st[ty][tx]);
      }
      for (int k = f.k0(); k <= f.k1(); ++k) {
        POM_TILE_LOOP {
          POM_TILE_IJ;
          double v[F::NV];
          for (int n = 0; n < F::NV; ++n) v[n] = 0.;
          if (ins[ty][tx]) f.stage(i, j, k, st[ty][tx], GlobalOp<F>{fld, f.g, i, j, k}, v);
          for (int n = 0; n < F::NV; ++n) S[(n * F::TY + ty) * TILE_X + tx] = v[n];
        }
        POM_TILE_LOOP {
          POM_TILE_IJ;
          if (outm[ty][tx]) f.combine(i, j, k, st[ty][tx], GlobalOp<F>{fld, f.g, i, j, k}, Tile2{S, tx, ty, F::TY});
        }
      }
      POM_TILE_LOOP {
        POM_TILE_IJ;
        if (outm[ty][tx]) f.post(i, j, st[ty][tx]);
      }
    }
#else
  if (c->prof_on) {
    const KInfo& k = F::info();
    double cols = (double)(i1 - i0 + 1) * (j1 - j0 + 1);
    prof_before(c, &k, 8. * cols * ((k.r3 + k.w3) * (double)c->g.kb + (k.r2 + k.w2)));
  }
  cudaSetDevice(c->device);
  dim3 b(TILE_X, F::TY), gr(nbx, nby);
  bool tma_ok = (c->g.im % 2 == 0) && !c->no_tma;
  if (tma_ok) {
    TmaMaps<F::NF> maps;
    const double* fld[F::NF];
    f.fields(fld);
    for (int n = 0; n < F::NF && tma_ok; ++n)
      if (tma_encode(c, &maps.m[n], fld[n], F::NK ? F::NK : c->g.kb, F::BW, F::BH)) tma_ok = false;
    if (tma_ok) {
      constexpr size_t smem = (size_t)(F::NS * F::NF * tma_plane(F::BW, F::BH) + 2 * F::NV * F::TY * TILE_X) * sizeof(double) + F::NS * 8;
      static DevOnce granted;
      if (granted.need(c->device)) cudaFuncSetAttribute(tmakernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      const dim3 bt(TILE_X, F::TY + POM_TILE_PRODUCER);     // (+ the producer warp)
      tmakernel<F><<<gr, bt, smem, (cudaStream_t)c->stream>>>(maps, f, i0, i1, j0, j1);
    }
  }
  if (!tma_ok) tilekernel_g<F><<<gr, b, 0, (cudaStream_t)c->stream>>>(f, i0, i1, j0, j1);
  if (c->prof_on) prof_after(c);
#endif
}

// ---- single-level tile kernel with TWO neighbour exchanges (the fused external substep) ----
// phaseA: own-point flux-like values v[NV] -> shared tile S; phaseB: one derived value per
// point from the neighbours' S (-> shared tile E) plus private results kept in State;
// phaseC: outputs from S, E and the staged operands.  F provides NF, NV, TY, OH*, BW, BH, NK,
// fields(), pre(i,j,inside,State&), phaseA(i,j,State&,op,v), phaseB(i,j,State&,op,Tile2 S) -> double,
// phaseC(i,j,State&,op,Tile2 S,Tile2 E).  Thread-tile halo is 1 on every side: phase B runs
// on tx<=30, ty<=14, phase C (the outputs) on 1<=tx<=30, 1<=ty<=14.
#ifndef POMGPU_EMU
template <class F, bool TMA>
__global__ void __launch_bounds__(TILE_X * F::TY, F::MINB)
tile3kernel(const __grid_constant__ TmaMaps<F::NF> maps, const F f, int i0, int i1, int j0, int j1) {
  constexpr int NF = F::NF, NV = F::NV, PL = tma_plane(F::BW, F::BH);
  constexpr int OX = TILE_X - 2, OY = F::TY - 2;
  extern __shared__ __align__(128) double pom_tsm[];
  double* ring = pom_tsm;                         // [NF][PL]   (TMA only)
  double* S = ring + (TMA ? NF * PL : 0);         // [NV][TY][TILE_X]
  double* E = S + NV * F::TY * TILE_X;           // [TILE_Y][TILE_X]
  uint64_t* bar = (uint64_t*)(E + F::TY * TILE_X);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ti0 = i0 + blockIdx.x * OX - 1, tj0 = j0 + blockIdx.y * OY - 1;
  const int i = ti0 + tx, j = tj0 + ty;
  const int n0 = ti0 - 1 - F::OHL, shift = n0 & 1;
  const bool inside = (i >= 1 && i <= f.g.im && j >= f.g.joff + 1 && j <= f.g.joff + f.g.jml);
  const bool out = (tx >= 1 && tx <= TILE_X - 2 && ty >= 1 && ty <= F::TY - 2 && i <= i1 && j <= j1 && inside);
  const bool bok = (tx <= TILE_X - 2 && ty <= F::TY - 2 && inside);
  typename F::State st;
  if (TMA) {
    static_assert(F::BW % 2 == 0 && F::BW >= TILE_X + F::OHL + F::OHR + 1, "TMA box too narrow");
    static_assert(F::BH >= F::TY + F::OHB + F::OHT, "TMA box too short");
    const bool leader = (tx == 0 && ty == 0);
    if (leader) {
      // two transactions: the first F::NFA fields are all that phases A and B read, so they start
      // while the operands only phase C needs are still in flight
      mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const int c0 = n0 - shift, c1 = tj0 - 1 - f.g.joff - F::OHB;
      mbar_expect_tx(&bar[0], (uint32_t)(F::NFA * F::BW * F::BH * sizeof(double)));
#pragma unroll
      for (int n = 0; n < F::NFA; ++n) tma_load_3d(ring + n * PL, &maps.m[n], &bar[0], c0, c1, 0);
      mbar_expect_tx(&bar[1], (uint32_t)((NF - F::NFA) * F::BW * F::BH * sizeof(double)));
#pragma unroll
      for (int n = F::NFA; n < NF; ++n) tma_load_3d(ring + n * PL, &maps.m[n], &bar[1], c0, c1, 0);
    }
    f.pre(i, j, inside, st);        // plain loads of the point-wise operands overlap the TMA
    __syncthreads();                // barrier init visible to every waiter
    mbar_wait(&bar[0], 0);
  } else {
    f.pre(i, j, inside, st);
  }
  const double* fld[NF];
  if (!TMA) f.fields(fld);
  const int own = (ty + F::OHB) * F::BW + tx + F::OHL + shift;
  const SmemOp<F> sop{ring + own, ring + own};
  const GlobalOp<F> gop{fld, f.g, i, j, 1};
  double v[NV];
#pragma unroll
  for (int n = 0; n < NV; ++n) v[n] = 0.;
  if (inside) { if (TMA) f.phaseA(i, j, st, sop, v); else f.phaseA(i, j, st, gop, v); }
#pragma unroll
  for (int n = 0; n < NV; ++n) S[(n * F::TY + ty) * TILE_X + tx] = v[n];
  __syncthreads();
  double e = 0.;
  if (bok) e = TMA ? f.phaseB(i, j, st, sop, Tile2{S, tx, ty, F::TY}, tx >= 1 && ty >= 1) : f.phaseB(i, j, st, gop, Tile2{S, tx, ty, F::TY}, tx >= 1 && ty >= 1);
  E[ty * TILE_X + tx] = e;
  if (TMA) mbar_wait(&bar[1], 0);
  __syncthreads();
  if (out) { if (TMA) f.phaseC(i, j, st, sop, Tile2{S, tx, ty, F::TY}, Tile2{E, tx, ty, F::TY}); else f.phaseC(i, j, st, gop, Tile2{S, tx, ty, F::TY}, Tile2{E, tx, ty, F::TY}); }
}
#endif

template <class F>
inline void launch_tile3(Ctx* c, const F& f, int i0, int i1, int j0, int j1) {
  if (i1 < i0 || j1 < j0) return;
  c->launches++;
  constexpr int OX = TILE_X - 2, OY = F::TY - 2;
  const int nbx = (i1 - i0 + OX) / OX, nby = (j1 - j0 + OY) / OY;
#ifdef POMGPU_EMU
  static typename F::State st[TILE_Y][TILE_X];
  static double S[F::NV * F::TY * TILE_X], E[F::TY * TILE_X];
  static bool ins[TILE_Y][TILE_X];
  const double* fld[F::NF];
  f.fields(fld);
  for (int by = 0; by < nby; ++by)
    for (int bx = 0; bx < nbx; ++bx) {
#define POM_T3_LOOP for (int ty = 0; ty < F::TY; ++ty) for (int tx = 0; tx < TILE_X; ++tx)
#define POM_T3_IJ const int i = i0 + bx * OX - 1 + tx, j = j0 + by * OY - 1 + ty
      POM_T3_LOOP {
        POM_T3_IJ;
        ins[ty][tx] = (i >= 1 && i <= f.g.im && j >= f.g.joff + 1 && j <= f.g.joff + f.g.jml);
        f.pre(i, j, ins[ty][tx], st[ty][tx]);
        double v[F::NV];
        for (int n = 0; n < F::NV; ++n) v[n] = 0.;
        if (ins[ty][tx]) f.phaseA(i, j, st[ty][tx], GlobalOp<F>{fld, f.g, i, j, 1}, v);
        for (int n = 0; n < F::NV; ++n) S[(n * F::TY + ty) * TILE_X + tx] = v[n];
      }
      POM_T3_LOOP {
        POM_T3_IJ;
        double e = 0.;
        if (tx <= TILE_X - 2 && ty <= F::TY - 2 && ins[ty][tx])
          e = f.phaseB(i, j, st[ty][tx], GlobalOp<F>{fld, f.g, i, j, 1}, Tile2{S, tx, ty, F::TY}, tx >= 1 && ty >= 1);
        E[ty * TILE_X + tx] = e;
      }
      POM_T3_LOOP {
        POM_T3_IJ;
        if (tx >= 1 && tx <= TILE_X - 2 && ty >= 1 && ty <= F::TY - 2 && i <= i1 && j <= j1 && ins[ty][tx])
          f.phaseC(i, j, st[ty][tx], GlobalOp<F>{fld, f.g, i, j, 1}, Tile2{S, tx, ty, F::TY}, Tile2{E, tx, ty, F::TY});
      }
    }
#else
  if (c->prof_on) {
    const KInfo& k = F::info();
    double cols = (double)(i1 - i0 + 1) * (j1 - j0 + 1);
    prof_before(c, &k, 8. * cols * ((k.r3 + k.w3) * (double)c->g.kb + (k.r2 + k.w2)));
  }
  cudaSetDevice(c->device);
  dim3 b(TILE_X, F::TY), gr(nbx, nby);
  TmaMaps<F::NF> maps;
  bool tma_ok = (c->g.im % 2 == 0) && !c->no_tma;
  if (tma_ok) {
    const double* fld[F::NF];
    f.fields(fld);
    for (int n = 0; n < F::NF && tma_ok; ++n)
      if (tma_encode(c, &maps.m[n], fld[n], 1, F::BW, F::BH)) tma_ok = false;
  }
  constexpr size_t sm_se = (size_t)((F::NV + 1) * F::TY * TILE_X) * sizeof(double) + 16;
  constexpr size_t sm_tma = sm_se + (size_t)F::NF * tma_plane(F::BW, F::BH) * sizeof(double);
  static DevOnce granted;
  if (granted.need(c->device)) {
    cudaFuncSetAttribute(tile3kernel<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_tma);
    cudaFuncSetAttribute(tile3kernel<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_se);
  }
  if (tma_ok) tile3kernel<F, true><<<gr, b, sm_tma, (cudaStream_t)c->stream>>>(maps, f, i0, i1, j0, j1);
  else tile3kernel<F, false><<<gr, b, sm_se, (cudaStream_t)c->stream>>>(maps, f, i0, i1, j0, j1);
  if (c->prof_on) prof_after(c);
#endif
}

template <class F>
inline void launch_tma_cols(Ctx* c, const F& f, int i0, int i1, int j0, int j1) {
  if (i1 < i0 || j1 < j0) return;
  c->launches++;
#ifdef POMGPU_EMU
  const double* fld[F::NF];
  f.fields(fld);
  static typename F::State st;
  static LocalCols<F::NVEC> cm;
  for (int j = j0; j <= j1; ++j)
    for (int i = i0; i <= i1; ++i) {
      f.pre(i, j, st, cm);
      for (int k = f.k0(); k <= f.k1(); ++k) f.level(i, j, k, st, cm, GlobalOp<F>{fld, f.g, i, j, k});
      f.post(i, j, st, cm);
    }
#else
  if (c->prof_on) {
    const KInfo& k = F::info();
    double cols = (double)(i1 - i0 + 1) * (j1 - j0 + 1);
    prof_before(c, &k, 8. * cols * ((k.r3 + k.w3) * (double)c->g.kb + (k.r2 + k.w2)));
  }
  cudaSetDevice(c->device);
  dim3 b(TILE_X, F::TY), gr((i1 - i0 + TILE_X) / TILE_X, (j1 - j0 + F::TY) / F::TY);
  bool tma_ok = (c->g.im % 2 == 0) && !c->no_tma;
  if (tma_ok) {
    TmaMaps<F::NF> maps;
    const double* fld[F::NF];
    f.fields(fld);
    for (int n = 0; n < F::NF && tma_ok; ++n)
      if (tma_encode(c, &maps.m[n], fld[n], F::NK ? F::NK : c->g.kb, F::BW, F::BH)) tma_ok = false;
    if (tma_ok) {
      constexpr size_t smem = (size_t)(F::NS * F::NF * tma_plane(F::BW, F::BH)) * sizeof(double) + F::NS * 8;
      static DevOnce granted;
      if (granted.need(c->device)) cudaFuncSetAttribute(tmacolkernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      tmacolkernel<F><<<gr, b, smem, (cudaStream_t)c->stream>>>(maps, f, i0, i1, j0, j1);
    }
  }
  if (!tma_ok) colkernel_g<F><<<gr, b, 0, (cudaStream_t)c->stream>>>(f, i0, i1, j0, j1);
  if (c->prof_on) prof_after(c);
#endif
}


// ---- two-sweep column kernel with the first sweep's results PARKED in shared memory ----------
// For column operations whose second sweep needs a depth sum of the first (the u, v Asselin filter with
// its depth-mean removal, advance.f:469-514): the operands are read from HBM ONCE.  PERSISTENT,
// WARP-SPECIALISED blocks, one per SM: TY consumer warps own 32 x TY columns of ONE component per tile,
// one producer warp feeds the NF operand planes of every level through a TMA ring of `ns` stages (as deep
// as the shared memory left over allows) guarded by full/empty mbarrier pairs -- no __syncthreads in the
// loop, and the producer runs ahead into the NEXT tile while the consumers are still in sweep 2 of the
// current one, so the loads never drain.  Sweep 1 reduces the staged operands and parks NP values per level
// and column in shared memory ([level][value][thread]: thread-private slots, conflict-free), sweep 2
// combines the parked values with the column sums and streams the results out.  F provides NC, NF, NP,
// fields(comp, b), ktab() (a k-only table, handed to sweep1 level by level), State, NPL + preload(comp,i,j,pl) (the
// column's plain loads, issued one tile ahead), pre(comp,i,j,State&,pl) (no memory access; also runs for columns
// outside the rectangle, whose results are never stored), fin(comp,i,j,State&,pl),
// sweep1(comp,i,j,k,State&,const double* f,double* pk,double tk),
// mid(comp,i,j,State&), sweep2(comp,i,j,k,State&,const double* pk), fin(comp,i,j,State&); levels 1..kb-1;
// and NSIDE extra warps per block that run side(w, nw, lane) = the w-th of nw shares of whatever columns the
// rectangle leaves over (the Orlanski frame of the filter), concurrently with the streaming warps.
#ifndef POMGPU_EMU
template <class F, int TY, int KL>
__global__ void __launch_bounds__(TILE_X * TY + 32 + 32 * F::NSIDE, 1)
tmaparkkernel(const __grid_constant__ TmaMaps<F::NC * F::NF> maps, const F f, int i0, int i1, int j0, int j1, int ns,
              int nbx, int ntiles) {
  constexpr int NF = F::NF, NP = F::NP, NC = F::NC, NT = TILE_X * TY;   // box = thread tile x KL levels
  constexpr int STG = NF * KL * NT;                                    // doubles per stage: [field][level][thread]
  extern __shared__ __align__(128) double pom_tsm[];
  const int nk = f.g.kb - 1, nst = (nk + KL - 1) / KL;                 // stages per tile
  double* ring = pom_tsm;                        // [ns][NF][KL][NT]
  double* park = ring + ns * STG;                // [nk][NP][NT]
  uint64_t* full = (uint64_t*)(park + nk * NP * NT);   // [ns] stage filled (TMA complete_tx)
  uint64_t* empty = full + ns;                         // [ns] stage released by the TY consumer warps
  double* kt = (double*)(empty + ns);                  // [nk] the functor's k-only table (dz), one LDS instead of a global load per level
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < ns; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], TY); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int k = tid; k < nk; k += blockDim.x) kt[k] = f.ktab()[k];
  __syncthreads();
  if (warp > TY) {                               // ---- side warps: the functor's left-over columns, hidden behind the stream ----
    f.side(blockIdx.x * F::NSIDE + (warp - TY - 1), gridDim.x * F::NSIDE, lane);
    return;
  }
  // ring position: stage s, parity of its current use (no integer division in the loops)
  int s = 0, ph = 0;
  if (warp == TY) {                              // ---- producer warp: lane n issues the box of field n ----
    if (lane >= NF) return;
    bool first = true;                           // first pass over the ring: the stages are free
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int bx = t % nbx, r = t / nbx, comp = r % NC, by = r / NC;
      const int c0 = i0 + bx * TILE_X - 1, c1 = j0 + by * TY - 1 - f.g.joff;   // even i-origin: the launcher checks i0
      const CUtensorMap* m = &maps.m[comp * NF + lane];
      for (int q = 0; q < nst; ++q) {
        if (lane == 0) {
          if (!first) mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], (uint32_t)(STG * sizeof(double)));
        }
        __syncwarp((1u << NF) - 1);
        tma_load_3d(ring + s * STG + lane * KL * NT, m, &full[s], c0, c1, q * KL);
        if (++s == ns) { s = 0; ph ^= 1; first = false; }
      }
    }
    return;
  }
  // the column's own 2-D / bottom-level operands (F::NPL plain loads) are fetched ONE TILE AHEAD, while the
  // previous tile is in its second sweep, so that no tile starts or ends on an exposed HBM round trip
  double pl[F::NPL], pln[F::NPL];
#pragma unroll
  for (int n = 0; n < F::NPL; ++n) pln[n] = 0.;
  {
    const int t = blockIdx.x;
    if (t < ntiles) {
      const int bx = t % nbx, r = t / nbx, comp = r % NC, by = r / NC;
      const int i = i0 + bx * TILE_X + lane, j = j0 + by * TY + warp;
      if (i <= i1 && j <= j1) f.preload(comp, i, j, pln);
    }
  }
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int bx = t % nbx, r = t / nbx, comp = r % NC, by = r / NC;
    const int i = i0 + bx * TILE_X + lane, j = j0 + by * TY + warp;
    const bool active = (i <= i1 && j <= j1);
#pragma unroll
    for (int n = 0; n < F::NPL; ++n) pl[n] = pln[n];
    typename F::State st;
    f.pre(comp, i, j, st, pl);
    double* pp = park + tid;
    for (int q = 0; q < nst; ++q) {
      mbar_wait(&full[s], ph);
      const double* rp = ring + s * STG + tid;
      double v[KL][NF], kk[KL];
#pragma unroll
      for (int l = 0; l < KL; ++l) {
#pragma unroll
        for (int n = 0; n < NF; ++n) v[l][n] = rp[(n * KL + l) * NT];
        kk[l] = kt[min(q * KL + l, nk - 1)];
      }
#pragma unroll
      for (int l = 0; l < KL; ++l) {
        const int k = q * KL + l + 1;
        if (k <= nk) {
          double pk[NP];
          f.sweep1(comp, i, j, k, st, v[l], pk, kk[l]);
#pragma unroll
          for (int n = 0; n < NP; ++n) pp[n * NT] = pk[n];
          pp += NP * NT;
        }
      }
      // release the stage only after the values read from it have been USED (the loads have landed)
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == ns) { s = 0; ph ^= 1; }
    }
    {
      const int tn = t + gridDim.x;
      if (tn < ntiles) {
        const int bxn = tn % nbx, rn = tn / nbx, compn = rn % NC, byn = rn / NC;
        const int in = i0 + bxn * TILE_X + lane, jn = j0 + byn * TY + warp;
        if (in <= i1 && jn <= j1) f.preload(compn, in, jn, pln);
      }
    }
    if (active) {
      f.mid(comp, i, j, st);
      pp = park + tid;
#pragma unroll 4
      for (int k = 1; k <= nk; ++k) {
        double pk[NP];
#pragma unroll
        for (int n = 0; n < NP; ++n) pk[n] = pp[n * NT];
        pp += NP * NT;
        f.sweep2(comp, i, j, k, st, pk);
      }
      f.fin(comp, i, j, st, pl);
    }
  }
}

inline int park_bps() {   // blocks per SM (tuning knob; 1 = one block owns the SM's whole shared memory)
  static const int b = getenv("POMGPU_PARK_BPS") ? atoi(getenv("POMGPU_PARK_BPS")) : 1;
  return b < 1 ? 1 : (b > 8 ? 8 : b);
}
template <class F, int TY>
inline size_t park_ring_bytes(const Ctx* c) {       // shared memory left for the ring with 32 x TY columns parked (0: does not fit)
  const size_t park = (size_t)(c->g.kb - 1) * F::NP * TILE_X * TY * sizeof(double), stage = (size_t)F::NF * TILE_X * TY * sizeof(double) + 16;
  // 228 kB per SM shared by the resident blocks, 1 kB of each reserved by the driver; less the k-table
  const size_t cap = (233472 / park_bps() - 1024 - 512) & ~(size_t)127;
  return park + 4 * stage > cap ? 0 : cap - park;
}
template <class F, int TY, int KL>
inline void launch_park_ty(Ctx* c, const F& f, const TmaMaps<F::NC * F::NF>& maps, int i0, int i1, int j0, int j1) {
  constexpr int NT = TILE_X * TY;
  const size_t park = (size_t)(c->g.kb - 1) * F::NP * NT * sizeof(double), stage = (size_t)F::NF * KL * NT * sizeof(double) + 16;
  int ns = (int)(park_ring_bytes<F, TY>(c) / stage);
  if (ns > 32) ns = 32;
  const size_t smem = park + ns * stage + (size_t)(c->g.kb - 1) * sizeof(double);
  static DevOnce granted;
  if (granted.need(c->device)) cudaFuncSetAttribute(tmaparkkernel<F, TY, KL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  const int nbx = (i1 - i0 + TILE_X) / TILE_X, nby = (j1 - j0 + TY) / TY, ntiles = nbx * nby * F::NC;
  const int nres = c->nsm * park_bps();
  const int nblk = ntiles < nres ? ntiles : nres;
  tmaparkkernel<F, TY, KL><<<nblk, NT + 32 + 32 * F::NSIDE, smem, (cudaStream_t)c->stream>>>(maps, f, i0, i1, j0, j1, ns, nbx, ntiles);
}
template <class F, int TY>
inline void launch_park_kl(Ctx* c, const F& f, const TmaMaps<F::NC * F::NF>& maps, int i0, int i1, int j0, int j1, int kl) {
  if (kl == 4) launch_park_ty<F, TY, 4>(c, f, maps, i0, i1, j0, j1);
  else if (kl == 2) launch_park_ty<F, TY, 2>(c, f, maps, i0, i1, j0, j1);
  else launch_park_ty<F, TY, 1>(c, f, maps, i0, i1, j0, j1);
}
#endif

// returns false (nothing launched) when the layout rules the TMA path out: the caller falls back
template <class F>
inline bool launch_tma_park(Ctx* c, const F& f, int i0, int i1, int j0, int j1) {
  if (i1 < i0 || j1 < j0) return true;
#ifdef POMGPU_EMU
  c->launches++;
  static double park[KMAX * F::NP];
  typename F::State st;
  const int nk = c->g.kb - 1;
  for (int comp = 0; comp < F::NC; ++comp) {
    const double* fld[F::NF];
    f.fields(comp, fld);
    for (int j = j0; j <= j1; ++j)
      for (int i = i0; i <= i1; ++i) {
        const Geo& g = f.g;
        double pl[F::NPL];
        f.preload(comp, i, j, pl);
        f.pre(comp, i, j, st, pl);
        for (int k = 1; k <= nk; ++k) {
          double v[F::NF];
          for (int n = 0; n < F::NF; ++n) v[n] = fld[n][POM_I3(i, j, k)];
          f.sweep1(comp, i, j, k, st, v, park + (k - 1) * F::NP, f.ktab()[k - 1]);
        }
        f.mid(comp, i, j, st);
        for (int k = 1; k <= nk; ++k) f.sweep2(comp, i, j, k, st, park + (k - 1) * F::NP);
        f.fin(comp, i, j, st, pl);
      }
  }
  f.side_host();
  return true;
#else
  if ((c->g.im % 2) || c->no_tma || !(i0 & 1)) return false;   // box origin i0-1 must be even
  // thread tile: the tallest whose ring still holds ~90 kB in flight per SM (what the HBM latency needs at
  // full rate); else the tallest that fits at all
  const size_t r8 = park_ring_bytes<F, 8>(c), r6 = park_ring_bytes<F, 6>(c), r4 = park_ring_bytes<F, 4>(c), r2 = park_ring_bytes<F, 2>(c);
  const size_t want = 90 * 1024;
  int ty = r8 >= want ? 8 : r6 >= want ? 6 : r4 >= want ? 4 : r8 ? 8 : r6 ? 6 : r4 ? 4 : r2 ? 2 : 0;
  static const int force = getenv("POMGPU_PARK_TY") ? atoi(getenv("POMGPU_PARK_TY")) : 0;   // tuning experiments
  if (force == 8 && r8) ty = 8; else if (force == 6 && r6) ty = 6; else if (force == 4 && r4) ty = 4; else if (force == 2 && r2) ty = 2;
  if (!ty) return false;
  static const int fkl = getenv("POMGPU_PARK_KL") ? atoi(getenv("POMGPU_PARK_KL")) : 0;
  int kl = (fkl == 1 || fkl == 2 || fkl == 4) ? fkl : 4;   // levels per TMA box / ring stage (measured: 4 > 2 > 1)
  {   // at least 3 stages in the ring
    const size_t ring = ty == 8 ? r8 : ty == 6 ? r6 : ty == 4 ? r4 : r2;
    while (kl > 1 && ring < 3 * ((size_t)F::NF * kl * TILE_X * ty * sizeof(double) + 16)) kl /= 2;
  }
  TmaMaps<F::NC * F::NF> maps;
  for (int comp = 0; comp < F::NC; ++comp) {
    const double* fld[F::NF];
    f.fields(comp, fld);
    for (int n = 0; n < F::NF; ++n)
      if (tma_encode(c, &maps.m[comp * F::NF + n], fld[n], c->g.kb, TILE_X, ty, kl)) return false;
  }
  c->launches++;
  if (c->prof_on) {
    const KInfo& k = F::info();
    prof_before(c, &k, 8. * f.columns(i0, i1, j0, j1) * ((k.r3 + k.w3) * (double)c->g.kb + (k.r2 + k.w2)));
  }
  cudaSetDevice(c->device);
  if (ty == 8) launch_park_kl<F, 8>(c, f, maps, i0, i1, j0, j1, kl);
  else if (ty == 6) launch_park_kl<F, 6>(c, f, maps, i0, i1, j0, j1, kl);
  else if (ty == 4) launch_park_kl<F, 4>(c, f, maps, i0, i1, j0, j1, kl);
  else launch_park_kl<F, 2>(c, f, maps, i0, i1, j0, j1, kl);
  if (c->prof_on) prof_after(c);
  return true;
#endif
}

}  // namespace pom
