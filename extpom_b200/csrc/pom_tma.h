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
int tma_encode(Ctx* c, CUtensorMap* m, const double* base, int nk, int bw, int bh);   // pom_state.cu

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
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

// the same load with an L2 eviction-priority hint (createpolicy): the operand planes of the column
// kernels are streamed once, so they are marked evict_first and leave the L2 to the column scratch
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_3d_hint(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(pol) : "memory");
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

template <class F>
__global__ void __launch_bounds__(TILE_X * F::TY, F::MINB)
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
    for (int n = 0; n < NF; ++n) tma_load_3d(ring + (s * NF + n) * PL, &maps.m[n], &bar[s], c0, c1, L - 1);
  };
  if (leader)
    for (int L = k0; L < k0 + NS && L <= kl1; ++L) issue(L);
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
    if (leader && k - 1 >= k0 && k - 1 + NS <= kl1) issue(k - 1 + NS);
    if (out) f.combine(i, j, k, st, op, Tile2{Sb, tx, ty, F::TY});
    buf ^= 1;
  }
  if (out) f.post(i, j, st);
}

#endif   // !POMGPU_EMU

// ---- per-thread column vectors (the eliminated Thomas coefficients ee/gg) ---------------------
// A column functor declares NVEC vectors of kb doubles per thread and uses them only through
// put(v,k,x) / get(v,k).  LocalCols keeps them in per-thread local memory (host-emulated build and
// the direct-load fallback).  ScratchCols keeps them in an explicit global scratch laid out
// [block slot][vector][level][thread] (each warp access is one 256-byte run), read and written
// through L2 only: the persistent kernel below re-uses a block's slot for every tile it processes
// and DISCARDS the lines (discard.global.L2: dropped without write-back) once the upward sweep
// has read them, so that the vectors make their round trip through L2 instead of HBM (measured,
// scripts/probes/l2scratch_probe.cu: the write-backs of the local-memory variant disappear).
template <int NVEC>
struct LocalCols {
  double a[NVEC][KMAX];
  POM_HD void put(int v, int k, double x) { a[v][k] = x; }
  POM_HD double get(int v, int k) const { return a[v][k]; }
};

#ifndef POMGPU_EMU
template <int NVEC>
struct ScratchCols {
  double* base;   // this thread's element of level 0 of vector 0
  int nt, kbs;    // threads per block, levels per vector
  uint64_t pol;   // L2 evict_last policy
  __device__ __forceinline__ void put(int v, int k, double x) {
#ifdef POM_SCR_PLAIN
    __stcg(base + (size_t)(v * kbs + k) * nt, x);
#else
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(base + (size_t)(v * kbs + k) * nt), "d"(x), "l"(pol) : "memory");
#endif
  }
  __device__ __forceinline__ double get(int v, int k) const {
#ifdef POM_SCR_PLAIN
    return __ldcg(base + (size_t)(v * kbs + k) * nt);
#else
    double x;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(x) : "l"(base + (size_t)(v * kbs + k) * nt), "l"(pol));
    return x;
#endif
  }
};

// Persistent column kernel on the TMA ring: one block per resident slot, each looping over tiles
// (tile n of block b = b + n*gridDim.x).  The ring runs ACROSS tiles: while a tile's upward sweep
// (`post`) runs, the first NS levels of the block's next tile are already in flight.  No flux
// exchange between threads (thread tile = output tile), only the operand staging.  F provides NF,
// NS, TY, OHL/OHR/OHB/OHT, BW, BH, UP, NK, NVEC, fields(), k0(), k1(), kl1(),
// pre(i,j,State&,CM&), level(i,j,k,State&,CM&,op), post(i,j,State&,CM&) with CM = the column
// vectors above.  `post` runs the upward sweeps (Thomas back-substitution) on the thread's own column.
template <class F>
__global__ void __launch_bounds__(TILE_X * F::TY, F::MINB)
tmacolkernel(const __grid_constant__ TmaMaps<F::NF> maps, const F f, int i0, int i1, int j0, int j1,
             int nbx, int ntiles, double* scratch, int kbs, unsigned long long* ringctl, int nslots) {
  constexpr int NF = F::NF, NS = F::NS, PL = tma_plane(F::BW, F::BH), NT = TILE_X * F::TY;
  extern __shared__ __align__(128) double pom_tsm[];
  double* ring = pom_tsm;
  uint64_t* bar = (uint64_t*)(ring + NS * NF * PL);
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TILE_X + tx;
  static_assert(F::BW % 2 == 0 && F::BW >= TILE_X + F::OHL + F::OHR + 1, "TMA box too narrow");
  static_assert(F::BH >= F::TY + F::OHB + F::OHT, "TMA box too short");
  const int k0 = f.k0(), k1 = f.k1(), kl1 = f.kl1();
  const int nl = kl1 - k0 + 1;                                        // levels staged per tile
  const int ntl = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this block
  const int total = ntl * nl;                                         // stage loads of this block
  const bool leader = (tid == 0);
  // With one tile per block (gridDim.x == ntiles) the block borrows a scratch slot from a ring of
  // free slot numbers: ringctl[0] counts the slots taken, ringctl[1] the slots given back, entry
  // t % nslots of the ring holds (slot | generation t/nslots << 16).  At most nslots blocks are
  // resident, so a free entry always exists; the generation tag makes a taker wait for the write of
  // a giver that has claimed its entry but not stored it yet.  The persistent variant (gridDim.x <=
  // nslots) simply uses slot blockIdx.x.
  __shared__ int s_slot;
  const bool borrowed = ((int)gridDim.x > nslots);
  if (leader) {
    for (int s = 0; s < NS; ++s) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int slot = (int)blockIdx.x;
    if (borrowed) {
      const unsigned long long h = atomicAdd(&ringctl[0], 1ull);
      volatile unsigned long long* ent = ringctl + 2 + (h % (unsigned long long)nslots);
      unsigned long long e;
      do { e = *ent; } while ((e >> 16) != h / (unsigned long long)nslots);
      slot = (int)(e & 0xffffull);
    }
    s_slot = slot;
  }
  __syncthreads();
  const int slot = s_slot;
#ifdef POM_TMA_PLAIN
  const uint64_t pol_ef = 0;
#else
  const uint64_t pol_ef = l2_policy_evict_first();
#endif
  // stage load number q of this block = level k0 + q%nl of its tile q/nl
  auto issue = [&](int q) {
    const int n = q / nl, L = k0 + (q - n * nl), t = (int)blockIdx.x + n * (int)gridDim.x;
    const int by = t / nbx, bx = t - by * nbx;
    const int n0 = i0 + bx * TILE_X - 1 - F::OHL, c0 = n0 - (n0 & 1), c1 = j0 + by * F::TY - 1 - f.g.joff - F::OHB;
    const int s = q % NS;
    mbar_expect_tx(&bar[s], (uint32_t)(NF * F::BW * F::BH * sizeof(double)));
#pragma unroll
    for (int m = 0; m < NF; ++m) {
#ifdef POM_TMA_PLAIN
      tma_load_3d(ring + (s * NF + m) * PL, &maps.m[m], &bar[s], c0, c1, L - 1);
#else
      tma_load_3d_hint(ring + (s * NF + m) * PL, &maps.m[m], &bar[s], c0, c1, L - 1, pol_ef);
#endif
    }
  };
  int qi = 0;                                                         // next load to issue (leader)
  if (leader)
    for (; qi < NS && qi < total; ++qi) issue(qi);
#ifdef POM_COLS_LOCAL   // (experiment) the vectors in per-thread local memory, like the round-1 kernels
  LocalCols<F::NVEC> cm;
#else
  ScratchCols<F::NVEC> cm{scratch + (size_t)slot * F::NVEC * kbs * NT + tid, NT, kbs, l2_policy_evict_last()};
#endif
  for (int n = 0; n < ntl; ++n) {
    const int t = (int)blockIdx.x + n * (int)gridDim.x;
    const int by = t / nbx, bx = t - by * nbx;
    const int ti0 = i0 + bx * TILE_X, tj0 = j0 + by * F::TY;
    const int i = ti0 + tx, j = tj0 + ty;
    const int shift = (ti0 - 1 - F::OHL) & 1;
    const bool active = (i <= i1 && j <= j1);
    typename F::State st;     // scalars: registers
    if (active) f.pre(i, j, st, cm);
    const int own = (ty + F::OHB) * F::BW + tx + F::OHL + shift;
    for (int k = k0; k <= k1; ++k) {
      const int q0 = n * nl + (k - k0), s0 = q0 % NS;
      mbar_wait(&bar[s0], (q0 / NS) & 1);
      int s1 = s0;
      if (F::UP && k + 1 <= kl1) {
        s1 = (q0 + 1) % NS;
        mbar_wait(&bar[s1], ((q0 + 1) / NS) & 1);
      }
      const SmemOp<F> op{ring + s0 * NF * PL + own, ring + s1 * NF * PL + own};
      if (active) f.level(i, j, k, st, cm, op);
      __syncthreads();
      // everyone is done with the stage of level k (and, after the last computed level, with the
      // stages that were only looked at through op.up): refill them with the loads NS ahead
      if (leader) {
        const int qfree = (k == k1) ? n * nl + nl - 1 : q0;
        for (; qi <= qfree + NS && qi < total; ++qi) issue(qi);
      }
    }
    if (active) f.post(i, j, st, cm);
    // drop this tile's column vectors from L2 without a write-back: every thread of the block has
    // read its own (first barrier); nobody writes the next tile's before the lines are gone (second)
#if !defined(POM_COLS_LOCAL) && !defined(POM_NO_DISCARD)
    __syncthreads();
    {
      constexpr int LPL = NT / 16;                                    // 128-byte lines per level of one vector
      const int nlines = F::NVEC * kbs * LPL;
      const double* mine = scratch + (size_t)slot * F::NVEC * kbs * NT;
      for (int ln = tid; ln < nlines; ln += NT)
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(mine + (size_t)ln * 16) : "memory");
    }
    __syncthreads();
#endif
  }
  __syncthreads();
  if (borrowed && leader) {   // give the slot back: nobody of this block uses it any more
    __threadfence();
    const unsigned long long t = atomicAdd(&ringctl[1], 1ull);
    volatile unsigned long long* ent = ringctl + 2 + (t % (unsigned long long)nslots);
    *ent = (unsigned long long)slot | ((t / (unsigned long long)nslots) << 16);
    __threadfence();
  }
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
      tmakernel<F><<<gr, b, smem, (cudaStream_t)c->stream>>>(maps, f, i0, i1, j0, j1);
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
  double* S = ring + (TMA ? NF * PL : 0);         // [NV][TILE_Y][TILE_X]
  double* E = S + NV * TILE_Y * TILE_X;           // [TILE_Y][TILE_X]
  uint64_t* bar = (uint64_t*)(E + TILE_Y * TILE_X);
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
    // Programmatic dependent launch: this grid may become resident while the previous kernel of the
    // stream (the previous external substep) is still draining.  Everything that kernel does not
    // write -- the operands F::is_static() names: metrics, depth, Coriolis, forcing -- is fetched
    // BEFORE griddepcontrol.wait, the time-stepped fields after it.
    if (leader) {
      // two transactions: the first F::NFA fields are all that phases A and B read, so they start
      // while the operands only phase C needs are still in flight
      mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_expect_tx(&bar[0], (uint32_t)(F::NFA * F::BW * F::BH * sizeof(double)));
      mbar_expect_tx(&bar[1], (uint32_t)((NF - F::NFA) * F::BW * F::BH * sizeof(double)));
      const int c0 = n0 - shift, c1 = tj0 - 1 - f.g.joff - F::OHB;
#pragma unroll
      for (int n = 0; n < NF; ++n)
        if (F::is_static(n)) tma_load_3d(ring + n * PL, &maps.m[n], &bar[n < F::NFA ? 0 : 1], c0, c1, 0);
    }
    f.pre(i, j, inside, st);        // plain loads of the static point-wise operands overlap the TMA
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (leader) {
      const int c0 = n0 - shift, c1 = tj0 - 1 - f.g.joff - F::OHB;
#pragma unroll
      for (int n = 0; n < NF; ++n)
        if (!F::is_static(n)) tma_load_3d(ring + n * PL, &maps.m[n], &bar[n < F::NFA ? 0 : 1], c0, c1, 0);
    }
    f.pre_dynamic(i, j, inside, st);
    __syncthreads();                // barrier init visible to every waiter
    mbar_wait(&bar[0], 0);
  } else {
    f.pre(i, j, inside, st);
    f.pre_dynamic(i, j, inside, st);
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
  for (int n = 0; n < NV; ++n) S[(n * TILE_Y + ty) * TILE_X + tx] = v[n];
  __syncthreads();
  double e = 0.;
  if (bok) e = TMA ? f.phaseB(i, j, st, sop, Tile2{S, tx, ty, TILE_Y}, tx >= 1 && ty >= 1) : f.phaseB(i, j, st, gop, Tile2{S, tx, ty, TILE_Y}, tx >= 1 && ty >= 1);
  E[ty * TILE_X + tx] = e;
  if (TMA) mbar_wait(&bar[1], 0);
  __syncthreads();
  if (out) { if (TMA) f.phaseC(i, j, st, sop, Tile2{S, tx, ty, TILE_Y}, Tile2{E, tx, ty, TILE_Y}); else f.phaseC(i, j, st, gop, Tile2{S, tx, ty, TILE_Y}, Tile2{E, tx, ty, TILE_Y}); }
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
  static double S[F::NV * TILE_Y * TILE_X], E[TILE_Y * TILE_X];
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
        for (int n = 0; n < F::NV; ++n) S[(n * TILE_Y + ty) * TILE_X + tx] = v[n];
      }
      POM_T3_LOOP {
        POM_T3_IJ;
        double e = 0.;
        if (tx <= TILE_X - 2 && ty <= F::TY - 2 && ins[ty][tx])
          e = f.phaseB(i, j, st[ty][tx], GlobalOp<F>{fld, f.g, i, j, 1}, Tile2{S, tx, ty, TILE_Y}, tx >= 1 && ty >= 1);
        E[ty * TILE_X + tx] = e;
      }
      POM_T3_LOOP {
        POM_T3_IJ;
        if (tx >= 1 && tx <= TILE_X - 2 && ty >= 1 && ty <= F::TY - 2 && i <= i1 && j <= j1 && ins[ty][tx])
          f.phaseC(i, j, st[ty][tx], GlobalOp<F>{fld, f.g, i, j, 1}, Tile2{S, tx, ty, TILE_Y}, Tile2{E, tx, ty, TILE_Y});
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
  constexpr size_t sm_se = (size_t)((F::NV + 1) * TILE_Y * TILE_X) * sizeof(double) + 16;
  constexpr size_t sm_tma = sm_se + (size_t)F::NF * tma_plane(F::BW, F::BH) * sizeof(double);
  static DevOnce granted;
  if (granted.need(c->device)) {
    cudaFuncSetAttribute(tile3kernel<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_tma);
    cudaFuncSetAttribute(tile3kernel<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_se);
  }
  if (tma_ok) {
    // launched with programmatic stream serialization: its blocks may start their static prologue
    // while the previous kernel drains (they wait for it in griddepcontrol.wait)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = gr; cfg.blockDim = b; cfg.dynamicSmemBytes = sm_tma; cfg.stream = (cudaStream_t)c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = c->no_pdl ? 0 : 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, tile3kernel<F, true>, maps, f, i0, i1, j0, j1);
  } else tile3kernel<F, false><<<gr, b, sm_se, (cudaStream_t)c->stream>>>(maps, f, i0, i1, j0, j1);
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
  constexpr int NT = TILE_X * F::TY;
  const int nbx = (i1 - i0 + TILE_X) / TILE_X, nby = (j1 - j0 + F::TY) / F::TY;
  dim3 b(TILE_X, F::TY);
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
      static int resident[64];                       // blocks per SM of this kernel, per device
      static unsigned long long* ringmem[64];        // free-slot ring of this kernel, per device (one tile per block mode)
      const int dv = c->device & 63;
      if (granted.need(c->device)) {
        cudaFuncSetAttribute(tmacolkernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tmacolkernel<F>, NT, smem);
        resident[dv] = occ > 0 ? occ : 1;
        // ring: [taken, given back, entries...]; initially every slot is free with generation 0
        const int ns = c->nsm * resident[dv];
        unsigned long long* h = (unsigned long long*)malloc((size_t)(ns + 2) * 8);
        h[0] = 0; h[1] = (unsigned long long)ns;
        for (int q = 0; q < ns; ++q) h[2 + q] = (unsigned long long)q;
        if (cudaMalloc((void**)&ringmem[dv], (size_t)(ns + 2) * 8) == cudaSuccess)
          cudaMemcpy(ringmem[dv], h, (size_t)(ns + 2) * 8, cudaMemcpyHostToDevice);
        else ringmem[dv] = nullptr;
        free(h);
      }
      const int slots = c->nsm * resident[dv], ntiles = nbx * nby;
#ifdef POM_NONPERSIST   // one tile per block, scratch slots borrowed from the ring
      const int grid = ringmem[dv] ? ntiles : (ntiles < slots ? ntiles : slots);
#else
      const int grid = ntiles < slots ? ntiles : slots;
#endif
      const size_t need = (size_t)slots * F::NVEC * c->g.kb * NT;
      if (need > c->colscr_cap) {
        if (c->colscr) { cudaStreamSynchronize((cudaStream_t)c->stream); cudaFree(c->colscr); c->colscr = nullptr; c->colscr_cap = 0; }
        if (cudaMalloc((void**)&c->colscr, need * sizeof(double)) == cudaSuccess) c->colscr_cap = need;
        else { (void)cudaGetLastError(); tma_ok = false; }   // no room for the scratch: direct-load kernel below
      }
      if (tma_ok)
        tmacolkernel<F><<<grid, b, smem, (cudaStream_t)c->stream>>>(maps, f, i0, i1, j0, j1, nbx, ntiles, c->colscr, c->g.kb,
                                                                    ringmem[dv], slots);
    }
  }
  if (!tma_ok) colkernel_g<F><<<dim3(nbx, nby), b, 0, (cudaStream_t)c->stream>>>(f, i0, i1, j0, j1);
  if (c->prof_on) prof_after(c);
#endif
}

}  // namespace pom
