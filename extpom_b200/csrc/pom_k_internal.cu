// pom_k_internal.cu -- the 3-D internal mode (advance.f:356-537) and the
// solver.f kernels it drives.  One thread per (i,j) column marching in k; the
// tridiagonal (Thomas) sweeps keep their ee/gg coefficients in per-thread
// arrays so no a/c/ee/gg 3-D temporaries (solver.f:1224-1230,1552-1554,
// 1692-1693) ever reach HBM.  Expression order follows the Fortran.
#include "pom_core.h"
#include "pom_tma.h"
#include "pom_names.h"
#include "pom_bcond.h"

// min blocks/SM of the plain column kernels whose occupancy (not traffic) limits them: capping
// them at 64 registers (4 x 256 threads) was measured at 0.43 -> 0.28 ms for realvertvl
#ifndef POM_PROFT_MINB
#define POM_PROFT_MINB 4
#endif
#ifndef POM_ADVPROF_MINB
#define POM_ADVPROF_MINB 3
#endif
#ifndef POM_ADVPROF_NS
#define POM_ADVPROF_NS 3
#endif
#ifndef POM_PROFQ_MINB
#define POM_PROFQ_MINB 3
#endif
#ifndef POM_PROFQ_NS
#define POM_PROFQ_NS 3
#endif
#ifndef POM_PROFQ_TY
#define POM_PROFQ_TY 6
#endif
#ifndef POM_PROFT_TY
#define POM_PROFT_TY 8
#endif
#ifndef POM_ADVPROF_TY
#define POM_ADVPROF_TY 8
#endif
#ifndef POM_TILE_TY
#define POM_TILE_TY 16
#define POM_TILE_MINB 1
#define POM_TILE_NS 4
#endif
#ifndef POM_PFD
#define POM_PFD 2
#endif
namespace pom {

// ---------------------------------------------------------------------------
// advance.f:365-393: adjust u(z), v(z) so that their depth means match (utb+utf)
struct UvAdjustK : KBase {
  POM_KINFO("uv_adjust", 2, 2, 5, 0)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    double tu = 0., tv = 0.;
    for (int k = 1; k <= kbm1; ++k) {
      PF3(p.u,i,j,k+3); PF3(p.v,i,j,k+3);
      tu=tu+u(i,j,k)*dz(k);
      tv=tv+v(i,j,k)*dz(k);
    }
    if (i >= 2) {
      const double r=(utb(i,j)+utf(i,j))/(dt(i,j)+dt(i-1,j));
      for (int k = 1; k <= kbm1; ++k) u(i,j,k)=(u(i,j,k)-tu)+r;
    }
    if (j >= 2) {
      const double r=(vtb(i,j)+vtf(i,j))/(dt(i,j)+dt(i,j-1));
      for (int k = 1; k <= kbm1; ++k) v(i,j,k)=(v(i,j,k)-tv)+r;
    }
  }
};

// ---------------------------------------------------------------------------
// vertvl (solver.f:1970-2021) + bcondorl(5) (bounds_forcing.f:553-559)
struct VertvlK : KBase {
  POM_KINFO("vertvl", 2, 1, 8, 0)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    const double m = fsm(i,j);
    if (i >= 2 && i <= imm1 && j >= 2 && j <= jmm1) {
      double wk=0.5*(vfluxb(i,j)+vfluxf(i,j));                          // :2004
      RDiv ddxy; ddxy.set(dx(i,j)*dy(i,j));
      const double de=(etf(i,j)-etb(i,j))/dti2;
      // .25*(dy+dy)*(dt+dt) of the four faces (:1984-1985,1993-1994), constant along k
      const double cW=.25*(dy(i,j)+dy(i-1,j))*(dt(i,j)+dt(i-1,j));
      const double cE=.25*(dy(i+1,j)+dy(i,j))*(dt(i+1,j)+dt(i,j));
      const double cS=.25*(dx(i,j)+dx(i,j-1))*(dt(i,j)+dt(i,j-1));
      const double cN=.25*(dx(i,j+1)+dx(i,j))*(dt(i,j+1)+dt(i,j));
      for (int k = 1; k <= kbm1; ++k) {
        PF3(p.u,i,j,k+2); PF3(p.v,i,j,k+2); PF3(p.v,i,j+1,k+2);
        w(i,j,k)=wk*m;
        wk=wk+dz(k)*(ddxy(cE*u(i+1,j,k)-cW*u(i,j,k)+cN*v(i,j+1,k)-cS*v(i,j,k))+de);   // :2011-2015
      }
      w(i,j,kb)=wk;
    } else {
      for (int k = 1; k <= kbm1; ++k) w(i,j,k)=w(i,j,k)*m;
    }
  }
};

// ---------------------------------------------------------------------------
// advance.f:365-393 + vertvl (solver.f:1970-2021) + bcondorl(5) in ONE sweep.  The depth sums
// tps=sum_k u*dz(k), sum_k v*dz(k) come in as 2-D fields (s2c, s2d): the previous step's uv_filter
// accumulated them, in the same k order, while it produced the arrays that are now u and v
// (UvSumK recomputes them when u or v were pushed since).  With the sums known, the adjusted
// velocity of the east / north neighbour that vertvl's flux difference needs is a point-wise
// expression, so u, v are read once (2R+3W instead of 6+3 passes).  The adjusted fields go to
// s3a, s3b (neighbours still read the raw u, v); the caller swaps the buffers.
// A TMA column kernel (pom_tma.h: tmacolkernel): u, v of every level staged with their east / north neighbours,
// two or more levels ahead (plain-load column kernel: 0.38 against 0.32 ms).
#ifndef POM_UVADJ_TY
#define POM_UVADJ_TY 4
#define POM_UVADJ_MINB 8
#define POM_UVADJ_NS 4
#endif
struct UvAdjVertvlTK : KBase {
  POM_KINFO("uvadjust_vertvl", 2, 3, 14, 0)
  using KBase::KBase;
  static constexpr int TY = POM_UVADJ_TY, MINB = POM_UVADJ_MINB;
  static constexpr int NF = 2, NS = POM_UVADJ_NS, OHL = 0, OHR = 1, OHB = 0, OHT = 1, BW = 34, BH = TY + 1, NK = 0;
  static constexpr bool UP = false;
  static constexpr int NVEC = 1;   // (unused)
  enum { U, V };
  POM_HD void fields(const double** b) const { b[U] = p.u; b[V] = p.v; }
  struct State { double m, tu, tv, ru, rv, tuE, ruE, tvN, rvN, wk, de, cW, cE, cS, cN; RDiv ddxy; bool interior, au, av; };
  POM_HD int k0() const { return 1; }
  POM_HD int k1() const { return g.kb; }
  POM_HD int kl1() const { return g.kb; }
  template <class CM>
  POM_HD void pre(int i, int j, State& s, CM&) const {
    POM_DIMS;
    s.m = fsm(i,j);
    s.interior = (i >= 2 && i <= imm1 && j >= 2 && j <= jmm1);
    s.au = (i >= 2); s.av = (j >= 2);                                     // :373, :386
    s.tu=A2(p.s2c,i,j); s.tv=A2(p.s2d,i,j);
    s.ru = s.au ? (utb(i,j)+utf(i,j))/(dt(i,j)+dt(i-1,j)) : 0.;
    s.rv = s.av ? (vtb(i,j)+vtf(i,j))/(dt(i,j)+dt(i,j-1)) : 0.;
    s.tuE = 0.; s.ruE = 0.; s.tvN = 0.; s.rvN = 0.; s.wk = 0.; s.de = 0.; s.cW = 0.; s.cE = 0.; s.cS = 0.; s.cN = 0.;
    s.ddxy.set(s.interior ? dx(i,j)*dy(i,j) : 1.);
    if (s.interior) {
      s.tuE=A2(p.s2c,i+1,j); s.ruE=(utb(i+1,j)+utf(i+1,j))/(dt(i+1,j)+dt(i,j));
      s.tvN=A2(p.s2d,i,j+1); s.rvN=(vtb(i,j+1)+vtf(i,j+1))/(dt(i,j+1)+dt(i,j));
      s.wk=0.5*(vfluxb(i,j)+vfluxf(i,j));                               // solver.f:2004
      s.de=(etf(i,j)-etb(i,j))/dti2;
      s.cW=.25*(dy(i,j)+dy(i-1,j))*(dt(i,j)+dt(i-1,j));                 // :1984-1985,1993-1994
      s.cE=.25*(dy(i+1,j)+dy(i,j))*(dt(i+1,j)+dt(i,j));
      s.cS=.25*(dx(i,j)+dx(i,j-1))*(dt(i,j)+dt(i,j-1));
      s.cN=.25*(dx(i,j+1)+dx(i,j))*(dt(i,j+1)+dt(i,j));
    }
  }
  template <class Op, class CM>
  POM_HD void level(int i, int j, int k, State& s, CM&, const Op& o) const {
    const double u0=o(U,0,0), v0=o(V,0,0);
    if (k == g.kb) {
      A3(p.s3a,i,j,k)=u0;
      A3(p.s3b,i,j,k)=v0;
      if (s.interior) w(i,j,k)=s.wk;
      return;
    }
    const double ua = s.au ? (u0-s.tu)+s.ru : u0;                       // advance.f:374-375
    const double va = s.av ? (v0-s.tv)+s.rv : v0;                       // :387-388
    A3(p.s3a,i,j,k)=ua;
    A3(p.s3b,i,j,k)=va;
    if (s.interior) {
      const double uE=(o(U,1,0)-s.tuE)+s.ruE, vN=(o(V,0,1)-s.tvN)+s.rvN;
      w(i,j,k)=s.wk*s.m;                                                // bounds_forcing.f:553-559
      s.wk=s.wk+dz(k)*(s.ddxy(s.cE*uE-s.cW*ua+s.cN*vN-s.cS*va)+s.de);   // solver.f:2011-2015
    } else {
      w(i,j,k)=w(i,j,k)*s.m;
    }
  }
  template <class CM>
  POM_HD void post(int, int, State&, CM&) const {}
};

// the depth sums of u and v (advance.f:367-369,380-382) when uv_filter's are not current
struct UvSumK : KBase {
  POM_KINFO("uv_sum", 2, 0, 0, 2)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    double tu = 0., tv = 0.;
    for (int k = 1; k <= kbm1; ++k) {
      PF3(p.u,i,j,k+3); PF3(p.v,i,j,k+3);
      tu=tu+u(i,j,k)*dz(k);
      tv=tv+v(i,j,k)*dz(k);
    }
    A2(p.s2c,i,j)=tu;
    A2(p.s2d,i,j)=tv;
  }
};

// ---------------------------------------------------------------------------
// advq (solver.f:411-477) for q2 -> uf and q2l -> vf in one pass, as a tile kernel: each
// thread evaluates the x/y fluxes of both quantities at its own point once per level.
struct AdvqK : KBase {
  POM_KINFO("advq", 8, 2, 9, 0)
  using KBase::KBase;
  static constexpr int NV = 4, HL = 0, HR = 1, HB = 0, HT = 1, TY = POM_TILE_TY, MINB = POM_TILE_MINB;
#ifndef POM_NS_ADVQ
#define POM_NS_ADVQ POM_TILE_NS
#endif
  static constexpr int NF = 8, NS = POM_NS_ADVQ, OHL = 1, OHR = 0, OHB = 1, OHT = 0, BW = 34, BH = TY + 1, NK = 0;
  static constexpr bool UP = true;
  static constexpr bool FULL = false;   // stage() may run for every thread and assigns every v[]
  enum { Q2, Q2B, Q2L, Q2LB, U, V, AAM, W };
  enum { XA, YA, XB, YB };
  POM_HD void fields(const double** b) const {
    b[Q2] = p.q2; b[Q2B] = p.q2b; b[Q2L] = p.q2l; b[Q2LB] = p.q2lb; b[U] = p.u; b[V] = p.v; b[AAM] = p.aam; b[W] = p.w;
  }
  struct State {
    double dtx, dty, hx, hy, dumc, dvmc, hdy, hdx;   // (dt+dt), (h+h), masks, .5*(dy+dy), .5*(dx+dx)
    RDiv ddxs, ddys, dhf;                             // (dx+dx(i-1)), (dy+dy(j-1)), (h+etf)*art
    double hb, ar;                                    // (h+etb)*art, art
    double um, vm, a0m, aWm, aSm;                     // level k-1: u, v, aam(i,j), aam(i-1,j), aam(i,j-1)
    double wm, w0, qam, qbm, qa0, qb0;                // w(k-1), w(k), q(k-1), q(k) of q2 / q2l (output columns)
    double qab, qbb;                                  // q2b, q2lb of this level (stage -> combine)
    bool fxa, fya, interior;
  };
  POM_HD int k0() const { return 2; }
  POM_HD int k1() const { return g.kb - 1; }
  POM_HD int kl1() const { return g.kb; }
  POM_HD void pre(int i, int j, bool inside, bool out, State& s) const {
    POM_DIMS;
    const int jlo = g.joff + 1;
    s.fxa = inside && i >= 2 && j >= 2;                       // :426-427 (j<=jm, i<=im)
    s.fya = inside && i >= 2 && j >= 2 && j - 1 >= jlo;
    s.interior = out && i >= 2 && i <= imm1 && j >= 2 && j <= jmm1;
    if (out) {                                                // advance.f:403-404 and levels 1, kb
      uf(i,j,1)=0.; vf(i,j,1)=0.; uf(i,j,kb)=0.; vf(i,j,kb)=0.;
    }
    if (!(s.fxa || s.fya)) return;
    s.dtx = dt(i,j)+dt(i-1,j); s.hx = h(i,j)+h(i-1,j); s.dumc = dum(i,j);
    s.hdy = .5*(dy(i,j)+dy(i-1,j)); s.ddxs.set(dx(i,j)+dx(i-1,j));
    s.um = u(i,j,1); s.a0m = aam(i,j,1); s.aWm = aam(i-1,j,1);
    if (s.fya) {
      s.dty = dt(i,j)+dt(i,j-1); s.hy = h(i,j)+h(i,j-1); s.dvmc = dvm(i,j);
      s.hdx = .5*(dx(i,j)+dx(i,j-1)); s.ddys.set(dy(i,j)+dy(i,j-1));
      s.vm = v(i,j,1); s.aSm = aam(i,j-1,1);
    }
    if (s.interior) {
      s.ar = art(i,j);
      s.hb = (h(i,j)+etb(i,j))*s.ar;
      s.dhf.set((h(i,j)+etf(i,j))*s.ar);
      s.wm = w(i,j,1); s.w0 = w(i,j,2); s.qam = q2(i,j,1); s.qbm = q2l(i,j,1);
      s.qa0 = q2(i,j,2); s.qb0 = q2l(i,j,2);
    }
  }
  template <class Op>
  POM_HD void stage(int i, int j, int k, State& s, const Op& o, double* v) const {
    if (!(s.fxa || s.fya)) return;
    const double qa=o(Q2,0,0), qab=o(Q2B,0,0), qb=o(Q2L,0,0), qbb=o(Q2LB,0,0), a0=o(AAM,0,0);
    s.qab=qab; s.qbb=qbb;
    if (s.fxa) {
      const double aW=o(AAM,-1,0), u0=o(U,0,0);
      const double a4=a0+aW+s.a0m+s.aWm;                       // :441-442
      const double us=u0+s.um;
      double a=.125*(qa+o(Q2,-1,0))*s.dtx*us;                  // :428-429
      a=a-s.ddxs(.25*a4*s.hx*(qab-o(Q2B,-1,0))*s.dumc);        // :440-445
      v[XA]=s.hdy*a;                                           // :452
      double b=.125*(qb+o(Q2L,-1,0))*s.dtx*us;
      b=b-s.ddxs(.25*a4*s.hx*(qbb-o(Q2LB,-1,0))*s.dumc);
      v[XB]=s.hdy*b;
      s.um=u0; s.aWm=aW;
    }
    if (s.fya) {
      const double aS=o(AAM,0,-1), v0=o(V,0,0);
      const double a4=a0+aS+s.a0m+s.aSm;                       // :447-448
      const double vs=v0+s.vm;
      double a=.125*(qa+o(Q2,0,-1))*s.dty*vs;                  // :430-431
      a=a-s.ddys(.25*a4*s.hy*(qab-o(Q2B,0,-1))*s.dvmc);        // :446-451
      v[YA]=s.hdx*a;                                           // :453
      double b=.125*(qb+o(Q2L,0,-1))*s.dty*vs;
      b=b-s.ddys(.25*a4*s.hy*(qbb-o(Q2LB,0,-1))*s.dvmc);
      v[YB]=s.hdx*b;
      s.vm=v0; s.aSm=aS;
    }
    s.a0m=a0;
  }
  template <class Op>
  POM_HD void combine(int i, int j, int k, State& s, const Op& o, const Tile2& tl) const {
    double a = 0., b = 0.;
    if (s.interior) {
      const double w1=o.up(W), qa1=o.up(Q2), qb1=o.up(Q2L);
      const double dzk=dz(k)+dz(k-1);
      double ra=pdiv((s.wm*s.qam-w1*qa1)*s.ar,dzk)
                +tl(XA,1,0)-tl(XA,0,0)+tl(YA,0,1)-tl(YA,0,0);  // :465-468
      a=s.dhf(s.hb*s.qab-dti2*ra);                             // :469-471
      double rb=pdiv((s.wm*s.qbm-w1*qb1)*s.ar,dzk)
                +tl(XB,1,0)-tl(XB,0,0)+tl(YB,0,1)-tl(YB,0,0);
      b=s.dhf(s.hb*s.qbb-dti2*rb);
      s.wm=s.w0; s.w0=w1;
      s.qam=s.qa0; s.qbm=s.qb0; s.qa0=qa1; s.qb0=qb1;
    }
    uf(i,j,k)=a;
    vf(i,j,k)=b;
  }
  POM_HD void post(int, int, State&) const {}
};

// ---------------------------------------------------------------------------
// profq (solver.f:1212-1538): Mellor-Yamada 2.5 with the wave-breaking surface
// condition; two Thomas solves per column (q2 -> uf, q2l -> vf), then km,kh,kq.
struct ProfqK : KBase {
  POM_KINFO("profq", 13, 8, 7, 0)
  double cgg, const1;
  double zr[KMAX];   // 1/|z(k)-z(1)| + 1/|z(k)-z(kb)| (:1429-1431), k-only: tabulated on the host
  int fuse;          // also bcond(6) + the Asselin filter of q2,q2l (advance.f:414-417) in the upward sweep
  ProfqK(const Ctx* x, const double* hz, int fz) : KBase(x), fuse(fz) {
    // solver.f:1297: (15.8*cbcnst)**(2./3.) with single-precision literals promoted
    // to double; solver.f:1273 const1
    const double cbcnst = 100.;
    cgg = pow((double)15.8f * cbcnst, (double)(2.f / 3.f));
    const1 = pow(16.6, 2. / 3.) * 1.;
    const int kb = x->g.kb;
    for (int k = 0; k < KMAX; ++k) zr[k] = 0.;
    for (int k = 2; k <= kb - 1 && k < KMAX; ++k) zr[k] = 1. / fabs(hz[k - 1] - hz[0]) + 1. / fabs(hz[k - 1] - hz[kb - 1]);
  }
  // TMA-fed column kernel: the 13 operand fields of every level are staged in shared memory
  // by the TMA two levels ahead; one downward sweep computes, level by level, the speed of
  // sound, buoyancy gradient, length scale, gh, production, the forward eliminations of BOTH
  // tridiagonal systems and the km/kh/kq update (in place, old kq kept in rolling registers);
  // `post` back-substitutes.  Only the four ee/gg vectors live in per-thread memory.
  static constexpr int TY = POM_PROFQ_TY, MINB = POM_PROFQ_MINB;
  static constexpr int NF = 13, NS = POM_PROFQ_NS, OHL = 0, OHR = 1, OHB = 0, OHT = 1, BW = 34, BH = TY + 1, NK = 0;
  static constexpr bool UP = true;
  enum { T, S, RHO, Q2B, Q2LB, Q2, U, V, KM, KH, KQ, UF, VF };
  POM_HD void fields(const double** b) const {
    b[T] = p.t; b[S] = p.s; b[RHO] = p.rho; b[Q2B] = p.q2b; b[Q2LB] = p.q2lb; b[Q2] = p.q2; b[U] = p.u; b[V] = p.v;
    b[KM] = p.km; b[KH] = p.kh; b[KQ] = p.kq; b[UF] = p.uf; b[VF] = p.vf;
  }
  static constexpr int NVEC = 4;   // eliminated coefficients of the two systems (ee, gg of q2; ee, gg of q2l)
  enum { C_EE, C_GG, C_E2, C_G2 };
  struct State {
    double hh, dh, kl0, m, ccm, kqm, kq0, rhom, um, uEm, vm, vNm, eem, ggm, e2m, g2m, q2_2, ufkb;
    int il, ir, jl, jr;
    bool interior;
  };
  POM_HD int k0() const { return 1; }
  POM_HD int k1() const { return g.kb; }
  POM_HD int kl1() const { return g.kb; }
  // new km,kh,kq of one level: own cell masked; boundary neighbours get the unmasked value
  // times their own mask (N,S,E,W copies of :1510-1529 = index clamped into the interior)
  POM_HD void put_k(int i, int j, int k, const State& st, double nkq, double nkm, double nkh) const {
    kq(i,j,k)=nkq*st.m; km(i,j,k)=nkm*st.m; kh(i,j,k)=nkh*st.m;
    if (st.il | st.ir | st.jl | st.jr)
      for (int dj = -st.jl; dj <= st.jr; ++dj)
        for (int di = -st.il; di <= st.ir; ++di) {
          if (di == 0 && dj == 0) continue;
          const double me = fsm(i+di,j+dj);
          kq(i+di,j+dj,k)=nkq*me; km(i+di,j+dj,k)=nkm*me; kh(i+di,j+dj,k)=nkh*me;
        }
  }
  template <class CM>
  POM_HD void pre(int i, int j, State& st, CM& cm) const {
    POM_DIMS;
    const double surfl = 2.e5;                                                // :1244
    st.interior = (i >= 2 && i <= imm1 && j >= 2 && j <= jmm1);
    st.hh=h(i,j);
    st.dh=st.hh+etf(i,j);                                                     // :1248
    double utau2 = 0.;
    st.ufkb = 0.;
    if (i <= imm1 && j <= jmm1) {                                             // :1281-1288
      double su=.5*(wusurf(i,j)+wusurf(i+1,j)), sv=.5*(wvsurf(i,j)+wvsurf(i,j+1));
      utau2=sqrt(su*su+sv*sv);
      double bu=.5*(wubot(i,j)+wubot(i+1,j)), bv=.5*(wvbot(i,j)+wvbot(i,j+1));
      st.ufkb=sqrt(bu*bu+bv*bv)*const1;
      uf(i,j,kb)=st.ufkb;
    }
    const double l0=surfl*utau2/grav;                                         // :1299
    st.kl0=kappa*l0;
    l(i,j,1)=st.kl0; l(i,j,kb)=0.;                                            // :1351-1352
    st.il = (i == 2) ? 1 : 0; st.ir = (i == imm1) ? 1 : 0;
    st.jl = (j == 2) ? 1 : 0; st.jr = (j == jmm1) ? 1 : 0;
    st.m = fsm(i,j);
    st.eem = 0.; st.ggm = cgg*utau2;                                          // ee(1), gg(1) :1296-1297
    st.e2m = 0.; st.g2m = 0.;
    cm.put(C_EE,1,st.eem); cm.put(C_GG,1,st.ggm);
  }
  template <class Op, class CM>
  POM_HD void level(int i, int j, int k, State& st, CM& cm, const Op& o) const {
    POM_DIMS;
    const double a1 = 0.92, b1 = 16.6, a2 = 0.74, b2 = 10.1, c1 = 0.08;     // :1241
    const double e1 = 1.8, e2 = 1.33, sef = 1., shiw = 0.;                   // :1242-1244
    const double coef4=18.*a1*a1+9.*a1*a2, coef5=9.*a1*a2;                   // :1474-1475
    const double coef1=a2*(1.-6.*a1/b1*1.);                                  // :1481-1483 (stf=1)
    const double coef2=3.*a2*b2/1.+18.*a1*a2;
    const double coef3=a1*(1.-3.*c1-6.*a1/b1*1.);
    if (!st.interior) {
      // boundary columns: only l is kept; the reference's solves there are overwritten by
      // bcond(6) (advance.f:414) and km,kh,kq by the copies of :1510-1529
      if (k >= 2 && k <= kbm1) {
        double qb=fabs(o(Q2B,0,0)), qlb=fabs(o(Q2LB,0,0));
        if (!fuse) { q2b(i,j,k)=qb; q2lb(i,j,k)=qlb; }                        // :1325-1326
        double ll=fabs(qlb/qb);
        if (z(k) > -0.5) ll=fmax(ll,st.kl0);
        l(i,j,k)=ll;
      }
      return;
    }
    const double hh=st.hh, dh=st.dh;
    if (k == 1) {
      {
        double tp=o(T,0,0)+tbias, sp=o(S,0,0)+sbias;                          // cc(1) (:1304-1319)
        double pp=grav*rhoref*(-zz(1)*hh)*1.e-4;
        double cv=1449.1+.00821*pp+4.55*tp-.045*(tp*tp)+1.34*(sp-35.0);
        st.ccm=cv/sqrt((1.-.01642*pp/cv)*(1.-0.40*pp/(cv*cv)));
      }
      st.kqm=o(KQ,0,0); st.kq0=o.up(KQ);                                      // old kq(1), kq(2)
      {
        // gh(1)=0 (:1353): sh=coef1, sm=coef3 (:1484-1486 with gh=0)
        double sh=coef1/(1.-coef2*0.);
        double sm=coef3+sh*coef4*0.;
        sm=sm/(1.-coef5*0.);
        double pr=st.kl0*sqrt(fabs(o(Q2,0,0)));                               // :1499
        const double nq=(pr*.41*sh+st.kqm)*.5, nm=(pr*sm+o(KM,0,0))*.5, nh=(pr*sh+o(KH,0,0))*.5;   // :1500-1503
        put_k(i,j,1,st,nq,nm,nh);
      }
      st.rhom=o(RHO,0,0);
      st.um=o(U,0,0); st.uEm=o(U,1,0); st.vm=o(V,0,0); st.vNm=o(V,0,1);
      st.q2_2=o.up(Q2);
      return;
    }
    if (k == kb) {                                                            // l=0, gh=0
      double sh=coef1/(1.-coef2*0.);
      double sm=coef3+sh*coef4*0.;
      sm=sm/(1.-coef5*0.);
      double pq=0.*sqrt(fabs(o(Q2,0,0)));
      const double nq=(pq*.41*sh+st.kq0)*.5, nm=(pq*sm+o(KM,0,0))*.5, nh=(pq*sh+o(KH,0,0))*.5;
      put_k(i,j,kb,st,nq,nm,nh);
      return;
    }
    // ---- levels 2..kbm1 ----
    double tp=o(T,0,0)+tbias, sp=o(S,0,0)+sbias;
    double pp=grav*rhoref*(-zz(k)*hh)*1.e-4;
    double cv=1449.1+.00821*pp+4.55*tp-.045*(tp*tp)+1.34*(sp-35.0);
    double cck=pdiv(cv,sqrt((1.-pdiv(.01642*pp,cv))*(1.-pdiv(0.40*pp,cv*cv))));
    double qb=fabs(o(Q2B,0,0)), qlb=fabs(o(Q2LB,0,0));
    if (!fuse) {   // fused: the upward sweep takes the abs again and stores the FILTERED values
      q2b(i,j,k)=qb;                                                          // :1325-1326
      q2lb(i,j,k)=qlb;
    }
    const double rhok=o(RHO,0,0);
    double boygr=pdiv(grav*(st.rhom-rhok),dzz(k-1)*hh)
                 +pdiv((grav*grav)*2.,st.ccm*st.ccm+cck*cck);                     // :1327-1330
    st.ccm=cck; st.rhom=rhok;
    double ll=fabs(pdiv(qlb,qb));                                                   // :1338
    if (z(k) > -0.5) ll=fmax(ll,st.kl0);                                      // :1339
    double gh=pdiv((ll*ll)*boygr,qb);                                               // :1343
    gh=fmin(gh,.028);                                                         // :1344
    l(i,j,k)=ll;
    const double u0=o(U,0,0), uE=o(U,1,0), v0=o(V,0,0), vN=o(V,0,1);
    const double kmk=o(KM,0,0), khk=o(KH,0,0);
    double pr;
    {
      double su=u0-st.um+uE-st.uEm;
      double sv=v0-st.vm+vN-st.vNm;
      double dd=dzz(k-1)*dh;
      pr=pdiv(kmk*.25*sef*(su*su+sv*sv),dd*dd)-shiw*kmk*boygr;                    // :1362-1369
      pr=pr+khk*boygr;                                                        // :1370
    }
    st.um=u0; st.uEm=uE; st.vm=v0; st.vNm=vN;
    double dtf=pdiv(sqrt(fabs(qb))*1.,b1*ll+small);                               // :1388-1389 (stf=1)
    // tridiagonal coefficients from the OLD kq (:1258-1267)
    const double kqp=o.up(KQ);
    const double kq0=st.kq0;
    const double a=pdiv(-dti2*(kqp+kq0+2.*umol)*.5,dzz(k-1)*dz(k)*dh*dh);
    const double cq=pdiv(-dti2*(st.kqm+kq0+2.*umol)*.5,dzz(k-1)*dz(k-1)*dh*dh);
    // q2 forward elimination (:1394-1404)
    {
      double gi=pdiv(1.,a+cq*(1.-st.eem)-(2.*dti2*dtf+1.));
      st.eem=a*gi;
      st.ggm=(-2.*dti2*pr+cq*st.ggm-o(UF,0,0))*gi;
      cm.put(C_EE,k,st.eem); cm.put(C_GG,k,st.ggm);
    }
    // q2l forward elimination (:1417-1446)
    {
      double r=pdiv(zr[k]*ll,dh*kappa);
      double dtf2=dtf*(1.+e2*(r*r));                                          // :1429-1432
      if (k == 2) {
        st.e2m=0.;                                                            // :1421
        st.g2m=-kappa*z(2)*dh*st.q2_2;                                        // :1422
      } else {
        // :1423 assigns vf(kbm1)=kappa*(1+z(kbm1))*dh*q2(kbm1) before this sweep reads it
        double vk=(k == kbm1) ? kappa*(1+z(kbm1))*dh*o(Q2,0,0) : o(VF,0,0);
        double gi=pdiv(1.,a+cq*(1.-st.e2m)-(dti2*dtf2+1.));
        st.e2m=a*gi;
        st.g2m=(dti2*(-pr*ll*e1)+cq*st.g2m-vk)*gi;
      }
      cm.put(C_E2,k,st.e2m); cm.put(C_G2,k,st.g2m);
    }
    // km, kh, kq (:1478-1506) -- in place; kq's old value stays in kqm for level k+1
    {
      double sh=pdiv(coef1,1.-coef2*gh);
      double sm=coef3+sh*coef4*gh;
      sm=pdiv(sm,1.-coef5*gh);
      double pq=ll*sqrt(fabs(o(Q2,0,0)));
      const double nq=(pq*.41*sh+kq0)*.5, nm=(pq*sm+kmk)*.5, nh=(pq*sh+khk)*.5;
      put_k(i,j,k,st,nq,nm,nh);
    }
    st.kqm=kq0; st.kq0=kqp;
  }
  // ---- back-substitutions (:1406-1413, :1448-1455) and abs (:1460-1471) ----
  // the recurrences use the signed iterate; abs() is applied by the reference afterwards
  // fused tail of one level: mask + 1e-10 (bounds_forcing.f:318-319), new q2,q2l stay in uf,vf,
  // filtered q2,q2l go to q2b,q2lb (advance.f:416-417); q2b,q2lb were left un-abs'ed on purpose
  POM_HD void emit(int i, int j, int k, double a, double b, double m) const {
    const int kbm1 = g.kb - 1;
    a=a*m+1.e-10;
    b=b*m+1.e-10;
    uf(i,j,k)=a;
    vf(i,j,k)=b;
    double qb=q2b(i,j,k), qlb=q2lb(i,j,k);
    if (k >= 2 && k <= kbm1) { qb=fabs(qb); qlb=fabs(qlb); }            // solver.f:1325-1326
    const double q=q2(i,j,k), ql=q2l(i,j,k);
    q2b(i,j,k)=q+.5*smoth*(a+qb-2.*q);                                  // advance.f:416
    q2lb(i,j,k)=ql+.5*smoth*(b+qlb-2.*ql);                              // advance.f:417
  }
  template <class CM>
  POM_HD void post(int i, int j, State& st, CM& cm) const {
    POM_DIMS;
    if (!st.interior) {
      if (!fuse) return;
      // boundary columns: bcond(6) upstream values (bounds_forcing.f:264-311)
      for (int k = kb; k >= 1; --k) {
        double a, b;
        bcond6_edge(*this, i, j, k, a, b);                                // bounds_forcing.f:257-311
        emit(i,j,k,a,b,st.m);
      }
      return;
    }
    if (!fuse) {
      double up=st.ufkb;
      for (int ki = kbm1; ki >= 1; --ki) {
        up=cm.get(C_EE,ki)*up+cm.get(C_GG,ki);
        uf(i,j,ki)=(ki >= 2) ? fabs(up) : up;
      }
      double vp = 0.;                                                         // vf(kb)=0 (:1420)
      vf(i,j,kb)=0.;
      for (int ki = kbm1; ki >= 2; --ki) {
        vp=cm.get(C_E2,ki)*vp+cm.get(C_G2,ki);
        vf(i,j,ki)=fabs(vp);
      }
      vf(i,j,1)=0.;                                                           // :1419
      return;
    }
    double up=st.ufkb, vp = 0.;
    emit(i,j,kb,up,0.,st.m);                                                  // uf(kb) (:1285), vf(kb)=0 (:1420)
    // the eliminated coefficients come back from local memory (L2): fetch them one level ahead
    double e1n=cm.get(C_EE,kbm1), g1n=cm.get(C_GG,kbm1), e2n=cm.get(C_E2,kbm1), g2n=cm.get(C_G2,kbm1);
    for (int ki = kbm1; ki >= 2; --ki) {
      const double e1c=e1n, g1c=g1n, e2c=e2n, g2c=g2n;
      e1n=cm.get(C_EE,ki-1); g1n=cm.get(C_GG,ki-1);
      if (ki > 2) { e2n=cm.get(C_E2,ki-1); g2n=cm.get(C_G2,ki-1); }
      {   // operands of the level below, towards L1 while this one is finished
        const int o = POM_I3(i,j,ki-1);
        POM_PREFETCH(p.q2+o); POM_PREFETCH(p.q2b+o); POM_PREFETCH(p.q2l+o); POM_PREFETCH(p.q2lb+o);
      }
      up=e1c*up+g1c;
      vp=e2c*vp+g2c;
      emit(i,j,ki,fabs(up),fabs(vp),st.m);
    }
    up=e1n*up+g1n;
    emit(i,j,1,up,0.,st.m);                                                   // vf(1)=0 (:1419)
  }
};

// ---------------------------------------------------------------------------
// bcond(6) (bounds_forcing.f:257-324) + Asselin filter and rotation of q2,q2l
// (advance.f:416-421).  Filtered q2 -> q2b buffer, new q2 stays in uf; the host
// rotates pointers.
struct QFilterK : KBase {
  POM_KINFO("q_filter", 6, 4, 1, 0)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    const double m = fsm(i,j);
    for (int k = 1; k <= kb; ++k) {
      double a=uf(i,j,k), b=vf(i,j,k);
      bcond6_edge(*this, i, j, k, a, b);                                // bounds_forcing.f:257-311
      a=a*m+1.e-10;                                                     // :318-319
      b=b*m+1.e-10;
      uf(i,j,k)=a;
      vf(i,j,k)=b;
      q2b(i,j,k)=q2(i,j,k)+.5*smoth*(a+q2b(i,j,k)-2.*q2(i,j,k));        // advance.f:416,418
      q2lb(i,j,k)=q2l(i,j,k)+.5*smoth*(b+q2lb(i,j,k)-2.*q2l(i,j,k));    // advance.f:417,420
    }
  }
};

// ---------------------------------------------------------------------------
// advt2 with nitera=1 (solver.f:577-731 + the fsm mask of smol_adif :1898-1900):
// upstream advection, leapfrog update, horizontal diffusion of (fb-fclim); tile kernel:
// every thread evaluates the upwind and the diffusive x/y fluxes of its own point once.
template <int NT, bool UPW = true>   // NT tracers in one pass (T alone, or T and S sharing u, v, w, aam and the metrics)
struct AdvT2K : KBase {
  static const KInfo& info() {
    static const KInfo k1{"advt2", 6, 1, 10, 0}, k2{"advt2_ts", 8, 2, 10, 0};
    return NT == 1 ? k1 : k2;
  }
  const double *fb_[NT], *f_[NT], *fc_[NT];
  double* ff_[NT];
  static constexpr bool upw = UPW;   // false: only the horizontal diffusion of (fb-fclim) is added to ff (:691-726, nitera>1)
  AdvT2K(const Ctx* x, const double* fb, const double* f, const double* fc, double* ff) : KBase(x) {
    fb_[0] = fb; f_[0] = f; fc_[0] = fc; ff_[0] = ff;
  }
  AdvT2K(const Ctx* x) : KBase(x) {   // T -> uf and S -> vf (advance.f:430-431)
    fb_[0] = x->p.tb; f_[0] = x->p.t; fc_[0] = x->p.tclim; ff_[0] = x->p.uf;
    fb_[NT - 1] = x->p.sb; f_[NT - 1] = x->p.s; fc_[NT - 1] = x->p.sclim; ff_[NT - 1] = x->p.vf;
  }
  static constexpr int NV = 4 * NT, HL = 0, HR = 1, HB = 0, HT = 1, TY = POM_TILE_TY, MINB = POM_TILE_MINB;
  // operands staged by the TMA: box = thread tile + one column W and one row S (34 x 17)
  static constexpr int NF = UPW ? 2 * NT + 4 : 2 * NT + 1;   // (the diffusion-only pass reads no u, v, w)
  static constexpr int NS = POM_TILE_NS, OHL = 1, OHR = 0, OHB = 1, OHT = 0, BW = 34, BH = TY + 1, NK = 0;
  static constexpr bool UP = true;
  static constexpr bool FULL = true;   // stage() may run for every thread and assigns every v[]
  enum { AAM = 2 * NT, U, V, W };     // FB(t) = 2t, FC(t) = 2t+1
  enum { XF, YF, XD, YD };            // + 4t
  POM_HD void fields(const double** b) const {
    for (int t = 0; t < NT; ++t) { b[2 * t] = fb_[t]; b[2 * t + 1] = fc_[t]; }
    b[AAM] = p.aam;
    if (UPW) { b[U] = p.u; b[V] = p.v; b[W] = p.w; }
  }
  struct State {
    double cx, cy, hx, hy, dumc, dvmc, dys, dxs;   // .25*(dy+dy)*(dt+dt), (h+h), masks, (dy+dy(i-1)), (dx+dx(j-1))
    RDiv ddxs, ddys, def;                          // (dx+dx(i-1)), (dy+dy(j-1)), (h+etf)*art
    double eb, ar, m;                              // (h+etb)*art, art, fsm
    double zk[NT], fb0[NT];                        // zflux(k), fb(i,j,k) of each tracer
    bool fxa, fya, interior;
  };
  POM_HD int k0() const { return 1; }
  POM_HD int k1() const { return g.kb - 1; }
  POM_HD int kl1() const { return g.kb - 1; }
  POM_HD void pre(int i, int j, bool inside, bool out, State& s) const {
    POM_DIMS;
    const int jlo = g.joff + 1;
    s.fxa = inside && i >= 2 && j >= 2 && j <= jmm1;                    // xmassflux range :603-608
    s.fya = inside && i >= 2 && i <= imm1 && j >= 2 && j - 1 >= jlo;    // ymassflux range :610-615
    s.interior = out && i >= 2 && i <= imm1 && j >= 2 && j <= jmm1;
    s.m = out ? fsm(i,j) : 0.;
    if (s.fxa) {
      s.cx = 0.25*(dy(i-1,j)+dy(i,j))*(dt(i-1,j)+dt(i,j));              // :605-606
      s.hx = h(i,j)+h(i-1,j); s.dumc = dum(i,j); s.dys = dy(i,j)+dy(i-1,j);
      s.ddxs.set(dx(i,j)+dx(i-1,j));
    }
    if (s.fya) {
      s.cy = 0.25*(dx(i,j-1)+dx(i,j))*(dt(i,j-1)+dt(i,j));              // :612-613
      s.hy = h(i,j)+h(i,j-1); s.dvmc = dvm(i,j); s.dxs = dx(i,j)+dx(i,j-1);
      s.ddys.set(dy(i,j)+dy(i,j-1));
    }
    if (s.interior) {
      s.ar = art(i,j);
      s.eb = (h(i,j)+etb(i,j))*s.ar;
      s.def.set((h(i,j)+etf(i,j))*s.ar);
      const double w1 = w(i,j,1);
#pragma unroll
      for (int t = 0; t < NT; ++t) s.zk[t] = w1*A3(f_[t],i,j,1)*s.ar;   // :648 (itera==1)
    }
  }
  // Branch-free (see AdvctK::stage): the definition-range flags select, nothing is skipped.
  template <class Op>
  POM_HD void stage(int i, int j, int k, State& s, const Op& o, double* v) const {
    const double a0=o(AAM,0,0);
    const double xm=upw ? s.cx*o(U,0,0) : 0., xd=0.5*(a0+o(AAM,-1,0));   // :605-606, :696
    const double ym=upw ? s.cy*o(V,0,0) : 0., yd=0.5*(a0+o(AAM,0,-1));   // :612-613, :697
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const double fb0=o(2*t,0,0);
      const double fd0=fb0-o(2*t+1,0,0);                                // fb-fclim (:691)
      s.fb0[t]=fb0;
      const double fbW=o(2*t,-1,0);
      v[4*t+XF]=s.fxa ? 0.5*((xm+fabs(xm))*fbW+(xm-fabs(xm))*fb0) : 0.;            // :631-635
      v[4*t+XD]=s.fxa ? s.ddxs(-xd*s.hx*tprni*(fd0-(fbW-o(2*t+1,-1,0)))*s.dumc*s.dys*0.5) : 0.;   // :705-707
      const double fbS=o(2*t,0,-1);
      v[4*t+YF]=s.fya ? 0.5*((ym+fabs(ym))*fbS+(ym-fabs(ym))*fb0) : 0.;            // :637-641
      v[4*t+YD]=s.fya ? s.ddys(-yd*s.hy*tprni*(fd0-(fbS-o(2*t+1,0,-1)))*s.dvmc*s.dxs*0.5) : 0.;   // :708-710
    }
  }
  template <class Op>
  POM_HD void combine(int i, int j, int k, State& s, const Op& o, const Tile2& tl) const {
    if (!s.interior) {
      // ff is not assigned here by the reference (bcond(4) sets it afterwards); only the
      // smol_adif mask applies
      if (upw) {
#pragma unroll
        for (int t = 0; t < NT; ++t) A3(ff_[t],i,j,k)=A3(ff_[t],i,j,k)*s.m;
      }
      return;
    }
    if (!upw) {   // diffusion only: ff=ff-dti2*div/((h+etf)*art) (:718-726)
#pragma unroll
      for (int t = 0; t < NT; ++t)
        A3(ff_[t],i,j,k)=A3(ff_[t],i,j,k)-s.def(dti2*(tl(4*t+XD,1,0)-tl(4*t+XD,0,0)+tl(4*t+YD,0,1)-tl(4*t+YD,0,0)));
      return;
    }
    const bool more = (k + 1 <= g.kb - 1);
    const double w1 = more ? o.up(W) : 0.;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      double zk1 = 0.;                                                  // zflux(k+1); :651 at kb
      if (more) {
        const double fb1=o.up(2*t);
        zk1=0.5*((w1+fabs(w1))*fb1+(w1-fabs(w1))*s.fb0[t]);             // :656-660
        zk1=zk1*s.ar;                                                   // :661
      }
      double q=tl(4*t+XF,1,0)-tl(4*t+XF,0,0)+tl(4*t+YF,0,1)-tl(4*t+YF,0,0)+pdiv(s.zk[t]-zk1,dz(k));   // :670-672
      q=s.def(s.fb0[t]*s.eb-dti2*q);                                    // :673-674
      q=q*s.m;                                                          // smol_adif :1899
      q=q-s.def(dti2*(tl(4*t+XD,1,0)-tl(4*t+XD,0,0)+tl(4*t+YD,0,1)-tl(4*t+YD,0,0)));   // :721-723
      POM_STCS(&A3(ff_[t],i,j,k),q);
      s.zk[t]=zk1;
    }
  }
  POM_HD void post(int i, int j, State& s) const {
    const int kb = g.kb;
    if (!upw) return;
#pragma unroll
    for (int t = 0; t < NT; ++t) A3(ff_[t],i,j,kb)=A3(ff_[t],i,j,kb)*s.m;   // smol_adif mask, level kb
  }
};

// ---------------------------------------------------------------------------
// advt2 with nitera > 1 (solver.f:577-731 + smol_adif :1880-1967): the Smolarkiewicz
// iterations need the anti-diffusive mass fluxes and the previous iterate as arrays, so the
// scheme runs as mass -> { upwind step -> smol_adif } x nitera -> diffusion on three scratch
// flux fields (xm, ym, zw) and a ping-pong pair for the iterate.
// One upwind step (:628-677) followed by the mask of smol_adif (:1898-1900), and smol_adif's anti-diffusive
// mass fluxes (:1903-1964), as TMA column kernels (pom_tma.h: tmacolkernel): the iterate and the three mass-flux
// fields of every level staged with the neighbours the fluxes need.  `stale` is what the reference's ff array
// holds where the step does not assign it (boundary columns, level kb).  In the FIRST iteration the mass fluxes
// are those of :602-621 -- 0.25*(dy+dy)*(dt+dt)*u, 0.25*(dx+dx)*(dt+dt)*v, w -- and are formed on the fly from
// the staged u, v, w (same expressions), so that no mass-flux kernel and no round trip of three 3-D arrays is
// needed; smol_adif then leaves its anti-diffusive fluxes in xm, ym, zw for the next iteration.
#ifndef POM_SMOL_TY
#define POM_SMOL_TY 4
#define POM_SMOL_MINB 6
#define POM_SMOL_NS 4
#endif
struct AdvT2UpTK : KBase {
  POM_KINFO("advt2_up", 4, 1, 6, 0)
  const double *fbm_, *f_, *xm_, *ym_, *zw_, *stale_;
  double* ff_;
  int first;
  AdvT2UpTK(const Ctx* x, const double* fbm, const double* f, const double* xm, const double* ym, const double* zw,
            const double* stale, double* ff, int fst)
      : KBase(x), fbm_(fbm), f_(f), xm_(xm), ym_(ym), zw_(zw), stale_(stale), ff_(ff), first(fst) {}
  static constexpr int TY = POM_SMOL_TY, MINB = POM_SMOL_MINB;
  static constexpr int NF = 4, NS = POM_SMOL_NS, OHL = 1, OHR = 1, OHB = 1, OHT = 1, BW = 36, BH = TY + 2, NK = 0;
  static constexpr bool UP = true;
  static constexpr int NVEC = 1;   // (unused)
  enum { FBM, XM, YM, ZW };
  POM_HD void fields(const double** b) const {
    b[FBM] = fbm_; b[XM] = first ? p.u : xm_; b[YM] = first ? p.v : ym_; b[ZW] = first ? p.w : zw_;
  }
  struct State { double m, ar, hb, zk, cxC, cxE, cyC, cyN; RDiv dhf; bool interior; };
  POM_HD int k0() const { return 1; }
  POM_HD int k1() const { return g.kb - 1; }
  POM_HD int kl1() const { return g.kb - 1; }
  template <class CM>
  POM_HD void pre(int i, int j, State& s, CM&) const {
    POM_DIMS;
    s.m=fsm(i,j);
    s.interior = (i >= 2 && i <= imm1 && j >= 2 && j <= jmm1);
    if (!s.interior) return;
    s.ar=art(i,j);
    const double eta=first ? etb(i,j) : etf(i,j);                        // :620,684
    s.hb=(h(i,j)+eta)*s.ar;
    s.dhf.set((h(i,j)+etf(i,j))*s.ar);
    s.zk=first ? w(i,j,1)*A3(f_,i,j,1)*s.ar : 0.;                        // :646-650
    if (first) {                                                         // :605-606, :612-613 at (i,j), (i+1,j), (i,j+1)
      s.cxC=0.25*(dy(i-1,j)+dy(i,j))*(dt(i-1,j)+dt(i,j));
      s.cxE=0.25*(dy(i,j)+dy(i+1,j))*(dt(i,j)+dt(i+1,j));
      s.cyC=0.25*(dx(i,j-1)+dx(i,j))*(dt(i,j-1)+dt(i,j));
      s.cyN=0.25*(dx(i,j)+dx(i,j+1))*(dt(i,j)+dt(i,j+1));
    }
  }
  template <class Op, class CM>
  POM_HD void level(int i, int j, int k, State& s, CM&, const Op& o) const {
    POM_DIMS;
    if (!s.interior) { A3(ff_,i,j,k)=A3(stale_,i,j,k)*s.m; return; }
    const double f0=o(FBM,0,0);
    double zk1 = 0.;                                                     // :651
    if (k + 1 <= kbm1) {
      const double zw1=o.up(ZW);
      zk1=0.5*((zw1+fabs(zw1))*o.up(FBM)+(zw1-fabs(zw1))*f0);            // :656-660
      zk1=zk1*s.ar;                                                      // :661
    }
    const double mE=first ? s.cxE*o(XM,1,0) : o(XM,1,0), mC=first ? s.cxC*o(XM,0,0) : o(XM,0,0);
    const double nN=first ? s.cyN*o(YM,0,1) : o(YM,0,1), nC=first ? s.cyC*o(YM,0,0) : o(YM,0,0);
    const double xE=0.5*((mE+fabs(mE))*f0+(mE-fabs(mE))*o(FBM,1,0));     // :631-635 at i+1
    const double xC=0.5*((mC+fabs(mC))*o(FBM,-1,0)+(mC-fabs(mC))*f0);
    const double yN=0.5*((nN+fabs(nN))*f0+(nN-fabs(nN))*o(FBM,0,1));     // :637-641 at j+1
    const double yC=0.5*((nC+fabs(nC))*o(FBM,0,-1)+(nC-fabs(nC))*f0);
    double q=xE-xC+yN-yC+pdiv(s.zk-zk1,dz(k));                           // :670-672
    q=s.dhf(f0*s.hb-dti2*q);                                             // :673-674
    A3(ff_,i,j,k)=q*s.m;                                                 // smol_adif :1899
    s.zk=zk1;
  }
  template <class CM>
  POM_HD void post(int i, int j, State& s, CM&) const {
    const int kb = g.kb;
    A3(ff_,i,j,kb)=A3(stale_,i,j,kb)*s.m;
  }
};

struct SmolAdifTK : KBase {
  POM_KINFO("smol_adif", 4, 3, 4, 0)
  const double* ff_;
  double *xm_, *ym_, *zw_;
  int first;   // the incoming mass fluxes are those of :602-621, formed here from u, v, w
  SmolAdifTK(const Ctx* x, const double* ff, double* xm, double* ym, double* zw, int fst)
      : KBase(x), ff_(ff), xm_(xm), ym_(ym), zw_(zw), first(fst) {}
  static constexpr int TY = POM_SMOL_TY, MINB = POM_SMOL_MINB;
  static constexpr int NF = 4, NS = POM_SMOL_NS, OHL = 1, OHR = 0, OHB = 1, OHT = 0, BW = 34, BH = TY + 1, NK = 0;
  static constexpr bool UP = false;
  static constexpr int NVEC = 1;   // (unused)
  enum { FF, XM, YM, ZW };
  POM_HD void fields(const double** b) const {
    b[FF] = ff_; b[XM] = first ? p.u : xm_; b[YM] = first ? p.v : ym_; b[ZW] = first ? p.w : zw_;
  }
  struct State { RDiv dax, day; double dtc, fU, cx, cy; bool fx, fy, fz; };
  POM_HD int k0() const { return 1; }
  POM_HD int k1() const { return g.kb - 1; }
  POM_HD int kl1() const { return g.kb - 1; }
  template <class CM>
  POM_HD void pre(int i, int j, State& s, CM&) const {
    POM_DIMS;
    s.fx = (i >= 2 && j >= 2 && j <= jmm1); s.fy = (i >= 2 && i <= imm1 && j >= 2);
    s.fz = (i >= 2 && i <= imm1 && j >= 2 && j <= jmm1);
    s.dax.set(s.fx ? aru(i,j)*(dt(i-1,j)+dt(i,j)) : 1.);
    s.day.set(s.fy ? arv(i,j)*(dt(i,j-1)+dt(i,j)) : 1.);
    s.dtc=dt(i,j);
    s.fU = 0.;
    s.cx = (first && s.fx) ? 0.25*(dy(i-1,j)+dy(i,j))*(dt(i-1,j)+dt(i,j)) : 0.;   // :605-606
    s.cy = (first && s.fy) ? 0.25*(dx(i,j-1)+dx(i,j))*(dt(i,j-1)+dt(i,j)) : 0.;   // :612-613
  }
  template <class Op, class CM>
  POM_HD void level(int i, int j, int k, State& s, CM&, const Op& o) const {
    const double value_min = 1.e-9, epsilon = 1.0e-14;
    const double f0=o(FF,0,0);
    if (s.fx) {                                                          // :1903-1922
      const double fW=o(FF,-1,0), xm=first ? s.cx*o(XM,0,0) : o(XM,0,0);
      const double udx=fabs(xm);
      const double u2dt=s.dax(dti2*xm*xm*2.);
      const double mol=pdiv(f0-fW,fW+f0+epsilon);
      double r=(udx-u2dt)*mol*sw;
      r=(fabs(udx) < fabs(u2dt)) ? 0. : r;
      A3(xm_,i,j,k)=(f0 < value_min || fW < value_min) ? 0. : r;
    }
    if (s.fy) {                                                          // :1924-1943
      const double fS=o(FF,0,-1), ym=first ? s.cy*o(YM,0,0) : o(YM,0,0);
      const double vdy=fabs(ym);
      const double v2dt=s.day(dti2*ym*ym*2.);
      const double mol=pdiv(f0-fS,fS+f0+epsilon);
      double r=(vdy-v2dt)*mol*sw;
      r=(fabs(vdy) < fabs(v2dt)) ? 0. : r;
      A3(ym_,i,j,k)=(f0 < value_min || fS < value_min) ? 0. : r;
    }
    if (s.fz && k >= 2) {                                                // :1945-1964
      const double zw=o(ZW,0,0);
      const double wdz=fabs(zw);
      const double w2dt=pdiv(dti2*zw*zw,dzz(k-1)*s.dtc);
      const double mol=pdiv(s.fU-f0,f0+s.fU+epsilon);
      double r=(wdz-w2dt)*mol*sw;
      r=(fabs(wdz) < fabs(w2dt)) ? 0. : r;
      A3(zw_,i,j,k)=(f0 < value_min || s.fU < value_min) ? 0. : r;
    }
    s.fU=f0;
  }
  template <class CM>
  POM_HD void post(int, int, State&, CM&) const {}
};

// advt1 (solver.f:480-574): centred advection + diffusion of (fb-fclim)
struct AdvT1K : KBase {
  POM_KINFO("advt1", 7, 1, 10, 0)
  const double *fb_, *f_, *fc_;
  double* ff_;
  AdvT1K(const Ctx* x, const double* fb, const double* f, const double* fc, double* ff)
      : KBase(x), fb_(fb), f_(f), fc_(fc), ff_(ff) {}
  POM_HD double fbd(int i, int j, int k) const { return A3(fb_,i,j,k)-A3(fc_,i,j,k); }   // :511
  POM_HD double xfl(int i, int j, int k) const {
    double a=.25*((dt(i,j)+dt(i-1,j))*(A3(f_,i,j,k)+A3(f_,i-1,j,k))*u(i,j,k));          // :502-503
    a=a-.5*(aam(i,j,k)+aam(i-1,j,k))*(h(i,j)+h(i-1,j))*tprni
        *(fbd(i,j,k)-fbd(i-1,j,k))*dum(i,j)/(dx(i,j)+dx(i-1,j));                        // :516-520
    return .5*(dy(i,j)+dy(i-1,j))*a;                                                    // :526
  }
  POM_HD double yfl(int i, int j, int k) const {
    double a=.25*((dt(i,j)+dt(i,j-1))*(A3(f_,i,j,k)+A3(f_,i,j-1,k))*v(i,j,k));          // :504-505
    a=a-.5*(aam(i,j,k)+aam(i,j-1,k))*(h(i,j)+h(i,j-1))*tprni
        *(fbd(i,j,k)-fbd(i,j-1,k))*dvm(i,j)/(dy(i,j)+dy(i,j-1));                        // :521-525
    return .5*(dx(i,j)+dx(i,j-1))*a;                                                    // :527
  }
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    if (!(i >= 2 && i <= imm1 && j >= 2 && j <= jmm1)) return;
    const double ar=art(i,j);
    double zk=A3(f_,i,j,1)*w(i,j,1)*ar;                                 // :537
    for (int k = 1; k <= kbm1; ++k) {
      double zk1 = 0.;                                                  // :538 at kb
      if (k + 1 <= kbm1) zk1=.5*(A3(f_,i,j,k)+A3(f_,i,j,k+1))*w(i,j,k+1)*ar;   // :545
      double r=xfl(i+1,j,k)-xfl(i,j,k)+yfl(i,j+1,k)-yfl(i,j,k)+(zk-zk1)/dz(k);   // :563-565
      // (fb-fclim)+fclim is what fb holds when :566 reads it (:511,:532)
      double fbr=fbd(i,j,k)+A3(fc_,i,j,k);
      A3(ff_,i,j,k)=(fbr*(h(i,j)+etb(i,j))*ar-dti2*r)/((h(i,j)+etf(i,j))*ar);   // :566-570
      zk=zk1;
    }
  }
};

// ---------------------------------------------------------------------------
// proft (solver.f:1541-1683): implicit vertical diffusion of one tracer
struct ProftK : KBase {
  POM_KINFO("proft", 2, 1, 3, 0)
  double* f_;
  const double *wfsurf_, *fsurf_;
  int nbc;
  double rn, ad1n, ad2n;   // Jerlov water type ntp (:1568-1575)
  ProftK(const Ctx* x, double* f, const double* wf, const double* fs, int nb)
      : KBase(x), f_(f), wfsurf_(wf), fsurf_(fs), nbc(nb) {
    const double r[5] = {.58, .62, .67, .77, .78};
    const double ad1[5] = {.35, .60, 1.0, 1.5, 1.4};
    const double ad2[5] = {23., 20., 17., 14., 7.9};
    const int n = (x->c.ntp >= 1 && x->c.ntp <= 5) ? x->c.ntp - 1 : 1;
    rn = r[n]; ad1n = ad1[n]; ad2n = ad2[n];
  }
  // short-wave penetration rad(k) (:1604-1615); rad(kb)=0
  POM_HD double rad(int k, double dh, double sw0) const {
    if (k >= g.kb) return 0.;
    return sw0*(rn*exp(z(k)*dh/ad1n)+(1.-rn)*exp(z(k)*dh/ad2n));
  }
  POM_HD void operator()(int i, int j) const {
    double ee_[KMAX], gg_[KMAX];
#define EE(k) ee_[k]
#define GG(k) gg_[k]
    POM_DIMS;
    const double dh=h(i,j)+etf(i,j);                                    // :1580
    const bool pen = (nbc == 2 || nbc == 4);
    const double sw0 = pen ? A2(p.swrad,i,j) : 0.;
    double radk = 0., radk1 = 0.;                                       // rad(k), rad(k+1)
    // a(k-1), c(k) from kh(k), k=2..kbm1 (:1589-1598); a(kbm1)=0, c(1)=0 (zero fill)
    double khk=POM_LDCS(&kh(i,j,2));
    double ak=-dti2*(khk+umol)/(dz(1)*dzz(1)*dh*dh);                    // a(1)
    double eem, ggm;
    const double f1=POM_LDCS(&A3(f_,i,j,1));
    if (nbc == 1) {                                                     // :1619-1625
      eem=ak/(ak-1.);
      ggm=dti2*A2(wfsurf_,i,j)/(dz(1)*dh)-f1;
      ggm=ggm/(ak-1.);
    } else if (nbc == 2) {                                              // :1629-1637
      radk=rad(1,dh,sw0); radk1=rad(2,dh,sw0);
      eem=ak/(ak-1.);
      ggm=dti2*(A2(wfsurf_,i,j)+radk-radk1)/(dz(1)*dh)-f1;
      ggm=ggm/(ak-1.);
    } else if (nbc == 3 || nbc == 4) {                                  // :1641-1646
      eem=0.;
      ggm=A2(fsurf_,i,j);
    } else {
      eem=0.; ggm=0.;
    }
    EE(1)=eem; GG(1)=ggm;
    if (pen) radk1=rad(2,dh,sw0);
    for (int q = 2; q < 2 + POM_PFD; ++q) { PF3(p.kh,i,j,q+1); PF3(f_,i,j,q); }
    for (int k = 2; k <= kbm2; ++k) {                                   // :1650-1661
      PF3(p.kh,i,j,k+1+POM_PFD); PF3(f_,i,j,k+POM_PFD);
      const double khn=POM_LDCS(&kh(i,j,k+1));
      ak=-dti2*(khn+umol)/(dz(k)*dzz(k)*dh*dh);                         // a(k)
      double ck=-dti2*(khk+umol)/(dz(k)*dzz(k-1)*dh*dh);                // c(k)
      khk=khn;
      double gi=1./(ak+ck*(1.-eem)-1.);
      eem=ak*gi;
      double rhs=ck*ggm-POM_LDCS(&A3(f_,i,j,k));
      if (pen) { radk=radk1; radk1=rad(k+1,dh,sw0); rhs=rhs+dti2*(radk-radk1)/(dh*dz(k)); }
      ggm=rhs*gi;
      EE(k)=eem; GG(k)=ggm;
    }
    {                                                                   // :1664-1671
      double ck=-dti2*(khk+umol)/(dz(kbm1)*dzz(kbm2)*dh*dh);            // c(kbm1), khk=kh(kbm1)
      double rhs=ck*ggm-POM_LDCS(&A3(f_,i,j,kbm1));
      if (pen) { radk=radk1; rhs=rhs+dti2*(radk-0.)/(dh*dz(kbm1)); }
      double fk=rhs/(ck*(1.-eem)-1.);
      POM_STCS(&A3(f_,i,j,kbm1),fk);
      for (int ki = kb-2; ki >= 1; --ki) {                              // :1673-1680
        fk=EE(ki)*fk+GG(ki);
        POM_STCS(&A3(f_,i,j,ki),fk);
      }
    }
#undef EE
#undef GG
  }
};

template <class K>
POM_HD void ts_level(const K& kk, int i, int j, int k, double a, double b, double m, double fold, double fnew, int with_dens);

// proft for T (in uf) and S (in vf) in one TMA-fed column kernel (advance.f:440-441): kh is read
// once, and the matrix coefficients a, c and -- when both tracers have the same class of
// surface condition -- the eliminated ee and the pivots are shared between the two systems.
template <bool SAME>   // both tracers have the same class of surface condition: one ee vector serves both systems
struct ProftTSK : KBase {
  POM_KINFO("proft_ts", 3, 2, 6, 0)
  double rn, ad1n, ad2n;   // Jerlov water type ntp (:1568-1575)
  int fuse;                // the upward sweep also does bcond(4) + t/s filter + restore + dens (advance.f:442-454)
  double fold, fnew;
  ProftTSK(const Ctx* x, int fz = 0) : KBase(x), fuse(fz) {
    const double trst = 30.;                                             // bounds_forcing.f:1033
    int ntime = (int)(x->c.time / trst);
    fnew = x->c.time / trst - ntime;
    fold = 1. - fnew;
    const double r[5] = {.58, .62, .67, .77, .78};
    const double ad1[5] = {.35, .60, 1.0, 1.5, 1.4};
    const double ad2[5] = {23., 20., 17., 14., 7.9};
    const int n = (x->c.ntp >= 1 && x->c.ntp <= 5) ? x->c.ntp - 1 : 1;
    rn = r[n]; ad1n = ad1[n]; ad2n = ad2[n];
  }
  static constexpr int TY = POM_PROFT_TY, MINB = POM_PROFT_MINB;
  static constexpr int NF = 3, NS = 4, OHL = 0, OHR = 0, OHB = 0, OHT = 0, BW = 34, BH = TY, NK = 0;
  static constexpr bool UP = true;
  enum { FT, FS, KH };
  POM_HD void fields(const double** b) const { b[FT] = p.uf; b[FS] = p.vf; b[KH] = p.kh; }
  static constexpr int NVEC = SAME ? 3 : 4;   // ee of T (shared with S when SAME), gg of T, gg of S [, ee of S]
  enum { C_EET, C_GGT, C_GGS, C_EES };
  struct State { double dh, swT, swS, radT, radS, eeT, eeS, ggT, ggS, fT, fS; };
  POM_HD int k0() const { return 1; }
  POM_HD int k1() const { return g.kb - 1; }
  POM_HD int kl1() const { return g.kb - 1; }
  POM_HD double rad(int k, double dh, double sw0) const {                // :1604-1615; rad(kb)=0
    if (k >= g.kb) return 0.;
    return sw0*(rn*exp(z(k)*dh/ad1n)+(1.-rn)*exp(z(k)*dh/ad2n));
  }
  template <class CM>
  POM_HD void pre(int i, int j, State& s, CM&) const {
    s.dh=h(i,j)+etf(i,j);                                               // :1580
    s.swT=(c.nbct == 2 || c.nbct == 4) ? A2(p.swrad,i,j) : 0.;
    s.swS=(c.nbcs == 2 || c.nbcs == 4) ? A2(p.swrad,i,j) : 0.;
    s.radT = 0.; s.radS = 0.;
  }
  // surface condition of one tracer (:1617-1648): ee(1), gg(1)
  POM_HD void surface(int i, int j, int nbc, double ak, double f1, double wfs, double fsf, double sw0, double dh,
                      double& ee1, double& gg1) const {
    if (nbc == 1) {                                                     // :1619-1625
      ee1=ak/(ak-1.);
      gg1=dti2*wfs/(dz(1)*dh)-f1;
      gg1=gg1/(ak-1.);
    } else if (nbc == 2) {                                              // :1629-1637
      ee1=ak/(ak-1.);
      gg1=dti2*(wfs+rad(1,dh,sw0)-rad(2,dh,sw0))/(dz(1)*dh)-f1;
      gg1=gg1/(ak-1.);
    } else if (nbc == 3 || nbc == 4) {                                  // :1641-1646
      ee1=0.;
      gg1=fsf;
    } else {
      ee1=0.; gg1=0.;
    }
  }
  template <class Op, class CM>
  POM_HD void level(int i, int j, int k, State& s, CM& cm, const Op& o) const {
    POM_DIMS;
    const double dh=s.dh;
    const bool penT = (c.nbct == 2 || c.nbct == 4), penS = (c.nbcs == 2 || c.nbcs == 4);
    constexpr bool same = SAME;
    if (k == 1) {
      // a(k-1), c(k) from kh(k), k=2..kbm1 (:1589-1598); a(kbm1)=0, c(1)=0 (zero fill)
      const double ak=-dti2*(o.up(KH)+umol)/(dz(1)*dzz(1)*dh*dh);       // a(1)
      surface(i,j,c.nbct,ak,o(FT,0,0),wtsurf(i,j),tsurf(i,j),s.swT,dh,s.eeT,s.ggT);
      surface(i,j,c.nbcs,ak,o(FS,0,0),wssurf(i,j),ssurf(i,j),s.swS,dh,s.eeS,s.ggS);
      cm.put(C_EET,1,s.eeT); cm.put(C_GGT,1,s.ggT); cm.put(C_GGS,1,s.ggS);
      if (!same) cm.put(C_EES,1,s.eeS);
      if (penT) s.radT=rad(2,dh,s.swT);
      if (penS) s.radS=rad(2,dh,s.swS);
      return;
    }
    const double ck=pdiv(-dti2*(o(KH,0,0)+umol),dz(k)*dzz(k-1)*dh*dh);  // c(k)
    if (k <= kbm2) {                                                    // :1650-1661
      const double ak=pdiv(-dti2*(o.up(KH)+umol),dz(k)*dzz(k)*dh*dh);   // a(k)
      const double giT=pdiv(1.,ak+ck*(1.-s.eeT)-1.);
      const double giS=same ? giT : pdiv(1.,ak+ck*(1.-s.eeS)-1.);
      s.eeT=ak*giT;
      s.eeS=same ? s.eeT : ak*giS;
      double rT=ck*s.ggT-o(FT,0,0), rS=ck*s.ggS-o(FS,0,0);
      if (penT) { const double r0=s.radT; s.radT=rad(k+1,dh,s.swT); rT=rT+dti2*(r0-s.radT)/(dh*dz(k)); }
      if (penS) { const double r0=s.radS; s.radS=rad(k+1,dh,s.swS); rS=rS+dti2*(r0-s.radS)/(dh*dz(k)); }
      s.ggT=rT*giT; s.ggS=rS*giS;
      cm.put(C_EET,k,s.eeT); cm.put(C_GGT,k,s.ggT); cm.put(C_GGS,k,s.ggS);
      if (!same) cm.put(C_EES,k,s.eeS);    // shared with T otherwise: one vector less to park
    } else {                                                            // k == kbm1 (:1664-1671)
      double rT=ck*s.ggT-o(FT,0,0), rS=ck*s.ggS-o(FS,0,0);
      if (penT) rT=rT+dti2*(s.radT-0.)/(dh*dz(kbm1));
      if (penS) rS=rS+dti2*(s.radS-0.)/(dh*dz(kbm1));
      s.fT=rT/(ck*(1.-s.eeT)-1.);
      s.fS=rS/(ck*(1.-s.eeS)-1.);
    }
  }
  template <class CM>
  POM_HD void post(int i, int j, State& s, CM& cm) const {
    POM_DIMS;
    constexpr bool same = SAME;
    double fT=s.fT, fS=s.fS;
    if (fuse) {
      // the new T,S of a level go straight into the filter (which writes uf, vf, tb, sb, rho)
      const double m = fsm(i,j);
      // the eliminated coefficients come back from local memory (L2): fetch them one level ahead
      double eTn=cm.get(C_EET,kb-2), gTn=cm.get(C_GGT,kb-2), gSn=cm.get(C_GGS,kb-2), eSn=same ? eTn : cm.get(C_EES,kb-2);
      ts_level(*this, i, j, kbm1, fT, fS, m, fold, fnew, 1);
      for (int ki = kb-2; ki >= 1; --ki) {                              // :1673-1680
        {   // the filter's operands of the level below, towards L1
          const int o = POM_I3(i,j,ki-1);
          if (ki > 1) { POM_PREFETCH(p.t+o); POM_PREFETCH(p.s+o); POM_PREFETCH(p.tb+o); POM_PREFETCH(p.sb+o);
                        POM_PREFETCH(p.tclim+o); POM_PREFETCH(p.sclim+o); }
        }
        const double eT=eTn, gT=gTn, eS=eSn, gS=gSn;
        if (ki > 1) { eTn=cm.get(C_EET,ki-1); gTn=cm.get(C_GGT,ki-1); gSn=cm.get(C_GGS,ki-1); eSn=same ? eTn : cm.get(C_EES,ki-1); }
        fT=eT*fT+gT;
        fS=eS*fS+gS;
        ts_level(*this, i, j, ki, fT, fS, m, fold, fnew, 1);
      }
      return;
    }
    POM_STCS(&uf(i,j,kbm1),fT);
    POM_STCS(&vf(i,j,kbm1),fS);
    for (int ki = kb-2; ki >= 1; --ki) {                                // :1673-1680
      const double eT=cm.get(C_EET,ki);
      fT=eT*fT+cm.get(C_GGT,ki);
      fS=(same ? eT : cm.get(C_EES,ki))*fS+cm.get(C_GGS,ki);
      POM_STCS(&uf(i,j,ki),fT);
      POM_STCS(&vf(i,j,ki),fS);
    }
  }
};

// the fused variant under its own name / algorithmic byte count in the per-kernel profile
template <bool SAME>
struct ProftFilterK : ProftTSK<SAME> {
  POM_KINFO("proft_tsfilter", 9, 5, 9, 0)
  ProftFilterK(const Ctx* x) : ProftTSK<SAME>(x, 1) {}
};

// x**1.5 for x>=0, rounded like a correctly-rounded pow(x,1.5): sqrt is IEEE-exact to
// 0.5 ulp, its residual and the product are recovered exactly with fma.
POM_HD double pow15(double x) {
  if (!(x > 0.)) return 0.;
  double sq=sqrt(x);
  double res=fma(-sq,sq,x);          // x - sq*sq exactly
  double ds=pdiv(res,2.*sq);             // sqrt(x) = sq + ds
  double pr=x*sq;
  double er=fma(x,sq,-pr);           // x*sq = pr + er exactly
  return pr+(er+x*ds);
}

// the equation of state of one cell (solver.f:1175-1203): T, S (+bias), pressure from -zz*h
POM_HD double dens_point(double tr, double sr, double pp, double rhoref_, double m) {
  double tr2=tr*tr, tr3=tr2*tr, tr4=tr3*tr;
  double rhor=-0.157406+6.793952e-2*tr-9.095290e-3*tr2+1.001685e-4*tr3
              -1.120083e-6*tr4+6.536332e-9*tr4*tr;
  rhor=rhor+(0.824493-4.0899e-3*tr+7.6438e-5*tr2-8.2467e-7*tr3+5.3875e-9*tr4)*sr
           +(-5.72466e-3+1.0227e-4*tr-1.6546e-6*tr2)*pow15(fabs(sr))
           +4.8314e-4*sr*sr;
  double cr=1449.1+.0821*pp+4.55*tr-.045*tr2+1.34*(sr-35.);
  rhor=rhor+pdiv(1.e5*pp,cr*cr)*(1.-pdiv(2.*pp,cr*cr));
  return pdiv(rhor,rhoref_)*m;
}

// One level of bcond(4) (bounds_forcing.f:151-242) + Asselin filter/rotation of t,s
// (advance.f:444-449) + restore_interior arithmetic and mask (bounds_forcing.f:1083-1118)
// [+ dens(s,t,rho) of the new t,s (advance.f:454)] for the new values a (T), b (S) that proft
// left at (i,j,k).  Shared by the stand-alone filter kernel and by proft's fused upward sweep.
template <class K>
POM_HD void ts_level(const K& kk, int i, int j, int k, double a, double b, double m, double fold, double fnew, int with_dens) {
  const Geo& g = kk.g; const Ptrs& p = kk.p; const Consts& c = kk.c; const KTab& kt = kk.kt; (void)kt;
  POM_DIMS;
    bcond4_edge(kk, i, j, k, a, b);                                   // bounds_forcing.f:151-231
    a=a*m;                                                            // :236-237
    b=b*m;
    // tb,sb as left behind by advt1/advt2: (fb-fclim)+fclim (solver.f:691,715)
    double tbr=(tb(i,j,k)-tclim(i,j,k))+tclim(i,j,k);
    double sbr=(sb(i,j,k)-sclim(i,j,k))+sclim(i,j,k);
    double tf=t(i,j,k)+.5*smoth*(a+tbr-2.*t(i,j,k));                  // advance.f:444
    double sf=s(i,j,k)+.5*smoth*(b+sbr-2.*s(i,j,k));                  // advance.f:445
    if (c.lrestore) {                                                 // bounds_forcing.f:1083-1110
      double tr=fold*trstrb(i,j,k)+fnew*trstrf(i,j,k);
      double sr=fold*srstrb(i,j,k)+fnew*srstrf(i,j,k);
      double ta=fold*taurstrb(i,j,k)+fnew*taurstrf(i,j,k);
      a=a+2.*dti/86400.*ta*(tr-a);
      tf=tf+2.*dti/86400.*ta*(tr-tf);
      b=b+2.*dti/86400.*ta*(sr-b);
      sf=sf+2.*dti/86400.*ta*(sr-sf);
    }
    a=a*m; b=b*m;                                                     // :1113-1118
    uf(i,j,k)=a;
    vf(i,j,k)=b;
    tb(i,j,k)=tf*m;
    sb(i,j,k)=sf*m;
    if (with_dens)                                                    // solver.f:1175-1205
      rho(i,j,k)=dens_point(a+tbias,b+sbias,grav*rhoref*(-zz(k)*h(i,j))*1.e-5,rhoref,m);
}

// ---------------------------------------------------------------------------
// bcond(4) + filter of t,s as a stand-alone kernel (the un-fused stage hook); new t,s stay in
// uf,vf, filtered t,s go to tb,sb; the host rotates pointers.
struct TsFilterK : KBase {
  POM_KINFO("ts_filter", 8, 5, 3, 0)
  double fold, fnew;
  int with_dens;   // also dens(s,t,rho) of the new t,s (advance.f:454), saving a pass over both
  TsFilterK(const Ctx* x, int wd) : KBase(x), with_dens(wd) {
    const double trst = 30.;                                             // bounds_forcing.f:1033
    int ntime = (int)(x->c.time / trst);
    fnew = x->c.time / trst - ntime;
    fold = 1. - fnew;
  }
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    const double m = fsm(i,j);
    for (int k = 1; k <= kbm1; ++k) ts_level(*this, i, j, k, uf(i,j,k), vf(i,j,k), m, fold, fnew, with_dens);
  }
};

// dens (solver.f:1162-1209)
struct DensK : KBase {
  POM_KINFO("dens", 2, 1, 2, 0)
  const double *si_, *ti_;
  double* ro_;
  DensK(const Ctx* x, const double* si, const double* ti, double* ro) : KBase(x), si_(si), ti_(ti), ro_(ro) {}
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    const double m=fsm(i,j), hh=h(i,j);
    for (int k = 1; k <= kbm1; ++k) {
      PF3(ti_,i,j,k+2); PF3(si_,i,j,k+2);
      const double pp=grav*rhoref*(-zz(k)*hh)*1.e-5;
      A3(ro_,i,j,k)=dens_point(A3(ti_,i,j,k)+tbias,A3(si_,i,j,k)+sbias,pp,rhoref,m);
    }
  }
};

// ---------------------------------------------------------------------------
// advu / advv (solver.f:734-845)
struct AdvuK : KBase {
  POM_KINFO("advu", 6, 1, 10, 0)
  using KBase::KBase;
  POM_HD double vfl(int i, int j, int k) const {                        // :747-748
    return .25*(w(i,j,k)+w(i-1,j,k))*(u(i,j,k)+u(i,j,k-1));
  }
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    if (i >= 2 && i <= imm1 && j >= 2 && j <= jmm1) {
      const double ar=aru(i,j);
      const double sl=grav*.125*(dt(i,j)+dt(i-1,j))
                      *(egf(i,j)-egf(i-1,j)+egb(i,j)-egb(i-1,j)
                        +(e_atmos(i,j)-e_atmos(i-1,j))*2.)
                      *(dy(i,j)+dy(i-1,j));                             // :765-768
      const double hb=(h(i,j)+etb(i,j)+h(i-1,j)+etb(i-1,j))*ar;
      const double hf=(h(i,j)+etf(i,j)+h(i-1,j)+etf(i-1,j))*ar;
      double fk = 0.;                                                   // uf(i,j,1)=0 (:742)
      for (int k = 1; k <= kbm1; ++k) {
        PF3(p.w,i,j,k+3); PF3(p.u,i,j,k+3); PF3(p.v,i,j,k+2); PF3(p.v,i,j+1,k+2);
        PF3(p.advx,i,j,k+2); PF3(p.drhox,i,j,k+2); PF3(p.ub,i,j,k+2);
        double fk1 = (k + 1 <= kbm1) ? vfl(i,j,k+1) : 0.;
        double r=advx(i,j,k)+(fk-fk1)*ar/dz(k)
                 -ar*.25*(cor(i,j)*dt(i,j)*(v(i,j+1,k)+v(i,j,k))
                          +cor(i-1,j)*dt(i-1,j)*(v(i-1,j+1,k)+v(i-1,j,k)))
                 +sl+drhox(i,j,k);                                      // :758-769
        uf(i,j,k)=(hb*ub(i,j,k)-2.*dti2*r)/hf;                          // :778-782
        fk=fk1;
      }
      uf(i,j,kb)=0.;
    } else {
      // outside the interior only the vertical-flux intermediate survives (:744-751)
      uf(i,j,1)=0.; uf(i,j,kb)=0.;
      for (int k = 2; k <= kbm1; ++k) uf(i,j,k)=(i >= 2) ? vfl(i,j,k) : 0.;
    }
  }
};

struct AdvvK : KBase {
  POM_KINFO("advv", 6, 1, 10, 0)
  using KBase::KBase;
  POM_HD double vfl(int i, int j, int k) const {                        // :804-805
    return .25*(w(i,j,k)+w(i,j-1,k))*(v(i,j,k)+v(i,j,k-1));
  }
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    if (i >= 2 && i <= imm1 && j >= 2 && j <= jmm1) {
      const double ar=arv(i,j);
      const double sl=grav*.125*(dt(i,j)+dt(i,j-1))
                      *(egf(i,j)-egf(i,j-1)+egb(i,j)-egb(i,j-1)
                        +(e_atmos(i,j)-e_atmos(i,j-1))*2.)
                      *(dx(i,j)+dx(i,j-1));                             // :822-825
      const double hb=(h(i,j)+etb(i,j)+h(i,j-1)+etb(i,j-1))*ar;
      const double hf=(h(i,j)+etf(i,j)+h(i,j-1)+etf(i,j-1))*ar;
      double fk = 0.;
      for (int k = 1; k <= kbm1; ++k) {
        PF3(p.w,i,j,k+3); PF3(p.w,i,j-1,k+3); PF3(p.v,i,j,k+3); PF3(p.u,i,j,k+2); PF3(p.u,i,j-1,k+2);
        PF3(p.advy,i,j,k+2); PF3(p.drhoy,i,j,k+2); PF3(p.vb,i,j,k+2);
        double fk1 = (k + 1 <= kbm1) ? vfl(i,j,k+1) : 0.;
        double r=advy(i,j,k)+(fk-fk1)*ar/dz(k)
                 +ar*.25*(cor(i,j)*dt(i,j)*(u(i+1,j,k)+u(i,j,k))
                          +cor(i,j-1)*dt(i,j-1)*(u(i+1,j-1,k)+u(i,j-1,k)))
                 +sl+drhoy(i,j,k);                                      // :815-826
        vf(i,j,k)=(hb*vb(i,j,k)-2.*dti2*r)/hf;                          // :835-839
        fk=fk1;
      }
      vf(i,j,kb)=0.;
    } else {
      vf(i,j,1)=0.; vf(i,j,kb)=0.;
      for (int k = 2; k <= kbm1; ++k) vf(i,j,k)=(j >= 2) ? vfl(i,j,k) : 0.;
    }
  }
};

// ---------------------------------------------------------------------------
// profu / profv (solver.f:1686-1877): implicit vertical viscosity + bottom drag
struct ProfuK : KBase {
  POM_KINFO("profu", 2, 1, 5, 1)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    if (!(i >= 2 && i <= imm1 && j >= 2 && j <= jmm1)) return;
    double ee[KMAX], gg[KMAX];
    const double dh=(h(i,j)+etf(i,j)+h(i-1,j)+etf(i-1,j))*.5;           // :1703
    double cn=(km(i,j,2)+km(i-1,j,2))*.5;                               // :1715 at k=2
    double ak=-dti2*(cn+umol)/(dz(1)*dzz(1)*dh*dh);                     // a(1) (:1723)
    ee[1]=ak/(ak-1.);                                                   // :1733
    gg[1]=(-dti2*wusurf(i,j)/(-dz(1)*dh)-uf(i,j,1))/(ak-1.);            // :1734-1736
    double ck=-dti2*(cn+umol)/(dz(2)*dzz(1)*dh*dh);                     // c(2) (:1725)
    for (int k = 2; k <= kbm2; ++k) {                                   // :1740-1748
      PF3(p.km,i,j,k+3); PF3(p.uf,i,j,k+2);
      cn=(km(i,j,k+1)+km(i-1,j,k+1))*.5;
      ak=-dti2*(cn+umol)/(dz(k)*dzz(k)*dh*dh);                          // a(k)
      double gi=1./(ak+ck*(1.-ee[k-1])-1.);
      ee[k]=ak*gi;
      gg[k]=(ck*gg[k-1]-uf(i,j,k))*gi;
      ck=-dti2*(cn+umol)/(dz(k+1)*dzz(k)*dh*dh);                        // c(k+1)
    }
    // ck == c(kbm1)
    double vbar=.25*(vb(i,j,kbm1)+vb(i,j+1,kbm1)+vb(i-1,j,kbm1)+vb(i-1,j+1,kbm1));
    double tp=0.5*(cbc(i,j)+cbc(i-1,j))
              *sqrt(ub(i,j,kbm1)*ub(i,j,kbm1)+vbar*vbar);               // :1752-1755
    double fk=(ck*gg[kbm2]-uf(i,j,kbm1))
              /(tp*dti2/(-dz(kbm1)*dh)-1.-(ee[kbm2]-1.)*ck);            // :1756-1758
    const double m=dum(i,j);
    fk=fk*m;                                                            // :1759
    uf(i,j,kbm1)=fk;
    wubot(i,j)=-tp*fk;                                                  // :1774
    for (int ki = kb-2; ki >= 1; --ki) {                                // :1763-1770
      fk=(ee[ki]*fk+gg[ki])*m;
      uf(i,j,ki)=fk;
    }
  }
};

struct ProfvK : KBase {
  POM_KINFO("profv", 2, 1, 5, 1)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    if (!(i >= 2 && i <= imm1 && j >= 2 && j <= jmm1)) return;
    double ee[KMAX], gg[KMAX];
    const double dh=.5*(h(i,j)+etf(i,j)+h(i,j-1)+etf(i,j-1));           // :1801
    double cn=(km(i,j,2)+km(i,j-1,2))*.5;                               // :1813
    double ak=-dti2*(cn+umol)/(dz(1)*dzz(1)*dh*dh);                     // :1821
    ee[1]=ak/(ak-1.);                                                   // :1831
    gg[1]=(-dti2*wvsurf(i,j)/(-dz(1)*dh)-vf(i,j,1))/(ak-1.);            // :1832-1833
    double ck=-dti2*(cn+umol)/(dz(2)*dzz(1)*dh*dh);                     // :1823
    for (int k = 2; k <= kbm2; ++k) {                                   // :1837-1845
      PF3(p.km,i,j,k+3); PF3(p.km,i,j-1,k+3); PF3(p.vf,i,j,k+2);
      cn=(km(i,j,k+1)+km(i,j-1,k+1))*.5;
      ak=-dti2*(cn+umol)/(dz(k)*dzz(k)*dh*dh);
      double gi=1./(ak+ck*(1.-ee[k-1])-1.);
      ee[k]=ak*gi;
      gg[k]=(ck*gg[k-1]-vf(i,j,k))*gi;
      ck=-dti2*(cn+umol)/(dz(k+1)*dzz(k)*dh*dh);
    }
    double ubar=.25*(ub(i,j,kbm1)+ub(i+1,j,kbm1)+ub(i,j-1,kbm1)+ub(i+1,j-1,kbm1));
    double tp=0.5*(cbc(i,j)+cbc(i,j-1))
              *sqrt(ubar*ubar+vb(i,j,kbm1)*vb(i,j,kbm1));               // :1849-1852
    double fk=(ck*gg[kbm2]-vf(i,j,kbm1))
              /(tp*dti2/(-dz(kbm1)*dh)-1.-(ee[kbm2]-1.)*ck);            // :1853-1855
    const double m=dvm(i,j);
    fk=fk*m;                                                            // :1856
    vf(i,j,kbm1)=fk;
    wvbot(i,j)=-tp*fk;                                                  // :1871
    for (int ki = kb-2; ki >= 1; --ki) {                                // :1860-1867
      fk=(ee[ki]*fk+gg[ki])*m;
      vf(i,j,ki)=fk;
    }
  }
};

// ---------------------------------------------------------------------------
// advu + profu (VC=false) / advv + profv (VC=true) in one TMA-fed column kernel: the explicit
// tendency uf(k) of solver.f:734-788 (791-845) is consumed by the forward elimination of
// solver.f:1686-1780 (1783-1877) at the same level, so it never makes the round trip through
// HBM.  B = the "back" neighbour (i-1 for u, j-1 for v), O = the other direction (j+1 / i+1).
template <bool VC>
struct AdvProfUVK : KBase {
  static const KInfo& info() {
    static const KInfo ku{"advu_profu", 7, 1, 14, 1}, kv{"advv_profv", 7, 1, 14, 1};
    return VC ? kv : ku;
  }
  using KBase::KBase;
  static constexpr int TY = POM_ADVPROF_TY, MINB = POM_ADVPROF_MINB;
  static constexpr int NF = 7, NS = POM_ADVPROF_NS, OHL = 1, OHR = 1, OHB = 1, OHT = 1, BW = 36, BH = TY + 2, NK = 0;
  static constexpr bool UP = true;
  static constexpr int BI = VC ? 0 : -1, BJ = VC ? -1 : 0, OI = VC ? 1 : 0, OJ = VC ? 0 : 1;
  enum { W, X, Y, ADV, DRHO, XB, KM };
  POM_HD void fields(const double** b) const {
    b[W] = p.w; b[X] = VC ? p.v : p.u; b[Y] = VC ? p.u : p.v; b[ADV] = VC ? p.advy : p.advx;
    b[DRHO] = VC ? p.drhoy : p.drhox; b[XB] = VC ? p.vb : p.ub; b[KM] = p.km;
  }
  POM_HD double* xf() const { return VC ? p.vf : p.uf; }
  static constexpr int NVEC = 2;
  enum { C_EE, C_GG };
  struct State { double ar, sl, hb, hf, c0, cB, dh, fk, xm, eem, ggm, ck, xfl; bool interior, live; };
  POM_HD int k0() const { return 1; }
  POM_HD int k1() const { return g.kb - 1; }
  POM_HD int kl1() const { return g.kb - 1; }
  template <class CM>
  POM_HD void pre(int i, int j, State& s, CM&) const {
    POM_DIMS;
    s.interior = (i >= 2 && i <= imm1 && j >= 2 && j <= jmm1);
    s.live = VC ? (j >= 2) : (i >= 2);          // the vertical-flux intermediate exists (:744-751, :801-808)
    s.fk = 0.; s.xm = 0.;
    double* f = xf();
    A3(f,i,j,kb)=0.;
    if (!s.interior) { A3(f,i,j,1)=0.; return; }
    const int ib = i + BI, jb = j + BJ;
    s.ar = VC ? arv(i,j) : aru(i,j);
    const double dd = VC ? (dx(i,j)+dx(ib,jb)) : (dy(i,j)+dy(ib,jb));
    s.sl=grav*.125*(dt(i,j)+dt(ib,jb))
         *(egf(i,j)-egf(ib,jb)+egb(i,j)-egb(ib,jb)
           +(e_atmos(i,j)-e_atmos(ib,jb))*2.)
         *dd;                                                           // :765-768 / :822-825
    s.hb=(h(i,j)+etb(i,j)+h(ib,jb)+etb(ib,jb))*s.ar;
    s.hf=(h(i,j)+etf(i,j)+h(ib,jb)+etf(ib,jb))*s.ar;
    s.c0=cor(i,j)*dt(i,j); s.cB=cor(ib,jb)*dt(ib,jb);
    s.dh=(h(i,j)+etf(i,j)+h(ib,jb)+etf(ib,jb))*.5;                      // :1703 / :1801
  }
  template <class Op, class CM>
  POM_HD void level(int i, int j, int k, State& s, CM& cm, const Op& o) const {
    POM_DIMS;
    double* f = xf();
    const double x0=o(X,0,0);
    if (!s.interior) {
      // outside the interior only the vertical-flux intermediate survives (:744-751)
      if (k >= 2) A3(f,i,j,k)=s.live ? .25*(o(W,0,0)+o(W,BI,BJ))*(x0+s.xm) : 0.;
      s.xm=x0;
      return;
    }
    // ---- explicit tendency (advu :753-785 / advv :810-842) ----
    const double fk1 = (k + 1 <= kbm1) ? .25*(o.up(W)+o.up(W,BI,BJ))*(o.up(X)+x0) : 0.;   // :747-748
    const double ct=s.c0*(o(Y,OI,OJ)+o(Y,0,0))+s.cB*(o(Y,BI+OI,BJ+OJ)+o(Y,BI,BJ));
    double r=o(ADV,0,0)+pdiv((s.fk-fk1)*s.ar,dz(k));
    r = VC ? r+s.ar*.25*ct : r-s.ar*.25*ct;                             // :760-763 / :817-820
    r=r+s.sl+o(DRHO,0,0);                                               // :764-769
    const double xk=pdiv(s.hb*o(XB,0,0)-2.*dti2*r,s.hf);                // :778-782
    s.fk=fk1;
    // ---- forward elimination (profu :1711-1748 / profv :1809-1845) ----
    const double dh=s.dh;
    if (k == 1) {
      const double cn=(o.up(KM)+o.up(KM,BI,BJ))*.5;                     // :1715 at k=2
      const double ak=-dti2*(cn+umol)/(dz(1)*dzz(1)*dh*dh);             // a(1) (:1723)
      const double ws = VC ? wvsurf(i,j) : wusurf(i,j);
      s.eem=ak/(ak-1.);                                                 // :1733
      s.ggm=(-dti2*ws/(-dz(1)*dh)-xk)/(ak-1.);                          // :1734-1736
      s.ck=-dti2*(cn+umol)/(dz(2)*dzz(1)*dh*dh);                        // c(2) (:1725)
      cm.put(C_EE,1,s.eem); cm.put(C_GG,1,s.ggm);
    } else if (k <= kbm2) {                                             // :1740-1748
      const double cn=(o.up(KM)+o.up(KM,BI,BJ))*.5;
      const double ak=pdiv(-dti2*(cn+umol),dz(k)*dzz(k)*dh*dh);         // a(k)
      const double gi=pdiv(1.,ak+s.ck*(1.-s.eem)-1.);
      s.eem=ak*gi;
      s.ggm=(s.ck*s.ggm-xk)*gi;
      s.ck=pdiv(-dti2*(cn+umol),dz(k+1)*dzz(k)*dh*dh);                  // c(k+1)
      cm.put(C_EE,k,s.eem); cm.put(C_GG,k,s.ggm);
    } else {
      s.xfl=xk;                                                         // uf(kbm1), used by the bottom condition
    }
  }
  template <class CM>
  POM_HD void post(int i, int j, State& s, CM& cm) const {
    POM_DIMS;
    if (!s.interior) return;
    double* f = xf();
    const int ib = i + BI, jb = j + BJ;
    const double* yb = VC ? p.ub : p.vb;
    const double* xb = VC ? p.vb : p.ub;
    const double ybar=.25*(A3(yb,i,j,kbm1)+A3(yb,i+OI,j+OJ,kbm1)+A3(yb,ib,jb,kbm1)+A3(yb,ib+OI,jb+OJ,kbm1));
    const double xo=A3(xb,i,j,kbm1);
    const double sp = VC ? sqrt(ybar*ybar+xo*xo) : sqrt(xo*xo+ybar*ybar);
    const double tp=0.5*(cbc(i,j)+cbc(ib,jb))*sp;                       // :1752-1755 / :1849-1852
    double fk=(s.ck*s.ggm-s.xfl)
              /(tp*dti2/(-dz(kbm1)*s.dh)-1.-(s.eem-1.)*s.ck);           // :1756-1758
    const double m = VC ? dvm(i,j) : dum(i,j);
    fk=fk*m;                                                            // :1759
    A3(f,i,j,kbm1)=fk;
    if (VC) wvbot(i,j)=-tp*fk; else wubot(i,j)=-tp*fk;                  // :1774 / :1871
    for (int ki = kb-2; ki >= 1; --ki) {                                // :1763-1770
      fk=(cm.get(C_EE,ki)*fk+cm.get(C_GG,ki))*m;
      A3(f,i,j,ki)=fk;
    }
  }
};

// ---------------------------------------------------------------------------
// bcondorl(3) (bounds_forcing.f:418-487) + Asselin filter with depth-mean
// removal and rotation of u,v (advance.f:469-514).  The Orlanski points read ub,vb
// of their neighbours, so the filtered u,v go to the scratch buffers s3a,s3b (which
// become ub,vb); new u,v stay in uf,vf; the host rotates pointers.
struct UvFilterK : KBase {
  POM_KINFO("uv_filter", 6, 2, 2, 2)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    const bool jin = (j >= 2 && j <= jmm1), iin = (i >= 2 && i <= imm1);
    const double mu=dum(i,j), mv=dvm(i,j);
    double su = 0., sv = 0., tu = 0., tv = 0.;
    double nu[KMAX], nv[KMAX];
    const bool edge0 = !(iin && jin) || i == 2 || j == 2;
    for (int k = 1; k <= kbm1; ++k) {
      PF3(p.uf,i,j,k+2); PF3(p.vf,i,j,k+2); PF3(p.ub,i,j,k+2); PF3(p.vb,i,j,k+2); PF3(p.u,i,j,k+2); PF3(p.v,i,j,k+2);
      double a=uf(i,j,k), b=vf(i,j,k);
      if (edge0) bcondorl3_edge(*this, i, j, k, a, b);                    // bounds_forcing.f:418-474
      a=a*mu;                                                           // :481-482
      b=b*mv;
      if (edge0) { nu[k]=a; nv[k]=b; }
      su=su+(a+ub(i,j,k)-2.*u(i,j,k))*dz(k);                            // advance.f:474-475
      sv=sv+(b+vb(i,j,k)-2.*v(i,j,k))*dz(k);                            // advance.f:495-496
      tu=tu+a*dz(k);                                                    // next step's advance.f:367-369:
      tv=tv+b*dz(k);                                                    // uf, vf become u, v (:512,514)
    }
    A2(p.s2c,i,j)=tu;
    A2(p.s2d,i,j)=tv;
    const bool edge = !(iin && jin) || i == 2 || j == 2;
    for (int k = 1; k <= kbm1; ++k) {
      // away from the open boundaries the masked tendency is recomputed from uf,vf (one more
      // read) instead of making the round trip through the per-thread arrays (a write + a read)
      double a = edge ? nu[k] : uf(i,j,k)*mu, b = edge ? nv[k] : vf(i,j,k)*mv;
      double un=u(i,j,k)+.5*smoth*(a+ub(i,j,k)-2.*u(i,j,k)-su);         // advance.f:483-485
      double vn=v(i,j,k)+.5*smoth*(b+vb(i,j,k)-2.*v(i,j,k)-sv);         // advance.f:504-506
      A3(p.s3a,i,j,k)=un;
      A3(p.s3b,i,j,k)=vn;
      if (edge) { uf(i,j,k)=a; vf(i,j,k)=b; }   // interior values are already masked (profu :1767)
    }
    A3(p.s3a,i,j,kb)=u(i,j,kb);                                         // advance.f:511,513
    A3(p.s3b,i,j,kb)=v(i,j,kb);
  }
};

// what is left of a window [jw0,jw1] around the parked kernel's rectangle [3,im-1] x [ja0,ja1]: whole rows
// below and above it, and the columns i = 1, 2, im beside it, enumerated through a virtual (i, row) space
struct UvFilterFrameK : UvFilterK {
  int jw0, ja0, ja1, nlow, nfull, nin;
  UvFilterFrameK(const Ctx* x, int w0, int w1, int a0, int a1) : UvFilterK(x), jw0(w0), ja0(a0), ja1(a1) {
    nlow = a0 - w0; nfull = nlow + (w1 - a1); nin = a1 - a0 + 1;
  }
  POM_HD int rows() const { return nfull + (3 * nin + g.im - 1) / g.im; }
  POM_HD bool cell(int vi, int vj, int& i, int& j) const {
    if (vj <= nlow) { i = vi; j = jw0 + vj - 1; return true; }
    if (vj <= nfull) { i = vi; j = ja1 + (vj - nlow); return true; }
    const int t = (vj - nfull - 1) * g.im + (vi - 1);
    if (t >= 3 * nin) return false;
    const int col = t / nin;
    i = col == 0 ? 1 : (col == 1 ? 2 : g.im); j = ja0 + t % nin;
    return true;
  }
  POM_HD void operator()(int vi, int vj) const {
    int i, j;
    if (cell(vi, vj, i, j)) UvFilterK::operator()(i, j);
  }
#if !defined(POMGPU_EMU) && defined(__CUDACC__)
  // One WARP per frame column, lane = level (k = lane+1, lane+33): every operand of the column is loaded at
  // once instead of along a chain of 2 x (kb-1) dependent iterations; the depth sums are then accumulated in
  // the reference's order k = 1, 2, ... from the lanes' products (shuffle broadcasts), so every bit is the
  // two-sweep kernel's.  All frame columns are Orlanski/edge columns (edge0 of UvFilterK).
  __device__ void warp_column(int i, int j, int lane) const {
    POM_DIMS;
    const double mu=dum(i,j), mv=dvm(i,j);
    double a[2], b[2], uo[2], vo[2], xu[2], xv[2], pu[2], pv[2], qu[2], qv[2];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      const int k = lane + 1 + 32 * sl;
      a[sl] = b[sl] = uo[sl] = vo[sl] = xu[sl] = xv[sl] = pu[sl] = pv[sl] = qu[sl] = qv[sl] = 0.;
      if (k <= kbm1) {
        double aa=uf(i,j,k), bb=vf(i,j,k);
        bcondorl3_edge(*this, i, j, k, aa, bb);                           // bounds_forcing.f:418-474
        aa=aa*mu;                                                         // :481-482
        bb=bb*mv;
        const double uu=u(i,j,k), vv=v(i,j,k);
        const double x=aa+ub(i,j,k)-2.*uu, y=bb+vb(i,j,k)-2.*vv;
        a[sl]=aa; b[sl]=bb; uo[sl]=uu; vo[sl]=vv; xu[sl]=x; xv[sl]=y;
        pu[sl]=x*dz(k); pv[sl]=y*dz(k); qu[sl]=aa*dz(k); qv[sl]=bb*dz(k);
      }
    }
    double su = 0., sv = 0., tu = 0., tv = 0.;
    for (int k = 1; k <= kbm1; ++k) {                                     // advance.f:474-475, 495-496
      const int src = (k - 1) & 31;
      const bool hi = k > 32;
      su=su+__shfl_sync(0xffffffffu, hi ? pu[1] : pu[0], src);
      sv=sv+__shfl_sync(0xffffffffu, hi ? pv[1] : pv[0], src);
      tu=tu+__shfl_sync(0xffffffffu, hi ? qu[1] : qu[0], src);
      tv=tv+__shfl_sync(0xffffffffu, hi ? qv[1] : qv[0], src);
    }
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      const int k = lane + 1 + 32 * sl;
      if (k <= kbm1) {
        A3(p.s3a,i,j,k)=uo[sl]+.5*smoth*(xu[sl]-su);                      // advance.f:483-485
        A3(p.s3b,i,j,k)=vo[sl]+.5*smoth*(xv[sl]-sv);                      // advance.f:504-506
        uf(i,j,k)=a[sl]; vf(i,j,k)=b[sl];
      }
    }
    if (lane == 0) {
      A2(p.s2c,i,j)=tu; A2(p.s2d,i,j)=tv;
      A3(p.s3a,i,j,kb)=u(i,j,kb);                                         // advance.f:511,513
      A3(p.s3b,i,j,kb)=v(i,j,kb);
    }
  }
#endif
};

// The interior columns of the same filter (3<=i<=im-1, 3<=j<=jm-1: no Orlanski point, uf, vf already
// masked by profu/profv) with the six operands read ONCE: sweep 1 parks (uf*mask+ub-2u, u) of every level
// in shared memory while it forms the depth sums, sweep 2 finishes from the parked values
// (pom_tma.h: tmaparkkernel; one component per tile).  Same expressions in the same order as UvFilterK.  The
// frame around the rectangle is worked off by the side warps of the same launch.
struct UvFilterParkK : UvFilterFrameK {
  POM_KINFO("uv_filter", 6, 2, 2, 2)
  using UvFilterFrameK::UvFilterFrameK;
  static constexpr int NC = 2, NF = 3, NP = 2;
  static constexpr int NSIDE = 2;   // warps per block that work the frame off while the rectangle streams
  double columns(int i0, int i1, int j0, int j1) const {
    return (double)(i1 - i0 + 1) * (j1 - j0 + 1) + (double)nfull * g.im + 3. * nin;
  }
#if !defined(POMGPU_EMU) && defined(__CUDACC__)
  __device__ void side(int w, int nw, int lane) const {
    const int nv = rows() * g.im;
    for (int q = w; q < nv; q += nw) {
      int i, j;
      if (cell(q % g.im + 1, q / g.im + 1, i, j)) warp_column(i, j, lane);
    }
  }
#endif
  void side_host() const {
    const int nr = rows();
    for (int vj = 1; vj <= nr; ++vj)
      for (int vi = 1; vi <= g.im; ++vi) UvFilterFrameK::operator()(vi, vj);
  }
  POM_HD void fields(int comp, const double** b) const {
    b[0] = comp ? p.vf : p.uf; b[1] = comp ? p.vb : p.ub; b[2] = comp ? p.v : p.u;
  }
  struct State { double m, su, tu; };
  static constexpr int NPL = 2;   // the column's mask and its bottom-level velocity
  POM_HD void preload(int comp, int i, int j, double* pl) const {
    const int kb = g.kb;
    pl[0] = comp ? dvm(i,j) : dum(i,j);
    pl[1] = comp ? v(i,j,kb) : u(i,j,kb);
  }
  POM_HD void pre(int, int, int, State& s, const double* pl) const { s.m = pl[0]; s.su = 0.; s.tu = 0.; }
  POM_HD const double* ktab() const { return p.dz; }
  POM_HD void sweep1(int, int, int, int, State& s, const double* f, double* pk, double dzk) const {
    const double a=f[0]*s.m;                                            // bounds_forcing.f:481-482
    const double x=a+f[1]-2.*f[2];
    s.su=s.su+x*dzk;                                                    // advance.f:474-475 / 495-496
    s.tu=s.tu+a*dzk;                                                    // next step's advance.f:367-369
    pk[0]=x; pk[1]=f[2];
  }
  POM_HD void mid(int comp, int i, int j, State& s) const {
    if (comp) A2(p.s2d,i,j)=s.tu; else A2(p.s2c,i,j)=s.tu;
  }
  POM_HD void sweep2(int comp, int i, int j, int k, State& s, const double* pk) const {
    const double un=pk[1]+.5*smoth*(pk[0]-s.su);                        // advance.f:483-485 / 504-506
    if (comp) A3(p.s3b,i,j,k)=un; else A3(p.s3a,i,j,k)=un;
  }
  POM_HD void fin(int comp, int i, int j, State&, const double* pl) const {
    const int kb = g.kb;
    if (comp) A3(p.s3b,i,j,kb)=pl[1]; else A3(p.s3a,i,j,kb)=pl[1];          // advance.f:511,513
  }
};

// ---------------------------------------------------------------------------
// advance.f:525-531: end-of-step 2-D rotations
struct EndStep2dK : KBase {
  POM_KINFO("endstep2d", 0, 0, 7, 7)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    egb(i,j)=egf(i,j);
    etb(i,j)=et(i,j);
    const double e=etf(i,j);
    et(i,j)=e;
    dt(i,j)=h(i,j)+e;
    utb(i,j)=utf(i,j);
    vtb(i,j)=vtf(i,j);
    vfluxb(i,j)=vfluxf(i,j);
  }
};

// realvertvl (solver.f:2024-2066) as a TMA column kernel: w (levels k and k+1), u, v staged with the halo the
// clamped index needs (one point west / south, two east / north).
#ifndef POM_RVV_TY
#define POM_RVV_TY 4
#define POM_RVV_MINB 8
#define POM_RVV_NS 4
#endif
struct RealvertvlTK : KBase {
  POM_KINFO("realvertvl", 3, 1, 7, 0)
  using KBase::KBase;
  static constexpr int TY = POM_RVV_TY, MINB = POM_RVV_MINB;
  static constexpr int NF = 3, NS = POM_RVV_NS, OHL = 1, OHR = 2, OHB = 1, OHT = 2, BW = 36, BH = TY + 3, NK = 0;
  static constexpr bool UP = true;
  static constexpr int NVEC = 1;   // (unused)
  enum { W, U, V };
  POM_HD void fields(const double** b) const { b[W] = p.w; b[U] = p.u; b[V] = p.v; }
  struct State { double m, dxr, dxl, dyt, dyb, dt0, dtE, dtW, dtN, dtS, et0, etE, etW, etN, etS, de, w0; RDiv ddti2; int di, dj; };
  POM_HD int k0() const { return 1; }
  POM_HD int k1() const { return g.kb - 1; }
  POM_HD int kl1() const { return g.kb; }
  template <class CM>
  POM_HD void pre(int i, int j, State& s, CM&) const {
    POM_DIMS;
    // edge copies S,N then W,E (:2057-2060) = value at the index clamped inside
    const int ic = i < 2 ? 2 : (i > imm1 ? imm1 : i);
    const int jc = j < 2 ? 2 : (j > jmm1 ? jmm1 : j);
    s.di = ic - i; s.dj = jc - j;
    s.m=fsm(i,j);
    s.dxr=2.0/(dx(ic+1,jc)+dx(ic,jc));
    s.dxl=2.0/(dx(ic,jc)+dx(ic-1,jc));
    s.dyt=2.0/(dy(ic,jc+1)+dy(ic,jc));
    s.dyb=2.0/(dy(ic,jc)+dy(ic,jc-1));
    s.dt0=dt(ic,jc); s.dtE=dt(ic+1,jc); s.dtW=dt(ic-1,jc); s.dtN=dt(ic,jc+1); s.dtS=dt(ic,jc-1);
    s.et0=et(ic,jc); s.etE=et(ic+1,jc); s.etW=et(ic-1,jc); s.etN=et(ic,jc+1); s.etS=et(ic,jc-1);
    s.de=etf(ic,jc)-etb(ic,jc);
    s.ddti2.set(dti2);
    wr(i,j,kb)=0.;
  }
  template <class Op, class CM>
  POM_HD void level(int i, int j, int k, State& s, CM&, const Op& o) const {
    const int di = s.di, dj = s.dj;
    if (k == 1) s.w0=o(W,di,dj);
    const double zk=zz(k);
    const double tp0=zk*s.dt0+s.et0;                                    // :2036
    const double w1=o.up(W,di,dj);
    const double r=0.5*(s.w0+w1)+0.5*
         (o(U,di+1,dj)*((zk*s.dtE+s.etE)-tp0)*s.dxr+
          o(U,di,dj)*(tp0-(zk*s.dtW+s.etW))*s.dxl+
          o(V,di,dj+1)*((zk*s.dtN+s.etN)-tp0)*s.dyt+
          o(V,di,dj)*(tp0-(zk*s.dtS+s.etS))*s.dyb)
         +s.ddti2((1.0+zk)*s.de);                                       // :2045-2050
    wr(i,j,k)=s.m*r;                                                    // :2063
    s.w0=w1;
  }
  template <class CM>
  POM_HD void post(int, int, State&, CM&) const {}
};

// in-place (fb-fclim)+fclim and fb(kb)=fb(kbm1): the side effects advt1/advt2 leave on
// their fb argument (solver.f:496,511,532 / 618,691,715); used by the unit-mode entry
struct FbRoundTripK : KBase {
  POM_KINFO("fb_roundtrip", 2, 1, 0, 0)
  double* fb_;
  const double* fc_;
  double* f_;
  FbRoundTripK(const Ctx* x, double* fb, const double* fc, double* f) : KBase(x), fb_(fb), fc_(fc), f_(f) {}
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    A3(fb_,i,j,kb)=A3(fb_,i,j,kbm1);
    if (f_) A3(f_,i,j,kb)=A3(f_,i,j,kbm1);
    for (int k = 1; k <= kb; ++k) A3(fb_,i,j,k)=(A3(fb_,i,j,k)-A3(fc_,i,j,k))+A3(fc_,i,j,k);
  }
};

// ---- domain_stats (advance.f:644-755): per-row partial sums ---------------------------------
// One block per owned row j; a fixed-shape tree reduction, so the 7 partial sums of a row do not
// depend on which strip holds it; the caller adds the rows in global order (deterministic and
// decomposition-invariant).  Per row: atot, eavg*atot, vtot, mtot, tavg*vtot, stot, ekin.
POM_HD void dstats_point(const KBase& kb_, int i, int j, double* acc) {
  const Geo& g = kb_.g; const Ptrs& p = kb_.p; const Consts& c = kb_.c; const KTab& kt = kb_.kt;
  const int im = g.im, jm = g.jmg, imm1 = im - 1, jmm1 = jm - 1, kbm1 = g.kb - 1;
  const bool jin = (j >= 2 && j <= jmm1), iin = (i >= 2 && i <= imm1);
  if (!(jin || iin)) return;                                  // the four corners are never counted
  const double darea=dx(i,j)*dy(i,j)*fsm(i,j);                // :668
  acc[0]+=darea;                                              // :669-673
  acc[1]+=et(i,j)*darea;                                      // :675-679
  if (!(jin && iin)) return;                                  // dvol is zero outside the interior (:693-696)
  for (int k = 1; k <= kbm1; ++k) {
    const double dvol=darea*dt(i,j)*dz(k);
    const double dmass=dvol*(rho(i,j,k)*rhoref+1000.);        // :704-705
    acc[2]+=dvol;
    acc[3]+=dmass;
    acc[4]+=tb(i,j,k)*dvol;                                   // :709
    acc[5]+=sb(i,j,k)*dvol;                                   // :710
    acc[6]+=.5*(dmass*(u(i,j,k)*u(i,j,k)+v(i,j,k)*v(i,j,k)));   // :742-744
  }
}
#ifndef POMGPU_EMU
__global__ void __launch_bounds__(256) dstats_kernel(const KBase kb_, int j0, double* out) {
  __shared__ double sh[7][256];
  double acc[7] = {0., 0., 0., 0., 0., 0., 0.};
  const int j = j0 + blockIdx.x;
  for (int i = 1 + threadIdx.x; i <= kb_.g.im; i += 256) dstats_point(kb_, i, j, acc);
  for (int q = 0; q < 7; ++q) sh[q][threadIdx.x] = acc[q];
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w)
      for (int q = 0; q < 7; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x < 7) out[blockIdx.x * 7 + threadIdx.x] = sh[threadIdx.x][0];
}
#endif
int domain_stats_rows(Ctx* c, double* rows) {
  const int nrow = c->jown1 - c->jown0 + 1;
  KBase kb_(c);
#ifdef POMGPU_EMU
  for (int r = 0; r < nrow; ++r) {
    // same tree as the device: 256 strided lanes, then pairwise halving
    double sh[7][256];
    for (int t = 0; t < 256; ++t) {
      double acc[7] = {0., 0., 0., 0., 0., 0., 0.};
      for (int i = 1 + t; i <= c->g.im; i += 256) dstats_point(kb_, i, c->jown0 + r, acc);
      for (int q = 0; q < 7; ++q) sh[q][t] = acc[q];
    }
    for (int w = 128; w > 0; w >>= 1)
      for (int t = 0; t < w; ++t)
        for (int q = 0; q < 7; ++q) sh[q][t] += sh[q][t + w];
    for (int q = 0; q < 7; ++q) rows[r * 7 + q] = sh[q][0];
  }
  return 0;
#else
  cudaSetDevice(c->device);
  double* d = nullptr;
  if (dev_alloc(c, &d, (size_t)nrow * 7)) return 1;
  c->launches++;
  dstats_kernel<<<nrow, 256, 0, (cudaStream_t)c->stream>>>(kb_, c->jown0, d);
  int rc = dev_d2h(c, rows, d, (size_t)nrow * 7);
  dev_free(c, d);
  return rc;
#endif
}

#define ALLI 1, c->g.im
void run_uvadjust(Ctx* c, int j0, int j1) { launch_cols(c, UvAdjustK(c), ALLI, j0, j1); }
void run_vertvl(Ctx* c, int j0, int j1) { launch_cols(c, VertvlK(c), ALLI, j0, j1); }
void run_uvadj_vertvl(Ctx* c, int j0, int j1) { launch_tma_cols(c, UvAdjVertvlTK(c), ALLI, j0, j1); }
void run_uvsum(Ctx* c, int j0, int j1) { launch_cols(c, UvSumK(c), ALLI, j0, j1); }
void run_advq(Ctx* c, int j0, int j1) { launch_tma_tiles(c, AdvqK(c), ALLI, j0, j1); }
// the fused variant under its own name / algorithmic byte count in the per-kernel profile
struct ProfqFilterK : ProfqK {
  POM_KINFO("profq_qfilter", 14, 8, 8, 0)
  ProfqFilterK(const Ctx* x, const double* hz) : ProfqK(x, hz, 1) {}
};
void run_profq(Ctx* c, int fuse_filter, int j0, int j1) {
  if (fuse_filter) launch_tma_cols(c, ProfqFilterK(c, c->hz), ALLI, j0, j1);
  else launch_tma_cols(c, ProfqK(c, c->hz, 0), ALLI, j0, j1);
}
// caller swaps q2<->uf, q2l<->vf (advance.f:418-421)
void run_qfilter(Ctx* c, int j0, int j1) { launch_cols(c, QFilterK(c), ALLI, j0, j1); }
void run_advt(Ctx* c, int nadv, const double* fb, const double* f, const double* fc, double* ff, int j0, int j1) {
  if (nadv == 1) launch_cols(c, AdvT1K(c, fb, f, fc, ff), ALLI, j0, j1);
  else launch_tma_tiles(c, AdvT2K<1>(c, fb, f, fc, ff), ALLI, j0, j1);
}
// advt2 of T (-> uf) and S (-> vf) in one pass (nitera=1)
void run_advt2_ts(Ctx* c, int j0, int j1) { launch_tma_tiles(c, AdvT2K<2>(c), ALLI, j0, j1); }
void run_advt2_up(Ctx* c, const double* fbm, const double* f, const double* xm, const double* ym, const double* zw,
                  const double* stale, double* ff, int first, int j0, int j1) {
  launch_tma_cols(c, AdvT2UpTK(c, fbm, f, xm, ym, zw, stale, ff, first), ALLI, j0, j1);
}
void run_smol_adif(Ctx* c, const double* ff, double* xm, double* ym, double* zw, int first, int j0, int j1) {
  launch_tma_cols(c, SmolAdifTK(c, ff, xm, ym, zw, first), ALLI, j0, j1);
}
void run_advt2_diff(Ctx* c, const double* fb, const double* fc, double* ff, int j0, int j1) {
  launch_tma_tiles(c, AdvT2K<1, false>(c, fb, fb, fc, ff), ALLI, j0, j1);   // the tile kernel's diffusion half
}
void run_fb_roundtrip(Ctx* c, double* fb, const double* fc, double* f, int j0, int j1) {
  launch_cols(c, FbRoundTripK(c, fb, fc, f), ALLI, j0, j1);
}
void run_proft(Ctx* c, double* f, const double* wf, const double* fs, int nbc, int j0, int j1) {
  launch_cols(c, ProftK(c, f, wf, fs, nbc), ALLI, j0, j1);
}
void run_proft_ts(Ctx* c, int fuse, int j0, int j1) {
  const bool same = ((c->c.nbct == 1 || c->c.nbct == 2) == (c->c.nbcs == 1 || c->c.nbcs == 2));
  if (fuse) { if (same) launch_tma_cols(c, ProftFilterK<true>(c), ALLI, j0, j1); else launch_tma_cols(c, ProftFilterK<false>(c), ALLI, j0, j1); }
  else { if (same) launch_tma_cols(c, ProftTSK<true>(c), ALLI, j0, j1); else launch_tma_cols(c, ProftTSK<false>(c), ALLI, j0, j1); }
}
// caller swaps t<->uf, s<->vf (advance.f:446-449)
void run_tsfilter(Ctx* c, int with_dens, int j0, int j1) { launch_cols(c, TsFilterK(c, with_dens), ALLI, j0, j1); }
void run_dens(Ctx* c, const double* si, const double* ti, double* ro, int j0, int j1) {
  launch_cols(c, DensK(c, si, ti, ro), ALLI, j0, j1);
}
void run_advu(Ctx* c, int j0, int j1) { launch_cols(c, AdvuK(c), ALLI, j0, j1); }
void run_advv(Ctx* c, int j0, int j1) { launch_cols(c, AdvvK(c), ALLI, j0, j1); }
void run_advprof_u(Ctx* c, int j0, int j1) { launch_tma_cols(c, AdvProfUVK<false>(c), ALLI, j0, j1); }
void run_advprof_v(Ctx* c, int j0, int j1) { launch_tma_cols(c, AdvProfUVK<true>(c), ALLI, j0, j1); }
void run_profu(Ctx* c, int j0, int j1) { launch_cols(c, ProfuK(c), ALLI, j0, j1); }
void run_profv(Ctx* c, int j0, int j1) { launch_cols(c, ProfvK(c), ALLI, j0, j1); }
// caller swaps u<->uf, v<->vf, ub<->s3a, vb<->s3b (advance.f:511-514)
void run_uvfilter(Ctx* c, int j0, int j1) {
  // interior rectangle streamed once through the parked kernel, the frame around it (Orlanski rows / columns
  // and their neighbours) on its side warps; layouts the TMA cannot address keep the two-sweep kernel
  const int ja0 = j0 > 3 ? j0 : 3, ja1 = j1 < c->g.jmg - 1 ? j1 : c->g.jmg - 1;
#ifndef POM_UVF_TWOSWEEP
  if (c->g.im >= 4 && ja1 >= ja0 && c->g.kb - 1 <= KMAX &&
      launch_tma_park(c, UvFilterParkK(c, j0, j1, ja0, ja1), 3, c->g.im - 1, ja0, ja1)) return;
#endif
  launch_cols(c, UvFilterK(c), ALLI, j0, j1);
}
void run_endstep2d(Ctx* c, int j0, int j1) { launch_cols(c, EndStep2dK(c), ALLI, j0, j1); }
void run_realvertvl(Ctx* c, int j0, int j1) { launch_tma_cols(c, RealvertvlTK(c), ALLI, j0, j1); }

}  // namespace pom
