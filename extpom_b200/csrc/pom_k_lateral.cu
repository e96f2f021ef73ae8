// pom_k_lateral.cu -- lateral_viscosity group (advance.f:96-141):
//   advct  (solver.f:201-409)  + vertical integrals adx2d/ady2d (advance.f:161-162)
//   baropg (solver.f:848-940)  + vertical integrals drx2d/dry2d (advance.f:163-164)
//   Smagorinsky aam (advance.f:122-136) + aam2d (advance.f:165)
// One thread per (i,j) column marching k; flux intermediates of the reference
// (xflux,yflux,curv automatic arrays) are recomputed in registers instead of
// being staged through HBM.  Expression order follows the Fortran so that the
// results are bit-identical to a no-FMA evaluation.
#include "pom_core.h"
#include "pom_tma.h"
#include "pom_names.h"

#ifndef POM_ADVCT_TY
#define POM_ADVCT_TY 16
#define POM_ADVCT_MINB 1
#endif

namespace pom {

// advct as a shared-memory tile kernel: per level every thread evaluates the five fluxes
// of its own point once -- x-part xflux (X), x-part yflux (Y), y-part xflux (XP), y-part
// yflux (YP), and the curvature products (CV, CU) -- and the tile interior differences them.
struct AdvctK : KBase {
  POM_KINFO("advct", 5, 2, 5, 2)
  using KBase::KBase;
#ifndef POM_TILE_TY
#define POM_TILE_TY 16
#define POM_TILE_MINB 1
#define POM_TILE_NS 4
#endif
  static constexpr int NV = 6, HL = 1, HR = 1, HB = 1, HT = 1, TY = POM_TILE_TY, MINB = POM_TILE_MINB;
  // operands staged by the TMA: thread tile + one point all around (36 x 18 box)
#ifndef POM_NS_ADVCT
#define POM_NS_ADVCT POM_TILE_NS
#endif
  static constexpr int NF = 5, NS = POM_NS_ADVCT, OHL = 1, OHR = 1, OHB = 1, OHT = 1, BW = 36, BH = TY + 2, NK = 0;
  static constexpr bool UP = false;
  static constexpr bool FULL = true;   // stage() may run for every thread and assigns every v[]
  enum { U, V, UB, VB, AAM };
  enum { X, Y, XP, YP, CV, CU };
  POM_HD void fields(const double** b) const { b[U] = p.u; b[V] = p.v; b[UB] = p.ub; b[VB] = p.vb; b[AAM] = p.aam; }
  POM_HD int kl1() const { return g.kb - 1; }
  struct State {
    double dtE, dtW, dtS, dtN, dtSW, dtWS, q4, dtc, dxc, dyc, qdx4, qdy4, dyd, dxd;
    RDiv ddx, ddy, ddy4, ddx4, ddxdy;   // hoisted divisors dx, dy, dy4, dx4, dx*dy
    double aru25, arv25, sx, sy;
    bool fx, fy, fxp, fyp, interior, i3, j3;
  };
  POM_HD int k0() const { return 1; }
  POM_HD int k1() const { return g.kb - 1; }
  POM_HD void pre(int i, int j, bool inside, bool out, State& s) const {
    POM_DIMS;
    const int jlo = g.joff + 1, jhi = g.joff + g.jml;
    s.sx = 0.; s.sy = 0.;
    // definition ranges (zero fill elsewhere, solver.f:213-216,319-321), limited to rows in memory
    s.fx  = inside && i >= 2 && i <= imm1 && j >= 2 && j <= jmm1;                 // :234-277
    s.fy  = inside && i >= 2 && i <= imm1 && j >= 2 && j <= jm && j - 1 >= jlo;   // :244-276
    s.fxp = inside && i >= 2 && i <= im && j >= 2 && j <= jmm1 && j - 1 >= jlo;   // :324-366
    s.fyp = inside && i >= 2 && i <= imm1 && j >= 2 && j <= jmm1 && j - 1 >= jlo && j + 1 <= jhi;   // :334-366, curv :218-228
    s.interior = out && i >= 2 && i <= imm1 && j >= 2 && j <= jmm1;
    s.i3 = i >= 3; s.j3 = j >= 3;
    if (!(s.fx || s.fy || s.fxp || s.fyp)) return;
    s.dtc = dt(i,j); s.dxc = dx(i,j); s.dyc = dy(i,j);
    s.ddx.set(s.dxc); s.ddy.set(s.dyc);
    s.dtW = dt(i,j)+dt(i-1,j);
    if (s.fx || s.fyp) s.dtE = dt(i+1,j)+dt(i,j);
    if (s.fy || s.fxp) {
      s.dtS = dt(i,j)+dt(i,j-1);
      s.dtSW = dt(i-1,j)+dt(i-1,j-1);
      s.dtWS = dt(i,j-1)+dt(i-1,j-1);
      s.q4 = .25*(dt(i,j)+dt(i-1,j)+dt(i,j-1)+dt(i-1,j-1));
      const double dy4 = dy(i,j)+dy(i-1,j)+dy(i,j-1)+dy(i-1,j-1);
      const double dx4 = dx(i,j)+dx(i-1,j)+dx(i,j-1)+dx(i-1,j-1);
      s.ddy4.set(dy4); s.ddx4.set(dx4);
      s.qdx4 = .25*dx4; s.qdy4 = .25*dy4;
    }
    if (s.fyp) {
      s.dtN = dt(i,j+1)+dt(i,j);
      s.dyd = dy(i+1,j)-dy(i-1,j);
      s.dxd = dx(i,j+1)-dx(i,j-1);
      s.ddxdy.set(dx(i,j)*dy(i,j));
    }
    if (s.interior) { s.aru25 = aru(i,j)*.25; s.arv25 = arv(i,j)*.25; }
  }
  // Branch-free: every thread evaluates every flux (operands outside the arrays read as zero, the
  // hoisted metrics of a thread outside a definition range are zero) and the definition-range
  // flags select the result; the arithmetic of a selected value is unchanged.
  template <class Op>
  POM_HD void stage(int i, int j, int k, State& s, const Op& o, double* v) const {
    const double u00 = o(U,0,0), v00 = o(V,0,0), ub00 = o(UB,0,0), vb00 = o(VB,0,0), a00 = o(AAM,0,0);
    const double uE = o(U,1,0), vN = o(V,0,1);
    {                                                                    // :237-239,258-260,272
      double a=.125*(s.dtE*uE+s.dtW*u00)*(uE+u00);
      a=a-s.ddx(s.dtc*a00*2.*(o(UB,1,0)-ub00));
      v[X]=s.fx ? s.dyc*a : 0.;
    }
    {
      const double uS = o(U,0,-1), vW = o(V,-1,0);
      const double dtaam=s.q4*(a00+o(AAM,-1,0)+o(AAM,0,-1)+o(AAM,-1,-1));  // :261-263,348-350
      const double cd=dtaam*(s.ddy4(ub00-o(UB,0,-1))+s.ddx4(vb00-o(VB,-1,0)));   // :265-270,352-357
      const double ay=.125*(s.dtS*v00+s.dtSW*vW)*(u00+uS);                  // :247-249,273-274
      v[Y]=s.fy ? s.qdx4*(ay-cd) : 0.;
      const double ax=.125*(s.dtW*u00+s.dtWS*uS)*(v00+vW);                  // :327-329,362-363
      v[XP]=s.fxp ? s.qdy4*(ax-cd) : 0.;
    }
    {
      double a=.125*(s.dtN*vN+s.dtS*v00)*(vN+v00);                        // :337-339
      a=a-s.ddy(s.dtc*a00*2.*(o(VB,0,1)-vb00));                            // :358-360
      v[YP]=s.fyp ? s.dxc*a : 0.;                                         // :364
      const double cv=s.ddxdy(.25*((vN+v00)*s.dyd-(uE+u00)*s.dxd));         // :221-225
      v[CV]=s.fyp ? cv*s.dtc*(vN+v00) : 0.;                               // :297-298
      v[CU]=s.fyp ? cv*s.dtc*(uE+u00) : 0.;                               // :387-388
    }
  }
  template <class Op>
  POM_HD void combine(int i, int j, int k, State& s, const Op&, const Tile2& tl) const {
    double ax=tl(X,0,0)-tl(X,-1,0)+tl(Y,0,1)-tl(Y,0,0);                      // :285-286
    const double cx=s.aru25*(tl(CV,0,0)+tl(CV,-1,0));                     // :293-300 (n_west==-1)
    ax=s.i3 ? ax-cx : ax;
    double ay=tl(XP,1,0)-tl(XP,0,0)+tl(YP,0,0)-tl(YP,0,-1);                  // :375-376
    const double cy=s.arv25*(tl(CU,0,0)+tl(CU,0,-1));                     // :383-390 (n_south==-1)
    ay=s.j3 ? ay+cy : ay;
    ax=s.interior ? ax : 0.;
    ay=s.interior ? ay : 0.;
    advx(i,j,k)=ax;
    advy(i,j,k)=ay;
    s.sx=s.sx+ax*dz(k);                                                   // advance.f:161-162
    s.sy=s.sy+ay*dz(k);
  }
  POM_HD void post(int i, int j, State& s) const {
    const int kb = g.kb;
    advx(i,j,kb)=0.;
    advy(i,j,kb)=0.;
    adx2d(i,j)=s.sx;
    ady2d(i,j)=s.sy;
  }
};

// solver.f:848-940.  rho is rewritten as (rho-rmean)+rmean (:854,937) into the
// alternate buffer rho2 (neighbours still read the old rho); the caller swaps.
// A TMA column kernel (pom_tma.h: tmacolkernel): rho and rmean of every level are staged with their west and
// south neighbours, two or more levels ahead (the plain-load column kernel of round 1 took 0.42 against 0.33 ms).
#ifndef POM_BAROPG_TY
#define POM_BAROPG_TY 4
#define POM_BAROPG_MINB 8
#define POM_BAROPG_NS 4
#endif
struct BaropgTK : KBase {
  POM_KINFO("baropg", 2, 3, 5, 2)
  using KBase::KBase;
  static constexpr int TY = POM_BAROPG_TY, MINB = POM_BAROPG_MINB;
  static constexpr int NF = 2, NS = POM_BAROPG_NS, OHL = 1, OHR = 0, OHB = 1, OHT = 0, BW = 34, BH = TY + 1, NK = 0;
  static constexpr bool UP = false;
  static constexpr int NVEC = 1;   // (unused)
  enum { RHO, RMEAN };
  POM_HD void fields(const double** b) const { b[RHO] = p.rho; b[RMEAN] = p.rmean; }
  struct State { double dtx, dty, ddx, ddy, dyx, dxy, mu, mv, a0, ax, ay, px, py, sx, sy; bool interior; };
  POM_HD int k0() const { return 1; }
  POM_HD int k1() const { return g.kb; }
  POM_HD int kl1() const { return g.kb; }
  template <class CM>
  POM_HD void pre(int i, int j, State& s, CM&) const {
    POM_DIMS;
    s.interior = (i >= 2 && i <= imm1 && j >= 2 && j <= jmm1);
    s.sx = 0.; s.sy = 0.;
    if (!s.interior) return;
    s.dtx=dt(i,j)+dt(i-1,j); s.dty=dt(i,j)+dt(i,j-1);
    s.ddx=dt(i,j)-dt(i-1,j); s.ddy=dt(i,j)-dt(i,j-1);
    s.dyx=dy(i,j)+dy(i-1,j); s.dxy=dx(i,j)+dx(i,j-1);
    s.mu=dum(i,j); s.mv=dvm(i,j);
  }
  template <class Op, class CM>
  POM_HD void level(int i, int j, int k, State& s, CM&, const Op& o) const {
    POM_DIMS;
    const double rmk=o(RMEAN,0,0);
    const double b0=o(RHO,0,0)-rmk;
    rho2(i,j,k)=b0+rmk;                                                 // :854,937 folded into the sweep
    if (!s.interior) {
      if (k <= kbm1) {   // edges keep their content (initialize.f:307-308)
        s.sx=s.sx+drhox(i,j,k)*dz(k);
        s.sy=s.sy+drhoy(i,j,k)*dz(k);
      }
      return;
    }
    if (k == kb) {                                                      // :928-932 (k=kb)
      drhox(i,j,kb)=ramp*drhox(i,j,kb);
      drhoy(i,j,kb)=ramp*drhoy(i,j,kb);
      return;
    }
    const double bx=o(RHO,-1,0)-o(RMEAN,-1,0);
    const double by=o(RHO,0,-1)-o(RMEAN,0,-1);
    if (k == 1) {
      s.px=.5*grav*(-zz(1))*s.dtx*(b0-bx);                              // :859-860
      s.py=.5*grav*(-zz(1))*s.dty*(b0-by);                              // :895-896
    } else {
      s.px=s.px+grav*.25*(zz(k-1)-zz(k))*s.dtx*(b0-bx+s.a0-s.ax)
               +grav*.25*(zz(k-1)+zz(k))*s.ddx*(b0+bx-s.a0-s.ax);       // :867-875
      s.py=s.py+grav*.25*(zz(k-1)-zz(k))*s.dty*(b0-by+s.a0-s.ay)
               +grav*.25*(zz(k-1)+zz(k))*s.ddy*(b0+by-s.a0-s.ay);       // :903-911
    }
    s.a0=b0; s.ax=bx; s.ay=by;
    const double ox=ramp*(.25*s.dtx*s.px*s.mu*s.dyx);                   // :883-885,931
    const double oy=ramp*(.25*s.dty*s.py*s.mv*s.dxy);                   // :919-921,932
    drhox(i,j,k)=ox;
    drhoy(i,j,k)=oy;
    s.sx=s.sx+ox*dz(k);                                                 // advance.f:163-164
    s.sy=s.sy+oy*dz(k);
  }
  template <class CM>
  POM_HD void post(int i, int j, State& s, CM&) const { drx2d(i,j)=s.sx; dry2d(i,j)=s.sy; }
};

// baropg_mcc (solver.f:943-1159, npg=2): 4th-order McCalpin pressure gradient.  Same outputs and
// side effects as BaropgK; reads rho-rmean and d two cells away (the order2d/3d_mpi width-2 halo
// of the reference = ghost rows here).  Single-precision literals (1./24.), (1./16.) kept.
struct BaropgMccK : KBase {
  POM_KINFO("baropg_mcc", 2, 3, 6, 2)
  using KBase::KBase;
  POM_HD double rr(int i, int j, int k) const { return rho(i,j,k)-rmean(i,j,k); }   // :954
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    const double c24 = (double)(1.f / 24.f), c16 = (double)(1.f / 16.f);
    const bool interior = (i >= 2 && i <= imm1 && j >= 2 && j <= jmm1);
    double sx = 0., sy = 0.;
    if (interior) {
      const bool cx = (i >= 3), cy = (j >= 3);            // n_west == -1 / n_south == -1 ranges (:982,:1080)
      const double mu=dum(i,j), mv=dvm(i,j);
      const double muE=dum(i+1,j), muW=dum(i-1,j), mvN=dvm(i,j+1), mvS=dvm(i,j-1);
      double ddxx=(d(i,j)-d(i-1,j))*mu, d4x=.5*(d(i,j)+d(i-1,j))*mu;                  // :976-977
      double ddxy=(d(i,j)-d(i,j-1))*mv, d4y=.5*(d(i,j)+d(i,j-1))*mv;                  // :1074-1075
      if (cx) {                                                                       // :994-1001
        ddxx=ddxx-c24*(muE*(d(i+1,j)-d(i,j))-2*(d(i,j)-d(i-1,j))+muW*(d(i-1,j)-d(i-2,j)));
        d4x=d4x+c16*(muE*(d(i,j)-d(i+1,j))+muW*(d(i-1,j)-d(i-2,j)));
      }
      if (cy) {                                                                       // :1092-1099
        ddxy=ddxy-c24*(mvN*(d(i,j+1)-d(i,j))-2*(d(i,j)-d(i,j-1))+mvS*(d(i,j-1)-d(i,j-2)));
        d4y=d4y+c16*(mvN*(d(i,j)-d(i,j+1))+mvS*(d(i,j-1)-d(i,j-2)));
      }
      const double dtx=dt(i,j)+dt(i-1,j), dty=dt(i,j)+dt(i,j-1);
      const double dyx=dy(i,j)+dy(i-1,j), dxy=dx(i,j)+dx(i,j-1);
      double px = 0., py = 0., drxm = 0., rhxm = 0., drym = 0., rhym = 0.;
      for (int k = 1; k <= kbm1; ++k) {
        const double r0=rr(i,j,k), rW=rr(i-1,j,k), rS=rr(i,j-1,k);
        double drx=(r0-rW)*mu, rhx=0.5*(r0+rW)*mu;                                     // :972-973
        double dry=(r0-rS)*mv, rhy=.5*(r0+rS)*mv;                                      // :1070-1071
        if (cx) {                                                                     // :985-992
          const double rE=rr(i+1,j,k), rWW=rr(i-2,j,k);
          drx=drx-c24*(muE*(rE-r0)-2*(r0-rW)+muW*(rW-rWW));
          rhx=rhx+c16*(muE*(r0-rE)+muW*(rW-rWW));
        }
        if (cy) {                                                                     // :1083-1090
          const double rN=rr(i,j+1,k), rSS=rr(i,j-2,k);
          dry=dry-c24*(mvN*(rN-r0)-2*(r0-rS)+mvS*(rS-rSS));
          rhy=rhy+c16*(mvN*(r0-rN)+mvS*(rS-rSS));
        }
        if (k == 1) {
          px=grav*(-zz(1))*d4x*drx;                                                   // :1031
          py=grav*(-zz(1))*d4y*dry;                                                   // :1129
        } else {
          px=px+grav*0.5*dzz(k-1)*d4x*(drxm+drx)+grav*0.5*(zz(k-1)+zz(k))*ddxx*(rhx-rhxm);   // :1038-1042
          py=py+grav*0.5*dzz(k-1)*d4y*(drym+dry)+grav*0.5*(zz(k-1)+zz(k))*ddxy*(rhy-rhym);   // :1136-1140
        }
        drxm=drx; rhxm=rhx; drym=dry; rhym=rhy;
        const double ox=ramp*(.25*dtx*px*mu*dyx);                                     // :1050-1052,1160
        const double oy=ramp*(.25*dty*py*mv*dxy);                                     // :1148-1150,1161
        drhox(i,j,k)=ox;
        drhoy(i,j,k)=oy;
        sx=sx+ox*dz(k);                                                               // advance.f:163-164
        sy=sy+oy*dz(k);
      }
      drhox(i,j,kb)=ramp*drhox(i,j,kb);
      drhoy(i,j,kb)=ramp*drhoy(i,j,kb);
    } else {
      for (int k = 1; k <= kbm1; ++k) {  // edges keep their content (initialize.f:307-308)
        sx=sx+drhox(i,j,k)*dz(k);
        sy=sy+drhoy(i,j,k)*dz(k);
      }
    }
    drx2d(i,j)=sx;
    dry2d(i,j)=sy;
    for (int k = 1; k <= kb; ++k)
      rho2(i,j,k)=(rho(i,j,k)-rmean(i,j,k))+rmean(i,j,k);                             // :954,1166
  }
};

// advance.f:122-136 + aam2d (advance.f:165)
// A TMA column kernel (pom_tma.h: tmacolkernel): u, v of every level are staged with one point of halo all
// around, two or more levels ahead, instead of ten dependent plain loads per cell (0.31 -> 0.22 ms).
#ifndef POM_SMAG_TY
#define POM_SMAG_TY 4
#endif
#ifndef POM_SMAG_NS
#define POM_SMAG_NS 4
#endif
#ifndef POM_SMAG_MINB
#define POM_SMAG_MINB 8
#endif
struct SmagTK : KBase {
  POM_KINFO("smagorinsky", 2, 1, 2, 1)
  using KBase::KBase;
  static constexpr int TY = POM_SMAG_TY, MINB = POM_SMAG_MINB;
  static constexpr int NF = 2, NS = POM_SMAG_NS, OHL = 1, OHR = 1, OHB = 1, OHT = 1, BW = 36, BH = TY + 2, NK = 0;
  static constexpr bool UP = false;
  static constexpr int NVEC = 1;   // (unused)
  enum { U, V };
  POM_HD void fields(const double** b) const { b[U] = p.u; b[V] = p.v; }
  struct State { RDiv ddx, ddy; double hdd, sa; bool interior; };
  POM_HD int k0() const { return 1; }
  POM_HD int k1() const { return g.kb - 1; }
  POM_HD int kl1() const { return g.kb - 1; }
  template <class CM>
  POM_HD void pre(int i, int j, State& s, CM&) const {
    POM_DIMS;
    s.interior = (i >= 2 && i <= imm1 && j >= 2 && j <= jmm1);
    s.sa = 0.; s.hdd = 0.;
    if (s.interior) { s.ddx.set(dx(i,j)); s.ddy.set(dy(i,j)); s.hdd=horcon*dx(i,j)*dy(i,j); }
  }
  template <class Op, class CM>
  POM_HD void level(int i, int j, int k, State& s, CM&, const Op& o) const {
    double a;
    if (s.interior) {
      const double u00=o(U,0,0), v00=o(V,0,0);
      double a1=s.ddx(o(U,1,0)-u00);
      double a2=s.ddy(o(V,0,1)-v00);
      double a3=s.ddy(.25*(o(U,0,1)+o(U,1,1)-o(U,0,-1)-o(U,1,-1)))
               +s.ddx(.25*(o(V,1,0)+o(V,1,1)-o(V,-1,0)-o(V,-1,1)));
      a=s.hdd*sqrt(a1*a1+a2*a2+.5*(a3*a3));
      aam(i,j,k)=a;
    } else {
      a=aam(i,j,k);
    }
    s.sa=s.sa+a*dz(k);
  }
  template <class CM>
  POM_HD void post(int i, int j, State& s, CM&) const { aam2d(i,j)=s.sa; }
};

void run_advct(Ctx* c, int j0, int j1) { launch_tma_tiles(c, AdvctK(c), 1, c->g.im, j0, j1); }
// the caller swaps rho <-> rho2 afterwards
void run_baropg(Ctx* c, int j0, int j1) { launch_tma_cols(c, BaropgTK(c), 1, c->g.im, j0, j1); }
void run_baropg_mcc(Ctx* c, int j0, int j1) { launch_cols(c, BaropgMccK(c), 1, c->g.im, j0, j1); }
void run_smag(Ctx* c, int j0, int j1) { launch_tma_cols(c, SmagTK(c), 1, c->g.im, j0, j1); }

}  // namespace pom
