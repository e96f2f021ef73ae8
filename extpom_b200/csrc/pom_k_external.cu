// pom_k_external.cu -- the 2-D external mode (advance.f:144-353):
//   advave (solver.f:6-121), mode_interaction tail (advance.f:172-196),
//   mode_external (advance.f:211-350) with bcond(1), bcond(2)
//   (bounds_forcing.f:18-83) folded into the producing kernels.
// ONE kernel per substep (ExtStepK).  The reference's Asselin filter + time rotation by
// whole-array copies (advance.f:321-330) becomes a write of the filtered field into a dead
// buffer followed by a pointer rotation on the host.
#include "pom_core.h"
#include "pom_tma.h"
#include "pom_names.h"

namespace pom {

// ------------------------------------------------------------------ advave ----
// advave (solver.f:16-121): every thread evaluates the four fluxes of its own point -- u-half
// fluxua / fluxva, v-half fluxua / fluxva -- once from the staged operands, sharing tps
// (solver.f:47-53,106) between the two halves like the reference does.  Writes v[V0..V0+3].
// Operand ids D,UA,VA,UAB,VAB,AAM2D,DX,DY = 0..7 in both users (AdvaveK, ExtStepK).
enum { OP_D, OP_UA, OP_VA, OP_UAB, OP_VAB, OP_AAM2D, OP_DX, OP_DY };
template <int V0, class Op>
POM_HD void advave_own_fluxes(const Geo& g, int i, int j, const Op& o, double* v) {
  enum { FXU = V0, FYU, FXV, FYV };
  const int imm1 = g.im - 1, jmm1 = g.jmg - 1;
  const int jlo = g.joff + 1, jhi = g.joff + g.jml;
  if (i < 2 || j < 2 || j - 1 < jlo) return;
  const double d00=o(OP_D,0,0), dW=o(OP_D,-1,0), dS=o(OP_D,0,-1), dSW=o(OP_D,-1,-1);
  const double ua00=o(OP_UA,0,0), va00=o(OP_VA,0,0), uaS=o(OP_UA,0,-1), vaW=o(OP_VA,-1,0);
  const double dx00=o(OP_DX,0,0), dy00=o(OP_DY,0,0);
  const double dx4=dx00+o(OP_DX,-1,0)+o(OP_DX,0,-1)+o(OP_DX,-1,-1);
  const double dy4=dy00+o(OP_DY,-1,0)+o(OP_DY,0,-1)+o(OP_DY,-1,-1);
  const double am00=o(OP_AAM2D,0,0), uab00=o(OP_UAB,0,0), vab00=o(OP_VAB,0,0);
  // tps(i,j), 2<=i<=im, 2<=j<=jm (:47-53)
  const double tp=.25*(d00+dW+dS+dSW)
                  *(am00+o(OP_AAM2D,0,-1)+o(OP_AAM2D,-1,0)+o(OP_AAM2D,-1,-1))
                  *(pdiv(uab00-o(OP_UAB,0,-1),dy4)+pdiv(vab00-o(OP_VAB,-1,0),dx4));
  {   // u half fluxva (:30-32,55-56) and v half fluxua (:80-82,106-107)
    double a=.125*((d00+dS)*va00+(dW+dSW)*vaW)*(ua00+uaS);
    v[FYU]=(a-tp)*.25*dx4;
    double b=.125*((d00+dW)*ua00+(dS+dSW)*uaS)*(vaW+va00);
    v[FXV]=(b-tp)*.25*dy4;
  }
  if (i <= imm1) {   // u half fluxua (:22-24,39-41,54)
    const double dE=o(OP_D,1,0), uaE=o(OP_UA,1,0);
    double a=.125*((dE+d00)*uaE+(d00+dW)*ua00)*(uaE+ua00);
    a=a-pdiv(d00*2.*am00*(o(OP_UAB,1,0)-uab00),dx00);
    v[FXU]=a*dy00;
  }
  if (j <= jmm1 && j + 1 <= jhi) {   // v half fluxva (:88-90,97-99,105)
    const double dN=o(OP_D,0,1), vaN=o(OP_VA,0,1);
    double a=.125*((dN+d00)*vaN+(d00+dS)*va00)*(vaN+va00);
    a=a-pdiv(d00*2.*am00*(o(OP_VAB,0,1)-vab00),dy00);
    v[FYV]=a*dx00;
  }
}

// stand-alone advave (the call in mode_interaction, advance.f:170, and the C-ABI entry)
struct AdvaveK : KBase {
  POM_KINFO("advave", 0, 0, 8, 2)
  using KBase::KBase;
  static constexpr int NV = 4, HL = 1, HR = 1, HB = 1, HT = 1, TY = 16, MINB = 1;
  static constexpr int NF = 8, NS = 1, OHL = 1, OHR = 1, OHB = 1, OHT = 1, BW = 36, BH = 18, NK = 1;
  static constexpr bool UP = false;
  static constexpr bool FULL = false;   // stage() may run for every thread and assigns every v[]
  enum { FXU, FYU, FXV, FYV };
  POM_HD void fields(const double** b) const {
    b[OP_D] = p.d; b[OP_UA] = p.ua; b[OP_VA] = p.va; b[OP_UAB] = p.uab; b[OP_VAB] = p.vab; b[OP_AAM2D] = p.aam2d;
    b[OP_DX] = p.dx; b[OP_DY] = p.dy;
  }
  struct State { bool interior; };
  POM_HD int k0() const { return 1; }
  POM_HD int k1() const { return 1; }
  POM_HD int kl1() const { return 1; }
  POM_HD void pre(int i, int j, bool, bool out, State& s) const {
    POM_DIMS;
    s.interior = out && i >= 2 && i <= imm1 && j >= 2 && j <= jmm1;
  }
  template <class Op>
  POM_HD void stage(int i, int j, int, State&, const Op& o, double* v) const { advave_own_fluxes<0>(g, i, j, o, v); }
  template <class Op>
  POM_HD void combine(int i, int j, int, State& s, const Op&, const Tile2& tl) const {
    double au = 0., av = 0.;
    if (s.interior) {
      au=tl(FXU,0,0)-tl(FXU,-1,0)+tl(FYU,0,1)-tl(FYU,0,0);       // :65-66
      av=tl(FXV,1,0)-tl(FXV,0,0)+tl(FYV,0,0)-tl(FYV,0,-1);       // :116-117
    }
    advua(i,j)=au;
    advva(i,j)=av;
  }
  POM_HD void post(int, int, State&) const {}
};

// ------------------------------------------- mode_interaction tail -------------
// advance.f:172-196: adx2d-=advua, ady2d-=advva, egf, utf, vtf
struct ModeInterTailK : KBase {
  POM_KINFO("mode_interaction", 0, 0, 8, 5)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    if (c.mode != 2) {
      adx2d(i,j)=adx2d(i,j)-advua(i,j);
      ady2d(i,j)=ady2d(i,j)-advva(i,j);
    }
    egf(i,j)=el(i,j)*ispi;
    if (i >= 2) utf(i,j)=ua(i,j)*(d(i,j)+d(i-1,j))*isp2i;
    if (j >= 2) vtf(i,j)=va(i,j)*(d(i,j)+d(i,j-1))*isp2i;
  }
};

// ------------------------------------------------------------------------------------------
// One external substep in ONE kernel (advance.f:211-350): elf + bcond(1), advave, uaf/vaf +
// bcond(2), etf, Asselin filter, running means.  Three phases on a TMA-staged tile with two
// neighbour exchanges through shared memory (pom_tma.h: tile3kernel):
//   A  own-point transports fua,fva (advance.f:211-218) and the four advave fluxes
//   B  elf(i,j) with bcond(1) from the neighbours' transports; advua, advva (kept in registers)
//   C  uaf,vaf (+bcond(2)) from elf(i,j), elf(i-1,j), elf(i,j-1); everything point-wise after
// 13 fields are read with neighbours (staged, halo 1), 18 at the own point only.  The filtered
// uab,vab go to the scratch buffers s2a,s2b (advave reads uab,vab of the neighbours); the
// caller rotates ua<->uaf, va<->vaf, uab<->s2a, vab<->s2b, elb<->el2, el<->elf, d<->d2.
template <bool M2>   // M2: mode=2 build with advave's bottom-stress / curvature block (solver.f:123-195)
struct ExtStepK : KBase {
  POM_KINFO("ext_step", 0, 0, 31, 12)
  int iext, do_adv;
  ExtStepK(const Ctx* x, int ie, int adv) : KBase(x), iext(ie), do_adv(adv) {}
#ifndef POM_EXT_TY
#define POM_EXT_TY 12
#define POM_EXT_MINB 2
#endif
  static constexpr int NV = M2 ? 7 : 6, TY = POM_EXT_TY, MINB = POM_EXT_MINB;
  static constexpr int NF = 13, OHL = 1, OHR = 1, OHB = 1, OHT = 1, BW = 36, BH = POM_EXT_TY + 2, NK = 1;
  enum { D = OP_D, UA = OP_UA, VA = OP_VA, UAB = OP_UAB, VAB = OP_VAB, AAM2D = OP_AAM2D, DX = OP_DX, DY = OP_DY,
         ELB, EL, H, COR, EATM };
  static constexpr int NFA = 9;   // D..DY, ELB: everything phases A and B read
  enum { FUA, FVA, FXU, FYU, FXV, FYV, CURV };   // CURV: curv2d of the mode=2 block (solver.f:145-152)
  POM_HD void fields(const double** b) const {
    b[D] = p.d; b[UA] = p.ua; b[VA] = p.va; b[UAB] = p.uab; b[VAB] = p.vab; b[AAM2D] = p.aam2d; b[DX] = p.dx;
    b[DY] = p.dy; b[EL] = p.el; b[ELB] = p.elb; b[H] = p.h; b[COR] = p.cor; b[EATM] = p.e_atmos;
  }
  struct State { double au, av, artc, vflc, fsm0, wub, wvb; bool m2; };
  POM_HD void pre(int i, int j, bool inside, State& s) const {
    s.au = 0.; s.av = 0.; s.artc = 1.; s.vflc = 0.; s.fsm0 = 0.; s.wub = 0.; s.wvb = 0.; s.m2 = false;
    if (inside) {   // phase B's point-wise operands, in flight while the TMA stages the rest
      const int imm1 = g.im - 1, jmm1 = g.jmg - 1;
      const int ic = i < 2 ? 2 : (i > imm1 ? imm1 : i);
      const int jc = j < 2 ? 2 : (j > jmm1 ? jmm1 : j);
      s.artc = POM_LDG(&art(ic,jc)); s.vflc = POM_LDG(&vfluxf(ic,jc)); s.fsm0 = POM_LDG(&fsm(i,j));
    }
    // the 18 point-wise operands of phase C: start them towards L2 now, while the TMA stages
    // the stencil operands (one thread per 128-byte line)
    if (inside && ((i - 1) & 15) == 0) {
      const int o = POM_I2(i,j);
      POM_PREFETCH_L2(p.aru+o);
      POM_PREFETCH_L2(p.arv+o); POM_PREFETCH_L2(p.adx2d+o); POM_PREFETCH_L2(p.ady2d+o); POM_PREFETCH_L2(p.drx2d+o);
      POM_PREFETCH_L2(p.dry2d+o); POM_PREFETCH_L2(p.wusurf+o); POM_PREFETCH_L2(p.wubot+o); POM_PREFETCH_L2(p.wvsurf+o);
      POM_PREFETCH_L2(p.wvbot+o); POM_PREFETCH_L2(p.dum+o); POM_PREFETCH_L2(p.dvm+o); POM_PREFETCH_L2(p.egf+o);
      POM_PREFETCH_L2(p.utf+o); POM_PREFETCH_L2(p.vtf+o);
      if (iext >= c.isplit - 1) POM_PREFETCH_L2(p.etf+o);
    }
  }
  template <class Op>
  POM_HD void phaseA(int i, int j, State&, const Op& o, double* v) const {
    POM_DIMS;
    const int jlo = g.joff + 1;
    const double d00=o(D,0,0);
    if (i >= 2) v[FUA]=.25*(d00+o(D,-1,0))*(o(DY,0,0)+o(DY,-1,0))*o(UA,0,0);          // advance.f:213-214
    if (j >= 2 && j - 1 >= jlo) v[FVA]=.25*(d00+o(D,0,-1))*(o(DX,0,0)+o(DX,0,-1))*o(VA,0,0);   // :215-216
    if (do_adv) advave_own_fluxes<FXU>(g, i, j, o, v);       // advave (solver.f:16-121)
    if (M2 && do_adv && i >= 2 && i <= imm1 && j >= 2 && j <= jmm1 && j + 1 <= g.joff + g.jml)
      v[CURV]=.25*((o(VA,0,1)+o(VA,0,0))*(o(DY,1,0)-o(DY,-1,0))
                  -(o(UA,1,0)+o(UA,0,0))*(o(DX,0,1)-o(DX,0,-1)))
              /(o(DX,0,0)*o(DY,0,0));                      // solver.f:145-152
  }
  // elf(i,j) after bcond(1) (bounds_forcing.f:21-39): W,E copies then S,N copies = the value
  // at the index clamped into the interior, times fsm; advua/advva of interior points
  template <class Op>
  POM_HD double phaseB(int i, int j, State& s, const Op& o, const Tile2& S, bool nbr) const {
    POM_DIMS;
    const int ic = i < 2 ? 2 : (i > imm1 ? imm1 : i);
    const int jc = j < 2 ? 2 : (j > jmm1 ? jmm1 : j);
    const int di = ic - i, dj = jc - j;
    const double ef=o(ELB,di,dj)+dte2*(pdiv(-(S(FUA,di+1,dj)-S(FUA,di,dj)+S(FVA,di,dj+1)-S(FVA,di,dj)),s.artc)
                                     -s.vflc);                            // advance.f:224-227
    if (nbr && i >= 2 && i <= imm1 && j >= 2 && j <= jmm1) {
      if (do_adv) {
        s.au=S(FXU,0,0)-S(FXU,-1,0)+S(FYU,0,1)-S(FYU,0,0);                // solver.f:65-66
        s.av=S(FXV,1,0)-S(FXV,0,0)+S(FYV,0,0)-S(FYV,0,-1);                // :116-117
      } else {
        s.au=advua(i,j); s.av=advva(i,j);
      }
      if (M2 && do_adv) {
        // mode 2 (solver.f:123-195): bottom stress from the 2-D velocities, curvature terms
        s.m2 = true;
        const double d00=o(D,0,0), uab00=o(UAB,0,0), vab00=o(VAB,0,0);
        const double qv=.25*(vab00+o(VAB,0,1)+o(VAB,-1,0)+o(VAB,-1,1));
        s.wub=-0.5*(cbc(i,j)+cbc(i-1,j))*sqrt(uab00*uab00+qv*qv)*uab00;          // :125-133
        const double qu=.25*(uab00+o(UAB,1,0)+o(UAB,0,-1)+o(UAB,1,-1));
        s.wvb=-0.5*(cbc(i,j)+cbc(i,j-1))*sqrt(vab00*vab00+qu*qu)*vab00;          // :135-143
        if (i >= 3)                                                               // :155-172 (n_west==-1)
          s.au=s.au-aru(i,j)*.25*(S(CURV,0,0)*d00*(o(VA,0,1)+o(VA,0,0))
                                  +S(CURV,-1,0)*o(D,-1,0)*(o(VA,-1,1)+o(VA,-1,0)));
        if (j >= 3)                                                               // :174-191 (n_south==-1)
          s.av=s.av+arv(i,j)*.25*(S(CURV,0,0)*d00*(o(UA,1,0)+o(UA,0,0))
                                  +S(CURV,0,-1)*o(D,0,-1)*(o(UA,1,-1)+o(UA,0,-1)));
      }
    }
    return ef*s.fsm0;
  }
  template <class Op>
  POM_HD void phaseC(int i, int j, State& s, const Op& o, const Tile2&, const Tile2& E) const {
    POM_DIMS;
    const bool jin = (j >= 2 && j <= jmm1), iin = (i >= 2 && i <= imm1);
    // every point-wise operand is loaded up front (before the first store), so that the
    // loads are in flight together instead of one L2 round trip after the other
    const double aru0=POM_LDG(&aru(i,j)), arv0=POM_LDG(&arv(i,j)), adx0=POM_LDG(&adx2d(i,j)), ady0=POM_LDG(&ady2d(i,j));
    const double drx0=POM_LDG(&drx2d(i,j)), dry0=POM_LDG(&dry2d(i,j)), wus0=POM_LDG(&wusurf(i,j));
    const double wub0=(M2 && s.m2) ? s.wub : POM_LDG(&wubot(i,j)), wvb0=(M2 && s.m2) ? s.wvb : POM_LDG(&wvbot(i,j));
    const double wvs0=POM_LDG(&wvsurf(i,j)), dum0=POM_LDG(&dum(i,j)), dvm0=POM_LDG(&dvm(i,j));
    const double egf0=egf(i,j), utf0=utf(i,j), vtf0=vtf(i,j);
    const double etf0=(iext >= c.isplit-1) ? etf(i,j) : 0.;
    const double ef=E(0,0);
    const double d00=o(D,0,0), el00=o(EL,0,0), elb00=o(ELB,0,0), h00=o(H,0,0), ua00=o(UA,0,0), va00=o(VA,0,0);
    elf(i,j)=ef;
    if (do_adv) { advua(i,j)=s.au; advva(i,j)=s.av; }
    if (M2 && s.m2) { wubot(i,j)=s.wub; wvbot(i,j)=s.wvb; }
    double un, vn;
    // ---- uaf(i,j) after bcond(2); cells never assigned keep uaf's content ----
    if (jin) {
      if (i == 1 || i == 2) {                                             // bounds_forcing.f:48-50
        const int q = 2 - i;
        un=ramp*(uabw(j)-rfw*sqrt(grav/o(D,q,0))*(o(EL,q,0)-elw(j)));
      } else if (i == im) {                                               // :57-59
        un=ramp*(uabe(j)+rfe*sqrt(grav/o(D,-1,0))*(o(EL,-1,0)-ele(j)));
      } else {                                                            // advance.f:239-260
        const double dW=o(D,-1,0), ar=aru0, efW=E(-1,0), elbW=o(ELB,-1,0), hW=o(H,-1,0);
        double r=adx0+s.au
                 -ar*.25
                   *(o(COR,0,0)*d00*(o(VA,0,1)+va00)
                    +o(COR,-1,0)*dW*(o(VA,-1,1)+o(VA,-1,0)))
                 +.25*grav*(o(DY,0,0)+o(DY,-1,0))
                   *(d00+dW)
                   *((1.-2.*alpha)*(el00-o(EL,-1,0))
                     +alpha*(elb00-elbW+ef-efW)
                     +o(EATM,0,0)-o(EATM,-1,0))
                 +drx0+ar*(wus0-wub0);
        un=pdiv((h00+elb00+hW+elbW)*ar*o(UAB,0,0)
                -4.*dte*r,
                (h00+ef+hW+efW)*ar);
      }
    } else if (iin) {
      un = (j == 1) ? uabs(i) : uabn(i);
    } else {
      un = uaf(i,j);   // four corners: never assigned (advance.f:237-262, bcond(2))
    }
    if (iin) {
      if (j == 1 || j == 2) {                                             // bounds_forcing.f:65-67
        const int q = 2 - j;
        vn=ramp*(vabs(i)-rfs*sqrt(grav/o(D,0,q))*(o(EL,0,q)-els(i)));
      } else if (j == jm) {                                               // :74-76
        vn=ramp*(vabn(i)+rfn*sqrt(grav/o(D,0,-1))*(o(EL,0,-1)-eln(i)));
      } else {                                                            // advance.f:266-286
        const double dS=o(D,0,-1), ar=arv0, efS=E(0,-1), elbS=o(ELB,0,-1), hS=o(H,0,-1);
        double r=ady0+s.av
                 +ar*.25
                   *(o(COR,0,0)*d00*(o(UA,1,0)+ua00)
                    +o(COR,0,-1)*dS*(o(UA,1,-1)+o(UA,0,-1)))
                 +.25*grav*(o(DX,0,0)+o(DX,0,-1))
                   *(d00+dS)
                   *((1.-2.*alpha)*(el00-o(EL,0,-1))
                     +alpha*(elb00-elbS+ef-efS)
                     +o(EATM,0,0)-o(EATM,0,-1))
                 +dry0+ar*(wvs0-wvb0);
        vn=pdiv((h00+elb00+hS+elbS)*ar*o(VAB,0,0)
                -4.*dte*r,
                (h00+ef+hS+efS)*ar);
      }
    } else if (jin) {
      vn = (i == 1) ? vabw(j) : vabe(j);
    } else {
      vn = vaf(i,j);
    }
    un=un*dum0;                                          // bounds_forcing.f:80-81
    vn=vn*dvm0;
    uaf(i,j)=un;
    vaf(i,j)=vn;
    // ---- etf accumulation on the last three substeps (advance.f:295-318) ----
    if (iext == c.isplit-2) etf(i,j)=.25*smoth*ef;
    else if (iext == c.isplit-1) etf(i,j)=etf0+.5*(1.-.5*smoth)*ef;
    else if (iext == c.isplit) etf(i,j)=(etf0+.5*ef)*s.fsm0;
    // ---- Asselin filter (advance.f:321-323): the filtered n-level goes to buffers no
    //      neighbour reads in this kernel ----
    A2(p.s2a,i,j)=ua00+.5*smoth*(o(UAB,0,0)-2.*ua00+un);
    A2(p.s2b,i,j)=va00+.5*smoth*(o(VAB,0,0)-2.*va00+vn);
    el2(i,j)=el00+.5*smoth*(elb00-2.*el00+ef);
    const double dn=h00+ef;                              // advance.f:326
    d2(i,j)=dn;
    // the four corners of uaf/vaf are never assigned by the reference and so persist;
    // seed the buffer that becomes uaf/vaf after the rotation (no stencil ever reads them)
    if (!jin && !iin) { ua(i,j)=un; va(i,j)=vn; }
    // ---- running means (advance.f:332-347) ----
    if (iext != c.isplit) {
      egf(i,j)=egf0+ef*ispi;
      if (i >= 2) utf(i,j)=utf0+un*(dn+(o(H,-1,0)+E(-1,0)))*isp2i;
      if (j >= 2) vtf(i,j)=vtf0+vn*(dn+(o(H,0,-1)+E(0,-1)))*isp2i;
    }
  }
};

void run_advave(Ctx* c, int j0, int j1) { launch_tma_tiles(c, AdvaveK(c), 1, c->g.im, j0, j1); }
void run_mode_inter_tail(Ctx* c, int j0, int j1) { launch_cols(c, ModeInterTailK(c), 1, c->g.im, j0, j1); }
}  // namespace pom

namespace pom {
// fused external substep; the caller rotates the time levels afterwards
void run_ext_step(Ctx* c, int iext, int do_adv, int j0, int j1) {
  if (c->c.mode == 2) launch_tile3(c, ExtStepK<true>(c, iext, do_adv), 1, c->g.im, j0, j1);
  else launch_tile3(c, ExtStepK<false>(c, iext, do_adv), 1, c->g.im, j0, j1);
}
}  // namespace pom
