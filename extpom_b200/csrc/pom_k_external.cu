// pom_k_external.cu -- the 2-D external mode (advance.f:144-353):
//   advave (solver.f:6-121), mode_interaction tail (advance.f:172-196),
//   mode_external (advance.f:211-350) with bcond(1), bcond(2)
//   (bounds_forcing.f:18-83) folded into the producing kernels.
// Three kernels per substep: ExtElfK -> AdvaveK -> ExtUvK.  The reference's
// Asselin filter + time rotation by whole-array copies (advance.f:321-330)
// becomes a write of the filtered field into the dead "b" buffer followed by a
// pointer rotation on the host.
#include "pom_core.h"
#include "pom_names.h"

namespace pom {

// ------------------------------------------------------------------ advave ----
// Tile kernel (single "level"): every thread evaluates the four fluxes of its own point --
// u-half fluxua (FXU) / fluxva (FYU), v-half fluxua (FXV) / fluxva (FYV) -- once, sharing
// tps (solver.f:47-53,106) between the two halves like the reference does.
struct AdvaveK : KBase {
  POM_KINFO("advave", 0, 0, 8, 2)
  using KBase::KBase;
  static constexpr int NV = 4, HL = 1, HR = 1, HB = 1, HT = 1, TY = 16, MINB = 2;
  enum { FXU, FYU, FXV, FYV };
  struct State { bool interior; };
  struct Regs {};
  POM_HD int k0() const { return 1; }
  POM_HD int k1() const { return 1; }
  POM_HD void pre(int i, int j, bool, bool out, State& s) const {
    POM_DIMS;
    s.interior = out && i >= 2 && i <= imm1 && j >= 2 && j <= jmm1;
  }
  POM_HD void fetch(int, int, int, const State&, Regs&) const {}
  POM_HD void stage(int i, int j, int, State&, const Regs&, double* v) const {
    POM_DIMS;
    const int jlo = g.joff + 1, jhi = g.joff + g.jml;
    if (i < 2 || j < 2 || j - 1 < jlo) return;
    const double d00=d(i,j), dW=d(i-1,j), dS=d(i,j-1), dSW=d(i-1,j-1);
    const double ua00=ua(i,j), va00=va(i,j), uaS=ua(i,j-1), vaW=va(i-1,j);
    const double dx4=dx(i,j)+dx(i-1,j)+dx(i,j-1)+dx(i-1,j-1);
    const double dy4=dy(i,j)+dy(i-1,j)+dy(i,j-1)+dy(i-1,j-1);
    // tps(i,j), 2<=i<=im, 2<=j<=jm (:47-53)
    const double tp=.25*(d00+dW+dS+dSW)
                    *(aam2d(i,j)+aam2d(i,j-1)+aam2d(i-1,j)+aam2d(i-1,j-1))
                    *((uab(i,j)-uab(i,j-1))/dy4+(vab(i,j)-vab(i-1,j))/dx4);
    {   // u half fluxva (:30-32,55-56) and v half fluxua (:80-82,106-107)
      double a=.125*((d00+dS)*va00+(dW+dSW)*vaW)*(ua00+uaS);
      v[FYU]=(a-tp)*.25*dx4;
      double b=.125*((d00+dW)*ua00+(dS+dSW)*uaS)*(vaW+va00);
      v[FXV]=(b-tp)*.25*dy4;
    }
    if (i <= imm1) {   // u half fluxua (:22-24,39-41,54)
      const double dE=d(i+1,j), uaE=ua(i+1,j);
      double a=.125*((dE+d00)*uaE+(d00+dW)*ua00)*(uaE+ua00);
      a=a-d00*2.*aam2d(i,j)*(uab(i+1,j)-uab(i,j))/dx(i,j);
      v[FXU]=a*dy(i,j);
    }
    if (j <= jmm1 && j + 1 <= jhi) {   // v half fluxva (:88-90,97-99,105)
      const double dN=d(i,j+1), vaN=va(i,j+1);
      double a=.125*((dN+d00)*vaN+(d00+dS)*va00)*(vaN+va00);
      a=a-d00*2.*aam2d(i,j)*(vab(i,j+1)-vab(i,j))/dy(i,j);
      v[FYV]=a*dx(i,j);
    }
  }
  POM_HD void combine(int i, int j, int, State& s, const Regs&, const Tile& tl) const {
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

// ------------------------------------------------------------ elf + bcond(1) ----
struct ExtElfK : KBase {
  POM_KINFO("ext_elf", 0, 0, 9, 1)
  using KBase::KBase;
  POM_HD double fua(int i, int j) const {   // advance.f:213-214
    return .25*(d(i,j)+d(i-1,j))*(dy(i,j)+dy(i-1,j))*ua(i,j);
  }
  POM_HD double fva(int i, int j) const {   // advance.f:215-216
    return .25*(d(i,j)+d(i,j-1))*(dx(i,j)+dx(i,j-1))*va(i,j);
  }
  POM_HD double elfi(int i, int j) const {  // advance.f:224-227 at an interior point
    return elb(i,j)+dte2*(-(fua(i+1,j)-fua(i,j)+fva(i,j+1)-fva(i,j))/art(i,j)-vfluxf(i,j));
  }
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    // bcond(1) (bounds_forcing.f:21-37): W,E copies then S,N copies = value at the
    // index clamped into the interior; then *fsm (:39)
    int ic = i < 2 ? 2 : (i > imm1 ? imm1 : i);
    int jc = j < 2 ? 2 : (j > jmm1 ? jmm1 : j);
    elf(i,j)=elfi(ic,jc)*fsm(i,j);
  }
};

// ------------------------------- uaf/vaf + bcond(2) + etf + filter + running means
struct ExtUvK : KBase {
  POM_KINFO("ext_uv", 0, 0, 32, 10)
  using KBase::KBase;
  int iext;
  ExtUvK(const Ctx* x, int ie) : KBase(x), iext(ie) {}
  // advance.f:239-260 at 2<=i<=im, 2<=j<=jmm1
  POM_HD double uafi(int i, int j) const {
    double r=adx2d(i,j)+advua(i,j)
             -aru(i,j)*.25
               *(cor(i,j)*d(i,j)*(va(i,j+1)+va(i,j))
                +cor(i-1,j)*d(i-1,j)*(va(i-1,j+1)+va(i-1,j)))
             +.25*grav*(dy(i,j)+dy(i-1,j))
               *(d(i,j)+d(i-1,j))
               *((1.-2.*alpha)*(el(i,j)-el(i-1,j))
                 +alpha*(elb(i,j)-elb(i-1,j)+elf(i,j)-elf(i-1,j))
                 +e_atmos(i,j)-e_atmos(i-1,j))
             +drx2d(i,j)+aru(i,j)*(wusurf(i,j)-wubot(i,j));
    return ((h(i,j)+elb(i,j)+h(i-1,j)+elb(i-1,j))*aru(i,j)*uab(i,j)
            -4.*dte*r)
           /((h(i,j)+elf(i,j)+h(i-1,j)+elf(i-1,j))*aru(i,j));
  }
  // advance.f:266-286 at 2<=i<=imm1, 2<=j<=jm
  POM_HD double vafi(int i, int j) const {
    double r=ady2d(i,j)+advva(i,j)
             +arv(i,j)*.25
               *(cor(i,j)*d(i,j)*(ua(i+1,j)+ua(i,j))
                +cor(i,j-1)*d(i,j-1)*(ua(i+1,j-1)+ua(i,j-1)))
             +.25*grav*(dx(i,j)+dx(i,j-1))
               *(d(i,j)+d(i,j-1))
               *((1.-2.*alpha)*(el(i,j)-el(i,j-1))
                 +alpha*(elb(i,j)-elb(i,j-1)+elf(i,j)-elf(i,j-1))
                 +e_atmos(i,j)-e_atmos(i,j-1))
             +dry2d(i,j)+arv(i,j)*(wvsurf(i,j)-wvbot(i,j));
    return ((h(i,j)+elb(i,j)+h(i,j-1)+elb(i,j-1))*arv(i,j)*vab(i,j)
            -4.*dte*r)
           /((h(i,j)+elf(i,j)+h(i,j-1)+elf(i,j-1))*arv(i,j));
  }
  // bcond(2) Flather values (bounds_forcing.f:48-50,57-59,65-67,74-76)
  POM_HD double uaf_w(int j) const { return ramp*(uabw(j)-rfw*sqrt(grav/d(2,j))*(el(2,j)-elw(j))); }
  POM_HD double uaf_e(int j) const { const int imm1=g.im-1;
    return ramp*(uabe(j)+rfe*sqrt(grav/d(imm1,j))*(el(imm1,j)-ele(j))); }
  POM_HD double vaf_s(int i) const { return ramp*(vabs(i)-rfs*sqrt(grav/d(i,2))*(el(i,2)-els(i))); }
  POM_HD double vaf_n(int i) const { const int jmm1=g.jmg-1;
    return ramp*(vabn(i)+rfn*sqrt(grav/d(i,jmm1))*(el(i,jmm1)-eln(i))); }

  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    const bool jin = (j >= 2 && j <= jmm1), iin = (i >= 2 && i <= imm1);
    // ---- uaf(i,j) after bcond(2); cells never assigned keep uaf's content ----
    double un, vn;
    if (jin) {
      if (i == 1 || i == 2) un = uaf_w(j);
      else if (i == im) un = uaf_e(j);
      else un = uafi(i,j);
    } else if (iin) {
      un = (j == 1) ? uabs(i) : uabn(i);
    } else {
      un = uaf(i,j);   // four corners: never assigned (advance.f:237-262, bcond(2))
    }
    if (iin) {
      if (j == 1 || j == 2) vn = vaf_s(i);
      else if (j == jm) vn = vaf_n(i);
      else vn = vafi(i,j);
    } else if (jin) {
      vn = (i == 1) ? vabw(j) : vabe(j);
    } else {
      vn = vaf(i,j);
    }
    un=un*dum(i,j);                                      // bounds_forcing.f:80-81
    vn=vn*dvm(i,j);
    uaf(i,j)=un;
    vaf(i,j)=vn;
    // ---- etf accumulation on the last three substeps (advance.f:295-318) ----
    const double ef=elf(i,j);
    if (iext == c.isplit-2) etf(i,j)=.25*smoth*ef;
    else if (iext == c.isplit-1) etf(i,j)=etf(i,j)+.5*(1.-.5*smoth)*ef;
    else if (iext == c.isplit) etf(i,j)=(etf(i,j)+.5*ef)*fsm(i,j);
    // ---- Asselin filter (advance.f:321-323); filtered n-level goes to the dead
    //      "b" buffers; uab,vab are only read at (i,j) in this kernel, elb/d are
    //      read at neighbours so their new values go to el2/d2 ----
    uab(i,j)=ua(i,j)+.5*smoth*(uab(i,j)-2.*ua(i,j)+un);
    vab(i,j)=va(i,j)+.5*smoth*(vab(i,j)-2.*va(i,j)+vn);
    el2(i,j)=el(i,j)+.5*smoth*(elb(i,j)-2.*el(i,j)+ef);
    const double dn=h(i,j)+ef;                           // advance.f:326
    d2(i,j)=dn;
    // the four corners of uaf/vaf are never assigned by the reference and so persist;
    // seed the buffer that becomes uaf/vaf after the rotation (read only by this thread)
    if (!jin && !iin) { ua(i,j)=un; va(i,j)=vn; }
    // ---- running means (advance.f:332-347) ----
    if (iext != c.isplit) {
      egf(i,j)=egf(i,j)+ef*ispi;
      if (i >= 2) utf(i,j)=utf(i,j)+un*(dn+(h(i-1,j)+elf(i-1,j)))*isp2i;
      if (j >= 2) vtf(i,j)=vtf(i,j)+vn*(dn+(h(i,j-1)+elf(i,j-1)))*isp2i;
    }
  }
};

void run_advave(Ctx* c, int j0, int j1) { launch_tiles(c, AdvaveK(c), 1, c->g.im, j0, j1); }
void run_mode_inter_tail(Ctx* c, int j0, int j1) { launch_cols(c, ModeInterTailK(c), 1, c->g.im, j0, j1); }
void run_ext_elf(Ctx* c, int j0, int j1) { launch_cols(c, ExtElfK(c), 1, c->g.im, j0, j1); }

// The caller then rotates the time levels by pointer swaps (advance.f:324-330):
// ua<->uaf, va<->vaf, elb<->el2, el<->elf, d<->d2
void run_ext_uv(Ctx* c, int iext, int j0, int j1) { launch_cols(c, ExtUvK(c, iext), 1, c->g.im, j0, j1); }

}  // namespace pom
