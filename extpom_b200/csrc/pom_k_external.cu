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
struct AdvaveK : KBase {
  POM_KINFO("advave", 0, 0, 8, 2)
  using KBase::KBase;
  POM_HD double dx4(int i, int j) const { return dx(i,j)+dx(i-1,j)+dx(i,j-1)+dx(i-1,j-1); }
  POM_HD double dy4(int i, int j) const { return dy(i,j)+dy(i-1,j)+dy(i,j-1)+dy(i-1,j-1); }
  // tps(i,j), 2<=i<=im, 2<=j<=jm (solver.f:47-53); reused by the v half (:106)
  POM_HD double tpsf(int i, int j) const {
    return .25*(d(i,j)+d(i-1,j)+d(i,j-1)+d(i-1,j-1))
           *(aam2d(i,j)+aam2d(i,j-1)+aam2d(i-1,j)+aam2d(i-1,j-1))
           *((uab(i,j)-uab(i,j-1))/dy4(i,j)
             +(vab(i,j)-vab(i-1,j))/dx4(i,j));
  }
  // u half: fluxua(i,j) for 1<=i<=imm1, 2<=j<=jm (:22-24,39-41,54); fluxua(1,j)=0
  POM_HD double fxu(int i, int j) const {
    if (i < 2) return 0.;
    double a=.125*((d(i+1,j)+d(i,j))*ua(i+1,j)+(d(i,j)+d(i-1,j))*ua(i,j))
                 *(ua(i+1,j)+ua(i,j));
    a=a-d(i,j)*2.*aam2d(i,j)*(uab(i+1,j)-uab(i,j))/dx(i,j);
    return a*dy(i,j);
  }
  // u half: fluxva(i,j), 2<=i<=im, 2<=j<=jm (:30-32,55-56)
  POM_HD double fyu(int i, int j) const {
    double a=.125*((d(i,j)+d(i,j-1))*va(i,j)+(d(i-1,j)+d(i-1,j-1))*va(i-1,j))
                 *(ua(i,j)+ua(i,j-1));
    return (a-tpsf(i,j))*.25*dx4(i,j);
  }
  // v half: fluxua(i,j), 2<=i<=im, 2<=j<=jm (:80-82,106-107)
  POM_HD double fxv(int i, int j) const {
    double a=.125*((d(i,j)+d(i-1,j))*ua(i,j)+(d(i,j-1)+d(i-1,j-1))*ua(i,j-1))
                 *(va(i-1,j)+va(i,j));
    return (a-tpsf(i,j))*.25*dy4(i,j);
  }
  // v half: fluxva(i,j), 1<=j<=jmm1 (:88-90,97-99,105); fluxva(i,1)=0
  POM_HD double fyv(int i, int j) const {
    if (j < 2) return 0.;
    double a=.125*((d(i,j+1)+d(i,j))*va(i,j+1)+(d(i,j)+d(i,j-1))*va(i,j))
                 *(va(i,j+1)+va(i,j));
    a=a-d(i,j)*2.*aam2d(i,j)*(vab(i,j+1)-vab(i,j))/dy(i,j);
    return a*dx(i,j);
  }
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    double au = 0., av = 0.;
    if (i >= 2 && i <= imm1 && j >= 2 && j <= jmm1) {
      au=fxu(i,j)-fxu(i-1,j)+fyu(i,j+1)-fyu(i,j);       // :65-66
      av=fxv(i+1,j)-fxv(i,j)+fyv(i,j)-fyv(i,j-1);       // :116-117
    }
    advua(i,j)=au;
    advva(i,j)=av;
  }
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

void run_advave(Ctx* c, int j0, int j1) { launch_cols(c, AdvaveK(c), 1, c->g.im, j0, j1); }
void run_mode_inter_tail(Ctx* c, int j0, int j1) { launch_cols(c, ModeInterTailK(c), 1, c->g.im, j0, j1); }
void run_ext_elf(Ctx* c, int j0, int j1) { launch_cols(c, ExtElfK(c), 1, c->g.im, j0, j1); }

// After the kernel: time rotation by pointer swaps (advance.f:324-330)
void run_ext_uv(Ctx* c, int iext, int j0, int j1) {
  launch_cols(c, ExtUvK(c, iext), 1, c->g.im, j0, j1);
  Ptrs& p = c->p;
  double* t;
  // (uab,ua,uaf) <- (filtered[in uab], uaf, old ua as next uaf buffer)
  t = p.ua; p.ua = p.uaf; p.uaf = t;
  t = p.va; p.va = p.vaf; p.vaf = t;
  // (elb,el,elf,el2) <- (el2[filtered], elf, old el, old elb)
  t = p.elb; p.elb = p.el2; p.el2 = t;
  t = p.el; p.el = p.elf; p.elf = t;
  t = p.d; p.d = p.d2; p.d2 = t;
}

}  // namespace pom
