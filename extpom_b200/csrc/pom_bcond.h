// pom_bcond.h -- the open-boundary updates called INSIDE the internal step, as point functions shared
// by the fused kernels (pom_k_internal.cu: profq/proft/uv_filter upward sweeps) and by the stand-alone
// bcond / bcondorl kernels (pom_k_bcond.cu: the reference's entry points of the same names,
// pom/bounds_forcing.f:6,331).  Each takes the value a (b) the interior scheme left at (i,j,k) and
// returns what the boundary code assigns there BEFORE the mask pass; interior points are untouched.
// Corner precedence follows the reference's loop order: the east/west loop runs first over all j,
// the south/north loop then overwrites the corners (bounds_forcing.f:155-231, 260-313).
// Include after pom_names.h, inside a translation unit of kernel bodies.
#pragma once

namespace pom {

// Orlanski radiation value from the point `1` cell inside (xf1,xb1), two inside (x2),
// and the boundary point's own xb0 and x at one inside (x1)  (bounds_forcing.f:425-434)
POM_HD double orl(double xf1, double xb1, double x2, double xb0, double x1) {
  double denom=(xf1+xb1-2.*x2);
  if (denom == 0.) denom=0.01;
  double cl=(xb1-xf1)/denom;
  if (cl > 1.) cl=1.;
  if (cl < 0.) cl=0.;
  return (xb0*(1.-cl)+2.*cl*x1)/(1.+cl);
}

// bcond(4): upstream advection of T (a) and S (b) on the four open edges, with the vertical-advection
// correction for outflow (bounds_forcing.f:151-231)
template <class K>
POM_HD void bcond4_edge(const K& kk, int i, int j, int k, double& a, double& b) {
  const Geo& g = kk.g; const Ptrs& p = kk.p; const Consts& c = kk.c; const KTab& kt = kk.kt; (void)kt;
  POM_DIMS;
  const bool vadv = (k != 1 && k != kbm1);
  if (j == 1) {                                                     // south (:196-211)
    double u1=2.*v(i,2,k)*dti/(dy(i,1)+dy(i,2));
    if (u1 >= 0.) {
      a=t(i,1,k)-u1*(t(i,1,k)-tbs(i,k));
      b=s(i,1,k)-u1*(s(i,1,k)-sbs(i,k));
    } else {
      a=t(i,1,k)-u1*(t(i,2,k)-t(i,1,k));
      b=s(i,1,k)-u1*(s(i,2,k)-s(i,1,k));
      if (vadv) {
        double wm=.5*(w(i,2,k)+w(i,2,k+1))*dti/((zz(k-1)-zz(k+1))*dt(i,2));
        a=a-wm*(t(i,2,k-1)-t(i,2,k+1));
        b=b-wm*(s(i,2,k-1)-s(i,2,k+1));
      }
    }
  } else if (j == jm) {                                             // north (:214-229)
    double u1=2.*v(i,jm,k)*dti/(dy(i,jm)+dy(i,jmm1));
    if (u1 <= 0.) {
      a=t(i,jm,k)-u1*(tbn(i,k)-t(i,jm,k));
      b=s(i,jm,k)-u1*(sbn(i,k)-s(i,jm,k));
    } else {
      a=t(i,jm,k)-u1*(t(i,jm,k)-t(i,jmm1,k));
      b=s(i,jm,k)-u1*(s(i,jm,k)-s(i,jmm1,k));
      if (vadv) {
        double wm=.5*(w(i,jmm1,k)+w(i,jmm1,k+1))*dti/((zz(k-1)-zz(k+1))*dt(i,jmm1));
        a=a-wm*(t(i,jmm1,k-1)-t(i,jmm1,k+1));
        b=b-wm*(s(i,jmm1,k-1)-s(i,jmm1,k+1));
      }
    }
  } else if (i == im) {                                             // east (:158-173)
    double u1=2.*u(im,j,k)*dti/(dx(im,j)+dx(imm1,j));
    if (u1 <= 0.) {
      a=t(im,j,k)-u1*(tbe(j,k)-t(im,j,k));
      b=s(im,j,k)-u1*(sbe(j,k)-s(im,j,k));
    } else {
      a=t(im,j,k)-u1*(t(im,j,k)-t(imm1,j,k));
      b=s(im,j,k)-u1*(s(im,j,k)-s(imm1,j,k));
      if (vadv) {
        double wm=.5*(w(imm1,j,k)+w(imm1,j,k+1))*dti/((zz(k-1)-zz(k+1))*dt(imm1,j));
        a=a-wm*(t(imm1,j,k-1)-t(imm1,j,k+1));
        b=b-wm*(s(imm1,j,k-1)-s(imm1,j,k+1));
      }
    }
  } else if (i == 1) {                                              // west (:176-191)
    double u1=2.*u(2,j,k)*dti/(dx(1,j)+dx(2,j));
    if (u1 >= 0.) {
      a=t(1,j,k)-u1*(t(1,j,k)-tbw(j,k));
      b=s(1,j,k)-u1*(s(1,j,k)-sbw(j,k));
    } else {
      a=t(1,j,k)-u1*(t(2,j,k)-t(1,j,k));
      b=s(1,j,k)-u1*(s(2,j,k)-s(1,j,k));
      if (vadv) {
        double wm=.5*(w(2,j,k)+w(2,j,k+1))*dti/((zz(k-1)-zz(k+1))*dt(2,j));
        a=a-wm*(t(2,j,k-1)-t(2,j,k+1));
        b=b-wm*(s(2,j,k-1)-s(2,j,k+1));
      }
    }
  }
}

// bcond(6): upstream advection of q2 (a) and q2l (b) on the four open edges (bounds_forcing.f:257-311)
template <class K>
POM_HD void bcond6_edge(const K& kk, int i, int j, int k, double& a, double& b) {
  const Geo& g = kk.g; const Ptrs& p = kk.p; const Consts& c = kk.c; const KTab& kt = kk.kt; (void)kt;
  POM_DIMS;
  if (j == 1) {                                                     // south (:290-299)
    double u1=2.*v(i,2,k)*dti/(dy(i,1)+dy(i,2));
    if (u1 >= 0.) { a=q2(i,1,k)-u1*(q2(i,1,k)-small); b=q2l(i,1,k)-u1*(q2l(i,1,k)-small); }
    else { a=q2(i,1,k)-u1*(q2(i,2,k)-q2(i,1,k)); b=q2l(i,1,k)-u1*(q2l(i,2,k)-q2l(i,1,k)); }
  } else if (j == jm) {                                             // north (:302-311)
    double u1=2.*v(i,jm,k)*dti/(dy(i,jm)+dy(i,jmm1));
    if (u1 <= 0.) { a=q2(i,jm,k)-u1*(small-q2(i,jm,k)); b=q2l(i,jm,k)-u1*(small-q2l(i,jm,k)); }
    else { a=q2(i,jm,k)-u1*(q2(i,jm,k)-q2(i,jmm1,k)); b=q2l(i,jm,k)-u1*(q2l(i,jm,k)-q2l(i,jmm1,k)); }
  } else if (i == 1) {                                              // west (:264-273)
    double u1=2.*u(2,j,k)*dti/(dx(1,j)+dx(2,j));
    if (u1 >= 0.) { a=q2(1,j,k)-u1*(q2(1,j,k)-small); b=q2l(1,j,k)-u1*(q2l(1,j,k)-small); }
    else { a=q2(1,j,k)-u1*(q2(2,j,k)-q2(1,j,k)); b=q2l(1,j,k)-u1*(q2l(2,j,k)-q2l(1,j,k)); }
  } else if (i == im) {                                             // east (:276-285)
    double u1=2.*u(im,j,k)*dti/(dx(im,j)+dx(imm1,j));
    if (u1 <= 0.) { a=q2(im,j,k)-u1*(small-q2(im,j,k)); b=q2l(im,j,k)-u1*(small-q2l(im,j,k)); }
    else { a=q2(im,j,k)-u1*(q2(im,j,k)-q2(imm1,j,k)); b=q2l(im,j,k)-u1*(q2l(im,j,k)-q2l(imm1,j,k)); }
  }
}

// bcondorl(3): Orlanski radiation of the normal velocity, zero tangential velocity on the boundary
// line itself (bounds_forcing.f:418-474); a = uf, b = vf
template <class K>
POM_HD void bcondorl3_edge(const K& kk, int i, int j, int k, double& a, double& b) {
  const Geo& g = kk.g; const Ptrs& p = kk.p; const KTab& kt = kk.kt; (void)kt;
  POM_DIMS;
  const bool jin = (j >= 2 && j <= jmm1), iin = (i >= 2 && i <= imm1);
  if (jin) {
    if (i == im) {                                                  // east (:425-434)
      a=orl(uf(im-1,j,k),ub(im-1,j,k),u(im-2,j,k),ub(im,j,k),u(im-1,j,k));
      b=0.;
    } else if (i == 2 || i == 1) {                                  // west (:437-447)
      a=orl(uf(3,j,k),ub(3,j,k),u(4,j,k),ub(2,j,k),u(3,j,k));
      if (i == 1) b=0.;
    }
  }
  if (iin) {
    if (j == jm) {                                                  // north (:465-474)
      b=orl(vf(i,jm-1,k),vb(i,jm-1,k),v(i,jm-2,k),vb(i,jm,k),v(i,jm-1,k));
      a=0.;
    } else if (j == 2 || j == 1) {                                  // south (:452-462)
      b=orl(vf(i,3,k),vb(i,3,k),v(i,4,k),vb(i,2,k),v(i,3,k));
      if (j == 1) a=0.;
    }
  }

}

}  // namespace pom
