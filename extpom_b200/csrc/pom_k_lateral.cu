// pom_k_lateral.cu -- lateral_viscosity group (advance.f:96-141):
//   advct  (solver.f:201-409)  + vertical integrals adx2d/ady2d (advance.f:161-162)
//   baropg (solver.f:848-940)  + vertical integrals drx2d/dry2d (advance.f:163-164)
//   Smagorinsky aam (advance.f:122-136) + aam2d (advance.f:165)
// One thread per (i,j) column marching k; flux intermediates of the reference
// (xflux,yflux,curv automatic arrays) are recomputed in registers instead of
// being staged through HBM.  Expression order follows the Fortran so that the
// results are bit-identical to a no-FMA evaluation.
#include "pom_core.h"
#include "pom_names.h"

namespace pom {

struct AdvctK : KBase {
  POM_KINFO("advct", 5, 2, 5, 2)
  using KBase::KBase;
  // solver.f:221-225; zero outside 2..imm1 x 2..jmm1 (zero fill :213)
  POM_HD double curv(int i, int j, int k) const {
    POM_DIMS;
    if (i < 2 || i > imm1 || j < 2 || j > jmm1) return 0.;
    return .25*((v(i,j+1,k)+v(i,j,k))*(dy(i+1,j)-dy(i-1,j))
                -(u(i+1,j,k)+u(i,j,k))*(dx(i,j+1)-dx(i,j-1)))
           /(dx(i,j)*dy(i,j));
  }
  // corner diffusive term shared by yflux (x-part, :261-270) and xflux (y-part, :348-357)
  POM_HD double cornerdiff(int i, int j, int k, double dy4, double dx4) const {
    double dtaam=.25*(dt(i,j)+dt(i-1,j)+dt(i,j-1)+dt(i-1,j-1))
                 *(aam(i,j,k)+aam(i-1,j,k)+aam(i,j-1,k)+aam(i-1,j-1,k));
    return dtaam*((ub(i,j,k)-ub(i,j-1,k))/dy4
                  +(vb(i,j,k)-vb(i-1,j,k))/dx4);
  }
  // x-part xflux(i,j,k), 1<=i<=imm1 (:237-239,258-260,272); xflux(1,j,k)=0 (:215)
  POM_HD double xfx(int i, int j, int k) const {
    if (i < 2) return 0.;
    double a=.125*((dt(i+1,j)+dt(i,j))*u(i+1,j,k)
                   +(dt(i,j)+dt(i-1,j))*u(i,j,k))
                  *(u(i+1,j,k)+u(i,j,k));
    a=a-dt(i,j)*aam(i,j,k)*2.*(ub(i+1,j,k)-ub(i,j,k))/dx(i,j);
    return dy(i,j)*a;
  }
  // x-part yflux(i,j,k), 2<=i<=imm1, 2<=j<=jm (:247-249,264-274)
  POM_HD double yfx(int i, int j, int k) const {
    double dy4=dy(i,j)+dy(i-1,j)+dy(i,j-1)+dy(i-1,j-1);
    double dx4=dx(i,j)+dx(i-1,j)+dx(i,j-1)+dx(i-1,j-1);
    double a=.125*((dt(i,j)+dt(i,j-1))*v(i,j,k)
                   +(dt(i-1,j)+dt(i-1,j-1))*v(i-1,j,k))
                  *(u(i,j,k)+u(i,j-1,k));
    a=a-cornerdiff(i,j,k,dy4,dx4);
    return .25*dx4*a;
  }
  // y-part xflux(i,j,k), 2<=i<=im, 2<=j<=jmm1 (:327-329,351-363)
  POM_HD double xfy(int i, int j, int k) const {
    double dy4=dy(i,j)+dy(i-1,j)+dy(i,j-1)+dy(i-1,j-1);
    double dx4=dx(i,j)+dx(i-1,j)+dx(i,j-1)+dx(i-1,j-1);
    double a=.125*((dt(i,j)+dt(i-1,j))*u(i,j,k)
                   +(dt(i,j-1)+dt(i-1,j-1))*u(i,j-1,k))
                  *(v(i,j,k)+v(i-1,j,k));
    a=a-cornerdiff(i,j,k,dy4,dx4);
    return .25*dy4*a;
  }
  // y-part yflux(i,j,k), 1<=j<=jmm1 (:337-339,358-364); yflux(i,1,k)=0 (:321)
  POM_HD double yfy(int i, int j, int k) const {
    if (j < 2) return 0.;
    double a=.125*((dt(i,j+1)+dt(i,j))*v(i,j+1,k)
                   +(dt(i,j)+dt(i,j-1))*v(i,j,k))
                  *(v(i,j+1,k)+v(i,j,k));
    a=a-dt(i,j)*aam(i,j,k)*2.*(vb(i,j+1,k)-vb(i,j,k))/dy(i,j);
    return dx(i,j)*a;
  }
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    const bool interior = (i >= 2 && i <= imm1 && j >= 2 && j <= jmm1);
    double sx = 0., sy = 0.;
    for (int k = 1; k <= kbm1; ++k) {
      double ax = 0., ay = 0.;
      if (interior) {
        double c00 = curv(i,j,k);
        ax=xfx(i,j,k)-xfx(i-1,j,k)+yfx(i,j+1,k)-yfx(i,j,k);            // :285-286
        if (i >= 3)                                                     // :293-300 (n_west==-1)
          ax=ax-aru(i,j)*.25*(c00*dt(i,j)*(v(i,j+1,k)+v(i,j,k))
                              +curv(i-1,j,k)*dt(i-1,j)*(v(i-1,j+1,k)+v(i-1,j,k)));
        ay=xfy(i+1,j,k)-xfy(i,j,k)+yfy(i,j,k)-yfy(i,j-1,k);            // :375-376
        if (j >= 3)                                                     // :383-390 (n_south==-1)
          ay=ay+arv(i,j)*.25*(c00*dt(i,j)*(u(i+1,j,k)+u(i,j,k))
                              +curv(i,j-1,k)*dt(i,j-1)*(u(i+1,j-1,k)+u(i,j-1,k)));
      }
      advx(i,j,k)=ax;
      advy(i,j,k)=ay;
      sx=sx+ax*dz(k);                                                   // advance.f:161-162
      sy=sy+ay*dz(k);
    }
    advx(i,j,kb)=0.;
    advy(i,j,kb)=0.;
    adx2d(i,j)=sx;
    ady2d(i,j)=sy;
  }
};

// solver.f:848-940.  rho is rewritten as (rho-rmean)+rmean (:854,937) into the
// alternate buffer rho2 (neighbours still read the old rho); the caller swaps.
struct BaropgK : KBase {
  POM_KINFO("baropg", 2, 3, 5, 2)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    const bool interior = (i >= 2 && i <= imm1 && j >= 2 && j <= jmm1);
    double sx = 0., sy = 0.;
    if (interior) {
      const double dtx=dt(i,j)+dt(i-1,j), dty=dt(i,j)+dt(i,j-1);
      const double ddx=dt(i,j)-dt(i-1,j), ddy=dt(i,j)-dt(i,j-1);
      const double dyx=dy(i,j)+dy(i-1,j), dxy=dx(i,j)+dx(i,j-1);
      // rho-rmean at (i,j),(i-1,j),(i,j-1), levels k-1 (a*) and k (b*)
      double a0=rho(i,j,1)-rmean(i,j,1);
      double ax=rho(i-1,j,1)-rmean(i-1,j,1);
      double ay=rho(i,j-1,1)-rmean(i,j-1,1);
      double px=.5*grav*(-zz(1))*dtx*(a0-ax);                          // :859-860
      double py=.5*grav*(-zz(1))*dty*(a0-ay);                          // :895-896
      for (int k = 1; k <= kbm1; ++k) {
        if (k >= 2) {
          double b0=rho(i,j,k)-rmean(i,j,k);
          double bx=rho(i-1,j,k)-rmean(i-1,j,k);
          double by=rho(i,j-1,k)-rmean(i,j-1,k);
          px=px+grav*.25*(zz(k-1)-zz(k))*dtx*(b0-bx+a0-ax)
               +grav*.25*(zz(k-1)+zz(k))*ddx*(b0+bx-a0-ax);           // :867-875
          py=py+grav*.25*(zz(k-1)-zz(k))*dty*(b0-by+a0-ay)
               +grav*.25*(zz(k-1)+zz(k))*ddy*(b0+by-a0-ay);           // :903-911
          a0=b0; ax=bx; ay=by;
        }
        double ox=ramp*(.25*dtx*px*dum(i,j)*dyx);                      // :883-885,931
        double oy=ramp*(.25*dty*py*dvm(i,j)*dxy);                      // :919-921,932
        drhox(i,j,k)=ox;
        drhoy(i,j,k)=oy;
        sx=sx+ox*dz(k);                                                 // advance.f:163-164
        sy=sy+oy*dz(k);
      }
      drhox(i,j,kb)=ramp*drhox(i,j,kb);                                 // :928-932 (k=kb)
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
      rho2(i,j,k)=(rho(i,j,k)-rmean(i,j,k))+rmean(i,j,k);               // :854,937
  }
};

// advance.f:122-136 + aam2d (advance.f:165)
struct SmagK : KBase {
  POM_KINFO("smagorinsky", 2, 1, 2, 1)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    const bool interior = (i >= 2 && i <= imm1 && j >= 2 && j <= jmm1);
    double sa = 0.;
    for (int k = 1; k <= kbm1; ++k) {
      double a;
      if (interior) {
        double a1=(u(i+1,j,k)-u(i,j,k))/dx(i,j);
        double a2=(v(i,j+1,k)-v(i,j,k))/dy(i,j);
        double a3=.25*(u(i,j+1,k)+u(i+1,j+1,k)-u(i,j-1,k)-u(i+1,j-1,k))/dy(i,j)
                 +.25*(v(i+1,j,k)+v(i+1,j+1,k)-v(i-1,j,k)-v(i-1,j+1,k))/dx(i,j);
        a=horcon*dx(i,j)*dy(i,j)*sqrt(a1*a1+a2*a2+.5*(a3*a3));
        aam(i,j,k)=a;
      } else {
        a=aam(i,j,k);
      }
      sa=sa+a*dz(k);
    }
    aam2d(i,j)=sa;
  }
};

void run_advct(Ctx* c, int j0, int j1) { launch_cols(c, AdvctK(c), 1, c->g.im, j0, j1); }
void run_baropg(Ctx* c, int j0, int j1) {
  launch_cols(c, BaropgK(c), 1, c->g.im, j0, j1);
  double* tmp = c->p.rho; c->p.rho = c->p.rho2; c->p.rho2 = tmp;
}
void run_smag(Ctx* c, int j0, int j1) { launch_cols(c, SmagK(c), 1, c->g.im, j0, j1); }

}  // namespace pom
