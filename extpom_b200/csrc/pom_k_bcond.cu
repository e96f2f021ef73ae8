// pom_k_bcond.cu -- bcond(idx) / bcondorl(idx) (pom/bounds_forcing.f:6-324, 331-590) and smol_adif
// (pom/solver.f:1880-1967) as STAND-ALONE kernels: the reference's entry points of the same names for
// drivers that call them one by one (the gfortran-ABI library libpomgpu_f, unit-level parity tests).
// pomgpu_step never launches these: inside the step the same point functions (pom_bcond.h) are fused
// into the upward sweeps of profq / proft / uv_filter and into the external substep kernel.
//
// The reference assigns the edge cells first (from interior cells of the UNMASKED arrays) and then
// multiplies whole arrays by the masks; where an edge formula reads the array it is about to mask
// (elf, uf, vf), the edge assignment and the mask pass are two launches, like the two loop nests.
#include "pom_core.h"
#include "pom_names.h"
#include "pom_bcond.h"

namespace pom {

// bcond(1), edge part (bounds_forcing.f:21-37): zero-gradient elf on the four physical edges; W,E
// copies run over all j, then S,N overwrite the corners = the value at the index clamped inside
struct Bcond1EdgeK : KBase {
  POM_KINFO("bcond1_edge", 0, 0, 1, 1)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    const int ic = i < 2 ? 2 : (i > imm1 ? imm1 : i);
    const int jc = j < 2 ? 2 : (j > jmm1 ? jmm1 : j);
    if (ic != i || jc != j) elf(i,j)=elf(ic,jc);
  }
};
struct Bcond1MaskK : KBase {   // :39
  POM_KINFO("bcond1_mask", 0, 0, 2, 1)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const { elf(i,j)=elf(i,j)*fsm(i,j); }
};

// bcond(2) (bounds_forcing.f:47-81): Flather-type normal velocity on the four open edges from the
// prescribed uabw/uabe/vabs/vabn and elw/ele/els/eln, prescribed tangential velocity, then the masks.
// Every assigned value depends on d, el and the boundary arrays only, so one launch does both parts.
struct Bcond2K : KBase {
  POM_KINFO("bcond2", 0, 0, 6, 2)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    const bool jin = (j >= 2 && j <= jmm1), iin = (i >= 2 && i <= imm1);
    double un=uaf(i,j), vn=vaf(i,j);
    if (jin) {
      if (i == 1 || i == 2) un=ramp*(uabw(j)-rfw*sqrt(grav/d(2,j))*(el(2,j)-elw(j)));                 // :48-51
      else if (i == im) un=ramp*(uabe(j)+rfe*sqrt(grav/d(imm1,j))*(el(imm1,j)-ele(j)));              // :57-59
      if (i == 1) vn=vabw(j);                                                                        // :52
      else if (i == im) vn=vabe(j);                                                                  // :60
    }
    if (iin) {
      if (j == 1 || j == 2) vn=ramp*(vabs(i)-rfs*sqrt(grav/d(i,2))*(el(i,2)-els(i)));                 // :65-68
      else if (j == jm) vn=ramp*(vabn(i)+rfn*sqrt(grav/d(i,jmm1))*(el(i,jmm1)-eln(i)));              // :74-76
      if (j == 1) un=uabs(i);                                                                        // :69
      else if (j == jm) un=uabn(i);                                                                  // :77
    }
    uaf(i,j)=un*dum(i,j);                                                                            // :80-81
    vaf(i,j)=vn*dvm(i,j);
  }
};

// bcond(4) (bounds_forcing.f:151-242): T -> uf, S -> vf.  The edge values read t, s, u, v, w (never
// uf, vf), so edge assignment and mask are one launch.
struct Bcond4K : KBase {
  POM_KINFO("bcond4", 2, 2, 1, 0)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    const double m=fsm(i,j);
    for (int k = 1; k <= kbm1; ++k) {
      double a=uf(i,j,k), b=vf(i,j,k);
      bcond4_edge(*this, i, j, k, a, b);
      uf(i,j,k)=a*m;                                                      // :236-237
      vf(i,j,k)=b*m;
    }
  }
};

// bcond(5) / bcondorl(5) (bounds_forcing.f:244-255, 550-561): w*fsm for k=1..kbm1
struct Bcond5K : KBase {
  POM_KINFO("bcond5", 1, 1, 1, 0)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    const double m=fsm(i,j);
    for (int k = 1; k <= kbm1; ++k) w(i,j,k)=w(i,j,k)*m;
  }
};

// bcond(6) (bounds_forcing.f:257-324): q2 -> uf, q2l -> vf, k=1..kb, then *fsm + 1e-10
struct Bcond6K : KBase {
  POM_KINFO("bcond6", 2, 2, 1, 0)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    const double m=fsm(i,j);
    for (int k = 1; k <= kb; ++k) {
      double a=uf(i,j,k), b=vf(i,j,k);
      bcond6_edge(*this, i, j, k, a, b);
      uf(i,j,k)=a*m+1.e-10;                                               // :318-319
      vf(i,j,k)=b*m+1.e-10;
    }
  }
};

// bcondorl(3), edge part (bounds_forcing.f:418-474): the Orlanski values read uf / vf one cell
// inside, which the mask pass of the same call multiplies afterwards -> two launches.  Only cells on
// the boundary lines (and the second line on the west / south side) are assigned here.
struct BcondOrl3EdgeK : KBase {
  POM_KINFO("bcondorl3_edge", 0, 0, 0, 0)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    if (i > 2 && i < im && j > 2 && j < jm) return;
    for (int k = 1; k <= kbm1; ++k) {
      double a=uf(i,j,k), b=vf(i,j,k);
      bcondorl3_edge(*this, i, j, k, a, b);
      uf(i,j,k)=a;
      vf(i,j,k)=b;
    }
  }
};
struct BcondOrl3MaskK : KBase {   // :476-485
  POM_KINFO("bcondorl3_mask", 2, 2, 2, 0)
  using KBase::KBase;
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    const double mu=dum(i,j), mv=dvm(i,j);
    for (int k = 1; k <= kbm1; ++k) {
      uf(i,j,k)=uf(i,j,k)*mu;
      vf(i,j,k)=vf(i,j,k)*mv;
    }
  }
};

// the mask with which smol_adif begins (solver.f:1898-1900): ff*fsm over k=1..kb
struct MaskFsmK : KBase {
  POM_KINFO("mask_fsm", 1, 1, 1, 0)
  double* ff_;
  MaskFsmK(const Ctx* x, double* ff) : KBase(x), ff_(ff) {}
  POM_HD void operator()(int i, int j) const {
    POM_DIMS;
    const double m=fsm(i,j);
    for (int k = 1; k <= kb; ++k) A3(ff_,i,j,k)=A3(ff_,i,j,k)*m;
  }
};

#define ALLI 1, c->g.im
// the Orlanski edge kernel of bcondorl(3) reads uf, vf up to 3 cells inside and writes the boundary
// lines: it runs on the strips that hold a physical south / north edge and on every strip for the
// west / east columns; all of it is expressed on global indices, so any window is correct
int run_bcond(Ctx* c, int idx, int orl, int j0, int j1) {
  if (!orl) {
    switch (idx) {
      case 1: launch_cols(c, Bcond1EdgeK(c), ALLI, j0, j1); launch_cols(c, Bcond1MaskK(c), ALLI, j0, j1); return 0;
      case 2: launch_cols(c, Bcond2K(c), ALLI, j0, j1); return 0;
      case 4: launch_cols(c, Bcond4K(c), ALLI, j0, j1); return 0;
      case 5: launch_cols(c, Bcond5K(c), ALLI, j0, j1); return 0;
      case 6: launch_cols(c, Bcond6K(c), ALLI, j0, j1); return 0;
      default: return 2;   // bcond(3): never called by the step (advance.f:231,290,414,442)
    }
  }
  switch (idx) {
    case 3: launch_cols(c, BcondOrl3EdgeK(c), ALLI, j0, j1); launch_cols(c, BcondOrl3MaskK(c), ALLI, j0, j1); return 0;
    case 5: launch_cols(c, Bcond5K(c), ALLI, j0, j1); return 0;
    default: return 2;     // bcondorl(1,2,4,6): never called by the step (advance.f:398,464)
  }
}
void run_mask_fsm(Ctx* c, double* ff, int j0, int j1) { launch_cols(c, MaskFsmK(c, ff), ALLI, j0, j1); }

}  // namespace pom
