/* pomgpu_f.h -- the gfortran-ABI face of the B200 time-stepping core: libpomgpu_f.so.
 *
 * extPOM's hot path is a set of Fortran 77 EXTERNAL subroutines without interfaces or modules that
 * talk through COMMON blocks (pom.h_dist).  Under gfortran that ABI is: lower-case name + trailing
 * underscore, every argument by reference, arrays as bare `double*` (column-major, 1-based),
 * `integer` = 32-bit int, no hidden arguments (no character arguments on this path); a COMMON
 * block /blk3d/ is the global data symbol `blk3d_` whose members lie in declaration order.
 * libpomgpu_f exports exactly those symbols, so that the reference's driver (pom.f, initialize.f,
 * bounds_forcing.f, io_pnetcdf.F, parallel_mpi.f and the glue part of advance.f) links against it
 * INSTEAD of solver.o and the four step routines of advance.f -- no source edit, no ISO_C_BINDING
 * shim (INTEGRATION.md has the link line).  It is plain host C++ over the C ABI of pomgpu.h.
 *
 * State.  The library binds to the driver's COMMON blocks (weak references to blksiz_, blkcon_,
 * blk1d_, blk2d_, blk3d_, bdry_; resolved against the executable at load time) through the member
 * table generated from the model's own pom.h (include/pom_common_layout.h).  The driver owns host
 * memory; the library owns the HBM copies:
 *   step level  (lateral_viscosity_, mode_interaction_, mode_external_, mode_internal_): RESIDENT.
 *     The first call pushes the whole COMMON state once; lateral_viscosity_ then pushes, every step,
 *     blkcon and the forcing / open-boundary arrays the Fortran driver refreshes before the hot path
 *     (bounds_forcing.f:844-865,908-909,954-955,978); mode_internal_ ends with what the driver reads
 *     every step -- vaf for check_velocity (advance.f:52,619-629) and error_status -- and, on print
 *     and restart steps (mod(iint,iprint)==0, mod(iint,irestart)==0, advance.f:35-49), pulls the
 *     whole state back into COMMON.
 *   routine level (everything else below): UNIT MODE.  H2D of the arrays the routine reads (COMMON
 *     members and the array arguments, by address), the kernel, D2H of the arrays it writes: what
 *     `initialize` needs for dens / baropg (initialize.f:416,425,502-505) and what a per-routine
 *     parity test needs.  An argument that is not a COMMON member (a local array of the caller) is
 *     staged through a scratch field.  If step-level calls have left the device ahead of the host,
 *     the whole state is pulled first.
 * Errors follow the reference: blkcon's error_status=1 and a message on stderr; never exit().
 */
#ifndef POMGPU_F_H
#define POMGPU_F_H
#ifdef __cplusplus
extern "C" {
#endif

/* ---- step level: pom/advance.f ------------------------------------------------------- */
void lateral_viscosity_(void);   /* advance.f:96   advct, baropg|baropg_mcc, Smagorinsky aam   */
void mode_interaction_(void);    /* advance.f:144  vertical integrals, advave, egf/utf/vtf     */
void mode_external_(void);       /* advance.f:205  one external substep; iext from blkcon      */
void mode_internal_(void);       /* advance.f:356  the 3-D step; iint from blkcon              */

/* ---- routine level: pom/solver.f ---------------------------------------------------------- */
void advave_(void);              /* solver.f:6    */
void advct_(void);               /* solver.f:201  */
void advq_(double* qb, double* q, double* qf);                            /* solver.f:411  */
void advt1_(double* fb, double* f, double* fclim, double* ff);            /* solver.f:480  */
void advt2_(double* fb, double* f, double* fclim, double* ff);            /* solver.f:577  */
void advu_(void);                /* solver.f:734  */
void advv_(void);                /* solver.f:791  */
void baropg_(void);              /* solver.f:848  */
void baropg_mcc_(void);          /* solver.f:943  */
void dens_(double* si, double* ti, double* rhoo);                         /* solver.f:1162 */
void profq_(void);               /* solver.f:1212 */
void proft_(double* f, double* wfsurf, double* fsurf, int* nbc);          /* solver.f:1541 */
void profu_(void);               /* solver.f:1686 */
void profv_(void);               /* solver.f:1783 */
void smol_adif_(double* xmassflux, double* ymassflux, double* zwflux, double* ff);   /* solver.f:1880 */
void vertvl_(void);              /* solver.f:1970 */
void realvertvl_(void);          /* solver.f:2024 */

/* ---- routine level: pom/bounds_forcing.f -------------------------------------------------- */
void bcond_(int* idx);           /* bounds_forcing.f:6    idx 1,2,4,5,6 (3 is never called)    */
void bcondorl_(int* idx);        /* bounds_forcing.f:331  idx 3,5 (the ones advance.f calls)   */

/* ---- pom/parallel_mpi.f:154,242 -------------------------------------------------------------------
 * The halo swaps of a host array.  Inside the step the exchanges are the library's own business
 * (ghost rows of the j-strips, pom_halo.cu); these two exist so that a SINGLE-RANK build links
 * without parallel_mpi.o: with every neighbour -1 the reference's exchange is a no-op
 * (parallel_mpi.f:171,179,...), and so are these.  An MPI build keeps parallel_mpi.o, whose
 * definitions then take precedence. */
void exchange2d_mpi_(double* work, int* nx, int* ny);
void exchange3d_mpi_(double* work, int* nx, int* ny, int* nz);

/* ---- control (optional; callable from Fortran as `call pomgpu_f_...`) --------------------- */
/* extents the COMMON blocks were compiled with, when they differ from the pom.h the library's table
 * was generated from (include/pom_common_layout.h: POMF_IM_LOCAL, POMF_JM_LOCAL, POMF_KB); call
 * before any other entry point */
void pomgpu_f_set_dims_(const int* im_local, const int* jm_local, const int* kb);
void pomgpu_f_set_device_(const int* device);   /* CUDA device of this rank (default 0)        */
/* Several GPUs behind ONE Fortran process: the (single-rank) driver's domain is cut into n j-strips, strip r on
 * device `device + r`, inside the library; the COMMON arrays keep their global extents, pushes scatter their rows
 * (with the strips' ghost rows), pulls gather the owned rows, the halo exchanges between the strips are the
 * library's (device / peer copies).  Results are bitwise those of one device.  Also POMGPU_F_DEVICES=n in the
 * environment; n < 0: |n| strips all on `device` (tests on one GPU); POMGPU_F_GHOST = ghost rows per seam (8).
 * Routine-level entries other than dens_, baropg_, baropg_mcc_ then run on one device holding the whole domain.
 * Call before the first entry point. */
void pomgpu_f_set_devices_(const int* n);
void pomgpu_f_push_all_(void);                  /* COMMON -> HBM, everything                   */
void pomgpu_f_pull_all_(void);                  /* HBM -> COMMON, everything (output, restart) */
void pomgpu_f_push_(const double* member);      /* one COMMON array, by address (e.g. trstrb after restore_interior re-read it) */
void pomgpu_f_pull_(double* member);
/* restore_interior (bounds_forcing.f:1023-1118) is called by the reference from INSIDE mode_internal (advance.f:452)
 * and mixes netCDF reads with arithmetic.  The arithmetic (time interpolation of the records, nudging of t, tb, s, sb,
 * masks: :1083-1118) is part of the device step; the reads stay in Fortran: when the executable defines
 *     subroutine restore_interior_records        (written by scripts/make_glue.py: restore_interior's lines up to
 *                                                 "linear interpolation in time", nothing else changed)
 * mode_internal_ calls it at the place of advance.f:452 and pushes trstrb/f, srstrb/f, taurstrb/f on the steps the
 * routine re-reads them (`iint.eq.2 .or. mod(iint,irst).eq.0`, :1038,1053).  Nudging is then ON, as in the reference;
 * it is also switched on by pushing a restoring record by hand (pomgpu_f_push_(trstrb) ...).  pomgpu_f_set_restore_
 * overrides: 0 = off, 1 = on, -1 = the automatic rule. */
void pomgpu_f_set_restore_(const int* on);
void pomgpu_f_finalize_(void);
/* C-side access for tests: address of a COMMON member by name (NULL if unknown / unbound) */
void* pomgpu_f_member(const char* name, long* elems);
char pomgpu_f_member_type(const char* name);   /* 'd', 'i', 'l', 'c' or 0 */
const char* pomgpu_f_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
