/* pomgpu.h -- C ABI of libpomgpu, the B200-native drop-in for extPOM's
 * time-stepping hot path (pom/advance.f:21-32 and the pom/solver.f kernels it
 * drives).  Plain pointers and sizes only; no CUDA or torch types.
 *
 * The reference passes all model state through Fortran COMMON blocks with
 * compile-time extents (pom.h_dist:22-28,291-364,410-450,532-608) and calls
 * argument-less external subroutines (advance.f:21-32).  This ABI keeps the
 * same names and the same array layout (fp64, column-major, i fastest,
 * (im_local, jm_local[, kb])), but with run-time extents and an explicit
 * context that owns the HBM-resident copies.  Fields are addressed by their
 * COMMON member name ("u", "tb", "wusurf", "tbe", ...).
 *
 * Return codes: 0 = ok, 1 = CUDA failure (also sets blkcon's error_status=1,
 * the reference's error convention, advance.f:118,556-563), 2 = bad argument.
 * The library never calls exit() and has NO CPU fallback: pomgpu_create fails
 * when no CUDA device is present.
 */
#ifndef POMGPU_H
#define POMGPU_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pomgpu pomgpu_t;
typedef struct pomgpu_group pomgpu_group_t;

/* ---- lifecycle ------------------------------------------------------------
 * Replaces distribute_mpi (pom/parallel_mpi.f:34-122).  pomgpu_create holds the
 * whole (im, jm, kb) domain on one GPU.  pomgpu_create_strip holds global rows
 * j_first..j_last (1-based, inclusive) of a (im, jm_global, kb) domain plus
 * `ghost` rows on each interior seam; host arrays passed to push/pull then
 * have jm_local = pomgpu_local_rows() rows starting at pomgpu_row_offset(). */
pomgpu_t* pomgpu_create(int im, int jm, int kb, int device);
pomgpu_t* pomgpu_create_strip(int im, int jm_global, int kb, int j_first,
                              int j_last, int ghost, int device);
void pomgpu_destroy(pomgpu_t* ctx);
int pomgpu_local_rows(const pomgpu_t* ctx);
int pomgpu_row_offset(const pomgpu_t* ctx); /* global 0-based row of local row 0 */
const char* pomgpu_last_error(const pomgpu_t* ctx);

/* ---- blkcon scalars (pom.h_dist:69-198; set by read_input, pom/initialize.f:67-191)
 * by name: "dti2", "smoth", "isplit", "nadv", ... */
int pomgpu_set_const(pomgpu_t* ctx, const char* name, double value);
int pomgpu_get_const(pomgpu_t* ctx, const char* name, double* value);

/* ---- host <-> HBM, by COMMON member name.  Synchronous w.r.t. the host buffer.
 * push: driver-owned host array -> device (init, per-step forcing of
 *       pom/bounds_forcing.f:844-865,908-909,954-955,978).
 * pull: device -> host (output/restart steps, pom/advance.f:35-49). */
int pomgpu_push(pomgpu_t* ctx, const char* name, const double* host);
int pomgpu_pull(pomgpu_t* ctx, const char* name, double* host);
/* local rows [row0, row0+nrows) of a field from a host array with nrows rows (band-wise init of
 * large strips); fields without a j dimension are pushed whole */
int pomgpu_push_rows(pomgpu_t* ctx, const char* name, const double* host, int row0, int nrows);
long pomgpu_field_elems(pomgpu_t* ctx, const char* name); /* 0 if unknown */
/* The same for a host array with GLOBAL extents (im, jm_global[, kb]; (jm_global[, kb]) for the west / east edge
 * arrays) -- a COMMON member of a single-rank driver whose domain is spread over several strips (libpomgpu_f with more
 * than one device).  push takes the rows this strip HOLDS (owned + ghost rows), pull returns the rows it OWNS, so
 * the strips of a group assemble the array between them; arrays without a j dimension are pushed / pulled whole. */
long pomgpu_field_global_elems(pomgpu_t* ctx, const char* name);
int pomgpu_push_global(pomgpu_t* ctx, const char* name, const double* host_global);
int pomgpu_pull_global(pomgpu_t* ctx, const char* name, double* host_global);
/* enqueue-only push (no host wait) for the per-step forcing; the host buffer should be
 * page-locked (pomgpu_pin_host registers a driver-owned array, e.g. a COMMON block).  The copy
 * goes to a shadow buffer on a copy stream, overlapping the step still running; the data
 * becomes the field's content at the next pomgpu_step / pull / push (buffers are swapped). */
int pomgpu_push_async(pomgpu_t* ctx, const char* name, const double* host);
/* Time interpolation of the forcing on the device.  The reference reads a new
 * record every iwind / iheat / ibc internal steps but interpolates `x = fold*xb + fnew*xf`,
 * fold = 1.-fnew, on the host EVERY step (pom/bounds_forcing.f:841-865 lateral_bc, :904-909 wind,
 * :949-957 heat).  Here the driver pushes a record only when the reference reads one:
 *   pomgpu_push_record(ctx,"wusurf",slot,host)  slot 0 = older record (wusurfb), 1 = newer (wusurff);
 *                                               asynchronous when `host` is page-locked;
 *   pomgpu_rotate_record(ctx,"wusurf")          `wusurfb = wusurff` (:888-893) as a pointer swap;
 * and calls, every step, with fnew computed exactly as the reference does (time/twind-ntime):
 *   pomgpu_wind (wusurf,wvsurf), pomgpu_heat (wtsurf,swrad), pomgpu_lateral_bc (tbw,sbw,ubw,tbe,sbe,ube,
 *   tbn,sbn,vbn,tbs,sbs,vbs and the depth integrals uabw,uabe,vabn,vabs), or pomgpu_interp for one field.
 * Fields are shaped like pomgpu_push's (local rows of a strip).  Returns 2 if a record is missing. */
int pomgpu_push_record(pomgpu_t* ctx, const char* name, int slot, const double* host);
int pomgpu_rotate_record(pomgpu_t* ctx, const char* name);
int pomgpu_interp(pomgpu_t* ctx, const char* name, double fnew);
int pomgpu_wind(pomgpu_t* ctx, double fnew);
int pomgpu_heat(pomgpu_t* ctx, double fnew);
int pomgpu_lateral_bc(pomgpu_t* ctx, double fnew);
/* Self test of the two division helpers the kernels use to stay bit-identical to the reference's
 * IEEE divisions (pdiv: nvcc's own fp64 division sequence without its operand-range test; RDiv:
 * hoisted reciprocal + fma residual): number of operand pairs, out of n pseudo-random ones with
 * magnitudes 10^-emax..10^emax, whose quotient differs in any bit from `a/b` on the device. */
long pomgpu_selftest_pdiv(pomgpu_t* ctx, long n, unsigned long seed, int emax);
int pomgpu_pin_host(void* ptr, unsigned long bytes);
int pomgpu_unpin_host(void* ptr);

/* ---- resident-mode time stepping -------------------------------------------
 * pomgpu_step = pom/advance.f:21-32 (lateral_viscosity, mode_interaction,
 * isplit x mode_external, mode_internal) for internal step `iint`; `time` and
 * `ramp` are what get_time (advance.f:62-75) computed.  Asynchronous: returns
 * after enqueueing; pomgpu_sync / pull / check_velocity wait. */
int pomgpu_step(pomgpu_t* ctx, int iint, double time, double ramp);
int pomgpu_sync(pomgpu_t* ctx);
/* check_velocity (advance.f:611-641) as a device max-reduction: returns
 * max|vaf| and sets error_status when it exceeds vmaxl. */
double pomgpu_check_velocity(pomgpu_t* ctx);
/* the same without waiting for the step just enqueued: starts the reduction and returns the
 * value of the previous call (one step of lag; 0 on the first call) */
double pomgpu_check_velocity_lagged(pomgpu_t* ctx);
/* the same lagged one-scalar read-back for any field: max |x| over the rows this strip holds
 * (-1 for an unknown field); does not touch error_status */
double pomgpu_field_absmax_lagged(pomgpu_t* ctx, const char* name);
/* domain_stats (advance.f:644-755) as a device reduction: 7 partial sums for each OWNED row
 * (rows[owned_rows][7] = atot, sum(et*darea), vtot, mtot, sum(tb*dvol), sum(sb*dvol), ekin);
 * adding the rows in global order and dividing (eavg/atot, tavg/vtot, savg=stot/vtot) gives
 * the reference's eight numbers independently of the decomposition. */
int pomgpu_domain_stats_rows(pomgpu_t* ctx, double* rows);
long pomgpu_launch_count(pomgpu_t* ctx, int reset);
/* per-kernel device time from CUDA events recorded on the launch stream around every
 * kernel between begin and end; end writes a JSON array
 * [{"name","launches","ms","bytes"}] where bytes = algorithmic bytes */
int pomgpu_event_record(pomgpu_t* ctx, int slot);            /* slot 0..7, on the launch stream */
double pomgpu_event_elapsed_ms(pomgpu_t* ctx, int a, int b);  /* waits for event b */
int pomgpu_profile_begin(pomgpu_t* ctx);
int pomgpu_profile_end(pomgpu_t* ctx, char* json, int len);

/* ---- multi-GPU: j-strips with redundant ghost rows -----------------------------
 * Replaces exchange2d_mpi / exchange3d_mpi (pom/parallel_mpi.f:154-351).  A group is
 * the ordered (south -> north) list of strips held by THIS process: all strips of the
 * domain in one process (tests), or one strip per process/GPU connected over NCCL
 * (pomgpu_nccl_unique_id on rank 0, broadcast by the driver's MPI, then
 * pomgpu_group_connect_nccl everywhere).  pomgpu_group_step = pomgpu_step on the
 * group; halo exchanges are issued inside, only when a kernel would read a stale row.
 * An N-strip run is bitwise equal to the single-domain run. */
pomgpu_group_t* pomgpu_group_create(int n, pomgpu_t** strips);
void pomgpu_group_destroy(pomgpu_group_t* g);
int pomgpu_nccl_unique_id(void* out128);
int pomgpu_group_connect_nccl(pomgpu_group_t* g, const void* id128, int rank, int world);
/* host transport (gloo in the CPU tests; MPI_Sendrecv in a Fortran/MPI driver without NCCL):
 * called with the packed south / north send buffers and the receive buffers to fill; n_* =
 * doubles per direction (0 = no neighbour on that side).  All four pointers are HOST memory:
 * page-locked mirrors that the library copies from / to its device staging buffers around the
 * call.  A non-zero return marks the group failed: error_status=1 on every strip and no further
 * kernel is launched (the ghost rows would be stale). */
typedef int (*pomgpu_halo_cb)(void* user, const double* send_s, double* recv_s, long n_s,
                              const double* send_n, double* recv_n, long n_n);
void pomgpu_group_set_transport(pomgpu_group_t* g, pomgpu_halo_cb cb, void* user);
int pomgpu_group_step(pomgpu_group_t* g, int iint, double time, double ramp);
/* the two solver.f routines `initialize` calls (initialize.f:416,425,502), on a group */
int pomgpu_group_dens(pomgpu_group_t* g, const char* si, const char* ti, const char* rhoo);
int pomgpu_group_baropg(pomgpu_group_t* g);   /* baropg or baropg_mcc by npg (initialize.f:502-505) */
int pomgpu_group_baropg_kind(pomgpu_group_t* g, int npg);   /* 1: baropg (solver.f:848), 2: baropg_mcc (:943) */
/* the four step routines of advance.f:21-32 one by one on a group, for a driver that keeps the reference's own
 * `advance` (libpomgpu_f over several strips); pomgpu_group_error_status = OR of the strips' error_status */
int pomgpu_group_lateral_viscosity(pomgpu_group_t* g);            /* advance.f:96  */
int pomgpu_group_mode_interaction(pomgpu_group_t* g);             /* advance.f:144 */
int pomgpu_group_mode_external(pomgpu_group_t* g, int iext);      /* advance.f:205 */
int pomgpu_group_mode_internal(pomgpu_group_t* g, int iint);      /* advance.f:356 */
int pomgpu_group_error_status(pomgpu_group_t* g);
double pomgpu_group_check_velocity(pomgpu_group_t* g); /* max over this process's strips */
long pomgpu_group_exchanges(pomgpu_group_t* g, long* fields, int reset); /* halo messages so far */
/* how the seam rows travel: 0 device copies between strips of this process, 1 ncclSend/ncclRecv,
   2 peer copies into the neighbour's CUDA-IPC-mapped staging buffers, 3 host callback */
int pomgpu_group_transport(pomgpu_group_t* g);
/* developer tool (POMGPU_HALO_TRACE=1): one text line per exchange since the last call -- bytes per
 * direction and the device time of its pack, transfer and unpack from CUDA events; returns the length */
int pomgpu_group_halo_trace(pomgpu_group_t* g, char* buf, int len);

/* ---- the reference's subroutines on the resident state (same names) ---------
 * advance.f */
int pomgpu_lateral_viscosity(pomgpu_t* ctx);       /* advance.f:96  */
int pomgpu_mode_interaction(pomgpu_t* ctx);        /* advance.f:144 */
int pomgpu_mode_external(pomgpu_t* ctx, int iext); /* advance.f:205 */
int pomgpu_mode_internal(pomgpu_t* ctx, int iint); /* advance.f:356 */
/* one block of mode_internal (test hook): 0 u/v adjust :365-393, 1 vertvl+bcondorl(5), 2 advq x2,
 * 3 profq, 4 bcond(6)+filter, 5/6 advt T/S, 7/8 proft T/S, 9 bcond(4)+filter+restore, 10 dens,
 * 11 advu, 12 advv, 13 profu, 14 profv, 15 bcondorl(3)+filter, 16 2-D rotations :525-531, 17 realvertvl */
int pomgpu_internal_stage(pomgpu_t* ctx, int iint, int stage);
/* solver.f; array arguments are COMMON member names (the reference passes
 * q2b,q2,uf / tb,t,tclim,uf / s,t,rho / uf,wtsurf,tsurf,nbct, advance.f:407-454) */
int pomgpu_advave(pomgpu_t* ctx);                  /* solver.f:6    */
int pomgpu_advct(pomgpu_t* ctx);                   /* solver.f:201  */
int pomgpu_advq(pomgpu_t* ctx);                    /* solver.f:411, both q2->uf and q2l->vf */
int pomgpu_advt1(pomgpu_t* ctx, const char* fb, const char* f, const char* fclim, const char* ff); /* :480 */
int pomgpu_advt2(pomgpu_t* ctx, const char* fb, const char* f, const char* fclim, const char* ff); /* :577 */
int pomgpu_advu(pomgpu_t* ctx);                    /* solver.f:734  */
int pomgpu_advv(pomgpu_t* ctx);                    /* solver.f:791  */
int pomgpu_baropg(pomgpu_t* ctx);                  /* solver.f:848  */
int pomgpu_baropg_mcc(pomgpu_t* ctx);              /* solver.f:943 (npg=2) */
int pomgpu_dens(pomgpu_t* ctx, const char* si, const char* ti, const char* rhoo); /* solver.f:1162 */
int pomgpu_profq(pomgpu_t* ctx);                   /* solver.f:1212 */
int pomgpu_proft(pomgpu_t* ctx, const char* f, const char* wfsurf, const char* fsurf, int nbc); /* :1541 */
int pomgpu_profu(pomgpu_t* ctx);                   /* solver.f:1686 */
int pomgpu_profv(pomgpu_t* ctx);                   /* solver.f:1783 */
int pomgpu_vertvl(pomgpu_t* ctx);                  /* solver.f:1970 (+ bcondorl(5)) */
int pomgpu_realvertvl(pomgpu_t* ctx);              /* solver.f:2024 */
/* advq(qb,q,qf) for ONE quantity (the reference's signature, solver.f:411); pomgpu_advq above does
 * advq(q2b,q2,uf) and advq(q2lb,q2l,vf) in one pass, which is what the step runs */
int pomgpu_advq_fields(pomgpu_t* ctx, const char* qb, const char* q, const char* qf);
/* smol_adif(xmassflux,ymassflux,zwflux,ff) (solver.f:1880): ff*fsm, then the anti-diffusive mass
 * fluxes in place; arguments name 3-D fields (COMMON members or the scratch fields "s3a".."s3e") */
int pomgpu_smol_adif(pomgpu_t* ctx, const char* xmassflux, const char* ymassflux, const char* zwflux,
                     const char* ff);
/* bounds_forcing.f:6 bcond(idx) for idx = 1 (elf), 2 (uaf,vaf), 4 (T,S in uf,vf), 5 (w), 6 (q2,q2l in
 * uf,vf) and :331 bcondorl(idx) for idx = 3 (uf,vf Orlanski), 5 (w) as stand-alone calls: the
 * branches advance.f calls (:231,290,398,414,442,464).  Other idx return 2.  pomgpu_step runs the
 * same point functions fused into its kernels and never calls these. */
int pomgpu_bcond(pomgpu_t* ctx, int idx);
int pomgpu_bcondorl(pomgpu_t* ctx, int idx);

#ifdef __cplusplus
}
#endif
#endif
