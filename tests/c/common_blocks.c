/* common_blocks.c -- the COMMON blocks of pom.h as a shared object, for tests that play the Fortran driver
 * from Python (tests/fabi.py).
 *
 * Under gfortran a COMMON block /blk3d/ is the global data symbol `blk3d_` (pom.h_dist:46-54,69-78,142-198,
 * 208-212,291-364,410-450,532-608).  A Fortran executable defines them; here they are defined once, large
 * enough for every grid the tests use -- libpomgpu_f computes the member offsets at run time from the extents
 * given to pomgpu_f_set_dims_, so one set of blocks serves any (im_local, jm_local, kb) that fits.  The pages
 * are .bss: untouched ones cost nothing.
 * Build: gcc -shared -fPIC common_blocks.c -o libpom_common.so ; load with RTLD_GLOBAL BEFORE libpomgpu_f.
 */
#define MAXN2 (1L << 14)        /* im_local*jm_local        */
#define MAXN3 (1L << 19)        /* im_local*jm_local*kb     */
#define MAXEDGE (1L << 13)      /* max(im_local,jm_local)*kb */
#define MAXKB 256
int blksiz_[8];
int blkpar_[7 + 2 * 4096];
double blkcon_[64];
double blk1d_[4 * MAXKB];
double blk2d_[73 * MAXN2];
double blk3d_[40 * MAXN3];
double bdry_[76 * MAXEDGE];
/* `subroutine restore_interior_records` of the driver's glue, played by whatever Python
 * function tests/fabi.py hangs into the slot */
void (*pom_records_hook)(void);
void restore_interior_records_(void) { if (pom_records_hook) pom_records_hook(); }
long pom_common_max_n2(void) { return MAXN2; }
long pom_common_max_n3(void) { return MAXN3; }
long pom_common_max_edge(void) { return MAXEDGE; }
long pom_common_bytes(int which) {
  switch (which) {
    case 0: return sizeof blksiz_;
    case 1: return sizeof blkpar_;
    case 2: return sizeof blkcon_;
    case 3: return sizeof blk1d_;
    case 4: return sizeof blk2d_;
    case 5: return sizeof blk3d_;
    case 6: return sizeof bdry_;
  }
  return 0;
}
