// pom_halo.h -- strip group: ghost-row validity tracking + halo exchange (see pom_halo.cu)
#pragma once
#include "pom_core.h"

namespace pom {

constexpr int HALO_MAXF = 64;   // fields per pack kernel launch (every live field fits: one pack / transfer / unpack per exchange)

struct NcclId { char internal[128]; };

// host transport (gloo in the CPU tests, MPI in a Fortran driver): exchange with the south / north
// neighbour process.  The buffers handed to the callback are always HOST memory: the strips'
// staging buffers themselves in the emulation build, page-locked mirrors of the device staging
// buffers (copied D2H before and H2D after the call) in the CUDA build.
typedef int (*halo_cb)(void* user, const double* send_s, double* recv_s, long n_s,
                       const double* send_n, double* recv_n, long n_n);

struct Req { int f; int r; };   // field id, j-radius it is read with

struct Group {
  int n;                 // strips held by this process, south -> north
  Ctx* c[16];
  int ghost;
  bool seams;            // any interior seam at all (false: single domain, no bookkeeping)
  int valid[F_COUNT];    // valid ghost rows per field (identical on every seam by construction)
  double* buf[16][4];    // staging: send south, send north, recv south, recv north
  size_t bufcap[16][4];
  void* nccl;            // ncclComm_t (one process per GPU)
  int rank, world;
  // direct peer transport (one process per GPU, same box): the neighbours' receive buffers and flag
  // words are mapped through CUDA IPC; rows travel as ONE peer copy per direction on the copy engines
  // (no SMs, no host-side blocking) and arrival / consumption are signalled with stream memory ops
  int ipc;               // peer buffers are mapped and in use
  double* peer_rbuf[2];  // [0]: the south neighbour's receive-from-north buffer, [1]: the north neighbour's receive-from-south
  unsigned* flags;       // mine (device): [0] arrived from south, [1] arrived from north, [2] south consumed, [3] north consumed
  unsigned* peer_flags[2];
  unsigned seq;          // exchanges so far
  size_t ipc_cap;        // capacity of the staging buffers (doubles)
  halo_cb cb; void* cb_user;
  double* hbuf[4]; size_t hbufcap[4];   // pinned host mirrors for the callback transport (CUDA build)
  void* ev_pack[16]; void* ev_copy[16];  // cross-stream ordering of in-process seams between devices
  int failed;            // a transport failed: ghost rows are stale, nothing is launched any more
  // overlap of the exchange with interior compute: after the packs (compute stream) the transfer and
  // the unpacks run on the communication stream; the kernel that asked for the rows is launched on
  // the rows that read no ghost row first, then -- after ev_halo -- on the two seam bands
  int overlap;           // transport supports it (NCCL, strips of one device); 0: everything on the compute stream
  int ov_active;         // an exchange is in flight on the communication stream
  int ov_r;              // largest j-radius the kernel about to be launched reads with
  int will[32]; int nwill;   // fields the kernel about to be launched writes (WILL in pom_step.cu)
  long n_exchanges, n_fields_exchanged;
  // POMGPU_HALO_TRACE=1: CUDA events around pack / transfer / unpack of every exchange (developer tool)
  int trace; int ntrace; void* tev[256][4]; long tbytes[256];
  double host_ms[4]; long host_n;   // host time spent enqueueing packs / transport / unpacks (trace)
};

Group* group_create(int n, Ctx** ctxs);
void group_destroy(Group* G);
int group_connect_nccl(Group* G, const void* id128, int rank, int world);
void group_set_callback(Group* G, halo_cb cb, void* user);
int nccl_unique_id(void* out128);
int group_exchange(Group* G, const int* fields, int nf);
int group_need(Group* G, const Req* in, int n);
void group_will(Group* G, const int* out, int n);
void group_produced(Group* G, int e, const int* out, int n);
void group_swap(Group* G, int fa, int fb);
// launch windows of the kernel about to run (pom_step.cu's EACH): 1 part, or interior + two seam bands
int group_parts(Group* G);
bool group_window(Group* G, const Ctx* c, int e, int part, int nparts, int* j0, int* j1);
void group_wait_halo(Group* G);
void group_launched(Group* G);
int group_trace_report(Group* G, char* buf, int n);

}  // namespace pom
