"""j-strip domain decomposition across the GPUs of one box (replaces distribute_mpi,
pom/parallel_mpi.f:34-122, and the exchange2d/3d_mpi halo swaps, :154-351).

Memory is i-contiguous, so the domain is cut in j only: rank r owns a block of global
rows and holds `ghost` extra rows on each interior seam (csrc/pom_halo.cu).  One process
per GPU; the strips of the box are connected over NCCL (send/recv across NVLink), the
unique id being broadcast through torch.distributed -- the role MPI_COMM_WORLD plays in
the reference (parallel_mpi.f:6-31).
"""
import os

import numpy as np

from . import synthetic as syn
from .pomgpu import PomGpu, PomGroup

GHOST = int(os.environ.get("POMGPU_GHOST", "8"))   # rows of redundant computation per seam: one batched exchange every ~4 external substeps


def partition(jm, world):
    """Owned global rows (1-based, inclusive) of each rank; uneven strips allowed
    (the reference's uniform-block rule, parallel_mpi.f:54-65, cannot split 4094 rows 8 ways)."""
    base, rem = divmod(jm, world)
    out, j = [], 1
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((j, j + n - 1))
        j += n
    return out


def held_rows(jm, own, ghost):
    """Global rows a strip holds in memory: owned rows + ghost rows on interior seams."""
    lo = own[0] - (ghost if own[0] > 1 else 0)
    hi = own[1] + (ghost if own[1] < jm else 0)
    return max(lo, 1), min(hi, jm)


def make_strip(im, jm_global, kb, own, ghost, factory, **kw):
    """Generate the synthetic state of one strip (only its rows are ever materialised) and load
    it into a solver strip created by factory(im, jm_global, kb, strip=own, ghost=ghost)."""
    whole = own == (1, jm_global)
    g = factory(im, jm_global, kb, strip=None if whole else own, ghost=0 if whole else ghost)
    lo, hi = held_rows(jm_global, own, 0 if whole else ghost)
    band = kw.pop("band", 0)
    if band and hi - lo + 1 > band:
        # large strips: generate and push `band` rows at a time (host memory stays ~30 band-sized fields)
        st = None
        for b0 in range(lo, hi + 1, band):
            st = syn.make_state(im, jm_global, kb, rows=(b0, min(b0 + band - 1, hi)), **kw)
            g.load_rows(st, b0 - lo)
        for n in ("tbn", "sbn", "tbs", "sbs"):     # rows 1 / jm live in the first / last band only
            if lo == 1 and n[2] == "s":
                g.put(n, syn.make_state(im, jm_global, kb, rows=(1, min(band, hi)), **kw)["fields"][n])
            if hi == jm_global and n[2] == "n":
                g.put(n, st["fields"][n])
        return st, g
    st = syn.make_state(im, jm_global, kb, rows=(lo, hi), **kw)
    assert st["dims"][1] == g.jml, (st["dims"], g.jml)
    g.load(st)
    return st, g


class StripSet:
    """The strips of the synthetic seamount case held by this process, stepped as a group."""

    def __init__(self, strips, group, rank, world, rows):
        self.strips, self.group, self.rank, self.world, self.rows = strips, group, rank, world, rows
        self.gpu = strips[0]

    @classmethod
    def create(cls, im, jm_global, kb, rank=0, world=1, device=0, dist=None, ghost=GHOST, factory=None, **kw):
        kw.setdefault("band", 256 if im * kb >= 1024 * 41 else 0)
        """One strip per process (rank of world).  dist = an initialised torch.distributed module
        (NCCL backend on the GPU box) used once, to broadcast the NCCL unique id."""
        factory = factory or (lambda a, b, c, strip=None, ghost=0: PomGpu(a, b, c, device=device, strip=strip, ghost=ghost))
        own = partition(jm_global, world)[rank]
        ghost = max(2, min(ghost, min(b - a + 1 for a, b in partition(jm_global, world)) - 4))
        st, g = make_strip(im, jm_global, kb, own, ghost, factory, **kw)
        grp = PomGroup([g])
        if world > 1:
            cls._connect(grp, g, rank, world, dist)
        finish_init_group(st, grp)
        del st
        return cls([g], grp, rank, world, own)

    @staticmethod
    def _connect(grp, g, rank, world, dist):
        import torch
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.frombuffer(bytearray(PomGroup.nccl_unique_id(g.L)), dtype=torch.uint8).clone()
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        uid = uid.to(dev)
        dist.broadcast(uid, src=0)
        grp.connect_nccl(bytes(uid.cpu().numpy().tobytes()), rank, world)

    def step(self, iint):
        self.group.step(iint)

    def check_velocity(self):
        return self.group.check_velocity()


def finish_init_group(state, group):
    """synthetic.finish_init on a strip group: rmean/rho through dens, the first baropg
    (initialize.f:416,425,502); the kernels also leave the drx2d/dry2d sums (:510-517)."""
    group.dens("sclim", "tclim", "rmean")
    group.dens("sb", "tb", "rho")
    group.baropg()


def gloo_transport(dist, rank, world):
    """Halo transport over a torch.distributed CPU backend (tests): exchange the packed rows
    with the south (rank-1) and north (rank+1) neighbour."""
    import torch

    def fn(send_s, recv_s, send_n, recv_n):
        ops = []
        if send_s is not None:
            ops.append(dist.P2POp(dist.isend, torch.from_numpy(send_s.copy()), rank - 1))
            rs = torch.empty(recv_s.shape[0], dtype=torch.float64)
            ops.append(dist.P2POp(dist.irecv, rs, rank - 1))
        if send_n is not None:
            ops.append(dist.P2POp(dist.isend, torch.from_numpy(send_n.copy()), rank + 1))
            rn = torch.empty(recv_n.shape[0], dtype=torch.float64)
            ops.append(dist.P2POp(dist.irecv, rn, rank + 1))
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        if send_s is not None:
            recv_s[...] = rs.numpy()
        if send_n is not None:
            recv_n[...] = rn.numpy()
    return fn
