"""j-strip domain decomposition across the GPUs of one box (replaces distribute_mpi,
pom/parallel_mpi.f:34-122, and the exchange2d/3d_mpi halo swaps, :154-351).

Memory is i-contiguous, so the domain is cut in j only: rank r owns a block of global
rows and holds `ghost` extra rows on each interior seam.  One process per GPU.
"""
import numpy as np

from . import synthetic as syn
from .pomgpu import PomGpu


def partition(jm, world):
    """Owned global rows (1-based, inclusive) of each rank; uneven strips allowed."""
    base, rem = divmod(jm, world)
    out, j = [], 1
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((j, j + n - 1))
        j += n
    return out


class StripSet:
    """The strip of the synthetic seamount case held by this rank."""

    def __init__(self, gpu, rank, world, rows):
        self.gpu, self.rank, self.world, self.rows = gpu, rank, world, rows

    @classmethod
    def create(cls, im, jm_global, kb, rank=0, world=1, device=0, dist=None, ghost=4, **kw):
        rows = partition(jm_global, world)[rank]
        if world == 1:
            st, g = syn.seamount(im, jm_global, kb, lambda a, b, c: PomGpu(a, b, c, device=device), **kw)
            del st
            return cls(g, rank, world, rows)
        raise NotImplementedError("multi-GPU strips: halo exchange lands in the next commit")

    def step(self, iint):
        self.gpu.step(iint)
