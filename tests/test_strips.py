"""CPU tests of the j-strip decomposition (csrc/pom_halo.cu; replaces distribute_mpi and
exchange2d/3d_mpi, pom/parallel_mpi.f:34-122,154-351) on the host-emulated build.

With all neighbours -1 (n_proc=1) every reference exchange is a no-op, so the single-domain
run is the canonical answer and an N-strip run must reproduce it; here
bitwise, because every kernel computes with global-index semantics."""
import os
import socket
import sys

import numpy as np
import pytest

from extpom_b200 import strips as sp
from extpom_b200 import synthetic as syn
from extpom_b200.pomgpu import PomGpu, PomGroup
from tests import emu
from tests.common import F2, F3


def _factory(im, jm, kb, strip=None, ghost=0):
    return emu.EmuPom(im, jm, kb, strip=strip, ghost=ghost)


def _whole(dims, nstep, **kw):
    st, g = syn.seamount(*dims, lambda a, b, c: _factory(a, b, c), **kw)
    for i in range(1, nstep + 1):
        g.step(i)
    return g


def _group(dims, nstrips, ghost, **kw):
    im, jm, kb = dims
    strips = []
    for own in sp.partition(jm, nstrips):
        st, g = sp.make_strip(im, jm, kb, own, ghost, _factory, **kw)
        strips.append(g)
    grp = PomGroup(strips)
    sp.finish_init_group(None, grp)
    return grp


def _assert_same(whole, grp, names=F3 + F2):
    for n in names:
        a, b = whole.get(n), grp.gather(n)
        assert a.shape == b.shape, (n, a.shape, b.shape)
        if n in ("t", "tb", "s", "sb"):
            a, b = a[:, :, :-1], b[:, :, :-1]
        assert np.array_equal(a, b), (n, float(np.abs(a - b).max()))


def test_partition_covers_uneven_rows():
    rows = sp.partition(4094 + 2, 8)
    assert rows[0][0] == 1 and rows[-1][1] == 4096
    assert all(rows[r + 1][0] == rows[r][1] + 1 for r in range(7))
    assert max(b - a for a, b in rows) - min(b - a for a, b in rows) <= 1


@pytest.mark.parametrize("nstrips,ghost", [(2, 2), (2, 4), (3, 3), (4, 4)])
def test_strips_equal_single_domain_bitwise(nstrips, ghost):
    dims, nstep = (22, 41, 8), 6
    whole = _whole(dims, nstep, island=True)
    grp = _group(dims, nstrips, ghost, island=True)
    for i in range(1, nstep + 1):
        grp.step(i)
    _assert_same(whole, grp)
    assert whole.check_velocity() == grp.check_velocity()
    assert whole.domain_stats() == grp.domain_stats()      # per-row partial sums added in global order
    nex, nfields = grp.exchanges()
    # the reference does 29 3-D + 340 2-D exchanges per step
    print(f"{nstrips} strips ghost {ghost}: {nex / nstep:.1f} batched exchanges/step, {nfields / nstep:.0f} field-rows sets")
    assert 0 < nex / nstep < 369
    assert grp.transport() == "device copies"             # pomgpu_group_transport: strips of one process


@pytest.mark.parametrize("kw", [{"nadv": 1}, {"mode": 4}, {"nbct": 2, "ntp": 3}, {"nitera": 3, "sw": 1.0}, {"mode": 2}, {"npg": 2},
                                {"walls": False, "fluxes": True, "obc": True},      # open north / south sides inside the end strips
                                {"walls": False, "fluxes": True, "obc": True, "npg": 2, "island": True}])
def test_strips_namelist_variants(kw):
    dims, nstep = (20, 30, 7), 4
    whole = _whole(dims, nstep, **kw)
    grp = _group(dims, 2, 3, **kw)
    for i in range(1, nstep + 1):
        grp.step(i)
    _assert_same(whole, grp)


def test_strip_too_thin_is_rejected():
    im, jm, kb = 20, 24, 7
    strips = [sp.make_strip(im, jm, kb, own, 4, _factory)[1] for own in sp.partition(jm, 4)]
    with pytest.raises(Exception):
        PomGroup(strips)          # 6 owned rows < ghost + 4


def test_seam_without_transport_sets_error_status():
    im, jm, kb = 20, 30, 7
    st, g = sp.make_strip(im, jm, kb, (1, 15), 3, _factory)
    grp = PomGroup([g])
    with pytest.raises(Exception):
        for i in range(1, 4):
            grp.step(i)
    assert g.getc("error_status") == 1


def test_failing_transport_stops_the_step_and_keeps_the_error():
    """A transport failure must not be swallowed (ghost rows would be stale): the step raises,
    error_status=1 survives the next CSYNC of the group's scalars, later steps keep failing."""
    im, jm, kb = 20, 30, 7
    st, g = sp.make_strip(im, jm, kb, (1, 15), 3, _factory)
    grp = PomGroup([g])

    def broken(*_):
        raise RuntimeError("link down")
    grp.set_transport(broken)
    assert grp.transport() == "host callback"
    for attempt in range(2):
        with pytest.raises(Exception):
            for i in range(1, 4):
                grp.step(i)
        assert g.getc("error_status") == 1


# ---- two processes over gloo: the host-callback transport ---------------------------------
def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, dims, nstep, ghost, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        im, jm, kb = dims
        own = sp.partition(jm, world)[rank]
        st, g = sp.make_strip(im, jm, kb, own, ghost, _factory, island=True)
        grp = PomGroup([g])
        grp.set_transport(sp.gloo_transport(dist, rank, world))
        sp.finish_init_group(None, grp)
        for i in range(1, nstep + 1):
            grp.step(i)
        np.savez(os.path.join(out, f"rank{rank}.npz"), vamax=grp.check_velocity(),
                 **{n: grp.gather(n) for n in F3 + F2})
    finally:
        dist.destroy_process_group()


def test_two_processes_gloo_equal_single_domain(tmp_path):
    import torch.multiprocessing as mp
    dims, nstep, ghost, world = (20, 36, 7), 4, 3, 2
    mp.spawn(_worker, args=(world, _free_port(), dims, nstep, ghost, str(tmp_path)), nprocs=world, join=True)
    whole = _whole(dims, nstep, island=True)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for n in F3 + F2:
        a = whole.get(n)
        b = np.concatenate([p[n] for p in parts], axis=1)
        if n in ("t", "tb", "s", "sb"):
            a, b = a[:, :, :-1], b[:, :, :-1]
        assert np.array_equal(a, b), n
    assert max(float(p["vamax"]) for p in parts) == whole.check_velocity()


def test_bandwise_loading_equals_whole_loading():
    """pomgpu_push_rows: a strip filled band by band (large grids) holds exactly the same state."""
    dims = (20, 40, 7)
    st, a = sp.make_strip(*dims, (1, 40), 0, _factory, island=True)
    _, b = sp.make_strip(*dims, (1, 40), 0, _factory, island=True, band=7)
    for n in list(F3[:8]) + ["h", "fsm", "dum", "uab", "wusurf", "tsurf", "cbc", "aru"] + ["tbe", "tbw", "tbn", "tbs", "uabe", "vabn", "els", "z"]:
        assert np.array_equal(a.get(n), b.get(n)), n
    for g in (a, b):
        grp = PomGroup([g]); sp.finish_init_group(None, grp)
        for i in range(1, 4):
            grp.step(i)
    for n in F3 + F2:
        assert np.array_equal(a.get(n), b.get(n)), n


def test_forcing_interpolation_on_strips():
    """Device-side time interpolation (bounds_forcing.f:841-865,904-909,949-957) per strip: every strip
    holds the records of its own rows; the result equals the single-domain run bitwise."""
    dims, nstep = (20, 30, 7), 4
    im, jm, kb = dims
    rng = np.random.default_rng(3)
    whole = _whole(dims, 0)
    grp = _group(dims, 2, 3)
    shapes = {"wusurf": (im, jm), "wvsurf": (im, jm), "wtsurf": (im, jm), "swrad": (im, jm)}
    for e in ("w", "e"):
        shapes.update({"tb" + e: (jm, kb), "sb" + e: (jm, kb), "ub" + e: (jm, kb)})
    for e in ("n", "s"):
        shapes.update({"tb" + e: (im, kb), "sb" + e: (im, kb), "vb" + e: (im, kb)})
    base = {"t": 10.0, "s": 35.0, "u": 0.2, "v": 0.0, "w": 0.0}
    rec = [{n: np.asfortranarray(base[n[0]] + (1e-4 if n[0] == "w" else 1e-2) * rng.standard_normal(s))
            for n, s in shapes.items()} for _ in range(2)]
    for g in [whole] + grp.strips:
        for n in shapes:
            g.put_record(n, 0, rec[0][n]); g.put_record(n, 1, rec[1][n])
    for i in range(1, nstep + 1):
        fnew = i / float(nstep + 1)
        for g in [whole] + grp.strips:
            g.wind(fnew); g.heat(fnew); g.lateral_bc(fnew)
        whole.step(i); grp.step(i)
    _assert_same(whole, grp)
    for n in ("wusurf", "wtsurf"):
        assert np.array_equal(whole.get(n), grp.gather(n))
    for s in grp.strips:
        assert np.array_equal(whole.get("uabw")[s.joff:s.joff + s.jml], s.get("uabw"))
        assert np.array_equal(whole.get("vabn"), s.get("vabn"))
