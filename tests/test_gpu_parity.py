"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs; plus size-independent properties at BASELINE.json's full size."""
import numpy as np
import pytest

from extpom_b200 import synthetic as syn
from tests import parity_cases as pc
from tests.common import RTOL, digest

pytestmark = pytest.mark.gpu


def _factory(im, jm, kb):
    from extpom_b200.pomgpu import PomGpu
    return PomGpu(im, jm, kb)


@pytest.mark.parametrize("case", pc.STEP_CASES_GPU, ids=pc.case_id)
def test_steps_match_oracle(case):
    pc.check_steps(_factory, case)


def test_stage_by_stage_is_bitwise_equal():
    pc.check_stages(_factory, (26, 21, 10), nstep=3)


@pytest.mark.parametrize("routine", pc.ROUTINES)
def test_routine_matches_oracle(routine):
    pc.check_routine(_factory, routine, (28, 22, 10))


@pytest.mark.parametrize("name", pc.REF_GOLDEN_GPU)
def test_cuda_path_matches_the_references_own_output(name):
    """The CUDA path against the fields the reference's own Fortran source produced (tests/golden/ref_*.npz,
    scripts/make_ref_golden.py + oracle/f77ref.py); 1e-11: the only non-identical operation is |S|**1.5."""
    pc.check_ref_golden(_factory, name, tol=RTOL)


def test_cuda_path_matches_a_live_run_of_the_reference_source():
    """The reference's own Fortran source (oracle/f77ref.py; on the GPU box the translation that
    __graft_entry__.build() left in oracle/_ref) run NOW next to the CUDA path, three steps on a small island case."""
    from oracle.f77ref import F77Ref, Reference
    from scripts.make_ref_golden import ref_restore_setup
    from tests.common import F2, F3, KB_SCRATCH, rel_err
    dims, kw = (13, 11, 6), {"island": True}
    if not Reference.available(*dims):
        pytest.skip("neither the reference source nor its translation (oracle/_ref) is on this machine")
    _, r = syn.seamount(*dims, F77Ref, **kw)
    st = syn.make_state(*dims, **kw)
    g = _factory(*dims)
    g.load(st); syn.finish_init(st, g); ref_restore_setup(g, st)
    for i in (1, 2, 3):
        r.step(i); g.step(i)
    for n in F3 + F2:
        a, b = r.get(n), g.get(n)
        if n in KB_SCRATCH:
            a, b = a[:, :, :-1], b[:, :, :-1]
        assert rel_err(a, b) <= RTOL, (n, rel_err(a, b))


@pytest.mark.parametrize("name", pc.GOLDEN)
def test_matches_golden(name):
    pc.check_golden(_factory, name)


def test_quarter_of_reference_default_grid_40_steps():
    """BASELINE configs[0] is the reference's own 282x306x40 grid (pom.h_dist:22-28); a
    141x153x40 quarter of it keeps the oracle within CI time.  After 40 internal steps:
    (a) bitwise equal when |S|**1.5 uses the same routine on both sides,
    (b) <= 1e-8 (max-abs / field max) against libm pow -- 1-ulp density differences amplified
        by the turbulence closure (q2l is the most sensitive field, ~1.5e-9 measured)."""
    case = ((141, 153, 40), 40, {})
    assert pc.check_steps(_factory, case, pow_mode=1, tol=0.0) == 0.0
    worst = pc.check_steps(_factory, case, pow_mode=0, tol=1e-8)
    print("worst rel max-abs error after 40 steps vs libm-pow oracle:", worst)


def test_full_reference_default_grid_282x306x40_10_steps():
    """BASELINE configs[0]: the reference's own grid (pom.h_dist:22-28: im_global=282, jm_global=306,
    kb=40) as ONE sub-domain with the default namelist (mode=3, nadv=2, nitera=1, npg=1, isplit=30),
    10 internal steps: bitwise when |S|**1.5 uses the same routine on both sides; against the libm-pow
    oracle <= 1e-9 (max-abs / field max) on every compared field -- the 1-ulp density differences are
    amplified most by the turbulence closure (measured on B200: km, kh, kq 1.5e-10 max-abs, 8e-12
    relative L2; el, u, v, t, s, q2, q2l <= 1.2e-12 relative L2)."""
    case = ((282, 306, 40), 10, {})
    assert pc.check_steps(_factory, case, pow_mode=1, tol=0.0) == 0.0
    worst = pc.check_steps(_factory, case, pow_mode=0, tol=1e-9)
    print("282x306x40, 10 steps: worst rel max-abs error vs libm-pow oracle:", worst)


def test_kb61_128x128_columns_match_oracle():
    """BASELINE configs[4] has kb=61: the column solvers (profq, proft, profu/v) with 61 levels on
    128x128 columns, 3 internal steps (one cold-start step + two full ones)."""
    pc.check_steps(_factory, ((128, 128, 61), 3, {"island": True}))


def test_config1_1024x1024x41_matches_oracle():
    """BASELINE configs[1] at FULL size against the oracle: seamount 1024x1024x41, isplit=30, three
    internal steps (the first is the cold-start step that skips the 3-D block, advance.f:362), every
    prognostic and diagnostic field of tests/common.py within RTOL.  One generated state feeds both."""
    import gc
    from oracle.pomo import Oracle
    from tests.common import assert_close
    im = jm = 1024; kb = 41
    st = syn.make_state(im, jm, kb)
    o = Oracle(im, jm, kb)
    o.load(st)
    g = _factory(im, jm, kb)
    g.load(st)
    syn.finish_init(st, o)
    syn.finish_init(st, g)
    del st
    gc.collect()
    for i in (1, 2, 3):
        o.step(i)
        g.step(i)
    worst = assert_close(o, g, tol=RTOL)
    vo, vg = o.check_velocity(), g.check_velocity()
    assert abs(vo - vg) <= RTOL * max(1.0, abs(vo))
    print("1024x1024x41, 3 steps: worst rel max-abs error vs oracle:", worst)
    o.close(); g.close()


def test_two_nccl_ranks_equal_single_domain_bitwise():
    """SURVEY 8(e) acceptance through the NCCL transport: two processes, one strip per GPU, every
    rank compares its rows bitwise with a single-domain run (scripts/strip_check.py).  Needs two
    GPUs; bench.py repeats the same check before every multi-GPU timing (`nrank_parity`)."""
    import os, socket, subprocess, sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(root, "scripts", "strip_check.py"), "96", "160", "16", "6", "4"],
                       capture_output=True, text=True, cwd=root, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("BITWISE EQUAL") == 2, r.stdout


def test_full_size_properties():
    """1024x1024x41 (BASELINE configs[1]): (a) two independent runs are bitwise identical,
    (b) fields stay finite and below the blow-up bound, (c) land stays masked,
    (d) a strip of the result equals the oracle run on ... is covered at smaller sizes."""
    im = jm = 1024; kb = 41
    dig = []
    for rep in range(2):
        st, g = syn.seamount(im, jm, kb, _factory)
        for i in range(1, 6):
            g.step(i)
        v = g.check_velocity()
        assert np.isfinite(v) and v < 1.0
        fields = {n: g.get(n) for n in ("u", "t", "q2", "el")}
        dig.append({n: digest(a) for n, a in fields.items()})
        if rep == 0:
            fsm = st["fields"]["fsm"]
            for n, a in fields.items():
                assert np.isfinite(a).all(), n
            assert np.abs(fields["t"][:, :, :kb - 1][fsm == 0]).max() == 0.0
            assert np.abs(fields["el"][fsm == 0]).max() == 0.0
        del g, st
    assert dig[0] == dig[1]


def test_rest_state_at_scale():
    """Uniform T,S, no wind, no inflow on 512x512x41: stays at rest (|u| < 1e-10)."""
    im = jm = 512; kb = 41
    st = syn.make_state(im, jm, kb, wind=False, noise=False)
    f = st["fields"]
    for n in ("tb", "t", "tclim"): f[n][...] = 10.0
    for n in ("ub", "u", "uab", "ua", "uabe", "uabw"): f[n][...] = 0.0
    f["tsurf"][...] = 10.0
    for n in ("tbe", "tbw", "tbn", "tbs"): f[n][:, :kb - 1] = 10.0
    g = _factory(im, jm, kb)
    g.load(st)
    syn.finish_init(st, g)
    for i in range(1, 6):
        g.step(i)
    for n in ("u", "v", "ua", "el", "w"):
        assert np.abs(g.get(n)).max() <= 1e-10, n


@pytest.mark.parametrize("nstrips,ghost", [(2, 4), (3, 2)])
def test_strips_on_one_gpu_equal_single_domain_bitwise(nstrips, ghost):
    """The j-strip decomposition (csrc/pom_halo.cu) with all strips on this GPU: pack kernels,
    device-to-device halo copies and the validity bookkeeping; must reproduce the single-domain
    run bitwise (the NCCL transport between processes is checked by scripts/strip_check.py)."""
    from extpom_b200 import strips as sp
    from extpom_b200.pomgpu import PomGpu, PomGroup
    from tests.common import F2, F3
    dims, nstep = (64, 90, 12), 5
    fac = lambda a, b, c, strip=None, ghost=0: PomGpu(a, b, c, strip=strip, ghost=ghost)
    _, whole = syn.seamount(*dims, lambda a, b, c: PomGpu(a, b, c), island=True)
    strips = [sp.make_strip(*dims, own, ghost, fac, island=True)[1] for own in sp.partition(dims[1], nstrips)]
    grp = PomGroup(strips)
    sp.finish_init_group(None, grp)
    for i in range(1, nstep + 1):
        whole.step(i)
        grp.step(i)
    for n in F3 + F2:
        a, b = whole.get(n), grp.gather(n)
        if n in ("t", "tb", "s", "sb"):
            a, b = a[:, :, :-1], b[:, :, :-1]
        assert np.array_equal(a, b), n
    assert whole.check_velocity() == grp.check_velocity()


def test_async_forcing_push_equals_blocking_push():
    """pomgpu_push_async (shadow buffer + copy stream, swapped in by the next step) must give the
    same fields as the blocking push; the forcing changes every step like surface_forcing does
    (bounds_forcing.f:908-909)."""
    from extpom_b200.pomgpu import PomGpu
    dims = (96, 80, 16)
    digs = []
    for mode in ("sync", "async"):
        st, g = syn.seamount(*dims, lambda a, b, c: PomGpu(a, b, c))
        w0 = st["fields"]["wusurf"].copy(order="F")
        pin = g.pinned("wusurf")
        vl = []
        for i in range(1, 7):
            pin[...] = w0 * (1.0 + 0.1 * i)
            if mode == "sync":
                g.put("wusurf", pin)
            else:
                g.put_async("wusurf", pin)
                g.sync() if i == 3 else None      # the host buffer may be rewritten only after the copy
            g.step(i)
            vl.append(g.check_velocity_lagged() if mode == "async" else g.check_velocity())
            if mode == "async":
                g.sync()                           # pin is rewritten next iteration
        assert np.array_equal(g.get("wusurf"), w0 * 1.6)
        digs.append({n: digest(g.get(n)) for n in ("u", "v", "t", "q2", "el", "ua", "wubot")})
        if mode == "async":
            assert vl[0] == 0.0 and vl[1:] == sync_v[:-1]
        else:
            sync_v = vl
    assert digs[0] == digs[1]


def test_restore_interior_matches_oracle():
    pc.check_restore(_factory)


def test_domain_stats_matches_oracle():
    pc.check_domain_stats(_factory)


def test_forcing_interpolation_matches_oracle():
    pc.check_forcing_interp(_factory)


@pytest.mark.parametrize("emax", [30, 100, 140])
def test_division_helpers_are_ieee_exact(emax):
    """pdiv (nvcc's fp64 division sequence without the range test) and RDiv (hoisted reciprocal)
    against `/` on the device: 2^26 pseudo-random operand pairs per magnitude range, every bit."""
    g = _factory(8, 8, 4)
    bad = g.L.pomgpu_selftest_pdiv(g.h, 1 << 26, 12345 + emax, emax)
    assert bad == 0, bad


def test_direct_load_fallback_equals_tma_path(monkeypatch):
    """POMGPU_NO_TMA=1 (and every odd `im`) runs the same functors on direct global loads instead of
    TMA-staged tiles: the two device paths must agree bitwise."""
    dims, nstep = (48, 40, 12), 5
    _, a = syn.seamount(*dims, _factory, island=True)
    monkeypatch.setenv("POMGPU_NO_TMA", "1")
    _, b = syn.seamount(*dims, _factory, island=True)
    monkeypatch.delenv("POMGPU_NO_TMA")
    for i in range(1, nstep + 1):
        a.step(i); b.step(i)
    for n in pc.F3 + pc.F2:
        assert np.array_equal(a.get(n), b.get(n)), n


@pytest.mark.parametrize("dims", [(320, 296, 41), (130, 120, 61), (282, 306, 40)], ids=lambda d: "x".join(map(str, d)))
def test_persistent_tma_kernels_equal_direct_load_path_at_scale(dims, monkeypatch):
    """The persistent parked uv_filter (pom_tma.h: tmaparkkernel; several tiles per block, ring stages re-used,
    32x6 tiles at kb=41 / 32x4 at kb=61, a partial last TMA box at kb=40) and the other TMA-staged kernels against
    the direct-load path of the same functors (POMGPU_NO_TMA=1: two-sweep uv_filter), every field, every bit."""
    nstep = 3
    _, a = syn.seamount(*dims, _factory, island=True)
    monkeypatch.setenv("POMGPU_NO_TMA", "1")
    _, b = syn.seamount(*dims, _factory, island=True)
    monkeypatch.delenv("POMGPU_NO_TMA")
    for i in range(1, nstep + 1):
        a.step(i); b.step(i)
    for n in pc.F3 + pc.F2:
        assert np.array_equal(a.get(n), b.get(n)), n


def test_push_of_u_v_between_steps():
    pc.check_push_midrun(_factory)


def test_host_callback_transport_on_the_device_build():
    """pomgpu_group_set_transport with the CUDA library: the callback gets HOST buffers (pinned
    mirrors of the device staging buffers).  Two strips, each in its own group and stepped by its
    own thread, swap their packed halo rows through Python queues; the result must equal the
    single-domain run bitwise.  A failing transport must stop the step and set error_status."""
    import queue, threading
    from extpom_b200 import strips as sp
    from extpom_b200.pomgpu import PomGpu, PomGroup, PomGpuError
    from tests.common import F2, F3
    dims, nstep, ghost = (48, 60, 10), 4, 3
    fac = lambda a, b, c, strip=None, ghost=0: PomGpu(a, b, c, strip=strip, ghost=ghost)
    _, whole = syn.seamount(*dims, lambda a, b, c: PomGpu(a, b, c), island=True)
    owns = sp.partition(dims[1], 2)
    strips = [sp.make_strip(*dims, own, ghost, fac, island=True)[1] for own in owns]
    groups = [PomGroup([s]) for s in strips]
    box = [queue.Queue(), queue.Queue()]           # box[r]: rows on their way TO rank r

    def transport(rank):
        def fn(send_s, recv_s, send_n, recv_n):
            if rank == 0:                           # north neighbour only
                box[1].put(send_n.copy()); recv_n[...] = box[0].get(timeout=60)
            else:
                box[0].put(send_s.copy()); recv_s[...] = box[1].get(timeout=60)
        return fn

    errs = []

    def run(rank):
        try:
            grp = groups[rank]
            grp.set_transport(transport(rank))
            sp.finish_init_group(None, grp)
            for i in range(1, nstep + 1):
                grp.step(i)
            strips[rank].sync()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    for t in th: t.start()
    for t in th: t.join()
    assert not errs, errs
    for i in range(1, nstep + 1):
        whole.step(i)
    for n in F3 + F2:
        a = whole.get(n)
        b = np.concatenate([g.gather(n) for g in groups], axis=1)
        if n in ("t", "tb", "s", "sb"):
            a, b = a[:, :, :-1], b[:, :, :-1]
        assert np.array_equal(a, b), n

    def broken(*_):
        raise RuntimeError("link down")
    groups[0].set_transport(broken)
    with pytest.raises(PomGpuError):
        for i in range(nstep + 1, nstep + 12):
            groups[0].step(i)
    assert strips[0].getc("error_status") == 1


def test_strips_on_two_devices_in_one_process():
    """A group whose strips sit on DIFFERENT devices of one process: each strip launches on its own
    stream, the seam copies are peer copies ordered by events (pack -> copy -> next pack), the
    shared-memory opt-in of every kernel is granted per device.  Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from extpom_b200 import strips as sp
    from extpom_b200.pomgpu import PomGpu, PomGroup
    from tests.common import F2, F3
    dims, nstep, ghost = (96, 120, 16), 5, 4
    _, whole = syn.seamount(*dims, lambda a, b, c: PomGpu(a, b, c), island=True)
    owns = sp.partition(dims[1], 2)
    strips = []
    for dev, own in enumerate(owns):
        fac = lambda a, b, c, strip=None, ghost=0, d=dev: PomGpu(a, b, c, device=d, strip=strip, ghost=ghost)
        strips.append(sp.make_strip(*dims, own, ghost, fac, island=True)[1])
    grp = PomGroup(strips)
    sp.finish_init_group(None, grp)
    for i in range(1, nstep + 1):
        whole.step(i)
        grp.step(i)
    for n in F3 + F2:
        a, b = whole.get(n), grp.gather(n)
        if n in ("t", "tb", "s", "sb"):
            a, b = a[:, :, :-1], b[:, :, :-1]
        assert np.array_equal(a, b), n
