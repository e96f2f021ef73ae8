"""The gfortran-ABI drop-in (libpomgpu_f.so, include/pomgpu_f.h).

tests/c/fortran_abi_driver.c plays the reference's Fortran driver in C: it DEFINES the COMMON blocks
of pom.h, fills them, makes `initialize`'s solver.f calls and then `advance`'s hot path exactly as
advance.f:21-32 spells it (argument-less calls, everything through COMMON).  The fields it ends with
must be bitwise those of the library driven through its C ABI (pomgpu_step), which the parity tests
pin against the oracle.  No Fortran compiler is involved: symbol names, by-reference arguments and
the COMMON layout generated from pom.h_dist are the whole ABI."""
import os
import re
import subprocess

import numpy as np
import pytest

from extpom_b200 import synthetic as syn
from extpom_b200.pomgpu import LIBPATH, PomGpu
from scripts.dump_state import dump, read_out
from scripts.make_ref_golden import ref_restore_setup
from tests import emu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "pomgpu_f.h")
FLIB = os.path.join(os.path.dirname(LIBPATH), "libpomgpu_f.so")
FLIB_EMU = os.path.join(ROOT, "tests", "_emu", "libpomgpu_f_emu.so")

# what the reference's callers bind (advance.f:21-32,96-537; solver.f; bounds_forcing.f:6,331)
REQUIRED = ("lateral_viscosity_ mode_interaction_ mode_external_ mode_internal_ advave_ advct_ advq_ advt1_ advt2_ "
            "advu_ advv_ baropg_ baropg_mcc_ dens_ profq_ proft_ profu_ profv_ smol_adif_ vertvl_ realvertvl_ bcond_ "
            "bcondorl_ exchange2d_mpi_ exchange3d_mpi_").split()


def _exported(so):
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    return {l.split()[-1] for l in out.splitlines() if l.strip()}


def _declared():
    src = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    return set(re.findall(r"\b([a-z_0-9]+)\s*\(", src)) - {"defined"}


def test_library_exports_every_declared_symbol_and_the_gfortran_names():
    emu.build_emu()
    sym = _exported(FLIB_EMU)
    assert set(REQUIRED) <= sym, sorted(set(REQUIRED) - sym)
    assert _declared() <= sym, sorted(_declared() - sym)
    if os.path.exists(FLIB):
        s2 = _exported(FLIB)
        assert set(REQUIRED) <= s2 and _declared() <= s2
    # the COMMON blocks are weak references, not definitions: the driver's own blocks must win
    und = subprocess.run(["nm", "-D", "--undefined-only", FLIB_EMU], capture_output=True, text=True).stdout
    for blk in ("blksiz_", "blkcon_", "blk1d_", "blk2d_", "blk3d_", "bdry_"):
        assert re.search(rf"\bw {blk}\b", und), blk


def test_layout_header_matches_the_reference_include():
    """pom_common_layout.h is generated from pom.h_dist; the numbers SURVEY.md 8(b) quotes must come out:
    blkcon = 22 doubles, 4 ints, 16 doubles, 14 ints, with ispi / isp2i DOUBLE despite their names."""
    h = open(os.path.join(ROOT, "include", "pom_common_layout.h")).read()
    blk = re.search(r"POMF_BLKCON\[\] = \{(.*?)\};", h, re.S).group(1)
    types = "".join(re.findall(r"\{\"\w+\", '(\w)'", blk))
    assert types == "d" * 22 + "i" * 4 + "d" * 16 + "i" * 14
    assert re.search(r"\{\"ispi\", 'd'", blk) and re.search(r"\{\"isp2i\", 'd'", blk)
    for blk_, n in (("BLK2D", 73), ("BLK3D", 40), ("BDRY", 76), ("BLK1D", 4), ("BLKSIZ", 8)):
        body = re.search(rf"POMF_{blk_}\[\] = \{{(.*?)\}};", h, re.S).group(1)
        assert len(re.findall(r"\{\"", body)) == n, blk_
    assert "#define POMF_IM_LOCAL 142" in h and "#define POMF_JM_LOCAL 306" in h and "#define POMF_KB 40" in h
    ref = "/root/reference/pom.h_dist"
    if os.path.exists(ref):        # regenerate and compare (the reference is not on the GPU box)
        from scripts.gen_common_layout import emit
        tmp = os.path.join(ROOT, "tests", "_emu", "layout_check.h")
        emit(ref, tmp)
        assert open(tmp).read() == h


def _build_driver(tmp_path, dims, libdir, flib, lib, *defs):
    exe = str(tmp_path / "fortran_abi_driver")
    subprocess.check_call(["gcc", "-O1", f"-DIML={dims[0]}", f"-DJML={dims[1]}", f"-DKB={dims[2]}", *defs,
                           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "fortran_abi_driver.c"),
                           "-o", exe, "-L", libdir, "-l" + flib, "-l" + lib, "-lm", "-Wl,-rpath," + libdir])
    return exe


def _run_step(tmp_path, exe, factory, dims, nstep, restore=False, **kw):
    state, out = str(tmp_path / "state.bin"), str(tmp_path / "out.bin")
    dump(state, *dims, **kw)
    r = subprocess.run([exe, state, str(nstep), out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = read_out(out)
    st, g = syn.seamount(*dims, factory, **kw)
    if restore:
        ref_restore_setup(g, st)
    for i in range(1, nstep + 1):
        g.step(i)
    assert f"{g.check_velocity():.17g}" in r.stdout
    kb = dims[2]
    for n, a in got.items():
        # after the time rotations the reference's `f` arrays equal the `n` ones (advance.f:324-330,511-514)
        src = {"uaf": "ua", "vaf": "va", "elf": "el", "uf": "u", "vf": "v"}.get(n, n)
        b = g.get(src)
        a = a.reshape(b.shape, order="F")
        if n in ("t", "tb", "s", "sb"):
            a, b = a[:, :, :kb - 1], b[:, :, :kb - 1]
        assert np.array_equal(a, b), n
    return len(got)


def _run_unit(tmp_path, exe, factory, dims, nstep, **kw):
    state, out = str(tmp_path / "state.bin"), str(tmp_path / "out_unit.bin")
    dump(state, *dims, **kw)
    r = subprocess.run([exe, state, str(nstep), out, "unit"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = read_out(out)
    _, g = syn.seamount(*dims, factory, **kw)
    for i in range(1, nstep + 1):
        g.step(i)
    shp = g.get("u").shape
    g.advct()
    assert np.array_equal(got["advx"].reshape(shp, order="F"), g.get("advx"))
    assert np.array_equal(got["advy"].reshape(shp, order="F"), g.get("advy"))
    g.dens("s", "t", "s3c")
    rho = g.get("s3c")
    assert np.array_equal(got["rho_local"].reshape(shp, order="F")[:, :, :-1], rho[:, :, :-1])
    g.put("s3c", np.asfortranarray(got["rho_local"].reshape(shp, order="F")))
    g.advq_fields("q2b", "q2", "s3c")
    assert np.array_equal(got["qf_local"].reshape(shp, order="F"), g.get("s3c"))
    g.put("s3c", g.get("t"))
    g.proft("s3c", "wtsurf", "tsurf", 1)
    assert np.array_equal(got["proft_local"].reshape(shp, order="F"), g.get("s3c"))
    # the host copies of uf, vf equal u, v after the rotations (advance.f:511-514; mirror of the pull)
    g.put("uf", g.get("u")); g.put("vf", g.get("v"))
    g.bcond(6)
    assert np.array_equal(got["uf_bcond6"].reshape(shp, order="F"), g.get("uf"))
    g.bcondorl(3)
    assert np.array_equal(got["vf_bcondorl3"].reshape(shp, order="F"), g.get("vf"))


def test_c_driver_through_the_fortran_abi_on_host_emulation(tmp_path):
    emu.build_emu()
    dims = (22, 18, 8)
    exe = _build_driver(tmp_path, dims, os.path.dirname(FLIB_EMU), "pomgpu_f_emu", "pomgpu_emu")
    assert _run_step(tmp_path, exe, emu.EmuPom, dims, 4, island=True) >= 40
    _run_unit(tmp_path, exe, emu.EmuPom, dims, 3, island=True)


@pytest.mark.gpu
def test_c_driver_through_the_fortran_abi_on_gpu(tmp_path):
    dims = (64, 48, 14)
    exe = _build_driver(tmp_path, dims, os.path.dirname(FLIB), "pomgpu_f", "pomgpu")
    assert _run_step(tmp_path, exe, PomGpu, dims, 5, island=True) >= 40
    _run_unit(tmp_path, exe, PomGpu, dims, 3, island=True)


@pytest.mark.gpu
def test_fortran_abi_npg2_and_nadv1_on_gpu(tmp_path):
    dims = (40, 36, 12)
    exe = _build_driver(tmp_path, dims, os.path.dirname(FLIB), "pomgpu_f", "pomgpu")
    _run_step(tmp_path, exe, PomGpu, dims, 4, npg=2, nadv=1)


def test_restore_interior_records_are_called_back_from_mode_internal(tmp_path):
    """restore_interior is called from INSIDE mode_internal (advance.f:452).  An executable that defines
    `restore_interior_records_` (the record half, bounds_forcing.f:1023-1081, left in Fortran by make_glue.py) gets it
    called at that place, the records pushed on the steps the routine re-reads them, and the nudging switched on:
    equal, bit for bit, to the name-addressed run that was handed the same records."""
    emu.build_emu()
    dims = (22, 18, 8)
    exe = _build_driver(tmp_path, dims, os.path.dirname(FLIB_EMU), "pomgpu_f_emu", "pomgpu_emu", "-DWITH_RESTORE")
    assert _run_step(tmp_path, exe, emu.EmuPom, dims, 4, restore=True, island=True) >= 40
    # and the nudging did change the result: without the records routine the same driver gives another t
    exe0 = _build_driver(tmp_path, dims, os.path.dirname(FLIB_EMU), "pomgpu_f_emu", "pomgpu_emu")
    with pytest.raises(AssertionError):
        _run_step(tmp_path, exe0, emu.EmuPom, dims, 4, restore=True, island=True)


@pytest.mark.gpu
def test_restore_interior_records_on_gpu(tmp_path):
    dims = (40, 36, 12)
    exe = _build_driver(tmp_path, dims, os.path.dirname(FLIB), "pomgpu_f", "pomgpu", "-DWITH_RESTORE")
    assert _run_step(tmp_path, exe, PomGpu, dims, 4, restore=True) >= 40


def test_make_glue_cuts_exactly_the_four_step_routines(tmp_path):
    """scripts/make_glue.py on the reference's own advance.f: the four step
    routines disappear, every other line of the file survives verbatim.  The reference tree exists in
    the build container only."""
    src = "/root/reference/pom/advance.f"
    if not os.path.exists(src):
        pytest.skip("reference tree not present")
    from scripts.make_glue import CUT, cut
    text, removed = cut(open(src).read())
    assert sorted(removed) == sorted(CUT)
    subs = re.findall(r"^\s+subroutine\s+(\w+)", text, flags=re.M | re.I)
    assert subs == ["advance", "get_time", "surface_forcing", "print_section", "check_velocity", "domain_stats"]
    kept = [l for l in text.splitlines() if not l.startswith("! [")]
    orig = open(src).read().splitlines()
    it = iter(orig)
    assert all(any(l == o for o in it) for l in kept)      # kept lines are a subsequence of the original
    for call in ("call lateral_viscosity", "call mode_interaction", "call mode_external", "call mode_internal"):
        assert call in text                                  # `advance` still calls them (advance.f:21-32)
    # the record half of restore_interior (bounds_forcing.f:1023-1081) as `restore_interior_records`: the routine's
    # own lines up to "linear interpolation in time", nothing after
    from scripts.make_glue import restore_records
    bf = open("/root/reference/pom/bounds_forcing.f").read()
    rec = restore_records(bf)
    body = [l for l in rec.splitlines() if not l.startswith("! [")]
    assert body[0].split() == ["subroutine", "restore_interior_records"] and body[-2:] == ["      return", "      end"]
    ref_lines = bf.splitlines()
    it = iter(ref_lines)
    assert all(any(l == o for o in it) for l in body[1:-2])   # verbatim, in order
    assert "read_restore_ts_interior_pnetcdf" in rec and "trstrb(i,j,k)=trstrf(i,j,k)" in rec
    assert "fold*trstrb" not in rec and "taurstr(i,j,k)*" not in rec      # the arithmetic half is the library's
