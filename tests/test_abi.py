"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/pomgpu.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from extpom_b200 import pomgpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pomgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pomgpu_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(pomgpu.LIBPATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(pomgpu.LIBPATH)
    syms = declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_emulation_library_exports_the_same_abi():
    from tests.emu import build_emu
    lib = ctypes.CDLL(build_emu())
    assert not [s for s in declared_symbols() if not hasattr(lib, s)]


def test_no_cpu_fallback():
    """Without a GPU the product must fail loudly instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pomgpu.PomGpuError):
        pomgpu.PomGpu(16, 12, 6)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "extpom_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("CPU oracle", "") or f == "synthetic.py", f
                assert "libpomo" not in txt and "pomo.h" not in txt, f


def test_field_registry_matches_python_mirror():
    lib = ctypes.CDLL(pomgpu.LIBPATH)
    # pomgpu_field_elems needs a context; only check the name tables agree via the emu build
    from tests.emu import EmuPom
    g = EmuPom(12, 10, 6)
    g.L.pomgpu_field_elems.restype = ctypes.c_long
    for n, shp in g.shapes.items():
        want = 1
        for s in shp:
            want *= s
        assert g.L.pomgpu_field_elems(g.h, n.encode()) == want, n
    del lib


def test_kb_beyond_kmax_is_rejected():
    """The column solvers keep their eliminated coefficients in KMAX=64 entries per thread:
    pomgpu_create must refuse kb > 64 instead of overrunning them."""
    import ctypes as C
    from tests import emu
    L = C.CDLL(emu.build_emu())
    L.pomgpu_create.restype = C.c_void_p
    assert not L.pomgpu_create(12, 10, 65, 0)
    assert not L.pomgpu_create(4, 10, 10, 0)       # im < 6
