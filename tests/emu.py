"""Host-compiled build of libpomgpu's kernel bodies -- FOR CPU-SIDE TESTS ONLY.

The container that runs `pytest -m "not gpu"` has no GPU.  To still check the host
logic (orchestration, pointer rotations, C ABI) and the kernel bodies against the
oracle there, the same .cu sources are compiled with g++ and -DPOMGPU_EMU into
tests/_emu/libpomgpu_emu.so, where a "launch" is a plain loop over columns.  The
product (extpom_b200.PomGpu) never loads this library; it is reachable only through
this test helper.
"""
import os
import subprocess

from extpom_b200 import pomgpu as _pg

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_SO = os.path.join(_ROOT, "tests", "_emu", "libpomgpu_emu.so")


def build_emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "extpom_b200", "csrc"), "emu"])
    return EMU_SO


class EmuPom(_pg.PomGpu):
    """PomGpu bound to the host-emulated library (test infrastructure; not importable from the package)."""

    @staticmethod
    def _library():
        return _pg._lib(build_emu())
