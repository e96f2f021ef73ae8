"""Per-kernel CUDA-event times of a few steps (same numbers bench.py's roofline uses)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from extpom_b200 import synthetic as syn
from extpom_b200.pomgpu import PomGpu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
kb = int(sys.argv[2]) if len(sys.argv) > 2 else 41
lib = os.environ.get("POMGPU_LIB")      # a variant build of the CUDA library (extpom_b200/variants/*.so), for A/B timing
if lib:
    from extpom_b200 import pomgpu as _pg

    class PomGpu(_pg.PomGpu):             # noqa: F811  (developer tool: never part of the product path)
        @staticmethod
        def _library():
            return _pg._lib(lib)
st, g = syn.seamount(n, n, kb, PomGpu)
print("lib:", lib or "default")
del st
for i in range(1, 4): g.step(i)
g.sync()
g.event_record(0)
for i in range(4, 9): g.step(i)
g.event_record(1)
print("ms/step %.3f" % (g.event_elapsed_ms(0, 1) / 5))
g.profile_begin()
for i in range(9, 12): g.step(i)
prof = g.profile_end()
prof.sort(key=lambda r: -r["ms"])
tot = sum(r["ms"] for r in prof)
print("sum of kernels ms/step %.3f" % (tot / 3))
only = os.environ.get("KPROF_ONLY", "").split(",") if os.environ.get("KPROF_ONLY") else None
for r in prof:
    if only and r["name"] not in only: continue
    print("%-18s n=%3d  %8.3f ms/step  %7.1f GB/s  %5.1f%%" % (r["name"], r["launches"] // 3, r["ms"] / 3, r["bytes"] / r["ms"] / 1e6, 100 * r["ms"] / tot))
