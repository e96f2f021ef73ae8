"""Developer tool: a variant build of the CUDA library (POMGPU_LIB=...) must give bitwise the same
fields as the default build after a few steps.  Each library runs in its own process (two builds in one
process share the function-local statics of the inline launchers, e.g. the shared-memory opt-in record)."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

FIELDS = "u v t s q2 q2l km kh el w rho ub vb tb aam ua va".split()


def child(out):
    from extpom_b200 import synthetic as syn
    from extpom_b200 import pomgpu as _pg
    lib = os.environ.get("POMGPU_LIB")
    cls = _pg.PomGpu
    if lib:
        class Variant(_pg.PomGpu):
            @staticmethod
            def _library():
                return _pg._lib(lib)
        cls = Variant
    dims = tuple(int(x) for x in os.environ.get("VARIANT_DIMS", "256,200,41").split(","))
    _, a = syn.seamount(*dims, cls, island=True)
    for i in range(1, 5):
        a.step(i)
    np.savez(out, **{n: a.get(n) for n in FIELDS})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2])
        sys.exit(0)
    lib = os.environ["POMGPU_LIB"]
    with tempfile.TemporaryDirectory() as d:
        env = dict(os.environ)
        subprocess.check_call([sys.executable, __file__, "--child", d + "/b.npz"], env=env)
        env.pop("POMGPU_LIB")
        if env.get("POMGPU_LIB_A"):             # compare two variants with each other (or one with itself: determinism)
            env["POMGPU_LIB"] = env["POMGPU_LIB_A"]
        subprocess.check_call([sys.executable, __file__, "--child", d + "/a.npz"], env=env)
        a, b = np.load(d + "/a.npz"), np.load(d + "/b.npz")
        bad = [n for n in FIELDS if not np.array_equal(a[n], b[n])]
    print("variant", lib, "BITWISE EQUAL to default" if not bad else f"MISMATCH {bad}")
