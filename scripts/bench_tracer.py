"""BASELINE configs[2]: tracer-only bench -- advt2 (MPDATA) of T and S + proft vertical diffusion on a
2048 x 2048 x 41 state on one B200, for nitera=1 and nitera=3 (sw=1).  Only the fields these routines
read are generated (one 1.4 GB array at a time); cell-updates/s = im*jm*kb / time of the four routines."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from extpom_b200 import synthetic as syn  # noqa: E402
from extpom_b200.pomgpu import PomGpu  # noqa: E402
if os.environ.get("POMGPU_LIB"):           # a variant build of the CUDA library, for A/B timing (developer tool)
    from extpom_b200 import pomgpu as _pg
    _lib = os.environ["POMGPU_LIB"]

    class PomGpu(_pg.PomGpu):             # noqa: F811
        @staticmethod
        def _library():
            return _pg._lib(_lib)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
kb = int(sys.argv[2]) if len(sys.argv) > 2 else 41
reps = 10
st = syn.make_state(n, n, 4)                      # 2-D fields and constants (independent of kb)
f2, c = st["fields"], st["consts"]
g = PomGpu(n, n, kb)
for k, v in c.items():
    g.L.pomgpu_set_const(g.h, k.encode(), float(v))
z, zz, dz, dzz = syn.sigma_levels(kb)
for nme, a in (("z", z), ("zz", zz), ("dz", dz), ("dzz", dzz)):
    g.put(nme, a)
for nme in "dx dy h fsm dum dvm art aru arv cor dt etb etf wtsurf wssurf swrad".split():
    g.put(nme, f2[nme])
h, fsm, dum, dvm = f2["h"], f2["fsm"], f2["dum"], f2["dvm"]
rng = np.random.default_rng(syn.SEED)
shape = (n, n, kb)


def push(name, make):
    a = np.asfortranarray(make())
    g.put(name, a)
    return a


tb = push("tb", lambda: (5.0 + 15.0 * np.exp(zz[None, None, :] * h[:, :, None] / 1000.0)
                         + 1e-2 * rng.standard_normal(shape)) * fsm[:, :, None])
g.put("t", tb); g.put("tclim", tb); g.put("tsurf", np.asfortranarray(tb[:, :, 0]))
del tb
sb = push("sb", lambda: (35.0 + 1e-2 * rng.standard_normal(shape)) * fsm[:, :, None])
g.put("s", sb); g.put("sclim", sb); g.put("ssurf", np.asfortranarray(sb[:, :, 0]))
del sb
push("u", lambda: (0.2 + 1e-2 * rng.uniform(-1, 1, shape)) * dum[:, :, None])
push("v", lambda: 1e-2 * rng.uniform(-1, 1, shape) * dvm[:, :, None])
push("w", lambda: 1e-5 * rng.uniform(-1, 1, shape) * fsm[:, :, None])
push("aam", lambda: np.full(shape, 500.0))
push("kh", lambda: 1e-3 * (1.0 + rng.random(shape)))
q = np.full(shape, 1e-9, order="F")
g.put("q2", q); g.put("q2l", q)
del q
res = {}
for nitera, sw in ((1, 0.5), (3, 1.0)):
    g.set("nitera", nitera); g.set("sw", sw)

    def step():
        g.internal_stage(2, 105)      # advt2 of T (-> uf) and S (-> vf)
        g.internal_stage(2, 107)      # proft of both

    for _ in range(3):
        step()
    g.sync(); g.event_record(0)
    for _ in range(reps):
        step()
    g.event_record(1)
    ms = g.event_elapsed_ms(0, 1) / reps
    g.profile_begin(); step(); prof = g.profile_end()
    res[f"nitera={nitera}"] = {"ms": ms, "cell_updates_per_s": n * n * kb / (ms * 1e-3),
                               "kernels": {r["name"]: {"ms": r["ms"], "GBps": r["bytes"] / r["ms"] / 1e6} for r in prof}}
    print(f"tracer bench {n}x{n}x{kb} nitera={nitera} sw={sw}: {ms:.3f} ms  {n * n * kb / (ms * 1e-3):.3e} cell-updates/s")
    for r in prof:
        print(f"    {r['name']:14s} x{r['launches']} {r['ms']:.3f} ms  {r['bytes'] / r['ms'] / 1e6:7.1f} GB/s")
assert np.isfinite(g.get("uf")).all()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_tracer.json"), "w"), indent=1)
