"""BASELINE configs[0]: the reference's own default grid 282 x 306 x 40 (pom.h_dist:22-28) with the
pom.nml_dist defaults (mode=3, nadv=2, nitera=1, sw=0.5, npg=1, isplit=30), single sub-domain, on the
synthetic seamount state.  Per-field max-abs and relative L2 error of the CUDA path against the CPU
oracle (libm pow, i.e. what gfortran emits) after N = 1, 10 and 100 internal steps -- the tolerance
statement BASELINE.json's north_star asks for."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from extpom_b200 import synthetic as syn  # noqa: E402
from extpom_b200.pomgpu import PomGpu  # noqa: E402
from oracle.pomo import Oracle  # noqa: E402

im, jm, kb = 282, 306, 40
nmax = int(sys.argv[1]) if len(sys.argv) > 1 else 100
marks = [n for n in (1, 10, 100) if n <= nmax]
_, o = syn.seamount(im, jm, kb, Oracle)
_, g = syn.seamount(im, jm, kb, PomGpu)
out = {}
t0 = time.time()
for i in range(1, nmax + 1):
    o.step(i)
    g.step(i)
    if i in marks:
        row = {}
        for n in ("el", "u", "v", "t", "s", "q2", "q2l", "ua", "va", "w", "rho", "km", "kh"):
            a, b = o.get(n), g.get(n)
            if n in ("t", "s"):
                a, b = a[:, :, :-1], b[:, :, :-1]
            row[n] = {"max_abs": float(np.abs(a - b).max()), "field_max": float(np.abs(a).max()),
                      "rel_l2": float(np.sqrt(((a - b) ** 2).sum() / max((a ** 2).sum(), 1e-300)))}
        out[i] = row
        print(f"--- after {i} internal steps ({time.time() - t0:.0f} s); vamax oracle {o.check_velocity():.6f} "
              f"gpu {g.check_velocity():.6f}")
        for n, r in row.items():
            print(f"  {n:4s} max|diff| {r['max_abs']:.3e}  (field max {r['field_max']:.3e})  rel L2 {r['rel_l2']:.3e}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "config0_parity.json"), "w"), indent=1)
