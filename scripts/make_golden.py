"""Generate tests/golden/*.npz from the CPU oracle (the only runnable statement of the
reference here: no Fortran compiler in the image).  The fixtures pin the
oracle and the generator against regressions; they are NOT outputs of the Fortran code."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from extpom_b200 import synthetic as syn  # noqa: E402
from oracle.pomo import Oracle  # noqa: E402
from tests.common import F2, F3, digest  # noqa: E402

CASES = {
    "seamount_24x19x9_nadv2": dict(dims=(24, 19, 9), steps=6, kw=dict(island=True)),
    "seamount_24x19x9_nadv1": dict(dims=(24, 19, 9), steps=6, kw=dict(island=True, nadv=1)),
    "seamount_30x22x12_swrad": dict(dims=(30, 22, 12), steps=5, kw=dict(nbct=2, ntp=3)),
}

if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, c in CASES.items():
        st, o = syn.seamount(*c["dims"], Oracle, **c["kw"])
        inp = {k: digest(v) for k, v in sorted(st["fields"].items())}
        for i in range(1, c["steps"] + 1):
            o.step(i)
        out = {n: o.get(n) for n in F3 + F2}
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"),
                            input_digest=np.array([f"{k}:{v}" for k, v in inp.items()]),
                            vamax=o.check_velocity(), **out)
        print(name, "vamax", o.check_velocity())
