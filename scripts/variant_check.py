"""Developer tool: a variant build of the CUDA library (POMGPU_LIB=...) must give bitwise the same
fields as the default build after a few steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from extpom_b200 import synthetic as syn
from extpom_b200 import pomgpu as _pg

lib = os.environ["POMGPU_LIB"]


class Variant(_pg.PomGpu):
    @staticmethod
    def _library():
        return _pg._lib(lib)


dims = (256, 200, 41)
_, a = syn.seamount(*dims, _pg.PomGpu, island=True)
_, b = syn.seamount(*dims, Variant, island=True)
for i in range(1, 5):
    a.step(i); b.step(i)
bad = [n for n in "u v t s q2 q2l km kh el w rho ub tb".split() if not np.array_equal(a.get(n), b.get(n))]
print("variant", lib, "BITWISE EQUAL to default" if not bad else f"MISMATCH {bad}")
