"""Three internal steps at a given size -- the command profiled under ncu (profiles/README.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from extpom_b200 import synthetic as syn
from extpom_b200.pomgpu import PomGpu

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
kb = int(sys.argv[2]) if len(sys.argv) > 2 else 41
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
st, g = syn.seamount(n, n, kb, PomGpu)
del st
for i in range(1, steps + 1):
    g.step(i)
g.sync()
print("ok", g.check_velocity(), g.launch_count())
