import os, sys, time
sys.path.insert(0, "/root/repo")
from extpom_b200 import synthetic as syn
from extpom_b200.pomgpu import PomGpu
from extpom_b200.strips import StripSet
n, kb = 1024, 41
st, g = syn.seamount(n, n, kb, PomGpu)
del st
i = 0
for _ in range(3):
    i += 1; g.step(i)
for K in (5, 20, 20, 40):
    g.sync(); g.event_record(0)
    for _ in range(K):
        i += 1; g.step(i)
    g.event_record(1); g.sync()
    print("PomGpu   K=%2d  %.3f ms/step (iint up to %d)" % (K, g.event_elapsed_ms(0, 1) / K, i))
g.close()
m = StripSet.create(n, n, kb, 0, 1, device=0, dist=None)
g = m.gpu
i = 0
for _ in range(3):
    i += 1; m.step(i)
for K in (5, 20, 20, 40):
    g.sync(); g.event_record(0)
    for _ in range(K):
        i += 1; m.step(i)
    g.event_record(1); g.sync()
    print("StripSet K=%2d  %.3f ms/step (iint up to %d)" % (K, g.event_elapsed_ms(0, 1) / K, i))
