"""Ad-hoc GPU check: parity vs the oracle on a small grid, then wall-clock step time on a big one."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from extpom_b200 import synthetic as syn
from extpom_b200.pomgpu import PomGpu

NAMES = ("u ub v vb t tb s sb q2 q2b q2l q2lb w rho km kh kq aam advx advy drhox drhoy l wr el elb et etb "
         "etf ua uab va vab d dt egf egb utf vtf utb vtb wubot wvbot adx2d ady2d drx2d dry2d aam2d advua advva").split()

def parity(im, jm, kb, n, **kw):
    from oracle.pomo import Oracle
    st, o = syn.seamount(im, jm, kb, Oracle, **kw)
    st2, g = syn.seamount(im, jm, kb, PomGpu, **kw)
    for iint in range(1, n + 1):
        o.step(iint); g.step(iint)
    worst, wn = 0.0, ""
    for nme in NAMES:
        a = o.get(nme); b = g.get(nme)
        if nme in ("t", "tb", "s", "sb"): a = a[:, :, :kb - 1]; b = b[:, :, :kb - 1]
        r = np.abs(a - b).max() / (np.abs(a).max() + 1e-300)
        if not (r <= worst): worst, wn = r, nme
    print("PARITY", (im, jm, kb, n, kw), "worst rel", wn, "%.3e" % worst, "vamax", o.check_velocity(), g.check_velocity(), flush=True)

def timing(im, jm, kb, nsteps, **kw):
    t0 = time.time()
    st, g = syn.seamount(im, jm, kb, PomGpu, **kw)
    print("init s", time.time() - t0, flush=True)
    del st
    for iint in range(1, 4):
        g.step(iint)
    g.sync()
    t0 = time.time()
    for iint in range(4, 4 + nsteps):
        g.step(iint)
    g.sync()
    dt = (time.time() - t0) / nsteps
    print("TIMING", (im, jm, kb), "ms/step %.3f" % (dt * 1e3), "cells/s %.3e" % (im * jm * kb / dt), "vamax", g.check_velocity(), flush=True)

if __name__ == "__main__":
    parity(40, 31, 16, 30)
    parity(24, 19, 9, 8, island=True, nadv=1)
    parity(65, 49, 21, 50)
    if len(sys.argv) > 1:
        n = int(sys.argv[1])
        timing(n, n, 41, 10)
