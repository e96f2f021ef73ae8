#!/usr/bin/env python
"""fuzz_reference.py SEED N -- random differential runs against the reference's own source (build container only).

Draws N random grids (incl. odd im), namelist switches (mode, nadv, nitera, sw, npg, nbct, nbcs, ntp, ispadv, smoth,
horcon, tprni, umol, tbias/sbias, ramp, a restart time0) and state variants (island, all four sides open, non-zero
e_atmos / vflux / wssurf, open-boundary values) and takes three internal steps with
  * the reference's Fortran source, executed by oracle/f77ref.py,
  * the C oracle (must be BITWISE equal),
  * the host build of the CUDA kernel bodies through the C ABI (<= 1e-11),
  * the same bodies on TWO STRIPS behind the gfortran ABI (tests/fabi.py; <= 1e-11).
Round 2: seeds 1-5, 185 draws, 0 mismatches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, random, traceback
from oracle.pomo import Oracle
from oracle.f77ref import F77Ref
from tests.emu import EmuPom
from tests.fabi import FabiEmu, strips
from scripts import make_ref_golden as mrg
rnd=random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 1)
def combo():
    kw={}
    if rnd.random()<.5: kw["island"]=True
    if rnd.random()<.5: kw.update(walls=False)
    if rnd.random()<.6: kw["fluxes"]=True
    if rnd.random()<.6: kw["obc"]=True
    kw["nadv"]=rnd.choice([1,2,2]); kw["npg"]=rnd.choice([1,1,2]); kw["mode"]=rnd.choice([3,3,3,4,2])
    kw["nbct"]=rnd.choice([1,2,3,4]); kw["nbcs"]=rnd.choice([1,1,3]); kw["ntp"]=rnd.choice([1,2,3,4,5])
    if kw["nadv"]==2: kw["nitera"]=rnd.choice([1,1,2,3]); kw["sw"]=rnd.choice([0.5,1.0,0.8])
    s={}
    if rnd.random()<.4: s["ispadv"]=rnd.choice([1,2,5,7])
    if rnd.random()<.3: s["time0"]=rnd.choice([0.5,3.0])
    if rnd.random()<.3: s["ramp"]=rnd.choice([0.2,0.9])
    if rnd.random()<.3: s["smoth"]=rnd.choice([0.0,0.05,0.2])
    if rnd.random()<.3: s["horcon"]=rnd.choice([0.05,0.3])
    if rnd.random()<.3: s["tprni"]=rnd.choice([0.0,0.5])
    if rnd.random()<.2: s["umol"]=1e-5
    if rnd.random()<.2: s["tbias"]=2.0; s["sbias"]=1.0
    if s: kw["_set"]=s
    dims=(rnd.choice([12,13,16,17,20]), rnd.choice([14,15,16,19]), rnd.choice([6,7,9,12]))
    return dims, kw
nbad=0
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 20):
    dims,kw=combo()
    try:
        res={}
        facts=[("ref",F77Ref),("oracle",Oracle),("emu",EmuPom),("fabi2",strips(FabiEmu,2,ghost=2))]
        for name,F in facts:
            st,g=mrg.loaded(F,dims,kw)
            for i in range(1,4):
                if name!="ref": mrg.ref_restore_records(g,st,i)
                g.step(i)
            res[name]={n:g.get(n) for n in mrg.F3+mrg.F2}
        fin=all(np.isfinite(v).all() for v in res["ref"].values())
        msgs=[]
        for other in ("oracle","emu","fabi2"):
            bad={}
            for n in res["ref"]:
                if other!="oracle" and n in ("uf","vf"): continue
                a,b=res["ref"][n],res[other][n]
                if n in ("t","tb","s","sb") and other!="oracle": a,b=a[:,:,:-1],b[:,:,:-1]
                if other=="oracle":
                    if not np.array_equal(a,b): bad[n]=float(np.nanmax(np.abs(a-b))/(np.nanmax(np.abs(a))+1e-300))
                else:
                    e=np.nanmax(np.abs(a-b))/(np.nanmax(np.abs(a))+1e-300)
                    if not e<=1e-11: bad[n]=float(e)
            if bad: msgs.append((other,bad))
        status="OK" if not msgs else "MISMATCH"
        if msgs: nbad+=1
        print(status, "finite" if fin else "NONFINITE", dims, kw, msgs if msgs else "", flush=True)
    except Exception as e:
        nbad+=1
        print("EXC", dims, kw, repr(e)[:300], flush=True)
print("bad", nbad)
