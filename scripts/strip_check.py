"""torchrun --nproc-per-node N scripts/strip_check.py [im jm kb steps ghost]
One strip per GPU connected over NCCL; every rank compares its owned rows bitwise with a
single-domain run of the same case done on its own GPU (small grids)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from extpom_b200 import synthetic as syn
from extpom_b200.pomgpu import PomGpu
from extpom_b200.strips import StripSet

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
a = [int(x) for x in sys.argv[1:]]
im, jm, kb, steps, ghost = (a + [96, 160, 16, 6, 4][len(a):])[:5]
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m = StripSet.create(im, jm, kb, rank, world, device=local, dist=dist, ghost=ghost, island=True)
_, whole = syn.seamount(im, jm, kb, lambda x, y, z: PomGpu(x, y, z, device=local), island=True)
for i in range(1, steps + 1):
    m.step(i); whole.step(i)
bad = []
j0, j1 = m.rows
for n in "u v t s q2 q2l el ua va w km kh rho ub tb etf wubot aam advx".split():
    x = whole.get(n)[:, j0 - 1:j1]
    y = m.group.gather(n)
    if n in ("t", "tb", "s"): x, y = x[:, :, :-1], y[:, :, :-1]
    if not np.array_equal(x, y): bad.append((n, float(np.abs(x - y).max())))
nex, nf = m.group.exchanges()
print(f"rank {rank}/{world} rows {m.rows}: {'BITWISE EQUAL' if not bad else 'MISMATCH ' + str(bad)}; "
      f"{nex / steps:.1f} exchanges/step", flush=True)
t = torch.tensor([len(bad)], device="cuda"); dist.all_reduce(t)
dist.destroy_process_group()
sys.exit(1 if int(t.item()) else 0)
