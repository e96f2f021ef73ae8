"""torchrun --nproc-per-node N scripts/halo_trace.py [steps]: per-step device time and the halo
exchanges inside each step at 1024 x 1024N x 41 (developer tool; POMGPU_HALO_TRACE=1 must be set)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from extpom_b200.strips import StripSet
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m = StripSet.create(1024, 1024 * world, 41, rank, world, device=local, dist=dist)
g = m.gpu
buf = C.create_string_buffer(1 << 16)
for i in range(1, 7):
    m.step(i)
g.sync(); dist.barrier()
g.L.pomgpu_group_halo_trace(C.c_void_p(m.group.h), buf, len(buf))
out = []
for i in range(7, 7 + nsteps):
    g.sync(); dist.barrier()
    t0 = time.perf_counter()
    g.event_record(0)
    m.step(i)
    g.event_record(1)
    th = (time.perf_counter() - t0) * 1e3
    ms = g.event_elapsed_ms(0, 1)
    g.L.pomgpu_group_halo_trace(C.c_void_p(m.group.h), buf, len(buf))
    out.append((i, ms, th, buf.value.decode()))
if rank == 0:
    for i, ms, th, tr in out:
        big = [l for l in tr.splitlines() if float(l.split("ms")[1].split("MB")[0]) > 5]
        print("step %d: %.3f ms (host enqueue %.2f ms), %d exchanges, large ones:" % (i, ms, th, len(tr.splitlines())))
        for l in big:
            print("    " + l)
dist.destroy_process_group()
