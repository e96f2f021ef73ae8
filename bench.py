#!/usr/bin/env python
"""bench.py -- grid-cell updates/s of one internal step of the extPOM hot path on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: CPU restatement

A "step" is one pass of advance.f:21-32 (lateral_viscosity, mode_interaction, isplit x
mode_external, mode_internal) over the synthetic seamount state of BASELINE.json
configs[1] (1024 x 1024 x 41, isplit=30) per GPU.  `value` = im*jm*kb*K / device time of
the K steps with the state resident in HBM; `e2e` = the same K steps through the public
API with the per-step forcing pushed from pinned host buffers and check_velocity's scalar
read back, host<->device copies inside the timed region.

The reference (Fortran+MPI+PnetCDF) cannot be built in this image, so the reference arm
and `cpu_baseline` time the C restatement under oracle/ (kind "port") on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid-cell updates/sec per internal step"
UNIT = "cell-updates/s"
FORCING_2D = ("wusurf", "wvsurf", "wtsurf", "swrad", "tsurf")          # bounds_forcing.f:908-909,954-955,978
FORCING_BDY = ("tbe", "sbe", "tbw", "sbw", "tbn", "sbn", "tbs", "sbs",  # bounds_forcing.f:844-865
               "uabe", "uabw", "vabn", "vabs", "ele", "els")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--im", type=int, default=1024)
    ap.add_argument("--jm", type=int, default=1024, help="rows PER GPU (weak scaling)")
    ap.add_argument("--kb", type=int, default=41)
    ap.add_argument("--cpu-sample", type=int, default=512, help="im=jm of the CPU-baseline sample grid")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.p = [], None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(n, kb, steps, warmup):
    """Oracle (C restatement, OpenMP over all host cores) on an n x n x kb seamount sample."""
    from oracle.pomo import Oracle
    from extpom_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    st, o = syn.seamount(n, n, kb, Oracle)
    for i in range(1, warmup + 1):
        o.step(i)
    t0 = time.perf_counter()
    for i in range(warmup + 1, warmup + steps + 1):
        o.step(i)
    dt = time.perf_counter() - t0
    return n * n * kb * steps / dt, dt / steps, cores


def run_reference(a, rank):
    if rank != 0:
        return
    val, spt, cores = cpu_reference_rate(a.cpu_sample, a.kb, a.steps, max(a.warmup, 1))
    sample = (f"{a.steps} internal steps of the seamount state at {a.cpu_sample}x{a.cpu_sample}x{a.kb} "
              f"(isplit=30) after {max(a.warmup, 1)} warm-up steps; C restatement of advance.f/solver.f, "
              f"gcc -O2 -ffp-contract=off, OpenMP x{cores}")
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": spt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"seamount {a.im}x{a.jm}x{a.kb} per GPU, isplit=30 (BASELINE configs[1])",
                   "l2": "inputs larger than L2"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


_REAL_STDOUT = None


def _guard_stdout():
    """stdout carries exactly ONE JSON line: anything a library prints there while the job runs
    (e.g. NCCL's version banner) is sent to stderr; emit() writes to the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(obj):
    f = _REAL_STDOUT or sys.stdout
    f.write(json.dumps(obj) + "\n")
    f.flush()


def main():
    a = parse()
    _guard_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        return run_reference(a, rank)

    import numpy as np
    import torch
    from extpom_b200 import synthetic as syn
    from extpom_b200.pomgpu import PomGpu
    from extpom_b200.strips import StripSet

    dist = None
    if world > 1:
        # NCCL_DEBUG=VERSION (set on some boxes) makes NCCL print its banner to stdout
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG")
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, W = a.steps, max(a.warmup, 3)
    im, kb = a.im, a.kb
    jm_global = a.jm * world            # weak scaling: a.jm rows per GPU, strips stacked in j
    model = StripSet.create(im, jm_global, kb, rank, world, device=local, dist=dist)
    g = model.gpu
    cells = im * jm_global * kb

    def barrier():
        g.sync()
        if dist is not None:
            dist.barrier()

    def maxreduce(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    iint = 0
    for _ in range(W):
        iint += 1
        model.step(iint)
    barrier()

    # ---- timed region 1: state resident in HBM --------------------------------------
    sampler = ClockSampler(local)
    g.launch_count(reset=True)
    barrier()
    t0 = time.time()
    g.event_record(0)
    th0 = time.perf_counter()
    for _ in range(K):
        iint += 1
        model.step(iint)
    host_ms = (time.perf_counter() - th0) * 1e3 / K      # host time to ENQUEUE one step (asynchronous)
    g.event_record(1)
    barrier()
    t1 = time.time()
    ms = maxreduce(g.event_elapsed_ms(0, 1))
    launches = g.launch_count(reset=True)
    clocks = sampler.stop(t0, t1)
    value = cells * K / (ms * 1e-3)

    # ---- timed region 2: end to end through the public API with host buffers --------
    pins = {n: g.pinned(n) for n in FORCING_2D + FORCING_BDY}
    for n, buf in pins.items():
        buf[...] = g.get(n)
    h2d = sum(b.nbytes for b in pins.values())

    def e2e_step():
        nonlocal iint
        iint += 1
        for n, buf in pins.items():
            g.put_async(n, buf)                 # per-step forcing (bounds_forcing.f:844-865,908-978)
        model.step(iint)
        # advance.f:52: one scalar back per step -- read with one step of lag, so that the host
        # can enqueue the next step's forcing copies while this step still computes
        return g.check_velocity_lagged()

    for _ in range(2):                          # untimed: allocates the shadow buffers of the async pushes
        e2e_step()
    barrier()
    g.event_record(2)
    vmax = 0.0
    for _ in range(K):
        vmax = max(vmax, e2e_step())
    vmax = max(vmax, g.check_velocity())       # the last step's own value (waits)
    g.event_record(3)
    barrier()
    ms_e2e = maxreduce(g.event_elapsed_ms(2, 3))
    e2e = cells * K / (ms_e2e * 1e-3)

    # ---- per-kernel CUDA-event times for the roofline of the dominant kernel ----------
    g.profile_begin()
    nprof = 3
    for _ in range(nprof):
        iint += 1
        model.step(iint)
    prof = g.profile_end()
    prof.sort(key=lambda r: -r["ms"])
    peaks, peak_src = None, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "measured"
    except Exception:
        peak = 6650.0
    top = prof[0]
    ach = top["bytes"] / (top["ms"] * 1e-3) / 1e9
    traffic = None
    try:   # dram bytes per launch of the dominant kernel from the committed ncu capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = tr.get(top["name"], {}).get(f"{im}x{a.jm}x{kb}")
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": top["name"], "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
            "ms_per_launch": top["ms"] / top["launches"],
            "algorithmic_bytes_per_launch": top["bytes"] / top["launches"],
            "share_of_step": top["ms"] / sum(r["ms"] for r in prof)}
    tot_ms = sum(r["ms"] for r in prof)
    kernels = [{"name": r["name"], "ms_per_step": r["ms"] / nprof, "launches_per_step": r["launches"] / nprof,
                "GBps": r["bytes"] / (r["ms"] * 1e-3) / 1e9, "frac": r["bytes"] / (r["ms"] * 1e-3) / 1e9 / peak,
                "share": r["ms"] / tot_ms} for r in prof]

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"seamount {im}x{a.jm}x{kb} per GPU (global {im}x{jm_global}x{kb}), isplit=30, "
                               "nadv=2 nitera=1 mode=3" + (" (BASELINE configs[1])" if (im, a.jm, kb) == (1024, 1024, 41) else ""),
                   "parallelism": f"j-strips x{world}",
                   "l2": f"inputs larger than L2 ({41 * im * a.jm * kb * 8 / 1e9:.0f} GB state per GPU)",
                   "seed": syn.SEED},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                "ms_per_step": ms_e2e / K, "vamax": vmax},
        "gpu_launches": launches,
        "host_enqueue_ms_per_step": host_ms,
        "roofline": roof,
        "kernels": kernels,
    }
    if world == 1 and not a.no_cpu_baseline:
        val, spt, cores = cpu_reference_rate(a.cpu_sample, kb, 8, 1)
        out["cpu_baseline"] = {
            "value": val, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"8 internal steps at {a.cpu_sample}x{a.cpu_sample}x{kb} (isplit=30) after 1 warm-up; "
                      "C restatement of advance.f/solver.f (oracle/), gcc -O2 -ffp-contract=off, OpenMP"}
    emit(out)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
