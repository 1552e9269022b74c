#!/usr/bin/env python
"""bench.py — MLUPS of the D2Q9 lattice update on N B200s, with the HBM roofline and the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--kernel NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --single-process --gpus N ...   # one process, one blbm_create_group handle over N GPUs
    python bench.py --impl reference ...            # the reference's algorithm on the host cores (oracle)

One "step" is one lattice timestep (collide + stream of every cell of the lattice, lbm.rs:1112-1116).
Workloads (BASELINE.json configs):
    porous16384  configs[2]: 16384 x 16384 per GPU, random porous mask (15 % solid, splitmix64 seed 0x5EED),
                 u0 = 0.05, omega = 1.0 — the single-B200 bandwidth-roofline configuration (default)
    channel16384 the same lattice without the porous mask (isolates the cost of the mask)
    cavity4096   configs[1]: 4096 x 4096 closed box
    cylinder512  configs[0]: 512 x 256 channel past a cylinder (L2-resident, launch-bound; not a roofline case)
    cylinder32768 configs[3]: 32768^2 cylinder wake, strong scaling (the lattice is fixed, N slabs)
    channel65536 configs[4]: 65536 x 8192 per GPU (65536^2 on 8 GPUs), one cylinder
At N > 1 the lattice is N slabs stacked in y (weak scaling: per-GPU work fixed), linked by direct NVLink
halo stores from the step kernel; no NCCL collective is on the data path.  Before anything is timed at N > 1 the
linked slabs are checked bit for bit against the undivided lattice (`multi_gpu_parity`), and after the headline
workload the north star's own multi-GPU configurations are timed in the same run (`north_star`).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BYTES_PER_CELL = 72.0  # 9 fp32 loads + 9 fp32 stores per cell update
L2_BYTES = 126 << 20

WORKLOADS = {
    #  name          W      rows/GPU  omega  u0    mask
    "porous16384": (16384, 16384, 1.0, 0.05, "porous"),
    "channel16384": (16384, 16384, 1.0, 0.05, "none"),
    "cavity4096": (4096, 4096, 1.25, 0.1, "box"),
    "cylinder512": (512, 256, 1.0 / (3 * 0.02 + 0.5), 0.1, "cylinder"),
    "channel65536": (65536, 8192, 1.0, 0.05, "cylinder"),
    # configs[3]: 32768 x 32768 cylinder wake, Re 1000 (nu = 0.0512); STRONG scaling: the lattice is fixed and
    # split into N slabs (rows/GPU below is the N = 1 value)
    "cylinder32768": (32768, 32768, 1.0 / (3 * 0.0512 + 0.5), 0.1, "cylinder"),
}
STRONG = {"cylinder32768"}


def splitmix64(x):
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def mask_rows(kind, w, h_total, r0, r1, discs=()):
    """uint8 barrier mask of global rows [r0, r1): walls on rows 0 and H-1 plus the workload's obstacles;
    discs: extra (cx, cy, radius) obstacles."""
    r0, r1 = max(r0, 0), min(r1, h_total)
    m = np.zeros((r1 - r0, w), np.uint8)
    ys = np.arange(r0, r1, dtype=np.int64)
    if kind == "porous":
        xs = np.arange(w, dtype=np.uint64)
        thr = np.uint64(int(0.15 * 2.0 ** 64))
        for j, y in enumerate(ys):  # row by row keeps the temporary small
            idx = np.uint64(y) * np.uint64(w) + xs
            row = splitmix64(idx ^ np.uint64(0x5EED)) < thr
            row[:2] = False
            row[w - 1:] = False
            m[j] = row
    elif kind == "box":
        m[:, 1] = 1
        m[:, w - 1] = 1
    elif kind == "cylinder":
        discs = tuple(discs) + ((w // 4, h_total // 2, max(4, min(w, h_total) // 16)),)
    for cx, cy, rad in discs:
        for y in range(max(r0, cy - rad), min(r1, cy + rad + 1)):
            dy = abs(int(y) - cy)
            half = int(np.floor(np.sqrt(rad * rad - dy * dy)))
            m[y - r0, max(0, cx - half):cx + half + 1] = 1
    m[ys == 0] = 1
    m[ys == h_total - 1] = 1
    return r0, m


def workload_config(name, world, strong=None):
    """The `config` object of the JSON line: the workload only — identical in both arms (ours and --impl reference)."""
    w, rows_gpu, omega, u0, kind = WORKLOADS[name]
    strong = name in STRONG if strong is None else strong
    h_total = rows_gpu if strong else rows_gpu * world
    per_gpu = h_total // world
    resident = w * per_gpu * 18 * 4  # the two population lattices alone
    return {"workload": name, "W": w, "H": h_total, "H_per_gpu": per_gpu, "omega": omega, "u0": u0, "mask": kind,
            "l2": ("inputs larger than L2: %.1f GiB of populations per GPU vs 126 MB, no flush between steps"
                   % (resident / 2 ** 30)) if resident > 4 * L2_BYTES
            else "lattice fits L2 (launch-bound case, not a roofline case); no flush"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons of one GPU, stamped with the host clock; summary() keeps the samples
    that fall inside the timed windows."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=50):
        super().__init__(daemon=True)
        self.index, self.period_ms, self.rows, self._stop_evt = index, period_ms, [], threading.Event()
        self.windows, self._open = [], None

    def run(self):
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                  "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        while not self._stop_evt.is_set():
            line = p.stdout.readline()
            if not line:
                break
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))
        p.terminate()

    def stop(self):
        self._stop_evt.set()

    def begin(self):
        self._open = time.monotonic()

    def end(self):
        if self._open is not None:
            self.windows.append((self._open, time.monotonic()))
            self._open = None

    def summary(self, windows=None):
        windows = self.windows if windows is None else windows
        sm, mx, reasons, power = [], 0.0, set(), 0.0
        for t, r in self.rows:
            if not any(a <= t <= b + 0.03 for a, b in windows):
                continue
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                power = max(power, float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": power or None,
                "window_s": round(sum(b - a for a, b in windows), 3)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def gpu_numa_cpus(local):
    """CPUs of the NUMA node the GPU hangs off (pinned buffers allocated from there avoid the inter-socket hop),
    or None when the topology is not exposed."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None, None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        return (node, cpus) if cpus else (None, None)
    except Exception:
        return None, None


def cpu_baseline(w, omega, u0, kind, seconds=12.0):
    """The oracle (CPU restatement of the reference's 8-pass WGSL pipeline, OpenMP) on a strip of the same
    workload: same W, same mask generator, fewer rows."""
    from oracle.lbm_oracle import Oracle, threads
    rows = min(2048, max(16, (32 << 20) // w))
    _, m = mask_rows(kind, w, rows, 0, rows)
    o = Oracle(omega, w, rows, inflow_ux=u0)
    loc = np.flatnonzero(m.reshape(-1)).astype(np.uint32)
    o.draw_points(np.stack([loc, np.ones_like(loc)], 1))
    o.iterate(1)
    t0 = time.perf_counter()
    o.iterate(2)
    per = (time.perf_counter() - t0) / 2
    steps = int(max(3, min(200, seconds / max(per, 1e-6))))
    t0 = time.perf_counter()
    o.iterate(steps)
    dt = time.perf_counter() - t0
    o.close()
    return {"value": w * rows * steps / dt / 1e6, "unit": "MLUPS", "cores": threads(), "kind": "port",
            "sample": f"{w}x{rows} strip of the workload ({kind} mask), {steps} steps, "
                      f"oracle/lbm_oracle.c 8-pass structure, OpenMP"}


def run_reference(args):
    """--impl reference: the reference's own algorithm (its 8 passes per step, restated in C — the crate is
    Rust->wasm32 + WGSL and cannot run here) on all host cores.  Each step is one GPU's share of the workload
    (the full 16384 x 16384 lattice of the default workload) unless time or host memory bound it to a strip."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.lbm_oracle import Oracle, threads, use_all_cores
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        use_all_cores()  # torchrun exported OMP_NUM_THREADS=1; the other ranks have exited, rank 0 owns the host
    w, rows_gpu, omega, u0, kind = WORKLOADS[args.workload]
    if args.workload in STRONG or args.strong:
        rows_gpu //= max(1, args.gpus)
    # calibrate, then bound the sample so that (steps + warmup) steps take at most about four minutes
    calib_rows = max(16, (4 << 20) // w)
    o = Oracle(omega, w, calib_rows, inflow_ux=u0)
    o.iterate(1)
    t0 = time.perf_counter()
    o.iterate(2)
    rate = w * calib_rows * 2 / (time.perf_counter() - t0)
    o.close()
    warm = min(args.warmup, 3)
    budget_s = 240.0
    rows = int(rate * budget_s / max(1, args.steps + warm) / w)
    try:
        import psutil
        mem_rows = int(psutil.virtual_memory().available * 0.6) // (w * 100)  # ~92 B per cell in the oracle
    except Exception:
        mem_rows = (24 << 30) // (w * 100)
    rows = max(16, min(rows, rows_gpu, mem_rows))
    _, m = mask_rows(kind, w, rows, 0, rows)
    o = Oracle(omega, w, rows, inflow_ux=u0)
    loc = np.flatnonzero(m.reshape(-1)).astype(np.uint32)
    o.draw_points(np.stack([loc, np.ones_like(loc)], 1))
    del m, loc
    o.iterate(warm)
    t0 = time.perf_counter()
    o.iterate(args.steps)
    dt = time.perf_counter() - t0
    o.close()
    val = w * rows * args.steps / dt / 1e6
    sample = (f"{w}x{rows} rows of {args.workload} per step ({'one GPU share in full' if rows == rows_gpu else 'strip'}), "
              f"{args.steps} steps after {warm} warm-up; CPU restatement of the WGSL pipeline (8 passes/step), not lavapipe")
    print(json.dumps({
        "impl": "reference", "metric": "MLUPS (D2Q9 fp32)", "value": val, "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if (args.workload in STRONG or args.strong) else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus, args.workload in STRONG or args.strong),
        "implementation": {"kernel": "cpu oracle (8 passes per step)", "parallelism": f"{threads()} host threads"},
        "cpu_baseline": {"value": val, "unit": "MLUPS", "cores": threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------------------------------------------
class Job:
    """One process's view of the run: rank / world, the torch.distributed plumbing (start-up exchange of the IPC
    blobs, barriers, max over ranks) and the lattice factory."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.args = args
        self.single = args.single_process
        self.rank = 0 if self.single else int(os.environ.get("RANK", "0"))
        self.world = 1 if self.single else int(os.environ.get("WORLD_SIZE", "1"))
        self.local = 0 if self.single else int(os.environ.get("LOCAL_RANK", "0"))
        self.ngpu = args.gpus if self.single else self.world  # GPUs the lattice spans
        if not self.single and self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def barrier(self, lbm=None):
        if lbm is not None:
            lbm.synchronize()
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()

    def max_over_ranks(self, v):
        if not self.dist:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather(self, obj):
        if not self.dist:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def lattice(self, w, h_total, omega, u0, kind, discs=(), part=None):
        """The lattice W x h_total over this job's GPUs: rank r owns rows [r*h/n, (r+1)*h/n) (torchrun), or one
        group handle over all GPUs (--single-process).  part=(index, count) overrides the decomposition (count 1:
        the undivided lattice on this rank's GPU)."""
        from lbm_b200 import LBM, Kernel
        from lbm_b200.lbm import slab_rows
        args = self.args
        kernel = {"auto": Kernel.Auto, "scalar": Kernel.Scalar, "vec4": Kernel.Vec4}[args.kernel]
        lazy = None if args.lazy < 0 else args.lazy
        idx, cnt = part if part is not None else (self.rank, self.world)
        if self.single and part is None and self.ngpu > 1:
            lbm = LBM(omega, w, h_total, inflow_ux=u0, devices=list(range(self.ngpu)), kernel=kernel, lazy_barriers=lazy)
            ranges = slab_rows(h_total, self.ngpu)
            for q, (a, b) in enumerate(ranges):  # each slab takes its own window of the mask
                mr0, m = mask_rows(kind, w, h_total, a - 2, b + 2, discs)
                lbm.slab(q).write_barrier_rows(mr0, m)
            r0, r1 = 0, h_total
        else:
            r0, r1 = slab_rows(h_total, cnt)[idx]
            lbm = LBM(omega, w, h_total, inflow_ux=u0, device=self.local, rows=(r0, r1), kernel=kernel, lazy_barriers=lazy)
            if cnt > 1:
                blobs = self.gather(lbm.export_peer())
                if idx > 0:
                    lbm.link_peer(0, blobs[idx - 1])
                if idx < cnt - 1:
                    lbm.link_peer(1, blobs[idx + 1])
                self.dist.barrier()
            mr0, m = mask_rows(kind, w, h_total, r0 - 2, r1 + 2, discs)
            lbm.write_barrier_rows(mr0, m)
            del m
        if args.block_rows:
            lbm.set_tuning(0, args.block_rows)
        for knob, val in ((4, args.dense), (5, args.graphs), (6, args.packed), (7, args.index32), (8, args.link_in_kernel),
                          (9, args.pdl)):
            if val >= 0:
                lbm.set_tuning(knob, val)
        return lbm, r0, r1


def timed_steps(job, lbm, ncells_total, ncells_gpu, steps, warmup, sampler=None):
    """warm-up, then exactly `steps` steps timed on the device (max over ranks), then the same number of step
    launches alone between two events on the launching stream (the roofline figure)."""
    lbm.iterate(warmup)
    job.barrier(lbm)
    if sampler:
        sampler.begin()
    l0 = lbm.launch_count()
    ms = lbm.iterate_timed(steps)  # K fused step launches + the summary launch, CUDA events on our stream
    job.barrier(lbm)
    launches = lbm.launch_count() - l0
    ms = job.max_over_ranks(ms)
    lbm.timer_start()
    lbm.advance(steps)
    ms_kernel = lbm.timer_stop()
    job.barrier(lbm)
    if sampler:
        sampler.end()
    peak, peak_src = measured_peak_gbs()
    achieved = BYTES_PER_CELL * ncells_gpu * steps / (ms_kernel * 1e-3) / 1e9
    return {"ms": ms, "value": ncells_total * steps / (ms * 1e-3) / 1e6, "launches": launches,
            "ms_kernel": ms_kernel, "achieved": achieved, "peak": peak, "peak_src": peak_src}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8)).hexdigest()


def state_digests(lbm, row_ranges):
    """sha256 of all 18 populations, the three moments and the output field, per row range"""
    arrays = [lbm.read_population(k, b) for b in (0, 1) for k in range(9)]
    arrays += list(lbm.read_moments()) + [lbm.read_output()]
    r_base = lbm.row_begin
    return [[sha(a[r0 - r_base:r1 - r_base]) for a in arrays] for r0, r1 in row_ranges]


def multi_gpu_parity(job):
    """Before anything is timed: a verification lattice on the N linked slabs against the same lattice undivided
    on rank 0's GPU — porous mask plus a cylinder straddling every slab boundary, 400 steps (slabs of 256 rows step in
    ~50 us, so this is 400 back-to-back halo handshakes at the protocol's latency floor) with a paint and an erase
    across the boundaries on the way — compared slab by slab through sha256 of all 18 populations, the moments
    and the curl output ("2-GPU vs 1-GPU bit-identical populations")."""
    from lbm_b200.lbm import slab_rows
    n = job.ngpu
    w, rows = 16384, 256
    h = rows * n
    omega, u0 = 1.0, 0.05
    ranges = slab_rows(h, n)
    discs = [(w // 5 + 997 * q, ranges[q][1], 40 + 3 * q) for q in range(n - 1)]
    stroke_rows = [r1 + d for _, r1 in ranges[:-1] for d in (-2, -1, 0, 1)]
    loc = np.array([y * w + x for y in stroke_rows for x in (0, 1, 2, 5000, 5001, w - 2, w - 1)], np.uint64)
    paint = np.stack([loc, np.ones_like(loc)], 1)
    erase = np.stack([loc[::3], np.zeros_like(loc[::3])], 1)

    def run(lbm):
        lbm.iterate(150)
        lbm.draw_points(paint)
        lbm.iterate(137)
        lbm.draw_points(erase)
        lbm.iterate(113)

    lbm, r0, r1 = job.lattice(w, h, omega, u0, "porous", discs)
    run(lbm)
    chain = lbm.lazy_barriers_active()  # (before the read-backs: reading populations returns the table to the planes)
    mine = state_digests(lbm, ranges if job.single else [(r0, r1)])
    job.barrier(lbm)
    lbm.close()
    got = mine if job.single else [d[0] for d in job.gather(mine)]
    ok, bad = True, []
    if job.rank == 0:
        ref, _, _ = job.lattice(w, h, omega, u0, "porous", discs, part=(0, 1))
        run(ref)
        want = state_digests(ref, ranges)
        ref.close()
        names = [f"f{b}[{k}]" for b in (0, 1) for k in range(9)] + ["mx", "my", "rho", "out"]
        for q in range(n):
            bad += [f"slab{q}:{nm}" for nm, a, b in zip(names, got[q], want[q]) if a != b]
        ok = not bad
    if job.dist:
        flag = job.torch.tensor([1 if ok else 0], device="cuda")
        job.dist.broadcast(flag, 0)
        ok = bool(flag.item())
    res = {"bit_identical": ok, "cells": w * h, "W": w, "H": h, "slabs": n, "steps": 400, "paints": 2,
           "mask": "porous 15 % + one disc centred on every slab boundary",
           "compared": "sha256 of 18 populations + mx, my, rho + curl output per slab vs the undivided lattice on one GPU",
           "chain_table_active": bool(chain)}
    if bad:
        res["mismatch"] = bad[:12]
    return res


def field_digests(lbm, row_ranges):
    """sha256 of the density and of the curl output, per row range (full-size lattices: two planes, not twenty-two)"""
    rho = lbm.read_density()
    out = lbm.read_output()
    base = lbm.row_begin
    return [[sha(rho[a - base:b - base]), sha(out[a - base:b - base])] for a, b in row_ranges]


def time_workload(job, name, steps, warmup, sampler, part=None, digest_steps=0, digest_ranges=None):
    """one timed pass of a workload on this job's GPUs (or, part=(0, 1), on rank 0's GPU alone).  digest_steps > 0:
    after that many steps from the initial state, sha256 of density and curl per row range go into the result."""
    w, rows_gpu, omega, u0, kind = WORKLOADS[name]
    strong = name in STRONG
    n = job.ngpu if part is None else part[1]
    h_total = rows_gpu if strong else rows_gpu * n
    lbm, r0, r1 = job.lattice(w, h_total, omega, u0, kind, part=part)
    ncells_gpu = w * (h_total // n)
    digests = None
    if digest_steps:
        lbm.iterate(digest_steps)
        digests = field_digests(lbm, digest_ranges if digest_ranges is not None else [(r0, r1)])
    w0 = len(sampler.windows)
    if part is None:
        t = timed_steps(job, lbm, w * h_total, ncells_gpu, steps, warmup, sampler)
    else:
        solo = Job.__new__(Job)  # a view of this job without the collective plumbing
        solo.__dict__.update(job.__dict__, dist=None)
        t = timed_steps(solo, lbm, w * h_total, ncells_gpu, steps, warmup, sampler)
    lbm.synchronize()
    lbm.close()
    res = {"workload": name, "W": w, "H": h_total, "n_gpus": n, "steps": steps, "warmup": warmup,
           "value": t["value"], "unit": "MLUPS", "ms_per_step": t["ms"] / steps,
           "frac": t["achieved"] / t["peak"], "achieved_gbs_per_gpu": t["achieved"],
           "clocks": sampler.summary(sampler.windows[w0:])}
    if digests is not None:
        res["_digests"] = digests
    return res


def north_star(job, sampler):
    """The north star's own multi-GPU configurations, timed in the same run as the headline workload:
    configs[4] channel65536 (65536 x 8192 per GPU: the 65536^2 lattice at N = 8), 50 steps, and configs[3]
    cylinder32768 strong scaling (rank 0 alone first for T1, then all N slabs), 100 steps."""
    out = {}
    ws = time_workload(job, "channel65536", 50, 5, sampler)
    ws["scaling"] = "weak (65536 x 8192 per GPU)"
    out["channel65536"] = ws
    job.barrier()
    # parity at full size: after 20 steps from the initial state, density and curl of the N slabs against the same rows
    # of the undivided 32768^2 lattice (1.07 G cells) on rank 0's GPU
    from lbm_b200.lbm import slab_rows
    h32 = WORKLOADS["cylinder32768"][1]
    ranges = slab_rows(h32, job.ngpu)
    t1 = None
    if job.rank == 0:
        t1 = time_workload(job, "cylinder32768", 40, 4, sampler, part=(0, 1), digest_steps=20, digest_ranges=ranges)
    job.barrier()
    tn = time_workload(job, "cylinder32768", 100, 10, sampler, digest_steps=20,
                       digest_ranges=ranges if job.single else None)
    t1 = job.gather(t1)[0]
    got = tn.pop("_digests")
    got = got if job.single else [d[0] for d in job.gather(got)]
    want = t1.pop("_digests")
    tn["parity_at_size"] = {"bit_identical": got == want, "cells": WORKLOADS["cylinder32768"][0] * h32, "steps": 20,
                            "compared": "sha256 of density and curl output per slab vs the undivided lattice on one GPU"}
    tn["scaling"] = "strong"
    tn["t1_ms_per_step"] = t1["ms_per_step"]
    tn["t1_value"] = t1["value"]
    tn["efficiency"] = t1["ms_per_step"] / (job.ngpu * tn["ms_per_step"])
    out["cylinder32768"] = tn
    return out


def run_e2e(job, lbm, w, r0, r1, ncells_total, args, sampler):
    """End to end through the C ABI with host buffers: every frame = paint a stroke (H2D from pinned host memory),
    blbm_iterate(n) (n steps + summary), read the output field back (D2H into pinned host memory), as
    lib.rs:108-199 does per redraw.  The number of frames does not depend on --steps: at least 16, so that the
    double-buffered read-back is in its steady state and the drain of the last copy (which nothing can hide) weighs
    what it would in a running application."""
    torch = job.torch
    rows = r1 - r0
    fs = max(1, args.frame_steps)
    frames = max(16, args.steps // fs)
    node, cpus = gpu_numa_cpus(job.local)
    keep_affinity = os.sched_getaffinity(0)
    if cpus and not job.single:
        os.sched_setaffinity(0, cpus)  # first touch of the pinned buffers on the GPU's own NUMA node
    out_host = [torch.empty((rows, w), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    for t in out_host:
        t.zero_()
    stroke = torch.empty((64, 2), dtype=torch.int64).pin_memory()
    os.sched_setaffinity(0, keep_affinity)
    y_mid = (r0 + r1) // 2
    loc = np.array([(y_mid + j // 8) * w + (w // 2 + j % 8) for j in range(64)], dtype=np.int64)
    stroke[:, 0] = torch.from_numpy(loc)
    stroke_np = stroke.numpy().view(np.uint64)
    L = lbm._L
    out_bytes = out_host[0].numel() * 4

    # the host link's ceiling: the same D2H volume from every GPU at once, nothing else running
    ceil_ms = d2h_ceiling_gbs = None
    if not (job.single and job.ngpu > 1):
        dev = torch.empty((rows, w), dtype=torch.float32, device="cuda")
        dev.zero_()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        out_host[0].copy_(dev, non_blocking=True)
        job.barrier()
        ev[0].record()
        for q in range(3):
            out_host[q & 1].copy_(dev, non_blocking=True)
        ev[1].record()
        torch.cuda.synchronize()
        ceil_ms = job.max_over_ranks(ev[0].elapsed_time(ev[1]) / 3)
        del dev
        d2h_ceiling_gbs = job.world * out_bytes / (ceil_ms * 1e-3) / 1e9

    def frame(fr):
        stroke_np[:, 1] = fr & 1 ^ 1  # draw the blob, erase it next frame
        assert L.blbm_draw_points64(lbm._h, stroke_np.ctypes.data, 64) == 0
        lbm.iterate(fs)
        # asynchronous read-back into alternating pinned buffers: the copy of frame f overlaps the steps
        # of frame f+1 (a renderer would consume buffer f%2 while frame f+1 computes)
        assert L.blbm_read_output_async(lbm._h, out_host[fr & 1].data_ptr()) == 0

    # one frame alone: the latency a caller sees from the paint to the field in host memory (after one untimed frame:
    # the first paint of a handle allocates its device list and rebuilds the class words of the whole lattice)
    frame(0)
    job.barrier(lbm)
    lbm.timer_start()
    frame(1)
    first_ms = job.max_over_ranks(lbm.timer_stop())
    job.barrier(lbm)
    sampler.begin()
    lbm.timer_start()
    for fr in range(frames):
        frame(fr)
    ms_e2e = lbm.timer_stop()
    job.barrier(lbm)
    sampler.end()
    ms_e2e = job.max_over_ranks(ms_e2e)
    nsteps = frames * fs
    d2h_gbs = job.world * out_bytes * frames / (ms_e2e * 1e-3) / 1e9
    return {"value": ncells_total * nsteps / (ms_e2e * 1e-3) / 1e6, "unit": "MLUPS",
            "h2d_bytes_per_step": stroke_np.nbytes * frames / nsteps, "d2h_bytes_per_step": out_bytes * frames / nsteps,
            "frames": frames, "steps_per_frame": fs, "nsteps": nsteps, "ms_per_frame": ms_e2e / frames,
            "first_frame_ms": first_ms,
            "d2h_ceiling_gbs": d2h_ceiling_gbs, "d2h_achieved_gbs": d2h_gbs,
            "d2h_ceiling_what": f"{job.world} concurrent cudaMemcpyAsync D2H of {out_bytes / 2**30:.2f} GiB each into "
                                f"pinned host memory, nothing else running (whole-node aggregate)",
            "d2h_ms_per_frame_at_ceiling": ceil_ms,
            "pinned_numa_node": node,
            "what": "per frame: blbm_draw_points64(64-point stroke, pinned host) + blbm_iterate(n) + "
                    "blbm_read_output_async(W*H fp32 to pinned host, double-buffered); no host synchronisation inside "
                    "the loop; the stopwatch stops after the last copy has landed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--warmup", type=int, default=15)
    ap.add_argument("--impl", default="blbm", choices=["blbm", "reference"])
    ap.add_argument("--workload", default="porous16384", choices=sorted(WORKLOADS))
    ap.add_argument("--kernel", default="auto", choices=["auto", "scalar", "vec4"])
    ap.add_argument("--single-process", action="store_true",
                    help="one process drives all --gpus devices through one blbm_create_group handle")
    ap.add_argument("--frame-steps", type=int, default=15, help="steps per frame of the e2e loop (lib.rs:17)")
    ap.add_argument("--block-rows", type=int, default=0, help="vec4 kernel rows per block (4, 8, 16); 0 = default")
    ap.add_argument("--graphs", type=int, default=-1, help="CUDA graphs for the step loop: -1 auto, 0 off, 1 on")
    ap.add_argument("--dense", type=int, default=-1, help="vec4 bounce flavour: -1 auto, 0 sparse, 1 dense, 2 dense + cp.async staging")
    ap.add_argument("--packed", type=int, default=-1, help="vec4 packed fp32 adds (FADD2): -1 default (off: measured slower), 0, 1")
    ap.add_argument("--index32", type=int, default=-1, help="vec4 32-bit plane offsets: -1 auto, 0, 1")
    ap.add_argument("--link-in-kernel", type=int, default=-1,
                    help="linked slabs: halo epoch handshake inside the step kernel (1, default) or by wait/signal kernels (0)")
    ap.add_argument("--pdl", type=int, default=-1,
                    help="programmatic dependent launch of the fused step: -1 auto (on outside CUDA graphs), 0, 1")
    ap.add_argument("--lazy", type=int, default=-1, help="barrier-chain table: 0 never, 1 always, 2 auto (default)")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling of any workload: the lattice of the N = 1 case is split into N slabs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU parity check (N > 1)")
    ap.add_argument("--north-star", type=int, default=-1,
                    help="also time channel65536 and cylinder32768-strong: -1 auto (on at N = 8 with the default "
                         "workload), 0 off, 1 on")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return

    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        args.single_process = True  # launched without torchrun: one process drives all N GPUs through a group handle
    job = Job(args)
    n = job.ngpu
    w, rows_gpu, omega, u0, kind = WORKLOADS[args.workload]
    strong = args.workload in STRONG or args.strong
    h_total = rows_gpu if strong else rows_gpu * n

    parity = None
    if n > 1 and not args.no_parity:
        parity = multi_gpu_parity(job)
        if not parity["bit_identical"]:
            if job.rank == 0:
                print(json.dumps({"error": "multi-GPU parity check failed", "multi_gpu_parity": parity}))
            raise SystemExit(3)

    sampler = ClockSampler(job.local)
    sampler.start()
    lbm, r0, r1 = job.lattice(w, h_total, omega, u0, kind)
    ncells_gpu = w * (h_total // n)
    ncells = w * h_total
    t = timed_steps(job, lbm, ncells, ncells_gpu, args.steps, args.warmup, sampler)
    kname = lbm.get_kernel().name.lower()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get(args.workload, {}).get(kname)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": t["achieved"], "peak": t["peak"], "unit": "GB/s",
                "frac": t["achieved"] / t["peak"], "traffic": traffic, "peak_source": t["peak_src"],
                "frac_of_nominal_8tbs": t["achieved"] / 8000.0,
                "kernel": f"step_{kname}_kernel", "bytes_per_cell": BYTES_PER_CELL, "cells_per_launch": ncells_gpu,
                "avg_launch_ms": t["ms_kernel"] / args.steps}
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(job, lbm, w, r0, r1, ncells, args, sampler)
    device_bytes = lbm.device_bytes()
    lbm.synchronize()
    job.barrier()
    lbm.close()

    line = {
        "metric": "MLUPS (D2Q9 fp32)", "value": t["value"], "unit": "MLUPS", "n_gpus": n, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t["ms"] / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, n, strong),
        "implementation": {"kernel": kname, "halo_handshake": (None if n == 1 else "wait/signal kernels" if args.link_in_kernel == 0
                                                               else "inside the step kernel"),
                           "parallelism": (f"y-slabs x{n}, one blbm_create_group handle in one process" if job.single
                                           else f"y-slabs x{n}, one process per GPU (CUDA IPC peers)"),
                           "device_bytes_per_gpu": device_bytes // (n if job.single else 1)},
        "roofline": roofline, "e2e": e2e, "gpu_launches": t["launches"], "clocks": sampler.summary(),
    }
    if parity is not None:
        line["multi_gpu_parity"] = parity
    want_ns = args.north_star == 1 or (args.north_star < 0 and n == 8 and args.workload == "porous16384")
    if want_ns and n > 1:
        line["north_star"] = north_star(job, sampler)
        if not line["north_star"]["cylinder32768"]["parity_at_size"]["bit_identical"]:
            if job.rank == 0:
                print(json.dumps({"error": "multi-GPU parity at full size failed", "north_star": line["north_star"]}))
            raise SystemExit(3)
    sampler.stop()
    if job.rank == 0 and n == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(w, omega, u0, kind)
    if job.rank == 0:
        print(json.dumps(line))
    if job.dist:
        job.dist.destroy_process_group()


if __name__ == "__main__":
    main()
