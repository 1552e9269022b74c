#!/usr/bin/env python
"""bench.py — MLUPS of the D2Q9 lattice update on N B200s, with the HBM roofline and the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--kernel NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's algorithm on the host cores (oracle)

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
halo stores from the step kernel; no NCCL collective is on the data path.
"""
import argparse
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


def mask_rows(kind, w, h_total, r0, r1):
    """uint8 barrier mask of global rows [r0, r1): walls on rows 0 and H-1 plus the workload's obstacles."""
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
        cx, cy, rad = w // 4, h_total // 2, max(4, min(w, h_total) // 16)
        for j, y in enumerate(ys):
            dy = abs(int(y) - cy)
            if dy <= rad:
                half = int(np.floor(np.sqrt(rad * rad - dy * dy)))
                m[j, cx - half:cx + half + 1] = 1
    m[ys == 0] = 1
    m[ys == h_total - 1] = 1
    return r0, m


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons of one GPU while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                  "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                 stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        while not self._stop_evt.is_set():
            line = p.stdout.readline()
            if not line:
                break
            self.rows.append([c.strip() for c in line.split(",")])
        p.terminate()

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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
    Rust->wasm32 + WGSL and cannot run here) on all host cores.  Each step is a bounded strip of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.lbm_oracle import Oracle, threads, use_all_cores
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        use_all_cores()  # torchrun exported OMP_NUM_THREADS=1; the other ranks have exited, rank 0 owns the host
    w, rows_gpu, omega, u0, kind = WORKLOADS[args.workload]
    # size the strip so that (steps + warmup) steps take about two minutes
    calib_rows = max(16, (4 << 20) // w)
    o = Oracle(omega, w, calib_rows, inflow_ux=u0)
    o.iterate(1)
    t0 = time.perf_counter()
    o.iterate(2)
    rate = w * calib_rows * 2 / (time.perf_counter() - t0)
    o.close()
    budget_s = 100.0
    rows = int(rate * budget_s / max(1, args.steps + args.warmup) / w)
    rows = max(16, min(rows, rows_gpu * args.gpus, (24 << 30) // (w * 100)))
    _, m = mask_rows(kind, w, rows, 0, rows)
    o = Oracle(omega, w, rows, inflow_ux=u0)
    loc = np.flatnonzero(m.reshape(-1)).astype(np.uint32)
    o.draw_points(np.stack([loc, np.ones_like(loc)], 1))
    o.iterate(args.warmup)
    t0 = time.perf_counter()
    o.iterate(args.steps)
    dt = time.perf_counter() - t0
    o.close()
    val = w * rows * args.steps / dt / 1e6
    sample = (f"{w}x{rows} strip of {args.workload} per step, {args.steps} steps; CPU restatement of the WGSL "
              f"pipeline (8 passes/step), not lavapipe")
    print(json.dumps({
        "impl": "reference", "metric": "MLUPS (D2Q9 fp32)", "value": val, "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.workload in STRONG else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "W": w, "H": rows_gpu if args.workload in STRONG else rows_gpu * args.gpus,
                   "H_per_gpu": rows_gpu // args.gpus if args.workload in STRONG else rows_gpu, "omega": omega,
                   "u0": u0, "mask": kind, "kernel": "cpu oracle (8 passes per step)",
                   "parallelism": f"{threads()} host threads"},
        "cpu_baseline": {"value": val, "unit": "MLUPS", "cores": threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--warmup", type=int, default=15)
    ap.add_argument("--impl", default="blbm", choices=["blbm", "reference"])
    ap.add_argument("--workload", default="porous16384", choices=sorted(WORKLOADS))
    ap.add_argument("--kernel", default="auto", choices=["auto", "scalar", "vec4", "tma"])
    ap.add_argument("--frame-steps", type=int, default=15, help="steps per frame of the e2e loop (lib.rs:17)")
    ap.add_argument("--block-rows", type=int, default=0, help="vec4 kernel rows per block (4, 8, 16); 0 = default")
    ap.add_argument("--graphs", type=int, default=-1, help="CUDA graphs for the step loop: -1 auto, 0 off, 1 on")
    ap.add_argument("--dense", type=int, default=-1, help="vec4 bounce flavour: -1 auto, 0 sparse, 1 dense, 2 dense + cp.async staging")
    ap.add_argument("--packed", type=int, default=-1, help="vec4 packed fp32 adds (FADD2): -1 default (off: measured slower), 0, 1")
    ap.add_argument("--index32", type=int, default=-1, help="vec4 32-bit plane offsets: -1 auto, 0, 1")
    ap.add_argument("--tma-rows", type=int, default=0)
    ap.add_argument("--tma-stages", type=int, default=0)
    ap.add_argument("--tma-ctas", type=int, default=0)
    ap.add_argument("--lazy", type=int, default=-1, help="barrier-chain table: 0 never, 1 always, 2 auto (default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from lbm_b200 import LBM, Kernel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w, rows_gpu, omega, u0, kind = WORKLOADS[args.workload]
    strong = args.workload in STRONG
    if strong:
        h_total = rows_gpu
        rows_gpu = h_total // world
    else:
        h_total = rows_gpu * world
    r0, r1 = rank * rows_gpu, (rank + 1) * rows_gpu
    kernel = {"auto": Kernel.Auto, "scalar": Kernel.Scalar, "vec4": Kernel.Vec4, "tma": Kernel.Tma}[args.kernel]
    lbm = LBM(omega, w, h_total, inflow_ux=u0, device=local, rows=(r0, r1), kernel=kernel,
              lazy_barriers=None if args.lazy < 0 else args.lazy)
    if args.block_rows:
        lbm.set_tuning(0, args.block_rows)
    if args.dense >= 0:
        lbm.set_tuning(4, args.dense)
    if args.graphs >= 0:
        lbm.set_tuning(5, args.graphs)
    if args.packed >= 0:
        lbm.set_tuning(6, args.packed)
    if args.index32 >= 0:
        lbm.set_tuning(7, args.index32)
    for knob, val in ((1, args.tma_rows), (2, args.tma_stages), (3, args.tma_ctas)):
        if val:
            lbm.set_tuning(knob, val)
    if world > 1:
        blobs = [None] * world
        dist.all_gather_object(blobs, lbm.export_peer())
        if rank > 0:
            lbm.link_peer(0, blobs[rank - 1])
        if rank < world - 1:
            lbm.link_peer(1, blobs[rank + 1])
        dist.barrier()
    mr0, m = mask_rows(kind, w, h_total, r0 - 2, r1 + 2)
    lbm.write_barrier_rows(mr0, m)
    del m
    ncells_gpu = w * rows_gpu
    ncells = w * h_total

    def barrier():
        lbm.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up, then exactly K steps timed on the device, max over ranks -------------------------------
    lbm.iterate(args.warmup)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lbm.launch_count()
    ms = lbm.iterate_timed(args.steps)  # K fused step launches + the summary launch, CUDA events on our stream
    barrier()
    launches = lbm.launch_count() - launches0
    ms = max_over_ranks(ms)
    value = ncells * args.steps / (ms * 1e-3) / 1e6

    # ---- the step kernel alone (roofline): K launches between two events on the launching stream ------------
    barrier()
    lbm.timer_start()
    lbm.advance(args.steps)
    ms_kernel = lbm.timer_stop()
    barrier()
    sampler.stop()
    peak, peak_src = measured_peak_gbs()
    achieved = BYTES_PER_CELL * ncells_gpu * args.steps / (ms_kernel * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get(args.workload, {}).get(lbm.get_kernel().name.lower())
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": f"step_{lbm.get_kernel().name.lower()}_kernel",
                "bytes_per_cell": BYTES_PER_CELL, "cells_per_launch": ncells_gpu,
                "avg_launch_ms": ms_kernel / args.steps}

    # ---- end to end through the C ABI with host buffers: every frame = paint a stroke (H2D), iterate(n)
    #      (n steps + summary), read the output field back (D2H), as lib.rs:108-199 does per redraw ----------
    e2e = None
    if not args.no_e2e:
        fs = max(1, min(args.frame_steps, args.steps))
        frames = max(1, args.steps // fs)
        out_host = [torch.empty((rows_gpu, w), dtype=torch.float32, pin_memory=True) for _ in range(2)]
        stroke = torch.empty((64, 2), dtype=torch.int64).pin_memory()
        y_mid = (r0 + r1) // 2
        loc = np.array([(y_mid + j // 8) * w + (w // 2 + j % 8) for j in range(64)], dtype=np.int64)
        stroke[:, 0] = torch.from_numpy(loc)
        stroke_np = stroke.numpy().view(np.uint64)
        C = __import__("ctypes")
        L = lbm._L
        h2d = d2h = 0
        barrier()
        lbm.timer_start()
        for fr in range(frames):
            stroke_np[:, 1] = fr & 1 ^ 1  # draw the blob, erase it next frame
            rc = L.blbm_draw_points64(lbm._h, stroke_np.ctypes.data, 64)
            assert rc == 0
            h2d += stroke_np.nbytes
            lbm.iterate(fs)
            # asynchronous read-back into alternating pinned buffers: the copy of frame f overlaps the steps
            # of frame f+1 (a renderer would consume buffer f%2 while frame f+1 computes)
            rc = L.blbm_read_output_async(lbm._h, out_host[fr & 1].data_ptr())
            assert rc == 0
            d2h += out_host[0].numel() * 4
        ms_e2e = lbm.timer_stop()
        barrier()
        ms_e2e = max_over_ranks(ms_e2e)
        nsteps = frames * fs
        e2e = {"value": ncells * nsteps / (ms_e2e * 1e-3) / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": h2d / nsteps, "d2h_bytes_per_step": d2h / nsteps,
               "frames": frames, "steps_per_frame": fs,
               "what": "per frame: blbm_draw_points64(64-point stroke, pinned host) + blbm_iterate(n) + "
                       "blbm_read_output_async(W*H fp32 to pinned host, double-buffered); the stopwatch stops after the "
                       "last copy has landed"}

    line = {
        "metric": "MLUPS (D2Q9 fp32)", "value": value, "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "W": w, "H": h_total, "H_per_gpu": rows_gpu, "omega": omega, "u0": u0,
                   "mask": kind, "kernel": lbm.get_kernel().name.lower(), "parallelism": f"y-slabs x{world}",
                   "l2": f"working set {lbm.device_bytes() / 2**30:.1f} GiB per GPU >> 126 MB L2 (no flush needed)"
                   if lbm.device_bytes() > (1 << 30) else "working set fits L2: launch-bound, not a roofline case"},
        "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary(),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(w, omega, u0, kind)
    if rank == 0:
        print(json.dumps(line))
    lbm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
