#!/usr/bin/env python
"""Does peer traffic between two B200s of the box ride NVLink?  NVML's per-link counters read N/A on this pool and ncu
cannot collect nvltx/nvlrx on the linked kernel, so the evidence is indirect: the bandwidth of a device-to-device copy
between GPU 0 and GPU 1 (PCIe Gen5 x16 tops out near 55 GB/s; NVLink 5 at several hundred), the peer attributes CUDA
reports, and the latency of a remote 8-byte store + flag round trip as the halo handshake does it."""
import json
import time

import torch

assert torch.cuda.device_count() >= 2
res = {"can_access_peer": torch.cuda.can_device_access_peer(0, 1)}
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda:0")
b = torch.empty(n, dtype=torch.uint8, device="cuda:1")
a.fill_(1)
b.fill_(0)
for direction, (src, dst) in {"0->1": (a, b), "1->0": (b, a)}.items():
    dst.copy_(src)
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    dev = src.device
    with torch.cuda.device(dev):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dst.copy_(src, non_blocking=True)
        e1.record()
        e1.synchronize()
        res[f"copy_{direction}_gbs"] = round(10 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)
# both directions at once
torch.cuda.synchronize(0)
torch.cuda.synchronize(1)
c = torch.empty(n, dtype=torch.uint8, device="cuda:0")
d = torch.empty(n, dtype=torch.uint8, device="cuda:1")
t0 = time.perf_counter()
s0, s1 = torch.cuda.Stream(device=0), torch.cuda.Stream(device=1)
for _ in range(10):
    with torch.cuda.stream(s0):
        b.copy_(a, non_blocking=True)
    with torch.cuda.stream(s1):
        c.copy_(d, non_blocking=True)
torch.cuda.synchronize(0)
torch.cuda.synchronize(1)
res["copy_both_directions_gbs_total"] = round(20 * n / (time.perf_counter() - t0) / 1e9, 1)
print(json.dumps(res))
