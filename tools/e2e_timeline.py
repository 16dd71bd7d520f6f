"""Device-side accounting of one dcn_filter_batch call (CUPTI through torch.profiler): busy time per kind of work
(H2D copies, D2H copies, kernels by name) inside the call's window, and how much of the window each kind covers."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deacon_server_b200 as d  # noqa: E402

threads = int(os.environ.get("THREADS", "12"))
dev = torch.device("cuda:0")
torch.manual_seed(1)
gpu = d.DeaconGpu(0)
keys = torch.randint(-2**63, 2**63 - 1, (380_000_000,), dtype=torch.int64, device=dev)
gpu.index_upload_device(keys, d.IndexHeader(2, 31, 15))
del keys
NP = 5_000_000
NR = 2 * NP
nb = NR * 150
lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
hbases = lut[torch.randint(0, 4, (nb,), device=dev)].cpu().pin_memory()
hoff = (torch.arange(NR + 1, dtype=torch.int64) * 150).pin_memory()
hk = torch.zeros(NP, dtype=torch.uint8).pin_memory()
hh = torch.zeros(NP, dtype=torch.int32).pin_memory()
ht = torch.zeros(NP, dtype=torch.int32).pin_memory()
gpu.host_pack_threads(threads)


def e2e():
    gpu.filter_batch_ptr(hbases.data_ptr(), hoff.data_ptr(), NR, True, 0, 2, 0.01, True, hk.data_ptr(), hh.data_ptr(), ht.data_ptr())


for _ in range(3):
    e2e()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    e2e()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
t0 = min(e.time_range.start for e in ev)
t1 = max(e.time_range.end for e in ev)
print(f"window {(t1 - t0) / 1e3:.2f} ms, {len(ev)} device activities")


def kind(e):
    n = e.name
    if n.startswith("Memcpy HtoD"):
        return "H2D"
    if n.startswith("Memcpy DtoH"):
        return "D2H"
    if n.startswith("Memset"):
        return "memset"
    return n.split("(")[0].replace("void ", "").replace("dcn::", "")[:40]


def union(iv):
    iv = sorted(iv)
    tot, cs, ce = 0.0, None, None
    for a, b in iv:
        if cs is None or a > ce:
            if cs is not None:
                tot += ce - cs
            cs, ce = a, b
        else:
            ce = max(ce, b)
    if cs is not None:
        tot += ce - cs
    return tot


groups = {}
for e in ev:
    groups.setdefault(kind(e), []).append((e.time_range.start, e.time_range.end))
for k, iv in sorted(groups.items(), key=lambda kv: -sum(b - a for a, b in kv[1])):
    s = sum(b - a for a, b in iv)
    print(f"{k:42s} n={len(iv):4d}  sum {s / 1e3:7.2f} ms  covered {union(iv) / 1e3:7.2f} ms  mean {s / len(iv):7.1f} us")
kern = [iv for k, v in groups.items() if k not in ("H2D", "D2H", "memset") for iv in v]
print(f"all kernels: sum {sum(b - a for a, b in kern) / 1e3:.2f} ms, covered {union(kern) / 1e3:.2f} ms")
# idle gaps of the copy engine and of the SMs over the window, in 1 ms bins
import math
nbin = int(math.ceil((t1 - t0) / 1e3))
for name, iv in (("H2D", groups.get("H2D", [])), ("kernels", kern)):
    bins = [0.0] * nbin
    for a, b in iv:
        x = a
        while x < b:
            i = int((x - t0) // 1e3)
            e_ = min(b, t0 + (i + 1) * 1e3)
            bins[i] += e_ - x
            x = e_
    print(name, "busy us per ms bin (sum over concurrent):", " ".join(f"{v:.0f}" for v in bins))
