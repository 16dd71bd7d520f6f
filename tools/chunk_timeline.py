"""Kernel timeline (CUPTI through torch.profiler) of a few back-to-back small dcn_filter_batch_device calls."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deacon_server_b200 as d  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(1)
gpu = d.DeaconGpu(0)
keys = torch.randint(-2**63, 2**63 - 1, (380_000_000,), dtype=torch.int64, device=dev)
gpu.index_upload_device(keys, d.IndexHeader(2, 31, 15))
del keys
pairs = int(os.environ.get("PAIRS", "56000"))
calls = int(os.environ.get("CALLS", "8"))
NP = pairs * calls
NR = 2 * NP
lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
bases = lut[torch.randint(0, 4, (NR * 150,), device=dev)]
off = torch.arange(NR + 1, device=dev, dtype=torch.int64) * 150
keep = torch.zeros(NP, dtype=torch.uint8, device=dev)
hits = torch.zeros(NP, dtype=torch.int32, device=dev)
tot = torch.zeros(NP, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream


def run():
    for c in range(calls):
        u0 = c * pairs
        gpu.filter_batch_device(bases[2 * u0 * 150:], off[:2 * pairs + 1], 2 * pairs, 2 * pairs * 150, keep[u0:], hits[u0:], tot[u0:],
                                paired=True, deplete=True, stream=st, max_unit_len=300)


run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
prev_end = t0
for e in ev:
    print(f"{(e.time_range.start - t0):9.1f} us  +{e.time_range.start - prev_end:6.1f} gap  {e.time_range.end - e.time_range.start:8.1f} us  {e.name[:70]}")
    prev_end = e.time_range.end
