"""GPU cost of one pipeline chunk as a function of its size: back-to-back dcn_filter_batch_device calls on one stream
over slices of a resident batch (what a stage of the host-pointer pipeline enqueues, less the copies)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deacon_server_b200 as d  # noqa: E402

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
bases = lut[torch.randint(0, 4, (nb,), device=dev)]
off = torch.arange(NR + 1, device=dev, dtype=torch.int64) * 150
keep = torch.zeros(NP, dtype=torch.uint8, device=dev)
hits = torch.zeros(NP, dtype=torch.int32, device=dev)
tot = torch.zeros(NP, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
hint = int(os.environ.get("HINT", "300"))
SIZES = [int(x) for x in os.environ.get("PAIRS", "14000,28000,56000,112000,224000,448000,5000000").split(",")]
for pairs in SIZES:
    n_calls = min(NP // pairs, int(os.environ.get("CALLS", "100000")))

    def run():
        for c in range(n_calls):
            u0 = c * pairs
            gpu.filter_batch_device(bases[2 * u0 * 150:], off[:2 * pairs + 1], 2 * pairs, 2 * pairs * 150, keep[u0:], hits[u0:], tot[u0:],
                                    paired=True, deplete=True, stream=st, max_unit_len=hint)
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"chunk {pairs * 300 / 1e6:7.1f} Mbp x {n_calls:4d} calls: {ms / n_calls * 1e3:8.1f} us per chunk, {n_calls * pairs * 300 / ms / 1e6:7.1f} Gbp/s", flush=True)
