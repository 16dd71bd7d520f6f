"""Host-ingest sweep for N ranks on one box (run under torchrun, or alone for N = 1): what do packer threads buy
when the ranks share the host?  Every rank filters its own 1.5 Gbp batch of 2x150 pairs from pinned host memory
through dcn_filter_batch with `t` packer threads, for each t in --threads; the slowest rank's time counts (as in
bench.py's e2e leg).  Also: concurrent pinned H2D rate (parallel.h2d_probe), packing-only rate of t threads per rank
with no copy running, and the caller-packed form.  One JSON line per setting on rank 0; `--out` appends them to a file.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/ingest_sweep.py --threads 0,1,2,3 --out gpurun_out/sweep8.jsonl
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deacon_server_b200 as d  # noqa: E402
from deacon_server_b200 import parallel as par  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--threads", default="0,1,2,3,4")
ap.add_argument("--pairs-m", type=float, default=5)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--keys-m", type=float, default=380)
ap.add_argument("--out", default="")
ap.add_argument("--no-bind", action="store_true")
args = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
bound = par.bind_to_gpu(local) if world > 1 and not args.no_bind else []
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def max_over_ranks(x):
    v = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return float(v.item())


def sum_over_ranks(x):
    v = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
    return float(v.item())


def emit(obj):
    if rank == 0:
        line = json.dumps(obj)
        print(line, flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(line + "\n")


emit({"what": "box", "world": world, "cpus": os.cpu_count(), "allowed": len(os.sched_getaffinity(0)), "bound": len(bound)})

# ---- index: random keys (the probe pattern of a real index: uniform), reads: random bases (host ingest does not
# depend on the hit rate; the device-resident kernel is measured by bench.py)
torch.manual_seed(1 + rank)
gpu = d.DeaconGpu(local)
keys = torch.randint(-2**63, 2**63 - 1, (int(args.keys_m * 1e6),), dtype=torch.int64, device=dev)
gpu.index_upload_device(keys, d.IndexHeader(2, 31, 15))
del keys
NP = int(args.pairs_m * 1e6)
NR = 2 * NP
nb = NR * 150
lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
bases = lut[torch.randint(0, 4, (nb,), device=dev)]
off = torch.arange(NR + 1, device=dev, dtype=torch.int64) * 150
keep = torch.zeros(NP, dtype=torch.uint8, device=dev)
hits = torch.zeros(NP, dtype=torch.int32, device=dev)
tot = torch.zeros(NP, dtype=torch.int32, device=dev)
gpu.filter_batch_device(bases, off, NR, nb, keep, hits, tot, paired=True, deplete=True,
                        stream=torch.cuda.current_stream().cuda_stream, max_unit_len=300)
torch.cuda.synchronize()
hbases = bases.cpu().pin_memory()
hoff = off.cpu().pin_memory()
hk = torch.zeros(NP, dtype=torch.uint8).pin_memory()
hh = torch.zeros(NP, dtype=torch.int32).pin_memory()
ht = torch.zeros(NP, dtype=torch.int32).pin_memory()
del bases

probe = par.h2d_probe(dev)
emit({"what": "h2d_probe", **probe})
# every rank's share with all ranks copying (the ranks of one host do not get equal shares)
shares = [probe["concurrent_gbs"]]
if world > 1:
    t = torch.tensor([probe["concurrent_gbs"]], dtype=torch.float64, device=dev)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    shares = [round(float(x.item()), 2) for x in parts]
emit({"what": "h2d_share_per_rank", "gbs": shares})

# ---- packing alone: t threads per rank run dcn_pack_ascii over slices of the batch, every rank at once
lib = d.load()
codes = torch.empty(nb // 4 + 64, dtype=torch.uint8).pin_memory()
inv = torch.empty(nb // 8 + 64, dtype=torch.uint8).pin_memory()


def pack_rate(t, seconds=0.6):
    per = (nb // t) & ~63
    done = [0] * t
    stop = [False]

    def worker(i):
        b = hbases.data_ptr() + i * per
        c = codes.data_ptr() + i * per // 4
        v = inv.data_ptr() + i * per // 8
        while not stop[0]:
            lib.dcn_pack_ascii(C.c_void_p(b), C.c_uint64(per), C.c_void_p(c), C.c_void_p(v))
            done[i] += per

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(t)]
    barrier()
    t0 = time.perf_counter()
    for th in ts:
        th.start()
    time.sleep(seconds)
    stop[0] = True
    for th in ts:
        th.join()
    return sum(done) / (time.perf_counter() - t0) / 1e9


for t in sorted({int(x) for x in args.threads.split(",") if x.isdigit() and int(x) > 0}):
    r = pack_rate(t)
    emit({"what": "pack_only", "threads_per_rank": t, "sum_gbp_per_s": round(sum_over_ranks(r), 1), "rank0_gbp_per_s": round(r, 1)})


def e2e():
    gpu.filter_batch_ptr(hbases.data_ptr(), hoff.data_ptr(), NR, True, 0, 2, 0.01, True, hk.data_ptr(), hh.data_ptr(), ht.data_ptr())


def threads_for(spec):
    """'3' -> 3 on every rank; 'slow2' -> 2 on the ranks whose share of the host is below the mean, 0 on the others;
    'eq' -> as many as lift the rank's share to the best rank's at ~8 GB/s per packer thread (at most 4)."""
    mean = sum(shares) / len(shares)
    if spec.startswith("slow"):
        return int(spec[4:]) if shares[rank] < mean else 0
    if spec == "eq":
        return min(4, max(0, round((max(shares) - shares[rank]) / 8.0)))
    return int(spec)


for spec in args.threads.split(","):
    t = threads_for(spec)
    gpu.host_pack_threads(t)
    gpu.host_pack_fraction(-1)
    e2e()
    e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e()
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0) / args.steps
    h2d = gpu.last_transfer_bytes()[0]
    assert torch.equal(hk, keep.cpu()) and torch.equal(hh, hits.cpu()) and torch.equal(ht, tot.cpu()), "host path != device path"
    emit({"what": "e2e", "pack_threads_per_rank": spec, "ms_per_step": round(dt * 1e3, 2), "sum_gbp_per_s": round(world * nb / dt / 1e9, 1),
          "h2d_mb_rank0": round(h2d / 1e6)})
    barrier()

from deacon_server_b200 import api as A  # noqa: E402
codes_np, inv_np = A.pack_ascii(hbases.numpy())
hc = torch.from_numpy(codes_np.view(np.int32)).pin_memory()
hi = torch.from_numpy(inv_np.view(np.int16)).pin_memory()


def e2e_packed():
    gpu.filter_batch_packed_ptr(hc.data_ptr(), hi.data_ptr(), None, hoff.data_ptr(), NR, True, 0, 2, 0.01, True, hk.data_ptr(), hh.data_ptr(), ht.data_ptr())


e2e_packed()
barrier()
t0 = time.perf_counter()
for _ in range(args.steps):
    e2e_packed()
torch.cuda.synchronize()
dt = max_over_ranks(time.perf_counter() - t0) / args.steps
emit({"what": "e2e_caller_packed", "ms_per_step": round(dt * 1e3, 2), "sum_gbp_per_s": round(world * nb / dt / 1e9, 1),
      "h2d_mb_rank0": round(gpu.last_transfer_bytes()[0] / 1e6)})
barrier()
if world > 1:
    dist.destroy_process_group()
