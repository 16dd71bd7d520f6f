#!/usr/bin/env python
"""BASELINE.json configs[4]: panmouse-scale (~550 M-minimizer) synthetic index, read-sharded filter through the
server / remote_filter split - the client extracts minimizer hashes (B3, dcn_extract_device), the server classifies the
pre-hashed records (B2, dcn_lookup_batch_device), the client owns the six counters; summed over ranks with NCCL.
Run single (python tools/bench_config5.py) or under torchrun (one rank per GPU).  Prints one JSON line from rank 0.
Device-resident reads; CUDA events; max over ranks."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
import deacon_server_b200 as d  # noqa: E402
from deacon_server_b200 import parallel as par  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genome-mbp", type=float, default=4400.0)
ap.add_argument("--pairs-m", type=float, default=5.0)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--warmup", type=int, default=2)
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
saved = os.dup(1); os.dup2(2, 1)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
G = int(args.genome_mbp * 1e6)
NP = int(args.pairs_m * 1e6); NR = 2 * NP; nb = NR * 150
genome = B.make_genome(torch, dev, G, 6)
coff = torch.from_numpy(B.contig_offsets(G, 6)).to(dev)
gpu = d.DeaconGpu(local)
st = torch.cuda.current_stream().cuda_stream
t0 = time.perf_counter()
n_keys = gpu.index_build_device(genome, coff, B.CONTIGS, G, 31, 15, 0.0, True, stream=st)
torch.cuda.synchronize()
t_index = time.perf_counter() - t0
batches = [B.make_pairs(torch, dev, genome, NP, 100 + 1000 * rank + b) for b in range(3)]
del genome
off = torch.arange(NR + 1, device=dev, dtype=torch.int64) * 150
cap = int(0.11 * nb)
d_h = torch.empty(cap, dtype=torch.int64, device=dev)
d_p = torch.empty(cap, dtype=torch.int32, device=dev)
d_o = torch.empty(NR + 1, dtype=torch.int64, device=dev)
keep = torch.zeros(NP, dtype=torch.uint8, device=dev)
hits = torch.zeros(NP, dtype=torch.int32, device=dev)
tot = torch.zeros(NP, dtype=torch.int32, device=dev)
n_min = 0


def step(i):
    global n_min
    bases = batches[i % len(batches)]
    n_min = gpu.extract_device(bases, off, NR, nb, d_h, d_p, d_o, stream=st)                   # client: B3
    pair_off = d_o[::2].contiguous()                                                           # pooled list per pair
    gpu.lookup_batch_device(d_h, pair_off, NP, keep, hits, tot, 2, 0.01, True, stream=st)      # server: B2
    gpu.stats_accumulate_device(off, NR, True, keep, stream=st)                                # client: counters


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()


for i in range(args.warmup):
    step(i)
barrier()
gpu.stats_reset()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(args.steps):
    step(i)
counters = par.reduce_counters(gpu.stats(), dev)
e1.record()
barrier()
tv = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(tv, op=dist.ReduceOp.MAX)
ms = float(tv.item())
# the split path must agree with the fused local filter on the last batch
k2, h2, t2 = torch.zeros_like(keep), torch.zeros_like(hits), torch.zeros_like(tot)
gpu.filter_batch_device(batches[(args.steps - 1) % len(batches)], off, NR, nb, k2, h2, t2, paired=True, deplete=True, stream=st)
torch.cuda.synchronize()
assert torch.equal(keep, k2) and torch.equal(hits, h2) and torch.equal(tot, t2), "B3 -> B2 differs from B1"
if rank == 0:
    out = {"what": "configs[4]: server/remote_filter split (B3 extract -> B2 lookup -> counters), device-resident",
           "n_gpus": world, "index_minimizers": n_keys, "table_bytes": gpu.index_info()["table_bytes"], "index_build_s": round(t_index, 3),
           "pairs_per_step_per_gpu": NP, "steps": args.steps, "ms_per_step": round(ms / args.steps, 3),
           "gbp_per_s": round(1e-9 * nb * args.steps * world / (ms * 1e-3), 2), "minimizers_per_step": n_min,
           "equals_local_filter": True, "counters": counters}
    os.write(saved, (json.dumps(out) + "\n").encode())
if world > 1:
    dist.destroy_process_group()
