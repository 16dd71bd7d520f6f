"""Early GPU look: kernel throughput on synthetic pairs, random-access ceiling, e2e with pinned host buffers.
Not the contract bench (bench.py); used while the kernels are being built."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import deacon_server_b200 as d  # noqa: E402
from oracle import oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genome-mbp", type=float, default=20)
ap.add_argument("--pad-keys-m", type=float, default=100)
ap.add_argument("--pairs-m", type=float, default=2)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--load", type=float, default=0.5)
ap.add_argument("--check", type=int, default=20000)
ap.add_argument("--no-e2e", action="store_true")
ap.add_argument("--no-hint", action="store_true", help="do not promise the unit length (dcn_filter_batch_device: one small sync per call)")
ap.add_argument("--sweep", default="", help="e2e ingest sweep: threads:fraction,threads:fraction,... (replaces the default three settings)")
args = ap.parse_args()
HINT = 0 if args.no_hint else 300

dev = torch.device("cuda:0")
torch.manual_seed(1)
G = int(args.genome_mbp * 1e6)
lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
genome = lut[torch.randint(0, 4, (G,), device=dev)]
t0 = time.time()
g_host = genome.cpu().numpy()
idx = O.index_build([g_host], 31, 15, threads=os.cpu_count())
keys = idx.keys()
print(f"oracle index: {len(keys)} keys from {G} bp in {time.time()-t0:.1f}s ({os.cpu_count()} cpus)")
npad = int(args.pad_keys_m * 1e6)
pad = torch.randint(-2**63, 2**63 - 1, (npad,), dtype=torch.int64, device=dev)
allkeys = torch.cat([torch.from_numpy(keys.view(np.int64)).to(dev), pad])
gpu = d.DeaconGpu(0)
gpu.set_load_factor(args.load)
t0 = time.time()
gpu.index_upload_device(allkeys, d.IndexHeader(2, 31, 15))
torch.cuda.synchronize()
print("table:", gpu.index_info(), f"built in {time.time()-t0:.2f}s")

# reads: 90% from genome (random strand, 0.5% subs), 10% random
NP = int(args.pairs_m * 1e6)
NR = 2 * NP
comp = torch.zeros(256, dtype=torch.uint8, device=dev)
for a, b in zip(b"ACGT", b"TGCA"):
    comp[a] = b
pos = torch.randint(0, G - 600, (NP,), device=dev)
ins = torch.randint(300, 400, (NP,), device=dev)
ar = torch.arange(150, device=dev)
m1 = genome[(pos[:, None] + ar[None, :])]
m2 = comp[genome[(pos + ins)[:, None] - 1 - ar[None, :]].long()]
reads = torch.stack([m1, m2], 1).reshape(NR, 150)
rnd = torch.rand(NP, device=dev) < 0.1
reads.view(NP, 300)[rnd] = lut[torch.randint(0, 4, (int(rnd.sum()), 300), device=dev)]
sub = torch.rand(NR, 150, device=dev) < 0.005
reads[sub] = lut[torch.randint(0, 4, (int(sub.sum()),), device=dev)]
bases = reads.reshape(-1).contiguous()
off = (torch.arange(NR + 1, device=dev, dtype=torch.int64) * 150)
keep = torch.zeros(NP, dtype=torch.uint8, device=dev)
hits = torch.zeros(NP, dtype=torch.int32, device=dev)
tot = torch.zeros(NP, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
nb = bases.numel()

def step():
    gpu.filter_batch_device(bases, off, NR, nb, keep, hits, tot, paired=True, deplete=True, stream=st, max_unit_len=HINT)

for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
print(f"device-resident: {ms:.3f} ms/step, {nb/ms/1e6:.2f} Gbp/s; minimizers/bp={float(tot.sum())/nb:.4f}; "
      f"kept={int(keep.sum())}/{NP}; hits/total={float(hits.sum())/float(tot.sum()):.3f}")

# parity on a prefix of the batch
C = min(args.check, NP)
hb = bases[:C * 300].cpu().numpy()
ho = off[:2 * C + 1].cpu().numpy().astype(np.uint64)
full = O.IndexSet(allkeys.cpu().numpy().view(np.uint64), threads=os.cpu_count())
ok, oh, ot = O.filter_batch(full, hb, ho, paired=True, deplete=True, threads=os.cpu_count())
assert np.array_equal(keep[:C].cpu().numpy(), ok) and np.array_equal(hits[:C].cpu().numpy().view(np.uint32), oh) \
    and np.array_equal(tot[:C].cpu().numpy().view(np.uint32), ot), "PARITY FAILURE"
print(f"parity ok on first {C} pairs")

# random access ceiling
n, rms = gpu.measure_random_access(1 << 28)
n, rms = gpu.measure_random_access(1 << 28)
print(f"random 32B sector ceiling: {n/rms/1e6:.2f} Gsectors/s ({n*32/rms/1e6:.1f} GB/s) over {gpu.index_info()['table_bytes']/1e9:.2f} GB")
probes = float(tot.sum())
print(f"fused kernel probe rate: {probes/ms/1e6:.2f} Gprobes/s")

if args.no_e2e:
    sys.exit(0)
# e2e with pinned host buffers
hbases = bases.cpu().pin_memory()
hoff = off.cpu().pin_memory()
hk = torch.zeros(NP, dtype=torch.uint8).pin_memory()
hh = torch.zeros(NP, dtype=torch.int32).pin_memory()
ht = torch.zeros(NP, dtype=torch.int32).pin_memory()
def e2e():
    gpu.filter_batch_ptr(hbases.data_ptr(), hoff.data_ptr(), NR, True, 0, 2, 0.01, True, hk.data_ptr(), hh.data_ptr(), ht.data_ptr())
settings = [(int(a), float(b)) for a, b in (x.split(":") for x in args.sweep.split(","))] if args.sweep else [(0, -1), (16, 1.0), (16, -1)]
for threads, fr in settings:
    gpu.host_pack_threads(threads)
    gpu.host_pack_fraction(fr)
    e2e()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e()
    dt = (time.perf_counter() - t0) / args.steps
    fms, fn = gpu.fused_time_take()   # the fused kernel's launches of the most recent calls (at most 256)
    print(f"e2e pinned, pack threads {threads} fraction {fr}: {dt*1e3:.2f} ms/step, {nb/dt/1e9:.2f} Gbp/s; timing {gpu.last_timing()} pack_ms {gpu.last_pack_ms():.2f}; "
          f"fused kernel {fms / max(fn, 1) * 1e3:.1f} us x {fn} launches; h2d {gpu.last_transfer_bytes()[0] / 1e6:.0f} MB")
    assert torch.equal(hk, keep.cpu()) and torch.equal(hh, hits.cpu())
# caller-packed input (dcn_filter_batch_packed), pinned
from deacon_server_b200 import api as A
codes_np, inv_np = A.pack_ascii(hbases.numpy())
hc = torch.from_numpy(codes_np.view(np.int32)).pin_memory()
hi = torch.from_numpy(inv_np.view(np.int16)).pin_memory()
def e2e_packed():
    gpu.filter_batch_packed_ptr(hc.data_ptr(), hi.data_ptr(), None, hoff.data_ptr(), NR, True, 0, 2, 0.01, True, hk.data_ptr(), hh.data_ptr(), ht.data_ptr())
e2e_packed()
t0 = time.perf_counter()
for _ in range(args.steps):
    e2e_packed()
dt = (time.perf_counter() - t0) / args.steps
print(f"e2e caller-packed pinned: {dt*1e3:.2f} ms/step, {nb/dt/1e9:.2f} Gbp/s; timing {gpu.last_timing()}")
assert torch.equal(hk, keep.cpu()) and torch.equal(hh, hits.cpu())
# pageable (non-pinned) caller buffers
pb, po = bases.cpu().numpy().copy(), off.cpu().numpy().copy()
gpu.host_pack_fraction(-1)
for threads in (0, 16):
    gpu.host_pack_threads(threads)
    t0 = time.perf_counter()
    k2, h2, t2 = gpu.filter_batch(pb, po.astype(np.uint64), paired=True, deplete=True)
    dt = time.perf_counter() - t0
    print(f"e2e pageable, pack threads {threads}: {dt*1e3:.2f} ms/step, {nb/dt/1e9:.2f} Gbp/s")
    assert np.array_equal(k2, keep.cpu().numpy())
t0 = time.perf_counter()
ob, oo = O.filter_batch(full, hb, ho, paired=True, deplete=True, threads=os.cpu_count())[:2]
dt = time.perf_counter() - t0
print(f"oracle CPU ({os.cpu_count()} threads): {C*300/dt/1e9:.3f} Gbp/s")
