"""Config-3 shape end to end (ONT-like long reads from pinned host memory through dcn_filter_batch), per pipeline form."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
import deacon_server_b200 as d  # noqa: E402

dev = torch.device("cuda:0")
G = 200_000_000
genome = B.make_genome(torch, dev, G, 1)
coff = torch.from_numpy(B.contig_offsets(G, 1)).to(dev)
gpu = d.DeaconGpu(0)
gpu.index_build_device(genome, coff, B.CONTIGS, G, 31, 15, 0.0, True, stream=torch.cuda.current_stream().cuda_stream)
bases, off, n, nb = B.make_long_reads(torch, dev, genome, 2_000_000_000)
keep = torch.zeros(n, dtype=torch.uint8, device=dev)
hits = torch.zeros(n, dtype=torch.int32, device=dev)
tot = torch.zeros(n, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
ms = B._timed(torch, lambda: gpu.filter_batch_device(bases, off, n, nb, keep, hits, tot, paired=False, deplete=False, stream=st), 5)
print(f"device-resident: {ms:.2f} ms, {nb / ms / 1e6:.1f} Gbp/s")
hb = bases[:nb].cpu().pin_memory()
ho = off.cpu().pin_memory()
hk = torch.zeros(n, dtype=torch.uint8).pin_memory()
hh = torch.zeros(n, dtype=torch.int32).pin_memory()
ht = torch.zeros(n, dtype=torch.int32).pin_memory()
for threads in [int(x) for x in os.environ.get("THREADS", "12").split(",")]:
    gpu.host_pack_threads(threads)
    ts = []
    for i in range(8):
        t0 = time.perf_counter()
        gpu.filter_batch_ptr(hb.data_ptr(), ho.data_ptr(), n, False, 0, 2, 0.01, False, hk.data_ptr(), hh.data_ptr(), ht.data_ptr())
        ts.append(time.perf_counter() - t0)
    assert torch.equal(hk, keep.cpu()) and torch.equal(hh, hits.cpu()) and torch.equal(ht, tot.cpu())
    print(f"pipeline {os.environ.get('DCN_PIPELINE', 'arena')} threads {threads}: calls ms " + " ".join(f"{t * 1e3:.1f}" for t in ts) +
          f" -> steady {nb / np.median(ts[3:]) / 1e9:.1f} Gbp/s; h2d {gpu.last_transfer_bytes()[0] / 1e6:.0f} MB")
