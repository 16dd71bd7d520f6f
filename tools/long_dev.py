"""Config-3 shape, device-resident: ONT-like long reads (2 Gbp) against the 3.1 Gbp index, search mode."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
import deacon_server_b200 as d  # noqa: E402

dev = torch.device("cuda:0")
G = int(float(os.environ.get("GENOME_MBP", "3100")) * 1e6)
genome = B.make_genome(torch, dev, G, 1)
coff = torch.from_numpy(B.contig_offsets(G, 1)).to(dev)
gpu = d.DeaconGpu(0)
gpu.index_build_device(genome, coff, B.CONTIGS, G, 31, 15, 0.0, True, stream=torch.cuda.current_stream().cuda_stream)
bases, off, n, nb = B.make_long_reads(torch, dev, genome, 2_000_000_000)
del genome
keep = torch.zeros(n, dtype=torch.uint8, device=dev)
hits = torch.zeros(n, dtype=torch.int32, device=dev)
tot = torch.zeros(n, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
ms = B._timed(torch, lambda: gpu.filter_batch_device(bases, off, n, nb, keep, hits, tot, paired=False, deplete=False, stream=st), 8)
print(f"DCN_DEDUP_LOCAL={os.environ.get('DCN_DEDUP_LOCAL', '1')}: {ms:.2f} ms/step, {nb / ms / 1e6:.1f} Gbp/s; kept {int(keep.sum())}/{n}; "
      f"hits {int(hits.sum())} total {int(tot.sum())}; checksum {int((hits.long() * 31 + tot.long()).sum())}")
