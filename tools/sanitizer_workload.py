import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
import deacon_server_b200 as d
from oracle import oracle as O
g = H.random_genome(120_000, 1)
idx = O.index_build([g], 31, 15, threads=4)
gpu = d.DeaconGpu(0)
gpu.index_upload(idx.keys(), d.IndexHeader(2, 31, 15))
reads = H.sample_reads(g, 1500, (1, 500), 2) + [g[:6000], g[10000:13000]] + H.sample_reads(g, 300, 150, 3)
reads += [np.frombuffer(b"A" * 900, np.uint8).copy()] * 4          # more picks than a warp pass holds: CTA tail
for paired in (False, True):
    recs = reads[: len(reads) // 2 * 2]
    bases, off = H.concat(recs)
    k, h, t = gpu.filter_batch(bases, off, paired=paired, deplete=True)
    ok, oh, ot = O.filter_batch(idx, bases, off, paired=paired, deplete=True, threads=4)
    assert np.array_equal(k, ok) and np.array_equal(h, oh) and np.array_equal(t, ot)
short = H.sample_reads(g, 3000, 150, 5)
bases, off = H.concat(short)
hh, pp, oo = gpu.extract(bases, off)
lists_off = oo
k, h, t = gpu.lookup_batch(hh, oo)
k2, h2, t2 = gpu.filter_batch(bases, off)
assert np.array_equal(k, k2) and np.array_equal(h, h2)
keys = gpu.index_build(np.concatenate([g, g[:1000]]), np.array([0, len(g), len(g) + 1000], np.uint64), 31, 15, 0.0, False)
assert np.array_equal(np.sort(keys), np.sort(O.index_build([g, g[:1000]], 31, 15).keys()))
print("sanitizer workload ok")
