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
# the arena form of the host pipeline (DCN_CHUNK_MB=1 in the environment: 128 KB atoms, so a few MB are hundreds of
# claims and dozens of launches), long units among the short ones, and the caller-packed sparse form
if os.environ.get("DCN_CHUNK_MB") == "1":
    from deacon_server_b200 import api as A
    rng = np.random.default_rng(9)
    n = 40_000
    lens = rng.integers(0, 400, n).astype(np.uint64)
    lens[::997] = 9_000
    off = np.zeros(n + 1, np.uint64); off[1:] = np.cumsum(lens)
    total = int(off[-1])
    start = rng.integers(0, len(g) - 300, total // 128 + 2)
    bases = g[(start[:, None] + np.arange(128)[None, :])].reshape(-1)[:total].copy()
    bases[rng.integers(0, total, total // 5000)] = ord("N")
    want = O.filter_batch(idx, bases, off, paired=True, deplete=True, threads=4)
    for threads in (0, 5):
        gpu.host_pack_threads(threads)
        got = gpu.filter_batch(bases, off, paired=True, deplete=True)
        for a, b in zip(got, want):
            assert np.array_equal(a, b), threads
    gpu.host_pack_threads(0)
    codes, exc, nl = A.pack_records_sparse(bases, off, 31, 0)
    got = gpu.filter_batch_packed_sparse(codes, exc, nl, off, paired=True, deplete=True)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
print("sanitizer workload ok")
