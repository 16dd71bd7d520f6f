"""Host-side multi-GPU logic on the CPU: world_size-2 `gloo` process group.  The per-shard engine is
the oracle here (there is no GPU in this container); on a GPU box the same functions are driven
with DeaconGpu by bench.py under torchrun."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from deacon_server_b200 import parallel as P
from oracle import oracle as O


class OracleEngine:
    def __init__(self, idx):
        self.idx = idx

    def filter_batch(self, bases, rec_off, paired=False, prefix_length=0, abs_threshold=2, rel_threshold=0.01, deplete=False):
        return O.filter_batch(self.idx, bases, rec_off, paired=paired, prefix_len=prefix_length, abs_thr=abs_threshold,
                              rel_thr=rel_threshold, deplete=deplete)


def _data(paired):
    g = H.random_genome(40_000, 5)
    reads = H.sample_reads(g, 1001 if not paired else 1002, (0, 300), 6)
    return g, H.concat(reads)


def _worker(rank, world, port, paired, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g, (bases, off) = _data(paired)
        eng = OracleEngine(O.index_build([g], 31, 15))
        (keep, hits, total), (u0, u1), counters = P.filter_sharded(eng, bases, off, paired=paired, deplete=True)
        n_units = (len(off) - 1) // (2 if paired else 1)
        full = P.gather_decisions(keep, n_units)
        q.put((rank, u0, u1, counters, full.tobytes(), hits.tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("paired", [False, True])
def test_two_rank_sharding_and_counter_reduce(paired):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, paired, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g, (bases, off) = _data(paired)
    idx = O.index_build([g], 31, 15)
    ok, oh, ot = O.filter_batch(idx, bases, off, paired=paired, deplete=True)
    want = P.counters_of(off, ok, paired)
    (r0, a0, b0, c0, f0, h0), (r1, a1, b1, c1, f1, h1) = res
    assert (a0, b1) == (0, len(ok)) and b0 == a1 and abs((b0 - a0) - (b1 - a1)) <= 1      # balanced, contiguous, complete
    assert c0 == c1 == want                                                               # all-reduced sum == single-process run
    assert np.array_equal(np.frombuffer(f0, np.uint8), ok) and f0 == f1                    # decisions re-assembled in input order
    assert np.array_equal(np.concatenate([np.frombuffer(h0, np.uint32), np.frombuffer(h1, np.uint32)]), oh)
    assert want["total_seqs"] == len(off) - 1 - ((len(off) - 1) % 2 if paired else 0)
    assert want["total_bp"] == want["output_bp"] + want["filtered_bp"]


def test_shard_units_partition():
    for n in (0, 1, 7, 8, 1000, 12345):
        for world in (1, 2, 3, 8):
            ranges = [P.shard_units(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        P.shard_units(10, 2, 2)


def test_counters_match_reference_semantics():
    """src/local_filter.rs:488-525: a pair counts as two sequences and both mates' bases."""
    off = np.array([0, 100, 250, 300, 420], np.uint64)
    c = P.counters_of(off, np.array([1, 0], np.uint8), paired=True)
    assert c == {"total_seqs": 4, "filtered_seqs": 2, "total_bp": 420, "output_bp": 250, "filtered_bp": 170, "output_seq_counter": 2}
    c = P.counters_of(off, np.array([0, 1, 1, 0], np.uint8), paired=False)
    assert c == {"total_seqs": 4, "filtered_seqs": 2, "total_bp": 420, "output_bp": 200, "filtered_bp": 220, "output_seq_counter": 2}


def test_pack_threads_for_rank(monkeypatch):
    """Host packing threads a rank gets (parallel.pack_threads_for_rank): its share of the CPUs minus four (at most 12),
    none when that leaves fewer than two; whether packing pays at all is read off the ingest probe (a rank that keeps
    >= 80 % of its solo H2D rate while all ranks copy is link-bound), and without a probe from the rank count."""
    import os
    from deacon_server_b200 import parallel as P
    monkeypatch.setattr(os, "cpu_count", lambda: 16)
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(16)))
    assert P.pack_threads_for_rank(1) == 12
    assert P.pack_threads_for_rank(2) == 4
    assert P.pack_threads_for_rank(8) == 0
    monkeypatch.setattr(os, "cpu_count", lambda: 224)
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(224)))
    assert P.pack_threads_for_rank(1) == 12
    assert P.pack_threads_for_rank(2) == 12
    assert P.pack_threads_for_rank(4) == 0
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(112)))     # bound to one of two sockets
    assert P.pack_threads_for_rank(2) == 12
    # with a probe the decision follows the measurement, not the rank count
    monkeypatch.setattr(os, "cpu_count", lambda: 64)
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(64)))
    link_bound = {"solo_gbs": 54.0, "concurrent_gbs": 52.0, "concurrent_sum_gbs": 104.0, "world": 2}
    four_links = {"solo_gbs": 54.0, "concurrent_gbs": 52.0, "concurrent_sum_gbs": 208.0, "world": 4}
    host_bound = {"solo_gbs": 54.0, "concurrent_gbs": 21.5, "concurrent_sum_gbs": 172.0, "world": 8}
    assert P.pack_threads_for_rank(2, link_bound) == 12
    assert P.pack_threads_for_rank(4, four_links) == 0                                   # every link is full, and so is the host: packers would take memory bandwidth from the copies
    assert P.pack_threads_for_rank(8, host_bound) == 0
    assert P.pack_threads_for_rank(8, dict(host_bound, concurrent_gbs=17.0)) == 2       # a rank well below the mean share (21.5): the step waits for it
    assert P.pack_threads_for_rank(4, dict(link_bound, world=4)) == 12                   # four ranks on a host whose links are slow: 64 / 4 - 4
    shared_root = {"solo_gbs": 55.0, "concurrent_gbs": 42.4, "concurrent_sum_gbs": 84.9, "world": 2}
    assert P.pack_threads_for_rank(2, shared_root) == 12                                 # two ranks slow each other down over a shared link; the host's memory has room
