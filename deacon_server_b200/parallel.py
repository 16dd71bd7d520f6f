"""Multi-GPU host logic of the filter path: one process per GPU, the index replicated
on every GPU, record batches sharded by rank, NO collective on the data path.  The only exchange is
the sum of the six ProcessingStats counters (src/local_filter.rs:179-187) at the end of a run - one
all-reduce of 6 x u64 (NCCL when the tensors live on a GPU, gloo on the CPU for the host-logic tests).

Nothing here computes minimizers or lookups: `engine` is a DeaconGpu (or anything with its
filter_batch signature).
"""
from __future__ import annotations

import numpy as np

COUNTER_NAMES = ("total_seqs", "filtered_seqs", "total_bp", "output_bp", "filtered_bp", "output_seq_counter")


def bind_to_gpu(local_rank: int) -> list[int]:
    """Pin this process to the CPUs NVML reports as local to its GPU, so the pinned staging buffers it allocates
    afterwards (first touch) sit on the GPU's NUMA node: with 8 GPUs on a two-socket host the H2D copies otherwise
    cross the socket interconnect.  Returns the CPU list (empty = left unchanged)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local_rank]) if vis and vis.replace(",", "").isdigit() else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = [i for i in range(n) if (mask[i // 64] >> (i % 64)) & 1]
        allowed = sorted(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return []


def h2d_probe(device, nbytes: int = 1 << 30, reps: int = 3) -> dict:
    """The host-ingest ceiling of this box for the ranks that are running: pinned host -> device copy bandwidth of
    this rank ALONE (ranks take turns) and with every rank copying AT THE SAME TIME.  An ASCII base costs one byte of
    this, so the concurrent sum in GB/s is the Gbp/s the ASCII route can reach at best; when `concurrent` falls well
    below `solo` the ranks are held back by what they share (the host's DRAM and PCIe root), not by their own links.
    Works with or without an initialised process group.  -> {"solo_gbs", "concurrent_gbs", "concurrent_sum_gbs", "world"}"""
    import torch
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    world = dist.get_world_size() if multi else 1
    rank = dist.get_rank() if multi else 0
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.fill_(65)
    d = torch.empty(nbytes, dtype=torch.uint8, device=device)

    def once():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d.copy_(h, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        return nbytes / e0.elapsed_time(e1) / 1e6

    def barrier():
        torch.cuda.synchronize()
        if multi:
            dist.barrier()

    once()
    solo = 0.0
    for r in range(world):          # one rank at a time
        barrier()
        if r == rank:
            solo = max(once() for _ in range(reps))
    barrier()
    conc = min(once() for _ in range(reps))      # everybody at once (the slowest repetition: the contended one)
    barrier()
    vec = torch.tensor([conc], dtype=torch.float64, device=device)
    lo, hi = vec.clone(), vec.clone()
    if multi:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    del h, d
    # the ranks do not get equal shares of the host (measured on an 8-GPU box: four at 20 GB/s, four at 35); min and max are
    # reported so that a caller can see it.  (Dealing the work out in proportion to the shares does not raise the total,
    # measured: the host's aggregate rate is what binds, bench.py e2e_balanced_shards.)
    return {"solo_gbs": round(solo, 2), "concurrent_gbs": round(conc, 2), "concurrent_sum_gbs": round(float(vec.item()), 2),
            "concurrent_min_gbs": round(float(lo.item()), 2), "concurrent_max_gbs": round(float(hi.item()), 2), "world": world}


HOST_BUSY_GBS = 150.0   # concurrent pinned H2D rate (sum over the ranks of a host) from which packer threads stop paying


def pack_threads_for_rank(local_world: int, probe: dict | None = None) -> int:
    """Host packing threads for one rank of `local_world` on this box (dcn_host_pack_threads).

    Packing trades host work for PCIe bytes: a packed base costs ~1.5 B of host DRAM traffic (read, write, DMA read)
    and 0.25 B of the link, a copied one 1.0 B of each.  It pays while a rank's own link is what holds it back and
    costs when the ranks are already held back by the DRAM they share.  `probe` (h2d_probe) tells which: with every
    rank copying at once, while the ranks together move less than HOST_BUSY_GBS the host's memory system has room
    -> pack with the CPUs this
    process may use, shared with the other ranks, minus four (caller, enqueueing thread, driver), at most 12;
    otherwise the ranks share a host that is the limit: 0 (ASCII route only: the copy engine needs no CPU), or 2 on a
    rank whose share is well below the mean (the step waits for it).  Without a probe the round-1 rule applies
    (no packing from four ranks per host: measured host-DRAM-bound on the 8-GPU boxes of this pool)."""
    import os
    if probe is not None and probe.get("solo_gbs"):
        # Packing pays while the DMA engines leave the host's memory system room: with all ranks copying at ~55 GB/s each, two
        # ranks move 111 GB/s and 8 packers per rank lift e2e from 111 to 135 Gbp/s; four ranks move 217 GB/s and packers
        # only take memory bandwidth away from the copies (183 -> 161 / 157 / 149 Gbp/s with 2 / 4 / 6 per rank: the cores
        # of these hosts read memory at ~135 GB/s in total, tools/ingest_sweep.py).
        # Ranks that slow each other down while the host is NOT busy (two GPUs behind one PCIe root: 42 GB/s each, 85 in
        # total) share a link, not the memory system: packing moves fewer bytes over that link, so they pack too.
        host_busy = probe.get("concurrent_sum_gbs", 0.0) >= HOST_BUSY_GBS
        if host_busy:
            # Host-bound.  The ranks' shares are not equal (one 8-GPU box: four ranks at 20 GB/s, four at 35) and with the
            # same work per rank the slow ones set the step time: two packers on a rank whose share is well below the mean
            # shorten its step a little (measured with tools/ingest_sweep.py: 166-172 -> 177-180 Gbp/s; three: no better);
            # on every rank they only add DRAM traffic (166 / 163 with two / three per rank).
            mean = probe.get("concurrent_sum_gbs", 0.0) / max(1, probe.get("world", 1))
            if probe["concurrent_gbs"] < 0.85 * mean:
                return 2
            return 0
    elif local_world >= 4:
        return 0
    allowed = len(os.sched_getaffinity(0))
    total = os.cpu_count() or allowed
    share = allowed
    if local_world > 1:
        # bound to one NUMA node: the node's ranks share it (ranks are dealt out evenly over the nodes)
        nodes = max(1, round(total / allowed)) if allowed < total else 1
        share = allowed // max(1, -(-local_world // nodes))
    n = min(12, share - 4)
    return n if n >= 2 else 0


def shard_units(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced unit range [u0, u1) of `rank`: sizes differ by at most one unit and a
    pair is never split (units, not records, are dealt out)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    base, extra = divmod(n_units, world)
    u0 = rank * base + min(rank, extra)
    return u0, u0 + base + (1 if rank < extra else 0)


def shard_batch(bases: np.ndarray, rec_off: np.ndarray, paired: bool, rank: int, world: int):
    """-> (bases view, rec_off rebased to 0, u0, u1) of this rank's share of a batch."""
    rec_off = np.ascontiguousarray(rec_off, np.uint64)
    rpu = 2 if paired else 1
    n_units = (len(rec_off) - 1) // rpu
    u0, u1 = shard_units(n_units, rank, world)
    lo, hi = int(rec_off[u0 * rpu]), int(rec_off[u1 * rpu])
    return bases[lo:hi], rec_off[u0 * rpu:u1 * rpu + 1] - np.uint64(lo), u0, u1


def counters_of(rec_off: np.ndarray, keep: np.ndarray, paired: bool) -> dict:
    """The six counters of one shard from its lengths and decisions (what stats_kernel accumulates on
    the device; src/local_filter.rs:347-371 single, 488-525 paired)."""
    rpu = 2 if paired else 1
    rec_off = np.asarray(rec_off, np.uint64)
    n_units = (len(rec_off) - 1) // rpu
    ulen = (rec_off[rpu::rpu][:n_units] - rec_off[0::rpu][:n_units]).astype(np.int64)
    k = np.asarray(keep[:n_units], bool)
    total_bp, out_bp = int(ulen.sum()), int(ulen[k].sum())
    n_keep = int(k.sum()) * rpu
    return {"total_seqs": n_units * rpu, "filtered_seqs": n_units * rpu - n_keep, "total_bp": total_bp,
            "output_bp": out_bp, "filtered_bp": total_bp - out_bp, "output_seq_counter": n_keep}


def reduce_counters(counters: dict, device=None) -> dict:
    """Sum the six counters over all ranks (the path's only collective).  Without an initialised
    process group (single GPU) this is the identity."""
    import torch
    import torch.distributed as dist
    vec = torch.tensor([int(counters[k]) for k in COUNTER_NAMES], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return {k: int(v) for k, v in zip(COUNTER_NAMES, vec.tolist())}


def gather_decisions(keep: np.ndarray, n_units_total: int, device=None) -> np.ndarray:
    """Re-assemble the per-unit keep flags of all shards in input order (the host re-interleaves
    outputs).  Used by drivers that write one output stream; not on the timed path."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(keep, np.uint8)
    world = dist.get_world_size()
    width = -(-n_units_total // world)
    mine = torch.zeros(width, dtype=torch.uint8, device=device)
    mine[:len(keep)] = torch.from_numpy(np.ascontiguousarray(keep, np.uint8)).to(mine.device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = np.empty(n_units_total, np.uint8)
    for r, p in enumerate(parts):
        u0, u1 = shard_units(n_units_total, r, world)
        out[u0:u1] = p[:u1 - u0].cpu().numpy()
    return out


def filter_sharded(engine, bases: np.ndarray, rec_off: np.ndarray, paired: bool = False, device=None, **kw):
    """Filter this rank's share of a batch with `engine.filter_batch` and reduce the counters.
    -> (keep, hits, total) of the local shard, (u0, u1), reduced counters."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    sb, so, u0, u1 = shard_batch(bases, rec_off, paired, rank, world)
    keep, hits, total = engine.filter_batch(sb, so, paired=paired, **kw)
    return (keep, hits, total), (u0, u1), reduce_counters(counters_of(so, keep, paired), device)
