#!/usr/bin/env python
"""bench.py -- `deacon filter` hot path on B200: filter Gbp/s (bit-exact decisions) vs reference CPU Gbp/s.

Workload (BASELINE.json configs[1]): paired-end 2x150 bp reads (50 M pairs per GPU, processed as
`--steps` batches of `--pairs-per-step` pairs) against a synthetic 3.1 Gbp / 24-contig random
reference indexed with k=31, w=15 (~390 M minimizers), `--deplete -a 2 -r 0.01`.  Data is synthetic
(seeded): 90 % of pairs are sampled from the reference (random strand, insert 300-400, 0.5 %
substitutions), 10 % are random sequence.

A "step" = one pass of the hot path (extract -> lookup -> classify) over one batch.
  value : whole-job Gbp/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e   : the same through the host-pointer C-ABI call (dcn_filter_batch) from pinned host buffers,
          H2D and D2H copies inside the timed region
  --impl reference : the CPU restatement of the reference path (oracle/) on all host cores

One process per GPU; under torchrun (WORLD_SIZE > 1) every rank runs the same per-GPU workload
("weak" scaling: index replicated, reads sharded) and the six summary counters are all-reduced
over NCCL at the end of the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONTIGS = 24
READ_LEN = 150
# The port is a scalar C restatement (no SIMD); the reference itself uses AVX2 simd-minimizers and its README claims
# more: keep that claim beside every CPU number so that nobody reads the GPU/CPU ratio as a GPU/reference ratio.
PUBLISHED_NOTE = ("the reference's own claim is \">2 Gbp/s\" filtering uncompressed long reads on unstated hardware "
                  "(/root/reference/README.md:14); it cannot be built in this image (no cargo/rustc), so the CPU arm is the "
                  "scalar C port and a GPU/port ratio overstates GPU/reference by the port's distance from that claim")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-mbp", type=float, default=3100.0, help="synthetic reference size (config: 3100)")
    ap.add_argument("--pairs-per-step", type=float, default=5e6, help="pairs per step (10 steps = the 50 M-pair job)")
    ap.add_argument("--batches", type=int, default=10, help="distinct resident batches cycled by the steps")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end leg (0 = same as --steps)")
    ap.add_argument("--cpu-sample-pairs", type=float, default=2e6, help="pairs of the cpu_baseline sample")
    ap.add_argument("--ref-pairs-per-step", type=float, default=1e6, help="--impl reference: pairs per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", default="lookup,long,build,config5",
                    help="extra legs recorded under \"extra\" at N=1 (comma list of lookup,long,build,config5; empty = none)")
    ap.add_argument("--long-gbp", type=float, default=2.0, help="extra leg 'long': bases of ONT-like reads per step")
    ap.add_argument("--seed", type=int, default=20261018)
    return ap.parse_args()


# --------------------------------------------------------------------------------------- data
def contig_offsets(total: int, seed: int) -> np.ndarray:
    """24 contigs of 50-250 Mbp (scaled) summing to `total` bases."""
    rng = np.random.default_rng(seed)
    w = rng.uniform(50, 250, CONTIGS)
    lens = np.floor(w / w.sum() * total).astype(np.int64)
    lens[-1] += total - int(lens.sum())
    off = np.zeros(CONTIGS + 1, np.int64)
    off[1:] = np.cumsum(lens)
    return off


def make_genome(torch, dev, total: int, seed: int):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    out = torch.empty(total, dtype=torch.uint8, device=dev)
    step = 1 << 28
    for s in range(0, total, step):
        n = min(step, total - s)
        out[s:s + n] = lut[torch.randint(0, 4, (n,), device=dev, generator=g)]
    return out


def make_pairs(torch, dev, genome, n_pairs: int, seed: int):
    """-> uint8 tensor [2 * n_pairs * 150]: mate 1 forward at p, mate 2 reverse complement ending at p + insert."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    G = genome.numel()
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    comp = torch.zeros(256, dtype=torch.uint8, device=dev)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    out = torch.empty(n_pairs * 2 * READ_LEN, dtype=torch.uint8, device=dev)
    ar = torch.arange(READ_LEN, device=dev)
    sub = 1 << 20
    for s in range(0, n_pairs, sub):
        n = min(sub, n_pairs - s)
        pos = torch.randint(0, G - 600, (n,), device=dev, generator=g)
        ins = torch.randint(300, 400, (n,), device=dev, generator=g)
        m1 = genome[pos[:, None] + ar[None, :]]
        m2 = comp[genome[(pos + ins)[:, None] - 1 - ar[None, :]].long()]
        swap = torch.rand(n, device=dev, generator=g) < 0.5          # random strand of the fragment
        pair = torch.where(swap[:, None, None], torch.stack([m2, m1], 1), torch.stack([m1, m2], 1))
        rnd = torch.rand(n, device=dev, generator=g) < 0.10          # 10 % non-host pairs
        nr = int(rnd.sum())
        if nr:
            pair[rnd] = lut[torch.randint(0, 4, (nr, 2, READ_LEN), device=dev, generator=g)]
        m = torch.rand(n, 2, READ_LEN, device=dev, generator=g) < 0.005   # substitutions
        nm = int(m.sum())
        if nm:
            pair[m] = lut[torch.randint(0, 4, (nm,), device=dev, generator=g)]
        out[s * 2 * READ_LEN:(s + n) * 2 * READ_LEN] = pair.reshape(-1)
    return out


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_capture():
    """What only a profiler can see (DRAM bytes per launch, pipe and issue utilisation) comes from the committed ncu
    capture of the SHIPPED kernel on this workload (profiles/traffic.json names the report it was read from);
    everything else in the roofline block is measured by this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


# --------------------------------------------------------------------------------------- extra legs (N = 1)
def _timed(torch, fn, steps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def extra_lookup(torch, dev, gpu, n_keys, pairs, steps, ceiling):
    """B2 dcn_lookup_batch_device on pre-hashed pairs (the server's request shape, src/remote_filter.rs:266-301):
    hash lists drawn from the index's own keys (hits) and random values (misses), ~28 per pair like 2x150 bp reads."""
    import ctypes
    st = torch.cuda.current_stream().cuda_stream
    rng = torch.Generator(device=dev); rng.manual_seed(7)
    per = torch.randint(24, 33, (pairs,), device=dev, generator=rng)
    off = torch.zeros(pairs + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(per, 0)
    nh = int(off[-1])
    keys = torch.empty(n_keys, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    ctypes.CDLL("libcudart.so.12").cudaMemcpy(ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(gpu.index_build_keys_ptr()),
                                             ctypes.c_size_t(n_keys * 8), 3)
    hashes = keys[torch.randint(0, n_keys, (nh,), device=dev, generator=rng)]
    miss = torch.rand(nh, device=dev, generator=rng) < 0.2
    hashes[miss] = torch.randint(-2**62, 2**62, (int(miss.sum()),), dtype=torch.int64, device=dev, generator=rng)
    del keys
    keep = torch.zeros(pairs, dtype=torch.uint8, device=dev)
    hits = torch.zeros(pairs, dtype=torch.int32, device=dev)
    tot = torch.zeros(pairs, dtype=torch.int32, device=dev)
    ms = _timed(torch, lambda: gpu.lookup_batch_device(hashes, off, pairs, keep, hits, tot, 2, 0.01, True, stream=st), steps)
    # parity of a sample against the oracle's classification of the same hash lists
    return {"what": "B2 dcn_lookup_batch_device (pre-hashed pairs, device-resident)", "records": pairs, "hashes": nh,
            "ms_per_step": round(ms, 3), "gprobes_per_s": round(nh / ms / 1e6, 2),
            "random_sector_ceiling_gsectors_per_s": round(ceiling, 2), "frac_random_sector": round(nh / ms / 1e6 / ceiling, 3),
            "equiv_gbp_per_s_at_0.0942_minimizers_per_bp": round(nh / 0.0942 / ms / 1e6, 1),
            "hit_fraction": round(float(hits.sum()) / nh, 3)}


def make_long_reads(torch, dev, genome, total, seed=5):
    """BASELINE configs[2] shape: lengths gamma(shape 2) mean 10 kbp clipped [200, 200 000], 50 % host-derived /
    50 % random, 5 % substitutions.  -> (bases u8 [padded to 16], off i64 [n + 1], n, nb)"""
    G = genome.numel()
    rs = np.random.default_rng(seed)
    lens = np.clip(rs.gamma(2.0, 5000.0, int(total / 10000 * 1.2)), 200, 200_000).astype(np.int64)
    lens = lens[np.cumsum(lens) <= total]
    n = len(lens)
    off_h = np.zeros(n + 1, np.int64); off_h[1:] = np.cumsum(lens)
    nb = int(off_h[-1])
    off = torch.from_numpy(off_h).to(dev)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    rng = torch.Generator(device=dev); rng.manual_seed(seed + 4)
    starts = torch.randint(0, G - 200_001, (n,), device=dev, generator=rng)
    host = torch.rand(n, device=dev, generator=rng) < 0.5
    rec_of = torch.repeat_interleave(torch.arange(n, device=dev), torch.from_numpy(lens).to(dev))
    src = starts[rec_of] + (torch.arange(nb, device=dev) - off[:-1][rec_of])
    bases = genome[src]
    rnd = ~host[rec_of]
    bases[rnd] = lut[torch.randint(0, 4, (int(rnd.sum()),), device=dev, generator=rng)]
    sub = torch.rand(nb, device=dev, generator=rng) < 0.05
    bases[sub] = lut[torch.randint(0, 4, (int(sub.sum()),), device=dev, generator=rng)]
    del rec_of, src, rnd, sub
    pad = (-nb) % 16
    if pad:
        bases = torch.cat([bases, torch.zeros(pad, dtype=torch.uint8, device=dev)])
    return bases, off, n, nb


def extra_long(torch, dev, gpu, genome, total, steps, oracle_idx=None):
    """configs[2]: ONT-like long reads, search mode: device-resident and end to end; first reads checked against the oracle."""
    st = torch.cuda.current_stream().cuda_stream
    bases, off, n, nb = make_long_reads(torch, dev, genome, total)
    keep = torch.zeros(n, dtype=torch.uint8, device=dev)
    hits = torch.zeros(n, dtype=torch.int32, device=dev)
    tot = torch.zeros(n, dtype=torch.int32, device=dev)
    ms = _timed(torch, lambda: gpu.filter_batch_device(bases, off, n, nb, keep, hits, tot, paired=False, deplete=False, stream=st), steps)
    out = {"what": "configs[2]: ONT-like long reads (gamma(2), mean 10 kbp, 5 % substitutions), search mode",
           "reads": n, "bases": nb, "ms_per_step": round(ms, 3), "value": round(nb / ms / 1e6, 2), "unit": "Gbp/s",
           "minimizers_per_bp": round(float(tot.sum()) / nb, 4), "kept": int(keep.sum()),
           "gprobes_per_s": round(float(tot.sum()) / ms / 1e6, 2)}
    hb = bases[:nb].cpu().pin_memory()
    ho = off.cpu().pin_memory()
    hk = torch.zeros(n, dtype=torch.uint8).pin_memory()
    hh = torch.zeros(n, dtype=torch.int32).pin_memory()
    ht = torch.zeros(n, dtype=torch.int32).pin_memory()

    def e2e_step():
        gpu.filter_batch_ptr(hb.data_ptr(), ho.data_ptr(), n, False, 0, 2, 0.01, False, hk.data_ptr(), hh.data_ptr(), ht.data_ptr())

    for _ in range(2):
        e2e_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    dt = (time.perf_counter() - t0) / steps
    assert torch.equal(hk, keep.cpu()) and torch.equal(hh, hits.cpu()) and torch.equal(ht, tot.cpu()), "config 3: e2e != device-resident"
    h2d, d2h = gpu.last_transfer_bytes()
    out["e2e"] = {"value": round(nb / dt / 1e9, 2), "unit": "Gbp/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}
    if oracle_idx is not None:
        from oracle import oracle as O
        m = min(n, 300)
        end = int(off[m])
        ok, oh, ot = O.filter_batch(oracle_idx, hb[:end].numpy(), ho[:m + 1].numpy().astype(np.uint64), paired=False,
                                    abs_thr=2, rel_thr=0.01, deplete=False, threads=os.cpu_count() or 1)
        good = bool(np.array_equal(hk[:m].numpy(), ok) and np.array_equal(hh[:m].numpy().view(np.uint32), oh)
                    and np.array_equal(ht[:m].numpy().view(np.uint32), ot))
        assert good, "config 3: GPU result differs from the oracle"
        out["parity_vs_oracle"] = f"bit-exact on the first {m} reads ({end / 1e6:.1f} Mbp)"
    return out


def extra_build(torch, dev, gpu, genome, coff, G):
    """configs[3]: index build with -e 0.5 on the reference with 2 % low-complexity inserts (GPU extraction + radix sort/unique)."""
    st = torch.cuda.current_stream().cuda_stream
    g2 = genome.clone()
    rs = np.random.default_rng(11)
    n_ins = int(G * 0.02 / 300)
    stride = G // n_ins   # one insert per stride: no two overlap, so the scatter below is deterministic (and the key counts reproducible)
    pos = torch.from_numpy(np.arange(n_ins, dtype=np.int64) * stride + rs.integers(0, stride - 400, n_ins)).to(dev)
    ar = torch.arange(300, device=dev)
    kind = torch.from_numpy(rs.integers(0, 2, n_ins)).to(dev)
    homo = torch.tensor([65, 84], dtype=torch.uint8, device=dev)[kind][:, None].expand(n_ins, 300)
    dinuc = torch.tensor([[65, 67], [71, 84]], dtype=torch.uint8, device=dev)[kind][:, ar % 2]
    pick = torch.from_numpy(rs.integers(0, 2, n_ins)).to(dev).bool()
    g2[(pos[:, None] + ar[None, :]).reshape(-1)] = torch.where(pick[:, None], homo, dinuc).reshape(-1)
    torch.cuda.synchronize()
    res = {}
    for thr in (0.0, 0.5):
        gpu.index_build_device(g2, coff, CONTIGS, G, 31, 15, thr, False, stream=st)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        nk = gpu.index_build_device(g2, coff, CONTIGS, G, 31, 15, thr, False, stream=st)
        torch.cuda.synchronize()
        res[str(thr)] = {"keys": nk, "seconds": round(time.perf_counter() - t1, 4)}
    del g2
    return {"what": "configs[3]: index build, 2 % low-complexity inserts (extract + radix sort + unique), device-resident FASTA",
            "reference_mbp": G / 1e6, "by_entropy_threshold": res,
            "gbp_per_s_at_e0.5": round(G / 1e9 / max(res["0.5"]["seconds"], 1e-9), 1)}


def extra_config5(torch, dev, local, pairs, steps, genome_mbp=4400.0, seed=6):
    """configs[4] on this GPU: ~550 M-minimizer index, the server/remote_filter split: B3 extract (client) -> B2 lookup
    (server) -> counters; must equal the fused local filter; a sample is checked against the oracle's hash lists."""
    import deacon_server_b200 as d
    g5 = d.DeaconGpu(local)
    try:
        st = torch.cuda.current_stream().cuda_stream
        G = int(genome_mbp * 1e6)
        genome = make_genome(torch, dev, G, seed)
        coff = torch.from_numpy(contig_offsets(G, seed)).to(dev)
        t0 = time.perf_counter()
        n_keys = g5.index_build_device(genome, coff, CONTIGS, G, 31, 15, 0.0, True, stream=st)
        torch.cuda.synchronize()
        t_index = time.perf_counter() - t0
        NR = 2 * pairs
        nb = NR * READ_LEN
        batches = [make_pairs(torch, dev, genome, pairs, 100 + b) for b in range(2)]
        del genome
        off = torch.arange(NR + 1, device=dev, dtype=torch.int64) * READ_LEN
        cap = int(0.11 * nb)
        d_h = torch.empty(cap, dtype=torch.int64, device=dev)
        d_p = torch.empty(cap, dtype=torch.int32, device=dev)
        d_o = torch.empty(NR + 1, dtype=torch.int64, device=dev)
        keep = torch.zeros(pairs, dtype=torch.uint8, device=dev)
        hits = torch.zeros(pairs, dtype=torch.int32, device=dev)
        tot = torch.zeros(pairs, dtype=torch.int32, device=dev)
        state = {"i": 0, "n_min": 0}

        def step():
            bases = batches[state["i"] % len(batches)]
            state["i"] += 1
            state["n_min"] = g5.extract_device(bases, off, NR, nb, d_h, d_p, d_o, stream=st)        # client: B3
            g5.lookup_batch_device(d_h, d_o[::2].contiguous(), pairs, keep, hits, tot, 2, 0.01, True, stream=st)   # server: B2
            g5.stats_accumulate_device(off, NR, True, keep, stream=st)                               # client: counters

        ms = _timed(torch, step, steps)
        k2, h2, t2 = torch.zeros_like(keep), torch.zeros_like(hits), torch.zeros_like(tot)
        g5.filter_batch_device(batches[(state["i"] - 1) % len(batches)], off, NR, nb, k2, h2, t2, paired=True, deplete=True, stream=st)
        torch.cuda.synchronize()
        assert torch.equal(keep, k2) and torch.equal(hits, h2) and torch.equal(tot, t2), "config 5: B3 -> B2 differs from B1"
        # oracle: extraction of a sample of the reads must give the same hash lists (order included)
        from oracle import oracle as O
        m = 2000
        hb = batches[(state["i"] - 1) % len(batches)][:m * READ_LEN].cpu().numpy()
        go = d_o[:m + 1].cpu().numpy().astype(np.int64)
        gh = d_h[:int(go[m])].cpu().numpy().view(np.uint64)
        for r in range(m):
            oh = O.extract_filter(hb[r * READ_LEN:(r + 1) * READ_LEN], 31, 15, 0)[0]
            assert np.array_equal(gh[go[r]:go[r + 1]], oh), "config 5: B3 hash list differs from the oracle"
        return {"what": "configs[4] shape on one GPU: B3 dcn_extract_device -> B2 dcn_lookup_batch_device -> counters, device-resident",
                "index_minimizers": n_keys, "table_bytes": g5.index_info()["table_bytes"], "index_build_s": round(t_index, 3),
                "pairs_per_step": pairs, "ms_per_step": round(ms, 3), "value": round(nb / ms / 1e6, 2), "unit": "Gbp/s",
                "minimizers_per_step": int(state["n_min"]), "equals_fused_filter": True,
                "parity_vs_oracle": f"B3 hash lists of the first {m} reads bit-exact (order included)"}
    finally:
        g5.close()


# --------------------------------------------------------------------------------------- arms
def run_ours(args):
    import torch
    import torch.distributed as dist
    import deacon_server_b200 as d
    from deacon_server_b200 import parallel as par

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bound = par.bind_to_gpu(local) if world > 1 and not os.environ.get("DCN_NO_BIND") else []
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    G = int(args.genome_mbp * 1e6)
    NP = int(args.pairs_per_step)
    NR = 2 * NP
    nb = NR * READ_LEN
    t_setup = time.time()
    genome = make_genome(torch, dev, G, args.seed)
    coff = torch.from_numpy(contig_offsets(G, args.seed)).to(dev)
    gpu = d.DeaconGpu(local)
    # host-ingest ceiling of this box with the ranks that are running (parallel.h2d_probe), and the packing policy it implies
    ingest = par.h2d_probe(dev)
    pack_threads = None   # library default (hardware threads this process may use - 4, at most 12)
    if not os.environ.get("DCN_PACK_THREADS"):
        pack_threads = par.pack_threads_for_rank(int(os.environ.get("LOCAL_WORLD_SIZE", world)), ingest)
        gpu.host_pack_threads(pack_threads)
    t0 = time.time()
    n_keys = gpu.index_build_device(genome, coff, CONTIGS, G, 31, 15, 0.0, True, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    t_index = time.time() - t0
    n_batches = max(1, min(args.batches, args.steps + args.warmup))
    batches = [make_pairs(torch, dev, genome, NP, args.seed + 1000 * rank + 1 + b) for b in range(n_batches)]
    off = torch.arange(NR + 1, device=dev, dtype=torch.int64) * READ_LEN
    keep = torch.zeros(NP, dtype=torch.uint8, device=dev)
    hits = torch.zeros(NP, dtype=torch.int32, device=dev)
    tot = torch.zeros(NP, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    t_setup = time.time() - t_setup
    stream = torch.cuda.current_stream().cuda_stream

    def step(i):
        gpu.filter_batch_device(batches[i % n_batches], off, NR, nb, keep, hits, tot, paired=True, abs_threshold=2,
                                rel_threshold=0.01, deplete=True, stream=stream, max_unit_len=2 * READ_LEN)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident leg
    for i in range(args.warmup):
        step(i)
    barrier()
    gpu.fused_time_take()
    gpu.stats_reset()
    launches0 = gpu.launch_count()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    counters = gpu.stats()                       # syncs; the six ProcessingStats counters of this rank
    counters = par.reduce_counters(counters, dev)   # the path's only collective: NCCL sum of 6 x u64
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = gpu.launch_count() - launches0
    fused_ms, fused_n = gpu.fused_time_take()
    minim_per_step = None
    tvec = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tvec, op=dist.ReduceOp.MAX)
    ms_total = float(tvec.item())
    total_minimizers_last = int(tot.sum().item())
    minim_per_step = total_minimizers_last
    kept_last = int(keep.sum().item())

    # ---- end-to-end leg: pinned host buffers through dcn_filter_batch (H2D + kernels + D2H per step)
    e2e_steps = args.e2e_steps or args.steps
    n_host = min(2, n_batches)
    hb = [batches[b].cpu().pin_memory() for b in range(n_host)]
    hoff = off.cpu().pin_memory()
    hk = torch.zeros(NP, dtype=torch.uint8).pin_memory()
    hh = torch.zeros(NP, dtype=torch.int32).pin_memory()
    ht = torch.zeros(NP, dtype=torch.int32).pin_memory()

    def e2e_step(i):
        gpu.filter_batch_ptr(hb[i % n_host].data_ptr(), hoff.data_ptr(), NR, True, 0, 2, 0.01, True,
                             hk.data_ptr(), hh.data_ptr(), ht.data_ptr())

    for i in range(min(3, args.warmup)):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_h2d, e2e_d2h = gpu.last_transfer_bytes()   # what the last call really moved (the library counts its copies)
    evec = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(evec, op=dist.ReduceOp.MAX)
    e2e_s = float(evec.item())
    clk = clocks.stop()

    # ---- the same end to end with the batch already packed by the caller (dcn_filter_batch_packed: what a
    # host-side FASTQ parser can emit directly); reported beside e2e, never instead of it
    from deacon_server_b200 import api as A
    codes_np, inv_np = A.pack_ascii(hb[0].numpy())
    hc = torch.from_numpy(codes_np.view(np.int32)).pin_memory()
    hi = torch.from_numpy(inv_np.view(np.int16)).pin_memory()

    def packed_step():
        gpu.filter_batch_packed_ptr(hc.data_ptr(), hi.data_ptr(), None, hoff.data_ptr(), NR, True, 0, 2, 0.01, True,
                                    hk.data_ptr(), hh.data_ptr(), ht.data_ptr())

    packed_step()
    barrier()
    p_steps = max(1, min(e2e_steps, 5))
    t0 = time.perf_counter()
    for i in range(p_steps):
        packed_step()
    torch.cuda.synchronize()
    pvec = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    p_h2d, p_d2h = gpu.last_transfer_bytes()
    if world > 1:
        dist.all_reduce(pvec, op=dist.ReduceOp.MAX)
    packed_s = float(pvec.item())
    step(0)
    torch.cuda.synchronize()
    assert torch.equal(hk, keep.cpu()) and torch.equal(hh, hits.cpu()) and torch.equal(ht, tot.cpu()), \
        "caller-packed path and device-pointer path disagree"

    # ---- ... and with the non-ACGT bits as the sparse list the library's own packers use (dcn_filter_batch_packed_sparse)
    _, exc_np, _ = A.pack_records_sparse(hb[0].numpy(), hoff.numpy().view(np.uint64))
    he = torch.from_numpy(np.ascontiguousarray(exc_np).view(np.int32).reshape(-1)).pin_memory() if len(exc_np) else None

    def sparse_step():
        gpu.filter_batch_packed_sparse_ptr(hc.data_ptr(), he.data_ptr() if he is not None else None, len(exc_np), None, hoff.data_ptr(),
                                           NR, True, 0, 2, 0.01, True, hk.data_ptr(), hh.data_ptr(), ht.data_ptr())

    hk.zero_()
    sparse_step()
    barrier()
    t0 = time.perf_counter()
    for i in range(p_steps):
        sparse_step()
    torch.cuda.synchronize()
    svec = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    s_h2d, s_d2h = gpu.last_transfer_bytes()
    if world > 1:
        dist.all_reduce(svec, op=dist.ReduceOp.MAX)
    sparse_s = float(svec.item())
    assert torch.equal(hk, keep.cpu()) and torch.equal(hh, hits.cpu()) and torch.equal(ht, tot.cpu()), \
        "caller-packed (sparse) path and device-pointer path disagree"
    # ---- the end-to-end leg once more with shards in proportion to each rank's share of the host (N > 1): does the rank with
    # the smallest share (on an 8-GPU box: four ranks at 20 GB/s, four at 35) hold the others back?  Measured: no -- 176.9
    # against 178.6 Gbp/s with equal shards: the host's aggregate rate is the limit, however the work is dealt out.
    balanced = None
    if world > 1:
        sh = torch.tensor([ingest["concurrent_gbs"]], dtype=torch.float64, device=dev)
        parts = [torch.empty_like(sh) for _ in range(world)]
        dist.all_gather(parts, sh)
        shares = [float(x.item()) for x in parts]
        my_pairs = max(2, int(NP * shares[rank] / max(shares)) & ~1)

        def balanced_step(i):
            gpu.filter_batch_ptr(hb[i % n_host].data_ptr(), hoff.data_ptr(), 2 * my_pairs, True, 0, 2, 0.01, True,
                                 hk.data_ptr(), hh.data_ptr(), ht.data_ptr())

        balanced_step(0)
        barrier()
        b_steps = max(1, min(e2e_steps, 5))
        t0 = time.perf_counter()
        for i in range(b_steps):
            balanced_step(i)
        torch.cuda.synchronize()
        bvec = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(bvec, op=dist.ReduceOp.MAX)
        tot_pairs = torch.tensor([my_pairs], dtype=torch.float64, device=dev)
        dist.all_reduce(tot_pairs, op=dist.ReduceOp.SUM)
        balanced = {"value": round(1e-9 * float(tot_pairs.item()) * 2 * READ_LEN * b_steps / float(bvec.item()), 3), "unit": "Gbp/s",
                    "steps": b_steps, "pairs_per_step_per_rank": [int(NP * x / max(shares)) & ~1 for x in shares],
                    "what": "dcn_filter_batch as in e2e, each rank's batch in proportion to its share of the host's concurrent H2D rate "
                            "(reported beside e2e, never instead of it)"}
    for i in range(1):
        e2e_step(e2e_steps - 1)   # leave the last e2e batch's result in the host buffers for the check below

    # parity of the last e2e step against the device-resident result of the same batch
    step((e2e_steps - 1) % n_host)
    torch.cuda.synchronize()
    assert torch.equal(hk, keep.cpu()) and torch.equal(hh, hits.cpu()) and torch.equal(ht, tot.cpu()), \
        "host-pointer path and device-pointer path disagree"

    # ---- random-access ceiling of the table (context for the lookup share of the kernel)
    nprobe, rms = gpu.measure_random_access(1 << 28)
    nprobe, rms = gpu.measure_random_access(1 << 28)

    # ---- cpu_baseline: the oracle (port of the reference path) on all host cores, rank 0, bounded sample
    cpu, oracle_idx = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, oracle_idx = cpu_baseline_leg(args, torch, gpu, batches[0], NP, keep_check=(step, keep, hits, tot))

    # ---- extra legs (N = 1): the other BASELINE.json configs and entry points, each with its own parity check
    extra = {}
    if rank == 0 and world == 1 and args.extras:
        ceiling = nprobe / (rms * 1e-3) / 1e9
        del hb, hc, hi
        legs = {
            "lookup": lambda: extra_lookup(torch, dev, gpu, n_keys, NP, 5, ceiling),
            "long": lambda: extra_long(torch, dev, gpu, genome, int(args.long_gbp * 1e9), 3, oracle_idx),
            "build": lambda: extra_build(torch, dev, gpu, genome, coff, G),
            "config5": lambda: extra_config5(torch, dev, local, NP, 4),
        }
        for name in [x for x in args.extras.split(",") if x]:
            t0 = time.time()
            try:
                if name == "config5":      # needs room for a second, larger index: let go of the big buffers first
                    batches.clear()
                    torch.cuda.empty_cache()
                extra[name] = legs[name]()
                extra[name]["leg_seconds"] = round(time.time() - t0, 1)
            except Exception as e:  # an extra leg must never take the contract line down with it
                extra[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()

    if rank == 0:
        gbp = 1e-9 * nb * args.steps * world
        value = gbp / (ms_total * 1e-3)
        e2e_value = 1e-9 * nb * e2e_steps * world / e2e_s
        peak, peak_src = measured_peaks()
        # algorithmic bytes of one fused-kernel launch: ASCII bases + record offsets + one 32-byte
        # sector per minimizer probed + 9 B of output per pair
        alg_bytes = nb * 1.0 + (NR + 1) * 8 + 32.0 * minim_per_step + 9.0 * NP
        fused_avg_ms = fused_ms / max(fused_n, 1)
        achieved = alg_bytes / (fused_avg_ms * 1e-3) / 1e9
        cap = ncu_capture() or {}
        ceiling = nprobe / (rms * 1e-3) / 1e9
        gprobes = minim_per_step / (fused_avg_ms * 1e-3) / 1e9
        out = {
            "metric": "filter Gbp/s (bit-exact decisions)", "value": round(value, 3), "unit": "Gbp/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "configs[1]: paired-end 2x150 bp, deplete, -a 2 -r 0.01, k=31 w=15",
                       "reference_mbp": args.genome_mbp, "contigs": CONTIGS, "index_minimizers": n_keys,
                       "pairs_per_step": NP, "pairs_total_per_gpu": NP * args.steps, "distinct_batches": n_batches,
                       "l2_policy": "inputs larger than L2 (1.5 GB batch, distinct batches cycled)",
                       "parallelism": f"read-sharded x{world}, index replicated",
                       "cpu_binding_rank0": f"{len(bound)} CPUs local to the GPU" if bound else "none",
                       "index_build_s": round(t_index, 3), "setup_s": round(t_setup, 2),
                       "table_bytes": gpu.index_info()["table_bytes"],
                       "device_api": "dcn_filter_batch_device_hint(max_unit_len = 300): enqueue only, no readback"},
            "e2e": {"value": round(e2e_value, 3), "unit": "Gbp/s", "h2d_bytes_per_step": int(e2e_h2d),
                    "d2h_bytes_per_step": int(e2e_d2h), "steps": e2e_steps, "host_buffers": "pinned",
                    "host_pack_threads": pack_threads if pack_threads is not None else os.environ.get("DCN_PACK_THREADS", "default"),
                    "caller_buffer_bytes_per_step": nb + (NR + 1) * 8,
                    # what the host can feed: pinned H2D GB/s of a rank alone and of all ranks at once (one ASCII base = one
                    # byte of it); packing on host threads is switched on while the ranks' own links are what binds
                    "ingest_ceiling": dict(ingest, ascii_route_gbp_per_s=ingest["concurrent_sum_gbs"],
                                           limiter=("the host the ranks share (DRAM / socket interconnect / PCIe root): its aggregate rate under the "
                                                    "pipeline is ~0.8 of the copy-only probe's sum, with equal shards and with shards in proportion to the "
                                                    "ranks' shares alike (e2e_balanced_shards); ASCII route, two packer threads on the ranks well below the "
                                                    "mean share (a packed base costs 1.5 B of host DRAM traffic, a copied one 1.0 B)"
                                                    if ingest["concurrent_sum_gbs"] >= par.HOST_BUSY_GBS
                                                    else "each rank's own PCIe link and the host memory traffic of its packer threads (DESIGN.md 5)")),
                    "frac_of_ascii_ceiling": round(e2e_value / max(ingest["concurrent_sum_gbs"], 1e-9), 3),
                    # e2e against world x the smallest share of the copy-only probe (> 1: the shares shift under the real pipeline)
                    "frac_of_slowest_rank_ceiling": round(e2e_value / max(world * ingest.get("concurrent_min_gbs", ingest["concurrent_gbs"]), 1e-9), 3),
                    "pipeline": "arena (dcn_api.cu filter_pipeline_arena): copies land in one device arena, kernels run over whatever "
                                "contiguous range of units has arrived" if (pack_threads is None or pack_threads > 0) and not os.environ.get("DCN_PIPELINE")
                                else "chunks (one kernel chain per 32 MB copy)",
                    "api": "dcn_filter_batch (C ABI), ASCII records + u64 offsets in host memory; bytes as counted by the "
                           "library (dcn_last_transfer_bytes) for the last step: part of the batch crosses as ASCII, part is packed "
                           "by host threads inside the call (the split is dynamic), and the offsets of "
                           "equal-length chunks are written on the device, not copied"},
            "e2e_packed_input": {"value": round(1e-9 * nb * p_steps * world / packed_s, 3), "unit": "Gbp/s",
                                 "h2d_bytes_per_step": int(p_h2d), "d2h_bytes_per_step": int(p_d2h), "steps": p_steps,
                                 "caller_buffer_bytes_per_step": int(codes_np.nbytes + inv_np.nbytes + (NR + 1) * 8),
                                 "api": "dcn_filter_batch_packed (C ABI): 2-bit codes + non-ACGT bits packed by the caller "
                                        "(packing time not included)"},
            "e2e_packed_sparse_input": {"value": round(1e-9 * nb * p_steps * world / sparse_s, 3), "unit": "Gbp/s",
                                        "h2d_bytes_per_step": int(s_h2d), "d2h_bytes_per_step": int(s_d2h), "steps": p_steps,
                                        "caller_buffer_bytes_per_step": int(codes_np.nbytes + exc_np.nbytes + (NR + 1) * 8),
                                        "api": "dcn_filter_batch_packed_sparse (C ABI): 2-bit codes + (block, mask) list of the non-ACGT "
                                               "bases packed by the caller with dcn_pack_records_sparse (packing time not included)"},
            **({"e2e_balanced_shards": balanced} if balanced else {}),
            "gpu_launches": int(launches),
            # `bound` keeps the contract's vocabulary: the path is nominally HBM work.  What actually binds it is stated
            # next to it, each as a fraction measured by THIS run unless it is prefixed ncu_ (then it is read from the
            # committed capture of the same kernel on the same workload, named in ncu_capture).
            "roofline": {"bound": "hbm", "kernel": "filter_warp_kernel<ASCII> (+ filter_tail_kernel: idle on this workload)",
                         "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": cap.get("dram_bytes_per_launch"),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes),
                         "kernel_ms_per_launch": round(fused_avg_ms, 4), "kernel_share_of_step": round(fused_ms / ms_total, 4),
                         "minimizers_per_bp": round(minim_per_step / nb, 5),
                         "binding": "random 32-byte-sector rate of HBM (every probe costs ~4 sectors of DRAM traffic) and the "
                                    "integer ALU pipe, both above 70 %; streaming HBM bandwidth is not the limit",
                         "frac_hbm": round(achieved / peak, 4),
                         "lookup_gprobes_per_s": round(gprobes, 3),
                         "random_sector_ceiling_gsectors_per_s": round(ceiling, 3),
                         "frac_random_sector": round(gprobes / ceiling, 4),
                         "frac_hbm_traffic": round(cap["dram_bytes_per_launch"] / (fused_avg_ms * 1e-3) / 1e9 / peak, 4) if cap.get("dram_bytes_per_launch") else None,
                         "ncu_alu_pipe_pct_of_peak": cap.get("alu_pipe_pct_of_peak"),
                         "ncu_issue_active_pct": cap.get("issue_active_pct"),
                         "ncu_capture": cap.get("report"),
                         "ncu_kernel_ms": cap.get("kernel_ms")},
            "cpu_baseline": cpu,
            "extra": extra,
            "clocks": clk,
            "counters": counters,
            "kept_pairs_last_step": kept_last,
        }
        _emit(out)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_leg(args, torch, gpu, batch0, NP, keep_check=None):
    """Oracle (port) on every host core over a bounded sample of the same workload; also a full
    bit-exact parity check of that sample against the GPU result."""
    from oracle import oracle as O
    threads = os.cpu_count() or 1
    S = int(min(args.cpu_sample_pairs, NP))
    t0 = time.time()
    nkeys = gpu.index_info()["n_keys"]
    keys_t = torch.empty(nkeys, dtype=torch.int64)
    gpu._check(gpu._lib.dcn_index_build_keys(gpu._ctx, keys_t.data_ptr(), nkeys))
    keys = keys_t.numpy().view(np.uint64)
    idx = O.IndexSet(keys, threads=threads)
    t_set = time.time() - t0
    hb = batch0[:S * 2 * READ_LEN].cpu().numpy()
    ho = (np.arange(2 * S + 1, dtype=np.uint64) * np.uint64(READ_LEN))
    t0 = time.perf_counter()
    ok, oh, ot = O.filter_batch(idx, hb, ho, paired=True, abs_thr=2, rel_thr=0.01, deplete=True, threads=threads)
    dt = time.perf_counter() - t0
    parity = None
    if keep_check is not None:
        step, keep, hits, tot = keep_check
        step(0)
        torch.cuda.synchronize()
        parity = bool(np.array_equal(keep[:S].cpu().numpy(), ok) and np.array_equal(hits[:S].cpu().numpy().view(np.uint32), oh)
                      and np.array_equal(tot[:S].cpu().numpy().view(np.uint32), ot))
        assert parity, "GPU result differs from the oracle on the cpu_baseline sample"
    return {"value": round(1e-9 * S * 2 * READ_LEN / dt, 4), "unit": "Gbp/s", "cores": threads, "kind": "port",
            "sample": f"first {S} pairs of batch 0 ({S * 2 * READ_LEN / 1e6:.0f} Mbp), oracle/deacon_oracle.c on {threads} threads, "
                      f"{dt:.2f} s; index set built in {t_set:.1f} s (not timed)",
            "parity_vs_gpu_on_sample": parity, "reference_published": PUBLISHED_NOTE}, idx


def run_reference(args):
    """The reference's own CPU path for this metric.  The reference is Rust and cannot be built in
    this image (no cargo/rustc), so this times the oracle restatement (kind = "port")
    on all host cores: extraction + FxHashSet-like lookup + classification, per step a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import oracle as O
    threads = os.cpu_count() or 1
    G = int(args.genome_mbp * 1e6)
    S = int(args.ref_pairs_per_step)
    use_cuda = torch.cuda.is_available()
    dev = torch.device("cuda", 0) if use_cuda else torch.device("cpu")
    t0 = time.time()
    genome = make_genome(torch, dev, G, args.seed)          # synthetic data only; the timed path is CPU
    n_b = max(1, min(args.batches, args.steps + args.warmup))
    batches = [make_pairs(torch, dev, genome, S, args.seed + 1 + b).cpu().numpy() for b in range(n_b)]
    g_host = genome.cpu().numpy()
    del genome
    coff = contig_offsets(G, args.seed).astype(np.uint64)
    t_data = time.time() - t0
    t0 = time.time()
    idx = O.index_build((g_host, coff), 31, 15, 0.0, threads=threads)   # CPU index build (oracle), not timed
    t_index = time.time() - t0
    ho = (np.arange(2 * S + 1, dtype=np.uint64) * np.uint64(READ_LEN))
    nb = 2 * S * READ_LEN

    def step(i):
        return O.filter_batch(idx, batches[i % n_b], ho, paired=True, abs_thr=2, rel_thr=0.01, deplete=True, threads=threads)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        k, h, t = step(args.warmup + i)
    dt = time.perf_counter() - t0
    value = 1e-9 * nb * args.steps / dt
    out = {"impl": "reference", "metric": "filter Gbp/s (bit-exact decisions)", "value": round(value, 4), "unit": "Gbp/s",
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u64", "data": "synthetic",
           "config": {"workload": "configs[1]: paired-end 2x150 bp, deplete, -a 2 -r 0.01, k=31 w=15",
                      "reference_mbp": args.genome_mbp, "contigs": CONTIGS, "index_minimizers": len(idx),
                      "pairs_per_step": S, "index_build_s": round(t_index, 1), "data_s": round(t_data, 1)},
           "cpu_baseline": {"value": round(value, 4), "unit": "Gbp/s", "cores": threads, "kind": "port",
                            "sample": f"{S} pairs ({nb / 1e6:.0f} Mbp) per step, oracle/deacon_oracle.c on {threads} threads "
                                      "(the Rust reference cannot be built here: no cargo/rustc)",
                            "reference_published": PUBLISHED_NOTE},
           "e2e": {"value": round(value, 4), "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(out)


def _emit(obj):
    """The one JSON line of the contract goes to the real stdout; everything else any library prints on
    fd 1 while we run (NCCL prints its version banner there) is sent to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


if __name__ == "__main__":
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
