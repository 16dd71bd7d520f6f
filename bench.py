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


def ncu_traffic():
    """Per-launch DRAM bytes of the fused kernel from the committed ncu capture, if it matches this workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


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
    pack_threads = None   # library default (hardware threads this process may use - 4, at most 16)
    if world > 1 and not os.environ.get("DCN_PACK_THREADS"):
        pack_threads = par.pack_threads_for_rank(int(os.environ.get("LOCAL_WORLD_SIZE", world)))
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
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg(args, torch, gpu, batches[0], NP, keep_check=(step, keep, hits, tot))

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
        traffic = ncu_traffic()
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
                       "table_bytes": gpu.index_info()["table_bytes"]},
            "e2e": {"value": round(e2e_value, 3), "unit": "Gbp/s", "h2d_bytes_per_step": int(e2e_h2d),
                    "d2h_bytes_per_step": int(e2e_d2h), "steps": e2e_steps, "host_buffers": "pinned",
                    "host_pack_threads": pack_threads if pack_threads is not None else os.environ.get("DCN_PACK_THREADS", "default"),
                    "caller_buffer_bytes_per_step": nb + (NR + 1) * 8,
                    "api": "dcn_filter_batch (C ABI), ASCII records + u64 offsets in host memory; bytes as counted by the "
                           "library (dcn_last_transfer_bytes) for the last step: part of the batch crosses as ASCII, part is packed "
                           "to 0.43 B/bp by host threads inside the call (the split is dynamic), and the offsets of "
                           "equal-length chunks are written on the device, not copied"},
            "e2e_packed_input": {"value": round(1e-9 * nb * p_steps * world / packed_s, 3), "unit": "Gbp/s",
                                 "h2d_bytes_per_step": int(p_h2d), "d2h_bytes_per_step": int(p_d2h), "steps": p_steps,
                                 "caller_buffer_bytes_per_step": int(codes_np.nbytes + inv_np.nbytes + (NR + 1) * 8),
                                 "api": "dcn_filter_batch_packed (C ABI): 2-bit codes + non-ACGT bits packed by the caller "
                                        "(packing time not included)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "filter_fused_kernel<Geo<31,15>>", "achieved": round(achieved, 2),
                         "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes),
                         "kernel_ms_per_launch": round(fused_avg_ms, 4), "kernel_share_of_step": round(fused_ms / ms_total, 4),
                         "minimizers_per_bp": round(minim_per_step / nb, 5),
                         "lookup_gprobes_per_s": round(minim_per_step / (fused_avg_ms * 1e-3) / 1e9, 3),
                         "random_sector_ceiling_gsectors_per_s": round(nprobe / (rms * 1e-3) / 1e9, 3),
                         "int_alu_pipe_pct_of_peak": traffic.get("alu_pipe_pct_of_peak") if traffic else None,
                         "issue_slots_pct": traffic.get("issue_active_pct") if traffic else None,
                         "note": "integer-ALU bound, not HBM bound (ncu: profiles/r1_final_fused_ncu_summary.json, captured one "
                                 "scheduling change earlier at 8.10 ms/launch: traffic and pipe percentages are that launch's); "
                                 "the lookup kernel alone (dcn_lookup_batch) runs at 89 % of the random-sector ceiling: DESIGN.md"},
            "cpu_baseline": cpu,
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
            "parity_vs_gpu_on_sample": parity}


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
                                      "(the Rust reference cannot be built here: no cargo/rustc)"},
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
