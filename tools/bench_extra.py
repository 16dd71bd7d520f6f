"""Measurements beside the contract bench (bench.py): the other BASELINE.json configs and entry points,
device-resident, CUDA events.  Prints one JSON object per line; results are quoted in DESIGN.md.
  --what lookup   B2 dcn_lookup_batch_device on pre-hashed pairs (config 5 server path): probes/s vs the random-sector ceiling
  --what long     config 3: ONT-like reads (gamma(2) lengths, mean 10 kbp, 5 % substitutions), search mode
  --what build    config 4: index build with -e 0.5 on a reference with low-complexity inserts
  --what idx      .idx container: GPU encode of the built key set and GPU decode + table build of that file
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
import deacon_server_b200 as d  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="lookup,long,build,idx")
ap.add_argument("--genome-mbp", type=float, default=3100.0)
ap.add_argument("--pairs-m", type=float, default=5.0)
ap.add_argument("--long-gbp", type=float, default=2.0, help="bases of long reads per step")
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
what = args.what.split(",")

dev = torch.device("cuda", 0)
G = int(args.genome_mbp * 1e6)
genome = B.make_genome(torch, dev, G, 20261018)
coff = torch.from_numpy(B.contig_offsets(G, 20261018)).to(dev)
gpu = d.DeaconGpu(0)
st = torch.cuda.current_stream().cuda_stream
n_keys = gpu.index_build_device(genome, coff, B.CONTIGS, G, 31, 15, 0.0, True, stream=st)
torch.cuda.synchronize()


def timed(fn, steps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


if "lookup" in what:
    # pre-hashed pairs: the server's request shape (src/remote_filter.rs:266-301).  Hash lists are made from the
    # index's own keys (hits) and random values (misses), ~28 per pair like 2x150 bp reads.
    NP = int(args.pairs_m * 1e6)
    rng = torch.Generator(device=dev); rng.manual_seed(7)
    per = torch.randint(24, 33, (NP,), device=dev, generator=rng)
    off = torch.zeros(NP + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(per, 0)
    nh = int(off[-1])
    keys_ptr = gpu.index_build_keys_ptr()
    keys = torch.empty(n_keys, dtype=torch.int64, device=dev)
    import ctypes
    torch.cuda.synchronize()
    ctypes.CDLL("libcudart.so.12").cudaMemcpy(ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(keys_ptr), ctypes.c_size_t(n_keys * 8), 3)
    sel = torch.randint(0, n_keys, (nh,), device=dev, generator=rng)
    hashes = keys[sel]
    miss = torch.rand(nh, device=dev, generator=rng) < 0.2
    hashes[miss] = torch.randint(-2**62, 2**62, (int(miss.sum()),), dtype=torch.int64, device=dev, generator=rng)
    del sel, keys
    keep = torch.zeros(NP, dtype=torch.uint8, device=dev)
    hits = torch.zeros(NP, dtype=torch.int32, device=dev)
    tot = torch.zeros(NP, dtype=torch.int32, device=dev)

    def step():
        rc = gpu._lib.dcn_lookup_batch_device(gpu._ctx, hashes.data_ptr(), off.data_ptr(), NP, 2, 0.01, 1, keep.data_ptr(),
                                              hits.data_ptr(), tot.data_ptr(), st)
        assert rc == 0

    ms = timed(step, args.steps)
    n, rms = gpu.measure_random_access(1 << 28)
    n, rms = gpu.measure_random_access(1 << 28)
    ceil = n / rms / 1e6
    print(json.dumps({"what": "lookup (B2, device-resident)", "records": NP, "hashes": nh, "ms": round(ms, 3),
                      "gprobes_per_s": round(nh / ms / 1e6, 2), "random_sector_ceiling_gsectors_per_s": round(ceil, 2),
                      "frac_of_ceiling": round(nh / ms / 1e6 / ceil, 3), "index_keys": n_keys,
                      "equiv_gbp_per_s_at_0.0942_minimizers_per_bp": round(nh / 0.0942 / ms / 1e6, 1),
                      "hit_fraction": round(float(hits.sum()) / nh, 3)}))
    del hashes

if "long" in what:
    total = int(args.long_gbp * 1e9)
    rs = np.random.default_rng(5)
    lens = np.clip(rs.gamma(2.0, 5000.0, int(total / 10000 * 1.2)), 200, 200_000).astype(np.int64)
    lens = lens[np.cumsum(lens) <= total]
    n = len(lens)
    off_h = np.zeros(n + 1, np.int64); off_h[1:] = np.cumsum(lens)
    nb = int(off_h[-1])
    off = torch.from_numpy(off_h).to(dev)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    bases = torch.empty(nb, dtype=torch.uint8, device=dev)
    rng = torch.Generator(device=dev); rng.manual_seed(9)
    starts = torch.randint(0, G - 200_001, (n,), device=dev, generator=rng)
    host = torch.rand(n, device=dev, generator=rng) < 0.5
    # gather per record in slabs
    rec_of = torch.repeat_interleave(torch.arange(n, device=dev), torch.from_numpy(lens).to(dev))
    pos_in = torch.arange(nb, device=dev) - off[:-1][rec_of]
    src = starts[rec_of] + pos_in
    bases = genome[src]
    rnd = ~host[rec_of]
    bases[rnd] = lut[torch.randint(0, 4, (int(rnd.sum()),), device=dev, generator=rng)]
    sub = torch.rand(nb, device=dev, generator=rng) < 0.05
    bases[sub] = lut[torch.randint(0, 4, (int(sub.sum()),), device=dev, generator=rng)]
    del rec_of, pos_in, src, rnd, sub
    pad = (-nb) % 16
    if pad:
        bases = torch.cat([bases, torch.zeros(pad, dtype=torch.uint8, device=dev)])
    keep = torch.zeros(n, dtype=torch.uint8, device=dev)
    hits = torch.zeros(n, dtype=torch.int32, device=dev)
    tot = torch.zeros(n, dtype=torch.int32, device=dev)

    def step():
        gpu.filter_batch_device(bases, off, n, nb, keep, hits, tot, paired=False, deplete=False, stream=st)

    ms = timed(step, args.steps)
    print(json.dumps({"what": "config 3: ONT-like long reads, search mode (device-resident)", "reads": n, "bases": nb,
                      "mean_len": round(nb / n), "ms": round(ms, 3), "gbp_per_s": round(nb / ms / 1e6, 2),
                      "minimizers_per_bp": round(float(tot.sum()) / nb, 4), "kept": int(keep.sum()),
                      "gprobes_per_s": round(float(tot.sum()) / ms / 1e6, 2)}))
    # the same batch end to end through dcn_filter_batch from pinned host buffers (two-route ingest)
    hb = bases[:nb].cpu().pin_memory()
    ho = off.cpu().pin_memory()
    hk = torch.zeros(n, dtype=torch.uint8).pin_memory()
    hh = torch.zeros(n, dtype=torch.int32).pin_memory()
    ht = torch.zeros(n, dtype=torch.int32).pin_memory()

    def e2e_step():
        gpu.filter_batch_ptr(hb.data_ptr(), ho.data_ptr(), n, False, 0, 2, 0.01, False, hk.data_ptr(), hh.data_ptr(), ht.data_ptr())

    for _ in range(3):
        e2e_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    dt = (time.perf_counter() - t0) / args.steps
    assert torch.equal(hk, keep.cpu()) and torch.equal(hh, hits.cpu()) and torch.equal(ht, tot.cpu())
    h2d, d2h = gpu.last_transfer_bytes()
    print(json.dumps({"what": "config 3 end to end (dcn_filter_batch, pinned host buffers)", "bases": nb, "ms": round(dt * 1e3, 2),
                      "gbp_per_s": round(nb / dt / 1e9, 2), "h2d_bytes": int(h2d), "d2h_bytes": int(d2h),
                      "results": "identical to the device-resident call"}))
    del bases

if "build" in what:
    g2 = genome.clone()
    rs = np.random.default_rng(11)
    n_ins = int(G * 0.02 / 300)
    pos = torch.from_numpy(rs.integers(0, G - 400, n_ins)).to(dev)
    ar = torch.arange(300, device=dev)
    kind = torch.from_numpy(rs.integers(0, 2, n_ins)).to(dev)
    homo = torch.tensor([65, 84], dtype=torch.uint8, device=dev)[kind][:, None].expand(n_ins, 300)
    dinuc = torch.tensor([[65, 67], [71, 84]], dtype=torch.uint8, device=dev)[kind][:, ar % 2]
    pick = torch.from_numpy(rs.integers(0, 2, n_ins)).to(dev).bool()
    ins = torch.where(pick[:, None], homo, dinuc)
    g2[(pos[:, None] + ar[None, :]).reshape(-1)] = ins.reshape(-1)
    torch.cuda.synchronize()
    res = {}
    for thr in (0.0, 0.5):
        t0 = time.perf_counter()
        nk = gpu.index_build_device(g2, coff, B.CONTIGS, G, 31, 15, thr, False, stream=st)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        nk = gpu.index_build_device(g2, coff, B.CONTIGS, G, 31, 15, thr, False, stream=st)
        torch.cuda.synchronize()
        res[str(thr)] = {"keys": nk, "seconds": round(time.perf_counter() - t1, 4), "first_call_seconds": round(t1 - t0, 4)}
    print(json.dumps({"what": "config 4: index build (extract + radix sort + unique), 2 % low-complexity inserts",
                      "reference_mbp": args.genome_mbp, "by_entropy_threshold": res}))

if "idx" in what:
    import ctypes as C
    nk = gpu.index_build_device(genome, coff, B.CONTIGS, G, 31, 15, 0.0, False, stream=st)
    torch.cuda.synchronize()
    ln = C.c_uint64()
    gpu._lib.dcn_idx_encode(gpu._ctx, None, 0, C.byref(ln))
    buf = torch.empty(ln.value, dtype=torch.uint8).pin_memory()
    t0 = time.perf_counter()
    gpu._check(gpu._lib.dcn_idx_encode(gpu._ctx, buf.data_ptr(), buf.numel(), C.byref(ln)))
    t_enc = time.perf_counter() - t0
    ver, k, w, nf, ns = C.c_uint8(), C.c_uint8(), C.c_uint8(), C.c_uint64(), C.c_uint64()
    t0 = time.perf_counter()
    gpu._check(gpu._lib.dcn_idx_decode(gpu._ctx, buf.data_ptr(), buf.numel(), 0, 1, C.byref(ver), C.byref(k), C.byref(w), C.byref(nf), C.byref(ns)))
    t_dec = time.perf_counter() - t0
    assert nf.value == ns.value == nk
    print(json.dumps({"what": ".idx container on the GPU (pinned host buffer)", "keys": nk, "file_bytes": ln.value,
                      "encode_s": round(t_enc, 3), "decode_sort_upload_table_s": round(t_dec, 3),
                      "decode_gb_per_s": round(ln.value / t_dec / 1e9, 1)}))
