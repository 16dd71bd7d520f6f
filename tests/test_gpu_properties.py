"""GPU property tests at (or near) the sizes BASELINE.json names, where the oracle would take too long:
size-independent identities the reference's semantics imply.  Data is generated on the GPU with torch
(seeded); every compute call goes through the C ABI.

  strand invariance     canonical k-mers + canonical minimizers: a read and its reverse complement give the same
                        (hits, total, keep) - what tests/filter_tests.rs:586-657 checks on one 60-mer
  batch invariance      decisions do not depend on how a batch is cut (chunks, shards, device vs host pointers)
  B3 -> B2 == B1        client extraction + server lookup == local filter (src/remote_filter.rs:762-790)
  set algebra           build(A) U build(B) == build(A ++ B); (A U B) - B == A - B; encode -> decode round trip
  counters              checksum of the per-unit outputs == the six device-side counters
"""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

G_BP = 200_000_000        # 200 Mbp reference -> ~25 M keys (config 2 shape, scaled so that the suite stays fast)
N_PAIRS = 2_000_000       # 600 Mbp of reads per call


def _dev():
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def world(gpu):
    from deacon_server_b200 import IndexHeader  # noqa: F401
    dev = _dev()
    gen = torch.Generator(device=dev)
    gen.manual_seed(77)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    genome = lut[torch.randint(0, 4, (G_BP,), device=dev, generator=gen)]
    coff = torch.tensor([0, G_BP // 3, G_BP // 2, G_BP], dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    n_keys = gpu.index_build_device(genome, coff, 3, G_BP, 31, 15, 0.0, True, stream=st)
    # pairs: mate 1 forward at p, mate 2 reverse complement ending at p + insert; 10 % random pairs; 0.5 % substitutions
    comp = torch.zeros(256, dtype=torch.uint8, device=dev)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    ar = torch.arange(150, device=dev)
    pos = torch.randint(0, G_BP - 600, (N_PAIRS,), device=dev, generator=gen)
    ins = torch.randint(300, 400, (N_PAIRS,), device=dev, generator=gen)
    m1 = genome[pos[:, None] + ar[None, :]]
    m2 = comp[genome[(pos + ins)[:, None] - 1 - ar[None, :]].long()]
    reads = torch.stack([m1, m2], 1).reshape(2 * N_PAIRS, 150).contiguous()
    rnd = torch.rand(N_PAIRS, device=dev, generator=gen) < 0.1
    reads.view(N_PAIRS, 300)[rnd] = lut[torch.randint(0, 4, (int(rnd.sum()), 300), device=dev, generator=gen)]
    sub = torch.rand(2 * N_PAIRS, 150, device=dev, generator=gen) < 0.005
    reads[sub] = lut[torch.randint(0, 4, (int(sub.sum()),), device=dev, generator=gen)]
    nmask = torch.rand(2 * N_PAIRS, device=dev, generator=gen) < 0.001           # a few reads carry an N
    reads[nmask, 75] = ord("N")
    return dict(genome=genome, coff=coff, n_keys=n_keys, reads=reads, comp=comp, stream=st)


def _filter_dev(gpu, reads2d, paired=True, deplete=True):
    dev = _dev()
    n_rec, ln = reads2d.shape
    bases = reads2d.reshape(-1).contiguous()
    if bases.data_ptr() % 16:                       # the device-pointer entry wants 16-byte aligned bases
        bases = bases.clone()
    off = torch.arange(n_rec + 1, device=dev, dtype=torch.int64) * ln
    nu = n_rec // 2 if paired else n_rec
    keep = torch.zeros(nu, dtype=torch.uint8, device=dev)
    hits = torch.zeros(nu, dtype=torch.int32, device=dev)
    tot = torch.zeros(nu, dtype=torch.int32, device=dev)
    gpu.filter_batch_device(bases, off, n_rec, bases.numel(), keep, hits, tot, paired=paired, deplete=deplete,
                            stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return keep, hits, tot


def test_strand_invariance_full_batch(gpu, world):
    reads = world["reads"]
    k1, h1, t1 = _filter_dev(gpu, reads)
    rc = world["comp"][reads.flip(1).long()]                      # reverse complement of every mate
    k2, h2, t2 = _filter_dev(gpu, rc)
    # An N breaks the symmetry by design: packed-seq's lossy code maps it to G on BOTH strands
    # (src/filter_common.rs:238), so windows near it may pick differently; every other pair must agree exactly.
    clean = ~(reads == ord("N")).any(1).view(-1, 2).any(1)
    assert int((~clean).sum()) > 1000
    assert torch.equal(h1[clean], h2[clean]) and torch.equal(t1[clean], t2[clean]) and torch.equal(k1[clean], k2[clean])
    # mates swapped: pooled counts are symmetric in the two mates (src/filter_common.rs:172-198), N or not
    sw = reads.view(-1, 2, 150).flip(1).reshape(-1, 150)
    k3, h3, t3 = _filter_dev(gpu, sw)
    assert torch.equal(h1, h3) and torch.equal(t1, t3) and torch.equal(k1, k3)
    assert 0.85 < float(k1.float().mean()) * 10 < 1.15              # ~10 % of the pairs are non-host and kept under --deplete
    assert int(h1.max()) >= 20 and int(t1.max()) <= 2 * 106


def test_batch_cut_invariance_and_counters(gpu, world):
    from deacon_server_b200 import parallel as P
    reads = world["reads"]
    gpu.stats_reset()
    k, h, t = _filter_dev(gpu, reads)
    st = gpu.stats()
    off = np.arange(reads.shape[0] + 1, dtype=np.uint64) * np.uint64(150)
    assert st == P.counters_of(off, k.cpu().numpy(), paired=True)       # checksum of the decisions == device counters
    # cut into three uneven device-resident parts
    cuts = [0, 333_334, 1_200_001, N_PAIRS]
    parts = [_filter_dev(gpu, reads[2 * a:2 * b]) for a, b in zip(cuts, cuts[1:])]
    assert torch.equal(torch.cat([p[0] for p in parts]), k) and torch.equal(torch.cat([p[1] for p in parts]), h)
    # host-pointer pipeline (32 MB chunks, both ingest routes) on a 1 M-pair slice
    n = 1_000_000
    hb = reads[:2 * n].reshape(-1).cpu().numpy()
    ho = off[:2 * n + 1].copy()
    for threads in (0, 8):
        gpu.host_pack_threads(threads)
        kk, hh, tt = gpu.filter_batch(hb, ho, paired=True, deplete=True)
        assert np.array_equal(kk, k[:n].cpu().numpy()) and np.array_equal(hh.view(np.int32), h[:n].cpu().numpy())
        assert np.array_equal(tt.view(np.int32), t[:n].cpu().numpy())
    # two-rank sharding of that slice == the unsharded result
    got = []
    for r in range(2):
        sb, so, u0, u1 = P.shard_batch(hb, ho, True, r, 2)
        got.append(gpu.filter_batch(sb, so, paired=True, deplete=True)[0])
    assert np.array_equal(np.concatenate(got), k[:n].cpu().numpy())
    # single-end: every mate on its own; a pair's total is the sum of its mates' totals (src/local_filter.rs:263)
    ks, hs, ts = _filter_dev(gpu, reads[:2 * n], paired=False, deplete=False)
    assert torch.equal(ts.view(-1, 2).sum(1), t[:n])
    assert bool((hs.view(-1, 2).sum(1) >= h[:n]).all()) and bool((hs.view(-1, 2).max(1).values <= h[:n]).all())


def test_extract_then_lookup_equals_filter_large(gpu, world):
    n = 500_000
    reads = world["reads"][:2 * n]
    k1, h1, t1 = _filter_dev(gpu, reads)
    hb = reads.reshape(-1).cpu().numpy()
    ho = np.arange(2 * n + 1, dtype=np.uint64) * np.uint64(150)
    hh, pp, oo = gpu.extract(hb, ho, 0, 31, 15, 0, cap=int(0.12 * len(hb)))
    assert np.array_equal(np.diff(oo.astype(np.int64)).reshape(-1, 2).sum(1), t1.cpu().numpy())      # totals == extraction counts
    assert bool((pp <= 150 - 31).all())
    k2, h2, t2 = gpu.lookup_batch(hh, oo[::2].copy(), 2, 0.01, True)
    assert np.array_equal(k2, k1.cpu().numpy()) and np.array_equal(h2.view(np.int32), h1.cpu().numpy())


def test_device_resident_client_server_chain(gpu, world):
    """dcn_extract_device -> dcn_lookup_batch_device -> dcn_stats_accumulate_device (the GPU-resident form of the
    remote engine, BASELINE configs[4]) == the fused local filter and its counters."""
    dev = _dev()
    reads = world["reads"][:2 * 1_000_000]
    n_rec = reads.shape[0]
    nu = n_rec // 2
    bases = reads.reshape(-1).contiguous()
    off = torch.arange(n_rec + 1, device=dev, dtype=torch.int64) * 150
    st = torch.cuda.current_stream().cuda_stream
    gpu.stats_reset()
    k1, h1, t1 = _filter_dev(gpu, reads)
    want = gpu.stats()
    d_h = torch.empty(int(0.11 * bases.numel()), dtype=torch.int64, device=dev)
    d_p = torch.empty(d_h.numel(), dtype=torch.int32, device=dev)
    d_o = torch.empty(n_rec + 1, dtype=torch.int64, device=dev)
    m = gpu.extract_device(bases, off, n_rec, bases.numel(), d_h, d_p, d_o, stream=st)
    torch.cuda.synchronize()
    assert m == int(d_o[-1]) == int(t1.sum())
    keep = torch.zeros(nu, dtype=torch.uint8, device=dev)
    hits = torch.zeros(nu, dtype=torch.int32, device=dev)
    tot = torch.zeros(nu, dtype=torch.int32, device=dev)
    gpu.lookup_batch_device(d_h, d_o[::2].contiguous(), nu, keep, hits, tot, 2, 0.01, True, stream=st)
    gpu.stats_reset()
    gpu.stats_accumulate_device(off, n_rec, True, keep, stream=st)
    assert torch.equal(keep, k1) and torch.equal(hits, h1) and torch.equal(tot, t1)
    assert gpu.stats() == want
    from deacon_server_b200 import DeaconCudaError
    with pytest.raises(DeaconCudaError, match="out_cap"):
        gpu.extract_device(bases, off, n_rec, bases.numel(), d_h[:1000], d_p, d_o, stream=st)


def test_index_set_algebra_large(gpu, world):
    genome, coff, st = world["genome"], world["coff"], world["stream"]
    n_all = world["n_keys"]
    dev = _dev()
    # A = contigs 0,1 ; B = contigs 1,2 (contig 1 shared)
    a_off = torch.tensor([0, int(coff[1]), int(coff[2])], dtype=torch.int64, device=dev)
    n_a = gpu.index_build_device(genome, a_off, 2, int(coff[2]), 31, 15, 0.0, False, stream=st)
    a_idx = gpu.idx_encode()
    b_lo = (int(coff[1]) // 16) * 16
    gb = genome[b_lo:]
    b_off = torch.tensor([0, int(coff[2]) - b_lo, G_BP - b_lo], dtype=torch.int64, device=dev)
    # records of B start at the 16-aligned offset b_lo <= coff[1]: a few extra bases of contig 0's tail belong to A as well
    n_b = gpu.index_build_device(gb, b_off, 2, G_BP - b_lo, 31, 15, 0.0, False, stream=st)
    b_idx = gpu.idx_encode()
    hdr, nf, ns = gpu.idx_decode(a_idx)
    assert nf == ns == n_a and len(a_idx) >= 9 * n_a
    n_union = gpu.index_union(b_idx)
    assert max(n_a, n_b) < n_union <= n_a + n_b
    union_keys = gpu.working_keys()
    assert bool((np.diff(union_keys.astype(np.float64)) > 0).all()) or bool((union_keys[1:] > union_keys[:-1]).all())
    # (A U B) - B == A - B, and |A - B| + |A n B| == |A|
    n_ab = gpu.index_diff(b_idx)
    gpu.idx_decode(a_idx)
    assert gpu.index_diff(b_idx) == n_ab
    a_minus_b = gpu.working_keys()
    gpu.idx_decode(a_idx)
    a_keys = gpu.working_keys()
    assert len(np.intersect1d(a_keys, a_minus_b)) == n_ab == len(a_minus_b)
    # encode -> decode round trip is the identity on the set, byte-stable on re-encode
    again = gpu.idx_encode()
    assert again == a_idx
    # build over all three contigs contains A and B (chunk seams between contigs can only ADD windows that span nothing here,
    # since records are independent: the union of per-contig builds equals the build of all contigs)
    gpu.idx_decode(a_idx)
    gpu.index_union(b_idx)
    assert gpu.working_set_info()["n_keys"] >= n_all
    # restore the resident index for any later test
    gpu.index_build_device(genome, coff, 3, G_BP, 31, 15, 0.0, True, stream=st)
