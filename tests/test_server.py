"""The wire front end (deacon_server_b200/server.py; reference: src/server.rs, src/server_common.rs, tests/server_tests.rs).

CPU part: HTTP plumbing, JSON and binary schemas against a scripted engine (a test double that records its calls).
GPU part (-m gpu): the real engine behind the same endpoints against the oracle's lookups."""
import hashlib
import json
import threading
import urllib.error
import urllib.request

import numpy as np
import pytest

from deacon_server_b200 import server as S


class ScriptedEngine:
    """Answers like a resident 7-key index of k=31, w=15 where every even hash is a hit; records what it was asked."""

    def __init__(self):
        self.calls = []

    def index_info(self):
        return {"n_keys": 7, "kmer_length": 31, "window_size": 15, "table_bytes": 0}

    def _decide(self, lists, abs_threshold, rel_threshold, deplete):
        out = []
        for h in lists:
            hits = len({int(x) for x in h if int(x) % 2 == 0})
            req = max(abs_threshold, max(1, round(rel_threshold * len(h))) if len(h) else 0)
            out.append(((hits < req) if deplete else (hits >= req), hits, len(h)))
        return out

    def unpaired_should_keep(self, recs, kmer_length, abs_threshold, rel_threshold, deplete, debug=False):
        self.calls.append(("unpaired", len(recs), kmer_length, debug))
        return [(k, h, t, ["ACGT"] if debug and h else []) for k, h, t in self._decide([r[0] for r in recs], abs_threshold, rel_threshold, deplete)]

    def paired_should_keep(self, recs, kmer_length, abs_threshold, rel_threshold, deplete, debug=False):
        self.calls.append(("paired", len(recs), kmer_length, debug))
        return [(k, h, t, []) for k, h, t in self._decide([r[0] for r in recs], abs_threshold, rel_threshold, deplete)]

    def lookup_batch(self, hashes, rec_off, abs_threshold=2, rel_threshold=0.01, deplete=False):
        self.calls.append(("binary", len(rec_off) - 1))
        res = self._decide([hashes[int(rec_off[i]):int(rec_off[i + 1])] for i in range(len(rec_off) - 1)], abs_threshold, rel_threshold, deplete)
        return (np.array([r[0] for r in res], np.uint8), np.array([r[1] for r in res], np.uint32), np.array([r[2] for r in res], np.uint32))


@pytest.fixture()
def served():
    def start(engine, version_path="/data/ref.idx", sha="ab" * 32):
        httpd = S.make_http_server(S.DeaconService(engine, version_path, sha), "127.0.0.1", 0)
        t = threading.Thread(target=httpd.serve_forever, daemon=True)
        t.start()
        started.append((httpd, t))
        return f"http://127.0.0.1:{httpd.server_address[1]}"
    started = []
    yield start
    for httpd, t in started:
        httpd.shutdown()
        httpd.server_close()
        t.join()


def http(url, body=None, ctype="application/json"):
    req = urllib.request.Request(url, data=body, headers={"Content-Type": ctype} if body is not None else {})
    try:
        with urllib.request.urlopen(req) as r:
            return r.status, r.headers.get("Content-Type"), r.read()
    except urllib.error.HTTPError as e:
        return e.code, e.headers.get("Content-Type"), e.read()


def unpaired_body(lists, **kw):
    p = {"abs_threshold": 2, "rel_threshold": 0.01, "deplete": False, "kmer_length": 31, "debug": False, **kw}
    return json.dumps({"input": [[h, list(range(len(h))), list(b"ACGTACGT")] for h in lists], **p}).encode()


def test_get_endpoints(served):
    url = served(ScriptedEngine())
    st, ct, body = http(url + "/")
    assert st == 200 and body == b"Index loaded with 7 minimizers and header: IndexHeader { format_version: 2, kmer_length: 31, window_size: 15 }"
    st, ct, body = http(url + "/index_header")
    assert st == 200 and ct == "application/json" and json.loads(body) == {"format_version": 2, "kmer_length": 31, "window_size": 15}
    st, ct, body = http(url + "/index_version")
    assert st == 200 and body == b"/data/ref.idx@" + b"ab" * 32
    assert http(url + "/nope")[0] == 404


def test_json_requests_follow_the_reference_schema(served):
    eng = ScriptedEngine()
    url = served(eng)
    lists = [[2, 4, 6, 7], [1, 3], [], [2, 2, 2, 8]]
    st, ct, body = http(url + "/should_output_unpaired", unpaired_body(lists))
    assert st == 200 and ct == "application/json"
    assert json.loads(body) == {"should_output": [[True, 3, 4, []], [False, 0, 2, []], [False, 0, 0, []], [True, 2, 4, []]]}
    st, ct, body = http(url + "/should_output_unpaired", unpaired_body(lists, deplete=True, debug=True))
    assert json.loads(body)["should_output"][0] == [False, 3, 4, ["ACGT"]] and json.loads(body)["should_output"][2] == [True, 0, 0, []]
    # u64 hashes above 2^63 survive the JSON round trip (serde_json writes them as plain integers)
    big = [[2 ** 64 - 2, 2 ** 63 + 2]]
    assert json.loads(http(url + "/should_output_unpaired", unpaired_body(big))[2])["should_output"] == [[True, 2, 2, []]]
    paired = json.dumps({"input": [[[2, 4, 5], [0, 1, 2], []]], "abs_threshold": 1, "rel_threshold": 0.0, "deplete": True, "kmer_length": 31,
                         "debug": False}).encode()
    assert json.loads(http(url + "/should_output_paired", paired)[2]) == {"should_output": [[False, 2, 3, []]]}
    assert [c[0] for c in eng.calls] == ["unpaired", "unpaired", "unpaired", "paired"] and eng.calls[1][3] is True
    # bodies that do not deserialise are refused, like axum's Json extractor (422), and the engine is not called
    n = len(eng.calls)
    assert http(url + "/should_output_unpaired", b"{\"input\": [[1, 2]]}")[0] == 422
    assert http(url + "/should_output_unpaired", b"not json")[0] == 422
    assert http(url + "/should_output_unpaired", unpaired_body(lists, kmer_length=300))[0] == 422
    assert http(url + "/frob", b"{}")[0] == 404 and len(eng.calls) == n


def test_binary_body_round_trip(served):
    eng = ScriptedEngine()
    url = served(eng)
    hashes = np.array([2, 4, 6, 7, 1, 3, 2, 2, 2, 8], np.uint64)
    off = np.array([0, 4, 6, 6, 10], np.uint64)
    st, ct, body = http(url + "/should_output_unpaired", S.pack_binary_request(hashes, off), S.BINARY_TYPE)
    assert st == 200 and ct == S.BINARY_TYPE
    keep, hits, total = S.parse_binary_response(body)
    assert keep.tolist() == [1, 0, 0, 1] and hits.tolist() == [3, 0, 0, 2] and total.tolist() == [4, 2, 0, 4]
    assert eng.calls == [("binary", 4)]
    bad = S.pack_binary_request(hashes, off)
    assert http(url + "/should_output_unpaired", bad[:-8], S.BINARY_TYPE)[0] == 422
    assert http(url + "/should_output_unpaired", b"XXXX" + bad[4:], S.BINARY_TYPE)[0] == 422
    assert http(url + "/should_output_unpaired", S.pack_binary_request(hashes, np.array([0, 6, 4, 10], np.uint64)), S.BINARY_TYPE)[0] == 422


@pytest.mark.gpu
def test_real_engine_behind_the_endpoints(served, tmp_path):
    """The server's answers are the oracle's lookups: JSON and binary forms, single and pooled-pair records, --debug k-mers."""
    import helpers as H
    from oracle import oracle as O
    import deacon_server_b200 as d

    g = H.random_genome(120_000, 3)
    idx = O.index_build([g], 31, 15, threads=8)
    path = tmp_path / "ref.idx"
    path.write_bytes(O.idx_encode(idx.keys(), 31, 15))
    data = path.read_bytes()
    gpu = d.DeaconGpu(0)
    try:
        gpu.idx_decode(data, 0, make_resident=True)
        url = served(gpu, str(path), hashlib.sha256(data).hexdigest())
        assert http(url + "/")[2].decode().startswith(f"Index loaded with {len(idx)} minimizers")
        assert http(url + "/index_version")[2].decode() == f"{path}@{hashlib.sha256(data).hexdigest()}"
        reads = H.sample_reads(g, 300, (31, 400), 9)
        recs, lists = [], []
        for r in reads:   # the client's side of the protocol: extraction (src/remote_filter.rs:762-790)
            h, p = O.extract_filter(r, 31, 15)
            recs.append([[int(x) for x in h], [int(x) for x in p], r.tolist()])
            lists.append(np.asarray(h, np.uint64))
        off = np.zeros(len(lists) + 1, np.uint64)
        off[1:] = np.cumsum([len(x) for x in lists])
        ok, oh, ot = O.lookup_batch(idx, np.concatenate(lists), off, threads=4)
        body = json.dumps({"input": recs, "abs_threshold": 2, "rel_threshold": 0.01, "deplete": False, "kmer_length": 31, "debug": True}).encode()
        out = json.loads(http(url + "/should_output_unpaired", body)[2])["should_output"]
        assert [(bool(a), b, c) for a, b, c, _ in out] == [(bool(a), int(b), int(c)) for a, b, c in zip(ok, oh, ot)]
        assert all(len(km) == h and all(len(s) == 31 for s in km) for (_, h, _, km) in out)   # one k-mer string per counted hit
        st, ct, bin_out = http(url + "/should_output_unpaired", S.pack_binary_request(np.concatenate(lists), off), S.BINARY_TYPE)
        keep, hits, total = S.parse_binary_response(bin_out)
        assert np.array_equal(keep, ok) and np.array_equal(hits, oh) and np.array_equal(total, ot)
        # pairs: one pooled hash list per pair, sequences always empty
        pl = [np.concatenate([lists[i], lists[i + 1]]) for i in range(0, len(lists) - 1, 2)]
        poff = np.zeros(len(pl) + 1, np.uint64)
        poff[1:] = np.cumsum([len(x) for x in pl])
        pk, ph, pt = O.lookup_batch(idx, np.concatenate(pl), poff, deplete=True, threads=4)
        body = json.dumps({"input": [[[int(x) for x in h], [0] * len(h), []] for h in pl], "abs_threshold": 2, "rel_threshold": 0.01,
                           "deplete": True, "kmer_length": 31, "debug": False}).encode()
        out = json.loads(http(url + "/should_output_paired", body)[2])["should_output"]
        assert [(bool(a), b, c, e) for a, b, c, e in out] == [(bool(a), int(b), int(c), []) for a, b, c in zip(pk, ph, pt)]
    finally:
        gpu.close()
