"""Wire-compatible front end of the batch engine: the reference's `deacon server` endpoints
(src/server.rs:48-58) over the GPU-resident index, plus a binary body for callers that do not want JSON.

  GET  /                        text  "Index loaded with <n> minimizers and header: IndexHeader { ... }"   (src/server.rs:90-104)
  GET  /index_header            json  {"format_version":2,"kmer_length":31,"window_size":15}               (src/server.rs:108-111)
  GET  /index_version           text  "<index path>@<sha256 of the file>"                                  (src/server.rs:68-73,115-118)
  POST /should_output_unpaired  json  UnpairedFilterRequest -> FilterResponse                              (src/server_common.rs:9-58)
  POST /should_output_paired    json  PairedFilterRequest   -> FilterResponse

Every decision comes from `engine.unpaired_should_keep / paired_should_keep` (deacon_server_b200.api.DeaconGpu: one
dcn_lookup_batch[_flags] call per request); requests are serialised by a lock, like the reference's INDEX mutex
(src/server.rs:121,145).  This is host plumbing: Python's http.server, no framework.

Binary body (either POST endpoint, `Content-Type: application/x-deacon-batch`), little-endian:
  request   "DCNB" | u32 n_rec | u32 abs_threshold | f64 rel_threshold | u8 deplete | u8 kmer_length | u16 0
            | u64 rec_off[n_rec + 1] | u64 hashes[rec_off[n_rec]]
  response  "DCNR" | u32 n_rec | u8 keep[n_rec] | u32 hits[n_rec] | u32 total[n_rec]
(no debug k-mers: those need the sequences, i.e. the JSON form).
"""
from __future__ import annotations

import hashlib
import json
import struct
import threading
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

import numpy as np

BODY_LIMIT = 2147483648   # DefaultBodyLimit::max, src/server.rs:58
BINARY_TYPE = "application/x-deacon-batch"
_REQ_HEAD = struct.Struct("<4sIIdBBH")


class BadRequest(ValueError):
    pass


def parse_json_request(body: bytes, paired: bool):
    """UnpairedFilterRequest / PairedFilterRequest (src/server_common.rs:9-50) -> (records, params).
    A record is (hashes, positions, sequence bytes | list of sequences): what unpaired_/paired_should_keep take."""
    try:
        req = json.loads(body)
        params = {"abs_threshold": int(req["abs_threshold"]), "rel_threshold": float(req["rel_threshold"]),
                  "deplete": bool(req["deplete"]), "kmer_length": int(req["kmer_length"]), "debug": bool(req["debug"])}
        records = []
        for rec in req["input"]:
            hashes, positions, seqs = rec
            h = np.asarray(hashes, dtype=np.uint64) if len(hashes) else np.zeros(0, np.uint64)
            p = np.asarray(positions, dtype=np.uint32) if len(positions) else np.zeros(0, np.uint32)
            s = [bytes(x) for x in seqs] if paired else bytes(seqs)
            records.append((h, p, s))
    except (KeyError, TypeError, ValueError, OverflowError) as e:   # axum answers 422 to a body that does not deserialise
        raise BadRequest(f"Failed to deserialize the JSON body into the target type: {e}") from e
    if params["abs_threshold"] < 0 or not 0 <= params["kmer_length"] <= 255:
        raise BadRequest("Failed to deserialize the JSON body into the target type: value out of range")
    return records, params


def format_json_response(results) -> bytes:
    """FilterResponse (src/server_common.rs:54-58): should_output = [(bool, hits, total, [kmers])]."""
    return json.dumps({"should_output": [[bool(k), int(h), int(t), list(km)] for k, h, t, km in results]},
                      separators=(",", ":")).encode()


def pack_binary_request(hashes, rec_off, abs_threshold=2, rel_threshold=0.01, deplete=False, kmer_length=31) -> bytes:
    hashes = np.ascontiguousarray(hashes, np.uint64)
    rec_off = np.ascontiguousarray(rec_off, np.uint64)
    return _REQ_HEAD.pack(b"DCNB", len(rec_off) - 1, abs_threshold, rel_threshold, int(deplete), kmer_length, 0) + \
        rec_off.tobytes() + hashes.tobytes()


def parse_binary_request(body: bytes):
    if len(body) < _REQ_HEAD.size:
        raise BadRequest("binary request shorter than its header")
    magic, n_rec, abs_thr, rel_thr, deplete, k, _ = _REQ_HEAD.unpack_from(body)
    if magic != b"DCNB":
        raise BadRequest("binary request does not start with DCNB")
    at = _REQ_HEAD.size
    if len(body) < at + 8 * (n_rec + 1):
        raise BadRequest("binary request truncated in rec_off")
    rec_off = np.frombuffer(body, np.uint64, n_rec + 1, at)
    at += 8 * (n_rec + 1)
    if rec_off[0] != 0 or np.any(rec_off[1:] < rec_off[:-1]):
        raise BadRequest("rec_off must start at 0 and be non-decreasing")
    n_h = int(rec_off[-1])
    if len(body) != at + 8 * n_h:
        raise BadRequest("binary request length does not match rec_off")
    hashes = np.frombuffer(body, np.uint64, n_h, at)
    return hashes, rec_off, {"abs_threshold": abs_thr, "rel_threshold": rel_thr, "deplete": bool(deplete), "kmer_length": k}


def pack_binary_response(keep, hits, total) -> bytes:
    keep = np.ascontiguousarray(keep, np.uint8)
    return b"DCNR" + struct.pack("<I", len(keep)) + keep.tobytes() + np.ascontiguousarray(hits, np.uint32).tobytes() + \
        np.ascontiguousarray(total, np.uint32).tobytes()


def parse_binary_response(body: bytes):
    if body[:4] != b"DCNR":
        raise ValueError("not a DCNR response")
    n, = struct.unpack_from("<I", body, 4)
    keep = np.frombuffer(body, np.uint8, n, 8)
    hits = np.frombuffer(body, np.uint32, n, 8 + n)
    total = np.frombuffer(body, np.uint32, n, 8 + 5 * n)
    return keep, hits, total


class DeaconService:
    """The endpoint logic, independent of the HTTP layer.  `engine` provides index_info(), header,
    unpaired_should_keep, paired_should_keep and lookup_batch (DeaconGpu does)."""

    def __init__(self, engine, index_path: str, index_sha256: str):
        self.engine = engine
        self.index_version = f"{index_path}@{index_sha256}"
        self.lock = threading.Lock()   # one request at a time on the ctx (src/server.rs:121,145)

    def get(self, path: str):
        if path == "/":
            with self.lock:
                info = self.engine.index_info()
            text = (f"Index loaded with {info['n_keys']} minimizers and header: IndexHeader {{ format_version: 2, "
                    f"kmer_length: {info['kmer_length']}, window_size: {info['window_size']} }}")
            return 200, "text/plain; charset=utf-8", text.encode()
        if path == "/index_header":
            with self.lock:
                info = self.engine.index_info()
            body = json.dumps({"format_version": 2, "kmer_length": info["kmer_length"], "window_size": info["window_size"]},
                              separators=(",", ":")).encode()
            return 200, "application/json", body
        if path == "/index_version":
            return 200, "text/plain; charset=utf-8", self.index_version.encode()
        return 404, "text/plain; charset=utf-8", b""

    def post(self, path: str, content_type: str, body: bytes):
        if path not in ("/should_output_unpaired", "/should_output_paired"):
            return 404, "text/plain; charset=utf-8", b""
        paired = path.endswith("_paired")
        try:
            if content_type.split(";")[0].strip() == BINARY_TYPE:
                hashes, rec_off, p = parse_binary_request(body)
                with self.lock:
                    keep, hits, total = self.engine.lookup_batch(hashes, rec_off, p["abs_threshold"], p["rel_threshold"], p["deplete"])
                return 200, BINARY_TYPE, pack_binary_response(keep, hits, total)
            records, p = parse_json_request(body, paired)
            fn = self.engine.paired_should_keep if paired else self.engine.unpaired_should_keep
            with self.lock:
                results = fn(records, p["kmer_length"], p["abs_threshold"], p["rel_threshold"], p["deplete"], p["debug"])
            return 200, "application/json", format_json_response(results)
        except BadRequest as e:
            return 422, "text/plain; charset=utf-8", str(e).encode()


def make_http_server(service: DeaconService, host: str = "0.0.0.0", port: int = 8888) -> ThreadingHTTPServer:
    class Handler(BaseHTTPRequestHandler):
        protocol_version = "HTTP/1.1"

        def _send(self, status, ctype, body):
            self.send_response(status)
            self.send_header("Content-Type", ctype)
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def do_GET(self):
            self._send(*service.get(self.path))

        def do_POST(self):
            n = int(self.headers.get("Content-Length") or 0)
            if n > BODY_LIMIT:
                self._send(413, "text/plain; charset=utf-8", b"length limit exceeded")
                return
            body = self.rfile.read(n)
            self._send(*service.post(self.path, self.headers.get("Content-Type") or "", body))

        def log_message(self, *args):   # quiet unless asked (the reference logs with RUST_LOG=trace)
            pass

    return ThreadingHTTPServer((host, port), Handler)


def run_server(index_path: str, port: int = 8888, device: int = 0, host: str = "0.0.0.0"):
    """`deacon server IDX -p PORT` (src/server.rs:38-64): load the index into HBM, then serve."""
    import sys
    from .api import DeaconGpu
    print(f"Loading index from: {index_path}", file=sys.stderr)
    data = open(index_path, "rb").read()
    gpu = DeaconGpu(device)
    gpu.idx_decode(data, 0, make_resident=True)
    print("Loaded index!", file=sys.stderr)
    httpd = make_http_server(DeaconService(gpu, str(index_path), hashlib.sha256(data).hexdigest()), host, port)
    try:
        httpd.serve_forever()
    finally:
        httpd.server_close()
        gpu.close()


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="deacon server (B200): hold an index in HBM and answer the reference's client")
    ap.add_argument("index")
    ap.add_argument("-p", "--port", type=int, default=8888)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    run_server(a.index, a.port, a.device)
