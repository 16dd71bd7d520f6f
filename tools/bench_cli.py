"""File-to-file throughput of the C++ host driver (deacon-b200): synthetic FASTA reference -> `index build`,
synthetic paired FASTQ (page cache) -> `filter --deplete`, output to a file in /dev/shm.  Prints one JSON line.
Not the contract bench (bench.py); the numbers are quoted in DESIGN.md."""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "deacon_server_b200", "deacon-b200")

ap = argparse.ArgumentParser()
ap.add_argument("--genome-mbp", type=float, default=200)
ap.add_argument("--pairs-m", type=float, default=5)
ap.add_argument("--dir", default="/dev/shm/dcn_cli")
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--devices", default="0")
ap.add_argument("--batch-mbp", type=int, default=64)
args = ap.parse_args()
os.makedirs(args.dir, exist_ok=True)
rng = np.random.default_rng(1)
ACGT = np.frombuffer(b"ACGT", np.uint8)
G = int(args.genome_mbp * 1e6)
genome = ACGT[rng.integers(0, 4, G, dtype=np.uint8)]
ref = os.path.join(args.dir, "ref.fa")
with open(ref, "wb") as f:
    f.write(b">chr1\n")
    lines = np.full((G // 80, 81), 10, np.uint8)
    lines[:, :80] = genome[: G // 80 * 80].reshape(-1, 80)
    f.write(lines.tobytes())


def fastq_file(path, reads, tag):
    n, ln = reads.shape
    w = 11 + ln + 3 + ln + 1
    rec = np.empty((n, w), np.uint8)
    rec[:, 0] = ord("@"); rec[:, 1] = tag
    ids = np.arange(n, dtype=np.int64)
    for d in range(8):
        rec[:, 9 - d] = 48 + (ids // 10 ** d) % 10
    rec[:, 10] = 10
    rec[:, 11:11 + ln] = reads
    rec[:, 11 + ln:14 + ln] = np.frombuffer(b"\n+\n", np.uint8)
    rec[:, 14 + ln:14 + 2 * ln] = ord("I")
    rec[:, -1] = 10
    with open(path, "wb") as f:
        f.write(rec.tobytes())


NP = int(args.pairs_m * 1e6)
pos = rng.integers(0, G - 600, NP)
ar = np.arange(150)
comp = np.zeros(256, np.uint8)
for a, b in zip(b"ACGT", b"TGCA"):
    comp[a] = b
m1 = genome[pos[:, None] + ar[None, :]]
m2 = comp[genome[(pos + 400)[:, None] - 1 - ar[None, :]]]
rnd = rng.random(NP) < 0.1
m1[rnd] = ACGT[rng.integers(0, 4, (int(rnd.sum()), 150), dtype=np.uint8)]
m2[rnd] = ACGT[rng.integers(0, 4, (int(rnd.sum()), 150), dtype=np.uint8)]
r1, r2 = os.path.join(args.dir, "r1.fq"), os.path.join(args.dir, "r2.fq")
fastq_file(r1, m1, ord("a"))
fastq_file(r2, m2, ord("b"))
del m1, m2, genome


def timed(*cmd):
    t0 = time.perf_counter()
    p = subprocess.run([BIN, *map(str, cmd)], capture_output=True)
    dt = time.perf_counter() - t0
    if p.returncode:
        sys.stderr.write(p.stderr.decode())
        sys.exit(1)
    return dt, p.stderr.decode()


idx = os.path.join(args.dir, "ref.idx")
t_build, err = timed("index", "build", "-q", "-o", idx, ref)
out = {"what": "deacon-b200 file to file", "genome_mbp": args.genome_mbp, "index_build_s": round(t_build, 2), "cpus": os.cpu_count(),
       "index_build_log": [ln for ln in err.splitlines() if ln.startswith(("Indexed", "Completed"))]}
extra = ["-t", args.threads, "--devices", args.devices, "--batch-mbp", args.batch_mbp]
summ = os.path.join(args.dir, "s.json")
for label, cmd in (("paired_deplete", ["filter", "-d", idx, r1, r2, "-o", os.path.join(args.dir, "o1.fq"), "-O", os.path.join(args.dir, "o2.fq")]),
                   ("single_search", ["filter", idx, r1, "-o", os.path.join(args.dir, "o.fq")])):
    timed(*cmd, "-s", summ, *extra)   # warm (page cache, driver)
    dt, err = timed(*cmd, "-s", summ, *extra)
    s = json.load(open(summ))
    load = [ln for ln in err.splitlines() if ln.startswith(("Loaded index", "host stages"))]
    phase = [ln for ln in load if "filter phase" in ln]
    out[label] = {"filter_phase_gbp_per_s": float(phase[0].split("(")[1].split(" ")[0]) if phase else None, "wall_s": round(dt, 2), "bp_in": s["bp_in"], "seqs_out": s["seqs_out"], "summary_time_s": round(s["time"], 2),
                  "gbp_per_s_incl_startup": round(s["bp_in"] / s["time"] / 1e9, 2), "loaded_index": load}
print(json.dumps(out))
