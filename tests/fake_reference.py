"""Writes files in the layout of tools/make_reference_fixtures.sh from the ORACLE's own output, so that
tests/test_reference_fixtures.py (parsers, id conventions, pair rules, summary fields) can be exercised on a box
without the reference binary.  The result proves nothing about parity: it is the oracle compared with itself."""
import gzip
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402


def main(out):
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "gen_fixture_inputs.py"), out])
    os.environ["DCN_REFERENCE_FIXTURES"] = out
    import test_reference_fixtures as T
    T.FIX = out
    genome = [np.frombuffer(s, np.uint8) for _id, s in T.read_fastx("genome.fa")]
    rng = np.random.default_rng(1)
    for name, k, w, e in T.INDEXES:
        keys = O.index_build(genome, k, w, e).keys()
        keys = keys[rng.permutation(len(keys))]          # a hash set's iteration order is arbitrary
        with open(os.path.join(out, name + ".idx"), "wb") as f:
            f.write(O.idx_encode(keys, k, w))
    for run in T.RUNS:
        stem, idxname, _reads, paired, prefix, abs_thr, rel, deplete = run
        _v, k, w, keys = T.load_idx(idxname)
        units, _ = T.expected_units(run, keys, k, w)
        lines = []
        for uid, hits, total, keep, kmers, _bp in units:
            if paired and hits == 0:
                continue
            lines.append(f"DEBUG: {uid} hits={hits}/{total} keep={'true' if keep else 'false'} kmers=[{','.join(kmers)}]")
        with gzip.open(os.path.join(out, stem + "_debug.txt.gz"), "wb") as f:
            f.write(("\n".join(lines) + "\n").encode())
        per = 2 if paired else 1
        s = {"seqs_in": per * len(units), "seqs_out": per * sum(1 for u in units if u[3]),
             "bp_in": sum(u[5] for u in units), "bp_out": sum(u[5] for u in units if u[3])}
        with open(os.path.join(out, stem + "_summary.json"), "w") as f:
            json.dump(s, f)
    for fn in os.listdir(out):
        if fn.endswith((".fa", ".fq")):
            subprocess.check_call(["gzip", "-9nf", os.path.join(out, fn)])


if __name__ == "__main__":
    main(sys.argv[1])
