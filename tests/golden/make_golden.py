"""Regenerates tests/golden/extract_vectors.json and xxh3_vectors.json.

extract_vectors: seeded random sequences -> (hashes, positions) computed by the INDEPENDENT
pure-Python restatement (oracle/py_oracle.py).  The reference itself (Rust, un-vendored crates)
cannot run in this container, so these pin the C oracle and the CUDA path against a second,
separately written restatement of SURVEY.md Appendix A -- not against a reference binary.
xxh3_vectors: from the `xxhash` Python module (the public XXH3 implementation).

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import py_oracle as P  # noqa: E402


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def main():
    rng = random.Random(20261018)
    cases = []
    for (k, w) in [(31, 15)] * 14 + [(5, 3), (5, 5), (31, 1), (21, 11), (15, 15), (41, 15), (57, 5), (33, 13)]:
        n = rng.choice([0, k - 1, k, k + w - 2, k + w - 1, k + w, 100, 150, 151, 300, 700])
        kind = rng.choice(["plain", "plain", "n", "lower", "lowcomplex", "newline", "iupac"])
        if kind == "lowcomplex":
            unit = rand_seq(rng, rng.choice([1, 2, 3, 5]))
            seq = (unit * (n // len(unit) + 1))[:n]
        else:
            seq = rand_seq(rng, n)
        if kind == "n" and n:
            s = list(seq)
            for _ in range(rng.randint(1, 3)):
                s[rng.randrange(n)] = "N"
            seq = "".join(s)
        if kind == "iupac" and n:
            s = list(seq)
            for _ in range(rng.randint(1, 6)):
                s[rng.randrange(n)] = rng.choice("RYSWKMBDHVNryswkmbdhvn-*")
            seq = "".join(s)
        if kind == "lower":
            seq = "".join(c.lower() if rng.random() < 0.3 else c for c in seq)
        if kind == "newline" and n:
            seq = seq[:-1] + "\n"
        prefix = rng.choice([0, 0, 0, 50, 80])
        hs, ps = P.extract_filter(seq.encode(), k, w, prefix)
        hi = P.extract_index(seq.encode(), k, w)
        cases.append({"seq": seq, "k": k, "w": w, "prefix": prefix, "kind": kind,
                      "filter_hashes": [hex(h) for h in hs], "filter_positions": ps,
                      "index_hashes": [hex(h) for h in hi]})
    with open(os.path.join(HERE, "extract_vectors.json"), "w") as f:
        json.dump({"_comment": __doc__.strip().split("\n\n")[0], "cases": cases}, f, indent=0)

    import xxhash
    vec = []
    for _ in range(64):
        v = rng.getrandbits(64)
        vec.append({"bytes": 8, "value": hex(v), "hash": hex(xxhash.xxh3_64_intdigest(v.to_bytes(8, "little")))})
        v = rng.getrandbits(114)
        vec.append({"bytes": 16, "value": hex(v), "hash": hex(xxhash.xxh3_64_intdigest(v.to_bytes(16, "little")))})
    with open(os.path.join(HERE, "xxh3_vectors.json"), "w") as f:
        json.dump({"_comment": "xxhash.xxh3_64_intdigest(value.to_bytes(n, 'little')), seed 0", "vectors": vec}, f, indent=0)
    print("wrote", len(cases), "extract cases and", len(vec), "xxh3 vectors")


if __name__ == "__main__":
    main()
