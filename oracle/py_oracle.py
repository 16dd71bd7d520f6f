"""Independent pure-Python restatement of SURVEY.md Appendix A (small inputs only).

TEST INFRASTRUCTURE ONLY.  Written from the closed forms, sharing no code with
deacon_oracle.c, so that a typo in one is caught by the other.  "parity unpinned" for
minimizer selection (see deacon_oracle.h).
"""
from __future__ import annotations

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF

# simd-minimizers ntHash seeds indexed by packed-seq code A=0 C=1 T=2 G=3
F = (0x95C60474, 0x62A02B4C, 0x82572324, 0x4BE24456)
ACGT = frozenset(b"ACGTacgt")

# src/minimizers.rs:24-43
_IUPAC = {}
for _s, _d in (("Aa", "A"), ("Cc", "C"), ("Gg", "G"), ("Tt", "T"), ("Rr", "G"), ("Yy", "C"),
               ("Ss", "G"), ("Ww", "A"), ("Kk", "G"), ("Mm", "C"), ("Bb", "C"), ("Dd", "G"),
               ("Hh", "C"), ("Vv", "G"), ("Nn", "C")):
    for _c in _s:
        _IUPAC[ord(_c)] = ord(_d)


def rotl32(x, r):
    r %= 32
    return ((x << r) | (x >> (32 - r))) & M32 if r else x


def rotl64(x, r):
    return ((x << r) | (x >> (64 - r))) & M64


def xxh3_u64(v):
    """XXH3_64bits(le64(v)), seed 0 (src/filter_common.rs:305)."""
    x = ((v >> 32) | (v << 32)) & M64
    x ^= 0xC73AB174C5ECD5A2
    x ^= rotl64(x, 49) ^ rotl64(x, 24)
    x = (x * 0x9FB21C651E98DF25) & M64
    x ^= ((x >> 35) + 8) & M64
    x = (x * 0x9FB21C651E98DF25) & M64
    return x ^ (x >> 28)


def xxh3_u128(v):
    """XXH3_64bits(le128(v)), seed 0 (src/filter_common.rs:296)."""
    lo = (v & M64) ^ 0x6782737BEA4239B9
    hi = ((v >> 64) & M64) ^ 0xAF56BC3B0996523A
    m = lo * hi
    sw = int.from_bytes(lo.to_bytes(8, "little"), "big")
    acc = (16 + sw + hi + ((m & M64) ^ (m >> 64))) & M64
    acc ^= acc >> 37
    acc = (acc * 0x165667919E3779F9) & M64
    return acc ^ (acc >> 32)


def codes_of(seq: bytes):
    return [(b >> 1) & 3 for b in seq]


def nthash(codes, p, k):
    fw = rc = 0
    for i in range(k):
        fw ^= rotl32(F[codes[p + i]], k - 1 - i)
        rc ^= rotl32(F[codes[p + i] ^ 2], i)
    return (fw + rc) & M32


def minimizer_positions(codes, k, w):
    """SURVEY A.3."""
    n = len(codes)
    l = k + w - 1
    if n < l:
        return []
    keys = [nthash(codes, p, k) >> 16 for p in range(n - k + 1)]
    out = []
    prev = None
    for j in range(n - l + 1):
        win = keys[j:j + w]
        m = min(win)
        left = j + win.index(m)
        right = j + (w - 1 - win[::-1].index(m))
        tg = sum((c >> 1) & 1 for c in codes[j:j + l])
        pick = left if 2 * tg > l else right
        if prev is None or pick != prev:
            out.append(pick)
        prev = pick
    return out


def canonical_value(codes, p, k):
    fw = rc = 0
    for i in range(k):
        fw |= codes[p + i] << (2 * i)
        rc |= (codes[p + k - 1 - i] ^ 2) << (2 * i)
    return min(fw, rc)


def kmer_hash(codes, p, k):
    v = canonical_value(codes, p, k)
    return xxh3_u64(v) if k <= 32 else xxh3_u128(v)


def extract_filter(seq: bytes, k=31, w=15, prefix_len=0):
    """src/filter_common.rs:211-310 -> (hashes, positions)."""
    if len(seq) < k:
        return [], []
    eff = seq[:prefix_len] if (prefix_len > 0 and len(seq) > prefix_len) else seq
    if eff.endswith(b"\n"):
        eff = eff[:-1]
    codes = codes_of(eff)
    hs, ps = [], []
    for p in minimizer_positions(codes, k, w):
        if all(b in ACGT for b in eff[p:p + k]):
            ps.append(p)
            hs.append(kmer_hash(codes, p, k))
    return hs, ps


def extract_index(seq: bytes, k=31, w=15):
    """src/minimizers.rs:125-191 with entropy_threshold == 0."""
    if len(seq) < k:
        return []
    mapped = bytes(_IUPAC.get(b, ord("C")) for b in seq)
    codes = codes_of(mapped)
    return [kmer_hash(codes, p, k) for p in minimizer_positions(codes, k, w)
            if all(b in ACGT for b in seq[p:p + k])]


def required_hits(abs_thr, rel_thr, total):
    """src/filter_common.rs:84-96; f64 round half away from zero."""
    import math
    if total == 0:
        rel = 0
    else:
        x = rel_thr * float(total)
        if x != x or x <= 0:
            r = 0
        else:
            r = math.floor(x)
            if x - r >= 0.5:  # exact in binary floating point
                r += 1
        rel = max(1, int(r))
    return max(abs_thr, rel)


def should_keep(index: set, hashes, abs_thr=2, rel_thr=0.01, deplete=False):
    """src/filter_common.rs:99-155 -> (keep, hits, total)."""
    hits = len({h for h in hashes if h in index})
    req = required_hits(abs_thr, rel_thr, len(hashes))
    return ((hits < req) if deplete else (hits >= req)), hits, len(hashes)
