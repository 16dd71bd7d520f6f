"""The C++ host driver (deacon_server_b200/host): CPU-side checks that need no GPU.

FASTA/FASTQ block reader (threads x block sizes, so that records straddle blocks and thread ranges), codec layers,
command-line surface (the reference's tests/cli_tests.rs), summary JSON helpers.  The GPU tests of the driver are in
test_host_cli_gpu.py.
"""
import gzip
import lzma
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "deacon_server_b200", "deacon-b200")


def run(*args, stdin=None, check=True):
    p = subprocess.run([BIN, *map(str, args)], input=stdin, capture_output=True)
    if check:
        assert p.returncode == 0, p.stderr.decode()
    return p


def parse_dump(out: bytes):
    recs = []
    for line in out.split(b"\n")[:-1]:
        i, s, q, v = line.split(b"\t")
        recs.append((i, s, q, v == b"1"))
    return recs


def rand_seq(rng, n):
    return bytes(rng.choice(b"ACGTN") for _ in range(n))


def make_fastq(rng, n, crlf=False, plus_id=False, final_newline=True):
    eol = b"\r\n" if crlf else b"\n"
    recs, blob = [], b""
    for i in range(n):
        ident = b"read%d some description/%d" % (i, i % 2 + 1)
        seq = rand_seq(rng, rng.choice([0, 1, 30, 31, 150, 151, 700]))
        # qualities that start with '@' or '+' are the classic traps for a block splitter
        qual = bytes(rng.choice(b"@+IJ#~") for _ in seq)
        plus = b"+" + (ident if plus_id else b"")
        blob += b"@" + ident + eol + seq + eol + plus + eol + qual + eol
        recs.append((ident, seq, qual))
    if not final_newline:
        blob = blob[: -len(eol)]
    return recs, blob


def make_fasta(rng, n, width, crlf=False):
    eol = b"\r\n" if crlf else b"\n"
    recs, blob = [], b""
    for i in range(n):
        ident = b"contig_%d len" % i
        seq = rand_seq(rng, rng.choice([0, 5, 59, 60, 61, 1000, 5000]))
        blob += b">" + ident + eol
        if width:
            for j in range(0, len(seq), width):
                blob += seq[j:j + width] + eol
        else:
            blob += seq + eol
        recs.append((ident, seq, b""))
    return recs, blob


@pytest.mark.parametrize("threads,block", [(1, 1 << 25), (4, 1 << 25), (3, 4096), (8, 700), (2, 64)])
@pytest.mark.parametrize("kind", ["fastq", "fastq_crlf_plusid", "fastq_no_final_newline", "fasta_multiline", "fasta_single", "fasta_crlf"])
def test_reader_matches_line_parser(tmp_path, threads, block, kind):
    rng = random.Random(hash((threads, block, kind)) & 0xFFFF)
    if kind == "fastq":
        recs, blob = make_fastq(rng, 400)
    elif kind == "fastq_crlf_plusid":
        recs, blob = make_fastq(rng, 200, crlf=True, plus_id=True)
    elif kind == "fastq_no_final_newline":
        recs, blob = make_fastq(rng, 50, final_newline=False)
    elif kind == "fasta_multiline":
        recs, blob = make_fasta(rng, 60, 60)
        blob += b"\n\n"   # trailing blank lines
    elif kind == "fasta_single":
        recs, blob = make_fasta(rng, 100, 0)
    else:
        recs, blob = make_fasta(rng, 40, 70, crlf=True)
    path = tmp_path / "in.fx"
    path.write_bytes(blob)
    got = parse_dump(run("_parse", path, threads, block).stdout)
    assert [(i, s, q) for i, s, q, _ in got] == recs
    if kind in ("fastq", "fasta_single"):
        # records already in the output layout are copied verbatim by the writer (unless they end the file without a line break)
        assert all(v for *_, v in got)
    if kind in ("fastq_crlf_plusid", "fasta_crlf"):
        assert not any(v for *_, v in got)
    if kind == "fasta_multiline":   # only sequences that fit one line are already in the output layout
        assert all(v == (0 < len(s) <= 60) for _, s, _, v in got[:-1])   # (the last record is followed by blank lines)


def test_large_records_grow_the_block(tmp_path):
    rng = random.Random(5)
    seq = rand_seq(rng, 300_000)
    blob = b">big one\n" + b"\n".join(seq[i:i + 80] for i in range(0, len(seq), 80)) + b"\n>small\nACGT\n"
    path = tmp_path / "big.fa"
    path.write_bytes(blob)
    got = parse_dump(run("_parse", path, 4, 1024).stdout)
    assert [(i, s) for i, s, _, _ in got] == [(b"big one", seq), (b"small", b"ACGT")]


@pytest.mark.parametrize("codec", ["gz", "bgzf_like", "xz", "zst"])
def test_compressed_inputs_and_outputs(tmp_path, codec):
    rng = random.Random(11)
    recs, blob = make_fastq(rng, 300)
    src = tmp_path / ("in.fq." + ("gz" if codec == "bgzf_like" else codec))
    if codec == "gz":
        src.write_bytes(gzip.compress(blob))
    elif codec == "bgzf_like":   # several gzip members back to back
        src.write_bytes(b"".join(gzip.compress(blob[i:i + 5000]) for i in range(0, len(blob), 5000)))
    elif codec == "xz":
        src.write_bytes(lzma.compress(blob))
    else:   # no zstd module in this image: make the file with the driver's own encoder and check the round trip
        plain = tmp_path / "plain.fq"
        plain.write_bytes(blob)
        run("_recode", plain, src, 3)
        assert src.read_bytes()[:4] == b"\x28\xb5\x2f\xfd"
    got = parse_dump(run("_parse", src, 2, 8192).stdout)
    assert [(i, s, q) for i, s, q, _ in got] == recs
    # the sinks: recode to each output format and read back with Python's own decoders where it has them
    out_gz, out_xz, out_plain = tmp_path / "o.gz", tmp_path / "o.xz", tmp_path / "o.txt"
    run("_recode", src, out_gz, 6)
    assert gzip.decompress(out_gz.read_bytes()) == blob
    run("_recode", src, out_xz, 1)
    assert lzma.decompress(out_xz.read_bytes()) == blob
    run("_recode", src, out_plain)
    assert out_plain.read_bytes() == blob


def test_invalid_inputs_fail_loudly(tmp_path):
    bad = tmp_path / "bad.fq"
    bad.write_bytes(b"@r1\nACGT\n+\nIII\n")   # quality shorter than the sequence
    p = run("_parse", bad, check=False)
    assert p.returncode == 1 and b"lengths differ" in p.stderr
    trunc = tmp_path / "trunc.fq"
    trunc.write_bytes(b"@r1\nACGT\n+\nIIII\n@r2\nAC")
    p = run("_parse", trunc, check=False)
    assert p.returncode == 1 and b"Truncated" in p.stderr
    other = tmp_path / "other.txt"
    other.write_bytes(b"hello\n")
    p = run("_parse", other, check=False)
    assert p.returncode == 1 and b"neither FASTA nor FASTQ" in p.stderr
    p = run("_parse", tmp_path / "missing.fq", check=False)
    assert p.returncode == 1 and b"Failed to open" in p.stderr
    p = run("_recode", bad, tmp_path / "x.gz", 12, check=False)   # validate_compression_level, src/local_filter.rs:94-107
    assert p.returncode == 1 and b"Invalid gzip compression level 12" in p.stderr


def test_cli_surface():
    # tests/cli_tests.rs: --version prints the version, no arguments is a usage failure
    p = run("--version")
    assert b"0.10.0" in p.stdout
    p = run(check=False)
    assert p.returncode != 0 and b"Usage" in p.stderr
    assert b"--abs-threshold" in run("filter", "--help").stdout
    assert b"union" in run("index", "--help").stdout
    p = run("filter", check=False)
    assert p.returncode == 2 and b"<INDEX>" in p.stderr
    p = run("filter", "x.idx", "-a", "0", check=False)   # clap range(1..), src/main.rs:44
    assert p.returncode == 2 and b"abs-threshold" in p.stderr
    p = run("index", "build", "-k", "58", "ref.fa", check=False)   # range(1..=57), src/main.rs:166
    assert p.returncode == 2
    p = run("frobnicate", check=False)
    assert p.returncode == 2 and b"unrecognized subcommand" in p.stderr
    p = run("server", "x.idx", check=False)
    assert p.returncode == 2 and b"not part of this build" in p.stderr
