#!/bin/bash
# Reference pin kit -- ONE command for whoever has a Rust toolchain (this image has none: no cargo / rustc, no network).
#
#   tools/make_reference_fixtures.sh [OUTDIR]          (default OUTDIR = tests/golden/reference_c1)
#
# Builds the UNMODIFIED reference (default features: the local engine), runs it on the inputs written by
# tests/gen_fixture_inputs.py and stores what it prints:
#   *.idx                    `deacon index build` output (k31 w15; k31 w15 -e 0.5; k41 w15; k21 w11)
#   *_debug.txt.gz           the `DEBUG: <id> hits=<h>/<n> keep=<bool> kmers=[...]` lines of `deacon filter --debug -t 1`
#                            (src/local_filter.rs:354-363: every single-end record; pairs with at least one hit)
#   *_summary.json           the -s summary (six counters, src/filter_common.rs:10-38)
# tests/test_reference_fixtures.py then checks the oracle (and, on a GPU box, the CUDA path through the C ABI)
# against them: key sets, per-record hits / total / keep, hit k-mers in order, counters.  Commit OUTDIR.
set -euo pipefail
HERE="$(cd "$(dirname "$0")/.." && pwd)"
OUT="${1:-$HERE/tests/golden/reference_c1}"
REF="${DEACON_REF:-/root/reference}"
command -v cargo >/dev/null || { echo "make_reference_fixtures: cargo not found (a Rust toolchain is required)" >&2; exit 2; }
[ -f "$REF/Cargo.toml" ] || { echo "make_reference_fixtures: reference checkout not found at $REF (set DEACON_REF)" >&2; exit 2; }
WORK="$(mktemp -d)"
trap 'rm -rf "$WORK"' EXIT
cp -r "$REF" "$WORK/deacon"                        # the reference tree may be read-only
(cd "$WORK/deacon" && cargo build --release --locked)
BIN="$WORK/deacon/target/release/deacon"
"$BIN" --version | tee "$WORK/version.txt"

mkdir -p "$OUT"
python "$HERE/tests/gen_fixture_inputs.py" "$OUT"
cp "$WORK/version.txt" "$OUT/reference_version.txt"
G="$OUT/genome.fa"

"$BIN" index build -q -k 31 -w 15 "$G" -o "$OUT/k31w15.idx"
"$BIN" index build -q -k 31 -w 15 -e 0.5 "$G" -o "$OUT/k31w15_e05.idx"
"$BIN" index build -q -k 41 -w 15 "$G" -o "$OUT/k41w15.idx"
"$BIN" index build -q -k 21 -w 11 "$G" -o "$OUT/k21w11.idx"

dbg() { grep '^DEBUG: ' | gzip -9n; }   # keep the per-record lines only
# single-end, search mode, defaults (-a 2 -r 0.01)
"$BIN" filter --debug -q -t 1 -s "$OUT/single_summary.json" "$OUT/k31w15.idx" "$OUT/reads_single.fq" -o /dev/null 2> >(dbg > "$OUT/single_debug.txt.gz")
# single-end, prefix 80, deplete
"$BIN" filter --debug -q -t 1 -d -p 80 -s "$OUT/single_p80_deplete_summary.json" "$OUT/k31w15.idx" "$OUT/reads_single.fq" -o /dev/null 2> >(dbg > "$OUT/single_p80_deplete_debug.txt.gz")
# paired, deplete
"$BIN" filter --debug -q -t 1 -d -s "$OUT/paired_deplete_summary.json" "$OUT/k31w15.idx" "$OUT/reads_r1.fq" "$OUT/reads_r2.fq" -o /dev/null -O /dev/null 2> >(dbg > "$OUT/paired_deplete_debug.txt.gz")
# long reads, search
"$BIN" filter --debug -q -t 1 -s "$OUT/long_summary.json" "$OUT/k31w15.idx" "$OUT/reads_long.fq" -o /dev/null 2> >(dbg > "$OUT/long_debug.txt.gz")
# k = 41 (u128 k-mer values) and k = 21, w = 11
"$BIN" filter --debug -q -t 1 -a 1 -s "$OUT/single_k41_summary.json" "$OUT/k41w15.idx" "$OUT/reads_single.fq" -o /dev/null 2> >(dbg > "$OUT/single_k41_debug.txt.gz")
"$BIN" filter --debug -q -t 1 -s "$OUT/single_k21_summary.json" "$OUT/k21w11.idx" "$OUT/reads_single.fq" -o /dev/null 2> >(dbg > "$OUT/single_k21_debug.txt.gz")
wait
gzip -9nf "$OUT"/genome.fa "$OUT"/reads_*.fq
ls -la "$OUT"
echo "fixtures written to $OUT: run  python -m pytest tests/test_reference_fixtures.py  and commit the directory"
