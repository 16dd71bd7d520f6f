"""How little the reference's own tests pin: of the 64 neighbouring hypotheses about the
un-vendored upstream arithmetic (ntHash seed order x low/high 32 bits x add/xor x top-16/full compare x TG/AC majority
x left/right on the canonical strand) every behavioural known-answer test of tests/filter_tests.rs rejects exactly ONE
family of 16; the other 48 pass.  The oracle and the CUDA kernels implement one of the 48 (the one recollection of the
upstream source points to).  A reference-derived fixture collapses the 48 to one: tests/test_reference_fixtures.py."""
import json
import os

from oracle import py_oracle as P
from oracle import py_variants as V

KATS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")))["cases"]


def test_working_hypothesis_is_the_oracles():
    """Variant() with all defaults is the restatement the rest of the suite checks the C oracle and the kernels against."""
    seq = (b"ATTAAAGGTTTATACCTTCCCAGGTAACAAACCAACCAACTTTCGATCTCTTGTAGATCTGTTCTCTAAACGAACTTTAAAATCTGTGTGGCTGTCACTCGGCTGCATGC"
           b"TTAGTGCACTNACGCAGTATAATTAATAACTAATTACTGTCGTTGACAGGACACGAGTAACTCGTCTATCTTCTGCAGGCTGCTTACGGTTTCGTCCGTGTTGCAGCCGA")
    for k, w in ((31, 15), (21, 11), (41, 15), (5, 5), (31, 1)):
        assert V.extract_filter(V.WORKING, seq, k, w) == P.extract_filter(seq, k, w)
        assert V.extract_index(V.WORKING, seq, k, w) == P.extract_index(seq, k, w)


def test_reference_tests_reject_one_family_of_sixteen():
    passing = [v for v in V.ALL if all(V.run_kat(v, c) for c in KATS)]
    failing = [v for v in V.ALL if v not in passing]
    assert V.WORKING in passing
    assert len(passing) == 48 and len(failing) == 16
    # the rejected family: every base takes its own seed (not the table indexed by packing code) with the low 32 seed bits
    assert all((not v.table_by_code) and v.low32 for v in failing)
    assert all(v.table_by_code or not v.low32 for v in passing)
