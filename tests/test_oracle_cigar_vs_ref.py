"""Pins the CIGAR / edit-distance restatement (oracle/smalt_oracle_cigar.c) against the UNMODIFIED reference
functions writeDiffStrCIGAR (through diffStrPrintfStr, diffstr.c:298-367, :1084-1121) and
diffStrGetLevenshteinDistance (diffstr.c:1496-1510) of oracle/_ref/libsmalt_ref.so, and against the known
answers of the reference's own test/bam_cigar_test.py."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from diffgen import encode_columns, random_bytes_string, random_columns
from golden_io import load_bam_cigar, parse_record
from oracle_lib import Oracle, RefLib, have_ref


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def _cases(rng):
    for n in (1, 2, 5, 30, 61, 62, 63, 64, 125, 150, 151, 300, 1000, 9000):
        for rep in range(6):
            px = (0.0, 0.02, 0.1, 0.5)[rep % 4]
            yield encode_columns(random_columns(rng, n, p_x=px, p_d=0.01 * rep, p_i=0.012 * rep))
    for cols in ("=", "X", "D", "I", "XX", "X=", "=X", "DI", "ID", "DDD", "=D=", "X" * 70, "=" * 62, "=" * 63, "=" * 124,
                 "=" * 125, "D" + "=" * 62 + "I", "XDXIX", "=" * 10 + "XX" + "=" * 10 + "DD" + "II" + "=" * 10):
        yield encode_columns(cols)
    for n in (1, 3, 10, 40, 200):
        for rep in range(20):
            yield random_bytes_string(rng, n)
    yield bytes([(3 << 6) | 0, 0])       # nothing but the closing byte
    yield bytes([(3 << 6) | 17, 0])


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_cigar_matches_reference(orc):
    ref = RefLib()
    rng = np.random.default_rng(77)
    n = 0
    for d in _cases(rng):
        for soft in (True, False):
            for xm in (False, True):
                cs, ce = (0, 0) if n % 3 == 0 else (int(rng.integers(0, 40)), int(rng.integers(0, 12000)))
                e, text, nm = ref.cigar(d, cs, ce, soft, xm)
                got_text, got_nm = orc.cigar(d, cs, ce, soft, xm)
                assert e == 0, (d, e)
                assert got_text == text, (d, cs, ce, soft, xm, text, got_text)
                assert got_nm == nm, (d, nm, got_nm)
                n += 1
    assert n > 700


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_cigar_error_codes_match_reference(orc):
    ref = RefLib()
    # empty string: ERRCODE_FAILURE (-1); a string that does not end with an S byte: ERRCODE_DIFFSTR (59)
    for d, want in ((b"\0", -1), (bytes([(0 << 6) | 5, 0]), 59), (bytes([(3 << 6) | 2, (1 << 6) | 1, 0]), 59),
                    (bytes([(2 << 6) | 0, 0]), 59)):
        e, _, _ = ref.cigar(d, 3, 4, True, False)
        assert e == want
        text, code = orc.cigar(d, 3, 4, True, False)
        assert text is None and code == (want if want < 0 else -want)


def test_cigar_known_answers(orc):
    """test/bam_cigar_test.py:3-51: CIGAR, X-CIGAR and NM of the alignments `smalt map` reports for the embedded
    reads, from the alignment strings of the recorded K3 calls (tests/golden/bam_cigar.json)"""
    g = load_bam_cigar()
    for key, xm in (("cigar", False), ("xcigar", True)):
        recs = [parse_record(l) for l in g["trace"][key]]
        for rd, sam in zip(g["reads"], g["sam"][key]):
            fld = sam.split("\t")
            qlen = len(rd["seq"])
            found = False
            for r in recs:
                if r["kind"] != "BA" or len(r["read"]) != qlen:
                    continue
                for (score, qs, qe, rs, re), dstr in r["results"]:
                    text, nm = orc.cigar(dstr, qs, qlen - 1 - qe, True, xm)
                    if text is not None and text.decode() == rd[key] and "AS:i:%d" % score in fld:
                        assert "NM:i:%d" % nm == rd["nm"]
                        found = True
            assert found, (rd[key], sam)
            assert fld[5] == rd[key] and rd["nm"] in fld


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@settings(max_examples=400, deadline=None)
@given(body=st.binary(min_size=0, max_size=60).map(lambda b: bytes(x or 1 for x in b)),
       last=st.integers(0, 63), cs=st.integers(0, 500), ce=st.integers(0, 20000), soft=st.booleans(), xm=st.booleans())
def test_cigar_fuzz_against_reference(body, last, cs, ce, soft, xm):
    """arbitrary non-zero bytes closed by an S byte: text, edit distance and error code of the restatement equal
    the reference's for strings no aligner writes"""
    orc, ref = Oracle(), RefLib()
    d = body + bytes([(3 << 6) | last, 0])
    e, text, nm = ref.cigar(d, cs, ce, soft, xm)
    got_text, got_nm = orc.cigar(d, cs, ce, soft, xm)
    assert e == 0 and got_text == text and got_nm == nm
