"""Pins the C restatement (oracle/smalt_oracle.c) against the UNMODIFIED reference
(oracle/_ref/libsmalt_ref.so built from /root/reference/src): DP part (K2, K2', K3)."""
import numpy as np
import pytest

from oracle_lib import Oracle, RefLib, have_ref
from seqgen import random_seq, read_window_pair

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def libs():
    return Oracle(), RefLib()


def test_sw_striped_matches_reference(libs):
    orc, ref = libs
    rng = np.random.default_rng(11)
    n = 0
    for qlen in (8, 15, 16, 17, 31, 32, 33, 64, 100, 150, 151, 250, 400):
        for rep in range(12):
            read, win = read_window_pair(rng, qlen, p_sub=0.03 * (rep % 4), p_ins=0.01 * (rep % 3),
                                         p_del=0.01 * (rep % 3))
            if rep % 5 == 4:
                read[rng.integers(0, len(read))] = 5
                win[rng.integers(0, len(win))] = 5
                # code 4 (X) cannot be produced by the reference codec (all non-ACGT
                # letters encode to N, sequence.c:303-304), so it is not exercised here
            assert orc.sw_striped(read, win) == ref.sw_striped(read, win)
            n += 1
    # unrelated sequences and scores above 255 (forces the reference's 16-bit retry)
    for qlen in (300, 600, 1200):
        read = random_seq(rng, qlen)
        other = random_seq(rng, qlen + 50)
        assert orc.sw_striped(read, other) == ref.sw_striped(read, other)
        win = np.concatenate([random_seq(rng, 20), read, random_seq(rng, 20)])
        e, s = ref.sw_striped(read, win)
        assert (e, s) == (0, qlen)
        assert orc.sw_striped(read, win) == (e, s)
    assert n > 100


def _band_args(rng, qlen, rlen, lf=None):
    style = rng.integers(0, 6)
    if style == 0 or style > 3:   # what rmap.c passes: whole read, band around the hit diagonal
        off = int(rng.integers(-30, 10)) if lf is None else -lf + int(rng.integers(-6, 7))
        w = int(rng.integers(2, 40))
        return off - w, off + w, 0, qlen - 1, 0, rlen - 1
    if style == 1:   # sub-ranges
        pl = int(rng.integers(0, qlen // 2))
        pr = int(rng.integers(qlen // 2, qlen + 3))
        ul = int(rng.integers(0, rlen // 2))
        ur = int(rng.integers(rlen // 2, rlen + 3))
        l = int(rng.integers(-rlen, qlen))
        return l, l + int(rng.integers(0, 60)), pl, pr, ul, ur
    if style == 2:   # degenerate / inverted band
        l = int(rng.integers(-50, 50))
        return l, l - int(rng.integers(0, 5)), 0, qlen - 1, 0, rlen - 1
    l = int(rng.integers(-2 * rlen, 2 * qlen))
    return l, l + int(rng.integers(0, 2 * qlen)), int(rng.integers(-2, qlen)), int(rng.integers(-2, qlen + 2)), \
        int(rng.integers(-2, rlen)), int(rng.integers(-2, rlen + 2))


def test_band_fast_matches_reference(libs):
    orc, ref = libs
    rng = np.random.default_rng(12)
    nfail = 0
    for it in range(600):
        qlen = int(rng.integers(20, 200))
        read, win, lf = read_window_pair(rng, qlen, with_flank=True, p_sub=0.03, p_ins=0.01, p_del=0.01)
        args = _band_args(rng, len(read), len(win), lf)
        eo, so, _ = orc.band_fast(read, win, *args)
        er, sr = ref.band_fast(read, win, *args)
        assert eo == er, (it, args)
        if er == 0:
            assert so == sr, (it, args)
        else:
            nfail += 1
    assert nfail < 500


def test_band_align_matches_reference(libs):
    orc, ref = libs
    rng = np.random.default_rng(13)
    nres_tot = 0
    for it in range(1500):
        qlen = int(rng.integers(20, 220))
        mut = dict(p_sub=0.04, p_ins=0.015, p_del=0.015) if it % 3 else dict(p_sub=0.1, p_ins=0.05, p_del=0.05)
        read, win, lf = read_window_pair(rng, qlen, with_flank=True, **mut)
        if it % 7 == 0:  # two copies in the window -> exercises the recursion
            win = np.concatenate([win, read_window_pair(rng, qlen)[1][:10], win])
        if it % 11 == 0:
            read[rng.integers(0, len(read))] = 5
        args = _band_args(rng, len(read), len(win), lf)
        minscore = int(rng.integers(1, 40))
        minscorlen = int(rng.integers(5, 30))
        eo, ro, _ = orc.band_align(read, win, *args, minscore, minscorlen)
        er, rr = ref.band_align(read, win, *args, minscore, minscorlen)
        assert eo == er, (it, args, minscore, minscorlen)
        assert ro == rr, (it, args, minscore, minscorlen)
        nres_tot += len(rr)
    assert nres_tot > 300
