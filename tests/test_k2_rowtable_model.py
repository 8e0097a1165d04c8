"""Model check of the K2 row-table score lookup (csrc/sw_score.cu, sw2_build_lut / sw2_qsel_tab): one byte permute
over the window row's score table must give the substitution scores of BOTH tasks of a packed cell pair as
sign-extended 16-bit values, for every base pair; padding columns must never score above 0.  PRMT is emulated as
the PTX ISA defines it (default mode: selector nibble bits 0-2 pick a byte of {b, a}, bit 3 replicates its sign).
The GPU tests compare the kernels with the oracle; this pins the selector / table constants on CPU."""
import itertools

import pytest


def prmt(a, b, sel):
    src = [(a >> (8 * i)) & 0xFF for i in range(4)] + [(b >> (8 * i)) & 0xFF for i in range(4)]
    out = 0
    for i in range(4):
        nib = (sel >> (4 * i)) & 0xF
        byte = src[nib & 7]
        if nib & 8:
            byte = 0xFF if byte & 0x80 else 0x00
        out |= byte << (8 * i)
    return out


def tab(x, match, mismatch):          # sw2_build_lut: score bytes s(q = 0..3, x) of a window base x
    t = 0
    for q in range(4):
        v = (match if q == x else mismatch) if x < 4 else (mismatch if x == 4 else 0)
        t |= (v & 0xFF) << (8 * q)
    return t


def qsel_tab(qa, qb):                 # sw2_qsel_tab: codes 0..3 bases, 8 = padding column
    lo = (qa | ((qa | 8) << 4)) if qa < 4 else 0x88
    hi = ((qb | 4) | ((qb | 12) << 4)) if qb < 4 else 0xCC
    return lo | (hi << 8)


def s16(x):
    return x - 0x10000 if x & 0x8000 else x


@pytest.mark.parametrize("match,mismatch", [(1, -2), (2, -3), (5, -4), (127, -127), (1, 0)])
def test_row_table_prmt_gives_both_scores(match, mismatch):
    for a, b in itertools.product(range(8), repeat=2):          # window bases of task A / B (4 = X, 5..7 = N, padding)
        if a == 4 or b == 4:
            continue                                            # X takes the per-cell table path
        ta, tb = tab(a, match, mismatch), tab(b, match, mismatch)
        for qa, qb in itertools.product((0, 1, 2, 3, 8), repeat=2):
            s2 = prmt(ta, tb, qsel_tab(qa, qb))
            for q, r, got in ((qa, a, s16(s2 & 0xFFFF)), (qb, b, s16(s2 >> 16))):
                if q == 8:                                      # padding column: anything <= 0
                    assert got in (0, -1)
                elif r >= 5:                                    # N / padding row: 0
                    assert got == 0
                else:
                    assert got == (match if q == r else mismatch)


def qsel_masked(qa, qb):              # sw2_qsel_masked: N and padding columns read the zero half of the table
    ia, ib = (qa if qa < 4 else 4), (qb if qb < 4 else 4)
    return (ia | ((ia | 8) << 4) | (ib << 8) | ((ib | 8) << 12)) ^ 0x4444


def wsel(a, b):                       # sw2_build_lut: row selector nibbles + N / padding masks of the masked form
    return (a * 0x11 if a < 4 else 0x00440000) | (b * 0x1100 if b < 4 else 0x44000000)


@pytest.mark.parametrize("match,mismatch", [(1, -2), (3, -5), (127, -127)])
def test_masked_form_scores_zero_for_n_in_reads_and_rows(match, mismatch):
    """the form pairs with an N in a READ take (short and long kernel): table {match, mismatch x3 | 0 x4} as the
    second PRMT source, index = (read selector ^ row selector) & ~row mask"""
    t0 = (match & 0xFF) | ((mismatch & 0xFF) * 0x01010100)
    for a, b in itertools.product((0, 1, 2, 3, 5, 6, 7), repeat=2):
        w = wsel(a, b)
        for qa, qb in itertools.product((0, 1, 2, 3, 5, 7, 8), repeat=2):
            s2 = prmt(0, t0, (qsel_masked(qa, qb) ^ w) & ~(w >> 16) & 0xFFFF)
            for q, r, got in ((qa, a, s16(s2 & 0xFFFF)), (qb, b, s16(s2 >> 16))):
                want = 0 if (q >= 4 or r >= 4) else (match if q == r else mismatch)
                assert got == want, (a, b, qa, qb, got, want)
