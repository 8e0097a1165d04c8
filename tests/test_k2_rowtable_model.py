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
