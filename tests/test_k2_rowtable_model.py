"""Model check of the K2 score lookup without a GPU.  The constants - row tables, row selectors and masks, column
selectors of both forms, entry numbers (csrc/sw2_lut.cuh, `__host__ __device__`) - are compiled for the host from
the very header the kernels include; PRMT is emulated as the PTX ISA defines it (default mode: selector nibble bits
0-2 pick a byte of {b, a}, bit 3 replicates its sign).  One byte permute must give the substitution scores of BOTH
tasks of a packed cell pair as sign-extended 16-bit values for every pair of bases; padding columns must never score
above 0 in the row-table form; N in a read or a row must score exactly 0 in the masked form.  (The GPU tests
compare the kernels with the oracle; this pins the constants on CPU.)"""
import ctypes as C
import itertools
import os
import shutil
import subprocess

import pytest

from oracle_lib import ROOT

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def lut(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("no nvcc")
    so = str(tmp_path_factory.mktemp("sw2") / "libsw2lut.so")
    r = subprocess.run([NVCC, "-std=c++17", "-O1", "-shared", "-Xcompiler", "-fPIC", "-gencode",
                        "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "smalt_b200", "csrc"),
                        os.path.join(ROOT, "tests", "c", "sw2_lut_host.cu"), "-o", so], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    lib = C.CDLL(so)
    for f in ("sw2h_lut_index", "sw2h_tab_word", "sw2h_wsel_word", "sw2h_qsel_tab", "sw2h_qsel_masked"):
        getattr(lib, f).restype = C.c_uint
    return lib


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


def s16(x):
    return x - 0x10000 if x & 0x8000 else x


@pytest.mark.parametrize("match,mismatch", [(1, -2), (2, -3), (5, -4), (127, -127), (1, 0)])
def test_row_table_prmt_gives_both_scores(lut, match, mismatch):
    for a, b in itertools.product(range(8), repeat=2):          # window bases of task A / B (4 = X, 5..7 = N, padding)
        if a == 4 or b == 4:
            continue                                            # X takes the per-cell table path
        ta, tb = lut.sw2h_tab_word(a, match, mismatch), lut.sw2h_tab_word(b, match, mismatch)
        for qa, qb in itertools.product((0, 1, 2, 3, 8), repeat=2):
            s2 = prmt(ta, tb, lut.sw2h_qsel_tab(qa, qb))
            for q, r, got in ((qa, a, s16(s2 & 0xFFFF)), (qb, b, s16(s2 >> 16))):
                if q == 8:                                      # padding column: anything <= 0
                    assert got in (0, -1)
                elif r >= 5:                                    # N / padding row: 0
                    assert got == 0
                else:
                    assert got == (match if q == r else mismatch)


@pytest.mark.parametrize("match,mismatch", [(1, -2), (3, -5), (127, -127)])
def test_masked_form_scores_zero_for_n_in_reads_and_rows(lut, match, mismatch):
    """the form pairs with an N in a READ take (short and long kernel): table {match, mismatch x3 | 0 x4} as the
    second PRMT source, index = (read selector ^ row selector) & ~row mask"""
    t0 = (match & 0xFF) | ((mismatch & 0xFF) * 0x01010100)
    for a, b in itertools.product((0, 1, 2, 3, 5, 6, 7), repeat=2):
        w = lut.sw2h_wsel_word(a, b)
        for qa, qb in itertools.product((0, 1, 2, 3, 5, 7, 8), repeat=2):
            s2 = prmt(0, t0, (lut.sw2h_qsel_masked(qa, qb) ^ w) & ~(w >> 16) & 0xFFFF)
            for q, r, got in ((qa, a, s16(s2 & 0xFFFF)), (qb, b, s16(s2 >> 16))):
                want = 0 if (q >= 4 or r >= 4) else (match if q == r else mismatch)
                assert got == want, (a, b, qa, qb, got, want)


def test_entry_numbers(lut):
    """distinct entries for the 64 base pairs inside the table; the 16 pairs of standard bases are the first 16
    8-byte elements (32 different banks)"""
    idx = {(a, b): lut.sw2h_lut_index(a, b) for a in range(8) for b in range(8)}
    assert len(set(idx.values())) == 64 and max(idx.values()) < lut.sw2h_lut_n()
    assert sorted(idx[(a, b)] for a in range(4) for b in range(4)) == list(range(16))
    assert max(idx.values()) * 8 < 65536     # rows are staged as 16-bit byte offsets
