"""The magic-number division by the sampling step in the candidate kernels (csrc/block.cu StepDiv: floor(x / d) =
umulhi(x, floor(2^32 / d) + 1) for x < 2^24, d <= 255) is exact on its whole guarded range, and the guard is
needed: the identity fails just above it for large steps.  umulhi(x, m) = floor(x / d + x * e / 2^32) with
e = m - 2^32 / d in (0, 1], so a wrong quotient shows first where x mod d = d - 1: those residues are checked for
every step, a set of steps exhaustively."""
import numpy as np


def _magic(x, d):
    return (x * np.uint64((1 << 32) // d + 1)) >> np.uint64(32)


def test_magic_division_exact_below_the_guard():
    full = np.arange(0, 1 << 24, dtype=np.uint64)
    for d in (2, 3, 6, 13, 127, 128, 129, 200, 251, 253, 254, 255):
        assert np.array_equal(_magic(full, d), full // np.uint64(d)), d
    for d in range(2, 256):
        x = np.arange(d - 1, 1 << 24, d, dtype=np.uint64)            # the residues that fail first
        assert np.array_equal(_magic(x, d), x // np.uint64(d)), d
        x = np.arange(0, 1 << 24, d, dtype=np.uint64)                # and the multiples
        assert np.array_equal(_magic(x, d), x // np.uint64(d)), d


def test_guard_is_not_slack_for_large_steps():
    x = np.arange(1 << 24, 1 << 25, dtype=np.uint64)
    wrong = np.nonzero(_magic(x, 255) != x // np.uint64(255))[0]
    assert len(wrong) and int(x[wrong[0]]) == 16_909_559
