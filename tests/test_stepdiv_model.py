"""The magic-number division by the sampling step in the candidate kernels (csrc/block.cu StepDiv: floor(x / d) =
umulhi(x, floor(2^32 / d) + 1) for x < 2^24, d <= 255) is exact on its whole guarded range - every d, every x - and
the guard is needed: the identity fails just above it for large steps."""
import numpy as np


def test_magic_division_exact_below_the_guard():
    x = np.arange(0, 1 << 24, dtype=np.uint64)
    for d in range(2, 256):
        m = np.uint64((1 << 32) // d + 1)
        assert np.array_equal((x * m) >> np.uint64(32), x // np.uint64(d)), d


def test_guard_is_not_slack_for_large_steps():
    x = np.arange(1 << 24, 1 << 25, dtype=np.uint64)
    m = np.uint64((1 << 32) // 255 + 1)
    wrong = np.nonzero(((x * m) >> np.uint64(32)) != x // np.uint64(255))[0]
    assert len(wrong) and int(x[wrong[0]]) == 16_909_559
