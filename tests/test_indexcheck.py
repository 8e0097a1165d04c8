"""The size-independent index checks (smalt_b200/indexcheck.py, used on the GPU at genome sizes where the host
builder is too slow) accept the tables of the host builder - itself byte-identical with `smalt index`
(tests/test_indexer.py) - and reject corrupted ones."""
import numpy as np
import pytest

from seqgen import random_seq
from smalt_b200 import indexcheck, indexer

CASES = [(13, 6, [300_000, 70_007], 0.001), (7, 1, [3_000, 2_999], 0.002), (20, 13, [400_000, 300_001], 0.0005),
         (8, 5, [1_000_003], 0.0), (11, 3, [30_011, 999], 0.01)]


@pytest.mark.parametrize("k,nskip,lens,p_n", CASES)
def test_checks_accept_the_host_builder_and_reject_corruption(k, nskip, lens, p_n):
    rng = np.random.default_rng(40 + k)
    seqs = [random_seq(rng, n, p_n=p_n) for n in lens]
    ix = indexer.build_index(seqs, k, nskip)
    indexcheck.check_structure(ix)
    ngrid = int(sum((n - k) // nskip + 1 for n in lens))
    assert indexcheck.check_samples(ix, seqs, k, nskip, nsample=4000) > 3000
    # two positions of different words exchanged: every grid position sampled -> found missing under its word
    bad = dict(ix)
    p = ix["pos"].copy()
    a, b = 5, int(ix["npos"]) - 7
    p[[a, b]] = p[[b, a]]
    bad["pos"] = p
    with pytest.raises(AssertionError):
        indexcheck.check_structure(bad)
        indexcheck.check_samples(bad, seqs, k, nskip, nsample=20 * ngrid)
    # an offset array that is not monotone
    bad = dict(ix)
    i2 = ix["idx"].copy()
    nz = np.nonzero(np.diff(i2.astype(np.int64)) > 0)[0]
    i2[nz[len(nz) // 2] + 1] += 1
    bad["idx"] = i2
    with pytest.raises(AssertionError):
        indexcheck.check_structure(bad)
        indexcheck.check_samples(bad, seqs, k, nskip, nsample=20 * ngrid)
