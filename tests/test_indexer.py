"""The host-side index builder/reader (smalt_b200/indexer.py, smifile.py) against the
reference's own `smalt index`: byte-identical .smi and .sma files."""
import os

import numpy as np
import pytest

from oracle_lib import ref_binary
from seqgen import random_seq
from smalt_b200 import indexer, smifile
from smalt_b200.seqpack import unpack3

LET = np.frombuffer(b"ACGTXN", np.uint8)


def write_fasta(path, names, seqs):
    with open(path, "w") as f:
        for n, s in zip(names, seqs):
            f.write(">%s\n" % n)
            a = LET[s].tobytes().decode()
            for k in range(0, len(a), 60):
                f.write(a[k:k + 60] + "\n")


CASES = [  # (seed, lengths, k, nskip)
    (1, [60011], 13, 6),
    (2, [30011, 20007, 999], 11, 3),
    (3, [5000, 13, 14, 4000], 13, 13),
    (4, [3000, 2999], 7, 1),      # perfect hash (4^7 <= 2*ntup)
    (5, [40000], 20, 13),         # k > 16: perfect low bits
    (6, [2500, 2500, 2500], 9, 2),
]


@pytest.mark.skipif(ref_binary("smalt") is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,lens,k,nskip", CASES)
def test_index_files_identical_to_reference(tmp_path, seed, lens, k, nskip):
    import subprocess
    rng = np.random.default_rng(seed)
    seqs = [random_seq(rng, n, p_n=0.001 if i % 2 else 0.0) for i, n in enumerate(lens)]
    names = ["s%d" % i for i in range(len(seqs))]
    fa = str(tmp_path / "g.fa")
    write_fasta(fa, names, seqs)
    subprocess.run([ref_binary("smalt"), "index", "-k", str(k), "-s", str(nskip), str(tmp_path / "ref"), fa],
                   check=True, capture_output=True)
    ix = indexer.build_index(seqs, k, nskip)
    indexer.write_smi(str(tmp_path / "mine"), ix)
    indexer.write_sma(str(tmp_path / "mine"), names, seqs)
    for ext in (".smi", ".sma"):
        a = open(str(tmp_path / "ref") + ext, "rb").read()
        b = open(str(tmp_path / "mine") + ext, "rb").read()
        assert a == b, ext
    # readers
    smi = smifile.read_smi(str(tmp_path / "ref"))
    ld = indexer.as_loaded(ix)
    for key in ("typ", "wordlen", "nskip", "nbits_key", "nbits_lo", "npos", "nwords", "maxpos"):
        assert smi[key] == ld[key]
    for key in ("idx", "pos", "wordidx", "posidx"):
        if ld[key] is None:
            assert smi[key] is None
        else:
            assert np.array_equal(smi[key], ld[key]), key
    sma = smifile.read_sma(str(tmp_path / "ref"))
    assert sma["names"] == names
    allc = np.concatenate(seqs)
    assert np.array_equal(unpack3(sma["words"], len(allc)), allc)
    assert list(sma["seq_offs"]) == list(np.concatenate([[0], np.cumsum(lens)]))


def test_pack_roundtrip():
    from smalt_b200.seqpack import pack3
    rng = np.random.default_rng(9)
    for n in (0, 1, 9, 10, 11, 1234):
        c = rng.integers(0, 8, n).astype(np.uint8)
        assert np.array_equal(unpack3(pack3(c), n), c)
