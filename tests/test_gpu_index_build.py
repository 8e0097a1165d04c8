"""Index construction on the GPU (csrc/index_build.cu, smb_index_build) against the host builder
that is byte-identical to `smalt index` (tests/test_indexer.py pins it against the compiled
reference): every array of the table, for perfect-hash and collision tables, several sequences
(k-mer grid carried across sequence boundaries), N bases, k up to 20; and the .smi file bytes."""
import os

import numpy as np
import pytest

from seqgen import random_seq
from smalt_b200 import indexer

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import smalt_b200
    c = smalt_b200.Context(0)
    yield c
    c.close()


CASES = [
    (13, 6, [300_000], 0.0),
    (13, 6, [50_011, 70_007, 999, 13], 0.001),
    (11, 3, [30_011, 20_007, 999], 0.01),
    (5, 1, [2_000], 0.0),              # perfect hash (the on-the-fly index of rmap.c:495-517)
    (7, 1, [3_000, 2_999], 0.002),     # perfect hash, two sequences
    (20, 13, [400_000, 300_001], 0.0005),
    (13, 13, [100_000], 0.0),
    (8, 5, [1_000_003], 0.0),          # perfect hash with sampling step
    (13, 3, [9_000_000, 4_000_001], 0.0001),   # 4.3 M positions: three levels of the scan, 1058 sort tiles
]


@pytest.mark.parametrize("k,nskip,lens,p_n", CASES)
def test_index_build_gpu_equals_host_builder(ctx, tmp_path, k, nskip, lens, p_n):
    rng = np.random.default_rng(500 + k + nskip)
    seqs = [random_seq(rng, n, p_n=p_n) for n in lens]
    want = indexer.build_index(seqs, k, nskip)
    got = indexer.build_index_gpu(ctx, seqs, k, nskip)
    for key in ("typ", "wordlen", "nskip", "nbits_key", "nbits_lo", "npos", "nwords", "maxpos", "nkeys"):
        assert got[key] == want[key], key
    for key in ("idx", "pos", "wordidx", "posidx"):
        if want[key] is None:
            assert got[key] is None
        else:
            assert np.array_equal(got[key], want[key]), key
    a, b = str(tmp_path / "host"), str(tmp_path / "gpu")
    indexer.write_smi(a, want)
    indexer.write_smi(b, got)
    assert open(a + ".smi", "rb").read() == open(b + ".smi", "rb").read()
    assert want["npos"] > 0


def test_index_build_gpu_large_genome_invariants(ctx):
    """120 Mb, 20 M positions (5000 sort tiles, three scan levels): too large for the numpy host builder, checked
    through the size-independent properties of smalt_b200/indexcheck.py (pinned against the host builder on CPU,
    tests/test_indexcheck.py): every offset array monotone and complete, words ascending inside a key, positions
    ascending inside a word (= the stable order of the sort), sampled grid positions found under their own words"""
    from smalt_b200 import indexcheck
    rng = np.random.default_rng(77)
    seqs = [rng.integers(0, 4, n, dtype=np.uint8) for n in (70_000_003, 49_999_999)]
    seqs[0][rng.integers(0, len(seqs[0]), 5000)] = 5
    ix = indexer.build_index_gpu(ctx, seqs, 13, 6)
    assert ix["typ"] == 1 and ix["npos"] > 19_000_000
    indexcheck.check_structure(ix)
    assert indexcheck.check_samples(ix, seqs, 13, 6, nsample=20000) > 19000


def _fasta(path, seqs):
    let = np.frombuffer(b"ACGTNN", np.uint8)
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">chr%d\n" % i)
            t = let[s].tobytes().decode()
            for j in range(0, len(t), 60):
                f.write(t[j:j + 60] + "\n")


@pytest.mark.parametrize("k,nskip,lens,p_n", [(13, 6, [400_000, 250_003, 1_777], 0.0005), (11, 2, [90_000], 0.0),
                                               (20, 13, [600_000, 500_000], 0.0), (7, 3, [60_000], 0.001)])
def test_cli_index_equals_reference(tmp_path, k, nskip, lens, p_n):
    """`smalt_b200 index` (hashTableSetUp on the GPU behind the reference's own driver and file
    writer) writes the same .smi and .sma bytes as the reference's `smalt index`"""
    import subprocess
    from oracle_lib import ROOT, ref_binary
    exe = os.path.join(ROOT, "smalt_b200", "bin", "smalt_b200")
    if ref_binary("smalt") is None or not os.path.exists(exe):
        pytest.skip("needs oracle/_ref and smalt_b200/bin")
    rng = np.random.default_rng(900 + k)
    seqs = [random_seq(rng, n, p_n=p_n) for n in lens]
    fa = str(tmp_path / "g.fa")
    _fasta(fa, seqs)
    for tag, binary in (("ref", ref_binary("smalt")), ("b200", exe)):
        r = subprocess.run([binary, "index", "-k", str(k), "-s", str(nskip), str(tmp_path / tag), fa],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (tag, r.stderr[-800:])
        if tag == "b200":
            assert "built on the GPU" in r.stderr, r.stderr[-800:]
    for ext in (".smi", ".sma"):
        assert open(str(tmp_path / "ref") + ext, "rb").read() == open(str(tmp_path / "b200") + ext, "rb").read(), ext
