"""Small-scale parity runs shaped like the BASELINE.json configs that are not bench lines:
C3 (paired-end 2x150 bp with insert size, through the reference's rmapPair on the one-call
shim), C4 (k=20 s=13 index) and C5 (multi-kb reads with indel-rich errors: wide bands, K2
column blocks, K3 thread-per-task path)."""
import os
import subprocess

import numpy as np
import pytest

from oracle_lib import ROOT, ref_binary
from seqgen import mutate, random_seq, revcomp
from smalt_b200 import indexer

pytestmark = pytest.mark.gpu
B200 = os.path.join(ROOT, "smalt_b200", "bin", "smalt_b200")
LET = np.frombuffer(b"ACGTNN", np.uint8)
needs = pytest.mark.skipif(ref_binary("smalt") is None or not os.path.exists(B200),
                           reason="needs oracle/_ref and smalt_b200/bin")


def _index(tmp_path, seqs, k, s):
    pref = str(tmp_path / "idx")
    indexer.write_smi(pref, indexer.build_index(seqs, k, s))
    indexer.write_sma(pref, ["chr%d" % i for i in range(len(seqs))], seqs)
    return pref


def _fq(path, names, reads):
    with open(path, "w") as f:
        for nm, r in zip(names, reads):
            s = LET[r].tobytes().decode()
            f.write("@%s\n%s\n+\n%s\n" % (nm, s, "I" * len(s)))


def _run_both(tmp_path, args, timeout=900):
    outs = {}
    for tag, exe in (("ref", ref_binary("smalt")), ("b200", B200)):
        out = str(tmp_path / (tag + ".sam"))
        r = subprocess.run([exe, "map", "-r", "7", "-o", out] + args, capture_output=True, text=True, timeout=timeout)
        assert r.returncode == 0, (tag, r.stderr[-1500:])
        outs[tag] = [l for l in open(out).read().splitlines() if not l.startswith("@PG")]
    return outs["ref"], outs["b200"]


@needs
def test_c3_paired_end(tmp_path):
    rng = np.random.default_rng(41)
    g = [random_seq(rng, 300_000), random_seq(rng, 200_000)]
    pref = _index(tmp_path, g, 13, 6)
    r1, r2, names = [], [], []
    for i in range(300):
        s = g[i % 2]
        ins = int(rng.normal(400, 40))
        st = int(rng.integers(0, len(s) - ins - 1))
        frag = s[st:st + ins]
        a = mutate(rng, frag[:150].copy(), p_sub=0.02, p_ins=0.002, p_del=0.002)
        b = revcomp(mutate(rng, frag[-150:].copy(), p_sub=0.02, p_ins=0.002, p_del=0.002))
        if i % 37 == 0:
            b = random_seq(rng, 150)       # mate that does not map
        r1.append(a); r2.append(b); names.append("p%d" % i)
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    _fq(f1, [n + "/1" for n in names], r1)
    _fq(f2, [n + "/2" for n in names], r2)
    ref, got = _run_both(tmp_path, ["-i", "600", "-j", "200", pref, f1, f2])
    assert len(ref) == len(got) == 2 * 300 + 3
    diff = [(a, b) for a, b in zip(ref, got) if a != b]
    assert not diff, "%d differing lines, first:\n%s\n%s" % (len(diff), diff[0][0], diff[0][1])
    proper = sum(1 for l in ref if not l.startswith("@") and int(l.split("\t")[1]) & 2)
    assert proper > 500


@needs
def test_c4_like_k20_s13(tmp_path):
    rng = np.random.default_rng(42)
    g = [random_seq(rng, 700_000), random_seq(rng, 600_000)]
    pref = _index(tmp_path, g, 20, 13)
    reads, names = [], []
    for i in range(3000):
        s = g[i % 2]
        st = int(rng.integers(0, len(s) - 150))
        rd = mutate(rng, s[st:st + 150].copy(), p_sub=0.02, p_ins=0.002, p_del=0.002)
        reads.append(revcomp(rd) if i % 2 else rd)
        names.append("q%d" % i)
    fq = str(tmp_path / "r.fq")
    _fq(fq, names, reads)
    ref, got = _run_both(tmp_path, [pref, fq])
    diff = [(a, b) for a, b in zip(ref, got) if a != b]
    assert not diff, "%d differing lines, first:\n%s\n%s" % (len(diff), diff[0][0], diff[0][1])
    assert sum(1 for l in ref if not l.startswith("@") and not int(l.split("\t")[1]) & 4) > 2500


@needs
def test_c5_like_long_reads(tmp_path):
    rng = np.random.default_rng(43)
    g = [random_seq(rng, 400_000)]
    pref = _index(tmp_path, g, 13, 6)
    reads, names = [], []
    for i in range(24):
        L = int(rng.integers(2000, 6000))
        st = int(rng.integers(0, len(g[0]) - L))
        rd = mutate(rng, g[0][st:st + L].copy(), p_sub=0.04, p_ins=0.04, p_del=0.04)   # 12 % errors, indel rich
        reads.append(revcomp(rd) if i % 2 else rd)
        names.append("long%d" % i)
    fq = str(tmp_path / "r.fq")
    _fq(fq, names, reads)
    ref, got = _run_both(tmp_path, [pref, fq])
    diff = [(a, b) for a, b in zip(ref, got) if a != b]
    assert not diff, "%d differing lines, first:\n%s\n%s" % (len(diff), diff[0][0][:300], diff[0][1][:300])
    assert sum(1 for l in ref if not l.startswith("@") and not int(l.split("\t")[1]) & 4) >= 20
