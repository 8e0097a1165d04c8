"""The in-process driver (include/smalt_b200_map.h, libsmalt_b200_map.so) returns the SAM
records the reference's `smalt map` prints for the same index and reads."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle_lib import ROOT, ref_binary
from seqgen import mutate, random_seq, revcomp
from smalt_b200 import indexer

pytestmark = pytest.mark.gpu
MAPLIB = os.path.join(ROOT, "smalt_b200", "libsmalt_b200_map.so")
LET = np.frombuffer(b"ACGTNN", np.uint8)


def _workload(tmp_path, seed, lens, k, s, nreads, qlen, err, fasta=False):
    rng = np.random.default_rng(seed)
    seqs = [random_seq(rng, n, p_n=0.0002) for n in lens]
    pref = str(tmp_path / "idx")
    indexer.write_smi(pref, indexer.build_index(seqs, k, s))
    indexer.write_sma(pref, ["chr%d" % i for i in range(len(seqs))], seqs)
    recs = []
    for i in range(nreads):
        g = seqs[int(rng.integers(0, len(seqs)))]
        L = int(rng.integers(qlen[0], qlen[1]))
        st = int(rng.integers(0, len(g) - L))
        rd = mutate(rng, g[st:st + L].copy(), p_sub=err, p_ins=err / 8, p_del=err / 8)
        if i % 2:
            rd = revcomp(rd)
        if i % 41 == 0:
            rd = random_seq(rng, len(rd))
        if i % 89 == 0:
            rd = rd[:7]
        sq = LET[rd].tobytes().decode()
        if fasta:
            recs.append(">q%d extra words\n%s\n" % (i, sq))
        else:
            recs.append("@q%d\n%s\n+\n%s\n" % (i, sq, "".join(chr(33 + 5 + (3 * i + j) % 26) for j in range(len(sq)))))
    text = "".join(recs).encode()
    fq = str(tmp_path / ("reads.fa" if fasta else "reads.fq"))
    open(fq, "wb").write(text)
    return pref, fq, text


def _ref_sam(tmp_path, pref, fq, extra=()):
    out = str(tmp_path / "ref.sam")
    # fixed seed for the draw among equally good hits (the default seeds it from the clock)
    r = subprocess.run([ref_binary("smalt"), "map", "-r", "7"] + list(extra) + ["-o", out, pref, fq],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1000:]
    lines = open(out).read().splitlines()
    return [l for l in lines if l.startswith("@") and not l.startswith("@PG")], [l for l in lines if not l.startswith("@")]


needs = pytest.mark.skipif(ref_binary("smalt") is None or not os.path.exists(MAPLIB),
                           reason="needs oracle/_ref/smalt and libsmalt_b200_map.so")


@needs
def test_mapper_matches_reference(tmp_path):
    from smalt_b200.mapper import Mapper
    pref, fq, text = _workload(tmp_path, 11, [300_000, 77_777], 13, 6, 6000, (36, 251), 0.025)
    hdr, want = _ref_sam(tmp_path, pref, fq)
    m = Mapper(pref, 1, ["-r", "7"])         # one worker: the same read (and random draw) order as the serial reference
    try:
        got = m.map_fastq(text).decode().splitlines()
        assert got == want
        assert m.stats.n_reads == 6000 and m.stats.k2_tasks > 0 and m.stats.k3_cells > 0
        # CIGAR and NM of every mapped record came from the device's output stage (csrc/cigar.cu), none from diffstr.c
        nmapped = sum(1 for l in want if l.split("\t")[5] != "*")
        assert m.stats.cigar_dev == nmapped > 5000 and m.stats.cigar_host == 0
        # a second call on the same mapper, a sub-range of the reads
        cut = text.index(b"@q3000\n")
        assert m.map_fastq(text[cut:]).decode().splitlines() == want[3000:]
        assert m.map_fastq(b"") == b""
        got_hdr = [l for l in m.sam_header().decode().splitlines() if not l.startswith("@PG")]
        assert got_hdr == hdr
    finally:
        m.close()
    # a fresh mapper in the same process, several workers, FASTA reads, other penalties
    pref2, fa, text2 = _workload(tmp_path, 12, [120_000], 11, 3, 3000, (50, 180), 0.04, fasta=True)
    opts = ["-S", "subst=-3", "-m", "40"]   # (the reference itself fails with ERRCODE_SWATSCOR for other gap penalties)
    _, want2 = _ref_sam(tmp_path, pref2, fa, opts)
    m = Mapper(pref2, 4, ["-r", "7"] + opts)
    try:
        got2 = m.map_fastq(text2).decode().splitlines()
    finally:
        m.close()
    assert len(got2) == len(want2)
    # several workers: reads with equally good hits are reported at a drand48-chosen one, whose
    # stream is shared by the workers (results.c:2298) - compare where MAPQ > 6 like the
    # reference's own test/mthread_test.py:40-101
    diff = [(a, b) for a, b in zip(got2, want2) if a != b and (int(a.split("\t")[4]) > 6 or int(b.split("\t")[4]) > 6)]
    assert not diff, diff[0]


@needs
@pytest.mark.parametrize("opts", [["-d", "3"], ["-d", "-1"], ["-y", "0.9"], ["-q", "10"], ["-f", "cigar"], ["-f", "ssaha"],
                                  ["-f", "sam:x"], ["-f", "sam:clip"], ["-f", "sam:x,clip"]])
def test_mapper_options(tmp_path, opts):
    """modes other than the default best-hit SAM: score range (-d), identity filter (-y), base
    quality threshold for seeds (-q), other text formats (-f); the SAM variants (X-CIGAR, hard clips) take their
    CIGAR / NM from the device's output stage"""
    from smalt_b200.mapper import Mapper
    pref, fq, text = _workload(tmp_path, 21, [150_000, 50_000], 13, 6, 2500, (60, 200), 0.03)
    _, want = _ref_sam(tmp_path, pref, fq, opts)
    m = Mapper(pref, 1, ["-r", "7"] + opts)
    try:
        got = m.map_fastq(text).decode().splitlines()
        if opts[0] == "-f" and opts[1].startswith("sam"):
            assert m.stats.cigar_dev > 2000 and m.stats.cigar_host == 0
    finally:
        m.close()
    assert got == want


@needs
@pytest.mark.parametrize("workers,force", [(1, "0"), (12, "1"), (12, None)])
def test_cigar_stage_placement(tmp_path, monkeypatch, workers, force):
    """the CIGAR / NM stage runs on the device up to 8 workers and on the host above (it follows the bottleneck,
    fastmap.inc.c); SMALT_B200_DEVCIGAR forces either - the records are the reference's in every case"""
    from smalt_b200.mapper import Mapper
    if force is None:
        monkeypatch.delenv("SMALT_B200_DEVCIGAR", raising=False)
    else:
        monkeypatch.setenv("SMALT_B200_DEVCIGAR", force)
    pref, fq, text = _workload(tmp_path, 31, [200_000], 13, 6, 20000, (80, 200), 0.03)
    _, want = _ref_sam(tmp_path, pref, fq)
    m = Mapper(pref, workers, ["-r", "7"])
    try:
        got = m.map_fastq(text).decode().splitlines()
        dev, host = m.stats.cigar_dev, m.stats.cigar_host
    finally:
        m.close()
    nmapped = sum(1 for l in got if l.split("\t")[5] != "*")
    if force == "1":
        assert dev == nmapped > 15000 and host == 0
    else:
        assert host == nmapped > 15000 and dev == 0
    assert len(got) == len(want)
    if workers == 1:
        assert got == want
    else:   # several workers: the draws among equally good hits are scheduling dependent (see above)
        diff = [(a, b) for a, b in zip(got, want) if a != b and (int(a.split("\t")[4]) > 6 or int(b.split("\t")[4]) > 6)]
        assert not diff, diff[0]


@needs
def test_mapreads_entry_point(tmp_path):
    """the multi-GPU entry point with world size 1 (sharding and merge are tested on CPU)"""
    pref, fq, text = _workload(tmp_path, 13, [200_000], 13, 6, 2000, (100, 101), 0.01)
    _, want = _ref_sam(tmp_path, pref, fq)
    out = str(tmp_path / "b200.sam")
    r = subprocess.run([sys.executable, "-m", "smalt_b200.mapreads", "-n", "1", "-r", "7", "-o", out, pref, fq],
                       capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    got = [l for l in open(out).read().splitlines() if not l.startswith("@")]
    assert got == want


@needs
def test_mapreads_two_gpus(tmp_path):
    """two ranks, two GPUs: reads sharded by rank, SAM merged in input order == reference"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    pref, fq, text = _workload(tmp_path, 14, [400_000], 13, 6, 20000, (150, 151), 0.02)
    _, want = _ref_sam(tmp_path, pref, fq)
    out = str(tmp_path / "b200.sam")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29561", "-m", "smalt_b200.mapreads",
                        "-n", "1", "-r", "7", "-o", out, pref, fq], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    got = [l for l in open(out).read().splitlines() if not l.startswith("@")]
    # each rank draws from its own drand48 stream: compare like the reference's threaded test
    assert len(got) == len(want)
    diff = [(a, b) for a, b in zip(got, want) if a != b and (int(a.split("\t")[4]) > 6 or int(b.split("\t")[4]) > 6)]
    assert not diff, diff[0]


@needs
def test_fast_record_loader_equals_reference_parser(tmp_path):
    """read names with comments, tabs and runs of blanks, '+name' separator lines, quality lines
    that start with '@' or '+', lower-case and ambiguous bases (encoded while they are loaded),
    CRLF records (those fall back to the reference parser)"""
    from smalt_b200.mapper import Mapper
    rng = np.random.default_rng(31)
    g = random_seq(rng, 80_000)
    pref = str(tmp_path / "idx")
    indexer.write_smi(pref, indexer.build_index([g], 13, 6))
    indexer.write_sma(pref, ["chrT"], [g])
    recs = []
    for i in range(1500):
        L = int(rng.integers(40, 130))
        st = int(rng.integers(0, len(g) - L))
        sq = LET[mutate(rng, g[st:st + L].copy(), p_sub=0.02, p_ins=0.002, p_del=0.002)].tobytes().decode()
        if i % 7 == 0:
            sq = sq.lower()                   # lower case, N, other IUPAC letters: the encoder's table
        if i % 11 == 0:
            k = int(rng.integers(0, len(sq)))
            sq = sq[:k] + "NnRy"[i % 4] + sq[k + 1:]
        q = "".join(chr(int(x)) for x in rng.integers(40, 74, len(sq)))
        if i % 3 == 0:
            q = "@" + q[1:]
        if i % 4 == 0:
            q = "+" + q[1:]
        name = ["r%d", "r%d  two  blanks ", "r%d\tafter tab", " r%d lead", "r%d/1 comment:x=1"][i % 5] % i
        eol = "\r\n" if 700 <= i < 720 else "\n"
        recs.append("@%s%s%s%s+%s%s%s%s" % (name, eol, sq, eol, ("r%d" % i) if i % 2 else "", eol, q, eol))
    text = "".join(recs).encode()
    fq = str(tmp_path / "t.fq")
    open(fq, "wb").write(text)
    _, want = _ref_sam(tmp_path, pref, fq)
    m = Mapper(pref, 1, ["-r", "7"])
    try:
        os.environ["SMALT_B200_BLOCK"] = "200"
        fast = m.map_fastq(text).decode().splitlines()
        os.environ["SMALT_B200_REFPARSE"] = "1"
        slow = m.map_fastq(text).decode().splitlines()
    finally:
        os.environ.pop("SMALT_B200_REFPARSE", None)
        os.environ.pop("SMALT_B200_BLOCK", None)
        m.close()
    assert slow == want
    assert fast == want
