"""Paired-end mapping (rmapPair, rmap.c:1744-2112) through the paired wave scheduler
(rmap_wave.c: four block-wide passes incl. the rescue with per-pair on-the-fly indexes on the
device) and through the fiber scheduler (reference per-item code, hot-path calls batched):
SAM identical to the reference's CPU smalt, line by line, with one worker and a fixed draw seed."""
import json
import os
import subprocess
import sys

import pytest

from oracle_lib import ROOT, ref_binary

sys.path.insert(0, os.path.join(ROOT, "tools"))
import paired_check as pc  # noqa: E402

pytestmark = pytest.mark.gpu
needs = pytest.mark.skipif(ref_binary("smalt") is None or not os.path.exists(pc.B200),
                           reason="needs oracle/_ref and smalt_b200/bin")


def _both(tmp, args, env=None):
    ref, _ = pc.run(pc.REF, args, os.path.join(tmp, "ref.sam"))
    stats = os.path.join(tmp, "stats.json")
    got, _ = pc.run(pc.B200, args, os.path.join(tmp, "b200.sam"), dict(env or {}, SMALT_B200_STATS=stats))
    st = {}
    for line in open(stats):
        st.update(json.loads(line))
    return ref, got, st


def _same(ref, got):
    assert len(ref) == len(got)
    diff = [(a, b) for a, b in zip(ref, got) if a != b]
    assert not diff, "%d differing SAM lines, first:\n%s\n%s" % (len(diff), diff[0][0], diff[0][1])


@needs
def test_pairs_wave_all_passes(tmp_path):
    """repeats, unmappable mates on either side: every pass of the scheduler is taken"""
    tmp = str(tmp_path)
    pref, f1, f2 = pc.make(tmp, 6000, 2.0, 3, seed=11)
    ref, got, st = _both(tmp, ["-n", "1", "-i", "600", "-j", "200", pref, f1, f2])
    _same(ref, got)
    assert st["pairs"] == 6000 and st["pairs_third_pass"] > 100 and st["pairs_fourth_pass"] > 100
    assert st["pairs_by_reference_code"] < 60
    proper = sum(1 for l in ref if not l.startswith("@") and int(l.split("\t")[1]) & 2)
    assert proper > 9000


@needs
def test_pairs_fibers_only(tmp_path):
    """the same through the fiber scheduler alone (reference rmapPair on fibers)"""
    tmp = str(tmp_path)
    pref, f1, f2 = pc.make(tmp, 1500, 1.0, 2, seed=12)
    ref, got, st = _both(tmp, ["-n", "1", "-i", "600", "-j", "200", pref, f1, f2],
                         {"SMALT_B200_FIBERS_ONLY": "1", "SMALT_B200_FIBERS": "256"})
    _same(ref, got)
    assert st["fiber"]["items"] == 1500 and st["fiber"]["hits"] > 0 and st["fiber"]["bandali"] > 0


@needs
def test_pairs_exhaustive_option(tmp_path):
    """-x (RMAPFLG_ALLPAIR | NOSHRTINFO: full seed tables, every pair searched both ways) is not
    restated by the wave passes: fibers"""
    tmp = str(tmp_path)
    pref, f1, f2 = pc.make(tmp, 400, 0.5, 2, seed=13)
    ref, got, st = _both(tmp, ["-n", "1", "-x", "-i", "600", "-j", "200", pref, f1, f2])
    _same(ref, got)
    assert st["fiber"]["items"] == 400


@needs
def test_many_reference_sequences(tmp_path):
    """>= 512 reference sequences: whole-set hit lists (hashCollectHitsUsingCutoff, K1 mode 2),
    single-end reads through the fibers"""
    tmp = str(tmp_path)
    pref, f1, f2 = pc.make(tmp, 1200, 1.2, 600, seed=14, repeats=False)
    ref, got, st = _both(tmp, ["-n", "1", pref, f1])
    _same(ref, got)
    assert st["fiber"]["items"] == 1200
    mapped = sum(1 for l in ref if not l.startswith("@") and not int(l.split("\t")[1]) & 4)
    assert mapped > 1000


@needs
def test_pairs_in_process_mapper(tmp_path):
    """library API (include/smalt_b200_map.h): smbm_open_paired + smbm_map_fastq_pairs with worker
    threads - both SAM records of every pair in input order, equal to the reference where the
    placement is not a random draw among equals (MAPQ > 6 on both mates, cf. mthread_test.py)"""
    from smalt_b200.mapper import Mapper
    tmp = str(tmp_path)
    pref, f1, f2 = pc.make(tmp, 3000, 1.5, 3, seed=15)
    ref, _ = pc.run(pc.REF, ["-n", "1", "-i", "600", "-j", "200", pref, f1, f2], os.path.join(tmp, "ref.sam"))
    ref = [l for l in ref if not l.startswith("@")]
    m = Mapper(pref, nthreads=4, options=["-r", "7", "-i", "600", "-j", "200"], paired=True)
    try:
        a, b = open(f1, "rb").read(), open(f2, "rb").read()
        for rep in range(2):       # worker state persists between calls
            got = m.map_fastq_pairs(a, b).decode().splitlines()
            assert len(got) == len(ref) == 6000
            for k in range(0, 6000, 2):
                r1, r2, g1, g2 = ref[k].split("\t"), ref[k + 1].split("\t"), got[k].split("\t"), got[k + 1].split("\t")
                assert r1[0] == g1[0] and r2[0] == g2[0]
                if min(int(r1[4]), int(r2[4])) > 6:
                    assert ref[k] == got[k] and ref[k + 1] == got[k + 1], (k, ref[k], got[k])
        assert m.stats.n_reads == 6000
    finally:
        m.close()


@needs
def test_sample_and_map_with_histogram(tmp_path):
    """`smalt sample` (insert-size histogram; RMAPFLG_BEST | ALLPAIR on every readskip-th pair,
    smalt.c:1397: fibers) writes the reference's histogram, and `map -g <histogram>` (pair scores
    from the histogram, resultpairs.c) gives the reference's SAM"""
    tmp = str(tmp_path)
    pref, f1, f2 = pc.make(tmp, 3000, 1.0, 2, seed=16)
    hist = {}
    for tag, exe in (("ref", pc.REF), ("b200", pc.B200)):
        hist[tag] = os.path.join(tmp, tag + ".hist")
        r = subprocess.run([exe, "sample", "-u", "10", "-n", "1", "-o", hist[tag], pref, f1, f2],
                           capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, (tag, r.stderr[-1500:])
    a, b = open(hist["ref"]).read(), open(hist["b200"]).read()
    assert a == b and len(a.splitlines()) > 5
    ref, got, st = _both(tmp, ["-n", "1", "-g", hist["ref"], pref, f1, f2])
    _same(ref, got)
    assert st["pairs"] == 3000


@needs
def test_mapreads_pairs(tmp_path):
    """the multi-GPU entry point with a mate file: pairs sharded by rank (both files cut at the same
    record numbers), paired-end path per rank, SAM merged in input order"""
    import torch
    tmp = str(tmp_path)
    pref, f1, f2 = pc.make(tmp, 4000, 1.5, 3, seed=17)
    ref, _ = pc.run(pc.REF, ["-n", "1", "-i", "600", "-j", "200", pref, f1, f2], os.path.join(tmp, "ref.sam"))
    ref = [l for l in ref if not l.startswith("@")]
    out = os.path.join(tmp, "b200.sam")
    r = subprocess.run([sys.executable, "-m", "smalt_b200.mapreads", "-n", "1", "-r", "7", "-i", "600", "-j", "200",
                        "-o", out, pref, f1, f2], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    got = [l for l in open(out).read().splitlines() if not l.startswith("@")]
    assert got == ref                      # one rank, one worker: byte for byte
    if torch.cuda.device_count() < 2:
        return
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29571", "-m", "smalt_b200.mapreads",
                        "-n", "2", "-r", "7", "-i", "600", "-j", "200", "-o", out, pref, f1, f2],
                       capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    got = [l for l in open(out).read().splitlines() if not l.startswith("@")]
    assert len(got) == len(ref)
    for k in range(0, len(ref), 2):        # pairs without a random draw on either side
        f = [x.split("\t", 5) for x in (ref[k], ref[k + 1], got[k], got[k + 1])]
        assert f[0][0] == f[2][0]
        if min(int(x[4]) for x in f) > 6:
            assert ref[k] == got[k] and ref[k + 1] == got[k + 1], (k, ref[k], got[k])
