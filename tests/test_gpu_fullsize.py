"""BASELINE.json configs[1] at FULL size (5 Mb genome, 1 M single-end 150 bp reads): the whole
program against the reference's own CPU program on the same files, plus properties that do not
need the reference - every read reported once and in input order, and the simulated origin of
the reads recovered."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle_lib import ROOT, ref_binary

sys.path.insert(0, ROOT)
import bench  # noqa: E402

pytestmark = pytest.mark.gpu
B200 = os.path.join(ROOT, "smalt_b200", "bin", "smalt_b200")


@pytest.mark.skipif(ref_binary("smalt") is None or not os.path.exists(B200), reason="needs oracle/_ref and smalt_b200/bin")
def test_c2_full_size(tmp_path):
    n = 1_000_000
    wl = bench.Workload(str(tmp_path), bench.CONFIGS["c2"], n)     # misc/simread 150 1000000 2.0 y 0 0 43
    pref, fq = wl.pref, wl.files[0]
    names = [ln[1:] for ln in wl.texts[0].split(b"\n")[0::4] if ln]
    cores = bench.host_threads()
    sams = {}
    for tag, exe in (("ref", ref_binary("smalt")), ("b200", B200)):
        out = str(tmp_path / (tag + ".sam"))
        r = subprocess.run([exe, "map", "-r", "7", "-n", str(cores), "-O", "-o", out, pref, fq],
                           capture_output=True, text=True, timeout=1800)
        assert r.returncode == 0, (tag, r.stderr[-1500:])
        sams[tag] = [l for l in open(out).read().split("\n") if l and not l.startswith("@PG")]
    ref, got = sams["ref"], sams["b200"]
    assert len(ref) == len(got)
    hdr = sum(1 for l in ref if l.startswith("@"))
    assert ref[:hdr] == got[:hdr] and len(got) - hdr == n
    # with worker threads the reference draws among equally good placements in scheduling order
    # (results.c:2298): like its own test/mthread_test.py compare the records with MAPQ > 6
    ndiff = nsame = 0
    near = mapped = 0
    for k in range(n):
        a, b = ref[hdr + k], got[hdr + k]
        fa, fb = a.split("\t", 5), b.split("\t", 5)
        assert fb[0].encode() == names[k]              # every read once, in input order
        if a == b:
            nsame += 1
        elif int(fa[4]) > 6 or int(fb[4]) > 6:
            ndiff += 1
        if not int(fb[1]) & 4:
            mapped += 1
            # simread names carry the origin: r_<no>_<sequence>_<position>_<mate>_<F|R>_<edits>
            f = fb[0].split("_")
            near += abs(int(fb[3]) - int(f[3])) <= 25 and bool(int(fb[1]) & 16) == (f[5] == "R")
    assert ndiff == 0, "%d records with MAPQ > 6 differ from the reference" % ndiff
    assert nsame > 0.999 * n
    assert mapped > 0.999 * n and near > 0.995 * n     # the simulated origins are found


@pytest.mark.skipif(ref_binary("smalt") is None or not os.path.exists(B200), reason="needs oracle/_ref and smalt_b200/bin")
def test_c3_scaled_pairs(tmp_path):
    """configs[2] with the genome scaled to 4 x 5 Mb and 250 k pairs of 2 x 150 bp (misc/simread, insert 400):
    whole program against the reference's, pair by pair"""
    npairs = 250_000
    cfg = dict(bench.CONFIGS["c3"], seqs=[5_000_000] * 4)
    wl = bench.Workload(str(tmp_path), cfg, npairs)
    pref, (f1, f2) = wl.pref, wl.files
    names = [ln[1:] for ln in wl.texts[0].split(b"\n")[0::4] if ln]
    cores = bench.host_threads()
    sams = {}
    for tag, exe in (("ref", ref_binary("smalt")), ("b200", B200)):
        out = str(tmp_path / (tag + ".sam"))
        r = subprocess.run([exe, "map", "-r", "7", "-n", str(cores), "-O", "-i", "600", "-j", "200", "-o", out, pref, f1, f2],
                           capture_output=True, text=True, timeout=1800)
        assert r.returncode == 0, (tag, r.stderr[-1500:])
        sams[tag] = [l for l in open(out).read().split("\n") if l and not l.startswith("@")]
    ref, got = sams["ref"], sams["b200"]
    assert len(ref) == len(got) == 2 * npairs
    ndiff = nsame = proper = 0
    for k in range(0, 2 * npairs, 2):
        same = ref[k] == got[k] and ref[k + 1] == got[k + 1]
        nsame += same
        f = [x.split("\t", 5) for x in (ref[k], ref[k + 1], got[k], got[k + 1])]
        assert f[2][0] == f[0][0] and f[3][0] == f[1][0] and names[k // 2].startswith(f[2][0].encode())   # input order
        proper += bool(int(f[2][1]) & 2)
        # a pair is comparable when no placement in it was a random draw among equals on either side
        if not same and min(int(x[4]) for x in f) > 6:
            ndiff += 1
    assert ndiff == 0, "%d confidently placed pairs differ from the reference" % ndiff
    assert nsame > 0.97 * npairs and proper > 0.95 * npairs
