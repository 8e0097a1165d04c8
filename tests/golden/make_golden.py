#!/usr/bin/env python
"""Regenerates the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Needs /root/reference (for the embedded known-answer sequences of the
reference's own test/bam_cigar_test.py) and oracle/_ref (reference binaries
built by `make -C oracle ref`).  Run in the build container only; the outputs
are committed so that the tests do not need the reference at run time.

Outputs
  bam_cigar.json    the reference's own known-answer set (test/bam_cigar_test.py:3-51):
                    reference + reads, expected CIGAR / X-CIGAR / NM, the SAM
                    fields the reference binary actually printed, and the DP
                    boundary calls it made (trace records)
  dp_trace_c1.txt   DP boundary calls (SW / BF / BA records, see oracle/ref_trace.c)
                    of `smalt map` on a C1-like workload (1 Mb random genome,
                    100 bp simread reads, k=13 s=6): a subsample
  dp_trace_hard.txt same on a repeat-rich genome with 10% error reads, short
                    reads (< 32 bp -> banded-fast path) and 600 bp reads
                    (scores > 255 -> 16-bit retry of the reference)
  sam_c1.txt        first SAM records (no header) of the C1-like run
"""
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFBIN = os.path.join(ROOT, "oracle", "_ref")
SMALT = os.path.join(REFBIN, "smalt_trace")
SIMREAD = os.path.join(REFBIN, "simread")
REFTEST = "/root/reference/test/bam_cigar_test.py"


def run(cmd, env=None, cwd=None):
    e = dict(os.environ)
    if env:
        e.update(env)
    subprocess.run(cmd, check=True, env=e, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)


def write_fasta(path, seqs, prefix="seq"):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">%s%d\n" % (prefix, i + 1))
            for k in range(0, len(s), 60):
                f.write(s[k:k + 60] + "\n")


def random_genome(seed, n, repeat_unit=0):
    rng = np.random.default_rng(seed)
    s = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
    if repeat_unit:  # plant diverged copies of a few units -> multi-hit seeds, several candidates
        for u in range(6):
            unit = s[1000 * (u + 1):1000 * (u + 1) + repeat_unit].copy()
            for c in range(8):
                p = int(rng.integers(0, n - repeat_unit))
                cp = unit.copy()
                m = rng.random(repeat_unit) < 0.03
                cp[m] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(m.sum()))]
                s[p:p + repeat_unit] = cp
    return s.tobytes().decode()


def sam_records(path):
    out = []
    for line in open(path):
        if not line.startswith("@"):
            out.append(line.rstrip("\n"))
    return out


def known_answer(tmp):
    src = open(REFTEST).read()
    head = src.split("PROGNAM")[0]  # only the data tuples (the rest is python-2 driver code)
    ns = {}
    exec(head, ns)
    refseq, reads, pairs = ns["REFSEQ"], ns["READSEQ"], ns["READSEQ_PAIR"]
    fa = os.path.join(tmp, "ka_ref.fa")
    with open(fa, "w") as f:
        for i, s in enumerate(refseq):
            f.write(">REF_%d\n%s\n" % (i + 1, s))
    rfa = os.path.join(tmp, "ka_reads.fa")
    with open(rfa, "w") as f:
        for i, r in enumerate(reads):
            f.write(">READ_%d\n%s\n" % (i + 1, r[0]))
    idx = os.path.join(tmp, "ka")
    run([SMALT, "index", "-k", "7", "-s", "1", idx, fa])
    out = {"refseq": list(refseq), "k": 7, "s": 1, "reads": [], "sam": {}, "trace": {}}
    for fmt, key, col in (("sam", "cigar", 0), ("sam:x", "xcigar", 1)):
        sam = os.path.join(tmp, "ka_%s.sam" % key)
        trace = os.path.join(tmp, "ka_%s.trace" % key)
        run([SMALT, "map", "-f", fmt, "-o", sam, idx, rfa], env={"SMALT_TRACE": trace})
        recs = sam_records(sam)
        assert len(recs) == len(reads)
        for rec, r in zip(recs, reads):
            fld = rec.split("\t")
            assert fld[5] == r[1][col], (fld[5], r[1][col])
            nm = [t for t in fld[11:] if t.startswith("NM:")]
            assert nm and nm[0] == r[2], (nm, r[2])
        out["sam"][key] = recs
        out["trace"][key] = open(trace).read().splitlines()
    for r in reads:
        out["reads"].append({"seq": r[0], "cigar": r[1][0], "xcigar": r[1][1], "nm": r[2]})
    out["pair_reads"] = [{"seq": r[0], "cigar": r[1][0], "xcigar": r[1][1], "nm": r[2]} for r in pairs]
    with open(os.path.join(HERE, "bam_cigar.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("bam_cigar.json: %d reads reproduce the reference's expected CIGAR/NM" % len(reads))


def subsample(lines, per_kind, keep_pred=None):
    seen = {}
    out = []
    for ln in lines:
        k = ln[:2]
        special = keep_pred(ln) if keep_pred else False
        if seen.get(k, 0) < per_kind or special:
            out.append(ln)
            seen[k] = seen.get(k, 0) + 1
    return out


def dp_traces(tmp):
    # C1-like
    fa = os.path.join(tmp, "g1.fa")
    write_fasta(fa, [random_genome(1, 1_000_000)], "chr")
    idx = os.path.join(tmp, "g1")
    run([SMALT, "index", "-k", "13", "-s", "6", idx, fa])
    run([SIMREAD, idx, "100", "3000", "1.0", "y", "0", "0", "42", "sim", os.path.join(tmp, "r1")])
    trace = os.path.join(tmp, "t1.txt")
    sam = os.path.join(tmp, "o1.sam")
    run([SMALT, "map", "-o", sam, idx, os.path.join(tmp, "r1.fq")], env={"SMALT_TRACE": trace})
    lines = open(trace).read().splitlines()
    with open(os.path.join(HERE, "dp_trace_c1.txt"), "w") as f:
        f.write("\n".join(subsample(lines, 150)) + "\n")
    with open(os.path.join(HERE, "sam_c1.txt"), "w") as f:
        f.write("\n".join(sam_records(sam)[:200]) + "\n")
    # hard cases
    fa = os.path.join(tmp, "g2.fa")
    write_fasta(fa, [random_genome(2, 300_000, 700), random_genome(3, 200_000, 400)], "ctg")
    idx = os.path.join(tmp, "g2")
    run([SMALT, "index", "-k", "11", "-s", "3", idx, fa])
    hard = []
    for rl, n, err, seed in ((150, 600, "10.0", 7), (28, 300, "2.0", 8), (600, 60, "3.0", 9),
                             (60, 300, "6.0", 10)):
        pref = os.path.join(tmp, "r2_%d" % rl)
        run([SIMREAD, idx, str(rl), str(n), err, "y", "0", "0", str(seed), "sim", pref])
        trace = os.path.join(tmp, "t2_%d.txt" % rl)
        run([SMALT, "map", "-o", os.path.join(tmp, "o2.sam"), idx, pref + ".fq"],
            env={"SMALT_TRACE": trace})
        lines = open(trace).read().splitlines()
        keep = lambda ln: ln.split(" ", 2)[1] != "0" or (ln.startswith("BA") and int(ln.rsplit(" ", 1)[0].count(" ")) < 0)
        hard += subsample(lines, 60 if rl != 600 else 25, keep)
    with open(os.path.join(HERE, "dp_trace_hard.txt"), "w") as f:
        f.write("\n".join(hard) + "\n")
    print("dp traces written:", len(hard), "hard records")


if __name__ == "__main__":
    if not os.path.exists(SMALT):
        sys.exit("build oracle/_ref first: make -C oracle ref")
    with tempfile.TemporaryDirectory() as tmp:
        known_answer(tmp)
        dp_traces(tmp)
