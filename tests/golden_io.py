"""Parsers for the golden fixtures under tests/golden/ (formats: oracle/ref_trace.c)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_CODE = {c: i for i, c in enumerate("ACGTXN")}


def enc(s):
    return np.array([_CODE.get(c, 5) for c in s.upper()], dtype=np.uint8)


def load_trace(name):
    """-> list of dict records (kind SW / BF / BA)."""
    path = name if os.path.isabs(name) else os.path.join(GOLDEN, name)
    recs = []
    for line in open(path):
        recs.append(parse_record(line))
    return recs


def parse_record(line):
    f = line.split()
    k = f[0]
    if k == "SW":
        return dict(kind=k, err=int(f[1]), score=int(f[2]), read=enc(f[5]), ref=enc(f[6]))
    if k == "BF":
        return dict(kind=k, err=int(f[1]), score=int(f[2]), args=tuple(int(x) for x in f[3:9]),
                    read=enc(f[11]), ref=enc(f[12]))
    if k == "BA":
        nres = int(f[14])
        res = []
        for i in range(nres):
            o = 15 + 6 * i
            res.append((tuple(int(x) for x in f[o:o + 5]), bytes.fromhex(f[o + 5])))
        return dict(kind=k, err=int(f[1]), args=tuple(int(x) for x in f[2:8]),
                    minscore=int(f[8]), minscorlen=int(f[9]), read=enc(f[12]), ref=enc(f[13]),
                    results=res)
    raise ValueError(line[:40])


def load_bam_cigar():
    return json.load(open(os.path.join(GOLDEN, "bam_cigar.json")))


def diffstr_to_cigar(dstr, clip_start=0, clip_end=0, xmismatch=False):
    """CIGAR of a forward DiffStr (diffstr.h:28-105): each byte = `count` matches then one
    column of its type (M match, S mismatch - or nothing when it closes the string,
    I insertion in the read, D deletion from the read)."""
    cols = []
    body = dstr[:dstr.index(0)] if 0 in dstr else dstr
    for n, b in enumerate(body):
        cnt, typ = b & 0x3F, b >> 6
        cols.append(("=", cnt))
        if typ == 0:
            cols.append(("=", 1))
        elif typ == 3:
            if n != len(body) - 1:
                cols.append(("X", 1))
        elif typ == 2:
            cols.append(("I", 1))
        else:
            cols.append(("D", 1))
    out = []
    for op, n in cols:
        if n == 0:
            continue
        if not xmismatch and op in "=X":
            op = "M"
        elif xmismatch and op == "=":
            op = "M"
        if out and out[-1][0] == op:
            out[-1][1] += n
        else:
            out.append([op, n])
    s = "".join("%d%s" % (n, op) for op, n in out)
    if clip_start:
        s = "%dS" % clip_start + s
    if clip_end:
        s += "%dS" % clip_end
    return s
