"""Seeded synthetic sequences shared by the tests (codes A0 C1 G2 T3 X4 N5)."""
import numpy as np


def random_seq(rng, n, p_n=0.0, p_x=0.0):
    s = rng.integers(0, 4, n).astype(np.uint8)
    if p_n > 0:
        s[rng.random(n) < p_n] = 5
    if p_x > 0:
        s[rng.random(n) < p_x] = 4
    return s


def mutate(rng, s, p_sub=0.02, p_ins=0.005, p_del=0.005, max_indel=4):
    out = []
    i = 0
    n = len(s)
    while i < n:
        r = rng.random()
        if r < p_sub:
            out.append((int(s[i]) + int(rng.integers(1, 4))) & 3)
            i += 1
        elif r < p_sub + p_ins:
            out.extend(int(x) for x in rng.integers(0, 4, int(rng.integers(1, max_indel + 1))))
        elif r < p_sub + p_ins + p_del:
            i += int(rng.integers(1, max_indel + 1))
        else:
            out.append(int(s[i]))
            i += 1
    return np.array(out, dtype=np.uint8)


def read_window_pair(rng, qlen, flank=(5, 40), with_flank=False, **mut):
    """A read and a reference window containing a mutated copy of it."""
    core = random_seq(rng, qlen)
    read = mutate(rng, core, **mut)
    if len(read) < 8:
        read = core.copy()
    lf, rf = (int(x) for x in rng.integers(flank[0], flank[1] + 1, 2))
    ref = np.concatenate([random_seq(rng, lf), core, random_seq(rng, rf)])
    if with_flank:
        return np.ascontiguousarray(read), np.ascontiguousarray(ref), lf
    return np.ascontiguousarray(read), np.ascontiguousarray(ref)


def revcomp(s):
    r = s[::-1].copy()
    m = r < 4
    r[m] = 3 - r[m]
    return np.ascontiguousarray(r)
