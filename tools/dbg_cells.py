import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import smalt_b200
from oracle_lib import Oracle
from seqgen import random_seq
import test_gpu_dp as T
ctx = smalt_b200.Context(0); orc = Oracle()
rng = np.random.default_rng(104)
trip = T._rand_pairs(rng, 600, 20, 260)
pairs, args = [], []
for k, (a, b, lf) in enumerate(trip):
    if k % 7 == 0:
        b = np.concatenate([b, random_seq(rng, 9), b])
    pairs.append((a, b)); args.append(T._band_args(rng, len(a), len(b), lf))
minscore = [int(x) for x in rng.integers(1, 40, len(pairs))]
minscorlen = [int(x) for x in rng.integers(5, 30, len(pairs))]
arena, offs = T._arena(pairs)
ctx.arena_upload(arena)
tasks = T._band_tasks(pairs, offs, args, minscore, minscorlen)
bad = 0
for i in range(len(pairs)):
    res, first, diff, errs, cells = ctx.band_align(tasks[i:i+1])
    e, want, c = orc.band_align(pairs[i][0], pairs[i][1], *args[i], minscore[i], minscorlen[i])
    if c != cells:
        bad += 1
        if bad < 8: print(i, "gpu", cells, "oracle", c, "args", args[i], "q,r", len(pairs[i][0]), len(pairs[i][1]), "ms", minscore[i], minscorlen[i], "nres", len(want), "err", e)
print("bad", bad)
